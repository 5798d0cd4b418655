// kernels_ipa.cuh — Invariant Point Attention (ipa_pytorch.py:170-329).
//
// Data flow of one IPA call (B samples, N residues, H=8 heads, C=256, Pq=8, Pv=12, c_z=128):
//   proj [M, 6816]   <- s @ Wcat^T + bcat                (one split-TF32 GEMM; Wcat = the four projection weights with rows permuted
//                                                          so that every head's q | q_pts, k | k_pts, v | v_pts are contiguous)
//   ipa_prep_kernel  : points -> global frame in place (R_i p + t_i); q_pts additionally scaled by gamma_h / s_qk; kn[j,h] = -gamma_h/2 |k_pts|^2
//   S[b,h,i,j]       <- s_qk * Q'_h . K'_h   (batched split-TF32 GEMM, K = 256 + 24):   s_qk q.k + gamma_h q_pts.k_pts
//                       ( -gamma/2 |q_p - k_p|^2 = gamma q_p.k_p - gamma/2 |k_p|^2 - gamma/2 |q_p|^2 ; the last term is constant in j
//                         and cancels in the softmax )
//   ipa_core_kernel  : logits = S + sqrt(1/3) (W_b z_ij + b_b) + kn[j,h] + 1e5 (m_i m_j - 1);  a = softmax_j;  a -> S (fp32);
//                      o_pair = down_z(sum_j a_ij z_ij)   (down_z is linear and sum_j a = 1, so it commutes)
//                      Both z contractions run on tcgen05 (fp16 z tile images straight from HBM via bulk copies, fp32 accumulate in TMEM):
//                        GEMM-b  D1[j, h]  = z_tile[j, c] . Wb[h, c]^T         (A K-major, N = 16: rows 0-7 fp16 hi, 8-15 fp16 lo of W_b)
//                        GEMM-o  D2[c, h] += z_tile[j, c]^T . P[h, j]^T        (A = the same smem tile read MN-major, B rows = hi | lo of a)
//                      z[b,i,:,:] is read from HBM once per (b,i): the tiles of a row stay resident in the smem ring for both GEMMs when
//                      N <= 128*ring slots, otherwise the second pass re-reads them from L2.
//   O'[b,i,h,0:292]  <- A_h [V_h | v_pts_h]   (batched split-TF32 GEMM with MN-major B, written straight into the concat buffer)
//   ipa_opt_kernel   : o_pt -> local frame R_i^T (p - t_i) in place, norms
//   out = linear_out(cat')  (GEMM; linear_out.weight columns permuted once to the cat' order)
//
// cat' column order: [h: o(256) | o_pt.x(12) | o_pt.y(12) | o_pt.z(12)] x 8 | norms (8 x 12) | o_pair (8 x 32)
#pragma once
#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace fdpt {

constexpr int QK_W = C_HID + PQ * 3;          // 280
constexpr int V_W = C_HID + PV * 3;           // 292
constexpr int PROJ_Q = 0, PROJ_K = NH * QK_W, PROJ_V = 2 * NH * QK_W, PROJ_W = 2 * NH * QK_W + NH * V_W;  // 6816
constexpr int CATP_NRM = NH * V_W;            // 2336
constexpr int CATP_PAIR = CATP_NRM + NH * PV;  // 2432
static_assert(PROJ_W == NH * (C_HID * 3 + (2 * PQ + PV) * 3), "projection width");
static_assert(CATP_PAIR + NH * (C_Z / 4) == CAT, "concat width");

// In-place frame application on the point slots of proj (planar x|y|z per head), ipa_pytorch.py:214-239, rigid_utils.py:82-106.
__global__ void __launch_bounds__(256) ipa_prep_kernel(int M, float* __restrict__ proj, const float* __restrict__ quats,
                                                       const float* __restrict__ trans, const float* __restrict__ head_w,
                                                       float* __restrict__ kn) {
  const int m = blockIdx.x, tid = threadIdx.x;
  __shared__ float R[9], t[3];
  if (tid == 0) {
    float q[4] = {quats[m * 4], quats[m * 4 + 1], quats[m * 4 + 2], quats[m * 4 + 3]};
    quat_to_rot(q, R);
    t[0] = trans[m * 3];
    t[1] = trans[m * 3 + 1];
    t[2] = trans[m * 3 + 2];
  }
  __syncthreads();
  float* row = proj + (long long)m * PROJ_W;
  float* slot;
  int h, p, np;
  if (tid < NH * PQ) {
    h = tid / PQ; p = tid % PQ; np = PQ;
    slot = row + PROJ_Q + h * QK_W + C_HID;
  } else if (tid < 2 * NH * PQ) {
    h = (tid - NH * PQ) / PQ; p = tid % PQ; np = PQ;
    slot = row + PROJ_K + h * QK_W + C_HID;
  } else if (tid < 2 * NH * PQ + NH * PV) {
    const int v = tid - 2 * NH * PQ;
    h = v / PV; p = v % PV; np = PV;
    slot = row + PROJ_V + h * V_W + C_HID;
  } else {
    return;
  }
  const float x = slot[p], y = slot[np + p], z = slot[2 * np + p];
  float gx = R[0] * x + R[1] * y + R[2] * z + t[0];
  float gy = R[3] * x + R[4] * y + R[5] * z + t[1];
  float gz = R[6] * x + R[7] * y + R[8] * z + t[2];
  const float gamma = log1pf(expf(head_w[h])) * sqrtf(1.0f / (3.f * (PQ * 9.0f / 2.f)));  // softplus(w_h) * sqrt(1/108)
  if (tid < NH * PQ) {
    const float sc = gamma / sqrtf(1.0f / (3.f * C_HID));  // the S GEMM multiplies the whole dot product by s_qk
    gx *= sc; gy *= sc; gz *= sc;
  } else if (tid < 2 * NH * PQ) {
    float d2 = gx * gx + gy * gy + gz * gz;
    d2 += __shfl_xor_sync(0xffffffffu, d2, 1);
    d2 += __shfl_xor_sync(0xffffffffu, d2, 2);
    d2 += __shfl_xor_sync(0xffffffffu, d2, 4);
    if (p == 0) kn[(long long)m * NH + h] = -0.5f * gamma * d2;
  }
  slot[p] = gx;
  slot[np + p] = gy;
  slot[2 * np + p] = gz;
}

// o_pt: global-frame sums -> local frame R_i^T (p - t_i) in place + norms (ipa_pytorch.py:302-316)
__global__ void __launch_bounds__(96) ipa_opt_kernel(int M, float* __restrict__ cat, const float* __restrict__ quats,
                                                     const float* __restrict__ trans) {
  const int m = blockIdx.x, tid = threadIdx.x;
  __shared__ float R[9], t[3];
  if (tid == 0) {
    float q[4] = {quats[m * 4], quats[m * 4 + 1], quats[m * 4 + 2], quats[m * 4 + 3]};
    quat_to_rot(q, R);
    t[0] = trans[m * 3];
    t[1] = trans[m * 3 + 1];
    t[2] = trans[m * 3 + 2];
  }
  __syncthreads();
  const int h = tid / PV, p = tid % PV;
  float* slot = cat + (long long)m * CAT + h * V_W + C_HID;
  const float x = slot[p] - t[0], y = slot[PV + p] - t[1], z = slot[2 * PV + p] - t[2];
  const float lx = R[0] * x + R[3] * y + R[6] * z;
  const float ly = R[1] * x + R[4] * y + R[7] * z;
  const float lz = R[2] * x + R[5] * y + R[8] * z;
  slot[p] = lx;
  slot[PV + p] = ly;
  slot[2 * PV + p] = lz;
  cat[(long long)m * CAT + CATP_NRM + h * PV + p] = sqrtf(lx * lx + ly * ly + lz * lz + 1e-8f);
}

struct IpaCoreArgs {
  int B, N, JB, ldS;
  float* S;              // [B,H,N,ldS] in: s_qk q.k + gamma q_pts.k_pts ; out: attention probabilities
  const __half* z;       // fp16 tile images [B][N][JB][2 k-blocks][128 rows][128 B, 128B-swizzled] (et_fused.cuh)
  const float* kn;       // [B*N, H]   -gamma_h/2 |k_pts|^2
  const float* mask;     // [B*N]
  const __half* Wb_img;  // 4 KB operand image [2 kb][16 rows][128 B]: rows 0-7 fp16 hi, rows 8-15 fp16 lo of linear_b.weight
  const float* bb;       // [H]
  const float* Wd;       // [32,128] down_z.weight
  const float* bd;       // [32]
  float* cat;            // [B*N, CAT] (cat' order)
  int rz, resident;      // ring slots; 1 = the JB tiles of a row stay in the ring for both GEMMs
  int rows;              // B*N
  int mn_swap;           // bring-up knob: swap LBO / SBO of the MN-major A descriptor
};

constexpr int IPA_TILE_BYTES = 32768;
constexpr int IPA_PIMG_TILE = 4096;   // [2 kb][16 rows][128 B]
constexpr int IPA_OZ_LD = 132;
constexpr int IPA_WDT_LD = 33;
constexpr int IPA_MAX_RZ = 6, IPA_MAX_JB = 8;

struct IpaSmemPlan {
  int rz, resident, ctas_per_sm;
  size_t bytes;
};
// Shared-memory plan.  The kernel is latency-bound per row (load -> GEMM-b -> logits -> softmax -> GEMM-o -> down_z), so two CTAs per
// SM are preferred whenever they fit (2-slot ring, second pass re-reads the row from L2); otherwise one CTA with the deepest ring.
inline IpaSmemPlan ipa_core_plan(int N, int max_smem, int max_smem_per_sm) {
  const int JB = (N + 127) / 128;
  const size_t lbytes = (size_t)NH * (JB * 128 > IPA_OZ_LD ? JB * 128 : IPA_OZ_LD) * 4;  // logits [H][ldL], re-used as ozs [H][132]
  const size_t other = (size_t)JB * IPA_PIMG_TILE + 4096 + lbytes + C_Z * IPA_WDT_LD * 4 + 512 + 1024;
  IpaSmemPlan p;
  const size_t two = other + 2 * (size_t)IPA_TILE_BYTES;
  if (2 * (two + 1024) <= (size_t)max_smem_per_sm) {
    p.rz = 2;
    p.ctas_per_sm = 2;
  } else {
    int rz = (int)(((size_t)max_smem - other) / IPA_TILE_BYTES);
    p.rz = rz > IPA_MAX_RZ ? IPA_MAX_RZ : rz;
    p.ctas_per_sm = 1;
  }
  p.resident = JB <= p.rz ? 1 : 0;
  p.bytes = other + (size_t)p.rz * IPA_TILE_BYTES;
  return p;
}

__global__ void __launch_bounds__(192, 2) ipa_core_kernel(IpaCoreArgs a) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int N = a.N, JB = a.JB, ldL = JB * 128;
  uint8_t* ring = smem;                                      // rz x 32 KB z tiles
  uint8_t* Pimg = ring + (size_t)a.rz * IPA_TILE_BYTES;      // JB x 4 KB probability images (B operand of GEMM-o)
  uint8_t* Wbs = Pimg + (size_t)JB * IPA_PIMG_TILE;          // 4 KB
  float* L = reinterpret_cast<float*>(Wbs + 4096);           // [H][ldL] logits / exp
  float* ozs = L;                                            // [H][IPA_OZ_LD], aliases L (dead once the probabilities are written)
  float* WdT = L + NH * (ldL > IPA_OZ_LD ? ldL : IPA_OZ_LD);  // [128][33] down_z.weight^T
  uint64_t* bars = reinterpret_cast<uint64_t*>(WdT + C_Z * IPA_WDT_LD);
  uint64_t* zfull = bars;                    // [IPA_MAX_RZ]
  uint64_t* zfree = zfull + IPA_MAX_RZ;      // [IPA_MAX_RZ]
  uint64_t* d1_full = zfree + IPA_MAX_RZ;    // [IPA_MAX_JB]
  uint64_t* p_full = d1_full + IPA_MAX_JB;   // [1]
  uint64_t* d2_full = p_full + 1;            // [1]
  uint64_t* wb_full = d2_full + 1;           // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wb_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;

  if (tid == 0) {
    for (int s = 0; s < IPA_MAX_RZ; ++s) {
      mbar_init(&zfull[s], 1);
      mbar_init(&zfree[s], 1);
    }
    for (int t = 0; t < IPA_MAX_JB; ++t) mbar_init(&d1_full[t], 1);
    mbar_init(p_full, 128);
    mbar_init(d2_full, 1);
    mbar_init(wb_full, 1);
    fence_barrier_init();
  }
  // rows 8..15 of the probability images hold the fp16 lo parts; every byte is rewritten per row, nothing to clear
  for (int k = tid; k < C_Z * (C_Z / 4); k += blockDim.x) {
    const int d = k / C_Z, c = k % C_Z;
    WdT[c * IPA_WDT_LD + d] = a.Wd[k];
  }
  if (warp == 4) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t D1 = tmem_base, D2 = tmem_base + 128;

  if (warp == 5) {
    // ============================ loader ============================
    if (lane == 0) {
      mbar_arrive_expect_tx(wb_full, 4096);
      bulk_g2s(Wbs, a.Wb_img, 4096, wb_full);
      uint32_t cnt = 0;
      const int passes = a.resident ? 1 : 2;
      for (int row = blockIdx.x; row < a.rows; row += gridDim.x) {
        const uint8_t* zrow = reinterpret_cast<const uint8_t*>(a.z) + (size_t)row * JB * IPA_TILE_BYTES;
        for (int pass = 0; pass < passes; ++pass)
          for (int t = 0; t < JB; ++t) {
            const uint32_t s = cnt % a.rz;
            mbar_wait(&zfree[s], ((cnt / a.rz) & 1) ^ 1);
            mbar_arrive_expect_tx(&zfull[s], IPA_TILE_BYTES);
            bulk_g2s(ring + (size_t)s * IPA_TILE_BYTES, zrow + (size_t)t * IPA_TILE_BYTES, IPA_TILE_BYTES, &zfull[s]);
            ++cnt;
          }
      }
    }
  } else if (warp == 4) {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      mbar_wait(wb_full, 0);
      tc_fence_after();
      const uint32_t idesc_b = make_idesc_f16(128, 16);
      const uint32_t idesc_o = make_idesc_f16(128, 16) | (1u << 15);  // A operand MN-major
      const uint32_t wb = smem_u32(Wbs), pimg = smem_u32(Pimg), ring_u = smem_u32(ring);
      const uint32_t lbo = a.mn_swap ? 1024u : 16384u, sbo = a.mn_swap ? 16384u : 1024u;
      uint32_t cnt = 0, it = 0;
      for (int row = blockIdx.x; row < a.rows; row += gridDim.x, ++it) {
        const uint32_t base_cnt = cnt;
        // GEMM-b: D1[t] = z_tile . Wb^T
        for (int t = 0; t < JB; ++t) {
          const uint32_t s = cnt % a.rz;
          mbar_wait(&zfull[s], (cnt / a.rz) & 1);
          tc_fence_after();
          const uint32_t zt = ring_u + s * IPA_TILE_BYTES;
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(D1 + t * 16, make_sw128_desc(zt + kb * 16384 + k * 32), make_sw128_desc(wb + kb * 2048 + k * 32), idesc_b,
                       (kb | k) ? 1u : 0u);
          umma_commit(&d1_full[t]);
          if (!a.resident) umma_commit(&zfree[s]);
          ++cnt;
        }
        // GEMM-o: D2 += z_tile^T . P^T  (after the softmax of this row has written the probability images)
        mbar_wait(p_full, it & 1);
        tc_fence_after();
        for (int t = 0; t < JB; ++t) {
          uint32_t s;
          if (a.resident) {
            s = (base_cnt + t) % a.rz;
          } else {
            s = cnt % a.rz;
            mbar_wait(&zfull[s], (cnt / a.rz) & 1);
            tc_fence_after();
            ++cnt;
          }
          const uint32_t zt = ring_u + s * IPA_TILE_BYTES;
#pragma unroll
          for (int k = 0; k < 8; ++k)  // 16 j-rows per step
            umma_f16(D2, make_sw128_desc_ls(zt + k * 2048, lbo, sbo),
                     make_sw128_desc(pimg + t * IPA_PIMG_TILE + (k >> 2) * 2048 + (k & 3) * 32), idesc_o, (t | k) ? 1u : 0u);
          umma_commit(&zfree[s]);
        }
        umma_commit(d2_full);
      }
    }
  } else {
    // ============================ logits / softmax / down_z (128 threads) ============================
    const int r = tid;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const float s_b = sqrtf(1.0f / 3.f);
    float bbv[NH];
#pragma unroll
    for (int h = 0; h < NH; ++h) bbv[h] = a.bb[h];
    const int d0 = (2 * r) & 31, hh_out = r >> 4;
    const float bd0 = a.bd[d0], bd1 = a.bd[d0 + 1];
    uint32_t it = 0;
    for (int row = blockIdx.x; row < a.rows; row += gridDim.x, ++it) {
      const int b = row / N, i = row - b * N;
      const float mi = a.mask[row];
      float* Srow0 = a.S + (((long long)b * NH) * N + i) * a.ldS;  // head h: + h * N * ldS
      const long long hs = (long long)N * a.ldS;
      for (int t = 0; t < JB; ++t) {
        const int j = t * 128 + r;
        const bool valid = j < N;
        float sv[NH];
        float4 k0 = make_float4(0.f, 0.f, 0.f, 0.f), k1 = k0;
        float mj = 0.f;
        if (valid) {
#pragma unroll
          for (int h = 0; h < NH; ++h) sv[h] = Srow0[h * hs + j];
          const float4* kp = reinterpret_cast<const float4*>(a.kn + ((long long)b * N + j) * NH);
          k0 = __ldg(kp);
          k1 = __ldg(kp + 1);
          mj = a.mask[(long long)b * N + j];
        }
        mbar_wait(&d1_full[t], it & 1);
        tc_fence_after();
        float bv[16];
        tmem_ld16(D1 + lane_base + t * 16, bv);
        tmem_ld_wait();
        const float knv[NH] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
        const float mterm = 1e5f * (mi * mj - 1.f);
#pragma unroll
        for (int h = 0; h < NH; ++h)
          L[h * ldL + j] = valid ? sv[h] + s_b * (bv[h] + bv[h + 8] + bbv[h]) + knv[h] + mterm : -INFINITY;
      }
      tc_fence_before();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // ---- softmax over j: warp w owns heads 2w, 2w+1
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int h = 2 * warp + hh;
        float* Lh = L + h * ldL;
        float mx = -INFINITY;
        for (int j = lane; j < ldL; j += 32) mx = fmaxf(mx, Lh[j]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int j = lane; j < ldL; j += 32) {
          const float e = expf(Lh[j] - mx);
          Lh[j] = e;
          sum += e;
        }
        sum = warp_sum(sum);
        const float inv = 1.f / sum;
        float* Sh = Srow0 + h * hs;
        for (int j = lane; j < ldL; j += 32) {
          const float p = Lh[j] * inv;
          if (j < N) Sh[j] = p;
          const __half ph = __float2half_rn(p);
          const __half pl = __float2half_rn(p - __half2float(ph));
          const int t = j >> 7, jj = j & 127;
          uint8_t* dst = Pimg + t * IPA_PIMG_TILE + (jj >> 6) * 2048 + h * 128 + ((((jj & 63) >> 3) ^ (h & 7)) << 4) + (jj & 7) * 2;
          *reinterpret_cast<__half*>(dst) = ph;
          *reinterpret_cast<__half*>(dst + 1024) = pl;  // row h + 8 (same swizzle phase)
        }
      }
      fence_proxy_async();
      mbar_arrive(p_full);
      // ---- o_pair: D2[c, h] (hi + lo columns) -> ozs[h][c] -> down_z
      mbar_wait(d2_full, it & 1);
      tc_fence_after();
      {
        float ov[16];
        tmem_ld16(D2 + lane_base, ov);
        tmem_ld_wait();
#pragma unroll
        for (int h = 0; h < NH; ++h) ozs[h * IPA_OZ_LD + r] = ov[h] + ov[h + 8];
      }
      tc_fence_before();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      {
        float acc0 = bd0, acc1 = bd1;
        const float* o = ozs + hh_out * IPA_OZ_LD;
        const float* w = WdT + d0;
#pragma unroll 8
        for (int c = 0; c < C_Z; ++c) {
          const float x = o[c];
          acc0 = fmaf(w[c * IPA_WDT_LD], x, acc0);
          acc1 = fmaf(w[c * IPA_WDT_LD + 1], x, acc1);
        }
        float* dst = a.cat + (long long)row * CAT + CATP_PAIR + hh_out * (C_Z / 4) + d0;
        *reinterpret_cast<float2*>(dst) = make_float2(acc0, acc1);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // ozs aliases L: the next row's logits may not land before every thread is done
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 256);
}

}  // namespace fdpt
