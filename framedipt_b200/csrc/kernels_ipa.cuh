// kernels_ipa.cuh — Invariant Point Attention (ipa_pytorch.py:170-329).
//
// Data flow of one IPA call (B samples, N residues, H=8 heads, C=256, Pq=8, Pv=12, c_z=128):
//   proj [M, 6816]   <- s @ Wcat^T + bcat                (one split-TF32 GEMM; Wcat = the four projection weights with rows permuted
//                                                          so that every head's q | q_pts, k | k_pts, v | v_pts are contiguous)
//   ipa_prep_kernel  : points -> global frame in place (R_i p + t_i); q_pts additionally scaled by gamma_h / s_qk;
//                      kbias[b,h,j] = -gamma_h/2 |k_pts_j|^2 + 1e5 (m_j - 1)
//   S[b,h,i,j]       <- s_qk * Q'_h . K'_h + kbias[b,h,j]   (batched split-precision GEMM, K = 256 + 24, per-batch column bias)
//                       ( -gamma/2 |q_p - k_p|^2 = gamma q_p.k_p - gamma/2 |k_p|^2 - gamma/2 |q_p|^2 ; the last term is constant in j
//                         and cancels in the softmax.  Mask term of the reference: 1e5 (m_i m_j - 1) = 1e5 (m_j - 1) on rows with
//                         m_i = 1; rows with m_i = 0 are zeroed after the IPA (ipa_pytorch.py:531), SURVEY V2 )
//   ipa_core_kernel  : logits = S + sqrt(1/3) (W_b z_ij + b_b);  a = softmax_j;  a -> S (fp32);
//                      o_pair = down_z(sum_j a_ij z_ij)   (down_z is linear and sum_j a = 1, so it commutes)
//                      Both z contractions run on tcgen05 (fp16 z tile images straight from HBM via bulk copies, fp32 accumulate in TMEM):
//                        GEMM-b  D1[j, h]  = z_tile[j, c] . Wb[h, c]^T         (A K-major, N = 16: rows 0-7 fp16 hi, 8-15 fp16 lo of W_b)
//                        GEMM-o  D2[c, h] += z_tile[j, c]^T . P[h, j]^T        (A = the same smem tile read MN-major, B rows = hi | lo of a)
//                      z[b,i,:,:] is read from HBM once per (b,i); the second pass (GEMM-o) re-reads the row from L2.
//   O'[b,i,h,0:292]  <- A_h [V_h | v_pts_h]   (batched split-TF32 GEMM with MN-major B, written straight into the concat buffer)
//   ipa_opt_kernel   : o_pt -> local frame R_i^T (p - t_i) in place, norms
//   out = linear_out(cat')  (GEMM; linear_out.weight columns permuted once to the cat' order)
//
// cat' column order: [h: o(256) | o_pt.x(12) | o_pt.y(12) | o_pt.z(12)] x 8 | norms (8 x 12) | o_pair (8 x 32)
#pragma once
#include <cuda_fp16.h>

#include "lin_tc.cuh"
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
                                                       const float* __restrict__ mask, int N, float* __restrict__ kbias) {
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
    if (p == 0) {  // kbias[b][h][j]
      const int b = m / N, j = m - b * N;
      kbias[((long long)b * NH + h) * N + j] = -0.5f * gamma * d2 + 1e5f * (mask[m] - 1.f);
    }
  }
  slot[p] = gx;
  slot[np + p] = gy;
  slot[2 * np + p] = gz;
}

// ---- operand-image form of the prep step (gemm_img.cuh) ---------------------------------------------------------------------
// Reads the fused projection row of residue m and writes, per head, the three operands of the attention GEMMs as fp16 hi | lo operand
// images [b, h][128-row tile][5 k-blocks of 64 columns][hi 16 KB | lo 16 KB][128 rows][128 B swizzled]:
//   Q' = [s_qk q (256) | gamma (R q_pts + t) (24) | 0 (40)]      K' = [k (256) | R k_pts + t (24) | 0]      V' = [v (256) | R v_pts + t (36) | 0]
// (points planar x | y | z like the projection) and kbias[b, h, j] = -gamma/2 |k_pts_j|^2 + 1e5 (m_j - 1).  Padding rows / columns are
// never written: the image buffers are zeroed once when the workspace is reserved.
struct IpaImgArgs {
  int M, N, JB;
  const float* proj; const float* quats; const float* trans; const float* head_w; const float* mask;
  float* kbias;
  uint8_t* Qimg; uint8_t* Kimg; uint8_t* Vimg;
};
constexpr int IPA_IMG_KB = 5;  // 320 columns per head and operand

FDPT_DEVINL void ipa_img_store8(uint8_t* img, int JB, int bh, int i, int c, const float x[8]) {
  uint4 hi, lo;
  tc::split8(make_float4(x[0], x[1], x[2], x[3]), make_float4(x[4], x[5], x[6], x[7]), hi, lo);
  uint8_t* d = img + (((size_t)bh * JB + (i >> 7)) * IPA_IMG_KB + (c >> 3)) * tc::LT_STAGE_BYTES + tc::sw128_chunk_off(i & 127, c & 7);
  *reinterpret_cast<uint4*>(d) = hi;
  *reinterpret_cast<uint4*>(d + 16384) = lo;
}

__global__ void __launch_bounds__(256) ipa_prep_img_kernel(IpaImgArgs a) {
  const int m = blockIdx.x, tid = threadIdx.x;
  __shared__ float R[9], t[3];
  if (tid == 0) {
    float q[4] = {a.quats[m * 4], a.quats[m * 4 + 1], a.quats[m * 4 + 2], a.quats[m * 4 + 3]};
    quat_to_rot(q, R);
    t[0] = a.trans[m * 3];
    t[1] = a.trans[m * 3 + 1];
    t[2] = a.trans[m * 3 + 2];
  }
  __syncthreads();
  const int b = m / a.N, i = m - b * a.N;
  const float* row = a.proj + (long long)m * PROJ_W;
  const float s_qk = sqrtf(1.0f / (3.f * C_HID));
  // scalar channels: 3 operands x 8 heads x 32 chunks of 8
  for (int idx = tid; idx < 3 * NH * 32; idx += 256) {
    const int kind = idx / (NH * 32), h = (idx >> 5) & (NH - 1), c = idx & 31;
    const float* src = row + (kind == 0 ? PROJ_Q + h * QK_W : kind == 1 ? PROJ_K + h * QK_W : PROJ_V + h * V_W) + 8 * c;
    const float4 x0 = *reinterpret_cast<const float4*>(src), x1 = *reinterpret_cast<const float4*>(src + 4);
    float x[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
    if (kind == 0) {
#pragma unroll
      for (int e = 0; e < 8; ++e) x[e] *= s_qk;
    }
    ipa_img_store8(kind == 0 ? a.Qimg : kind == 1 ? a.Kimg : a.Vimg, a.JB, b * NH + h, i, c, x);
  }
  // points: one thread per (operand, head)
  if (tid < 3 * NH) {
    const int kind = tid / NH, h = tid % NH;
    const float gamma = log1pf(expf(a.head_w[h])) * sqrtf(1.0f / (3.f * (PQ * 9.0f / 2.f)));  // softplus(w_h) * sqrt(1/108)
    if (kind < 2) {
      const float* slot = row + (kind == 0 ? PROJ_Q : PROJ_K) + h * QK_W + C_HID;
      float gx[PQ], gy[PQ], gz[PQ];
      float d2 = 0.f;
#pragma unroll
      for (int p = 0; p < PQ; ++p) {
        const float x = slot[p], y = slot[PQ + p], z = slot[2 * PQ + p];
        gx[p] = R[0] * x + R[1] * y + R[2] * z + t[0];
        gy[p] = R[3] * x + R[4] * y + R[5] * z + t[1];
        gz[p] = R[6] * x + R[7] * y + R[8] * z + t[2];
        d2 += gx[p] * gx[p] + gy[p] * gy[p] + gz[p] * gz[p];
      }
      if (kind == 0) {
#pragma unroll
        for (int p = 0; p < PQ; ++p) {
          gx[p] *= gamma; gy[p] *= gamma; gz[p] *= gamma;
        }
      } else {
        a.kbias[((long long)b * NH + h) * a.N + i] = -0.5f * gamma * d2 + 1e5f * (a.mask[m] - 1.f);
      }
      uint8_t* img = kind == 0 ? a.Qimg : a.Kimg;
      ipa_img_store8(img, a.JB, b * NH + h, i, 32, gx);
      ipa_img_store8(img, a.JB, b * NH + h, i, 33, gy);
      ipa_img_store8(img, a.JB, b * NH + h, i, 34, gz);
    } else {
      const float* slot = row + PROJ_V + h * V_W + C_HID;
      float g[40];
#pragma unroll
      for (int p = 0; p < PV; ++p) {
        const float x = slot[p], y = slot[PV + p], z = slot[2 * PV + p];
        g[p] = R[0] * x + R[1] * y + R[2] * z + t[0];
        g[PV + p] = R[3] * x + R[4] * y + R[5] * z + t[1];
        g[2 * PV + p] = R[6] * x + R[7] * y + R[8] * z + t[2];
      }
#pragma unroll
      for (int p = 3 * PV; p < 40; ++p) g[p] = 0.f;
#pragma unroll
      for (int c = 0; c < 5; ++c) ipa_img_store8(a.Vimg, a.JB, b * NH + h, i, 32 + c, g + 8 * c);
    }
  }
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

// ipa_opt + operand image of the concat row in one pass (gemm_img.cuh consumes it as the A operand of linear_out, K = 2688 = 42 k-blocks):
// the row of residue m is staged in shared memory, the point outputs are moved to the local frame and their norms written exactly as
// ipa_opt_kernel does, then the whole row is split into fp16 hi | lo 16-byte chunks of the image
// [128-row tile of the flattened residue index][42 k-blocks][hi 16 KB | lo 16 KB][128 rows][128 B swizzled].
__global__ void __launch_bounds__(256) ipa_opt_img_kernel(int M, const float* __restrict__ cat, const float* __restrict__ quats,
                                                          const float* __restrict__ trans, uint8_t* __restrict__ img) {
  const int m = blockIdx.x, tid = threadIdx.x;
  __shared__ __align__(16) float row[CAT];
  __shared__ float R[9], t[3];
  const float4* src = reinterpret_cast<const float4*>(cat + (long long)m * CAT);
  for (int k = tid; k < CAT / 4; k += 256) reinterpret_cast<float4*>(row)[k] = src[k];
  if (tid == 0) {
    float q[4] = {quats[m * 4], quats[m * 4 + 1], quats[m * 4 + 2], quats[m * 4 + 3]};
    quat_to_rot(q, R);
    t[0] = trans[m * 3];
    t[1] = trans[m * 3 + 1];
    t[2] = trans[m * 3 + 2];
  }
  __syncthreads();
  if (tid < NH * PV) {
    const int h = tid / PV, p = tid % PV;
    float* slot = row + h * V_W + C_HID;
    const float x = slot[p] - t[0], y = slot[PV + p] - t[1], z = slot[2 * PV + p] - t[2];
    const float lx = R[0] * x + R[3] * y + R[6] * z;
    const float ly = R[1] * x + R[4] * y + R[7] * z;
    const float lz = R[2] * x + R[5] * y + R[8] * z;
    slot[p] = lx;
    slot[PV + p] = ly;
    slot[2 * PV + p] = lz;
    row[CATP_NRM + h * PV + p] = sqrtf(lx * lx + ly * ly + lz * lz + 1e-8f);
  }
  __syncthreads();
  uint8_t* tile = img + (size_t)(m >> 7) * (CAT / 64) * tc::LT_STAGE_BYTES;
  const int r = m & 127;
  for (int c = tid; c < CAT / 8; c += 256) {
    const float4 x0 = reinterpret_cast<const float4*>(row)[2 * c], x1 = reinterpret_cast<const float4*>(row)[2 * c + 1];
    uint4 hi, lo;
    tc::split8(x0, x1, hi, lo);
    uint8_t* d = tile + (size_t)(c >> 3) * tc::LT_STAGE_BYTES + tc::sw128_chunk_off(r, c & 7);
    *reinterpret_cast<uint4*>(d) = hi;
    *reinterpret_cast<uint4*>(d + 16384) = lo;
  }
}

// 2^x for x <= 0 with the SFU instruction (2 ulp; results below the normal range flush to zero, which is what a softmax wants)
FDPT_DEVINL float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct IpaCoreArgs {
  int B, N, JB, ldS;
  float* S;              // [B,H,N,ldS] in: s_qk q.k + gamma q_pts.k_pts + kbias ; out: attention probabilities
  const __half* z;       // fp16 tile images [B][N][JB][2 k-blocks][128 rows][128 B, 128B-swizzled] (et_fused.cuh)
  const __half* Wb_img;  // 4 KB operand image [2 kb][16 rows][128 B]: rows 0-7 fp16 hi, rows 8-15 fp16 lo of linear_b.weight
  const float* bb;       // [H]
  const float* Wd;       // [32,128] down_z.weight
  const float* bd;       // [32]
  float* cat;            // [B*N, CAT] (cat' order)
  int rz, tmem_cols;     // z ring slots; TMEM columns to allocate (two D1 buffers + D2)
  int single_pass;       // 1: the ring holds two whole rows (rz = 2 JB): GEMM-o re-uses the tiles GEMM-b left in shared memory, every z
                         //    tile is fetched once; 0: the tiles of a row are fetched a second time for GEMM-o (L2 hits)
  int rows;              // B*N
  uint8_t* Pg;           // optional: probabilities as fp16 hi | lo operand images [b, h][i-tile][2 JB k-blocks][32 KB] for the A.V GEMM
                         // (gemm_img.cuh); S then keeps the logits (the fp32 probabilities are not written)
  int mn_swap;           // bring-up knob: swap LBO / SBO of the MN-major A descriptor
  long long* dbg;        // optional clock64 timeline of CTA 0 ([row iteration][48] stamps), or nullptr
};

#define IPA_TS(id)                                                                  \
  do {                                                                              \
    if (a.dbg && blockIdx.x == 0 && (it) < 8) a.dbg[(it) * 48 + (id)] = clock64();  \
  } while (0)

constexpr int IPA_TILE_BYTES = 32768;
constexpr int IPA_PIMG_TILE = 4096;   // [2 kb][16 rows][128 B]
constexpr int IPA_OZ_LD = 132;
constexpr int IPA_MAX_RZ = 6, IPA_MAX_JB = 8;
constexpr int IPA_EPI = 256;                 // epilogue threads (logits / softmax / down_z): two groups of 128
constexpr int IPA_THREADS = IPA_EPI + 64;    // + MMA issuer warp, loader warp

struct IpaSmemPlan {
  int rz, ctas_per_sm, tmem_cols, single_pass;
  size_t bytes;
};
__host__ __device__ inline size_t ipa_pimg_bytes(int JB) {  // probability images; the region doubles as ozs [H][132] fp32 once GEMM-o has read it
  const size_t p = (size_t)JB * IPA_PIMG_TILE, o = (size_t)NH * IPA_OZ_LD * 4;
  return ((p > o ? p : o) + 1023) / 1024 * 1024;
}
// Shared-memory / TMEM plan.  N <= 384: the ring holds two whole rows, so the tiles GEMM-b consumed are still there when the softmax of
// the row has produced P and GEMM-o runs without a second fetch (single pass over z, one CTA per SM).  Larger N: the tiles of a row
// are fetched twice (the second time from L2); the kernel is then latency-bound per row, so two CTAs per SM are preferred whenever
// they fit (2-slot z ring, <= 256 TMEM columns each), otherwise one CTA with the deepest ring.
inline IpaSmemPlan ipa_core_plan(int N, int max_smem, int max_smem_per_sm, bool force_two_pass = false) {
  const int JB = (N + 127) / 128;
  const size_t other_sp = ipa_pimg_bytes(JB) + 4096 + (size_t)NH * JB * 128 * 4 + 512 + 1024;  // single pass: down_z.weight in registers
  const size_t other = other_sp + C_Z * (C_Z / 4) * 4;                                          // else: staged in shared memory
  const int need_cols = 2 * JB * 16 + 16;  // two D1 buffers + D2
  IpaSmemPlan p;
  p.tmem_cols = need_cols <= 128 ? 128 : (need_cols <= 256 ? 256 : 512);
  p.single_pass = 0;
  const size_t two = other + 2 * (size_t)IPA_TILE_BYTES;
  if (!force_two_pass && 2 * JB <= IPA_MAX_RZ && other_sp + 2 * (size_t)JB * IPA_TILE_BYTES <= (size_t)max_smem) {
    // two whole rows resident (N <= 384): one CTA per SM, every z tile crosses HBM / L2 once
    p.rz = 2 * JB;
    p.ctas_per_sm = 1;
    p.single_pass = 1;
    p.bytes = other_sp + (size_t)p.rz * IPA_TILE_BYTES;
    return p;
  } else if (2 * (two + 1024) <= (size_t)max_smem_per_sm && p.tmem_cols <= 256) {
    p.rz = 2;
    p.ctas_per_sm = 2;
  } else {
    int rz = (int)(((size_t)max_smem - other) / IPA_TILE_BYTES);
    p.rz = rz > IPA_MAX_RZ ? IPA_MAX_RZ : rz;
    p.ctas_per_sm = 1;
  }
  p.bytes = other + (size_t)p.rz * IPA_TILE_BYTES;
  return p;
}

// One CTA streams rows (b, i).  Software pipeline across rows (it = row iteration of this CTA):
//   loader / MMA warp order:  b(0), b(1), { o(it), b(it+2) }          b = GEMM-b into D1[it & 1],  o = GEMM-o (second pass over the
//                                                                      row's z tiles, L2 hits) once the softmax of row it has stored P
//   epilogue order:           logits(0), { softmax(it), logits(it+1), down_z(it) }
// so the HBM loads and GEMM-b of the next rows run under the softmax of the current one, and the logits of the next row are computed
// while the tensor core does GEMM-o.
// SP = true (N <= 384, ring = 2 JB tiles, one CTA per SM): tile (it, t) stays in slot (it JB + t) % rz from its load until GEMM-o of
// row it has read it, so there is no second pass; the MMA warp runs an event loop (GEMM-o of the oldest row whenever its
// probabilities are there, else the next landed GEMM-b tile) and the epilogue keeps the down_z weights and the next S row in registers.
// SP (single pass): one CTA per SM, so a thread may keep its whole row of down_z.weight (128 floats) in registers
template <bool SP>
__global__ void __launch_bounds__(IPA_THREADS, SP ? 1 : 2) ipa_core_kernel(IpaCoreArgs a) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + smem_align_pad(smem_raw);  // offset arithmetic on the __shared__ symbol: accesses stay LDS / STS
  const int N = a.N, JB = a.JB, ldL = JB * 128;
  uint8_t* ring = smem;                                      // rz x 32 KB z tiles
  uint8_t* Pimg = ring + (size_t)a.rz * IPA_TILE_BYTES;      // JB x 4 KB probability images (B operand of GEMM-o)
  float* ozs = reinterpret_cast<float*>(Pimg);               // [H][132], aliases Pimg (dead once GEMM-o has completed)
  uint8_t* Wbs = Pimg + ipa_pimg_bytes(JB);                  // 4 KB
  float* L = reinterpret_cast<float*>(Wbs + 4096);           // [H][ldL] logits / exp
  float* Wd4 = L + NH * ldL;                                 // !SP: [32 channel groups][32 d][4] down_z.weight
  uint64_t* bars = reinterpret_cast<uint64_t*>(Wd4 + (SP ? 0 : C_Z * (C_Z / 4)));
  uint64_t* zfull = bars;                        // [IPA_MAX_RZ]
  uint64_t* zfree = zfull + IPA_MAX_RZ;          // [IPA_MAX_RZ]
  uint64_t* d1_full = zfree + IPA_MAX_RZ;        // [2][IPA_MAX_JB]
  uint64_t* d1_free = d1_full + 2 * IPA_MAX_JB;  // [2]
  uint64_t* p_full = d1_free + 2;                // [1]
  uint64_t* d2_full = p_full + 1;                // [1]
  uint64_t* wb_full = d2_full + 1;               // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wb_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int nrows = (a.rows - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // rows blockIdx.x, + gridDim.x, ...

  if (tid == 0) {
    for (int s = 0; s < IPA_MAX_RZ; ++s) {
      mbar_init(&zfull[s], 1);
      mbar_init(&zfree[s], 1);
    }
    for (int t = 0; t < 2 * IPA_MAX_JB; ++t) mbar_init(&d1_full[t], 1);
    mbar_init(&d1_free[0], IPA_EPI);
    mbar_init(&d1_free[1], IPA_EPI);
    mbar_init(p_full, IPA_EPI);
    mbar_init(d2_full, 1);
    mbar_init(wb_full, 1);
    fence_barrier_init();
  }
  if constexpr (!SP) {
    for (int k = tid; k < C_Z * (C_Z / 4); k += blockDim.x) {
      const int d = k / C_Z, c = k % C_Z;
      Wd4[((c >> 2) * (C_Z / 4) + d) * 4 + (c & 3)] = a.Wd[k];
    }
  }
  if (warp == IPA_EPI / 32) tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t D1 = tmem_base, D2 = tmem_base + 2 * JB * 16;  // D1 buffer b at D1 + b*JB*16

  if (warp == IPA_EPI / 32 + 1) {
    // ============================ loader ============================
    if (elect_one() && nrows > 0) {
      mbar_arrive_expect_tx(wb_full, 4096);
      bulk_g2s(Wbs, a.Wb_img, 4096, wb_full);
      uint32_t cnt = 0;
      auto load_row = [&](int it) {
        const int row = (int)blockIdx.x + it * (int)gridDim.x;
        const uint8_t* zrow = reinterpret_cast<const uint8_t*>(a.z) + (size_t)row * JB * IPA_TILE_BYTES;
        for (int t = 0; t < JB; ++t) {
          const uint32_t s = cnt % a.rz;
          mbar_wait(&zfree[s], ((cnt / a.rz) & 1) ^ 1);
          mbar_arrive_expect_tx(&zfull[s], IPA_TILE_BYTES);
          bulk_g2s(ring + (size_t)s * IPA_TILE_BYTES, zrow + (size_t)t * IPA_TILE_BYTES, IPA_TILE_BYTES, &zfull[s]);
          ++cnt;
        }
      };
      if (a.single_pass) {
        for (int it = 0; it < nrows; ++it) load_row(it);  // a slot is released by GEMM-o of its tile
      } else {
        load_row(0);
        if (nrows > 1) load_row(1);
        for (int it = 0; it < nrows; ++it) {
          load_row(it);                        // second pass of row it (GEMM-o): L2 hits
          if (it + 2 < nrows) load_row(it + 2);
        }
      }
    }
  } else if (warp == IPA_EPI / 32) {
    // ============================ MMA issuer ============================
    if (elect_one() && nrows > 0) {
      mbar_wait(wb_full, 0);
      tc_fence_after();
      const uint32_t idesc_b = make_idesc_f16(128, 16);
      const uint32_t idesc_o = make_idesc_f16(128, 16) | (1u << 15);  // A operand MN-major
      const uint32_t wb = smem_u32(Wbs), pimg = smem_u32(Pimg), ring_u = smem_u32(ring);
      const uint32_t lbo = a.mn_swap ? 1024u : 16384u, sbo = a.mn_swap ? 16384u : 1024u;
      uint32_t cnt = 0;
      auto gemm_b_tile = [&](int it, int t, uint32_t s) {  // tile t of row iteration it (landed in ring slot s): D1[it & 1][t] = z_tile . Wb^T
        const uint32_t zt = ring_u + s * IPA_TILE_BYTES;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(D1 + ((it & 1) * JB + t) * 16, make_sw128_desc(zt + kb * 16384 + k * 32), make_sw128_desc(wb + kb * 2048 + k * 32), idesc_b,
                     (kb | k) ? 1u : 0u);
        umma_commit(&d1_full[(it & 1) * IPA_MAX_JB + t]);
      };
      // GEMM-b of row iteration `it`: D1[it & 1][t] = z_tile . Wb^T
      auto issue_b = [&](int it) {
        const int buf = it & 1;
        if (it >= 2) {  // the logits of row it-2 have been read out of this D1 buffer
          mbar_wait(&d1_free[buf], ((it - 2) >> 1) & 1);
          tc_fence_after();
        }
        for (int t = 0; t < JB; ++t) {
          const uint32_t c_use = a.single_pass ? (uint32_t)(it * JB + t) : cnt;  // single pass: tile (it, t) lives in slot (it JB + t) % rz
          const uint32_t s = c_use % a.rz;
          mbar_wait(&zfull[s], (c_use / a.rz) & 1);
          tc_fence_after();
          gemm_b_tile(it, t, s);
          if (!a.single_pass) umma_commit(&zfree[s]);  // single pass: the tile stays for GEMM-o
          ++cnt;
          if (t == 0) IPA_TS(1);
        }
      };
      // GEMM-o of row iteration `it`: D2 += z_tile^T . P^T  (after the softmax of this row has written the probability images)
      auto issue_o = [&](int it) {
        if (!SP) mbar_wait(p_full, it & 1);  // SP: the caller has polled it
        tc_fence_after();
        IPA_TS(9);
        for (int t = 0; t < JB; ++t) {
          const uint32_t c_use = a.single_pass ? (uint32_t)(it * JB + t) : cnt;
          const uint32_t s = c_use % a.rz;
          if (!a.single_pass) {  // (single pass: GEMM-b of this tile has already waited for it)
            mbar_wait(&zfull[s], (c_use / a.rz) & 1);
            tc_fence_after();
          }
          const uint32_t zt = ring_u + s * IPA_TILE_BYTES;
          // one descriptor pair per tile; the 8 steps advance the start-address fields by compile-time constants (the 14-bit field
          // holds address >> 4 and shared-memory addresses stay below 256 KB, so the additions never carry out of it)
          const uint64_t dz0 = make_sw128_desc_ls(zt, lbo, sbo), dp0 = make_sw128_desc(pimg + t * IPA_PIMG_TILE);
#pragma unroll
          for (int k = 0; k < 8; ++k)  // 16 j-rows per step
            umma_f16(D2, dz0 + (uint64_t)((k * 2048) >> 4), dp0 + (uint64_t)(((k >> 2) * 2048 + (k & 3) * 32) >> 4), idesc_o, (t | k) ? 1u : 0u);
          umma_commit(&zfree[s]);
          ++cnt;
        }
        umma_commit(d2_full);
        IPA_TS(10);
      };
      if constexpr (SP) {
        // event loop: GEMM-o of the oldest row as soon as its probabilities are there (the epilogue's critical path), otherwise the next
        // GEMM-b tile that has landed -- a blocking wait for an HBM tile would delay GEMM-o by a whole memory latency
        int io = 0, bt = 0;
        const int total_b = nrows * JB;
        long long t_idle = clock64();  // bounded like mbar_wait: a protocol bug must trap instead of hanging the GPU
        while (io < nrows) {
          if (mbar_try_wait(p_full, io & 1)) {
            issue_o(io);
            ++io;
            t_idle = clock64();
            continue;
          }
          if (clock64() - t_idle > 4000000000LL) __trap();
          if (bt < total_b) {
            const int it_b = bt / JB, t = bt - it_b * JB;
            if (t == 0 && it_b >= 2 && !mbar_try_wait(&d1_free[it_b & 1], ((it_b - 2) >> 1) & 1)) continue;  // logits of row it_b - 2 not read yet
            const uint32_t s = (uint32_t)bt % a.rz;
            if (mbar_try_wait(&zfull[s], ((uint32_t)bt / a.rz) & 1)) {
              tc_fence_after();
              gemm_b_tile(it_b, t, s);
              t_idle = clock64();
              if (t == 0) {
                const int it = it_b;
                IPA_TS(1);
              }
              ++bt;
            }
          }
        }
      } else {
        issue_b(0);
        if (nrows > 1) issue_b(1);
        for (int it = 0; it < nrows; ++it) {
          issue_o(it);
          if (it + 2 < nrows) issue_b(it + 2);
        }
      }
    }
  } else {
    // ============================ logits / softmax / down_z (256 threads) ============================
    // two groups of 128 threads: thread r of either group <-> TMEM lane r (key j of a tile in D1, channel c in D2); group eg owns heads
    // 4 eg .. 4 eg + 3 wherever the work splits by head (logits, o_pair read-out), warp w owns head w in the softmax, and down_z has
    // one output per thread
    const int r = tid & 127, eg = tid >> 7;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const float s_b = sqrtf(1.0f / 3.f);
    float bbv[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) bbv[h] = a.bb[4 * eg + h];
    const int hh_out = tid >> 5, dq = tid & 31;
    const float bd0 = a.bd[dq];
    // SP: down_z from registers.  Warp w: outputs d in [16 (w & 1), +16) of heads 2 (w >> 1), 2 (w >> 1) + 1; lane: d = that range +
    // (lane & 15), channel half lane >> 4 (64 weights in registers), the two halves are added with one shuffle
    const int sp_d = 16 * (warp & 1) + (lane & 15), sp_half = lane >> 4, sp_h0 = 2 * (warp >> 1);
    float wdr[SP ? C_Z / 2 : 1];
    if constexpr (SP) {
#pragma unroll
      for (int c = 0; c < C_Z / 2; c += 4) {
        const float4 w = *reinterpret_cast<const float4*>(a.Wd + sp_d * C_Z + sp_half * (C_Z / 2) + c);
        wdr[c] = w.x; wdr[c + 1] = w.y; wdr[c + 2] = w.z; wdr[c + 3] = w.w;
      }
    }
    const long long hs = (long long)N * a.ldS;
    // S row of iteration `it` (8 heads x ldS floats) -> L, asynchronously (cp.async, 16 bytes per request)
    auto prefetch_S = [&](int it) {
      const int row = (int)blockIdx.x + it * (int)gridDim.x;
      const int b = row / N, i = row - b * N;
      const float* Srow0 = a.S + (((long long)b * NH) * N + i) * a.ldS;
      const int cpr = a.ldS >> 2;  // 16-byte chunks per head
      for (int k = tid; k < NH * cpr; k += IPA_EPI) {
        const int h = k / cpr, c = k - h * cpr;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(L + h * ldL + 4 * c)), "l"(Srow0 + h * hs + 4 * c) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // SP: the S values a thread adds its bias to (4 heads x JB keys) are loaded into registers one row ahead (issued before the softmax
    // of the previous row, consumed after it): no cp.async into L, whose single buffer is busy until that softmax has finished
    float sreg[SP ? 3 : 1][4];
    auto prefetch_S_regs = [&](int it) {
      if constexpr (SP) {
        const int row = (int)blockIdx.x + it * (int)gridDim.x;
        const int b = row / N, i = row - b * N;
        const float* Srow0 = a.S + (((long long)b * NH + 4 * eg) * N + i) * a.ldS;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          const int j = t * 128 + r;
#pragma unroll
          for (int h = 0; h < 4; ++h) sreg[t][h] = (t < JB && j < N) ? Srow0[h * hs + j] : 0.f;
        }
      }
    };
    // logits of row iteration `it`: L[h][j] = S'[h][j] + sqrt(1/3) (b_ij,h + b_b,h), -inf beyond N
    auto logits = [&](int it) {
      const int buf = it & 1;
      if constexpr (SP) {
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          if (t >= JB) break;
          const int j = t * 128 + r;
          mbar_wait(&d1_full[buf * IPA_MAX_JB + t], (it >> 1) & 1);
          tc_fence_after();
          float bv[16];
          tmem_ld16(D1 + lane_base + (buf * JB + t) * 16, bv);
          tmem_ld_wait();
#pragma unroll
          for (int h = 0; h < 4; ++h)
            L[(4 * eg + h) * ldL + j] = j < N ? sreg[t][h] + s_b * ((eg ? bv[4 + h] + bv[12 + h] : bv[h] + bv[8 + h]) + bbv[h]) : -INFINITY;
        }
        tc_fence_before();
        mbar_arrive(&d1_free[buf]);
        return;
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      asm volatile("bar.sync 1, %0;" ::"n"(IPA_EPI) : "memory");  // every thread's part of the S row has landed
      for (int t = 0; t < JB; ++t) {
        const int j = t * 128 + r;
        mbar_wait(&d1_full[buf * IPA_MAX_JB + t], (it >> 1) & 1);
        tc_fence_after();
        float bv[16];
        tmem_ld16(D1 + lane_base + (buf * JB + t) * 16, bv);
        tmem_ld_wait();
        if (j < N) {
#pragma unroll
          for (int h = 0; h < 4; ++h)  // constant register indices (a runtime index would put bv on the stack)
            L[(4 * eg + h) * ldL + j] += s_b * ((eg ? bv[4 + h] + bv[12 + h] : bv[h] + bv[8 + h]) + bbv[h]);
        } else {
#pragma unroll
          for (int h = 0; h < 4; ++h) L[(4 * eg + h) * ldL + j] = -INFINITY;
        }
      }
      tc_fence_before();
      mbar_arrive(&d1_free[buf]);
    };
    if (nrows > 0) {
      if constexpr (SP) prefetch_S_regs(0);
      else prefetch_S(0);
      logits(0);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(IPA_EPI) : "memory");
    for (int it = 0; it < nrows; ++it) {
      const int row = (int)blockIdx.x + it * (int)gridDim.x;
      const int b = row / N, i = row - b * N;
      float* Srow0 = a.S + (((long long)b * NH) * N + i) * a.ldS;
      if (tid == 0) IPA_TS(20);
      if (SP && it + 1 < nrows) prefetch_S_regs(it + 1);  // loads in flight under the softmax
      // ---- softmax over j: warp w owns head w; a lane owns 4 consecutive j per 128-column tile
      {
        const int h = warp;
        const float4* L4 = reinterpret_cast<const float4*>(L + h * ldL);
        float4 v[IPA_MAX_JB];
        float mx = -INFINITY;
#pragma unroll
        for (int q = 0; q < IPA_MAX_JB; ++q)
          if (q < JB) {
            v[q] = L4[lane + 32 * q];
            mx = fmaxf(mx, fmaxf(fmaxf(v[q].x, v[q].y), fmaxf(v[q].z, v[q].w)));
          }
        mx = warp_max(mx);
        float sum = 0.f;
        constexpr float LOG2E = 1.4426950408889634f;
#pragma unroll
        for (int q = 0; q < IPA_MAX_JB; ++q)
          if (q < JB) {
            v[q].x = ex2_approx((v[q].x - mx) * LOG2E);
            v[q].y = ex2_approx((v[q].y - mx) * LOG2E);
            v[q].z = ex2_approx((v[q].z - mx) * LOG2E);
            v[q].w = ex2_approx((v[q].w - mx) * LOG2E);
            sum += (v[q].x + v[q].y) + (v[q].z + v[q].w);
          }
        sum = warp_sum(sum);
        const float inv = 1.f / sum;
        float* Sh = Srow0 + h * hs;
#pragma unroll
        for (int q = 0; q < IPA_MAX_JB; ++q)
          if (q < JB) {
            const float4 p = make_float4(v[q].x * inv, v[q].y * inv, v[q].z * inv, v[q].w * inv);
            const int j = 4 * (lane + 32 * q);
            if (!a.Pg && j < a.ldS) *reinterpret_cast<float4*>(Sh + j) = p;  // columns in [N, ldS) are zero
            const __half2 h01 = __floats2half2_rn(p.x, p.y), h23 = __floats2half2_rn(p.z, p.w);
            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
            const __half2 l01 = __floats2half2_rn(p.x - f01.x, p.y - f01.y), l23 = __floats2half2_rn(p.z - f23.x, p.w - f23.y);
            if (a.Pg) {
              // operand image of the A.V GEMM: row i of tile i >> 7, k-block j >> 6; the low part carries the 2^11 scale of gemm_img's
              // split (its epilogue divides the cross accumulator by 2048); columns j >= N hold exact zeros (exp(-inf))
              const __half2 s01 = __floats2half2_rn((p.x - f01.x) * 2048.f, (p.y - f01.y) * 2048.f);
              const __half2 s23 = __floats2half2_rn((p.z - f23.x) * 2048.f, (p.w - f23.y) * 2048.f);
              uint8_t* gd = a.Pg + ((((size_t)b * NH + h) * JB + (i >> 7)) * (2 * JB) + (j >> 6)) * (size_t)32768 + (i & 127) * 128 +
                            (((((j & 63) >> 3)) ^ (i & 7)) << 4) + (j & 7) * 2;
              *reinterpret_cast<uint2*>(gd) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
              *reinterpret_cast<uint2*>(gd + 16384) = make_uint2(*reinterpret_cast<const uint32_t*>(&s01), *reinterpret_cast<const uint32_t*>(&s23));
            }
            const int jj = j & 127;
            uint8_t* dst = Pimg + q * IPA_PIMG_TILE + (jj >> 6) * 2048 + h * 128 + ((((jj & 63) >> 3) ^ (h & 7)) << 4) + (jj & 7) * 2;
            *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
            *reinterpret_cast<uint2*>(dst + 1024) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
          }
      }
      fence_proxy_async();
      mbar_arrive(p_full);
      if (tid == 0) IPA_TS(31);
      asm volatile("bar.sync 1, %0;" ::"n"(IPA_EPI) : "memory");  // every warp is done with L
      if (it + 1 < nrows) {                            // overlaps GEMM-o of this row
        if constexpr (!SP) prefetch_S(it + 1);
        logits(it + 1);
      }
      if (tid == 0) IPA_TS(30);
      // ---- o_pair: D2[c, h] (hi + lo columns) -> ozs[h][c] -> down_z
      mbar_wait(d2_full, it & 1);
      tc_fence_after();
      if (tid == 0) IPA_TS(32);
      {
        float ov[16];
        tmem_ld16(D2 + lane_base, ov);
        tmem_ld_wait();
#pragma unroll
        for (int h = 0; h < 4; ++h) ozs[(4 * eg + h) * IPA_OZ_LD + r] = eg ? ov[4 + h] + ov[12 + h] : ov[h] + ov[8 + h];
      }
      tc_fence_before();
      asm volatile("bar.sync 1, %0;" ::"n"(IPA_EPI) : "memory");
      if constexpr (SP) {
        float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const float4* o4 = reinterpret_cast<const float4*>(ozs + (sp_h0 + hh) * IPA_OZ_LD + sp_half * (C_Z / 2));
#pragma unroll
          for (int cgp = 0; cgp < C_Z / 8; ++cgp) {
            const float4 x = o4[cgp];  // two distinct addresses per warp (one per channel half): broadcasts
            acc[hh][0] = fmaf(wdr[4 * cgp], x.x, fmaf(wdr[4 * cgp + 1], x.y, acc[hh][0]));
            acc[hh][1] = fmaf(wdr[4 * cgp + 2], x.z, fmaf(wdr[4 * cgp + 3], x.w, acc[hh][1]));
          }
        }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float v = acc[hh][0] + acc[hh][1];
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if (sp_half == hh) a.cat[(long long)row * CAT + CATP_PAIR + (sp_h0 + hh) * (C_Z / 4) + sp_d] = v + a.bd[sp_d];
        }
      } else {
        float acc0 = bd0, acc2 = 0.f;
        const float4* o4 = reinterpret_cast<const float4*>(ozs + hh_out * IPA_OZ_LD);
        const float4* w4 = reinterpret_cast<const float4*>(Wd4);
#pragma unroll 8
        for (int cgp = 0; cgp < C_Z / 4; ++cgp) {
          const float4 x = o4[cgp];
          const float4 w0 = w4[cgp * (C_Z / 4) + dq];
          acc0 = fmaf(w0.x, x.x, fmaf(w0.y, x.y, acc0));
          acc2 = fmaf(w0.z, x.z, fmaf(w0.w, x.w, acc2));
        }
        a.cat[(long long)row * CAT + CATP_PAIR + hh_out * (C_Z / 4) + dq] = acc0 + acc2;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(IPA_EPI) : "memory");  // ozs aliases the probability images: the next softmax may not start earlier
      if (tid == 0) IPA_TS(33);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == IPA_EPI / 32) tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
}

}  // namespace fdpt
