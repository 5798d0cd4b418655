// kernels_ipa.cuh — Invariant Point Attention (ipa_pytorch.py:170-329), pair-side core.
//
// Data flow of one IPA call (B samples, N residues, H=8 heads, C=256, Pq=8, Pv=12, c_z=128):
//   q, kv, q_pts_raw, kv_pts_raw   <- Linear(s)                       (gemm)
//   q_pts/k_pts/v_pts              <- R_i p + t_i                      (ipa_points_kernel)
//   S[b,h,i,j]                     <- q . k                            (batched gemm, raw dot products)
//   ipa_core_kernel (one CTA per (b,i), one warp per head):
//        logits = sqrt(1/(3C)) S + sqrt(1/3) (W_b z_ij + b_b) - 0.5 gamma_h sum_p |q_p - k_p|^2 + 1e5 (m_i m_j - 1)
//        a = softmax_j(logits)  -> written back over S
//        o_pt  = R_i^T (sum_j a v_pts_j - t_i), |o_pt|          -> cat[:, 2048:2432]
//        o_pair = down_z(sum_j a_ij z_ij)  (down_z is linear and sum_j a = 1, so it commutes) -> cat[:, 2432:2688]
//   o[b,i,h,:] = sum_j a v                                            (batched gemm into cat[:, :2048])
//   out = linear_out(cat)                                             (gemm)
// z[b,i,:,:] (N x 512 B, contiguous) is streamed from HBM once per (b,i); the second use (o_pair) re-reads it
// while it is still L2 resident.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace fdpt {

// raw point projections (x-block | y-block | z-block, ipa_pytorch.py:214-239) -> global-frame points
//   q_pts [M,H,PQ,3], k_pts [M,H,PQ,3], v_pts [M,H,PV,3]
__global__ void ipa_points_kernel(int M, const float* __restrict__ qp_raw, const float* __restrict__ kvp_raw,
                                  const float* __restrict__ quats, const float* __restrict__ trans,
                                  float* __restrict__ q_pts, float* __restrict__ k_pts, float* __restrict__ v_pts) {
  const int m = blockIdx.x;
  __shared__ float R[9], t[3];
  if (threadIdx.x == 0) {
    float q[4] = {quats[m * 4], quats[m * 4 + 1], quats[m * 4 + 2], quats[m * 4 + 3]};
    quat_to_rot(q, R);
    t[0] = trans[m * 3];
    t[1] = trans[m * 3 + 1];
    t[2] = trans[m * 3 + 2];
  }
  __syncthreads();
  constexpr int NQ = NH * PQ, NKV = NH * (PQ + PV);
  for (int idx = threadIdx.x; idx < NQ + NKV; idx += blockDim.x) {
    float x, y, z;
    float* dst;
    if (idx < NQ) {
      const float* r = qp_raw + (long long)m * (3 * NQ);
      x = r[idx], y = r[NQ + idx], z = r[2 * NQ + idx];
      dst = q_pts + ((long long)m * NQ + idx) * 3;
    } else {
      const int k = idx - NQ;
      const float* r = kvp_raw + (long long)m * (3 * NKV);
      x = r[k], y = r[NKV + k], z = r[2 * NKV + k];
      const int h = k / (PQ + PV), p = k - h * (PQ + PV);
      dst = (p < PQ) ? k_pts + (((long long)m * NH + h) * PQ + p) * 3 : v_pts + (((long long)m * NH + h) * PV + (p - PQ)) * 3;
    }
    dst[0] = R[0] * x + R[1] * y + R[2] * z + t[0];
    dst[1] = R[3] * x + R[4] * y + R[5] * z + t[1];
    dst[2] = R[6] * x + R[7] * y + R[8] * z + t[2];
  }
}

struct IpaCoreArgs {
  int B, N;
  float* S;                 // [B,H,N,N] in: q.k ; out: attention probabilities
  const __half* z;          // fp16 tile images [B][N][JB][2 k-blocks][128 rows][128 B, 128B-swizzled] (et_fused.cuh)
  int JB;
  const float* q_pts;       // [B,N,H,PQ,3]
  const float* k_pts;       // [B,N,H,PQ,3]
  const float* v_pts;       // [B,N,H,PV,3]
  const float* quats;       // [B,N,4]
  const float* trans;       // [B,N,3]
  const float* mask;        // [B,N]
  const float* Wb;          // [H,128] linear_b.weight
  const float* bb;          // [H]
  const float* head_w;      // [H] raw head_weights (softplus applied here)
  const float* Wd;          // [32,128] down_z.weight
  const float* bd;          // [32]
  float* cat;               // [B*N, 2688]
};

constexpr int IPA_JC = 32;            // j-chunk
constexpr int IPA_ZLD = C_Z + 4;      // padded smem row (conflict-free float4 rows)

__global__ void __launch_bounds__(256) ipa_core_kernel(IpaCoreArgs a) {
  extern __shared__ float smem[];
  const int N = a.N;
  const int i = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, h = tid >> 5, lane = tid & 31;
  const long long m = (long long)b * N + i;

  float* L = smem;                          // [H][N]
  float* zt = L + NH * N;                   // [32][IPA_ZLD]
  float* wb = zt + IPA_JC * IPA_ZLD;        // [H][128]
  float* kp = wb + NH * C_Z;                // [32][H*PQ*3 = 192] (+1 pad per row)
  float* qp = kp + IPA_JC * (NH * PQ * 3 + 1);  // [H][24]
  float* oz = qp + NH * PQ * 3;             // [H][128]
  float* opt = oz + NH * C_Z;               // [H][36]

  for (int k = tid; k < NH * C_Z; k += 256) wb[k] = a.Wb[k];
  for (int k = tid; k < NH * PQ * 3; k += 256) qp[k] = a.q_pts[m * (NH * PQ * 3) + k];
  const float mi = a.mask[m];
  const float sp = log1pf(expf(a.head_w[h]));  // softplus
  const float gamma = sp * sqrtf(1.0f / (3.f * (PQ * 9.0f / 2.f)));
  const float s_qk = sqrtf(1.0f / (3.f * C_HID)), s_b = sqrtf(1.0f / 3.f);
  const float bbh = a.bb[h];
  float* Srow = a.S + (((long long)b * NH + h) * N + i) * N;
  const uint8_t* zrow = reinterpret_cast<const uint8_t*>(a.z) + (size_t)m * a.JB * 32768;
  __syncthreads();

  // ---- pass 1: logits -------------------------------------------------------------------------
  for (int j0 = 0; j0 < N; j0 += IPA_JC) {
    const int nj = min(IPA_JC, N - j0);
    for (int k = tid; k < IPA_JC * (C_Z / 8); k += 256) {  // one 16-byte chunk (8 halfs) per iteration
      const int jj = k / (C_Z / 8), kc = k % (C_Z / 8);
      float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (jj < nj) {
        const int j = j0 + jj, r = j & 127;
        const uint8_t* src = zrow + (size_t)(j >> 7) * 32768 + (kc >> 3) * 16384 + r * 128 + (((kc & 7) ^ (r & 7)) << 4);
        const uint4 u = *reinterpret_cast<const uint4*>(src);
        const __half2* hh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 t2 = __half22float2(hh[e]);
          f[2 * e] = t2.x;
          f[2 * e + 1] = t2.y;
        }
      }
      *reinterpret_cast<float4*>(zt + jj * IPA_ZLD + kc * 8) = make_float4(f[0], f[1], f[2], f[3]);
      *reinterpret_cast<float4*>(zt + jj * IPA_ZLD + kc * 8 + 4) = make_float4(f[4], f[5], f[6], f[7]);
    }
    const float* kpg = a.k_pts + ((long long)b * N + j0) * (NH * PQ * 3);
    for (int k = tid; k < IPA_JC * NH * PQ * 3; k += 256) {
      const int jj = k / (NH * PQ * 3), c = k % (NH * PQ * 3);
      kp[jj * (NH * PQ * 3 + 1) + c] = (jj < nj) ? kpg[k] : 0.f;
    }
    __syncthreads();
    if (lane < nj) {
      const int j = j0 + lane;
      float bias = 0.f;
      const float* zr = zt + lane * IPA_ZLD;
      const float* wr = wb + h * C_Z;
#pragma unroll 8
      for (int c = 0; c < C_Z; c += 4) {
        const float4 zv = *reinterpret_cast<const float4*>(zr + c);
        const float4 wv = *reinterpret_cast<const float4*>(wr + c);
        bias = fmaf(zv.x, wv.x, bias);
        bias = fmaf(zv.y, wv.y, bias);
        bias = fmaf(zv.z, wv.z, bias);
        bias = fmaf(zv.w, wv.w, bias);
      }
      bias += bbh;
      float d2 = 0.f;
      const float* kr = kp + lane * (NH * PQ * 3 + 1) + h * (PQ * 3);
      const float* qr = qp + h * (PQ * 3);
#pragma unroll
      for (int c = 0; c < PQ * 3; ++c) {
        const float d = qr[c] - kr[c];
        d2 = fmaf(d, d, d2);
      }
      const float mj = a.mask[(long long)b * N + j];
      L[h * N + j] = s_qk * Srow[j] + s_b * bias - 0.5f * gamma * d2 + 1e5f * (mi * mj - 1.f);
    }
    __syncthreads();
  }

  // ---- softmax over j (warp h owns row h) -------------------------------------------------------
  {
    float mx = -INFINITY;
    for (int j = lane; j < N; j += 32) mx = fmaxf(mx, L[h * N + j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < N; j += 32) {
      const float e = expf(L[h * N + j] - mx);
      L[h * N + j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int j = lane; j < N; j += 32) {
      const float p = L[h * N + j] * inv;
      L[h * N + j] = p;
      Srow[j] = p;
    }
  }
  __syncthreads();

  // ---- pass 2: o_pair accumulators (thread = (h, 4 channels)), o_pt partials (lane = j) -------------
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float pt[PV * 3];
#pragma unroll
  for (int k = 0; k < PV * 3; ++k) pt[k] = 0.f;
  for (int j0 = 0; j0 < N; j0 += IPA_JC) {
    const int nj = min(IPA_JC, N - j0);
    for (int k = tid; k < IPA_JC * (C_Z / 8); k += 256) {  // one 16-byte chunk (8 halfs) per iteration
      const int jj = k / (C_Z / 8), kc = k % (C_Z / 8);
      float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (jj < nj) {
        const int j = j0 + jj, r = j & 127;
        const uint8_t* src = zrow + (size_t)(j >> 7) * 32768 + (kc >> 3) * 16384 + r * 128 + (((kc & 7) ^ (r & 7)) << 4);
        const uint4 u = *reinterpret_cast<const uint4*>(src);
        const __half2* hh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 t2 = __half22float2(hh[e]);
          f[2 * e] = t2.x;
          f[2 * e + 1] = t2.y;
        }
      }
      *reinterpret_cast<float4*>(zt + jj * IPA_ZLD + kc * 8) = make_float4(f[0], f[1], f[2], f[3]);
      *reinterpret_cast<float4*>(zt + jj * IPA_ZLD + kc * 8 + 4) = make_float4(f[4], f[5], f[6], f[7]);
    }
    __syncthreads();
    for (int jj = 0; jj < nj; ++jj) {
      const float p = L[h * N + j0 + jj];
      const float4 zv = *reinterpret_cast<const float4*>(zt + jj * IPA_ZLD + lane * 4);
      acc[0] = fmaf(p, zv.x, acc[0]);
      acc[1] = fmaf(p, zv.y, acc[1]);
      acc[2] = fmaf(p, zv.z, acc[2]);
      acc[3] = fmaf(p, zv.w, acc[3]);
    }
    if (lane < nj) {
      const int j = j0 + lane;
      const float p = L[h * N + j];
      const float4* vp = reinterpret_cast<const float4*>(a.v_pts + (((long long)b * N + j) * NH + h) * (PV * 3));
#pragma unroll
      for (int k = 0; k < PV * 3 / 4; ++k) {
        const float4 v = __ldg(vp + k);
        pt[4 * k + 0] = fmaf(p, v.x, pt[4 * k + 0]);
        pt[4 * k + 1] = fmaf(p, v.y, pt[4 * k + 1]);
        pt[4 * k + 2] = fmaf(p, v.z, pt[4 * k + 2]);
        pt[4 * k + 3] = fmaf(p, v.w, pt[4 * k + 3]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) oz[h * C_Z + lane * 4 + k] = acc[k];
#pragma unroll
  for (int k = 0; k < PV * 3; ++k) {
    const float s = warp_sum(pt[k]);
    if (lane == 0) opt[h * (PV * 3) + k] = s;
  }
  __syncthreads();

  // ---- epilogue ----------------------------------------------------------------------------------
  float* cat = a.cat + m * CAT;
  {  // o_pair = down_z(sum_j a z): thread (h, c = lane)
    const float* w = a.Wd + lane * C_Z;
    const float* o = oz + h * C_Z;
    float s = 0.f;
#pragma unroll 8
    for (int c = 0; c < C_Z; ++c) s = fmaf(__ldg(w + c), o[c], s);
    cat[CAT_PAIR + h * (C_Z / 4) + lane] = s + a.bd[lane];
  }
  if (lane < PV) {  // o_pt: R^T (p - t), norms
    float q[4] = {a.quats[m * 4], a.quats[m * 4 + 1], a.quats[m * 4 + 2], a.quats[m * 4 + 3]};
    float R[9];
    quat_to_rot(q, R);
    const float* o = opt + h * (PV * 3) + lane * 3;
    const float x = o[0] - a.trans[m * 3], y = o[1] - a.trans[m * 3 + 1], zc = o[2] - a.trans[m * 3 + 2];
    const float lx = R[0] * x + R[3] * y + R[6] * zc;
    const float ly = R[1] * x + R[4] * y + R[7] * zc;
    const float lz = R[2] * x + R[5] * y + R[8] * zc;
    const int k = h * PV + lane;
    cat[CAT_OPT + k] = lx;
    cat[CAT_OPT + NH * PV + k] = ly;
    cat[CAT_OPT + 2 * NH * PV + k] = lz;
    cat[CAT_NRM + k] = sqrtf(lx * lx + ly * ly + lz * lz + 1e-8f);
  }
}

inline size_t ipa_core_smem_bytes(int N) {
  return sizeof(float) * ((size_t)NH * N + IPA_JC * IPA_ZLD + NH * C_Z + IPA_JC * (NH * PQ * 3 + 1) + NH * PQ * 3 + NH * C_Z +
                          NH * PV * 3);
}

}  // namespace fdpt
