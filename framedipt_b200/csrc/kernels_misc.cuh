// kernels_misc.cuh — LayerNorm, feature building, frame algebra, small elementwise kernels.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace fdpt {

// ------------------------------------------------------------------------------------------------
// LayerNorm over the last dim C (eps 1e-5, torch default), one warp per row.
//   y = LN(x) * gamma + beta ; optionally y *= rowmask[row] (node mask) or pair mask m[b,i]*m[b,j].
// x and y may alias.
// ------------------------------------------------------------------------------------------------
// y = LN(residual + inmask * (sum_k part_k + bias)): the reduce step of a split-K GEMM fused into the LayerNorm that follows it
// (IPA linear_out, K = 2688: the K thirds run as three batches of the GEMM kernel and cut its per-CTA latency; parts are `pstride`
// floats apart).
template <int C>
__global__ void __launch_bounds__(256) sumk_layernorm_kernel(const float* __restrict__ parts, long long pstride, int nparts,
                                                             const float* __restrict__ bias, const float* __restrict__ residual,
                                                             const float* __restrict__ inmask, float* __restrict__ y,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta, long long rows) {
  // a lane owns C / 32 CONSECUTIVE columns (16-byte loads); the loop over the K parts is the outer one, so that every part costs one
  // memory round trip for the whole row (with the part loop innermost the loads of a column formed a serial chain: 13.8 -> 5 us)
  constexpr int PER = C / 32, V4 = PER / 4;
  static_assert(PER % 4 == 0, "C must be a multiple of 128");
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float im = inmask ? inmask[row] : 1.f;
  const int c0 = lane * PER;
  float4 acc[V4];
#pragma unroll
  for (int i = 0; i < V4; ++i) acc[i] = *reinterpret_cast<const float4*>(parts + row * C + c0 + 4 * i);
  for (int k = 1; k < nparts; ++k) {  // fixed order: deterministic
    float4 t[V4];
#pragma unroll
    for (int i = 0; i < V4; ++i) t[i] = *reinterpret_cast<const float4*>(parts + k * pstride + row * C + c0 + 4 * i);
#pragma unroll
    for (int i = 0; i < V4; ++i) acc[i] = make_float4(acc[i].x + t[i].x, acc[i].y + t[i].y, acc[i].z + t[i].z, acc[i].w + t[i].w);
  }
  float v[PER];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V4; ++i) {
    const float4 bs = *reinterpret_cast<const float4*>(bias + c0 + 4 * i);
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (residual) r = *reinterpret_cast<const float4*>(residual + row * C + c0 + 4 * i);
    v[4 * i + 0] = (acc[i].x + bs.x) * im + r.x;
    v[4 * i + 1] = (acc[i].y + bs.y) * im + r.y;
    v[4 * i + 2] = (acc[i].z + bs.z) * im + r.z;
    v[4 * i + 3] = (acc[i].w + bs.w) * im + r.w;
    s += (v[4 * i] + v[4 * i + 1]) + (v[4 * i + 2] + v[4 * i + 3]);
  }
  const float mean = warp_sum(s) * (1.f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const float d = v[i] - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + 1e-5f);
#pragma unroll
  for (int i = 0; i < V4; ++i) {
    const float4 g = *reinterpret_cast<const float4*>(gamma + c0 + 4 * i), be = *reinterpret_cast<const float4*>(beta + c0 + 4 * i);
    *reinterpret_cast<float4*>(y + row * C + c0 + 4 * i) =
        make_float4((v[4 * i] - mean) * rstd * g.x + be.x, (v[4 * i + 1] - mean) * rstd * g.y + be.y, (v[4 * i + 2] - mean) * rstd * g.z + be.z,
                    (v[4 * i + 3] - mean) * rstd * g.w + be.w);
  }
}


template <int C>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* x, float* y, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, long long rows,
                                                        const float* __restrict__ rowmask,
                                                        const float* __restrict__ pairmask, int nres, long long row0) {
  constexpr int PER = C / 32;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + row * C;
  float v[PER];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    v[i] = xr[lane + 32 * i];
    s += v[i];
  }
  const float mean = warp_sum(s) * (1.f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const float d = v[i] - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + 1e-5f);
  float m = 1.f;
  if (rowmask) m = rowmask[row];
  if (pairmask) {
    const long long p = row0 + row;
    const long long nn = (long long)nres * nres;
    const long long b = p / nn;
    const int rem = (int)(p - b * nn);
    const int i = rem / nres, j = rem - i * nres;
    m = pairmask[b * nres + i] * pairmask[b * nres + j];
  }
  float* yr = y + row * C;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = lane + 32 * i;
    yr[c] = ((v[i] - mean) * rstd * gamma[c] + beta[c]) * m;
  }
}

// ------------------------------------------------------------------------------------------------
// Node features (Embedder.forward, score_network.py:153-182):
//   feat1d[m, :F1] = [onehot21(aatype) | t_emb (eps row on fixed residues) | fixed_mask]      (F1 = 54, or 33 without aatype)
//   node_feat[m, :F1+32] = [feat1d | idx_emb]
// ------------------------------------------------------------------------------------------------
__global__ void node_feats_kernel(int M, int N, int with_aatype, const int32_t* __restrict__ aatype,
                                  const float* __restrict__ fixed_mask, const float* __restrict__ t_emb,
                                  const float* __restrict__ t_emb_eps, const float* __restrict__ idx_emb,
                                  float* __restrict__ feat1d, float* __restrict__ node_feat) {
  const int m = blockIdx.x * blockDim.y + threadIdx.y;
  if (m >= M) return;
  const int b = m / N;
  const int F1 = with_aatype ? 54 : 33;
  const int FN = F1 + EMB;
  const float fm = fixed_mask[m];
  const int aa = with_aatype ? aatype[m] : 0;
  for (int c = threadIdx.x; c < FN; c += blockDim.x) {
    float v;
    int cc = c;
    if (with_aatype) {
      if (c < 21) {
        v = (c == aa) ? 1.f : 0.f;
        cc = -1;
      } else {
        cc = c - 21;
      }
    }
    if (cc >= 0) {
      if (cc < EMB) {
        v = (with_aatype && fm != 0.f) ? t_emb_eps[cc] : t_emb[b * EMB + cc];
      } else if (cc == EMB) {
        v = fm;
      } else {
        v = idx_emb[(long long)m * EMB + (cc - EMB - 1)];
      }
    }
    node_feat[(long long)m * FN + c] = v;
    if (c < F1) feat1d[(long long)m * F1 + c] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// small strided copy:  dst[m, dcol0 + c] = src[m, c] * (rowmask ? rowmask[m] : 1)   for c < cols
// ------------------------------------------------------------------------------------------------
__global__ void copy_cols_kernel(long long M, int cols, const float* __restrict__ src, int lds, float* __restrict__ dst,
                                 int ldd, int dcol0, const float* __restrict__ rowmask) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * cols) return;
  const long long m = idx / cols;
  const int c = (int)(idx - m * cols);
  float v = src[m * lds + c];
  if (rowmask) v *= rowmask[m];
  dst[m * ldd + dcol0 + c] = v;
}

// ------------------------------------------------------------------------------------------------
// frames init (IpaScore.forward, ipa_pytorch.py:517-524): quats = rigids_t[..., :4], trans = rigids_t[..., 4:] * 0.1
// plus diffuse_mask = (1 - fixed) * res_mask
// ------------------------------------------------------------------------------------------------
__global__ void init_frames_kernel(int M, const float* __restrict__ rigids, float scale, const float* __restrict__ res_mask,
                                   const float* __restrict__ fixed_mask, float* __restrict__ quats,
                                   float* __restrict__ trans, float* __restrict__ diffuse_mask) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
#pragma unroll
  for (int k = 0; k < 4; ++k) quats[m * 4 + k] = rigids[m * 7 + k];
#pragma unroll
  for (int k = 0; k < 3; ++k) trans[m * 3 + k] = rigids[m * 7 + 4 + k] * scale;
  diffuse_mask[m] = (1.f - fixed_mask[m]) * res_mask[m];
}

// final frames: rigids[m] = [quats | trans / cs]   (unscale_rigids, ipa_pytorch.py:495-507: fp32 divide)
__global__ void finish_frames_kernel(int M, const float* __restrict__ quats, const float* __restrict__ trans, float cs,
                                     float* __restrict__ rigids) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
#pragma unroll
  for (int k = 0; k < 4; ++k) rigids[m * 7 + k] = quats[m * 4 + k];
#pragma unroll
  for (int k = 0; k < 3; ++k) rigids[m * 7 + 4 + k] = __fdiv_rn(trans[m * 3 + k], cs);
}

// ------------------------------------------------------------------------------------------------
// Rigid.compose_q_update_vec (rigid_utils.py:1039-1063, 587-616, 266-275):
//   q' = normalise(q + m * (q (x) (0, u0..2)))  ;  t' = t + m * R(q) u3..5      (t in 0.1 A units)
// ------------------------------------------------------------------------------------------------
__global__ void compose_update_kernel(int M, const float* __restrict__ upd, const float* __restrict__ dmask,
                                      float* __restrict__ quats, float* __restrict__ trans) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float q[4], R[9];
#pragma unroll
  for (int k = 0; k < 4; ++k) q[k] = quats[m * 4 + k];
  const float u[6] = {upd[m * 6 + 0], upd[m * 6 + 1], upd[m * 6 + 2], upd[m * 6 + 3], upd[m * 6 + 4], upd[m * 6 + 5]};
  const float mk = dmask[m];
  const float v[4] = {0.f, u[0], u[1], u[2]};
  float dq[4];
  quat_mul(q, v, dq);
  float nq[4];
  float n2 = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    nq[k] = q[k] + dq[k] * mk;
    n2 += nq[k] * nq[k];
  }
  const float inv = 1.f / sqrtf(n2);
  quat_to_rot(q, R);
#pragma unroll
  for (int k = 0; k < 4; ++k) quats[m * 4 + k] = nq[k] * inv;
#pragma unroll
  for (int r = 0; r < 3; ++r)
    trans[m * 3 + r] += (R[r * 3 + 0] * u[3] + R[r * 3 + 1] * u[4] + R[r * 3 + 2] * u[5]) * mk;
}

// ------------------------------------------------------------------------------------------------
// Row softmax with scale and key mask (sequence transformer; boolean key-padding semantics).
// S[rows, N]: row = ((b*H + h)*N + i).  One warp per row, in place.
// ------------------------------------------------------------------------------------------------
// Rows have stride ld (multiple of 4, >= N, <= 1024); the row lives in registers between the single read and the single write.
// Pimg != nullptr: the probabilities are written as the fp16 hi | lo operand image of the P.V GEMM (gemm_img.cuh) instead of in place:
// [row / N = (b, h)][i-tile][2 JB k-blocks of 64 keys][hi 16 KB | lo 16 KB][128 rows][128 B]; the low part carries gemm_img's 2^11 scale.
// Columns beyond ld are never written (the buffer is zeroed when the workspace is reserved).
__global__ void __launch_bounds__(256) softmax_rows_kernel(float* S, long long rows, int N, int ld, int rows_per_batch,
                                                           float scale, const float* __restrict__ keymask, uint8_t* Pimg = nullptr, int JB = 0) {
  // key mask of the CTA's sample, staged once in shared memory (16-byte reads, no bank conflicts): read per element from global
  // memory it cost four L1 wavefronts per instruction, four times the traffic of the logits themselves.  The 8 rows of a CTA share
  // one sample whenever rows_per_batch is a multiple of 8; otherwise (odd N) the rows read the mask from global memory.
  __shared__ __align__(16) float km_s[1024];
  const long long row0 = (long long)blockIdx.x * 8;
  const bool one_sample = (rows_per_batch & 7) == 0;
  if (one_sample) {
    const float* kmg = keymask + (row0 / rows_per_batch) * N;
    for (int j = threadIdx.x; j < ((N + 3) & ~3); j += blockDim.x) km_s[j] = j < N ? kmg[j] : 0.f;
  }
  __syncthreads();
  const long long row = row0 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const long long b = row / rows_per_batch;
  float4* s4 = reinterpret_cast<float4*>(S + row * ld);
  const float* km = keymask + b * N;
  const int nch = ld >> 2;  // 16-byte chunks per row
  float4 v[8];
  float mx = -INFINITY;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int c = lane + 32 * q;
    v[q] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    if (c < nch) {
      const float4 x = s4[c];
      const int j = 4 * c;
      float4 k4;
      if (one_sample) {
        k4 = *reinterpret_cast<const float4*>(km_s + j);  // entries >= N are 0
      } else {
        k4.x = j < N ? km[j] : 0.f;
        k4.y = j + 1 < N ? km[j + 1] : 0.f;
        k4.z = j + 2 < N ? km[j + 2] : 0.f;
        k4.w = j + 3 < N ? km[j + 3] : 0.f;
      }
      v[q].x = k4.x != 0.f ? x.x * scale : -INFINITY;
      v[q].y = k4.y != 0.f ? x.y * scale : -INFINITY;
      v[q].z = k4.z != 0.f ? x.z * scale : -INFINITY;
      v[q].w = k4.w != 0.f ? x.w * scale : -INFINITY;
      mx = fmaxf(mx, fmaxf(fmaxf(v[q].x, v[q].y), fmaxf(v[q].z, v[q].w)));
    }
  }
  mx = warp_max(mx);
  if (mx == -INFINITY) mx = 0.f;  // fully masked row: all probabilities 0
  float sum = 0.f;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    v[q].x = expf(v[q].x - mx);
    v[q].y = expf(v[q].y - mx);
    v[q].z = expf(v[q].z - mx);
    v[q].w = expf(v[q].w - mx);
    sum += (v[q].x + v[q].y) + (v[q].z + v[q].w);
  }
  sum = warp_sum(sum);
  const float inv = sum > 0.f ? 1.f / sum : 0.f;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int c = lane + 32 * q;
    if (c >= nch) continue;
    const float4 p = make_float4(v[q].x * inv, v[q].y * inv, v[q].z * inv, v[q].w * inv);
    if (!Pimg) {
      s4[c] = p;
      continue;
    }
    const long long bh = row / N;
    const int i = (int)(row - bh * N), j = 4 * c;
    const __half2 h01 = __floats2half2_rn(p.x, p.y), h23 = __floats2half2_rn(p.z, p.w);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn((p.x - f01.x) * 2048.f, (p.y - f01.y) * 2048.f);
    const __half2 l23 = __floats2half2_rn((p.z - f23.x) * 2048.f, (p.w - f23.y) * 2048.f);
    uint8_t* d = Pimg + (((size_t)bh * JB + (i >> 7)) * (2 * JB) + (j >> 6)) * (size_t)32768 + (i & 127) * 128 + (((((j & 63) >> 3)) ^ (i & 7)) << 4) +
                 (j & 7) * 2;
    *reinterpret_cast<uint2*>(d) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
    *reinterpret_cast<uint2*>(d + 16384) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
  }
}

// ------------------------------------------------------------------------------------------------
// Torsion head epilogue + psi blend (ipa_pytorch.py:355-361, score_network.py:258-260):
//   psi = u / sqrt(max(|u|^2, 1e-8)) ; psi = dm * psi + (1 - dm) * gt_psi,  dm = 1 - fixed_mask
// ------------------------------------------------------------------------------------------------
__global__ void psi_kernel(int M, const float* __restrict__ u, const float* __restrict__ fixed_mask,
                           const float* __restrict__ gt_psi, float* __restrict__ psi) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float a = u[m * 2], b = u[m * 2 + 1];
  const float den = sqrtf(fmaxf(a * a + b * b, 1e-8f));
  const float dm = 1.f - fixed_mask[m];
  psi[m * 2 + 0] = dm * (a / den) + (1.f - dm) * gt_psi[m * 2 + 0];
  psi[m * 2 + 1] = dm * (b / den) + (1.f - dm) * gt_psi[m * 2 + 1];
}

}  // namespace fdpt
