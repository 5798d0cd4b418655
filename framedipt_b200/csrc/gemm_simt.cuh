// gemm_simt.cuh — fp32 SIMT tiled GEMM with fused epilogues (node-side Linear layers, batched QK^T / PV).
//
// C[M,N] = epi( alpha * A[M,K] @ op(B) )     op(B) = B^T for weight layout B[N,K] (KMAJOR) or B[K,N]
// Epilogue order (matches how the reference composes Linear -> ReLU -> mask -> residual):
//   v = alpha*acc + bias[n];  v += U[row_i(r), n] + V[row_j(r), n]  (pair broadcast);  relu;  v *= rowmask[r];
//   v += residual[r, n];  if (accumulate) v += C[r, n]
// The node side of the network needs fp32-class accuracy (SURVEY §7 hard part 1), hence plain FFMA here.
#pragma once
#include "common.cuh"

namespace fdpt {

struct GemmArgs {
  const float* A = nullptr; int lda = 0; long long sA1 = 0, sA2 = 0;
  const float* B = nullptr; int ldb = 0; long long sB1 = 0, sB2 = 0;
  float* C = nullptr; int ldc = 0; long long sC1 = 0, sC2 = 0;
  int M = 0, N = 0, K = 0;
  int batch2 = 1;  // blockIdx.z = b1 * batch2 + b2
  float alpha = 1.f;
  const float* bias = nullptr; long long sBias1 = 0, sBias2 = 0;  // per-column bias, optionally per batch
  const float* residual = nullptr; int ldr = 0;
  const float* rowmask = nullptr;
  // pair-broadcast add: global pair row p = row0 + r -> b = p / (nres*nres), i = (p / nres) % nres, j = p % nres
  const float* U = nullptr; const float* V = nullptr; int lduv = 0; int nres = 0; long long row0 = 0;
  int relu = 0, accumulate = 0;
};

constexpr int GB_M = 64, GB_N = 64, GB_K = 16;

template <bool B_KMAJOR>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmArgs g) {
  __shared__ float As[GB_K][GB_M + 4];
  __shared__ float Bs[GB_K][GB_N + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, 4x4 outputs each
  const int m0 = blockIdx.x * GB_M, n0 = blockIdx.y * GB_N;
  const int b1 = blockIdx.z / g.batch2, b2 = blockIdx.z % g.batch2;
  const float* __restrict__ A = g.A + b1 * g.sA1 + b2 * g.sA2;
  const float* __restrict__ B = g.B + b1 * g.sB1 + b2 * g.sB2;
  float* C = g.C + b1 * g.sC1 + b2 * g.sC2;  // may alias residual

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // loader mapping: 256 threads load 64x16 elements of A (4 each): k = tid % 16, rows (tid / 16) + 16 r
  const int lk = tid & 15, lr = tid >> 4;
  // for B in [K,N] layout: n = tid % 64, k = tid / 64 + 4 r
  const int bn = tid & 63, bk = tid >> 6;

  for (int k0 = 0; k0 < g.K; k0 += GB_K) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int m = m0 + lr + 16 * r, k = k0 + lk;
      As[lk][lr + 16 * r] = (m < g.M && k < g.K) ? __ldg(A + (long long)m * g.lda + k) : 0.f;
    }
    if (B_KMAJOR) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int n = n0 + lr + 16 * r, k = k0 + lk;
        Bs[lk][lr + 16 * r] = (n < g.N && k < g.K) ? __ldg(B + (long long)n * g.ldb + k) : 0.f;
      }
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int n = n0 + bn, k = k0 + bk + 4 * r;
        Bs[bk + 4 * r][bn] = (n < g.N && k < g.K) ? __ldg(B + (long long)k * g.ldb + n) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GB_K; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
    const float* urow = nullptr;
    const float* vrow = nullptr;
    if (g.U) {
      const long long p = g.row0 + m;
      const long long nn = (long long)g.nres * g.nres;
      const long long b = p / nn;
      const int rem = (int)(p - b * nn);
      const int ii = rem / g.nres, jj = rem - ii * g.nres;
      urow = g.U + (b * g.nres + ii) * (long long)g.lduv;
      vrow = g.V + (b * g.nres + jj) * (long long)g.lduv;
    }
    const float rm = g.rowmask ? g.rowmask[m] : 1.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = g.alpha * acc[i][j];
      if (g.bias) v += g.bias[b1 * g.sBias1 + b2 * g.sBias2 + n];
      if (urow) v += urow[n] + vrow[n];
      if (g.relu) v = fmaxf(v, 0.f);
      if (g.rowmask) v *= rm;
      if (g.residual) v += g.residual[(long long)m * g.ldr + n];
      float* c = C + (long long)m * g.ldc + n;
      if (g.accumulate) v += *c;
      *c = v;
    }
  }
}

inline cudaError_t launch_gemm(const GemmArgs& g, bool b_kmajor, int batch, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return cudaSuccess;
  dim3 grid((g.M + GB_M - 1) / GB_M, (g.N + GB_N - 1) / GB_N, batch);
  if (b_kmajor)
    gemm_simt_kernel<true><<<grid, 256, 0, st>>>(g);
  else
    gemm_simt_kernel<false><<<grid, 256, 0, st>>>(g);
  return cudaGetLastError();
}

}  // namespace fdpt
