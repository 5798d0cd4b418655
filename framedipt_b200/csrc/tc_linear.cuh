// tc_linear.cuh — Y[M,N] = act(X[M,K] @ W[N,K]^T + bias) on the 5th-gen tensor cores (tcgen05, fp16 operands,
// fp32 accumulation in TMEM).  Building block / bring-up kernel for the pair-side fused kernels: exercises the
// packed-weight image + bulk-copy ring, the swizzled A image written from registers, UMMA descriptors, TMEM
// accumulator stages and the tcgen05.ld epilogue.
//   grid = ceil(M/128) CTAs, 192 threads: warps 0-3 workers (A staging + epilogue), warp 4 MMA issuer + TMEM owner,
//   warp 5 weight loader.  K multiple of 64 (<= 512), N multiple of 128.
#pragma once
#include "tc_common.cuh"

namespace fdpt {
namespace tc {

constexpr int LIN_STAGES = 4;
constexpr int LIN_STAGE_BYTES = 128 * 128;  // 128 weight rows x one k-block

struct TcLinearArgs {
  const float* X; int ldx; int M; int K;
  const __half* Wimg; int N;
  const float* bias; int relu;
  float* Y; int ldy;
};

__global__ void __launch_bounds__(192, 1) tc_linear_kernel(TcLinearArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [A image: K/64 x 16 KB][weight stages: LIN_STAGES x 16 KB][barriers]
  uint8_t* smem = smem_raw + smem_align_pad(smem_raw);  // offset arithmetic on the __shared__ symbol: accesses stay LDS / STS
  const int nkb = a.K / KB;
  uint8_t* Aimg = smem;
  uint8_t* Wst = Aimg + (size_t)nkb * 16384;
  uint64_t* bars = reinterpret_cast<uint64_t*>(Wst + LIN_STAGES * LIN_STAGE_BYTES);
  uint64_t* full = bars;                    // [LIN_STAGES]
  uint64_t* empty = bars + LIN_STAGES;      // [LIN_STAGES]
  uint64_t* acc_full = empty + LIN_STAGES;  // [2]
  uint64_t* acc_empty = acc_full + 2;       // [2]
  uint64_t* a_full = acc_empty + 2;         // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;
  const int nchunks = a.N / 128;

  if (threadIdx.x == 0) {
    for (int s = 0; s < LIN_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 128);
    }
    mbar_init(a_full, 128);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 5) {
    // ---------------- weight loader ----------------
    if (lane == 0) {
      int it = 0;
      for (int nc = 0; nc < nchunks; ++nc) {
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % LIN_STAGES;
          const uint32_t ph = (it / LIN_STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], LIN_STAGE_BYTES);
          const uint8_t* src = reinterpret_cast<const uint8_t*>(a.Wimg) + ((size_t)kb * a.N + (size_t)nc * 128) * 128;
          bulk_g2s(Wst + s * LIN_STAGE_BYTES, src, LIN_STAGE_BYTES, &full[s]);
        }
      }
    }
  } else if (warp == 4) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(128, 128);
      mbar_wait(a_full, 0);
      tc_fence_after();
      int it = 0;
      for (int nc = 0; nc < nchunks; ++nc) {
        const int as = nc & 1;
        const uint32_t aph = (nc >> 1) & 1;
        mbar_wait(&acc_empty[as], aph ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % LIN_STAGES;
          const uint32_t ph = (it / LIN_STAGES) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(Aimg + (size_t)kb * 16384);
          const uint32_t b_addr = smem_u32(Wst + s * LIN_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base + as * 128, make_sw128_desc(a_addr + k * 32), make_sw128_desc(b_addr + k * 32), idesc, (kb | k) != 0);
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[as]);
      }
    }
  } else {
    // ---------------- workers: stage A (fp32 -> fp16 swizzled image), then epilogues ----------------
    const int t = threadIdx.x;  // 0..127
    const int chunks_per_row = a.K / 8;
    for (int idx = t; idx < 128 * chunks_per_row; idx += 128) {
      const int r = idx / chunks_per_row, kc = idx % chunks_per_row;
      const int kb = kc / 8, c = kc % 8;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (m0 + r < a.M) {
        const float4* src = reinterpret_cast<const float4*>(a.X + (size_t)(m0 + r) * a.ldx + kc * 8);
        const float4 x0 = __ldg(src), x1 = __ldg(src + 1);
        v = make_uint4(pack_half2(x0.x, x0.y), pack_half2(x0.z, x0.w), pack_half2(x1.x, x1.y), pack_half2(x1.z, x1.w));
      }
      *reinterpret_cast<uint4*>(Aimg + (size_t)kb * 16384 + sw128_chunk_off(r, c)) = v;
    }
    fence_proxy_async();
    mbar_arrive(a_full);
    const int row = warp * 32 + lane;
    for (int nc = 0; nc < nchunks; ++nc) {
      const int as = nc & 1;
      const uint32_t aph = (nc >> 1) & 1;
      mbar_wait(&acc_full[as], aph);
      tc_fence_after();
      float v[4][32];
#pragma unroll
      for (int q = 0; q < 4; ++q) tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + as * 128 + q * 32, v[q]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&acc_empty[as]);
      if (m0 + row < a.M) {
        float* y = a.Y + (size_t)(m0 + row) * a.ldy + nc * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            float o[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float x = v[q][e + u];
              if (a.bias) x += a.bias[nc * 128 + q * 32 + e + u];
              if (a.relu) x = fmaxf(x, 0.f);
              o[u] = x;
            }
            *reinterpret_cast<float4*>(y + q * 32 + e) = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 256);
}

inline size_t tc_linear_smem_bytes(int K) { return 1024 + (size_t)(K / KB) * 16384 + LIN_STAGES * LIN_STAGE_BYTES + 256; }

}  // namespace tc
}  // namespace fdpt
