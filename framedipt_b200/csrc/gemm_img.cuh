// gemm_img.cuh — batched fp32-class GEMM whose BOTH operands are ready fp16 hi | lo operand images:  C[z] = A[z] . op(B[z]) (+ bias).
//
// The two batched attention GEMMs of the IPA (per sample and head: S = Q' K'^T with K = 280, O' = P [V | v_pts] with K = N_res,
// ipa_pytorch.py:245-324) used to run on gemm_tc.cuh, whose 256 producer threads load fp32 tiles and split them into hi | lo on the
// fly — instruction-issue bound, and every tile was converted by three CTAs.  Here the producers of those operands write the images
// once (ipa_prep_img_kernel: Q', K', V' with the frames applied; ipa_core_kernel's softmax: P), in exactly the layout lin_tc.cuh
// uses — [128-row tile][64-column k-block][hi 16 KB | lo 16 KB][128 rows][128 B, SWIZZLE_128B] — and this kernel is pure bulk copy +
// tcgen05.mma: a loader warp streams one 64 KB stage per k-block (A hi|lo + B hi|lo) through a 3-stage ring, one thread issues the
// 12 MMAs of the stage (2-term split: hi.hi into the main accumulator, the two cross terms into a second one, as in gemm_tc.cuh),
// 8 epilogue warps drain the double-buffered TMEM accumulators of tile t while the MMAs of tile t+1 run.  Persistent: one CTA per SM.
//
// B comes in two forms:
//   K-major  (rows = n):  image [n-tile][k-block][hi|lo]            -> S = Q' K'^T            (B = K' rows j)
//   MN-major (rows = k):  image [k-tile][64-col block][hi|lo]       -> O' = P V'              (B = V' rows j: the SAME row-image format,
//                          read through an MN-major descriptor: 8 k-rows x 128 B atoms, SBO = 1024 B, LBO = next 64-column block)
#pragma once
#include "lin_tc.cuh"

namespace fdpt {
namespace tc {

constexpr int GI_STAGES = 3;
constexpr int GI_STAGE_BYTES = 65536;  // A hi 16 KB | A lo 16 KB | B hi 16 KB | B lo 16 KB
constexpr int GI_WORKERS = 256;
constexpr int GI_THREADS = GI_WORKERS + 64;

struct GemmImgArgs {
  const uint8_t* A; long long sA; int nkbA;  // per batch: [m_tiles][nkbA][32 KB]
  const uint8_t* B; long long sB; int nkbB;  // K-major: [n_tiles][nkbB][32 KB];  MN-major: [k_tiles][nkbB = 64-column blocks][32 KB]
  int b_mn;                                  // 1: B is MN-major (rows = k)
  int nkb;                                   // k-blocks (64 K elements) to accumulate
  int M, N;                                  // valid rows / columns of C
  int m_tiles, n_tiles, batch, batch2;
  const float* bias; long long sBias;        // per batch [N], or nullptr
  float* C; int ldc; long long sC1, sC2;     // C + (z / batch2) * sC1 + (z % batch2) * sC2
};

__global__ void __launch_bounds__(GI_THREADS, 1) gemm_img_kernel(const __grid_constant__ GemmImgArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + smem_align_pad(smem_raw);  // offset arithmetic on the __shared__ symbol: accesses stay LDS / STS
  uint8_t* ring = smem;                                                     // GI_STAGES x 64 KB
  float* Stg = reinterpret_cast<float*>(ring + (size_t)GI_STAGES * GI_STAGE_BYTES);  // 8 warps x 32 rows x 20 floats
  uint64_t* bars = reinterpret_cast<uint64_t*>(Stg + 8 * 32 * 20);
  uint64_t* s_full = bars;                     // [GI_STAGES]
  uint64_t* s_empty = s_full + GI_STAGES;      // [GI_STAGES]
  uint64_t* acc_full = s_empty + GI_STAGES;    // [2]
  uint64_t* acc_empty = acc_full + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const long long tiles = (long long)a.m_tiles * a.n_tiles * a.batch;
  const int ncb_total = (a.N + 63) >> 6;  // MN-major: 64-column blocks that hold valid columns

  if (tid == 0) {
    for (int s = 0; s < GI_STAGES; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], GI_WORKERS);
    }
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();  // both operands were written by predecessor kernels

  auto decode = [&](long long t, int& mt, int& nt, int& z) {
    mt = (int)(t % a.m_tiles);
    const long long r = t / a.m_tiles;
    nt = (int)(r % a.n_tiles);
    z = (int)(r / a.n_tiles);
  };
  // width of n-tile nt: MN-major images carry 64-column blocks, the last tile may hold a single one
  auto tile_bn = [&](int nt) { return (a.b_mn && 2 * nt + 1 >= ncb_total) ? 64 : 128; };

  if (warp == 9) {
    // ============================ loader ============================
    if (elect_one()) {
      uint32_t it = 0;
      for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        int mt, nt, z;
        decode(t, mt, nt, z);
        const uint8_t* Ab = a.A + (long long)z * a.sA + (size_t)mt * a.nkbA * LT_STAGE_BYTES;
        const uint8_t* Bb = a.B + (long long)z * a.sB;
        const int ncb_t = tile_bn(nt) >> 6;
        for (int kb = 0; kb < a.nkb; ++kb, ++it) {
          const uint32_t s = it % GI_STAGES;
          mbar_wait(&s_empty[s], ((it / GI_STAGES) & 1) ^ 1);
          uint8_t* st = ring + (size_t)s * GI_STAGE_BYTES;
          if (!a.b_mn) {
            mbar_arrive_expect_tx(&s_full[s], 2 * LT_STAGE_BYTES);
            bulk_g2s(st, Ab + (size_t)kb * LT_STAGE_BYTES, LT_STAGE_BYTES, &s_full[s]);
            bulk_g2s(st + LT_STAGE_BYTES, Bb + ((size_t)nt * a.nkbB + kb) * LT_STAGE_BYTES, LT_STAGE_BYTES, &s_full[s]);
          } else {
            mbar_arrive_expect_tx(&s_full[s], LT_STAGE_BYTES + ncb_t * 2 * 8192);
            bulk_g2s(st, Ab + (size_t)kb * LT_STAGE_BYTES, LT_STAGE_BYTES, &s_full[s]);
            const int kt = kb >> 1, half = kb & 1;  // 64 k-rows = half of a 128-row image tile
            for (int cbi = 0; cbi < ncb_t; ++cbi)
              for (int hl = 0; hl < 2; ++hl)
                bulk_g2s(st + LT_STAGE_BYTES + hl * 16384 + cbi * 8192,
                         Bb + ((size_t)kt * a.nkbB + (2 * nt + cbi)) * LT_STAGE_BYTES + hl * 16384 + half * 8192, 8192, &s_full[s]);
          }
        }
      }
    }
  } else if (warp == 8) {
    // ============================ MMA issuer ============================
    if (elect_one()) {
      uint32_t it = 0, tl = 0;
      for (long long t = blockIdx.x; t < tiles; t += gridDim.x, ++tl) {
        int mt, nt, z;
        decode(t, mt, nt, z);
        const int bn = tile_bn(nt);
        const uint32_t idesc = make_idesc_f16(128, bn) | (a.b_mn ? (1u << 16) : 0u);
        const int as = tl & 1;
        mbar_wait(&acc_empty[as], ((tl >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc_main = tmem_base + as * 256, acc_x = acc_main + 128;
        for (int kb = 0; kb < a.nkb; ++kb, ++it) {
          const uint32_t s = it % GI_STAGES;
          mbar_wait(&s_full[s], (it / GI_STAGES) & 1);
          tc_fence_after();
          const uint32_t ah = smem_u32(ring + (size_t)s * GI_STAGE_BYTES), al = ah + 16384, bh = ah + 32768, bl = bh + 16384;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t dah = make_sw128_desc(ah + k * 32), dal = make_sw128_desc(al + k * 32);
            uint64_t dbh, dbl;
            if (!a.b_mn) {
              dbh = make_sw128_desc(bh + k * 32);
              dbl = make_sw128_desc(bl + k * 32);
            } else {  // 16 k-rows per step = two 8-row atoms of 1024 B; the next 64-column block lies 8 KB further
              dbh = make_sw128_desc_ls(bh + k * 2048, 8192, 1024);
              dbl = make_sw128_desc_ls(bl + k * 2048, 8192, 1024);
            }
            const uint32_t first = (kb == 0 && k == 0) ? 0u : 1u;
            umma_f16(acc_x, dal, dbh, idesc, first);
            umma_f16(acc_x, dah, dbl, idesc, 1u);
            umma_f16(acc_main, dah, dbh, idesc, first);
          }
          umma_commit(&s_empty[s]);
        }
        umma_commit(&acc_full[as]);
      }
    }
  } else {
    // ============================ epilogue workers ============================
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int c_half = (warp >> 2) * 64;
    float* stg = Stg + warp * 32 * 20;
    const bool c_vec = ((reinterpret_cast<uintptr_t>(a.C) | (uintptr_t)(a.ldc * 4) | (uintptr_t)(a.sC1 * 4) | (uintptr_t)(a.sC2 * 4)) & 15) == 0;  // (N need not be a multiple of 4: the straddling float4 is stored by element; gemm_img has no residual / accumulate)
    uint32_t tl = 0;
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x, ++tl) {
      int mt, nt, z;
      decode(t, mt, nt, z);
      const int as = tl & 1;
      const int m0 = mt * 128, n0 = nt * 128;
      EpiArgs ep;
      ep.M = a.M; ep.N = a.N; ep.alpha = 1.f; ep.bias = a.bias ? a.bias + (long long)z * a.sBias : nullptr; ep.relu = 0; ep.rowmask = nullptr;
      ep.residual = nullptr; ep.ldr = 0; ep.accumulate = 0;
      ep.Y = a.C + (long long)(z / a.batch2) * a.sC1 + (long long)(z % a.batch2) * a.sC2;
      ep.ldy = a.ldc;
      const int mw = m0 + (warp & 3) * 32;
      mbar_wait(&acc_full[as], (tl >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int cb = c_half + q * 32;
        float v[32], x2[32];
        tmem_ld32(tmem_base + lane_base + as * 256 + cb, v);
        tmem_ld32(tmem_base + lane_base + as * 256 + 128 + cb, x2);
        tmem_ld_wait();
        if (q == 1) {
          tc_fence_before();
          mbar_arrive(&acc_empty[as]);
        }
        if (n0 + cb >= a.N) continue;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaf(x2[j], 1.0f / GT_LO_SCALE, v[j]);
        if (c_vec) {  // 16-byte accesses through a [32][20] patch, two 16-column passes
          store_transposed_v4<16>(ep, v, stg, lane, mw, n0 + cb);
          store_transposed_v4<16>(ep, v + 16, stg, lane, mw, n0 + cb + 16);
        } else {
          store_transposed<16>(ep, v, stg, lane, mw, n0 + cb);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 512);
}

inline size_t gemm_img_smem_bytes() { return 1024 + (size_t)GI_STAGES * GI_STAGE_BYTES + (size_t)8 * 32 * 20 * 4 + (2 * GI_STAGES + 4) * 8 + 64; }

}  // namespace tc
}  // namespace fdpt
