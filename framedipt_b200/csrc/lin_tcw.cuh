// lin_tcw.cuh — weight-resident variant of lin_tc.cuh for CTAs that own ONE 128-column n-tile (every node-side Linear with N <= 384).
//
// lin_tc keeps the activation tile resident (up to 160 KB) and streams the weights past it through what is left of shared memory:
// with K = 320 that is three 16 KB units, i.e. 1.5 k-blocks in flight, and the clock64 timeline shows the MMA issuer waiting on
// L2 latency for every k-block (8 k cycles for the 60 MMAs of a tile instead of 3.8 k).  Weights are static, and under programmatic
// dependent launch the CTA starts while its predecessor is still running — so here the loader warp fetches the WHOLE weight panel of
// the CTA's n-tile (K/64 x 32 KB, hi | lo) before the predecessor has even finished, and what streams after `griddepcontrol.wait` is the
// activation: 32 KB k-blocks through a two-slot ring (one bulk copy each when the producer handed over an operand image, else staged
// and split by the workers).  The first MMA issues one L2 round trip after the predecessor's last store.  Same arithmetic, same
// epilogues as lin_tc (fp32 rows through the transposition patch, and/or the next layer's operand image); the patches reuse the
// activation ring, which is idle once the accumulator is complete.
#pragma once
#include "lin_tc.cuh"

namespace fdpt {
namespace tc {

template <bool XIMG, bool YIMG>
__global__ void __launch_bounds__(LT_THREADS, 1) lin_tcw_kernel(LinTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + smem_align_pad(smem_raw);  // offset arithmetic on the __shared__ symbol: accesses stay LDS / STS
  uint8_t* Aring = smem;                                          // 2 x [hi 16 KB | lo 16 KB]; later the epilogue's transposition patches
  uint8_t* Wst = Aring + 2 * (size_t)LT_STAGE_BYTES;              // nkb x [hi 16 KB | lo 16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(Wst + (size_t)a.nkb * LT_STAGE_BYTES);
  uint64_t* w_full = bars;                 // [LT_MAX_KB]
  uint64_t* a_full = bars + LT_MAX_KB;     // [2]
  uint64_t* a_empty = a_full + 2;          // [2]
  uint64_t* acc_full = a_empty + 2;        // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int m0 = blockIdx.x * 128;
  const int nt = blockIdx.y;  // this CTA's n-tile
  if (tid == 0) LT_TS(0);

  if (tid == 0) {
    for (int k = 0; k < LT_MAX_KB; ++k) mbar_init(&w_full[k], 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_full[s], XIMG ? 1 : LT_WORKERS);
      mbar_init(&a_empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  if (tid == 0) LT_TS(1);

  if (warp == 9) {
    // ============================ weight loader: the whole panel, before the predecessor kernel has finished ============================
    if (elect_one()) {
      for (int kb = 0; kb < a.nkb; ++kb) {
        mbar_arrive_expect_tx(&w_full[kb], LT_STAGE_BYTES);
        bulk_g2s(Wst + (size_t)kb * LT_STAGE_BYTES, reinterpret_cast<const uint8_t*>(a.Wimg) + ((size_t)nt * a.nkb + kb) * LT_STAGE_BYTES,
                 LT_STAGE_BYTES, &w_full[kb]);
      }
    }
  } else if (warp == 8) {
    // ============================ MMA issuer ============================
    if (elect_one()) {
      const uint32_t idesc = make_idesc_f16(128, 128);
      const uint32_t acc_main = tmem_base, acc_x = tmem_base + 128;
      for (int kb = 0; kb < a.nkb; ++kb) {
        const int s = kb & 1;
        mbar_wait(&w_full[kb], 0);
        if (kb == 0) LT_TS(7);
        mbar_wait(&a_full[s], (kb >> 1) & 1);
        if (kb == 0) LT_TS(6);
        tc_fence_after();
        const uint32_t ah = smem_u32(Aring + (size_t)s * LT_STAGE_BYTES), al = ah + 16384;
        const uint32_t bh = smem_u32(Wst + (size_t)kb * LT_STAGE_BYTES), bl = bh + 16384;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t dah = make_sw128_desc(ah + k * 32), dal = make_sw128_desc(al + k * 32);
          const uint64_t dbh = make_sw128_desc(bh + k * 32), dbl = make_sw128_desc(bl + k * 32);
          const uint32_t first = (kb == 0 && k == 0) ? 0u : 1u;
          umma_f16(acc_x, dal, dbh, idesc, first);
          umma_f16(acc_x, dah, dbl, idesc, 1u);
          umma_f16(acc_main, dah, dbh, idesc, first);
        }
        umma_commit(&a_empty[s]);
      }
      umma_commit(acc_full);
      LT_TS(8);
    }
  } else {
    // ============================ workers: feed the activation ring, then the epilogue ============================
    pdl_wait();  // X (and residual / Y) belong to the predecessor kernel until it has completed
    if (tid == 0) LT_TS(2);
    if constexpr (XIMG) {
      if (tid == 0) {
        const uint8_t* src = reinterpret_cast<const uint8_t*>(a.X) + (size_t)blockIdx.x * a.nkb * LT_STAGE_BYTES;
        for (int kb = 0; kb < a.nkb; ++kb) {
          const int s = kb & 1;
          if (kb >= 2) mbar_wait(&a_empty[s], ((kb - 2) >> 1) & 1);
          mbar_arrive_expect_tx(&a_full[s], LT_STAGE_BYTES);
          bulk_g2s(Aring + (size_t)s * LT_STAGE_BYTES, src + (size_t)kb * LT_STAGE_BYTES, LT_STAGE_BYTES, &a_full[s]);
        }
      }
    } else {
      ChunkPlan pa;
      const int r0 = tid >> 3, c = tid & 7;
      pa.src = a.X + (long long)(m0 + r0) * a.ldx + 8 * c;
      pa.it_stride = 32LL * a.ldx; pa.kb_stride = GT_KB;
      pa.dst = sw128_chunk_off(r0, c); pa.dst_it_stride = 32 * 128; pa.lo_off = 16384;
      pa.iters = 4; pa.kmajor = 1;
      pa.row0 = m0 + r0; pa.row_step = 32; pa.row_lim = a.M; pa.col0 = 8 * c; pa.col_lim = a.K; pa.vec = a.x_vec;
      RegTile ta[2];
      load_tile(pa, 0, ta[0]);
      if (a.nkb > 1) load_tile(pa, 1, ta[1]);
#pragma unroll
      for (int kb = 0; kb < LT_MAX_KB; ++kb) {  // unrolled: the register tiles are indexed at compile time
        if (kb < a.nkb) {
          const int s = kb & 1;
          if (kb >= 2) mbar_wait(&a_empty[s], ((kb - 2) >> 1) & 1);
          store_tile(pa, Aring + (size_t)s * LT_STAGE_BYTES, ta[s]);
          fence_proxy_async();
          mbar_arrive(&a_full[s]);
          if (tid == 0 && kb == 0) LT_TS(4);
          if (kb + 2 < a.nkb) load_tile(pa, kb + 2, ta[s]);
        }
      }
    }
    if (tid == 0) LT_TS(5);
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int c_half = (warp >> 2) * 64;
    const int mw = m0 + (warp & 3) * 32;
    const int n0 = nt * 128;
    EpiArgs ep;
    ep.M = a.M; ep.N = a.N; ep.alpha = 1.f; ep.bias = a.bias; ep.relu = a.relu; ep.rowmask = a.rowmask; ep.residual = a.residual;
    ep.ldr = a.ldr; ep.accumulate = 0; ep.Y = a.Y; ep.ldy = a.ldy;
    mbar_wait(acc_full, 0);  // every MMA has completed: the activation ring is idle and becomes the transposition patches
    tc_fence_after();
    if (tid == 0) LT_TS(9);
    // private transposition patch of the warp (the vectorised path needs 32 x 36 floats, the scalar fallback 32 x 17)
    float* stg = reinterpret_cast<float*>(Aring) + warp * ST4_PATCH_FLOATS;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int cb = c_half + q * 32;
      float v[32], x2[32];
      tmem_ld32(tmem_base + lane_base + cb, v);
      tmem_ld32(tmem_base + lane_base + 128 + cb, x2);
      tmem_ld_wait();
      if (tid == 0 && q == 1) LT_TS(10);
      if (n0 + cb >= a.N) continue;
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaf(x2[j], 1.0f / GT_LO_SCALE, v[j]);
      if constexpr (YIMG) {
        const int r = (warp & 3) * 32 + lane;
        if (m0 + r < a.M) {
          const float rm = a.rowmask ? __ldg(a.rowmask + m0 + r) : 1.f;
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const int col = n0 + cb + 8 * cc;
            if (col >= a.N) break;
            float y[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float t = v[8 * cc + e] + ((a.bias && col + e < a.N) ? __ldg(a.bias + col + e) : 0.f);
              if (a.relu) t = fmaxf(t, 0.f);
              t *= rm;
              if (a.residual && col + e < a.N) t += a.residual[(long long)(m0 + r) * a.ldr + col + e];
              y[e] = (col + e < a.N) ? t : 0.f;
            }
            uint4 hi, lo;
            split8(make_float4(y[0], y[1], y[2], y[3]), make_float4(y[4], y[5], y[6], y[7]), hi, lo);
            uint8_t* dst = reinterpret_cast<uint8_t*>(a.dbg) + ((size_t)blockIdx.x * ((a.N + 63) >> 6) + (col >> 6)) * LT_STAGE_BYTES +
                           sw128_chunk_off(r, (col & 63) >> 3);
            *reinterpret_cast<uint4*>(dst) = hi;
            *reinterpret_cast<uint4*>(dst + 16384) = lo;
          }
        }
        if (!a.Y) continue;
      }
      if (a.y_vec) store_transposed_v4<32>(ep, v, stg, lane, mw, n0 + cb);
      else store_transposed<16>(ep, v, stg, lane, mw, n0 + cb);
    }
    if (tid == 0) LT_TS(11);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 256);
  if (tid == 0) LT_TS(12);
}

inline size_t lin_tcw_smem_bytes(int nkb) { return 1024 + (size_t)(2 + nkb) * LT_STAGE_BYTES + (LT_MAX_KB + 5) * 8 + 64; }

}  // namespace tc
}  // namespace fdpt
