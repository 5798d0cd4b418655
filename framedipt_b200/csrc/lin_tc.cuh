// lin_tc.cuh — fp32-class Linear layer  Y[M,N] = epi(X[M,K] @ W[N,K]^T)  for the node side, K <= 320, weights pre-packed.
//
// Same arithmetic as gemm_tc.cuh (2-term split fp16 with scaled low part, two TMEM accumulators; fp32-class accuracy), but organised
// around what bounds the node-side GEMMs on B200: L2 -> SM operand traffic and per-CTA fixed latency, not tensor throughput.
//   * weights are split ONCE (fdpt_finalize_params) into fp16 hi | lo operand images, laid out [n-tile][k-block][hi|lo][128 rows][128 B]
//     so that one 32 KB bulk copy (TMA unit, no thread work, deep prefetch) lands a ready-to-multiply stage;
//   * a CTA owns one 128-row tile of X for a contiguous range of n-tiles: X is loaded and split once into a resident operand image
//     (K/64 k-blocks x 32 KB) and every weight stage streams past it ("A-stationary"), instead of re-loading and re-splitting X for
//     every 64- or 128-column tile;
//   * two accumulator pairs in TMEM (2 x (main 128 + cross 128) = 512 columns): the epilogue of n-tile t overlaps the MMAs of t+1;
//   * the epilogue transposes each warp's 32 x 32 (or 32 x 16) accumulator block through a private padded shared-memory patch so that
//     global stores are row-contiguous 128-byte (64-byte) segments: with the TMEM-native "one thread = one row" mapping every store
//     instruction touched 32 different lines and the stores alone took 42 of 75 us on the IPA projection GEMM.
// Warps 0-7: X staging, then epilogues (warps w and w+4 share TMEM lanes, 64 columns each); warp 8: MMA issuer + TMEM owner;
// warp 9: weight loader.
#pragma once
#include "gemm_tc.cuh"

namespace fdpt {
namespace tc {

constexpr int LT_STAGE_BYTES = 32768;   // one (n-tile, k-block): hi 16 KB | lo 16 KB
constexpr int LT_UNIT_BYTES = 16384;    // ring granularity: the hi and the lo image of a k-block are separate ring units
constexpr int LT_WORKERS = 256;
constexpr int LT_THREADS = LT_WORKERS + 64;
constexpr int LT_MAX_KB = 5;            // K <= 320

// Epilogue of the fused IPA projection (kernels_ipa.cuh): the GEMM's output row of residue m is turned straight into the operand images
// of the attention GEMMs (gemm_img.cuh) — no fp32 projection buffer, no separate prep kernel.  Output columns are laid out as 24
// blocks of 320 = [operand (q, k, v)][head][256 scalar channels | points planar x | y | z (24 or 36) | zero padding]; a thread of the
// epilogue owns one residue (TMEM lane) and 64 consecutive columns, i.e. either 64 scalar channels or the whole point group of one
// (operand, head), to which it applies the residue's frame (ipa_pytorch.py:214-239, rigid_utils.py:82-106) in registers.
struct IpaProjEpi {
  const float* quats; const float* trans;   // [M,4], [M,3] frames (translations in 0.1 A units)
  const float* head_w; const float* mask;   // [H] head weights (softplus -> gamma), [M] residue mask
  float* kbias;                             // [B,H,N]  -gamma/2 |k_pts|^2 + 1e5 (m - 1)
  uint8_t *Qimg, *Kimg, *Vimg;              // [B*H][JB][5 k-blocks][hi|lo][128 rows][128 B]
  int n_res, JB;
};

struct LinTcArgs {
  const float* X; int ldx; int M, K, N;
  const __half* Wimg;          // [n_tiles][nkb][2][128][64] fp16 (swizzled rows)
  int nkb, n_tiles, tiles_per_cta, units;   // units: 16 KB ring slots (<= 8)
  int stg_cols;                             // epilogue staging width per warp: 32 or 16 columns
  const float* bias; int relu;
  const float* rowmask;
  const float* residual; int ldr;
  float* Y; int ldy;
  int x_vec, y_vec;
  int dbg_flags;               // bring-up: bit 0 skip the epilogue stores, bit 1 skip the MMAs
  long long* dbg;              // optional clock64 timeline of CTA (0,0): 16 stamps (tools/lin_timeline.py), or nullptr
  IpaProjEpi ipa;              // only read by the EPI_IPA instantiation
  // Operand-image chaining between consecutive Linear layers (both optional):
  //   XIMG: X points to the split operand image, i.e. X is already available as the split operand image [m-tile][k-block][hi|lo][128 rows][128 B] written by the previous
  //         layer's epilogue -> the staging phase (fp32 loads + split + swizzled stores by 256 threads) becomes nkb bulk copies;
  //   Yimg: the epilogue writes the result in that same layout for the next layer (k-block = 64 output columns; rows >= M and
  //         columns >= N are never written and stay zero from the allocation-time memset).  Y may then be null.
  // (no extra fields: the input image travels in X when the kernel is instantiated with XIMG, the output image in `dbg` when it is
  //  instantiated with YIMG -- a longer parameter block costs the plain instantiation 16 bytes of spills)
};

#define LT_TS(id)                                                                  \
  do {                                                                             \
    if (!YIMG && a.dbg && blockIdx.x == 0 && blockIdx.y == 0) a.dbg[(id)] = clock64(); \
  } while (0)

// YIMG: the epilogue also (or only) writes the operand image of the next layer; a separate instantiation so that the plain kernel keeps
// its register allocation
// EPI: 0 plain epilogue; 1 fused IPA projection (IpaProjEpi above); 2 in_proj of a sequence-transformer layer: the 960 output columns
// [q | k | v] x [4 heads x 80] are written as the per-(sample, head) Q / K / V operand images [b*4 + h][JB][2 k-blocks][32 KB] of the
// attention GEMMs (gemm_img.cuh); 80 = 10 chunks of 8 columns, so every 16-byte chunk belongs to exactly one head
template <bool XIMG, bool YIMG, int EPI = 0>
__global__ void __launch_bounds__(LT_THREADS, 1) lin_tc_kernel(LinTcArgs a) {
  constexpr bool EPI_IPA = EPI == 1;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + smem_align_pad(smem_raw);  // offset arithmetic on the __shared__ symbol: accesses stay LDS / STS
  uint8_t* Aimg = smem;                                        // nkb x [hi 16 KB | lo 16 KB]
  uint8_t* Wst = Aimg + (size_t)a.nkb * LT_STAGE_BYTES;        // units x 16 KB
  float* Stg = reinterpret_cast<float*>(Wst + (size_t)a.units * LT_UNIT_BYTES);  // 8 warps x 32 rows x (stg_cols + 1) floats
  uint64_t* bars = reinterpret_cast<uint64_t*>(Stg + 8 * 32 * (a.stg_cols + 1));
  uint64_t* w_full = bars;                 // [8]
  uint64_t* w_empty = bars + 8;            // [8]
  uint64_t* a_full = bars + 16;            // [LT_MAX_KB]
  uint64_t* acc_full = a_full + LT_MAX_KB; // [2]
  uint64_t* acc_empty = acc_full + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int m0 = blockIdx.x * 128;
  const int nt_begin = blockIdx.y * a.tiles_per_cta;
  const int nt_end = min(a.n_tiles, nt_begin + a.tiles_per_cta);
  const int ntiles = nt_end - nt_begin;
  if (tid == 0) LT_TS(0);

  if (tid == 0) {
    for (int s = 0; s < 8; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    for (int k = 0; k < LT_MAX_KB; ++k) mbar_init(&a_full[k], XIMG ? 1 : LT_WORKERS);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], LT_WORKERS);
    }
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  if (tid == 0) LT_TS(1);

  if (warp == 9) {
    // ============================ weight loader (weights are static: no need to wait for the predecessor kernel) ============================
    if (elect_one()) {
      int it = 0;  // ring unit counter: 2 per k-block (hi, lo)
      for (int nt = nt_begin; nt < nt_end; ++nt)
        for (int kb = 0; kb < a.nkb; ++kb)
          for (int hl = 0; hl < 2; ++hl, ++it) {
            const int s = it % a.units;
            mbar_wait(&w_empty[s], ((it / a.units) & 1) ^ 1);
            mbar_arrive_expect_tx(&w_full[s], LT_UNIT_BYTES);
            bulk_g2s(Wst + (size_t)s * LT_UNIT_BYTES,
                     reinterpret_cast<const uint8_t*>(a.Wimg) + ((size_t)nt * a.nkb + kb) * LT_STAGE_BYTES + hl * LT_UNIT_BYTES, LT_UNIT_BYTES,
                     &w_full[s]);
          }
    }
  } else if (warp == 8) {
    // ============================ MMA issuer ============================
    if (elect_one()) {
      const uint32_t idesc = make_idesc_f16(128, 128);
      int it = 0;
      for (int t = 0; t < ntiles; ++t) {
        const int as = t & 1;
        mbar_wait(&acc_empty[as], ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc_main = tmem_base + as * 256, acc_x = acc_main + 128;
        for (int kb = 0; kb < a.nkb; ++kb, it += 2) {
          if (t == 0) mbar_wait(&a_full[kb], 0);
          if (t == 0 && kb == 0) LT_TS(6);
          const int sh = it % a.units, sl = (it + 1) % a.units;
          mbar_wait(&w_full[sh], (it / a.units) & 1);
          mbar_wait(&w_full[sl], ((it + 1) / a.units) & 1);
          tc_fence_after();
          if (t == 0 && kb == 0) LT_TS(7);
          const uint32_t ah = smem_u32(Aimg + (size_t)kb * LT_STAGE_BYTES), al = ah + 16384;
          const uint32_t bh = smem_u32(Wst + (size_t)sh * LT_UNIT_BYTES), bl = smem_u32(Wst + (size_t)sl * LT_UNIT_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t dah = make_sw128_desc(ah + k * 32), dal = make_sw128_desc(al + k * 32);
            const uint64_t dbh = make_sw128_desc(bh + k * 32), dbl = make_sw128_desc(bl + k * 32);
            const uint32_t first = (kb == 0 && k == 0) ? 0u : 1u;
            if (a.dbg_flags & 2) continue;
            umma_f16(acc_x, dal, dbh, idesc, first);
            umma_f16(acc_x, dah, dbl, idesc, 1u);
            umma_f16(acc_main, dah, dbh, idesc, first);
          }
          umma_commit(&w_empty[sh]);
          umma_commit(&w_empty[sl]);
        }
        umma_commit(&acc_full[as]);
        if (t == 0) LT_TS(8);
      }
    }
  } else {
    // ============================ workers: stage X once, then epilogues ============================
    pdl_wait();  // X (and residual / Y) belong to the predecessor kernel until it has completed
    if (tid == 0) LT_TS(2);
    if constexpr (XIMG) {
      if (tid == 0) {
        const uint8_t* src = reinterpret_cast<const uint8_t*>(a.X) + (size_t)blockIdx.x * a.nkb * LT_STAGE_BYTES;
        for (int kb = 0; kb < a.nkb; ++kb) {
          mbar_arrive_expect_tx(&a_full[kb], LT_STAGE_BYTES);
          bulk_g2s(Aimg + (size_t)kb * LT_STAGE_BYTES, src + (size_t)kb * LT_STAGE_BYTES, LT_STAGE_BYTES, &a_full[kb]);
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
      if (tid == 0) LT_TS(3);
      for (int kb = 0; kb < a.nkb; kb += 2) {
        if (kb + 1 < a.nkb) load_tile(pa, kb + 1, ta[1]);
        store_tile(pa, Aimg + (size_t)kb * LT_STAGE_BYTES, ta[0]);
        fence_proxy_async();
        mbar_arrive(&a_full[kb]);
        if (tid == 0 && kb == 0) LT_TS(4);
        if (kb + 1 < a.nkb) {
          if (kb + 2 < a.nkb) load_tile(pa, kb + 2, ta[0]);
          store_tile(pa, Aimg + (size_t)(kb + 1) * LT_STAGE_BYTES, ta[1]);
          fence_proxy_async();
          mbar_arrive(&a_full[kb + 1]);
        }
      }
    }
    if (tid == 0) LT_TS(5);
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int c_half = (warp >> 2) * 64;
    const int sw = a.stg_cols, sld = sw + 1;           // staging width / padded row stride
    float* stg = Stg + warp * 32 * sld;                // this warp's private patch
    const int mw = m0 + (warp & 3) * 32;               // first row of this warp's block
    EpiArgs ep;
    ep.M = a.M; ep.N = a.N; ep.alpha = 1.f; ep.bias = a.bias; ep.relu = a.relu; ep.rowmask = a.rowmask; ep.residual = a.residual;
    ep.ldr = a.ldr; ep.accumulate = 0; ep.Y = a.Y; ep.ldy = a.ldy;
    for (int t = 0; t < ntiles; ++t) {
      const int as = t & 1;
      const int n0 = (nt_begin + t) * 128;
      mbar_wait(&acc_full[as], (t >> 1) & 1);
      tc_fence_after();
      if (tid == 0 && t == 0) LT_TS(9);
      float pt[EPI_IPA ? 32 : 1];  // EPI_IPA: first half of a point group (columns 256..287 of a block), kept for the second pass
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int cb = c_half + q * 32;
        float v[32], x2[32];
        tmem_ld32(tmem_base + lane_base + as * 256 + cb, v);
        tmem_ld32(tmem_base + lane_base + as * 256 + 128 + cb, x2);
        tmem_ld_wait();
        if (q == 1) {  // both halves of this thread's columns are in registers: the accumulator pair can be overwritten
          tc_fence_before();
          mbar_arrive(&acc_empty[as]);
          if (tid == 0 && t == 0) LT_TS(10);
        }
        if (n0 + cb >= a.N || (a.dbg_flags & 1)) continue;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaf(x2[j], 1.0f / GT_LO_SCALE, v[j]);
        if constexpr (EPI == 2) {
          const int r = (warp & 3) * 32 + lane;
          const int m = m0 + r;
          const int col0 = n0 + cb;
          if (m < a.M) {
            const int bsamp = m / a.ipa.n_res, i = m - bsamp * a.ipa.n_res;
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
              const int cg = (col0 >> 3) + cc;          // 8-column chunk index, 0..119
              if (cg >= 120) break;
              const int kind = cg / 40, within = cg - kind * 40, h = within / 10, c = within - h * 10;
              float y[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) y[e] = v[8 * cc + e] + __ldg(a.bias + col0 + 8 * cc + e);
              uint4 hi, lo;
              split8(make_float4(y[0], y[1], y[2], y[3]), make_float4(y[4], y[5], y[6], y[7]), hi, lo);
              uint8_t* img = kind == 0 ? a.ipa.Qimg : (kind == 1 ? a.ipa.Kimg : a.ipa.Vimg);
              uint8_t* d = img + (((size_t)(bsamp * 4 + h) * a.ipa.JB + (i >> 7)) * 2 + (c >> 3)) * (size_t)LT_STAGE_BYTES + sw128_chunk_off(i & 127, c & 7);
              *reinterpret_cast<uint4*>(d) = hi;
              *reinterpret_cast<uint4*>(d + 16384) = lo;
            }
          }
          continue;
        }
        if constexpr (EPI_IPA) {
          const int r = (warp & 3) * 32 + lane;
          const int m = m0 + r;
          const int col0 = n0 + cb;                  // multiple of 32
          const int blk = col0 / 320, within = col0 - blk * 320;
          const int kind = blk >> 3, h = blk & 7;    // blocks ordered [operand][head]
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += __ldg(a.bias + col0 + j);
          if (m < a.M) {
            const int bsamp = m / a.ipa.n_res, i = m - bsamp * a.ipa.n_res;
            uint8_t* img = kind == 0 ? a.ipa.Qimg : (kind == 1 ? a.ipa.Kimg : a.ipa.Vimg);
            uint8_t* tile = img + ((size_t)(bsamp * 8 + h) * a.ipa.JB + (i >> 7)) * (5 * (size_t)LT_STAGE_BYTES);
            auto put8 = [&](int c, const float* y) {  // chunk c (8 columns) of this (operand, head) block, row i
              uint4 hi, lo;
              split8(make_float4(y[0], y[1], y[2], y[3]), make_float4(y[4], y[5], y[6], y[7]), hi, lo);
              uint8_t* d = tile + (size_t)(c >> 3) * LT_STAGE_BYTES + sw128_chunk_off(i & 127, c & 7);
              *reinterpret_cast<uint4*>(d) = hi;
              *reinterpret_cast<uint4*>(d + 16384) = lo;
            };
            if (within < 256) {
#pragma unroll
              for (int cc = 0; cc < 4; ++cc) put8((within >> 3) + cc, v + 8 * cc);
            } else if (within == 256) {
#pragma unroll
              for (int j = 0; j < 32; ++j) pt[j] = v[j];
            } else {  // within == 288: the point group is complete (pt = columns 256..287, v = 288..319)
              float R[9];
              {
                const float qf[4] = {a.ipa.quats[m * 4], a.ipa.quats[m * 4 + 1], a.ipa.quats[m * 4 + 2], a.ipa.quats[m * 4 + 3]};
                quat_to_rot(qf, R);
              }
              const float t0 = a.ipa.trans[m * 3], t1 = a.ipa.trans[m * 3 + 1], t2 = a.ipa.trans[m * 3 + 2];
              const float gamma = log1pf(expf(__ldg(a.ipa.head_w + h))) * sqrtf(1.0f / 108.f);  // softplus(w_h) * sqrt(1 / (3 * 8 * 9 / 2))
              if (kind < 2) {  // 8 points: x = pt[0:8], y = pt[8:16], z = pt[16:24]
                float gx[8], gy[8], gz[8];
                float d2 = 0.f;
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                  const float x = pt[p], y = pt[8 + p], z = pt[16 + p];
                  gx[p] = R[0] * x + R[1] * y + R[2] * z + t0;
                  gy[p] = R[3] * x + R[4] * y + R[5] * z + t1;
                  gz[p] = R[6] * x + R[7] * y + R[8] * z + t2;
                  d2 += gx[p] * gx[p] + gy[p] * gy[p] + gz[p] * gz[p];
                }
                if (kind == 0) {
#pragma unroll
                  for (int p = 0; p < 8; ++p) {
                    gx[p] *= gamma; gy[p] *= gamma; gz[p] *= gamma;
                  }
                } else {
                  a.ipa.kbias[((long long)bsamp * 8 + h) * a.ipa.n_res + i] = -0.5f * gamma * d2 + 1e5f * (__ldg(a.ipa.mask + m) - 1.f);
                }
                put8(32, gx);
                put8(33, gy);
                put8(34, gz);
              } else {  // 12 points: x = cols 256..267, y = 268..279, z = 280..291 (z straddles the two passes)
                float g[40];
#pragma unroll
                for (int p = 0; p < 12; ++p) {
                  const float x = pt[p], y = pt[12 + p], z = (24 + p < 32) ? pt[24 + p] : v[24 + p - 32];
                  g[p] = R[0] * x + R[1] * y + R[2] * z + t0;
                  g[12 + p] = R[3] * x + R[4] * y + R[5] * z + t1;
                  g[24 + p] = R[6] * x + R[7] * y + R[8] * z + t2;
                }
#pragma unroll
                for (int p = 36; p < 40; ++p) g[p] = 0.f;
#pragma unroll
                for (int c = 0; c < 5; ++c) put8(32 + c, g + 8 * c);
              }
            }
          }
          continue;
        }
        if constexpr (YIMG) {
          // thread = row, 32 consecutive columns in registers = four 16-byte chunks of the next layer's operand image
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
        if (sw == 32)
          store_transposed<32>(ep, v, stg, lane, mw, n0 + cb);
        else
          store_transposed<16>(ep, v, stg, lane, mw, n0 + cb);
      }
      if (tid == 0 && t == ntiles - 1) LT_TS(11);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) LT_TS(12);
  if (warp == 8) tmem_dealloc(tmem_base, 512);
}

inline size_t lin_tc_fixed_bytes(int nkb, int stg_cols) {
  return 1024 + (size_t)nkb * LT_STAGE_BYTES + (size_t)8 * 32 * (stg_cols + 1) * 4 + (16 + LT_MAX_KB + 4) * 8 + 64;
}
inline size_t lin_tc_smem_bytes(int nkb, int units, int stg_cols) { return lin_tc_fixed_bytes(nkb, stg_cols) + (size_t)units * LT_UNIT_BYTES; }

// W[n, k] fp32 (row stride ldw) -> split operand images [n_tiles][nkb][hi|lo][128 rows][128 B]; rows >= N and columns >= K are zero.
__global__ void pack_weight_split_kernel(const float* __restrict__ W, int ldw, int N, int K, int nkb, int n_tiles, __half* __restrict__ img) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one 8-element chunk
  const long long total = (long long)n_tiles * nkb * 128 * 8;
  if (idx >= total) return;
  const int c = (int)(idx & 7);
  const int r = (int)((idx >> 3) & 127);
  const long long st = idx >> 10;  // nt * nkb + kb
  const int kb = (int)(st % nkb);
  const int nt = (int)(st / nkb);
  const int n = nt * 128 + r, k0 = kb * 64 + c * 8;
  float x[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) x[e] = (n < N && k0 + e < K) ? W[(long long)n * ldw + k0 + e] : 0.f;
  uint4 hi, lo;
  split8(make_float4(x[0], x[1], x[2], x[3]), make_float4(x[4], x[5], x[6], x[7]), hi, lo);
  uint8_t* dst = reinterpret_cast<uint8_t*>(img) + st * LT_STAGE_BYTES + sw128_chunk_off(r, c);
  *reinterpret_cast<uint4*>(dst) = hi;
  *reinterpret_cast<uint4*>(dst + 16384) = lo;
}

}  // namespace tc
}  // namespace fdpt
