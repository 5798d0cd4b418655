// node_chain.cuh — a whole chain of row-local node-side layers in ONE persistent tcgen05 kernel.
//
// The node side of a trunk block (ipa_pytorch.py:531-547) is a sequence of small layers that only mix channels of the same residue:
// LayerNorm, Linear (+ bias / ReLU / mask / residual), the quaternion update.  Launched one by one (lin_tc.cuh) each of them is bound by
// fixed per-launch latency (TMEM allocation, barrier set-up, operand staging, a grid of 22-66 CTAs on 148 SMs): 73 launches of
// 13-24 us per timestep.  Here a CTA owns one 128-row tile of residues for the WHOLE chain: it executes a small program of ops, one
// after the other, with
//   * the weight stream of every Linear prefetched through one bulk-copy ring by a loader warp that runs ahead across ops (weights are
//     static), pre-split fp16 hi | lo images exactly as in lin_tc.cuh (same arithmetic: 2-term split, two TMEM accumulators);
//   * the activation handed from one Linear to the next as a ready fp16 hi | lo operand image: written by the producing op's epilogue
//     (16-byte swizzled chunks, L2-resident scratch of this CTA's tile) and landed in shared memory by one bulk copy per k-block —
//     no fp32 round trip, no re-split, no staging phase;
//   * LayerNorm (optionally fused with the split-K reduction + bias + mask + residual of the IPA linear_out) as an op of the same
//     kernel: one warp per row, values in registers, fp32 result and/or operand image written;
//   * the backbone update (Linear 256 -> 6 + compose_q_update_vec, rigid_utils.py:1039-1063) folded into that Linear's epilogue;
//   * one TMEM allocation, one barrier set-up, one launch.
// Ops only touch rows of the CTA's own tile, so no inter-CTA synchronisation is needed; inside the CTA consecutive ops are separated
// by a barrier of the 256 worker threads (+ a generic->async proxy fence where the next op bulk-copies what this one stored).
//
// Warps 0-7: workers (operand staging from fp32 when an op has no image source, epilogues, LayerNorm, copies);
// warp 8: MMA issuer + TMEM owner; warp 9: weight loader.
#pragma once
#include "lin_tc.cuh"

namespace fdpt {
namespace tc {

enum { CH_LINEAR = 0, CH_LN = 1 };
enum { CH_EPI_PLAIN = 0, CH_EPI_COMPOSE = 1 };
constexpr int CH_MAX_OPS = 16;
constexpr int CH_UNITS = 3;  // 16 KB ring units next to a 5-k-block activation image (227 KB of shared memory)

struct ChainOp {
  int kind;
  // ---- CH_LINEAR: Y = epi(X @ W^T);  X is either fp32 rows (staged + split by the workers) or a ready operand image
  const float* X; int ldx;
  const __half* Ximg;            // [m-tile][nkb][hi|lo][128 rows][128 B] or nullptr
  int K, N, nkb, n_tiles;
  const __half* Wimg;            // [n_tiles][nkb][hi|lo][128][128 B]
  const float* bias; int relu;
  const float* rowmask;
  const float* residual; int ldr;
  float* Y; int ldy;             // fp32 result or nullptr
  __half* Yimg;                  // operand image of the result for the next Linear, or nullptr
  int epi;                       // CH_EPI_COMPOSE: N = 6 update vector -> quaternion / translation update in place (Y = quats, Y2 = trans,
                                 //                 rowmask = diffuse mask)
  // ---- CH_LN: y = LN(pre) * gamma + beta (* rowmask);  pre = X, or (sum_k X[k * pstride] + pre_bias) * pre_mask + pre_res
  int C;                         // 256 or 320
  int nparts; long long pstride;
  const float* pre_bias; const float* pre_mask; const float* pre_res;
  const float* gamma; const float* beta;
  float* Y2; int ldy2;           // optional second fp32 copy of the result (LN) / translations (compose)
};

struct ChainArgs {
  int M, nops;
  ChainOp ops[CH_MAX_OPS];
};

FDPT_DEVINL int aligned16_dev(const void* p, int ld) { return (reinterpret_cast<uintptr_t>(p) % 16 == 0) && (ld % 4 == 0); }

// ---- one LayerNorm row per warp; lane owns the 8-column chunks lane, lane + 32 ------------------------------------------------
FDPT_DEVINL void chain_ln_rows(const ChainOp& op, int M, int m0, int warp, int lane, int mtile) {
  const int C = op.C, nch = C >> 3, nkb_out = C >> 6;
  for (int rr = 0; rr < 16; ++rr) {
    const int r = warp * 16 + rr;
    const long long m = (long long)m0 + r;
    if (m >= M) break;
    float x[2][8];
    float s = 0.f;
    const float pm = op.pre_mask ? op.pre_mask[m] : 1.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = lane + 32 * i;
#pragma unroll
      for (int e = 0; e < 8; ++e) x[i][e] = 0.f;
      if (c < nch) {
        const float* src = op.X + m * op.ldx + 8 * c;
        float4 a0 = *reinterpret_cast<const float4*>(src), a1 = *reinterpret_cast<const float4*>(src + 4);
        for (int k = 1; k < op.nparts; ++k) {  // fixed order: deterministic
          const float4 b0 = *reinterpret_cast<const float4*>(src + k * op.pstride), b1 = *reinterpret_cast<const float4*>(src + k * op.pstride + 4);
          a0.x += b0.x; a0.y += b0.y; a0.z += b0.z; a0.w += b0.w;
          a1.x += b1.x; a1.y += b1.y; a1.z += b1.z; a1.w += b1.w;
        }
        x[i][0] = a0.x; x[i][1] = a0.y; x[i][2] = a0.z; x[i][3] = a0.w;
        x[i][4] = a1.x; x[i][5] = a1.y; x[i][6] = a1.z; x[i][7] = a1.w;
        if (op.pre_bias) {
#pragma unroll
          for (int e = 0; e < 8; ++e) x[i][e] = (x[i][e] + __ldg(op.pre_bias + 8 * c + e)) * pm;
        }
        if (op.pre_res) {
          const float* rs = op.pre_res + m * op.ldx + 8 * c;
#pragma unroll
          for (int e = 0; e < 8; ++e) x[i][e] += rs[e];
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) s += x[i][e];
      }
    }
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i)
      if (lane + 32 * i < nch) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float d = x[i][e] - mean;
          q += d * d;
        }
      }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + 1e-5f);
    const float mk = op.rowmask ? op.rowmask[m] : 1.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = lane + 32 * i;
      if (c < nch) {
        float y[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) y[e] = ((x[i][e] - mean) * rstd * __ldg(op.gamma + 8 * c + e) + __ldg(op.beta + 8 * c + e)) * mk;
        const float4 y0 = make_float4(y[0], y[1], y[2], y[3]), y1 = make_float4(y[4], y[5], y[6], y[7]);
        if (op.Y) {
          float* d = op.Y + m * op.ldy + 8 * c;
          *reinterpret_cast<float4*>(d) = y0;
          *reinterpret_cast<float4*>(d + 4) = y1;
        }
        if (op.Y2) {
          float* d = op.Y2 + m * op.ldy2 + 8 * c;
          *reinterpret_cast<float4*>(d) = y0;
          *reinterpret_cast<float4*>(d + 4) = y1;
        }
        if (op.Yimg) {
          uint4 hi, lo;
          split8(y0, y1, hi, lo);
          uint8_t* d = reinterpret_cast<uint8_t*>(op.Yimg) + ((size_t)mtile * nkb_out + (c >> 3)) * LT_STAGE_BYTES + sw128_chunk_off(r, c & 7);
          *reinterpret_cast<uint4*>(d) = hi;
          *reinterpret_cast<uint4*>(d + 16384) = lo;
        }
      }
    }
  }
}

// Rigid.compose_q_update_vec (rigid_utils.py:1039-1063, 587-616, 266-275) on one residue: q' = normalise(q + m (q (x) (0, u0..2))),
// t' = t + m R(q) u3..5  (same arithmetic as compose_update_kernel)
FDPT_DEVINL void chain_compose(const float u[6], float mk, float* __restrict__ quats, float* __restrict__ trans) {
  float q[4], R[9];
#pragma unroll
  for (int k = 0; k < 4; ++k) q[k] = quats[k];
  const float v[4] = {0.f, u[0], u[1], u[2]};
  float dq[4], nq[4];
  quat_mul(q, v, dq);
  float n2 = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    nq[k] = q[k] + dq[k] * mk;
    n2 += nq[k] * nq[k];
  }
  const float inv = 1.f / sqrtf(n2);
  quat_to_rot(q, R);
#pragma unroll
  for (int k = 0; k < 4; ++k) quats[k] = nq[k] * inv;
#pragma unroll
  for (int r = 0; r < 3; ++r) trans[r] += (R[r * 3 + 0] * u[3] + R[r * 3 + 1] * u[4] + R[r * 3 + 2] * u[5]) * mk;
}

__global__ void __launch_bounds__(LT_THREADS, 1) node_chain_kernel(const __grid_constant__ ChainArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Aimg = smem;                                               // LT_MAX_KB x [hi 16 KB | lo 16 KB]
  uint8_t* Wst = Aimg + (size_t)LT_MAX_KB * LT_STAGE_BYTES;           // CH_UNITS x 16 KB
  float* Stg = reinterpret_cast<float*>(Wst + (size_t)CH_UNITS * LT_UNIT_BYTES);  // 8 warps x 32 rows x 17 floats
  uint64_t* bars = reinterpret_cast<uint64_t*>(Stg + 8 * 32 * 17);
  uint64_t* w_full = bars;                        // [CH_UNITS]
  uint64_t* w_empty = w_full + CH_UNITS;          // [CH_UNITS]
  uint64_t* a_stg = w_empty + CH_UNITS;           // [LT_MAX_KB] 256 arrivals: k-block staged from fp32 by the workers
  uint64_t* a_img = a_stg + LT_MAX_KB;            // [LT_MAX_KB] 1 arrival + tx: k-block landed by a bulk copy
  uint64_t* acc_full = a_img + LT_MAX_KB;         // [2]
  uint64_t* acc_empty = acc_full + 2;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int mtile = blockIdx.x, m0 = blockIdx.x * 128;

  if (tid == 0) {
    for (int s = 0; s < CH_UNITS; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    for (int k = 0; k < LT_MAX_KB; ++k) {
      mbar_init(&a_stg[k], LT_WORKERS);
      mbar_init(&a_img[k], 1);
    }
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

  if (warp == 9) {
    // ============================ weight loader: runs ahead across ops ============================
    if (lane == 0) {
      int it = 0;
      for (int o = 0; o < a.nops; ++o) {
        const ChainOp& op = a.ops[o];
        if (op.kind != CH_LINEAR) continue;
        for (int nt = 0; nt < op.n_tiles; ++nt)
          for (int kb = 0; kb < op.nkb; ++kb)
            for (int hl = 0; hl < 2; ++hl, ++it) {
              const int s = it % CH_UNITS;
              mbar_wait(&w_empty[s], ((it / CH_UNITS) & 1) ^ 1);
              mbar_arrive_expect_tx(&w_full[s], LT_UNIT_BYTES);
              bulk_g2s(Wst + (size_t)s * LT_UNIT_BYTES,
                       reinterpret_cast<const uint8_t*>(op.Wimg) + ((size_t)nt * op.nkb + kb) * LT_STAGE_BYTES + hl * LT_UNIT_BYTES, LT_UNIT_BYTES,
                       &w_full[s]);
            }
      }
    }
  } else if (warp == 8) {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(128, 128);
      int it = 0, t = 0;
      uint32_t use_stg = 0, use_img = 0;  // per-k-block completed-phase parities (bit kb): ops differ in their k-block counts
      for (int o = 0; o < a.nops; ++o) {
        const ChainOp& op = a.ops[o];
        if (op.kind != CH_LINEAR) continue;
        const bool from_img = op.Ximg != nullptr;
        uint64_t* a_full = from_img ? a_img : a_stg;
        const uint32_t a_par = from_img ? use_img : use_stg;
        for (int nt = 0; nt < op.n_tiles; ++nt, ++t) {
          const int as = t & 1;
          mbar_wait(&acc_empty[as], ((t >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t acc_main = tmem_base + as * 256, acc_x = acc_main + 128;
          for (int kb = 0; kb < op.nkb; ++kb, it += 2) {
            if (nt == 0) mbar_wait(&a_full[kb], (a_par >> kb) & 1u);
            const int sh = it % CH_UNITS, sl = (it + 1) % CH_UNITS;
            mbar_wait(&w_full[sh], (it / CH_UNITS) & 1);
            mbar_wait(&w_full[sl], ((it + 1) / CH_UNITS) & 1);
            tc_fence_after();
            const uint32_t ah = smem_u32(Aimg + (size_t)kb * LT_STAGE_BYTES), al = ah + 16384;
            const uint32_t bh = smem_u32(Wst + (size_t)sh * LT_UNIT_BYTES), bl = smem_u32(Wst + (size_t)sl * LT_UNIT_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t dah = make_sw128_desc(ah + k * 32), dal = make_sw128_desc(al + k * 32);
              const uint64_t dbh = make_sw128_desc(bh + k * 32), dbl = make_sw128_desc(bl + k * 32);
              const uint32_t first = (kb == 0 && k == 0) ? 0u : 1u;
              umma_f16(acc_x, dal, dbh, idesc, first);
              umma_f16(acc_x, dah, dbl, idesc, 1u);
              umma_f16(acc_main, dah, dbh, idesc, first);
            }
            umma_commit(&w_empty[sh]);
            umma_commit(&w_empty[sl]);
          }
          umma_commit(&acc_full[as]);
        }
        const uint32_t used = (1u << op.nkb) - 1u;
        if (from_img) use_img ^= used; else use_stg ^= used;
      }
    }
  } else {
    // ============================ workers ============================
    pdl_wait();  // activations belong to the predecessor kernel until it has completed
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int c_half = (warp >> 2) * 64;
    float* stg = Stg + warp * 32 * 17;
    const int mw = m0 + (warp & 3) * 32;
    int t = 0;
    for (int o = 0; o < a.nops; ++o) {
      const ChainOp& op = a.ops[o];
      if (op.kind == CH_LN) {
        chain_ln_rows(op, a.M, m0, warp, lane, mtile);
      } else {
        // ---- A operand of this Linear
        if (op.Ximg) {
          if (tid == 0) {
            const uint8_t* src = reinterpret_cast<const uint8_t*>(op.Ximg) + (size_t)mtile * op.nkb * LT_STAGE_BYTES;
            for (int kb = 0; kb < op.nkb; ++kb) {
              mbar_arrive_expect_tx(&a_img[kb], LT_STAGE_BYTES);
              bulk_g2s(Aimg + (size_t)kb * LT_STAGE_BYTES, src + (size_t)kb * LT_STAGE_BYTES, LT_STAGE_BYTES, &a_img[kb]);
            }
          }
        } else {
          ChunkPlan pa;
          const int r0 = tid >> 3, c = tid & 7;
          pa.src = op.X + (long long)(m0 + r0) * op.ldx + 8 * c;
          pa.it_stride = 32LL * op.ldx; pa.kb_stride = GT_KB;
          pa.dst = sw128_chunk_off(r0, c); pa.dst_it_stride = 32 * 128; pa.lo_off = 16384;
          pa.iters = 4; pa.kmajor = 1;
          pa.row0 = m0 + r0; pa.row_step = 32; pa.row_lim = a.M; pa.col0 = 8 * c; pa.col_lim = op.K;
          pa.vec = aligned16_dev(op.X, op.ldx);
          RegTile ta[2];
          load_tile(pa, 0, ta[0]);
          for (int kb = 0; kb < op.nkb; kb += 2) {
            if (kb + 1 < op.nkb) load_tile(pa, kb + 1, ta[1]);
            store_tile(pa, Aimg + (size_t)kb * LT_STAGE_BYTES, ta[0]);
            fence_proxy_async();
            mbar_arrive(&a_stg[kb]);
            if (kb + 1 < op.nkb) {
              if (kb + 2 < op.nkb) load_tile(pa, kb + 2, ta[0]);
              store_tile(pa, Aimg + (size_t)(kb + 1) * LT_STAGE_BYTES, ta[1]);
              fence_proxy_async();
              mbar_arrive(&a_stg[kb + 1]);
            }
          }
        }
        // ---- epilogues
        EpiArgs ep;
        ep.M = a.M; ep.N = op.N; ep.alpha = 1.f; ep.bias = op.bias; ep.relu = op.relu; ep.rowmask = op.rowmask; ep.residual = op.residual;
        ep.ldr = op.ldr; ep.accumulate = 0; ep.Y = op.Y; ep.ldy = op.ldy;
        const int nkb_out = (op.N + 63) >> 6;
        for (int nt = 0; nt < op.n_tiles; ++nt, ++t) {
          const int as = t & 1;
          const int n0 = nt * 128;
          mbar_wait(&acc_full[as], (t >> 1) & 1);
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
            if (n0 + cb >= op.N) continue;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaf(x2[j], 1.0f / GT_LO_SCALE, v[j]);
            const int r = (warp & 3) * 32 + lane;
            if (op.epi == CH_EPI_COMPOSE) {
              if (cb == 0 && n0 == 0 && m0 + r < a.M) {
                float u[6];
#pragma unroll
                for (int e = 0; e < 6; ++e) u[e] = v[e] + __ldg(op.bias + e);
                const long long m = (long long)m0 + r;
                chain_compose(u, op.rowmask[m], op.Y + m * 4, op.Y2 + m * 3);
              }
              continue;
            }
            if (op.Yimg) {
              if (m0 + r < a.M) {
                const float rm = op.rowmask ? __ldg(op.rowmask + m0 + r) : 1.f;
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                  const int col = n0 + cb + 8 * cc;
                  if (col >= op.N) break;
                  float y[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) {
                    float tv = v[8 * cc + e] + ((op.bias && col + e < op.N) ? __ldg(op.bias + col + e) : 0.f);
                    if (op.relu) tv = fmaxf(tv, 0.f);
                    tv *= rm;
                    if (op.residual && col + e < op.N) tv += op.residual[(long long)(m0 + r) * op.ldr + col + e];
                    y[e] = (col + e < op.N) ? tv : 0.f;
                  }
                  uint4 hi, lo;
                  split8(make_float4(y[0], y[1], y[2], y[3]), make_float4(y[4], y[5], y[6], y[7]), hi, lo);
                  uint8_t* dst = reinterpret_cast<uint8_t*>(op.Yimg) + ((size_t)mtile * nkb_out + (col >> 6)) * LT_STAGE_BYTES +
                                 sw128_chunk_off(r, (col & 63) >> 3);
                  *reinterpret_cast<uint4*>(dst) = hi;
                  *reinterpret_cast<uint4*>(dst + 16384) = lo;
                }
              }
            }
            if (op.Y) store_transposed<16>(ep, v, stg, lane, mw, n0 + cb);
          }
        }
      }
      // ---- end of op: what this op stored (global fp32 rows / operand image) is read by the next ops of this CTA, possibly through
      //      the async proxy (bulk copy of the image): fence, then a barrier of the 256 workers
      fence_proxy_async_all();
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 512);
}

inline size_t node_chain_smem_bytes() {
  return 1024 + (size_t)LT_MAX_KB * LT_STAGE_BYTES + (size_t)CH_UNITS * LT_UNIT_BYTES + (size_t)8 * 32 * 17 * 4 + (2 * CH_UNITS + 2 * LT_MAX_KB + 4) * 8 + 64;
}

}  // namespace tc
}  // namespace fdpt
