// gemm_tc.cuh — fp32-class GEMM on the 5th-gen tensor cores: 2-term split fp16 (tcgen05.mma kind::f16, fp32 accumulation in TMEM).
//
//   C[M,N] = epi( alpha * A[M,K] @ op(B) )      same GemmArgs / epilogue contract as gemm_simt.cuh
//
// The node side of the network needs fp32-class products (single-pass TF32/fp16 fails the 1e-3 A budget, SURVEY §7 hard part 1).
// Every fp32 operand x is split on the fly into   x_hi = fp16(x)   and   x_lo' = fp16((x - x_hi) * 2^11)   (22 mantissa bits in all;
// the 2^11 scale keeps the low part out of fp16's subnormal range) and the product is evaluated as
//      A·B ~= A_hi·B_hi  +  2^-11 (A_lo'·B_hi + A_hi·B_lo')          (dropped term ~2^-22 relative)
// with the hi·hi products and the cross terms in two separate TMEM accumulators that are combined in the epilogue.  Besides the
// scaling, the second accumulator keeps two thirds of the additions away from the main one: the tensor core's fp32 accumulation
// truncates, and its bias grows with the number of accumulation steps.  Measured: ~1e-6 of max|C| (SIMT fp32: 7e-7; 1 pass: 5e-4).
// A first version used kind::tf32 (hi/lo as fp32, K=8 per MMA; git history): same accuracy, but twice the shared-memory bytes and
// twice the MMA count per K element; it was shared-memory-bandwidth bound at ~70 TFLOP/s.
//
// CTA = one 128 x BN output tile (BN = 128 or 64), 288 threads:
//   warps 0-7  producers: coalesced 32-byte global loads of the A / B k-block (64 K elements) into registers one k-block ahead
//              (double-buffered), split, 16-byte stores into K-major SWIZZLE_128B operand images (B may also be [K,N] row-major =
//              MN-major operand, used by P·V); later the epilogue (tcgen05.ld of both accumulators -> bias / relu / mask /
//              residual -> global; warps w and w+4 share TMEM lanes and take half of the columns each).
//   warp 8     MMA issuer (one elected lane) + TMEM owner.
// 3-stage smem ring (64 KB / stage at BN=128) with full/done mbarriers; arbitrary M, N, K (zero-filled edges), row strides and two
// batch strides, so the same kernel serves Linear layers, per-head Q·K^T and P·V.
#pragma once
#include "gemm_simt.cuh"
#include "tc_common.cuh"

namespace fdpt {
namespace tc {

constexpr int GT_STAGES = 2;
constexpr int GT_BM = 128;
constexpr int GT_KB = 64;          // K elements per k-block (128-byte fp16 rows)
constexpr int GT_PRODUCERS = 256;  // producer / epilogue threads (8 warps); warp 8 issues the MMAs
constexpr int GT_THREADS = GT_PRODUCERS + 32;
constexpr float GT_LO_SCALE = 2048.f;

struct GemmTcArgs {
  GemmArgs g;
  int bn;         // 128 or 64
  int b_kmajor;   // 1: B is [N,K] (weights, K^T);  0: B is [K,N] row-major (P·V)
  int a_vec, b_vec, c_vec;  // 16-byte aligned rows -> float4 path
};

// two fp32 -> packed fp16x2 (lo half = a, hi half = b), round-to-nearest, clamped to the finite fp16 range (one F2FP instruction)
FDPT_DEVINL uint32_t f2h2_sat(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
// 8 consecutive fp32 -> 8 fp16 hi + 8 fp16 scaled lo (one 16-byte chunk each)
FDPT_DEVINL void split8(const float4& x0, const float4& x1, uint4& hi, uint4& lo) {
  const float x[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    h[e] = f2h2_sat(x[2 * e], x[2 * e + 1]);
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h[e]));
    l[e] = f2h2_sat((x[2 * e] - hf.x) * GT_LO_SCALE, (x[2 * e + 1] - hf.y) * GT_LO_SCALE);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// ---- epilogue: row-contiguous global stores --------------------------------------------------------------------------
// With the TMEM-native mapping (thread = accumulator row) every store instruction of a warp touches 32 different lines.  Each warp
// therefore transposes its 32 x 32 block SW columns at a time through a private padded shared-memory patch and stores with
// lane = column: 128-byte (SW = 32) or 2 x 64-byte (SW = 16) segments per instruction; bias / relu / row mask / residual on the way.
struct EpiArgs {
  int M, N;
  float alpha;
  const float* bias; int relu;
  const float* rowmask;
  const float* residual; int ldr;
  int accumulate;
  float* Y; int ldy;
};
// Vectorised variant: the warp's 32 x 32 block goes through a [32 rows][36 floats] patch with 16-byte shared-memory accesses on both
// sides (conflict-free: 8 lanes = one 128-byte row segment) and leaves as float4 stores, 4 rows x 128 bytes per instruction: 8 STS.128 +
// 8 LDS.128 + 8 STG.128 per thread instead of 32 + 32 + 32 scalar ones.  Needs 16-byte aligned rows of Y / residual (ldy, ldr, c0
// multiples of 4) and N a multiple of 4; stg must be 16-byte aligned.
constexpr int ST4_PATCH_FLOATS = 32 * 36;
// COLS = 32: patch [32][36]; COLS = 16: patch [32][20] (both row strides keep the 16-byte accesses of a quarter warp on distinct banks),
// v holds the COLS columns starting at v[0]
template <int COLS = 32>
FDPT_DEVINL void store_transposed_v4(const EpiArgs& a, const float* v, float* stg, int lane, int mw, int c0) {
  constexpr int SLD = COLS + 4, LPR = COLS / 4, RPI = 32 / LPR, NIT = 32 / RPI;  // lanes per row, rows per instruction, instructions
  __syncwarp();
#pragma unroll
  for (int j = 0; j < COLS / 4; ++j) *reinterpret_cast<float4*>(stg + lane * SLD + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
  const int lrow = lane / LPR, lc = (lane % LPR) * 4;
  const int col = c0 + lc;
  const bool cok = col < a.N;
  float4 bs = make_float4(0.f, 0.f, 0.f, 0.f);
  if (a.bias && cok)
    bs = make_float4(__ldg(a.bias + col), col + 1 < a.N ? __ldg(a.bias + col + 1) : 0.f, col + 2 < a.N ? __ldg(a.bias + col + 2) : 0.f,
                     col + 3 < a.N ? __ldg(a.bias + col + 3) : 0.f);
  float4 x[NIT];
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const float4 t = *reinterpret_cast<const float4*>(stg + (it * RPI + lrow) * SLD + lc);
    x[it] = make_float4(fmaf(a.alpha, t.x, bs.x), fmaf(a.alpha, t.y, bs.y), fmaf(a.alpha, t.z, bs.z), fmaf(a.alpha, t.w, bs.w));
  }
  if (a.relu) {
#pragma unroll
    for (int it = 0; it < NIT; ++it) x[it] = make_float4(fmaxf(x[it].x, 0.f), fmaxf(x[it].y, 0.f), fmaxf(x[it].z, 0.f), fmaxf(x[it].w, 0.f));
  }
  if (a.rowmask) {
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int m = mw + it * RPI + lrow;
      const float rm = (m < a.M) ? __ldg(a.rowmask + m) : 0.f;
      x[it] = make_float4(x[it].x * rm, x[it].y * rm, x[it].z * rm, x[it].w * rm);
    }
  }
  if (a.residual) {
    float4 rr[NIT];
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int m = mw + it * RPI + lrow;
      rr[it] = (cok && m < a.M) ? *reinterpret_cast<const float4*>(a.residual + (long long)m * a.ldr + col) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int it = 0; it < NIT; ++it) x[it] = make_float4(x[it].x + rr[it].x, x[it].y + rr[it].y, x[it].z + rr[it].z, x[it].w + rr[it].w);
  }
  if (a.accumulate) {
    float4 rr[NIT];
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int m = mw + it * RPI + lrow;
      rr[it] = (cok && m < a.M) ? *reinterpret_cast<const float4*>(a.Y + (long long)m * a.ldy + col) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int it = 0; it < NIT; ++it) x[it] = make_float4(x[it].x + rr[it].x, x[it].y + rr[it].y, x[it].z + rr[it].z, x[it].w + rr[it].w);
  }
  if (cok && col + 3 >= a.N) {  // the float4 straddles N (N not a multiple of 4): element stores, columns >= N stay untouched
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int m = mw + it * RPI + lrow;
      if (m >= a.M) continue;
      float* yp = a.Y + (long long)m * a.ldy + col;
      const float e[4] = {x[it].x, x[it].y, x[it].z, x[it].w};
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (col + q < a.N) yp[q] = e[q];
    }
    return;
  }
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int m = mw + it * RPI + lrow;
    if (cok && m < a.M) *reinterpret_cast<float4*>(a.Y + (long long)m * a.ldy + col) = x[it];
  }
}

template <int SW>
FDPT_DEVINL void store_transposed(const EpiArgs& a, const float (&v)[32], float* stg, int lane, int mw, int c0) {
  constexpr int SLD = SW + 1, RPI = 32 / SW, NIT = 32 / RPI;  // rows per store instruction, store instructions per pass
  const int lrow = lane / SW, lcol = lane % SW;
#pragma unroll
  for (int h0 = 0; h0 < 32; h0 += SW) {
    __syncwarp();
#pragma unroll
    for (int j = 0; j < SW; ++j) stg[lane * SLD + j] = v[h0 + j];
    __syncwarp();
    const int col = c0 + h0 + lcol;
    const bool cok = col < a.N;
    const float bs = (a.bias && cok) ? __ldg(a.bias + col) : 0.f;
    float x[NIT];
#pragma unroll
    for (int it = 0; it < NIT; ++it) x[it] = fmaf(a.alpha, stg[(it * RPI + lrow) * SLD + lcol], bs);
    if (a.relu) {
#pragma unroll
      for (int it = 0; it < NIT; ++it) x[it] = fmaxf(x[it], 0.f);
    }
    if (a.rowmask) {
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int m = mw + it * RPI + lrow;
        x[it] *= (m < a.M) ? __ldg(a.rowmask + m) : 0.f;
      }
    }
    float* yp = a.Y + (long long)(mw + lrow) * a.ldy + col;
    if (a.residual) {
      const float* rp = a.residual + (long long)(mw + lrow) * a.ldr + col;
      float rr[NIT];
#pragma unroll
      for (int it = 0; it < NIT; ++it) rr[it] = (cok && mw + it * RPI + lrow < a.M) ? rp[(long long)it * RPI * a.ldr] : 0.f;
#pragma unroll
      for (int it = 0; it < NIT; ++it) x[it] += rr[it];
    }
    if (a.accumulate) {
      float rr[NIT];
#pragma unroll
      for (int it = 0; it < NIT; ++it) rr[it] = (cok && mw + it * RPI + lrow < a.M) ? yp[(long long)it * RPI * a.ldy] : 0.f;
#pragma unroll
      for (int it = 0; it < NIT; ++it) x[it] += rr[it];
    }
#pragma unroll
    for (int it = 0; it < NIT; ++it)
      if (cok && mw + it * RPI + lrow < a.M) yp[(long long)it * RPI * a.ldy] = x[it];
  }
}

// per-thread copy plan of one operand: `iters` 8-element chunks per k-block, affine in the iteration index
struct ChunkPlan {
  const float* src;        // first chunk of k-block 0
  long long it_stride;     // elements between consecutive iterations
  long long kb_stride;     // elements between consecutive k-blocks
  uint32_t dst, dst_it_stride, lo_off;  // byte offset of the hi chunk inside the stage, per-iteration stride, hi -> lo image distance
  int iters;
  int kmajor;              // 1: validity along the chunk = K tail, per-iteration = row;  0: chunk = N tail, per-iteration = k row
  int row0, row_step, row_lim;  // row (m / n) or k index of iteration 0, its step, and its limit
  int col0, col_lim;            // first element index along the chunk direction and its limit
  int vec;
};
struct RegTile {
  float4 v[4][2];
};

FDPT_DEVINL void load_tile(const ChunkPlan& p, int kb, RegTile& t) {
  const float* src = p.src + (long long)kb * p.kb_stride;
  const int cvalid = p.kmajor ? p.col_lim - (kb * GT_KB + p.col0) : p.col_lim - p.col0;
  const int roff = p.kmajor ? 0 : kb * GT_KB;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    t.v[it][0] = make_float4(0.f, 0.f, 0.f, 0.f);
    t.v[it][1] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (it < p.iters && (p.row0 + roff + it * p.row_step) < p.row_lim && cvalid > 0) {
      const float* s = src + it * p.it_stride;
      if (cvalid >= 8 && p.vec) {
        t.v[it][0] = __ldg(reinterpret_cast<const float4*>(s));
        t.v[it][1] = __ldg(reinterpret_cast<const float4*>(s) + 1);
      } else {
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = e < cvalid ? __ldg(s + e) : 0.f;
        t.v[it][0] = make_float4(x[0], x[1], x[2], x[3]);
        t.v[it][1] = make_float4(x[4], x[5], x[6], x[7]);
      }
    }
  }
}
FDPT_DEVINL void store_tile(const ChunkPlan& p, uint8_t* stage, const RegTile& t) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    if (it < p.iters) {
      uint4 hi, lo;
      split8(t.v[it][0], t.v[it][1], hi, lo);
      uint8_t* d = stage + p.dst + it * p.dst_it_stride;
      *reinterpret_cast<uint4*>(d) = hi;
      *reinterpret_cast<uint4*>(d + p.lo_off) = lo;
    }
  }
}

__global__ void __launch_bounds__(GT_THREADS, 2) gemm_tc_kernel(GemmTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + smem_align_pad(smem_raw);  // offset arithmetic on the __shared__ symbol: accesses stay LDS / STS
  const GemmArgs& g = a.g;
  const int BN = a.bn;
  const uint32_t a_bytes = GT_BM * 128, b_bytes = (uint32_t)BN * 128;
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;  // [A_hi][A_lo][B_hi][B_lo]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GT_STAGES * stage_bytes);
  uint64_t* full = bars;                 // [GT_STAGES] GT_PRODUCERS arrivals: operand images of k-block kb are ready
  uint64_t* done = bars + GT_STAGES;     // [GT_STAGES] tcgen05.commit: the MMAs of k-block kb have completed
  uint64_t* acc_full = done + GT_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int m0 = blockIdx.x * GT_BM, n0 = blockIdx.y * BN;
  const int b1 = blockIdx.z / g.batch2, b2 = blockIdx.z % g.batch2;
  const float* __restrict__ A = g.A + b1 * g.sA1 + b2 * g.sA2;
  const float* __restrict__ B = g.B + b1 * g.sB1 + b2 * g.sB2;
  float* C = g.C + b1 * g.sC1 + b2 * g.sC2;
  const int nkb = (g.K + GT_KB - 1) / GT_KB;
  const int n_atoms = BN / 64;  // MN-major B: 128-byte atoms along n

  if (tid == 0) {
    for (int s = 0; s < GT_STAGES; ++s) {
      mbar_init(&full[s], GT_PRODUCERS);
      mbar_init(&done[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == GT_PRODUCERS / 32) tmem_alloc(tmem_slot, (uint32_t)(2 * BN));  // [0,BN) hi*hi, [BN,2BN) cross terms
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  if (warp == GT_PRODUCERS / 32) {
    // ============================ MMA issuer ============================
    if (elect_one()) {
      const uint32_t idesc = make_idesc_f16(GT_BM, BN) | (a.b_kmajor ? 0u : (1u << 16));
      const uint32_t acc_main = tmem_base, acc_x = tmem_base + (uint32_t)BN;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % GT_STAGES;
        mbar_wait(&full[s], (kb / GT_STAGES) & 1);
        tc_fence_after();
        const uint32_t ah = smem_u32(smem + s * stage_bytes), al = ah + a_bytes, bh = al + a_bytes, bl = bh + b_bytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // 4 x k16 per 128-byte k-block
          const uint64_t dah = make_sw128_desc(ah + k * 32), dal = make_sw128_desc(al + k * 32);
          uint64_t dbh, dbl;
          if (a.b_kmajor) {
            dbh = make_sw128_desc(bh + k * 32);
            dbl = make_sw128_desc(bl + k * 32);
          } else {  // [k-atom (8 rows)][n-atom (64 cols)][1024 B]: LBO steps along n, SBO along k
            const uint32_t koff = (uint32_t)(2 * k * n_atoms) * 1024;
            dbh = make_sw128_desc_ls(bh + koff, 1024, (uint32_t)n_atoms * 1024);
            dbl = make_sw128_desc_ls(bl + koff, 1024, (uint32_t)n_atoms * 1024);
          }
          const uint32_t first = (kb == 0 && k == 0) ? 0u : 1u;
          umma_f16(acc_x, dal, dbh, idesc, first);
          umma_f16(acc_x, dah, dbl, idesc, 1u);
          umma_f16(acc_main, dah, dbh, idesc, first);
        }
        umma_commit(&done[s]);
      }
      umma_commit(acc_full);
    }
  } else {
    // ============================ producers (256 threads) ============================
    ChunkPlan pa, pb;
    {
      const int r0 = tid >> 3, c = tid & 7;
      pa.src = A + (long long)(m0 + r0) * g.lda + 8 * c;
      pa.it_stride = 32LL * g.lda; pa.kb_stride = GT_KB;
      pa.dst = sw128_chunk_off(r0, c); pa.dst_it_stride = 32 * 128; pa.lo_off = a_bytes;
      pa.iters = GT_BM / 32; pa.kmajor = 1;
      pa.row0 = m0 + r0; pa.row_step = 32; pa.row_lim = g.M; pa.col0 = 8 * c; pa.col_lim = g.K; pa.vec = a.a_vec;
      pb.lo_off = b_bytes;
      if (a.b_kmajor) {
        pb.src = B + (long long)(n0 + r0) * g.ldb + 8 * c;
        pb.it_stride = 32LL * g.ldb; pb.kb_stride = GT_KB;
        pb.dst = 2 * a_bytes + sw128_chunk_off(r0, c); pb.dst_it_stride = 32 * 128;
        pb.iters = BN / 32; pb.kmajor = 1;
        pb.row0 = n0 + r0; pb.row_step = 32; pb.row_lim = g.N; pb.col0 = 8 * c; pb.col_lim = g.K; pb.vec = a.b_vec;
      } else {
        const int cpr = BN / 8, kk0 = tid / cpr, cc = tid % cpr, kstep = GT_PRODUCERS / cpr;  // kstep = 16 (BN=128) or 32
        pb.src = B + (long long)kk0 * g.ldb + n0 + 8 * cc;
        pb.it_stride = (long long)kstep * g.ldb; pb.kb_stride = (long long)GT_KB * g.ldb;
        pb.dst = 2 * a_bytes + (uint32_t)(((kk0 >> 3) * n_atoms + (cc >> 3)) * 1024 + (kk0 & 7) * 128 + (((cc & 7) ^ (kk0 & 7)) << 4));
        pb.dst_it_stride = (uint32_t)(kstep / 8) * n_atoms * 1024;
        pb.iters = GT_KB / kstep; pb.kmajor = 0;
        pb.row0 = kk0; pb.row_step = kstep; pb.row_lim = g.K; pb.col0 = n0 + 8 * cc; pb.col_lim = g.N; pb.vec = a.b_vec;
      }
    }
    // one register tile per operand (<= 112 registers so that two CTAs share an SM: while one waits on its loads or runs its
    // epilogue, the other converts / multiplies)
    RegTile ta, tb;
    pdl_wait();  // operands (and residual / C) belong to the predecessor kernel until it has completed
    for (int kb = 0; kb < nkb; ++kb) {
      load_tile(pa, kb, ta);
      load_tile(pb, kb, tb);
      if (kb >= GT_STAGES) mbar_wait(&done[kb % GT_STAGES], ((kb - GT_STAGES) / GT_STAGES) & 1);
      uint8_t* stage = smem + (kb % GT_STAGES) * stage_bytes;
      store_tile(pa, stage, ta);
      store_tile(pb, stage, tb);
      fence_proxy_async();
      mbar_arrive(&full[kb % GT_STAGES]);
    }
    // ============================ epilogue ============================
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int c_begin = (warp >> 2) * (BN / 2), c_end = c_begin + BN / 2;
    // all MMAs have completed, so the operand ring is idle: its first bytes become the per-warp transposition patches
    float* stg = reinterpret_cast<float*>(smem) + warp * 32 * 33;
    const int mw = m0 + (warp & 3) * 32;
    EpiArgs ep;
    ep.M = g.M; ep.N = g.N; ep.alpha = g.alpha; ep.bias = g.bias ? g.bias + b1 * g.sBias1 + b2 * g.sBias2 : nullptr; ep.relu = g.relu; ep.rowmask = g.rowmask; ep.residual = g.residual;
    ep.ldr = g.ldr; ep.accumulate = g.accumulate; ep.Y = C; ep.ldy = g.ldc;
    for (int cb = c_begin; cb < c_end; cb += 32) {
      float v[32], x2[32];
      tmem_ld32(tmem_base + lane_base + cb, v);
      tmem_ld32(tmem_base + lane_base + BN + cb, x2);
      tmem_ld_wait();
      if (n0 + cb >= g.N) continue;
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaf(x2[j], 1.0f / GT_LO_SCALE, v[j]);
      store_transposed<16>(ep, v, stg, lane, mw, n0 + cb);  // 16-column passes: measured faster than one 32-column pass
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == GT_PRODUCERS / 32) tmem_dealloc(tmem_base, (uint32_t)(2 * BN));
}

inline size_t gemm_tc_smem_bytes(int bn) { return 1024 + (size_t)GT_STAGES * (2 * GT_BM * 128 + 2 * (size_t)bn * 128) + 128; }

inline bool aligned16(const void* p, long long ld, long long s1, long long s2) {
  return (reinterpret_cast<uintptr_t>(p) % 16 == 0) && (ld % 4 == 0) && (s1 % 4 == 0) && (s2 % 4 == 0);
}

// Same contract as launch_gemm (gemm_simt.cuh). Problems the tensor-core tile shape cannot serve well (N < 16, or the
// pair-broadcast epilogue) stay on the SIMT kernel.
static int g_force_bn = 0;  // bring-up knob (FDPT_OPT_DEBUG_FLAGS bit 3)
static int g_use_pdl = 1;   // programmatic dependent launch of the GEMM kernels (FDPT_OPT_DEBUG_FLAGS bit 4 turns it off)

inline cudaError_t launch_gemm_tc(const GemmArgs& g, bool b_kmajor, int batch, cudaStream_t st, int num_sms) {
  if (g.M <= 0 || g.N <= 0) return cudaSuccess;
  if (g.N < 16 || g.U != nullptr || g.K <= 0) return launch_gemm(g, b_kmajor, batch, st);
  GemmTcArgs a;
  a.g = g;
  a.b_kmajor = b_kmajor ? 1 : 0;
  const long long tiles128 = (long long)((g.M + GT_BM - 1) / GT_BM) * ((g.N + 127) / 128) * batch;
  (void)tiles128; (void)num_sms;
  // 64 columns: 96 KB of shared memory per CTA -> two CTAs per SM (best for the latency-bound single GEMMs); the batched attention
  // GEMMs (many tiles, L2-traffic bound) prefer 128 columns: each operand tile is loaded and split for half as many CTAs
  a.bn = ((batch >= 8 || g_force_bn == 128) && g.N > 128) ? 128 : 64;
  a.a_vec = aligned16(g.A, g.lda, g.sA1, g.sA2);
  a.b_vec = aligned16(g.B, g.ldb, g.sB1, g.sB2);
  a.c_vec = aligned16(g.C, g.ldc, g.sC1, g.sC2) && (!g.residual || aligned16(g.residual, g.ldr, 0, 0));
  dim3 grid((g.M + GT_BM - 1) / GT_BM, (g.N + a.bn - 1) / a.bn, batch);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(GT_THREADS);
  cfg.dynamicSmemBytes = gemm_tc_smem_bytes(a.bn);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_use_pdl;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, gemm_tc_kernel, a);
}

}  // namespace tc
}  // namespace fdpt
