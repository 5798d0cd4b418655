// edge_embed_fused.cuh — the pair half of Embedder.forward (score_network.py:114-127, 153-197; data/utils.py:541-550) as ONE persistent
// tcgen05 kernel: pair features are built on the fly, the 3-layer MLP is chained through TMEM / shared memory, LayerNorm + mask in
// the epilogue, and z0 is written once as fp16 tile images (256 B / pair).  Nothing else pair-sized touches HBM.
//
//   x_ij = [f_i (54) | f_j (54) | idx_emb(seq_i - seq_j) (32) | onehot22(bin(|ca_i - ca_j|))]            (162 columns of edge_embedder.0)
//   layer 1 is factorised:  W0 x = A f_i + b0   (per residue, fp32-class GEMM outside: PA_i, enters as an
//                                                        epilogue vector of the tile, all 128 pairs of a tile share i)
//                                                      + B f_j       (k-block 1 of GEMM0: an fp16 image per (b, j-block) shared by all i)
//                                                      + [C | D] [emb | onehot]   (k-block 0 of GEMM0, built per tile by the workers)
//   GEMM0  D0[128 x 128] = [emb | onehot | f_j] (K = 128) . [C | D | B]^T      E0: h1 = relu(D0 + PA_i)            -> fp16 A1
//   GEMM1  D1 = h1 . W2^T                                                        E1: h2 = relu(D1 + b2)              -> fp16 A2
//   GEMM2  D2 = h2 . W4^T                                                        E2: z0 = LN(D2 + b4) * m_i m_j      -> fp16 tile image
//
// All three weight matrices (96 KB as fp16 swizzled images) stay resident in shared memory.  Warps 0-15: workers in four groups of 128
// (thread <-> pair row <-> TMEM lane; group g owns columns [32g, 32g+32) of every buffer it writes; groups 0, 1 build the idx_emb
// half of the per-tile feature block, groups 2, 3 the distogram half), warp 16: MMA issuer + TMEM owner, warp 17: loader.  MMA issue order G1(t), G2(t), G0(t+1): the first GEMM
// of the next tile runs under the LayerNorm epilogue of this one and is done when the workers get there.  fp16 operands (10-bit mantissa = TF32 class, which the pair side
// tolerates: SURVEY §7 hard part 1), fp32 accumulate, positional tables evaluated on the host and gathered here.
#pragma once
#include "tc_common.cuh"

namespace fdpt {
namespace tc {

struct EeArgs {
  int B, N, JB;
  __half* z_out;              // tile images [B][N][JB][32 KB]
  const __half* f_img;        // [B][JB][16 KB] k-block image of the per-residue features f_j (zero padded to 64 columns / 128 rows)
  const float* PA;            // [B*N, 128]  A f_i + b0
  const __half* rel_tab;      // [rel_count][32] fp16 idx_emb(rel_min + r)
  const int32_t* seq_idx;     // [B*N]
  int rel_min, rel_count;
  const float* sc_ca;         // [B*N, 3]
  const float* bin_lower;     // [22]
  const float* b2; const float* b4; const float* ln_g; const float* ln_b;  // [128]
  const float* mask;          // [B*N]
  const __half* W0img;        // image [2 kb][128][128 B]: kb0 = [C | D | 0], kb1 = [B | 0]
  const __half* W2img;        // image [2 kb][128][128 B]
  const __half* W4img;
  long long tiles;            // B*N*JB
  long long* dbg;             // optional clock64 timeline of worker thread 0 of CTA 0: [tile][16] stamps (tools/ee_timeline.py), or nullptr
};

#define EE_TS(id)                                                                                                  \
  do {                                                                                                             \
    if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0 && (t - t_begin) < 8) a.dbg[(t - t_begin) * 16 + (id)] = clock64(); \
  } while (0)

constexpr int EE_W_BYTES = 32768;
constexpr int EE_GROUPS = 4;                      // worker groups of 128 threads (thread <-> pair row <-> TMEM lane)
constexpr int EE_GC = 128 / EE_GROUPS;            // columns of every 128-column buffer owned by one group
constexpr int EE_WORKERS = 128 * EE_GROUPS;
constexpr int EE_WW = EE_WORKERS / 32;            // worker warps; then the MMA warp and the loader warp
constexpr int EE_THREADS = EE_WORKERS + 64;

__global__ void __launch_bounds__(EE_THREADS, 1) ee_fused_kernel(EeArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + smem_align_pad(smem_raw);  // offset arithmetic on the __shared__ symbol: accesses stay LDS / STS
  uint8_t* W0s = smem;                      // 32 KB
  uint8_t* W2s = W0s + EE_W_BYTES;          // 32 KB
  uint8_t* W4s = W2s + EE_W_BYTES;          // 32 KB
  uint8_t* A0t = W4s + EE_W_BYTES;          // 2 x 16 KB per-tile k-block [emb | onehot]
  uint8_t* A0f = A0t + 2 * 16384;           // 16 KB per-(b, j-block) k-block f_j
  uint8_t* A1 = A0f + 16384;                // 32 KB h1 (also staging of the output tile)
  uint8_t* A2 = A1 + 32768;                 // 32 KB h2
  uint64_t* bars = reinterpret_cast<uint64_t*>(A2 + 32768);
  uint64_t* w_full = bars;          // [1]
  uint64_t* an_full = bars + 1;     // [1]
  uint64_t* an_free = bars + 2;     // [1]
  uint64_t* a0_full = bars + 3;     // [2]
  uint64_t* d0_full = bars + 5;     // [1]
  uint64_t* a1_full = bars + 6;     // [1]
  uint64_t* d1_full = bars + 7;     // [1]
  uint64_t* a2_full = bars + 8;     // [1]
  uint64_t* d2_full = bars + 9;     // [1]
  uint64_t* stg_full = bars + 10;   // [1] every worker has written its part of the output tile into A1
  uint64_t* a1_free = bars + 11;    // [1] the bulk store of the staged tile has finished reading A1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  float* PA_s = reinterpret_cast<float*>(tmem_slot + 4);  // [128]
  float* b2_s = PA_s + 128;
  float* b4_s = b2_s + 128;
  float* g_s = b4_s + 128;
  float* be_s = g_s + 128;
  float* lower_s = be_s + 128;  // [24]
  float* PA_s2 = lower_s + 24;  // [128] second PA buffer (tiles alternate between PA_s and PA_s2)
  float* red_s = PA_s2 + 128;   // [EE_GROUPS][128] LayerNorm partial means of the worker groups
  float* red_q = red_s + EE_GROUPS * 128;  // [EE_GROUPS][128] partial sums of squared deviations

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long per = (a.tiles + gridDim.x - 1) / gridDim.x;
  const long long t_begin = (long long)blockIdx.x * per;
  const long long t_end = min(a.tiles, t_begin + per);

  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    mbar_init(an_full, 1);
    mbar_init(an_free, 1);
    mbar_init(&a0_full[0], EE_WORKERS);
    mbar_init(&a0_full[1], EE_WORKERS);
    mbar_init(d0_full, 1);
    mbar_init(a1_full, EE_WORKERS);
    mbar_init(d1_full, 1);
    mbar_init(a2_full, EE_WORKERS);
    mbar_init(d2_full, 1);
    mbar_init(stg_full, EE_WORKERS);
    mbar_init(a1_free, 1);
    fence_barrier_init();
  }
  for (int k = threadIdx.x; k < 128; k += blockDim.x) {
    b2_s[k] = a.b2[k];
    b4_s[k] = a.b4[k];
    g_s[k] = a.ln_g[k];
    be_s[k] = a.ln_b[k];
  }
  if (threadIdx.x < 24) lower_s[threadIdx.x] = threadIdx.x < NBINS ? a.bin_lower[threadIdx.x] : 1e8f;  // [22] = top edge 1e8
  if (warp == EE_WW) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t D0 = tmem_base, D1 = tmem_base + 128, D2 = tmem_base + 256;

  // tiles are ordered (b, jb, i) with i fastest; the f_j image changes when (b, jb) changes
  // (32-bit arithmetic: tiles < 2^31, and a 64-bit division costs ~100 instructions -- three of them per tile and thread were 1.5 k
  //  of the 10.6 k cycles of a tile)
  auto tile_bjb = [&](long long t) -> long long { return (long long)((unsigned)t / (unsigned)a.N); };
  auto tile_m = [&](long long t, int& jb, int& b) -> long long {
    const unsigned bjb = (unsigned)t / (unsigned)a.N;
    const int i = (int)((unsigned)t - bjb * (unsigned)a.N);
    b = (int)(bjb / (unsigned)a.JB);
    jb = (int)(bjb - (unsigned)b * (unsigned)a.JB);
    return (long long)b * a.N + i;
  };

  if (warp == EE_WW + 1) {
    // ============================ loader ============================
    if (elect_one() && t_begin < t_end) {
      mbar_arrive_expect_tx(w_full, 3 * EE_W_BYTES);
      bulk_g2s(W0s, a.W0img, EE_W_BYTES, w_full);
      bulk_g2s(W2s, a.W2img, EE_W_BYTES, w_full);
      bulk_g2s(W4s, a.W4img, EE_W_BYTES, w_full);
      auto load_f = [&](long long t) {
        mbar_arrive_expect_tx(an_full, 16384);
        bulk_g2s(A0f, reinterpret_cast<const uint8_t*>(a.f_img) + tile_bjb(t) * 16384LL, 16384, an_full);
      };
      load_f(t_begin);
      uint32_t nfree = 0;
      for (long long t = t_begin; t < t_end; ++t) {
        if (t + 1 < t_end && tile_bjb(t + 1) != tile_bjb(t)) {
          mbar_wait(an_free, nfree & 1);  // GEMM0 of tile t (last reader of the old image) has completed
          ++nfree;
          load_f(t + 1);
        }
        // output tile t is staged in A1: bulk-store it and hand A1 back once the copy engine has read it
        mbar_wait(stg_full, (uint32_t)(t - t_begin) & 1);
        int jb, b;
        const long long m = tile_m(t, jb, b);
        uint8_t* dst = reinterpret_cast<uint8_t*>(a.z_out) + ((m * a.JB + jb) * 32768LL);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(A1)), "r"(32768) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mbar_arrive(a1_free);
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else if (warp == EE_WW) {
    // ============================ MMA issuer ============================
    if (elect_one() && t_begin < t_end) {
      const uint32_t idesc = make_idesc_f16(128, 128);
      mbar_wait(w_full, 0);
      tc_fence_after();
      const uint32_t w0 = smem_u32(W0s), w2 = smem_u32(W2s), w4 = smem_u32(W4s);
      const uint32_t a0t = smem_u32(A0t), a0f = smem_u32(A0f), a1 = smem_u32(A1), a2 = smem_u32(A2);
      uint32_t an_f = 0;
      auto gemm128 = [&](uint32_t d, uint32_t a_kb0, uint32_t a_kb1, uint32_t w) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(d, make_sw128_desc(a_kb0 + k * 32), make_sw128_desc(w + k * 32), idesc, k ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(d, make_sw128_desc(a_kb1 + k * 32), make_sw128_desc(w + 16384 + k * 32), idesc, 1u);
      };
      auto G0 = [&](long long t) {
        const uint32_t n = (uint32_t)(t - t_begin);
        if (t == t_begin || tile_bjb(t) != tile_bjb(t - 1)) {
          mbar_wait(an_full, an_f & 1);
          ++an_f;
        }
        mbar_wait(&a0_full[n & 1], (n >> 1) & 1);
        tc_fence_after();
        gemm128(D0, a0t + (n & 1) * 16384, a0f, w0);
        umma_commit(d0_full);
        if (t + 1 < t_end && tile_bjb(t + 1) != tile_bjb(t)) umma_commit(an_free);
      };
      G0(t_begin);
      for (long long t = t_begin; t < t_end; ++t) {
        const uint32_t ph = (uint32_t)(t - t_begin) & 1;
        mbar_wait(a1_full, ph);
        tc_fence_after();
        gemm128(D1, a1, a1 + 16384, w2);
        umma_commit(d1_full);
        mbar_wait(a2_full, ph);
        tc_fence_after();
        gemm128(D2, a2, a2 + 16384, w4);
        umma_commit(d2_full);
        if (t + 1 < t_end) G0(t + 1);  // runs under this tile's LayerNorm epilogue (the workers load D0 of this tile long before)
      }
    }
  } else {
    // ============================ workers (EE_GROUPS groups x 128 threads) ============================
    const int wg = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const int cg = wg * EE_GC;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    auto store_part = [&](uint8_t* buf, const float* v /*[32]*/) {  // fp16, swizzled: columns [cg, cg + 32) of row `row`
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float* p = v + c * 8;
        const uint4 u = make_uint4(pack_half2(p[0], p[1]), pack_half2(p[2], p[3]), pack_half2(p[4], p[5]), pack_half2(p[6], p[7]));
        *reinterpret_cast<uint4*>(buf + (wg >> 1) * 16384 + sw128_chunk_off(row, (wg & 1) * 4 + c)) = u;
      }
    };
    auto load_part = [&](uint32_t taddr, float* v /*[32]*/) {
      tmem_ld32(taddr + lane_base + cg, v);
      tmem_ld_wait();
    };
    // per-tile k-block [emb(rel) 32 | onehot(bin) 22 | 0 x 10] of row `row`: group g writes chunks 2g, 2g + 1 (groups 0, 1: the two
    // halves of the embedding row; groups 2, 3: bins 0-15 / bins 16-21 + padding of the distogram one-hot)
    // phases, so that the global-memory latencies do not sit on the workers' critical path: prep_a0 (top of the previous tile) only
    // ISSUES the loads of the residue indices / CA coordinates; after E0, code_a0 reduces them to one small code (groups 0, 1: row of
    // the relative-position table; groups 2, 3: distogram bin; -1: padding row) and emit_a0 gathers the table row and writes the chunks
    struct A0Raw {  // what prep_a0 leaves in registers (loads issued, nothing consumed)
      int si, sj, valid;
      float ci[3], cj[3];
    };
    auto prep_a0 = [&](long long t) -> A0Raw {
      A0Raw r;
      r.si = r.sj = 0;
      r.ci[0] = r.ci[1] = r.ci[2] = r.cj[0] = r.cj[1] = r.cj[2] = 0.f;
      int jb, b;
      const long long m = tile_m(t, jb, b);
      const int j = jb * 128 + row;
      r.valid = j < a.N;
      if (r.valid) {
        const long long mj = (long long)b * a.N + j;
        if (wg < 2) {
          r.si = a.seq_idx[m];
          r.sj = a.seq_idx[mj];
        } else {
#pragma unroll
          for (int e = 0; e < 3; ++e) {
            r.ci[e] = a.sc_ca[m * 3 + e];
            r.cj[e] = a.sc_ca[mj * 3 + e];
          }
        }
      }
      return r;
    };
    auto code_a0 = [&](const A0Raw& r) -> int {
      if (!r.valid) return -1;
      if (wg < 2) return min(max(r.si - r.sj - a.rel_min, 0), a.rel_count - 1);
      const float dx = r.ci[0] - r.cj[0], dy = r.ci[1] - r.cj[1], dz = r.ci[2] - r.cj[2];
      const float d = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
      int bin = -1;
#pragma unroll
      for (int k = 0; k < NBINS; ++k)
        if (d > lower_s[k] && d < lower_s[k + 1]) bin = k;  // strict on both sides (data/utils.py:547-549)
      return bin;
    };
    auto emit_a0 = [&](long long t, int code) {
      uint8_t* dst = A0t + ((t - t_begin) & 1) * 16384;
      uint4 ch[2];
      ch[0] = ch[1] = make_uint4(0, 0, 0, 0);
      if (code >= 0) {
        if (wg < 2) {
          const uint4* e = reinterpret_cast<const uint4*>(a.rel_tab + (long long)code * EMB) + 2 * wg;
          ch[0] = __ldg(e);
          ch[1] = __ldg(e + 1);
        } else {
          const int w0 = (wg - 2) * 8;  // first 32-bit word (= bin pair) of this group's two chunks
          if ((code >> 1) >= w0 && (code >> 1) < w0 + 8) {
            const uint32_t one = (code & 1) ? 0x3C000000u : 0x00003C00u;  // fp16 1.0 in the high / low half
            uint32_t w[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) w[k] = (w0 + k == (code >> 1)) ? one : 0u;
            ch[0] = make_uint4(w[0], w[1], w[2], w[3]);
            ch[1] = make_uint4(w[4], w[5], w[6], w[7]);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) *reinterpret_cast<uint4*>(dst + sw128_chunk_off(row, wg * 2 + c)) = ch[c];
      fence_proxy_async();
      mbar_arrive(&a0_full[(t - t_begin) & 1]);
    };

    // CTA-wide (worker) barriers per tile, none with a global-memory latency behind it: the per-tile vector PA_i is double buffered
    // and fetched one tile ahead, the pair mask is loaded at the top of the tile, the LayerNorm statistics of the four column groups
    // are combined through shared memory, and the wait for the asynchronous store of tile t sits in front of the first write into its
    // staging buffer in tile t+1.
    if (t_begin < t_end) {
      emit_a0(t_begin, code_a0(prep_a0(t_begin)));
      int jb0, b0;
      const long long m0 = tile_m(t_begin, jb0, b0);
      if (threadIdx.x < 128) PA_s[threadIdx.x] = a.PA[m0 * 128 + threadIdx.x];
      asm volatile("bar.sync 1, %0;" ::"n"(EE_WORKERS) : "memory");
    }
    for (long long t = t_begin; t < t_end; ++t) {
      const uint32_t ph = (uint32_t)(t - t_begin) & 1;
      int jb, b;
      const long long m = tile_m(t, jb, b);
      const int j = jb * 128 + row;
      const float* PA_t = ph ? PA_s2 : PA_s;
      float pa_next = 0.f;
      if (t + 1 < t_end && threadIdx.x < 128) {
        int jbn, bn;
        const long long mn = tile_m(t + 1, jbn, bn);
        pa_next = a.PA[mn * 128 + threadIdx.x];  // consumed at the end of this iteration
      }
      float mk_i = 0.f, mk_j = 0.f;  // consumed in E2 (loaded here, multiplied there: no scoreboard wait at the top of the tile)
      if (j < a.N) {
        mk_i = a.mask[m];
        mk_j = a.mask[(long long)b * a.N + j];
      }
      A0Raw raw_next;
      raw_next.valid = 0;
      if (t + 1 < t_end) raw_next = prep_a0(t + 1);
      float v[EE_GC];
      EE_TS(0);
      // ---- E0
      mbar_wait(d0_full, ph);
      tc_fence_after();
      EE_TS(1);
      load_part(D0, v);
      tc_fence_before();
#pragma unroll
      for (int n = 0; n < EE_GC; ++n) v[n] = fmaxf(v[n] + PA_t[cg + n], 0.f);
      EE_TS(2);
      if (t != t_begin) mbar_wait(a1_free, (uint32_t)(t - t_begin - 1) & 1);  // tile t-1's store has left A1
      EE_TS(3);
      store_part(A1, v);
      fence_proxy_async();
      mbar_arrive(a1_full);
      EE_TS(4);
      // the next tile's feature block is gathered here: its two dependent global loads (seq_idx -> rel_tab row) hide under GEMM 1 of
      // this tile, which the workers would otherwise just wait for; GEMM 0 of the next tile is issued after GEMM 1 anyway
      if (t + 1 < t_end) emit_a0(t + 1, code_a0(raw_next));
      EE_TS(5);
      // ---- E1
      mbar_wait(d1_full, ph);
      tc_fence_after();
      EE_TS(6);
      load_part(D1, v);
      tc_fence_before();
#pragma unroll
      for (int n = 0; n < EE_GC; ++n) v[n] = fmaxf(v[n] + b2_s[cg + n], 0.f);
      store_part(A2, v);
      fence_proxy_async();
      mbar_arrive(a2_full);
      EE_TS(7);
      // ---- E2: LayerNorm + mask -> fp16 tile image (staged in A1, free since GEMM1 of this tile has completed) -> bulk store.
      //      Each thread reduces its 32 columns (two-pass); the groups exchange (mean, sum of squared deviations) through shared
      //      memory and combine them with the pairwise update (as in et_fused.cuh)
      mbar_wait(d2_full, ph);
      tc_fence_after();
      EE_TS(8);
      load_part(D2, v);
      tc_fence_before();
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int n = 0; n < EE_GC; n += 2) {
        v[n] += b4_s[cg + n];
        v[n + 1] += b4_s[cg + n + 1];
        s0 += v[n];
        s1 += v[n + 1];
      }
      const float mh = (s0 + s1) * (1.f / EE_GC);
      float q0 = 0.f, q1 = 0.f;
#pragma unroll
      for (int n = 0; n < EE_GC; n += 2) {
        const float d0 = v[n] - mh, d1 = v[n + 1] - mh;
        q0 += d0 * d0;
        q1 += d1 * d1;
      }
      red_s[wg * 128 + row] = mh;
      red_q[wg * 128 + row] = q0 + q1;
      // the other PA buffer was last read in E0 of the previous tile; the barrier below orders this write before E0 of the next one
      if (threadIdx.x < 128 && t + 1 < t_end) (ph ? PA_s : PA_s2)[threadIdx.x] = pa_next;
      asm volatile("bar.sync 1, %0;" ::"n"(EE_WORKERS) : "memory");
      float mean = 0.f, m2 = 0.f;
#pragma unroll
      for (int g = 0; g < EE_GROUPS; ++g) mean += red_s[g * 128 + row];
      mean *= (1.f / EE_GROUPS);
#pragma unroll
      for (int g = 0; g < EE_GROUPS; ++g) {
        const float dg = red_s[g * 128 + row] - mean;
        m2 += red_q[g * 128 + row] + (float)EE_GC * dg * dg;
      }
      EE_TS(9);
      const float rstd = rsqrtf(m2 * (1.f / 128.f) + 1e-5f);
      const float mk = mk_i * mk_j;
#pragma unroll
      for (int n = 0; n < EE_GC; ++n) v[n] = ((v[n] - mean) * rstd * g_s[cg + n] + be_s[cg + n]) * mk;
      store_part(A1, v);
      fence_proxy_async();
      EE_TS(10);
      mbar_arrive(stg_full);  // the loader warp issues the bulk store
      EE_TS(11);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == EE_WW) tmem_dealloc(tmem_base, 512);
}

inline size_t ee_smem_bytes() { return 1024 + 3 * (size_t)EE_W_BYTES + 3 * 16384 + 2 * 32768 + 12 * 8 + 16 + (5 * 128 + 24 + 512 + 2 * EE_GROUPS * 128) * 4 + 64; }

// per-residue features feat1d [B*N, F1] fp32 -> per-(b, j-block) fp16 k-block images [B][JB][128 rows][128 B] (columns >= F1 and rows >= N zero)
__global__ void f_to_image_kernel(int B, int N, int JB, int F1, const float* __restrict__ feat1d, __half* __restrict__ img) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte chunk
  const long long total = (long long)B * JB * 128 * 8;
  if (idx >= total) return;
  const int c = (int)(idx & 7);
  const int r = (int)((idx >> 3) & 127);
  const long long tile = idx >> 10;  // b*JB + jb
  const int jb = (int)(tile % JB);
  const long long b = tile / JB;
  const int j = jb * 128 + r;
  uint32_t w[4] = {0, 0, 0, 0};
  if (j < N) {
    const float* src = feat1d + (b * N + j) * F1;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = c * 8 + 2 * e;
      w[e] = pack_half2(k < F1 ? src[k] : 0.f, k + 1 < F1 ? src[k + 1] : 0.f);
    }
  }
  uint8_t* dst = reinterpret_cast<uint8_t*>(img) + tile * 16384 + sw128_chunk_off(r, c);
  *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void f32_to_f16_kernel(long long n, const float* __restrict__ x, __half* __restrict__ y) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) y[idx] = __float2half_rn(x[idx]);
}

}  // namespace tc
}  // namespace fdpt
