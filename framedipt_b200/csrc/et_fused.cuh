// et_fused.cuh — EdgeTransition (ipa_pytorch.py:61-102) as ONE persistent tcgen05 kernel.
//
//   x = [z_ij | n_i | n_j] (384);  y = relu(W2 relu(W1 x + b1) + b2) + x;  z' = LN(Wf y + bf) * m_i m_j
//
// Per 128-pair tile (fixed b, i; 128 consecutive j) the three GEMMs are chained through TMEM/shared memory; only z is
// read (fp16 tile image, 32 KB) and written (32 KB) in HBM.  The n_i terms are per-residue and enter as epilogue
// vectors (U_i = W1[:,128:256] n_i + b1, Pf_i = Wf[:,128:256] n_i + bf); the n_j terms enter as two extra k-blocks of
// the A operand (a per-(b, j-block) fp16 image shared by all i).  The residual Wf x is folded into GEMM 3.
//
//   GEMM1  D1[128 x 384] = [z | n_j] (K=256) . W1cat^T          W1cat = [W1[:, 0:128] | W1[:, 256:384]]
//   E1     h1 = relu(D1 + U_i)                      -> fp16, 128-column chunks in shared memory
//   GEMM2  D2[128 x 384] += h1_chunk(c) . W2[:, chunk c]^T       (accumulated over the three K chunks as they appear)
//   E2     r2 = relu(D2 + b2)                       -> fp16, written back IN PLACE into D2's tensor-memory columns (tcgen05.st: two
//                                                      K elements per 32-bit column) = the A operand of GEMM3's partial products
//   GEMM3  D3[128 x 128] = [z | n_j] . W3cat[:, 384:640]^T + sum_c r2_chunk(c) . W3cat[:, chunk c]^T      (A from tensor memory)
//                                                     W3cat = [Wf | Wf[:, 0:128] | Wf[:, 256:384]]
//   E3     z' = LN(D3 + Pf_i) * mask                -> fp16 tile image, bulk store
//
// TMEM: columns [0,384) = D2 (then r2, packed fp16, in place), [384,512) = DS: staging of GEMM 1's chunks 0 and 2, then D3; D2's columns
// [256,384) double as the staging buffer of chunk 1 (they are idle between the previous tile's last partial product and the tile's own
// G2(0), and the tensor pipe executes in order), so chunks 0 and 1 of the next tile are issued while the workers are still in E3.
// Issue order per tile: G1(0) G1(1) G1(2) G2(0) G3[z] G2(1) G2(2) G3[n_j] G3[r2 chunks]; GEMM 2's output chunks are handed to E2 one by one.
// Shared memory (216 KB): A0z (32 KB, bulk-copied; the next tile's z is fetched as soon as G3[z] has read this one), A0n (32 KB), two
// 32 KB chunk buffers (h1 / output staging), 5-stage weight ring (80 KB: the ring depth, not the tensor pipe, paced this kernel with 3
// stages).  The 40 weight k-blocks of a tile are 8 full turns of the ring, so the issue sequence is fully unrolled with compile-time
// stage indices and barrier parities.
// Warps 0-15: epilogue workers in four groups of 128 (thread <-> tile row <-> TMEM lane; group g owns columns [32g, 32g+32) of every
// 128-column chunk: four warps per scheduler hide the TMEM-load / shared-memory latencies of each other); warp 16: MMA issuer (one
// elect.sync lane) + TMEM owner; warp 17: weight / z / n_j loader; warp 18: prefetches the per-tile epilogue vectors (U_i, Pf_i, pair
// mask) of the next tile into a double-buffered shared-memory slot so that the workers never wait on a global load between tiles,
// and issues the bulk store of the finished tile.
// All operands fp16 (10-bit mantissa = TF32 precision, which the pair side tolerates: SURVEY §7 hard part 1), fp32 accumulate.
#pragma once
#include "tc_common.cuh"
#include "tmem_a_test.cuh"

namespace fdpt {
namespace tc {

constexpr int ET_TILE_BYTES = 32768;    // 128 rows x 128 halfs (two k-blocks)
constexpr int ET_WSTAGES = 5;
constexpr int ET_STAGE_BYTES = 16384;   // 128 weight rows x one k-block
constexpr int ET_KB_PER_TILE = 40;      // weight k-blocks streamed per tile (G1 12 + G2 18 + G3 10)
// every tile starts at ring stage 0 with an even number of uses per stage, so stage index and barrier parity of each of the 40 uses
// are compile-time constants in the fully unrolled issue sequence (no modulo / phase arithmetic on the issuing thread's critical path)
static_assert(ET_KB_PER_TILE % ET_WSTAGES == 0 && (ET_KB_PER_TILE / ET_WSTAGES) % 2 == 0, "ring depth must divide the per-tile stream evenly");

struct EtArgs {
  int B, N, JB;                 // JB = ceil(N/128) j-blocks
  const __half* z_in;           // tile images [B][N][JB][32 KB]
  __half* z_out;                // may alias z_in
  const __half* n_img;          // [B][JB][32 KB] images of n_emb rows (zero padded)
  const float* Ui;              // [B*N, 384]
  const float* Pf;              // [B*N, 128]
  const float* b2;              // [384]
  const float* ln_g; const float* ln_b;  // [128]
  const float* mask;            // [B*N]
  const __half* W1cat;          // image [4 kb][384][128 B]
  const __half* W2;             // image [6 kb][384][128 B]
  const __half* W3cat;          // image [10 kb][128][128 B]
  long long tiles;              // B*N*JB
  int exp;                      // timing experiments only (results wrong): 1 no weight copies, 2 MMAs shrunk to N=16, 4 no shared-memory stores in E1/E3, 8 no z load/store
  long long* dbg;               // optional clock64 timeline of CTA 0 (bring-up / profiling aid): [tile][48] stamps, or nullptr
};

#define ET_TS(id)                                                                                  \
  do {                                                                                             \
    if (a.dbg && blockIdx.x == ((a.exp & 8) ? 77 : 0) && (t - t_begin) >= ((a.exp & 8) ? 40 : 0) && (t - t_begin) < ((a.exp & 8) ? 48 : 8)) \
      a.dbg[((t - t_begin) & 7) * 48 + (id)] = clock64();                                            \
  } while (0)

struct EtPhase {  // phase counters of one role
  uint32_t w = 0, az[2] = {0, 0}, an = 0, ds_full = 0, ds_empty = 0, buf_full[2] = {0, 0}, buf_free[2] = {0, 0}, d2_full = 0, d2_empty = 0;
};

constexpr int ET_GROUPS = 4;                      // epilogue worker groups (128 threads each)
constexpr int ET_GC = 128 / ET_GROUPS;            // columns of a 128-column chunk owned by one group
constexpr int ET_WORKERS = 128 * ET_GROUPS;
constexpr int ET_WW = ET_WORKERS / 32;            // worker warps; then: MMA warp, loader warp, epilogue-vector warp
constexpr int ET_THREADS = ET_WORKERS + 96;  // + MMA warp, weight loader warp, epilogue-vector prefetch warp

__global__ void __launch_bounds__(ET_THREADS, 1) et_fused_kernel(EtArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + smem_align_pad(smem_raw);  // offset arithmetic on the __shared__ symbol: accesses stay LDS / STS
  uint8_t* A0z = smem;                                 // 32 KB (single buffer: the next tile's z is fetched under G3's partial products + E3)
  uint8_t* A0n = A0z + ET_TILE_BYTES;                  // 32 KB
  uint8_t* BUF = A0n + ET_TILE_BYTES;                  // 2 x 32 KB
  uint8_t* WST = BUF + 2 * ET_TILE_BYTES;              // ET_WSTAGES x 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(WST + ET_WSTAGES * ET_STAGE_BYTES);
  uint64_t* w_full = bars;                   // [ET_WSTAGES]
  uint64_t* w_empty = w_full + ET_WSTAGES;   // [4]
  uint64_t* az_full = w_empty + ET_WSTAGES;  // [2]
  uint64_t* az_empty = az_full + 2;          // [2]
  uint64_t* an_full = az_empty + 2;          // [1]
  uint64_t* ds_full = an_full + 1;           // [1]
  uint64_t* ds_empty = ds_full + 1;          // [1]
  uint64_t* buf_full = ds_empty + 1;         // [2]
  uint64_t* buf_free = buf_full + 2;         // [2]
  uint64_t* d2_full = buf_free + 2;          // [3] output chunk n of GEMM 2 is complete (committed after each n-block of G2(2))
  uint64_t* sb_full = d2_full + 3;           // [1] G1(1)'s output is in the second staging buffer (D2's idle columns [256,384))
  uint64_t* sb_empty = sb_full + 1;          // [1] ... and the workers have drained it
  uint64_t* an_empty = sb_empty + 1;         // [1] the tile's last reader of the n_j image has completed
  uint64_t* vec_full = an_empty + 1;         // [2]
  uint64_t* vec_free = vec_full + 2;         // [2]
  uint64_t* stg_full = vec_free + 2;         // [1] all workers have written their part of the output tile into BUF[1]
  uint64_t* r2_full = stg_full + 1;          // [3] chunk c of r2 is in tensor memory (fp16, A operand of GEMM3)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(r2_full + 3);
  float* Ui_s = reinterpret_cast<float*>(tmem_slot + 4);  // [2][384]
  float* Pf_s = Ui_s + 2 * 384;                            // [2][128]
  float* b2_s = Pf_s + 2 * 128;                            // [384]
  float* g_s = b2_s + 384;                                 // [128]
  float* be_s = g_s + 128;                                 // [128]
  float* red_s = be_s + 128;                               // [ET_GROUPS][128] LayerNorm partial means of the worker groups
  float* red_q = red_s + ET_GROUPS * 128;                  // [ET_GROUPS][128] partial sums of squared deviations
  float* mk_s = red_q + ET_GROUPS * 128;                   // [2][128] pair mask of the tile's rows (prefetched with the epilogue vectors)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long per = (a.tiles + gridDim.x - 1) / gridDim.x;
  const long long t_begin = (long long)blockIdx.x * per;
  const long long t_end = min(a.tiles, t_begin + per);

  if (threadIdx.x == 0) {
    for (int s = 0; s < ET_WSTAGES; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&az_full[s], 1);
      mbar_init(&az_empty[s], 1);
      mbar_init(&buf_full[s], ET_WORKERS);
      mbar_init(&buf_free[s], 1);
    }
    mbar_init(an_full, 1);
    mbar_init(ds_full, 1);
    mbar_init(ds_empty, ET_WORKERS);
    for (int c = 0; c < 3; ++c) mbar_init(&d2_full[c], 1);
    mbar_init(sb_full, 1);
    mbar_init(sb_empty, ET_WORKERS);
    mbar_init(an_empty, 1);
    mbar_init(stg_full, ET_WORKERS);
    for (int c = 0; c < 3; ++c) mbar_init(&r2_full[c], ET_WORKERS);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&vec_full[s], 32);
      mbar_init(&vec_free[s], ET_WORKERS);
    }
    fence_barrier_init();
  }
  for (int k = threadIdx.x; k < 384; k += blockDim.x) b2_s[k] = a.b2[k];
  for (int k = threadIdx.x; k < 128; k += blockDim.x) {
    g_s[k] = a.ln_g[k];
    be_s[k] = a.ln_b[k];
  }
  if (warp == ET_WW) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t D2 = tmem_base, DS = tmem_base + 384, SB = tmem_base + 256;

  // tile -> (b*N+i, jb); the n_j image changes when (b, jb) changes. Tiles are ordered (b, jb, i) with i fastest.
  auto tile_bjb = [&](long long t) -> long long { return (long long)((unsigned)t / (unsigned)a.N); };  // b*JB + jb (tiles < 2^31)
  auto tile_mb = [&](long long t, int& jb, int& b) -> long long {             // returns b*N + i
    const unsigned bjb = (unsigned)t / (unsigned)a.N;
    const int i = (int)((unsigned)t - bjb * (unsigned)a.N);
    b = (int)(bjb / (unsigned)a.JB);
    jb = (int)(bjb - (unsigned)b * (unsigned)a.JB);
    return (long long)b * a.N + i;
  };
  auto tile_m = [&](long long t, int& jb) -> long long {
    int b;
    return tile_mb(t, jb, b);
  };

  if (warp == ET_WW + 1) {
    // ============================ loader ============================
    if (elect_one() && t_begin < t_end) {
      uint32_t wit = 0;      // weight stage counter of the tile
      auto load_z = [&](long long t) {
        mbar_arrive_expect_tx(&az_full[0], ET_TILE_BYTES);
        int jb;
        const long long m = tile_m(t, jb);
        bulk_g2s(A0z, reinterpret_cast<const uint8_t*>(a.z_in) + ((m * a.JB + jb) * (long long)ET_TILE_BYTES), ET_TILE_BYTES, &az_full[0]);
      };
      auto load_n = [&](long long t) {
        mbar_arrive_expect_tx(an_full, ET_TILE_BYTES);
        bulk_g2s(A0n, reinterpret_cast<const uint8_t*>(a.n_img) + tile_bjb(t) * (long long)ET_TILE_BYTES, ET_TILE_BYTES, an_full);
      };
      auto stage = [&](const __half* img, int rows_total, int row0, int kb) {
        const int s = wit % ET_WSTAGES;
        mbar_wait(&w_empty[s], ((wit / ET_WSTAGES) & 1) ^ 1);
        if (a.exp & 1) {
          mbar_arrive(&w_full[s]);
        } else {
          mbar_arrive_expect_tx(&w_full[s], ET_STAGE_BYTES);
          bulk_g2s(WST + s * ET_STAGE_BYTES, reinterpret_cast<const uint8_t*>(img) + ((size_t)kb * rows_total + row0) * 128, ET_STAGE_BYTES,
                   &w_full[s]);
        }
        ++wit;
      };
      load_z(t_begin);
      load_n(t_begin);
      for (long long t = t_begin; t < t_end; ++t) {
        // weight stream in the exact order the MMA warp consumes it
        wit = 0;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) stage(a.W1cat, 384, c * 128, kb);                           // G1(0), G1(1), G1(2)
#pragma unroll
        for (int n = 0; n < 3; ++n)
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) stage(a.W2, 384, n * 128, kb);                              // G2(0)
#pragma unroll
        for (int kb = 6; kb < 8; ++kb) stage(a.W3cat, 128, 0, kb);                                   // G3 static, z part
#pragma unroll
        for (int n = 0; n < 3; ++n)
#pragma unroll
          for (int kb = 2; kb < 4; ++kb) stage(a.W2, 384, n * 128, kb);                              // G2(1)
        if (t + 1 < t_end) {
          // tile t's last reader of z (G3 static, z part) precedes G2(1) in the tensor pipe: fetch the next tile's z now
          mbar_wait(&az_empty[0], (uint32_t)(t - t_begin) & 1);
          load_z(t + 1);
        }
#pragma unroll
        for (int n = 0; n < 3; ++n)
#pragma unroll
          for (int kb = 4; kb < 6; ++kb) stage(a.W2, 384, n * 128, kb);                              // G2(2)
#pragma unroll
        for (int kb = 8; kb < 10; ++kb) stage(a.W3cat, 128, 0, kb);                                  // G3 static, n_j part
#pragma unroll
        for (int kb = 0; kb < 6; ++kb) stage(a.W3cat, 128, 0, kb);                                   // G3 partials
        if (t + 1 < t_end && tile_bjb(t + 1) != tile_bjb(t)) {
          mbar_wait(an_empty, (uint32_t)(t - t_begin) & 1);  // the n_j image changes: its last reader (G3 static, n_j part) has completed
          load_n(t + 1);
        }
      }
    }
  } else if (warp == ET_WW + 2) {
    // ============================ epilogue-vector prefetcher + output store ============================
    // iteration n: fetch the vectors of tile n (one tile ahead of the workers), then wait until the workers have staged tile n-1's
    // output in BUF[1], bulk-store it and hand BUF[1] back once the copy engine has read it
    const long long ntl = t_end - t_begin;
    for (long long n = 0; n <= ntl && ntl > 0; ++n) {
      if (n < ntl) {
        const long long t = t_begin + n;
        const uint32_t buf = (uint32_t)n & 1;
        mbar_wait(&vec_free[buf], (((uint32_t)n >> 1) & 1) ^ 1);
        int jb, bsamp;
        const long long m = tile_mb(t, jb, bsamp);
        for (int k = lane; k < 384; k += 32) Ui_s[buf * 384 + k] = a.Ui[m * 384 + k];
        for (int k = lane; k < 128; k += 32) Pf_s[buf * 128 + k] = a.Pf[m * 128 + k];
        const float mi = a.mask[m];
        for (int k = lane; k < 128; k += 32) {
          const int j = jb * 128 + k;
          mk_s[buf * 128 + k] = j < a.N ? mi * a.mask[(long long)bsamp * a.N + j] : 0.f;
        }
        mbar_arrive(&vec_full[buf]);
      }
      if (n >= 1) {
        const long long t = t_begin + n - 1;
        mbar_wait(stg_full, (uint32_t)(n - 1) & 1);
        if (lane == 0) {
          int jb;
          const long long m = tile_m(t, jb);
          uint8_t* dst = reinterpret_cast<uint8_t*>(a.z_out) + ((m * a.JB + jb) * (long long)ET_TILE_BYTES);
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(BUF + ET_TILE_BYTES)), "r"(ET_TILE_BYTES)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          mbar_arrive(&buf_free[1]);
        }
        __syncwarp();
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else if (warp == ET_WW) {
    // ============================ MMA issuer ============================
    if (elect_one()) {
      const uint32_t idesc = (a.exp & 2) ? make_idesc_f16(128, 16) : make_idesc_f16(128, 128);
      uint32_t ds_e = 0, bf[2] = {0, 0}, an_f = 0;
      uint32_t wit = 0;  // weight stage counter of the tile (see ET_KB_PER_TILE: stage index and parity of every use are compile-time constants)
      auto gemm_kb = [&](uint32_t a_addr, uint32_t d_col, bool first_acc) {
        // one k-block: wait the weight stage, 4 x (128x128x16) MMAs, release the stage
        const int s = wit % ET_WSTAGES;
        mbar_wait(&w_full[s], (wit / ET_WSTAGES) & 1);
        tc_fence_after();
        const uint32_t b_addr = smem_u32(WST + s * ET_STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(d_col, make_sw128_desc(a_addr + k * 32), make_sw128_desc(b_addr + k * 32), idesc, (first_acc && k == 0) ? 0u : 1u);
        umma_commit(&w_empty[s]);
        ++wit;
      };
      for (long long t = t_begin; t < t_end; ++t) {
        wit = 0;
        mbar_wait(&az_full[0], (uint32_t)(t - t_begin) & 1);
        if (t == t_begin || tile_bjb(t) != tile_bjb(t - 1)) {
          mbar_wait(an_full, an_f & 1);
          ++an_f;
        }
        tc_fence_after();
        ET_TS(0);
        const uint32_t az = smem_u32(A0z), an = smem_u32(A0n);
        const uint32_t bufa[2] = {smem_u32(BUF), smem_u32(BUF + ET_TILE_BYTES)};
        auto a0_kb = [&](int kb) { return kb < 2 ? az + kb * 16384 : an + (kb - 2) * 16384; };
        auto wait_sa = [&]() {  // DS drained by the workers (E1(0) / E1(2) / E3 hold its content in registers)
          mbar_wait(ds_empty, (ds_e & 1) ^ 1);
          ++ds_e;
          tc_fence_after();
        };
        auto G2 = [&](int c) {
          const int b = c & 1;
          mbar_wait(&buf_full[b], bf[b] & 1);
          ++bf[b];
          tc_fence_after();
#pragma unroll
          for (int n = 0; n < 3; ++n) {
            if (c == 0 && n == 2) {  // D2[256,384) staged G1(1)'s output: wait until E1(1) has drained it
              mbar_wait(sb_empty, (uint32_t)(t - t_begin) & 1);
              tc_fence_after();
            }
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) gemm_kb(bufa[b] + kb * 16384, D2 + n * 128, c == 0 && kb == 0);
            if (c == 2) umma_commit(&d2_full[n]);  // E2(n) starts while the later output chunks are still accumulating
          }
          umma_commit(&buf_free[b]);
        };
        // All three chunks of GEMM 1 are issued back to back: chunk 0 into DS, chunk 1 into D2's columns [256,384) -- idle between the
        // previous tile's G3 partial products (in order in the tensor pipe, so no barrier) and this tile's G2(0) -- chunk 2 into DS again
        // once E1(0) holds chunk 0 in registers.  Chunks 0 and 1 therefore run while the workers are still in the previous tile's E3.
        wait_sa();
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) gemm_kb(a0_kb(kb), DS, kb == 0);
        umma_commit(ds_full);
        ET_TS(1);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) gemm_kb(a0_kb(kb), SB, kb == 0);
        umma_commit(sb_full);
        ET_TS(2);
        wait_sa();
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) gemm_kb(a0_kb(kb), DS, kb == 0);
        umma_commit(ds_full);
        ET_TS(4);
        G2(0);
        ET_TS(3);
        // G3 static, z part: z . W3cat[:, 384:512]^T -> DS (E1(2) holds chunk 2 in registers); releases the z tile for the next fetch
        wait_sa();
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) gemm_kb(a0_kb(kb), DS, kb == 0);
        umma_commit(&az_empty[0]);
        G2(1);
        ET_TS(5);
        G2(2);
        ET_TS(6);
        // G3 static, n_j part (fills the tensor pipe while E2(0) runs)
#pragma unroll
        for (int kb = 2; kb < 4; ++kb) gemm_kb(a0_kb(kb), DS, false);
        umma_commit(an_empty);
        ET_TS(7);
        // G3 partial products: A = r2 chunk c straight from tensor memory (the workers wrote it in place over D2: the 16 K-elements of
        // k-step j of the chunk sit packed in the 8 columns D2 + 128 c + ET_GC (16 j / ET_GC) + (16 j % ET_GC) / 2), B = W3cat k-block
        // 2c + kb from the weight ring
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          mbar_wait(&r2_full[c], (uint32_t)(t - t_begin) & 1);
          tc_fence_after();
          ET_TS(11 + c);
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const int s = wit % ET_WSTAGES;
            mbar_wait(&w_full[s], (wit / ET_WSTAGES) & 1);
            tc_fence_after();
            const uint32_t b_addr = smem_u32(WST + s * ET_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int kk = 64 * kb + 16 * k;  // first K element of the step within the chunk
              umma_f16_ts(DS, D2 + c * 128 + ET_GC * (kk / ET_GC) + (kk % ET_GC) / 2, make_sw128_desc(b_addr + k * 32), idesc, 1u);
            }
            umma_commit(&w_empty[s]);
            ++wit;
          }
          ET_TS(8 + c);
        }
        umma_commit(ds_full);
      }
    }
  } else {
    // ============================ epilogue workers (ET_GROUPS groups x 128 threads) ============================
    // thread <-> tile row <-> TMEM lane; group g owns columns [ET_GC g, ET_GC g + ET_GC) of every 128-column chunk
    const int wg = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const int cg = wg * ET_GC;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t ds_f = 0, fr[2] = {0, 0};
    auto wait_free = [&](int b) {
      mbar_wait(&buf_free[b], (fr[b] & 1) ^ 1);
      ++fr[b];
    };
    // 32 values -> fp16 -> columns [cg, cg + 32) of row `row` of a swizzled chunk buffer (k-block cg / 64, four 16-byte chunks)
    auto store_part = [&](uint8_t* buf, const float* v /*[32]*/) {
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
    for (long long t = t_begin; t < t_end; ++t) {
      int jb, bsamp;
      const long long m = tile_mb(t, jb, bsamp);
      // per-tile epilogue vectors (same i for the whole tile), prefetched by the vector warp
      const uint32_t vn = (uint32_t)(t - t_begin), vbuf = vn & 1;
      const float* Ui_t = Ui_s + vbuf * 384;
      const float* Pf_t = Pf_s + vbuf * 128;
      mbar_wait(&vec_full[vbuf], (vn >> 1) & 1);
      const float mk = mk_s[vbuf * 128 + row];
      float v[ET_GC];
      if (threadIdx.x == 0) ET_TS(16);
      // ---- E1: three chunks of h1
      for (int c = 0; c < 3; ++c) {
        if (c == 1) {
          mbar_wait(sb_full, vn & 1);
          tc_fence_after();
          if (threadIdx.x == 0) ET_TS(17 + 3 * c);
          load_part(SB, v);
          tc_fence_before();
          mbar_arrive(sb_empty);
        } else {
          mbar_wait(ds_full, ds_f & 1);
          ++ds_f;
          tc_fence_after();
          if (threadIdx.x == 0) ET_TS(17 + 3 * c);
          load_part(DS, v);
          tc_fence_before();
          mbar_arrive(ds_empty);
        }
#pragma unroll
        for (int n = 0; n < ET_GC; ++n) v[n] = fmaxf(v[n] + Ui_t[c * 128 + cg + n], 0.f);
        if (threadIdx.x == 0) ET_TS(18 + 3 * c);
        wait_free(c & 1);
        if (!(a.exp & 4)) store_part(BUF + (c & 1) * ET_TILE_BYTES, v);
        fence_proxy_async();
        mbar_arrive(&buf_full[c & 1]);
        if (threadIdx.x == 0) ET_TS(19 + 3 * c);
      }
      // ---- E2: three chunks of r2, written back in place (this thread's 32 fp32 columns become 16 packed fp16 columns)
      for (int c = 0; c < 3; ++c) {
        mbar_wait(&d2_full[c], vn & 1);
        tc_fence_after();
        if (threadIdx.x == 0 && c == 0) ET_TS(26);
        load_part(D2 + c * 128, v);
        uint32_t pk[ET_GC / 2];
#pragma unroll
        for (int n = 0; n < ET_GC / 2; ++n)
          pk[n] = pack_half2(fmaxf(v[2 * n] + b2_s[c * 128 + cg + 2 * n], 0.f), fmaxf(v[2 * n + 1] + b2_s[c * 128 + cg + 2 * n + 1], 0.f));
        tmem_st16(D2 + lane_base + c * 128 + cg, pk);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&r2_full[c]);
        if (threadIdx.x == 0) ET_TS(27 + c);
      }
      // ---- E3: LayerNorm + mask -> fp16 tile image -> bulk store staged in BUF[1] (free: its last reader, G2(1), completed long
      //      ago).  DS is released as soon as each thread holds its columns; the row statistics of the groups are combined through
      //      shared memory (one named barrier among the workers).
      mbar_wait(ds_full, ds_f & 1);
      ++ds_f;
      tc_fence_after();
      if (threadIdx.x == 0) ET_TS(30);
      load_part(DS, v);
      tc_fence_before();
      mbar_arrive(ds_empty);  // D3 is in registers: the next tile's first GEMM may overwrite DS while the LayerNorm runs
      if (threadIdx.x == 0) ET_TS(33);
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int n = 0; n < ET_GC; n += 2) {
        v[n] += Pf_t[cg + n];
        v[n + 1] += Pf_t[cg + n + 1];
        s0 += v[n];
        s1 += v[n + 1];
      }
      const float mh = (s0 + s1) * (1.f / ET_GC);
      float q0 = 0.f, q1 = 0.f;
#pragma unroll
      for (int n = 0; n < ET_GC; n += 2) {
        const float d0 = v[n] - mh, d1 = v[n + 1] - mh;
        q0 += d0 * d0;
        q1 += d1 * d1;
      }
      // exchange (mean, sum of squared deviations) of the row parts and combine them with the pairwise update (Chan et al.)
      red_s[wg * 128 + row] = mh;
      red_q[wg * 128 + row] = q0 + q1;
      if (threadIdx.x == 0) ET_TS(34);
      asm volatile("bar.sync 1, %0;" ::"n"(ET_WORKERS) : "memory");
      if (threadIdx.x == 0) ET_TS(35);
      float mean = 0.f, m2 = 0.f;
#pragma unroll
      for (int g = 0; g < ET_GROUPS; ++g) mean += red_s[g * 128 + row];
      mean *= (1.f / ET_GROUPS);
#pragma unroll
      for (int g = 0; g < ET_GROUPS; ++g) {
        const float dg = red_s[g * 128 + row] - mean;
        m2 += red_q[g * 128 + row] + (float)ET_GC * dg * dg;
      }
      const float rstd = rsqrtf(m2 * (1.f / 128.f) + 1e-5f);
#pragma unroll
      for (int n = 0; n < ET_GC; ++n) v[n] = ((v[n] - mean) * rstd * g_s[cg + n] + be_s[cg + n]) * mk;
      mbar_arrive(&vec_free[vbuf]);
      if (threadIdx.x == 0) ET_TS(31);
      wait_free(1);  // BUF[1]'s last reader was G2(1) (h1 chunk 1); its release is consumed here (already complete: D3 is)
      if (!(a.exp & 4)) store_part(BUF + ET_TILE_BYTES, v);
      fence_proxy_async();
      mbar_arrive(stg_full);  // the vector warp issues the bulk store of the staged tile
      if (threadIdx.x == 0) ET_TS(32);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == ET_WW) tmem_dealloc(tmem_base, 512);
}

inline size_t et_smem_bytes() {
  return 1024 + 4 * (size_t)ET_TILE_BYTES + ET_WSTAGES * ET_STAGE_BYTES + 44 * 8 + 16 + (2 * 384 + 2 * 128 + 384 + 128 + 128 + 2 * ET_GROUPS * 128 + 256) * 4 + 64 + 32;
}

// ---- layout helpers --------------------------------------------------------------------------------------------------
// fp32 z[B,N,N,128] -> fp16 tile images [B][N][JB][2 kb][128 rows][128 B swizzled]; rows j >= N are zero.
__global__ void z_to_image_kernel(int B, int N, int JB, const float* __restrict__ z, __half* __restrict__ img) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte chunk
  const long long total = (long long)B * N * JB * 128 * 16;
  if (idx >= total) return;
  const int kc = (int)(idx & 15);
  const int r = (int)((idx >> 4) & 127);
  const long long tile = idx >> 11;
  const int jb = (int)(tile % JB);
  const long long m = tile / JB;
  const int j = jb * 128 + r;
  uint4 u = make_uint4(0, 0, 0, 0);
  if (j < N) {
    const float4* src = reinterpret_cast<const float4*>(z + (m * N + j) * 128 + kc * 8);
    const float4 x0 = __ldg(src), x1 = __ldg(src + 1);
    u = make_uint4(pack_half2(x0.x, x0.y), pack_half2(x0.z, x0.w), pack_half2(x1.x, x1.y), pack_half2(x1.z, x1.w));
  }
  uint8_t* dst = reinterpret_cast<uint8_t*>(img) + tile * ET_TILE_BYTES + (kc >> 3) * 16384 + sw128_chunk_off(r, kc & 7);
  *reinterpret_cast<uint4*>(dst) = u;
}

__global__ void image_to_z_kernel(int B, int N, int JB, const __half* __restrict__ img, float* __restrict__ z) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * N * JB * 128 * 16;
  if (idx >= total) return;
  const int kc = (int)(idx & 15);
  const int r = (int)((idx >> 4) & 127);
  const long long tile = idx >> 11;
  const int jb = (int)(tile % JB);
  const long long m = tile / JB;
  const int j = jb * 128 + r;
  if (j >= N) return;
  const uint8_t* src = reinterpret_cast<const uint8_t*>(img) + tile * ET_TILE_BYTES + (kc >> 3) * 16384 + sw128_chunk_off(r, kc & 7);
  const uint4 u = *reinterpret_cast<const uint4*>(src);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
  float* dst = z + (m * N + j) * 128 + kc * 8;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 f = __half22float2(h[e]);
    dst[2 * e] = f.x;
    dst[2 * e + 1] = f.y;
  }
}

// n_emb [B*N,128] fp32 -> per-(b, j-block) fp16 images [B][JB][2 kb][128 rows][128 B]
__global__ void n_to_image_kernel(int B, int N, int JB, const float* __restrict__ n_emb, __half* __restrict__ img) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * JB * 128 * 16;
  if (idx >= total) return;
  const int kc = (int)(idx & 15);
  const int r = (int)((idx >> 4) & 127);
  const long long tile = idx >> 11;  // b*JB + jb
  const int jb = (int)(tile % JB);
  const long long b = tile / JB;
  const int j = jb * 128 + r;
  uint4 u = make_uint4(0, 0, 0, 0);
  if (j < N) {
    const float4* src = reinterpret_cast<const float4*>(n_emb + (b * N + j) * 128 + kc * 8);
    const float4 x0 = __ldg(src), x1 = __ldg(src + 1);
    u = make_uint4(pack_half2(x0.x, x0.y), pack_half2(x0.z, x0.w), pack_half2(x1.x, x1.y), pack_half2(x1.z, x1.w));
  }
  uint8_t* dst = reinterpret_cast<uint8_t*>(img) + tile * ET_TILE_BYTES + (kc >> 3) * 16384 + sw128_chunk_off(r, kc & 7);
  *reinterpret_cast<uint4*>(dst) = u;
}

}  // namespace tc
}  // namespace fdpt
