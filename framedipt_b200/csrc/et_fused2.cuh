// et_fused2.cuh — the EdgeTransition kernel of et_fused.cuh on CTA PAIRS (tcgen05 cta_group::2).
//
// et_fused_kernel is bound by shared-memory bandwidth: every 128x128x16 SS-mode MMA re-reads its A and B operand slices from the
// SM's shared memory, and the weight ring is written at the same time (profiles/r01_ncu_full_et_fused_kernel.txt).  Here two CTAs
// of a cluster (one TPC) each own one 128-pair tile and share every weight stage: the pair executes 256x128x16 MMAs whose B operand
// (128 weight rows) is split 64 / 64 between the two CTAs' shared memories.  Per CTA and tile that halves the bulk-copy writes of
// the weight stream (320 KB -> 160 KB) and the B-operand reads, and the freed space makes the weight ring 7 half-stages deep.
//
// Protocol (differences from et_fused.cuh; the GEMM chain, the TMEM map and the worker code are the same):
//   * only the leader CTA (cluster rank 0) issues MMAs; every tcgen05.commit is multicast to the same barrier in both CTAs, so all
//     "MMA -> somebody" barriers (ds_full, d2_full, buf_free, az_empty, w_empty) stay CTA-local for their waiters;
//   * "worker -> MMA" barriers (buf_full, ds_empty, d2_empty) live in the leader only: lane 0 of each of the 16 worker warps of the
//     pair arrives there (remote mbarrier.arrive for the peer), count 16;
//   * bulk copies can only signal a barrier of the CTA they write to, so lane 0 of the peer's (otherwise idle) MMA warp relays:
//     it waits on the peer's own az_full / an_full / w_full in consumption order and forwards one arrive to the leader's
//     *_peer barrier; the leader waits on its own and on the forwarded barrier before issuing.
//   * tiles are taken in pairs (2p, 2p+1); with an odd tile count the peer recomputes the last tile and skips the store.
#pragma once
#include "et_fused.cuh"

namespace fdpt {
namespace tc {

constexpr int ET2_WSTAGES = 7;
constexpr int ET2_HALF_BYTES = 8192;  // 64 weight rows x one k-block

FDPT_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
FDPT_DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
FDPT_DEVINL uint32_t mapa_u32(uint32_t cta_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
  return r;
}
FDPT_DEVINL void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default .release.cta semantics, as cutlass::arch::ClusterBarrier::arrive(cta_id) does: a .release.cluster arrive compiles to
  // MEMBAR.ALL.GPU + SYNCS.ARRIVE and costs ~1000 cycles per signal (measured: it serialised the whole weight pipeline)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
FDPT_DEVINL bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
FDPT_DEVINL void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) mbar_timeout(smem_u32(bar) | 0x800000u, parity);
  }
}
FDPT_DEVINL void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {  // the same warp of both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
FDPT_DEVINL void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
FDPT_DEVINL void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued pair MMAs arrive on the barrier at this shared-memory offset in BOTH CTAs when complete
FDPT_DEVINL void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

#define ET2_TS(id)                                                                      \
  do {                                                                                  \
    if (a.dbg && blockIdx.x == 0 && it < 8) a.dbg[it * 48 + (id)] = clock64();          \
  } while (0)

__global__ void __launch_bounds__(ET_THREADS, 1) et_fused2_kernel(EtArgs a, int flags) {  // flags: profiling experiments (1: no weight waits, 2: no MMAs, 4: every MMA twice)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* A0z = smem;                                 // 2 x 32 KB
  uint8_t* A0n = A0z + 2 * ET_TILE_BYTES;              // 32 KB
  uint8_t* BUF = A0n + ET_TILE_BYTES;                  // 2 x 32 KB
  uint8_t* WST = BUF + 2 * ET_TILE_BYTES;              // ET2_WSTAGES x 8 KB (this CTA's 64 rows of every weight stage)
  uint64_t* bars = reinterpret_cast<uint64_t*>(WST + ET2_WSTAGES * ET2_HALF_BYTES);
  uint64_t* w_full = bars;                    // [ET2_WSTAGES] local bulk copy landed
  uint64_t* w_peer = w_full + ET2_WSTAGES;    // [ET2_WSTAGES] (leader) the peer's half landed
  uint64_t* w_empty = w_peer + ET2_WSTAGES;   // [ET2_WSTAGES]
  uint64_t* az_full = w_empty + ET2_WSTAGES;  // [2]
  uint64_t* az_peer = az_full + 2;            // [2] (leader)
  uint64_t* az_empty = az_peer + 2;           // [2]
  uint64_t* an_full = az_empty + 2;           // [1]
  uint64_t* an_peer = an_full + 1;            // [1] (leader)
  uint64_t* ds_full = an_peer + 1;            // [1]
  uint64_t* ds_empty = ds_full + 1;           // [1] (leader) 16 warp arrivals
  uint64_t* buf_full = ds_empty + 1;          // [2] (leader) 16 warp arrivals
  uint64_t* buf_free = buf_full + 2;          // [2]
  uint64_t* d2_full = buf_free + 2;           // [1]
  uint64_t* d2_empty = d2_full + 1;           // [1] (leader) 16 warp arrivals
  uint64_t* vec_full = d2_empty + 1;          // [2]
  uint64_t* vec_free = vec_full + 2;          // [2]
  uint64_t* stg_full = vec_free + 2;          // [1]
  uint64_t* d3_full = stg_full + 1;           // [1] D3 (in D2 columns [0,128)) complete
  uint64_t* d2c0_empty = d3_full + 1;         // [1] (leader) 16 warp arrivals: E2 has drained D2 columns [0,128)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d2c0_empty + 1);
  float* Ui_s = reinterpret_cast<float*>(tmem_slot + 4);  // [2][384]
  float* Pf_s = Ui_s + 2 * 384;                            // [2][128]
  float* b2_s = Pf_s + 2 * 128;                            // [384]
  float* g_s = b2_s + 384;                                 // [128]
  float* be_s = g_s + 128;                                 // [128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const unsigned n_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
  const long long pairs = (a.tiles + 1) >> 1;
  const long long per = (pairs + n_clusters - 1) / n_clusters;
  const long long p_begin = (long long)cluster_id * per;
  const long long p_end = min(pairs, p_begin + per);
  // tile of CTA `r` in pair p (the last pair of an odd tile count repeats the last tile on the peer)
  auto tile_of = [&](long long p, uint32_t r) -> long long { return min(2 * p + (long long)r, a.tiles - 1); };

  if (threadIdx.x == 0) {
    for (int s = 0; s < ET2_WSTAGES; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_peer[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&az_full[s], 1);
      mbar_init(&az_peer[s], 1);
      mbar_init(&az_empty[s], 1);
      mbar_init(&buf_full[s], 16);
      mbar_init(&buf_free[s], 1);
      mbar_init(&vec_full[s], 32);
      mbar_init(&vec_free[s], ET_WORKERS);
    }
    mbar_init(an_full, 1);
    mbar_init(an_peer, 1);
    mbar_init(ds_full, 1);
    mbar_init(ds_empty, 16);
    mbar_init(d2_full, 1);
    mbar_init(d2_empty, 16);
    mbar_init(stg_full, ET_WORKERS);
    mbar_init(d3_full, 1);
    mbar_init(d2c0_empty, 16);
    fence_barrier_init();
  }
  for (int k = threadIdx.x; k < 384; k += blockDim.x) b2_s[k] = a.b2[k];
  for (int k = threadIdx.x; k < 128; k += blockDim.x) {
    g_s[k] = a.ln_g[k];
    be_s[k] = a.ln_b[k];
  }
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers exist before anything remote can touch them
  if (warp == 8) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both halves of the pair's tensor memory are allocated before the leader's first MMA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t D2 = tmem_base, DS = tmem_base + 384;

  auto tile_bjb = [&](long long t) -> long long { return (long long)((unsigned)t / (unsigned)a.N); };
  auto tile_mb = [&](long long t, int& jb, int& b) -> long long {
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
  // does CTA r load a new n_j image for pair p?
  auto n_changes = [&](long long p, uint32_t r) -> bool { return p == p_begin || tile_bjb(tile_of(p, r)) != tile_bjb(tile_of(p - 1, r)); };

  if (warp == 9) {
    // ============================ loader (both CTAs: own z tile, own n_j image, own half of every weight stage) ============
    if (lane == 0 && p_begin < p_end) {
      uint32_t wit = 0;
      auto load_z = [&](long long p) {
        const uint32_t n = (uint32_t)(p - p_begin);
        const int s = (int)(n & 1);
        mbar_wait_diag(&az_empty[s], ((n >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&az_full[s], ET_TILE_BYTES);
        int jb;
        const long long m = tile_m(tile_of(p, rank), jb);
        bulk_g2s(A0z + s * ET_TILE_BYTES, reinterpret_cast<const uint8_t*>(a.z_in) + ((m * a.JB + jb) * (long long)ET_TILE_BYTES), ET_TILE_BYTES,
                 &az_full[s]);
      };
      auto load_n = [&](long long p) {
        mbar_arrive_expect_tx(an_full, ET_TILE_BYTES);
        bulk_g2s(A0n, reinterpret_cast<const uint8_t*>(a.n_img) + tile_bjb(tile_of(p, rank)) * (long long)ET_TILE_BYTES, ET_TILE_BYTES, an_full);
      };
      auto stage = [&](const __half* img, int rows_total, int row0, int kb) {
        const int s = wit % ET2_WSTAGES;
        mbar_wait_diag(&w_empty[s], ((wit / ET2_WSTAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&w_full[s], ET2_HALF_BYTES);
        bulk_g2s(WST + s * ET2_HALF_BYTES, reinterpret_cast<const uint8_t*>(img) + ((size_t)kb * rows_total + row0 + 64 * rank) * 128,
                 ET2_HALF_BYTES, &w_full[s]);
        ++wit;
      };
      load_z(p_begin);
      load_n(p_begin);
      for (long long p = p_begin; p < p_end; ++p) {
        const uint32_t it = (uint32_t)(p - p_begin);
        for (int kb = 0; kb < 4; ++kb) {
          stage(a.W1cat, 384, 0, kb);                                                                // G1(0)
          ET2_TS(33 + kb);
        }
        for (int kb = 0; kb < 4; ++kb) stage(a.W1cat, 384, 128, kb);                                 // G1(1)
        if (p + 1 < p_end) load_z(p + 1);
        for (int n = 0; n < 3; ++n) for (int kb = 0; kb < 2; ++kb) stage(a.W2, 384, n * 128, kb);      // G2(0)
        for (int kb = 0; kb < 4; ++kb) stage(a.W1cat, 384, 256, kb);                                 // G1(2)
        for (int n = 0; n < 3; ++n) for (int kb = 2; kb < 4; ++kb) stage(a.W2, 384, n * 128, kb);      // G2(1)
        for (int n = 0; n < 3; ++n) for (int kb = 4; kb < 6; ++kb) stage(a.W2, 384, n * 128, kb);      // G2(2)
        for (int kb = 6; kb < 10; ++kb) stage(a.W3cat, 128, 0, kb);                                  // G3 static
        for (int kb = 0; kb < 6; ++kb) stage(a.W3cat, 128, 0, kb);                                   // G3 partials
        if (p + 1 < p_end && n_changes(p + 1, rank)) {
          // the n_j image changes: wait until this pair's last reader (G3 static) has completed
          const uint32_t n = (uint32_t)(p - p_begin);
          mbar_wait_diag(&az_empty[n & 1], (n >> 1) & 1);
          load_n(p + 1);
        }
      }
    }
  } else if (warp == 10) {
    // ============================ epilogue-vector prefetcher ============================
    for (long long p = p_begin; p < p_end; ++p) {
      const uint32_t n = (uint32_t)(p - p_begin), buf = n & 1;
      mbar_wait_diag(&vec_free[buf], ((n >> 1) & 1) ^ 1);
      int jb;
      const long long m = tile_m(tile_of(p, rank), jb);
      for (int k = lane; k < 384; k += 32) Ui_s[buf * 384 + k] = a.Ui[m * 384 + k];
      for (int k = lane; k < 128; k += 32) Pf_s[buf * 128 + k] = a.Pf[m * 128 + k];
      mbar_arrive(&vec_full[buf]);
    }
  } else if (warp == 8) {
    if (lane == 0 && !leader) {
      // ============================ peer: relay of the bulk-copy barriers to the leader ============================
      uint32_t wit = 0, an_f = 0;
      const uint32_t r_w = mapa_u32(smem_u32(w_peer), 0), r_az = mapa_u32(smem_u32(az_peer), 0), r_an = mapa_u32(smem_u32(an_peer), 0);
      for (long long p = p_begin; p < p_end; ++p) {
        const uint32_t n = (uint32_t)(p - p_begin);
        mbar_wait_diag(&az_full[n & 1], (n >> 1) & 1);
        mbar_arrive_cluster(r_az + (n & 1) * 8);
        if (n_changes(p, 1)) {
          mbar_wait_diag(an_full, an_f & 1);
          ++an_f;
          mbar_arrive_cluster(r_an);
        }
        for (int k = 0; k < 40; ++k) {  // 40 weight stages per tile (3 x 4 + 3 x 6 + 4 + 6)
          const int s = wit % ET2_WSTAGES;
          mbar_wait_diag(&w_full[s], (wit / ET2_WSTAGES) & 1);
          mbar_arrive_cluster(r_w + s * 8);
          ++wit;
        }
      }
    } else if (lane == 0) {
      // ============================ leader: MMA issuer for the pair ============================
      const uint32_t idesc = make_idesc_f16(256, 128);
      uint32_t wit = 0, ds_e = 0, bf[2] = {0, 0}, d2_e = 0, an_f = 0, an_pf = 0;
      int ts_slot = -1;
      uint32_t ts_it = 0;
      const bool prof = a.dbg != nullptr && blockIdx.x == 0;  // timeline / wait accounting of cluster 0 only
      long long ti = 0, tcm = 0;  // cycles spent issuing MMAs / commits
      long long wl = 0, wp = 0;  // cycles spent waiting for the local / the peer's half of the weight stages
      auto gemm_kb = [&](uint32_t a_addr, uint32_t d_col, bool first_acc) {
        const int s = wit % ET2_WSTAGES;
        const uint32_t ph = (wit / ET2_WSTAGES) & 1;
        if (!(flags & 1)) {
          const long long c0 = prof ? clock64() : 0;
          mbar_wait_diag(&w_full[s], ph);
          const long long c1 = prof ? clock64() : 0;
          mbar_wait_cluster(&w_peer[s], ph);
          const long long c2 = prof ? clock64() : 0;
          wl += c1 - c0;
          wp += c2 - c1;
          if (ts_slot >= 0 && a.dbg && blockIdx.x == 0 && ts_it < 8) {
            a.dbg[ts_it * 48 + ts_slot] = c1;
            a.dbg[ts_it * 48 + ts_slot + 1] = c2;
          }
        }
        tc_fence_after();
        const uint32_t b_addr = smem_u32(WST + s * ET2_HALF_BYTES);
        const long long i0 = prof ? clock64() : 0;
        const int reps = (flags & 2) ? 0 : (flags & 4) ? 2 : 1;  // profiling experiments: no MMAs / every MMA twice
        for (int r = 0; r < reps; ++r) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_pair(d_col, make_sw128_desc(a_addr + k * 32), make_sw128_desc(b_addr + k * 32), idesc, (first_acc && k == 0 && r == 0) ? 0u : 1u);
        }
        const long long i1 = prof ? clock64() : 0;
        umma_commit_pair(&w_empty[s]);
        tcm += (prof ? clock64() : 0) - i1;
        ti += i1 - i0;
        ++wit;
      };
      for (long long p = p_begin; p < p_end; ++p) {
        const uint32_t it = (uint32_t)(p - p_begin);
        const int zs = (int)(it & 1);
        const uint32_t zuse = it >> 1;
        mbar_wait_diag(&az_full[zs], zuse & 1);
        mbar_wait_cluster(&az_peer[zs], zuse & 1);
        if (n_changes(p, 0)) {
          mbar_wait_diag(an_full, an_f & 1);
          ++an_f;
        }
        if (n_changes(p, 1)) {
          mbar_wait_cluster(an_peer, an_pf & 1);
          ++an_pf;
        }
        tc_fence_after();
        ET2_TS(0);
        const uint32_t az = smem_u32(A0z + zs * ET_TILE_BYTES), an = smem_u32(A0n);
        const uint32_t bufa[2] = {smem_u32(BUF), smem_u32(BUF + ET_TILE_BYTES)};
        auto a0_kb = [&](int kb) { return kb < 2 ? az + kb * 16384 : an + (kb - 2) * 16384; };
        auto G1 = [&](int c) {
          mbar_wait_cluster(ds_empty, (ds_e & 1) ^ 1);
          ++ds_e;
          tc_fence_after();
          if (c == 0) ET2_TS(14);
          for (int kb = 0; kb < 4; ++kb) {
            ts_it = it;
            ts_slot = (c == 0 && kb < 3) ? 42 + 2 * kb : -1;
            gemm_kb(a0_kb(kb), DS, kb == 0);
            ts_slot = -1;
            if (c == 0 && kb == 0) ET2_TS(15);
            if (c == 0 && kb > 0) ET2_TS(36 + kb);
          }
          umma_commit_pair(ds_full);
        };
        auto G2 = [&](int c) {
          const int b = c & 1;
          mbar_wait_cluster(&buf_full[b], bf[b] & 1);
          ++bf[b];
          if (c == 0) {
            mbar_wait_cluster(d2_empty, (d2_e & 1) ^ 1);
            ++d2_e;
          }
          tc_fence_after();
          for (int n = 0; n < 3; ++n)
            for (int kb = 0; kb < 2; ++kb) gemm_kb(bufa[b] + kb * 16384, D2 + n * 128, c == 0 && kb == 0);
          umma_commit_pair(&buf_free[b]);
          if (c == 2) umma_commit_pair(d2_full);
        };
        G1(0);
        ET2_TS(1);
        G1(1);
        ET2_TS(2);
        G2(0);
        ET2_TS(3);
        G1(2);
        ET2_TS(4);
        G2(1);
        ET2_TS(5);
        G2(2);
        ET2_TS(6);
        // G3 static part -> D3 = D2 columns [0,128) once E2(0) has drained them; DS stays free, so the next tile's G1(0) is issued
        // right behind this tile's G3 and runs while the workers are in E3
        mbar_wait_cluster(d2c0_empty, it & 1);
        tc_fence_after();
        for (int kb = 0; kb < 4; ++kb) gemm_kb(a0_kb(kb), D2, kb == 0);
        umma_commit_pair(&az_empty[zs]);
        ET2_TS(7);
        for (int c = 0; c < 3; ++c) {
          const int b = c & 1;
          mbar_wait_cluster(&buf_full[b], bf[b] & 1);
          ++bf[b];
          tc_fence_after();
          if (c < 1) ET2_TS(11 + c);
          for (int kb = 0; kb < 2; ++kb) gemm_kb(bufa[b] + kb * 16384, D2, false);
          if (c != 1) umma_commit_pair(&buf_free[b]);  // BUF[1] becomes the output staging buffer: released by worker thread 0
          ET2_TS(8 + c);
        }
        umma_commit_pair(d3_full);
        if (a.dbg && blockIdx.x == 0 && it < 8) {
          a.dbg[it * 48 + 40] = wl;
          a.dbg[it * 48 + 41] = wp;
          a.dbg[it * 48 + 13] = ti;
          a.dbg[it * 48 + 12] = tcm;
        }
      }
    }
  } else {
    // ============================ epilogue workers (2 groups x 128 threads per CTA) ============================
    const int wg = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const int cg = wg * 64;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t l_ds_empty = mapa_u32(smem_u32(ds_empty), 0), l_buf_full = mapa_u32(smem_u32(buf_full), 0),
                   l_d2_empty = mapa_u32(smem_u32(d2_empty), 0), l_d2c0_empty = mapa_u32(smem_u32(d2c0_empty), 0);
    uint32_t ds_f = 0, fr[2] = {0, 0}, d2_f = 0;
    auto warp_arrive = [&](uint32_t leader_bar) {  // one arrival per warp on a leader barrier
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_bar);
    };
    auto wait_free = [&](int b) {
      mbar_wait_diag(&buf_free[b], (fr[b] & 1) ^ 1);
      ++fr[b];
    };
    auto store_half = [&](uint8_t* buf, const float* v /*[64]*/) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float* q = v + c * 8;
        const uint4 u = make_uint4(pack_half2(q[0], q[1]), pack_half2(q[2], q[3]), pack_half2(q[4], q[5]), pack_half2(q[6], q[7]));
        *reinterpret_cast<uint4*>(buf + wg * 16384 + sw128_chunk_off(row, c)) = u;
      }
    };
    auto load_half = [&](uint32_t taddr, float* v /*[64]*/) {
      tmem_ld32(taddr + lane_base + cg, v);
      tmem_ld32(taddr + lane_base + cg + 32, v + 32);
      tmem_ld_wait();
    };
    for (long long p = p_begin; p < p_end; ++p) {
      const uint32_t it = (uint32_t)(p - p_begin);
      const bool valid = 2 * p + rank < a.tiles;
      int jb, bsamp;
      const long long m = tile_mb(tile_of(p, rank), jb, bsamp);
      const int j = jb * 128 + row;
      const uint32_t vbuf = it & 1;
      const float* Ui_t = Ui_s + vbuf * 384;
      const float* Pf_t = Pf_s + vbuf * 128;
      float mk = 0.f;
      if (j < a.N) mk = a.mask[m] * a.mask[(long long)bsamp * a.N + j];
      mbar_wait_diag(&vec_full[vbuf], (it >> 1) & 1);
      float v[64];
      if (threadIdx.x == 0) ET2_TS(16);
      // ---- E1: three chunks of h1
      for (int c = 0; c < 3; ++c) {
        if (c == 1 && threadIdx.x == 0 && it != 0) {  // BUF[1] staged the previous tile's output: its bulk store has finished reading
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          mbar_arrive(&buf_free[1]);
        }
        mbar_wait_diag(ds_full, ds_f & 1);
        ++ds_f;
        tc_fence_after();
        if (threadIdx.x == 0) ET2_TS(17 + 3 * c);
        load_half(DS, v);
        tc_fence_before();
        warp_arrive(l_ds_empty);
#pragma unroll
        for (int n = 0; n < 64; ++n) v[n] = fmaxf(v[n] + Ui_t[c * 128 + cg + n], 0.f);
        if (threadIdx.x == 0) ET2_TS(18 + 3 * c);
        wait_free(c & 1);
        store_half(BUF + (c & 1) * ET_TILE_BYTES, v);
        fence_proxy_async();
        warp_arrive(l_buf_full + (c & 1) * 8);
        if (threadIdx.x == 0) ET2_TS(19 + 3 * c);
      }
      // ---- E2: three chunks of r2
      mbar_wait_diag(d2_full, d2_f & 1);
      ++d2_f;
      tc_fence_after();
      if (threadIdx.x == 0) ET2_TS(26);
      for (int c = 0; c < 3; ++c) {
        load_half(D2 + c * 128, v);
        if (c == 0) {
          tc_fence_before();
          warp_arrive(l_d2c0_empty);
        }
#pragma unroll
        for (int n = 0; n < 64; ++n) v[n] = fmaxf(v[n] + b2_s[c * 128 + cg + n], 0.f);
        wait_free(c & 1);
        store_half(BUF + (c & 1) * ET_TILE_BYTES, v);
        fence_proxy_async();
        warp_arrive(l_buf_full + (c & 1) * 8);
        if (threadIdx.x == 0) ET2_TS(27 + c);
      }
      // ---- E3: LayerNorm + mask -> fp16 tile image staged in BUF[1] -> bulk store (see et_fused.cuh)
      mbar_wait_diag(d3_full, it & 1);
      tc_fence_after();
      if (threadIdx.x == 0) ET2_TS(30);
      load_half(D2, v);
      float s0 = 0.f;
#pragma unroll
      for (int n = 0; n < 64; ++n) {
        v[n] += Pf_t[cg + n];
        s0 += v[n];
      }
      const float shift = s0 * (1.f / 64.f);
      float sd = 0.f, sq = 0.f;
#pragma unroll
      for (int n = 0; n < 64; ++n) {
        const float d = v[n] - shift;
        sd += d;
        sq += d * d;
      }
      const int og = 64 - cg;
      {
        float w[32];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          tmem_ld32(D2 + lane_base + og + 32 * h, w);
          tmem_ld_wait();
#pragma unroll
          for (int n = 0; n < 32; ++n) {
            const float d = w[n] + Pf_t[og + 32 * h + n] - shift;
            sd += d;
            sq += d * d;
          }
        }
      }
      tc_fence_before();
      warp_arrive(l_d2_empty);  // all of D2 (E2 drained [128,384) earlier) may be overwritten by the next tile's G2(0)
      const float dm = sd * (1.f / 128.f);
      const float mean = shift + dm;
      const float rstd = rsqrtf(fmaxf(sq * (1.f / 128.f) - dm * dm, 0.f) + 1e-5f);
#pragma unroll
      for (int n = 0; n < 64; ++n) v[n] = ((v[n] - mean) * rstd * g_s[cg + n] + be_s[cg + n]) * mk;
      mbar_arrive(&vec_free[vbuf]);
      if (threadIdx.x == 0) ET2_TS(31);
      store_half(BUF + ET_TILE_BYTES, v);
      fence_proxy_async();
      mbar_arrive(stg_full);
      if (threadIdx.x == 0) {
        mbar_wait_diag(stg_full, it & 1);
        if (valid) {
          uint8_t* dst = reinterpret_cast<uint8_t*>(a.z_out) + ((m * a.JB + jb) * (long long)ET_TILE_BYTES);
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(BUF + ET_TILE_BYTES)),
                       "r"(ET_TILE_BYTES)
                       : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        ET2_TS(32);
      }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's shared memory and barriers stay alive until the leader's last MMA / commit has landed
  if (warp == 8) tmem_dealloc_pair(tmem_base, 512);
}

inline size_t et2_smem_bytes() {
  return 1024 + 5 * (size_t)ET_TILE_BYTES + ET2_WSTAGES * ET2_HALF_BYTES + (3 * ET2_WSTAGES + 24) * 8 + 16 +
         (2 * 384 + 2 * 128 + 384 + 128 + 128) * 4 + 64;
}

}  // namespace tc
}  // namespace fdpt
