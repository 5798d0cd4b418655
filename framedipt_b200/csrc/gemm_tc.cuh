// gemm_tc.cuh — fp32-class GEMM on the 5th-gen tensor cores: 3-term split TF32 (tcgen05.mma kind::tf32, fp32 accumulation in TMEM).
//
//   C[M,N] = epi( alpha * A[M,K] @ op(B) )      same GemmArgs / epilogue contract as gemm_simt.cuh
//
// The node side of the network needs fp32-class products (single-pass TF32 fails the 1e-3 A budget, SURVEY §7 hard part 1),
// so every operand x is split on the fly into x_hi = x with the 13 low mantissa bits cleared (exactly representable in TF32)
// and x_lo = x - x_hi (exact in fp32), and   A·B ~= A_lo·B_hi + A_hi·B_lo + A_hi·B_hi   (dropped term ~2^-22 relative).
//
// CTA = one 128 x BN output tile (BN = 128 or 64), 288 threads:
//   warps 0-7  producers: cp.async (LDGSTS) of the raw fp32 A / B k-block (32 fp32 = 128 B per row) straight into the swizzled
//              K-major operand image, two k-blocks ahead; then an in-place split into the hi / lo images (B may also be
//              [K,N] row-major = MN-major operand in the SWIZZLE_128B_BASE32B layout, used by P·V); later the epilogue
//              (tcgen05.ld of both accumulators -> bias / relu / mask / residual -> global).
//   warp 8     MMA issuer (one elected lane) + TMEM owner.
// 5-deep raw/hi ring + 2-deep lo ring (224 KB at BN=128) with full/done mbarriers; the split of k-block kb overlaps the MMAs of kb-1; arbitrary M, N, K (zero-filled edges), row strides and
// two batch strides, so the same kernel serves Linear layers, per-head Q·K^T and P·V.
#pragma once
#include "gemm_simt.cuh"
#include "tc_common.cuh"

namespace fdpt {
namespace tc {

constexpr int GT_STAGES = 5;       // raw/hi ring depth (the lo images have their own 2-deep ring)
constexpr int GT_PRODUCERS = 256;  // producer / epilogue threads (8 warps); warp 8 issues the MMAs
constexpr int GT_THREADS = GT_PRODUCERS + 32;
constexpr int GT_BM = 128;
constexpr int GT_KB = 32;  // fp32 elements per k-block row (128 B)

// Instruction descriptor, kind::tf32: TF32 A/B (format 2), fp32 accumulate; b_mn = 1 -> B operand is MN-major.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
FDPT_DEVINL void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
struct GemmTcArgs {
  GemmArgs g;
  int bn;         // 128 or 64
  int b_kmajor;   // 1: B is [N,K] (weights, K^T);  0: B is [K,N] row-major (P·V)
  int a_vec, b_vec, c_vec;  // 16-byte aligned rows -> float4 path
  int mn_swap;    // bring-up knob: swap LBO / SBO of the MN-major descriptor
};

// ---- cp.async helpers (LDGSTS: global -> shared without register staging; bytes beyond `src_bytes` are zero-filled) ----
FDPT_DEVINL void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
FDPT_DEVINL void cp_async4(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
FDPT_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
FDPT_DEVINL void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// 4 consecutive fp32 of one row -> one 16-byte smem chunk; `valid` = number of in-range elements (<= 0: all zero)
FDPT_DEVINL void copy_chunk(uint32_t dst, const float* row_base, const float* src, int valid, int vec) {
  valid = max(0, min(4, valid));
  if (vec) {
    cp_async16(dst, valid > 0 ? (const void*)src : (const void*)row_base, valid * 4);
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) cp_async4(dst + 4 * e, e < valid ? (const void*)(src + e) : (const void*)row_base, e < valid ? 4 : 0);
  }
}

// in-place split of one raw fp32 chunk: hi (13 low mantissa bits cleared) stays, lo = x - hi goes to the second image
FDPT_DEVINL void split_chunk(uint8_t* hi_img, uint8_t* lo_img, uint32_t off) {
  const float4 v = *reinterpret_cast<const float4*>(hi_img + off);
  float4 h, l;
  h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
  h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
  h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
  h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
  l.x = v.x - h.x;
  l.y = v.y - h.y;
  l.z = v.z - h.z;
  l.w = v.w - h.w;
  *reinterpret_cast<float4*>(hi_img + off) = h;
  *reinterpret_cast<float4*>(lo_img + off) = l;
}

// MN-major tf32 operands only exist in the SWIZZLE_128B_BASE32B layout (layout type 1): atoms of 4 k-rows x 128 B (32 n),
// 32-byte units XOR-ed with (k & 3).  Tile image order: [k-atom][n-atom][512 B]  ->  LBO (n atoms) = 512, SBO (k atoms) = mn_atoms*512.
FDPT_DEVINL uint32_t mn32_chunk_off(int kk, int c, int mn_atoms) {
  const int ka = kk >> 2, kr = kk & 3, na = c >> 3, cc = c & 7;
  return (uint32_t)((ka * mn_atoms + na) * 512 + kr * 128 + ((((cc >> 1) ^ kr) << 5) | ((cc & 1) << 4)));
}
FDPT_DEVINL uint64_t make_mn32_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;  // SWIZZLE_128B_BASE32B
  return d;
}

// per-thread copy plan of one operand: `iters` chunks per k-block, affine in the iteration index
struct ChunkPlan {
  const float* src;        // first chunk of k-block 0
  long long it_stride;     // elements between consecutive iterations
  long long kb_stride;     // elements between consecutive k-blocks
  uint32_t dst, dst_it_stride;  // byte offset inside the operand image
  int iters;
  int kmajor;              // 1: validity along the chunk = K tail, per-iteration = row;  0: chunk = N tail, per-iteration = k row
  int row0, row_step, row_lim;  // row (m / n) or k index of iteration 0, its step, and its limit
  int col0, col_lim;            // first element index along the chunk direction and its limit
  int vec;
};

__global__ void __launch_bounds__(GT_THREADS) gemm_tc_kernel(GemmTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const GemmArgs& g = a.g;
  const int BN = a.bn;
  const uint32_t a_bytes = GT_BM * 128, b_bytes = (uint32_t)BN * 128;
  const uint32_t stage_bytes = a_bytes + b_bytes;          // one k-block of A and B (raw -> hi in place; or lo)
  uint8_t* raw_ring = smem;                                // GT_STAGES stages
  uint8_t* lo_ring = smem + GT_STAGES * stage_bytes;       // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(lo_ring + 2 * stage_bytes);
  uint64_t* full = bars;                 // [GT_STAGES] GT_PRODUCERS arrivals: hi and lo images of k-block kb are ready
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
  const int mn_atoms = BN / 32;

  if (tid == 0) {
    for (int s = 0; s < GT_STAGES; ++s) {
      mbar_init(&full[s], GT_PRODUCERS);
      mbar_init(&done[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  // two accumulators: [0,BN) hi*hi, [BN,2BN) the two cross terms (2^-11 smaller, so the tensor core's truncating fp32
  // accumulation costs 2^-11 less there and the main accumulator sees a third of the additions)
  if (warp == GT_PRODUCERS / 32) tmem_alloc(tmem_slot, (uint32_t)(2 * BN));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == GT_PRODUCERS / 32) {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(GT_BM, BN, a.b_kmajor ? 0 : 1);
      const uint32_t acc_main = tmem_base, acc_x = tmem_base + (uint32_t)BN;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % GT_STAGES;
        mbar_wait(&full[s], (kb / GT_STAGES) & 1);
        tc_fence_after();
        const uint32_t ah = smem_u32(raw_ring + s * stage_bytes), bh = ah + a_bytes;
        const uint32_t al = smem_u32(lo_ring + (kb & 1) * stage_bytes), bl = al + a_bytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // 4 x k8 per 128-byte k-block
          const uint64_t dah = make_sw128_desc(ah + k * 32), dal = make_sw128_desc(al + k * 32);
          uint64_t dbh, dbl;
          if (a.b_kmajor) {
            dbh = make_sw128_desc(bh + k * 32);
            dbl = make_sw128_desc(bl + k * 32);
          } else {
            const uint32_t koff = (uint32_t)(2 * k * mn_atoms) * 512, lbo = 512, sbo = (uint32_t)mn_atoms * 512;
            dbh = make_mn32_desc(bh + koff, lbo, sbo);
            dbl = make_mn32_desc(bl + koff, lbo, sbo);
          }
          const uint32_t first = (kb == 0 && k == 0) ? 0u : 1u;
          umma_tf32(acc_x, dal, dbh, idesc, first);
          umma_tf32(acc_x, dah, dbl, idesc, 1u);
          umma_tf32(acc_main, dah, dbh, idesc, first);
        }
        umma_commit(&done[s]);
      }
      umma_commit(acc_full);
    }
  } else {
    // ============================ producers (256 threads) ============================
    // chunk ownership is fixed per thread, so a thread only ever splits chunks it copied itself (cp.async.wait_group suffices)
    ChunkPlan pa, pb;
    {
      const int r0 = tid >> 3, c = tid & 7;
      pa.src = A + (long long)(m0 + r0) * g.lda + 4 * c;
      pa.it_stride = 32LL * g.lda; pa.kb_stride = GT_KB;
      pa.dst = sw128_chunk_off(r0, c); pa.dst_it_stride = 32 * 128;
      pa.iters = GT_BM / 32; pa.kmajor = 1;
      pa.row0 = m0 + r0; pa.row_step = 32; pa.row_lim = g.M; pa.col0 = 4 * c; pa.col_lim = g.K; pa.vec = a.a_vec;
      if (a.b_kmajor) {
        pb.src = B + (long long)(n0 + r0) * g.ldb + 4 * c;
        pb.it_stride = 32LL * g.ldb; pb.kb_stride = GT_KB;
        pb.dst = a_bytes + sw128_chunk_off(r0, c); pb.dst_it_stride = 32 * 128;
        pb.iters = BN / 32; pb.kmajor = 1;
        pb.row0 = n0 + r0; pb.row_step = 32; pb.row_lim = g.N; pb.col0 = 4 * c; pb.col_lim = g.K; pb.vec = a.b_vec;
      } else {
        const int cpr = BN / 4, kk0 = tid / cpr, cc = tid % cpr, kstep = GT_PRODUCERS / cpr;
        pb.src = B + (long long)kk0 * g.ldb + n0 + 4 * cc;
        pb.it_stride = (long long)kstep * g.ldb; pb.kb_stride = (long long)GT_KB * g.ldb;
        pb.dst = a_bytes + mn32_chunk_off(kk0, cc, mn_atoms); pb.dst_it_stride = (uint32_t)(kstep / 4) * mn_atoms * 512;
        pb.iters = 32 / kstep; pb.kmajor = 0;
        pb.row0 = kk0; pb.row_step = kstep; pb.row_lim = g.K; pb.col0 = n0 + 4 * cc; pb.col_lim = g.N; pb.vec = a.b_vec;
      }
    }
    auto issue_op = [&](const ChunkPlan& p, uint32_t stage_u32, int kb, const float* base) {
      const float* src = p.src + (long long)kb * p.kb_stride;
      const int cvalid = p.kmajor ? p.col_lim - (kb * GT_KB + p.col0) : p.col_lim - p.col0;
      const int roff = p.kmajor ? 0 : kb * GT_KB;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        if (it < p.iters) {
          const bool rv = (p.row0 + roff + it * p.row_step) < p.row_lim;
          copy_chunk(stage_u32 + p.dst + it * p.dst_it_stride, base, src + it * p.it_stride, rv ? cvalid : 0, p.vec);
        }
      }
    };
    auto split_op = [&](const ChunkPlan& p, uint8_t* hi, uint8_t* lo) {
#pragma unroll
      for (int it = 0; it < 4; ++it)
        if (it < p.iters) split_chunk(hi, lo, p.dst + it * p.dst_it_stride);
    };
    auto issue = [&](int kb) {
      const uint32_t st = smem_u32(raw_ring + (kb % GT_STAGES) * stage_bytes);
      issue_op(pa, st, kb, A);
      issue_op(pb, st, kb, B);
    };
    // prefetch distance GT_STAGES - 2: the split of k-block kb overlaps the MMAs of kb-1; a raw slot (and the lo slot of the same
    // parity) is recycled once the MMAs of k-block kb-2 are done
    for (int kb = 0; kb < GT_STAGES - 2; ++kb) {
      if (kb < nkb) issue(kb);
      cp_async_commit();
    }
    for (int kb = 0; kb < nkb; ++kb) {
      if (kb >= 2) mbar_wait(&done[(kb - 2) % GT_STAGES], ((kb - 2) / GT_STAGES) & 1);
      const int kn = kb + GT_STAGES - 2;
      if (kn < nkb) issue(kn);
      cp_async_commit();
      cp_async_wait<GT_STAGES - 2>();
      uint8_t* hi = raw_ring + (kb % GT_STAGES) * stage_bytes;
      uint8_t* lo = lo_ring + (kb & 1) * stage_bytes;
      split_op(pa, hi, lo);
      split_op(pb, hi, lo);
      fence_proxy_async();
      mbar_arrive(&full[kb % GT_STAGES]);
    }
    // ============================ epilogue (warps w and w+4 share TMEM lanes, each takes half of the columns) ============
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int r = (warp & 3) * 32 + lane, m = m0 + r;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const float rm = (g.rowmask && m < g.M) ? g.rowmask[m] : 1.f;
    const int c_begin = (warp >> 2) * (BN / 2), c_end = c_begin + BN / 2;
    for (int cb = c_begin; cb < c_end; cb += 32) {
      float v[32], x2[32];
      tmem_ld32(tmem_base + lane_base + cb, v);
      tmem_ld32(tmem_base + lane_base + BN + cb, x2);
      tmem_ld_wait();
      if (m >= g.M || n0 + cb >= g.N) continue;
      const int nv = min(32, g.N - (n0 + cb));
      float* crow = C + (long long)m * g.ldc + n0 + cb;
      const float* rrow = g.residual ? g.residual + (long long)m * g.ldr + n0 + cb : nullptr;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = g.alpha * (v[j] + x2[j]);
        if (g.bias && j < nv) x += __ldg(g.bias + n0 + cb + j);
        if (g.relu) x = fmaxf(x, 0.f);
        if (g.rowmask) x *= rm;
        v[j] = x;
      }
      if (nv == 32 && a.c_vec) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          if (rrow) {
            const float4 q = *reinterpret_cast<const float4*>(rrow + j);
            o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
          }
          if (g.accumulate) {
            const float4 q = *reinterpret_cast<const float4*>(crow + j);
            o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
          }
          *reinterpret_cast<float4*>(crow + j) = o;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (j < nv) {
            float x = v[j];
            if (rrow) x += rrow[j];
            if (g.accumulate) x += crow[j];
            crow[j] = x;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == GT_PRODUCERS / 32) tmem_dealloc(tmem_base, (uint32_t)(2 * BN));
}

inline size_t gemm_tc_smem_bytes(int bn) { return 1024 + (size_t)(GT_STAGES + 2) * (GT_BM * 128 + (size_t)bn * 128) + 128; }

inline bool aligned16(const void* p, long long ld, long long s1, long long s2) {
  return (reinterpret_cast<uintptr_t>(p) % 16 == 0) && (ld % 4 == 0) && (s1 % 4 == 0) && (s2 % 4 == 0);
}

// Same contract as launch_gemm (gemm_simt.cuh). Problems the tensor-core tile shape cannot serve well (N < 16, or the
// pair-broadcast epilogue) stay on the SIMT kernel.
inline cudaError_t launch_gemm_tc(const GemmArgs& g, bool b_kmajor, int batch, cudaStream_t st, int num_sms, int mn_swap = 0) {
  if (g.M <= 0 || g.N <= 0) return cudaSuccess;
  if (g.N < 16 || g.U != nullptr || g.K <= 0) return launch_gemm(g, b_kmajor, batch, st);
  GemmTcArgs a;
  a.g = g;
  a.b_kmajor = b_kmajor ? 1 : 0;
  const long long tiles128 = (long long)((g.M + GT_BM - 1) / GT_BM) * ((g.N + 127) / 128) * batch;
  a.bn = (g.N <= 64 || tiles128 < num_sms) ? 64 : 128;
  a.a_vec = aligned16(g.A, g.lda, g.sA1, g.sA2);
  a.b_vec = aligned16(g.B, g.ldb, g.sB1, g.sB2);
  a.c_vec = aligned16(g.C, g.ldc, g.sC1, g.sC2) && (!g.residual || aligned16(g.residual, g.ldr, 0, 0));
  a.mn_swap = mn_swap;
  dim3 grid((g.M + GT_BM - 1) / GT_BM, (g.N + a.bn - 1) / a.bn, batch);
  gemm_tc_kernel<<<grid, GT_THREADS, gemm_tc_smem_bytes(a.bn), st>>>(a);
  return cudaGetLastError();
}

}  // namespace tc
}  // namespace fdpt
