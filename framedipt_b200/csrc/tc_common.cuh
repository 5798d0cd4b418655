// tc_common.cuh — sm_100a tensor-core plumbing: mbarrier, bulk async copy (TMA, UBLKCP), TMEM alloc,
// tcgen05.mma / commit / ld, UMMA shared-memory + instruction descriptors, 128B-swizzled operand images.
//
// Operand convention used by every tcgen05 kernel in this library (kind::f16, fp16 inputs, fp32 accumulation):
//   * A [128 rows x K] and B [N rows x K] are K-major, stored in shared memory as K/64 "k-blocks"; one k-block is
//     rows x 64 halfs = rows x 128 B with the canonical SWIZZLE_128B pattern: 16-byte chunk c of row r sits at
//     chunk position (c ^ (r & 7)); 8 rows form one 1024-B swizzle atom, atoms are stacked along rows (SBO = 1024 B).
//   * k-block bases are 1024-B aligned; a k16 step inside a k-block advances the descriptor start address by 32 B.
//   * weights are pre-packed once into this exact image in global memory (pack_weight_image_kernel), so a plain 1-D
//     bulk copy (cp.async.bulk, no tensor map needed) lands them in shared memory ready for the MMA.
//   * D [128 x N] fp32 lives in TMEM: row r <-> lane r, column n <-> column base + n.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace fdpt {
namespace tc {

FDPT_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// bytes to skip so that the dynamic shared-memory window starts 1024-byte aligned (swizzle atoms).  Rounding the POINTER up through
// uintptr_t loses the address space: every later access becomes a generic LD / ST (long-scoreboard latency) instead of LDS / STS.
FDPT_DEVINL uint32_t smem_align_pad(const void* smem_raw) { return (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u; }

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its predecessor in the stream is still
// running (as soon as every predecessor CTA has called launch_dependents or exited); it must call pdl_wait() before it touches
// anything the predecessor wrote.  Used by the node-side GEMM kernels to overlap their prologue (barrier init, TMEM allocation, weight
// prefetch) with the tail of the previous kernel.
FDPT_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
FDPT_DEVINL void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// one lane of a converged warp (elect.sync): unlike `lane == 0`, ptxas knows that exactly one thread is active in the guarded region,
// so the warp-level tcgen05 / bulk-copy instructions in it are emitted once instead of inside a per-active-thread ELECT loop
FDPT_DEVINL bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ---------------------------------------------------------------------------------------
FDPT_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
FDPT_DEVINL void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
FDPT_DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// all state spaces: orders this thread's generic-proxy stores to GLOBAL memory (an operand image it has just written) before later
// async-proxy reads of them (a bulk copy issued after a barrier by another thread of the CTA)
FDPT_DEVINL void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
FDPT_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
FDPT_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
FDPT_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
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
// Bounded wait: a protocol bug must trap (reported as a CUDA error) instead of hanging the GPU.  mbar_wait_diag additionally records
// {block, thread, barrier offset, parity} of the first 31 waits that time out in a host-mapped buffer (FDPT_OPT_ET_TIMELINE).
__device__ unsigned long long* g_mbar_fail_buf = nullptr;
FDPT_DEVINL void mbar_timeout(uint32_t bar_addr, uint32_t parity) {  // inlined: a call here costs the hot kernels a stack frame
  unsigned long long* fb = g_mbar_fail_buf;
  if (fb) {
    const unsigned long long slot = atomicAdd(fb, 1ull);
    if (slot < 31)
      fb[1 + slot] = ((unsigned long long)blockIdx.x << 48) | ((unsigned long long)threadIdx.x << 32) | ((unsigned long long)(bar_addr & 0xFFFFFF) << 1) | parity;
    __threadfence_system();
    const long long t1 = clock64();
    while (clock64() - t1 < 400000000LL) {
    }
  }
  __trap();
}
FDPT_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
// same, leaving a record behind before the trap (bring-up of new protocols; costs the caller a few registers)
FDPT_DEVINL void mbar_wait_diag(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) mbar_timeout(smem_u32(bar), parity);
  }
}

// ---- bulk async copy global -> shared (TMA unit, 1-D), completion on an mbarrier -------------------------
FDPT_DEVINL void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- TMEM ---------------------------------------------------------------------------------------------
FDPT_DEVINL void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
FDPT_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
FDPT_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
FDPT_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread t of the warp <-> lane base + t)
FDPT_DEVINL void tmem_ld32(uint32_t taddr, float v[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 16 consecutive fp32 columns
FDPT_DEVINL void tmem_ld16(uint32_t taddr, float v[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
FDPT_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors ----------------------------------------------------------------------------------------
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (unused for swizzled K-major: 1) | [32,46) SBO >> 4 (1024 B -> 64)
//   [46,48) version = 1 (Blackwell) | [61,64) layout type = 2 (SWIZZLE_128B)
FDPT_DEVINL uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// SWIZZLE_128B descriptor with explicit leading / stride byte offsets.  MN-major operands (16-bit types) use both:
//   LBO = byte stride between 128-byte atoms along M/N (64 halfs), SBO = byte stride between 8-row atoms along K
//   (verified on B200 with the tf32 MN-major operand of gemm_tc.cuh: LBO steps along M/N, SBO along K).
FDPT_DEVINL uint64_t make_sw128_desc_ls(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor, kind::f16: fp16 A/B (format 0), fp32 accumulate, both operands K-major, shape M x N.
//   [4,6) c_format = 1 (F32) | [7,10) a_format | [10,13) b_format | [15] a_major | [16] b_major | [17,23) N >> 3 | [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
FDPT_DEVINL void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when complete (implies fence::before_thread_sync)
FDPT_DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- swizzled operand image helpers --------------------------------------------------------------------
constexpr int KB = 64;                 // halfs per k-block row (128 B)
// byte offset of 16-byte chunk `c` (8 halfs, k = 8c..8c+7 within the k-block) of row r inside a k-block image
FDPT_DEVINL uint32_t sw128_chunk_off(int r, int c) { return (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4); }

FDPT_DEVINL uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// fp32 weight W[n, k] (row stride ldw, optional column offset applied by the caller) -> fp16 swizzled image
//   image layout: [K/64 k-blocks][N rows][128 B], N multiple of 8, K multiple of 64 (zero padded by caller's K range)
__global__ void pack_weight_image_kernel(const float* __restrict__ W, int ldw, int N, int K, int Kvalid, __half* __restrict__ img) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte chunk per thread
  const long long chunks = (long long)N * (K / 8);
  if (idx >= chunks) return;
  const int n = (int)(idx / (K / 8));
  const int kc = (int)(idx % (K / 8));
  const int kb = kc / 8, c = kc % 8;
  uint32_t v[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int k = kc * 8 + 2 * e;
    const float a = (k < Kvalid) ? W[(long long)n * ldw + k] : 0.f;
    const float b = (k + 1 < Kvalid) ? W[(long long)n * ldw + k + 1] : 0.f;
    v[e] = pack_half2(a, b);
  }
  char* dst = reinterpret_cast<char*>(img) + (long long)kb * N * 128 + sw128_chunk_off(n, c);
  *reinterpret_cast<uint4*>(dst) = make_uint4(v[0], v[1], v[2], v[3]);
}

}  // namespace tc
}  // namespace fdpt
