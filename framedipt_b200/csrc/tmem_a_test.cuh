// tmem_a_test.cuh — bring-up unit of the "A operand from tensor memory" form of tcgen05.mma (kind::f16):
//   D[128 x 128] (fp32, TMEM) = A[128 x 64] (fp16, TMEM: lane = row, two K elements per 32-bit column) . B[128 x 64]^T (fp16, shared memory)
// A is written with tcgen05.st by the thread that owns the row.  Used by tests to pin the operand layout the fused EdgeTransition
// kernel relies on when it hands r2 from one GEMM to the next without a shared-memory round trip.
#pragma once
#include "tc_common.cuh"

namespace fdpt {
namespace tc {

FDPT_DEVINL void tmem_st32(uint32_t taddr, const uint32_t r[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]),
        "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
FDPT_DEVINL void tmem_st16(uint32_t taddr, const uint32_t r[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               :
               : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                 "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
               : "memory");
}
FDPT_DEVINL void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] . B[smem desc]^T
FDPT_DEVINL void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// one CTA of 128 threads: A [128, 64] fp32, B [128, 64] fp32 (row-major) -> D [128, 128] fp32
__global__ void __launch_bounds__(128) tmem_a_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + smem_align_pad(smem_raw);  // offset arithmetic on the __shared__ symbol: accesses stay LDS / STS
  uint8_t* Bimg = smem;  // 128 rows x 128 B, swizzled
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  // B row tid -> swizzled fp16 image
  for (int c = 0; c < 8; ++c) {
    const float* p = B + tid * 64 + c * 8;
    const uint4 u = make_uint4(pack_half2(p[0], p[1]), pack_half2(p[2], p[3]), pack_half2(p[4], p[5]), pack_half2(p[6], p[7]));
    *reinterpret_cast<uint4*>(Bimg + sw128_chunk_off(tid, c)) = u;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  // A row tid -> 32 packed columns at TMEM columns [128, 160)
  uint32_t r[32];
  for (int c = 0; c < 32; ++c) r[c] = pack_half2(A[tid * 64 + 2 * c], A[tid * 64 + 2 * c + 1]);
  tmem_st32(tmem_base + lane_base + 128, r);
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_f16(128, 128);
    for (int k = 0; k < 4; ++k) umma_f16_ts(tmem_base, tmem_base + 128 + 8 * k, make_sw128_desc(smem_u32(Bimg) + k * 32), idesc, k ? 1u : 0u);
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int q = 0; q < 4; ++q) {
    float v[32];
    tmem_ld32(tmem_base + lane_base + 32 * q, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[tid * 128 + 32 * q + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

}  // namespace tc
}  // namespace fdpt
