// Blackwell (sm_100a) primitives used by the tensor-core graph-linear kernel: mbarrier, TMA
// (cp.async.bulk.tensor), tcgen05.mma / commit / ld, TMEM allocation, UMMA descriptors.
// Descriptor bit layouts follow the PTX ISA "tcgen05 matrix descriptor" / "instruction
// descriptor" tables (same fields as cute::UMMA::SmemDescriptor / InstrDescriptor).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

namespace sd { namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    // suspend-time hint: the thread is parked by the hardware instead of re-issuing the test (waiting producer warps share
    // their schedulers with the epilogue warps; without the hint ~30 % of their instructions were this spin loop)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
    return ok;
}
// Bounded wait: a protocol bug traps (CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag = 0) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) {
            printf("skeldiff: mbarrier wait timed out (block %d thread %d bar %p parity %u tag %d state %016llx)\n", blockIdx.x, threadIdx.x,
                   (void*)bar, parity, tag, *reinterpret_cast<volatile unsigned long long*>(bar));
            __trap();
        }
    }
}

__device__ __forceinline__ void mbar_timeout(uint64_t* bar, uint32_t parity) {
    printf("skeldiff: mbarrier wait timed out (block %d thread %d bar %p parity %u state %016llx)\n", blockIdx.x, threadIdx.x,
           (void*)bar, parity, *reinterpret_cast<volatile unsigned long long*>(bar));
    __trap();
}
// The same bounded wait expanded AT THE CALL SITE, so that a profiler's source view attributes the waiting time to the line of
// the kernel that waits (every inlined mbar_wait shares the lines above).
#define MBAR_WAIT_AT(bar_, parity_) do { \
        uint64_t* _b = (bar_); const uint32_t _p = (parity_); uint32_t _ok, _spins = 0; \
        do { \
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}" \
                         : "=r"(_ok) : "r"(sd::tc::smem_u32(_b)), "r"(_p), "r"(0x989680u) : "memory"); \
            if (!_ok && ++_spins > (1u << 24)) sd::tc::mbar_timeout(_b, _p); \
        } while (!_ok); \
    } while (0)

// ------------------------------------------------------------------ TMA
__device__ __forceinline__ void tma_prefetch_desc(const void* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// ------------------------------------------------------------------ TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ------------------------------------------------------------------ UMMA
// K-major operand tile [rows][64 bf16] (128-byte rows), SWIZZLE_128B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);   // start address, bits [0,14)
    d |= (uint64_t)0 << 16;                          // leading byte offset: unused for swizzled K-major
    d |= (uint64_t)(1024u >> 4) << 32;               // stride byte offset (next 8-row group), bits [32,46)
    d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                          // layout type SWIZZLE_128B
    return d;
}
// kind::f16, A = B = bf16 (format 1), D = f32 (format 1), both operands K-major
__host__ __device__ constexpr uint32_t umma_idesc_bf16(uint32_t M, uint32_t N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// kind::f16, A = B = f16 (format 0), D = f32
__host__ __device__ constexpr uint32_t umma_idesc_f16(uint32_t M, uint32_t N) {
    return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives on `bar` when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread = TMEM lane)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// tanh to ~1e-6 absolute error in 5 instructions (MUFU.EX2 + MUFU.RCP): 1 - 2 / (exp(2x) + 1).
// Used by the fp32-grade tensor-core epilogues, where 32 inlined copies of libdevice tanhf (~25 instructions
// and a branch each) would blow the instruction cache of the warp-specialised kernels.
__device__ __forceinline__ float tanh_acc(float x) {
    return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f);
}

// tanh with ~1e-7 ABSOLUTE error in 7 instructions: (1 - t) / (1 + t), t = exp(-2|x|) (MUFU.EX2 + MUFU.RCP, no branch).
// Its consumers are linear layers, so the absolute error is what propagates; it equals the rounding of a value of magnitude 1.
__device__ __forceinline__ float tanh_ex2(float x) {
    const float t = exp2f(-2.8853900817779268f * fabsf(x));
    return copysignf(__fdividef(1.0f - t, 1.0f + t), x);
}

// logistic function through MUFU.EX2 + MUFU.RCP (the gates of the fused GRU step; same formula as gru_gates_kernel<FAST>)
__device__ __forceinline__ float sigmoid_ex2(float x) {
    return __fdividef(1.0f, 1.0f + exp2f(-1.4426950408889634f * x));
}

}}  // namespace sd::tc
