// Debug / measurement only: issue-rate probe of tcgen05.mma.kind::tf32 (one CTA per SM, one issuing thread).
// For a given N, number of independent accumulators (round-robin) and A source (shared memory descriptor or tensor
// memory), issues `count` MMAs of M = 128, K = 8, commits, waits for completion and reports cycles per instruction.
// Answers: is a chain of MMAs into ONE accumulator paced by the N/2-cycle floor or by a fixed dependent-issue latency?
#include <cuda.h>
#include "common.cuh"

namespace i2v {

__device__ __forceinline__ uint32_t p_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1)
mma_probe_kernel(int N, int accs, int a_tmem, int count, int issuers, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(p_smem_u32(&bar)), "r"(issuers) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(p_smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    // `issuers` warps (0, 2, 3 — warp 1 owns the allocation) issue concurrently, each into its own accumulators
    const int my = warp == 0 ? 0 : warp - 1;
    if ((threadIdx.x & 31) == 0 && warp != 1 && my < issuers) {
        auto desc = [](uint32_t addr) {
            uint64_t d = (uint64_t)((addr & 0x3FFFF) >> 4);
            d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61;
            return d;
        };
        const uint64_t da = desc(p_smem_u32(smem)), db = desc(p_smem_u32(smem + 16384));
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const long long t0 = clock64();
        for (int i = 0; i < count; ++i) {
            const uint32_t d = tmem + (uint32_t)(((my * accs + i % accs) * N) % 448);   // issuers * accs * N <= 448 columns
            const uint64_t a = da + 2 * (i & 3), b = db + 2 * (i & 3);
            if (a_tmem)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                             :: "r"(d), "r"(tmem + 448u + 8u * (i & 3)), "l"(b), "r"(idesc), "r"(1) : "memory");
            else
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                             :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(1) : "memory");
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(p_smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(p_smem_u32(&bar)), "r"(0) : "memory");
        const long long t2 = clock64();
        if (blockIdx.x == 0 && my == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
    }
}

}  // namespace i2v

using namespace i2v;

// out[0] = cycles spent issuing `count` MMAs, out[1] = cycles until all of them completed (device int64[2])
extern "C" int i2v_mma_probe(int N, int accs, int a_tmem, int count, int ctas, int issuers, long long* out, i2v_stream_t stream) {
    I2V_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0 && accs >= 1 && accs * N <= 384 && count >= 1 && out, "bad probe arguments");
    I2V_REQUIRE(issuers >= 1 && issuers <= 3 && issuers * accs * N <= 448, "1..3 issuing warps, issuers * accs * N <= 448 columns");
    const size_t smem = 16384 + 32768 + 1024;
    cudaError_t e = cudaFuncSetAttribute(mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "i2v_mma_probe (shared memory)");
    mma_probe_kernel<<<ctas > 0 ? ctas : 1, 128, smem, as_stream(stream)>>>(N, accs, a_tmem, count, issuers, out);
    I2V_LAUNCH_CHECK("i2v_mma_probe");
    return I2V_OK;
}
