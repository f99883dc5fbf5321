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

// Debug: does a SWIZZLE_128B K-major A descriptor whose start address is shifted by an arbitrary number of 128-byte ROWS
// (not a multiple of the 8-row swizzle atom) read the rows it points at?  A = 256 rows x 32 tf32 laid out the way TMA lays
// a tile out (16-byte chunk c of row r at chunk c ^ (r & 7), absolute address bits), B = 64 rows likewise;
// D[128 x 64] = A[shift .. shift + 127] B^T with the descriptor's base-offset field = 0 (mode 0) or (start >> 7) & 7 (mode 1).
__global__ void __launch_bounds__(128, 1)
mma_shift_probe_kernel(int shift, int mode, const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;                    // 256 rows x 128 B
    uint8_t* sb = smem + 32768;            // 64 rows x 128 B
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) {
        const int r = i >> 5, k = i & 31;
        *reinterpret_cast<float*>(sa + r * 128 + (((k >> 2) ^ (r & 7)) << 4) + (k & 3) * 4) = a[i];
    }
    for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) {
        const int r = i >> 5, k = i & 31;
        *reinterpret_cast<float*>(sb + r * 128 + (((k >> 2) ^ (r & 7)) << 4) + (k & 3) * 4) = b[i];
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(p_smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(p_smem_u32(&slot)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t a0 = p_smem_u32(sa) + (uint32_t)shift * 128u;
        auto desc = [&](uint32_t addr, bool with_off) {
            uint64_t d = (uint64_t)((addr & 0x3FFFF) >> 4);
            d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61;
            if (with_off) d |= (uint64_t)((addr >> 7) & 7) << 49;
            return d;
        };
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int kk = 0; kk < 4; ++kk) {
            const uint64_t da = desc(a0 + 32u * kk, mode == 1), db = desc(p_smem_u32(sb) + 32u * kk, false);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(kk ? 1 : 0) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(p_smem_u32(&bar)) : "memory");
    }
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(p_smem_u32(&bar)), "r"(0) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = threadIdx.x;
    for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                     "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) out[row * 64 + c0 + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(64) : "memory");
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

// a = [256][32], b = [64][32], out = [128][64] device f32; see mma_shift_probe_kernel
extern "C" int i2v_mma_shift_probe(int shift_rows, int base_offset_mode, const float* a, const float* b, float* out, i2v_stream_t stream) {
    I2V_REQUIRE(shift_rows >= 0 && shift_rows <= 128 && a && b && out, "bad probe arguments");
    const size_t smem = 32768 + 8192 + 1024;
    cudaError_t e = cudaFuncSetAttribute(mma_shift_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "i2v_mma_shift_probe (shared memory)");
    mma_shift_probe_kernel<<<1, 128, smem, as_stream(stream)>>>(shift_rows, base_offset_mode, a, b, out);
    I2V_LAUNCH_CHECK("i2v_mma_shift_probe");
    return I2V_OK;
}
