// Shared helpers for the i2v_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/i2v_b200.h"
#include "../../include/i2v_b200_debug.h"

namespace i2v {

// ---- error plumbing (thread-local text behind i2v_last_error) ---------------------------------
void set_error(const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what);

#define I2V_REQUIRE(cond, ...)                    \
    do {                                          \
        if (!(cond)) {                            \
            ::i2v::set_error(__VA_ARGS__);        \
            return I2V_EINVAL;                    \
        }                                         \
    } while (0)

#define I2V_LAUNCH_CHECK(what)                                 \
    do {                                                       \
        cudaError_t e__ = cudaGetLastError();                  \
        if (e__ != cudaSuccess) return ::i2v::cuda_fail(e__, what); \
    } while (0)

inline cudaStream_t as_stream(i2v_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// im2col / col2im passes of the tensor-core first-layer kernels (conv_stem.cu), launched by conv_tc.cu
int stem_im2col_launch(const float* x, float* col, int n0, int n, int H, int W, int P, int Q, int R, int stride, int pad, int Kp,
                       cudaStream_t st);
int stem_col2im_launch(const float* zt, float* dx, int N, int H, int W, int P, int Q, int R, int stride, int pad, cudaStream_t st);
int stem_pack_nhwc4_launch(const float* x, float* xp, int N, int H, int W, int Hp, int Wp, int pad, cudaStream_t st);

// 148 SMs on B200; queried once (falls back to 148 if the query fails before a device exists).
int sm_count();

// ---- ImageNet normalisation constants, rounded to f32 exactly as torch.as_tensor(list, f32) ----
// image_attacks.py:33-34 / base_attacks.py:39-40.  Channel 3 (NHWC4 padding) maps to mean 0 / std 1
// so that padded lanes stay exactly 0 through every kernel.
__device__ __forceinline__ float chan_mean(int c) {
    return c == 0 ? 0.485f : (c == 1 ? 0.456f : (c == 2 ? 0.406f : 0.0f));
}
__device__ __forceinline__ float chan_std(int c) {
    return c == 0 ? 0.229f : (c == 1 ? 0.224f : (c == 2 ? 0.225f : 1.0f));
}

// ---- 128-bit streaming loads / stores -----------------------------------------------------------
// Read-once data: bypass L1 allocation.  Read-modify-write state uses plain ld/st.
__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ld_plain(const float4* p) { return *p; }
// L2 eviction-priority hints (createpolicy + .L2::cache_hint): evict_last keeps lines that WILL be re-read resident while
// evict_first streams pass through (K1: the un-stashed tail of a frame slice against everything else)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float4 ld_hint(const float4* p, uint64_t policy) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(policy));
    return r;
}
__device__ __forceinline__ void st_hint(float4* p, const float4& v, uint64_t policy) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(policy) : "memory");
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// torch.clamp(x, lo, hi) == min(max(x, lo), hi) with NaN propagation; inputs here are never NaN.
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace i2v
