// K8: temporal-translation kernels (reference video_attacks.py, TemporalTranslation).
//
//   i2v_temporal_shift_stack_f32   the D cyclically shifted copies of the clip the attack feeds to the model each step
//                                  (video_attacks.py:93-105 `_cycle_move`, 192-200): out[d][b,c,(t + m_d) mod T] = adv[b,c,t]
//   i2v_temporal_combine_f32       the gradient augmentation (video_attacks.py:163-177 `_grad_augmentation` with
//                                  `_conv1d_frame` 80-91): with G_d the gradient of variant d,
//                                      s   = sum_d k_d * G_d[b,c,t]                     (same position, different frame)
//                                      dd  = sum_d k_d * G_d[b,c,(t + m_d) mod T]       (shifted back: same frame)
//                                      out = (1 - weight) * s + weight * dd
//
// Both are single-pass and HBM-bound: shift reads the clip once and writes D copies (4 + 4D bytes / element); combine
// reads every G_d element twice — the second read of a line is a frame or two away from the first and is served by L2 —
// and writes once (4D + 4 bytes / element of DRAM traffic).  128-bit accesses when the frame size allows.
#include "common.cuh"

namespace i2v {

constexpr int kMaxVariants = 32;

struct TemporalParams {
    int moves[kMaxVariants];
    float k[kMaxVariants];
};

__device__ __forceinline__ int wrap_frame(int t, int T) {
    t %= T;
    return t < 0 ? t + T : t;
}

// one thread per (b*c, t, vector) of the SOURCE clip; writes its D destinations
template <typename V>
__global__ void __launch_bounds__(256)
temporal_shift_stack_kernel(const V* __restrict__ adv, V* __restrict__ out, int64_t BC, int T, int64_t HWv, int D,
                            const TemporalParams prm) {
    const int64_t total = BC * T * HWv;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t hw = i % HWv;
        const int64_t r = i / HWv;
        const int t = (int)(r % T);
        const int64_t bc = r / T;
        const V v = adv[i];
        for (int d = 0; d < D; ++d) {
            const int tn = wrap_frame(t + prm.moves[d], T);
            out[(int64_t)d * total + (bc * T + tn) * HWv + hw] = v;
        }
    }
}

__device__ __forceinline__ float4 fma4(float k, float4 g, float4 acc) {
    return make_float4(__fmaf_rn(k, g.x, acc.x), __fmaf_rn(k, g.y, acc.y), __fmaf_rn(k, g.z, acc.z), __fmaf_rn(k, g.w, acc.w));
}
__device__ __forceinline__ float fma4(float k, float g, float acc) { return __fmaf_rn(k, g, acc); }
__device__ __forceinline__ float4 mix4(float w0, float4 s, float w1, float4 d) {
    return make_float4(__fadd_rn(__fmul_rn(w0, s.x), __fmul_rn(w1, d.x)), __fadd_rn(__fmul_rn(w0, s.y), __fmul_rn(w1, d.y)),
                       __fadd_rn(__fmul_rn(w0, s.z), __fmul_rn(w1, d.z)), __fadd_rn(__fmul_rn(w0, s.w), __fmul_rn(w1, d.w)));
}
__device__ __forceinline__ float mix4(float w0, float s, float w1, float d) {
    return __fadd_rn(__fmul_rn(w0, s), __fmul_rn(w1, d));
}
__device__ __forceinline__ float4 zero_of(float4) { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float zero_of(float) { return 0.f; }

template <typename V>
__global__ void __launch_bounds__(256)
temporal_combine_kernel(const V* __restrict__ grads, V* __restrict__ out, int64_t BC, int T, int64_t HWv, int D,
                        const TemporalParams prm, float w_same, float w_shift) {
    const int64_t total = BC * T * HWv;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t hw = i % HWv;
        const int64_t r = i / HWv;
        const int t = (int)(r % T);
        const int64_t bc = r / T;
        V s = zero_of(V()), dd = zero_of(V());
        for (int d = 0; d < D; ++d) {
            const V* g = grads + (int64_t)d * total;
            const int ts = wrap_frame(t + prm.moves[d], T);
            s = fma4(prm.k[d], g[i], s);
            dd = fma4(prm.k[d], g[(bc * T + ts) * HWv + hw], dd);
        }
        out[i] = mix4(w_same, s, w_shift, dd);
    }
}

static int temporal_grid(int64_t total) {
    const int64_t want = (total + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace i2v

using namespace i2v;

extern "C" int i2v_temporal_shift_stack_f32(const float* adv, float* out, int64_t BC, int T, int64_t HW, const int* moves,
                                            int D, i2v_stream_t stream) {
    I2V_REQUIRE(BC >= 0 && T >= 1 && HW >= 1 && D >= 1 && D <= kMaxVariants, "bad sizes (1 <= D <= 32)");
    if (BC == 0) return I2V_OK;
    I2V_REQUIRE(adv && out && moves && adv != out, "null or aliased pointer");
    TemporalParams prm{};
    for (int d = 0; d < D; ++d) prm.moves[d] = moves[d];
    const bool vec = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(adv) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (vec) {
        const int64_t total = BC * T * (HW / 4);
        temporal_shift_stack_kernel<float4><<<temporal_grid(total), 256, 0, as_stream(stream)>>>(
            reinterpret_cast<const float4*>(adv), reinterpret_cast<float4*>(out), BC, T, HW / 4, D, prm);
    } else {
        const int64_t total = BC * T * HW;
        temporal_shift_stack_kernel<float><<<temporal_grid(total), 256, 0, as_stream(stream)>>>(adv, out, BC, T, HW, D, prm);
    }
    I2V_LAUNCH_CHECK("i2v_temporal_shift_stack_f32");
    return I2V_OK;
}

extern "C" int i2v_temporal_combine_f32(const float* grads, const float* kernel, const int* moves, int D, double weight,
                                        float* out, int64_t BC, int T, int64_t HW, i2v_stream_t stream) {
    I2V_REQUIRE(BC >= 0 && T >= 1 && HW >= 1 && D >= 1 && D <= kMaxVariants, "bad sizes (1 <= D <= 32)");
    if (BC == 0) return I2V_OK;
    I2V_REQUIRE(grads && out && kernel && moves && grads != out, "null or aliased pointer");
    TemporalParams prm{};
    for (int d = 0; d < D; ++d) { prm.moves[d] = moves[d]; prm.k[d] = kernel[d]; }
    // (1 - weight) is formed in double like the Python expression `(1-self.weight)` and rounded once (video_attacks.py:176)
    const float w_same = (float)(1.0 - weight), w_shift = (float)weight;
    const bool vec = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (vec) {
        const int64_t total = BC * T * (HW / 4);
        temporal_combine_kernel<float4><<<temporal_grid(total), 256, 0, as_stream(stream)>>>(
            reinterpret_cast<const float4*>(grads), reinterpret_cast<float4*>(out), BC, T, HW / 4, D, prm, w_same, w_shift);
    } else {
        const int64_t total = BC * T * HW;
        temporal_combine_kernel<float><<<temporal_grid(total), 256, 0, as_stream(stream)>>>(grads, out, BC, T, HW, D, prm, w_same,
                                                                                           w_shift);
    }
    I2V_LAUNCH_CHECK("i2v_temporal_combine_f32");
    return I2V_OK;
}
