// K6: Dispersion-Reduction loss (reference image_attacks.py:129-234, `ImageGuidedStd_Adam`): the cost is the
// UNBIASED standard deviation of the whole hooked feature map [N,C,h,w] (216-220: `activations.std()`), so unlike
// the cosine loss it couples every frame of the call.  Three HBM-bound kernels:
//
//   std_partial   one pass over a slice of the feature map: per-thread float32 partial sums of x and x*x over
//                 <= 64 elements (4 x float4 x 4 accumulators), promoted to FP64, block-reduced in a fixed order,
//                 one (sum, sumsq) pair per block               4 B / element
//   std_combine   ONE thread adds the partials in block order to the running FP64 accumulator of the layer
//                 (fixed order => bit-reproducible for a given chunking)
//   std_finalize  mean = s1/n, var = (s2 - s1*s1/n)/(n-1), std = sqrt(var) in FP64; writes the float scalars the
//                 gradient kernel needs and adds the layer's std to the step's cost
//   std_grad      d std / d x_i = (x_i - mean) / ((n-1) std), times 1[x_i > 0] when the gradient is kept
//                 pre-activation (the native engine's convention)   8 B / element
//
// Post-ReLU features have mean ~ std, so the one-pass variance loses no more than a few of FP64's 53 bits.
#include "common.cuh"

namespace i2v {

constexpr int kStdThreads = 256;
constexpr int kStdMaxBlocks = 1184;     // 8 x 148: one full wave at 8 CTAs / SM

__global__ void __launch_bounds__(kStdThreads)
std_partial_kernel(const float* __restrict__ a, int64_t n, double* __restrict__ partials, int vec) {
    const int64_t n4 = vec ? n / 4 : 0;                      // unaligned slices (ragged frame chunks) take the scalar loop
    const float4* a4 = reinterpret_cast<const float4*>(a);
    double s1 = 0.0, s2 = 0.0;
    const int64_t stride = (int64_t)gridDim.x * kStdThreads;
    int64_t i = (int64_t)blockIdx.x * kStdThreads + threadIdx.x;
    while (i < n4) {
        // up to 16 float4 = 64 elements in float32, then promote
        float p1[4] = {0.f, 0.f, 0.f, 0.f}, p2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (i < n4) {
                    const float4 v = ld_stream(a4 + i);
                    p1[u] += (v.x + v.y) + (v.z + v.w);
                    p2[u] = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, p2[u]))));
                    i += stride;
                }
            }
        }
        s1 += ((double)p1[0] + (double)p1[1]) + ((double)p1[2] + (double)p1[3]);
        s2 += ((double)p2[0] + (double)p2[1]) + ((double)p2[2] + (double)p2[3]);
    }
    for (int64_t j = n4 * 4 + (int64_t)blockIdx.x * kStdThreads + threadIdx.x; j < n; j += stride) {   // tail / scalar path
        const double v = a[j];
        s1 += v;
        s2 += v * v;
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    __shared__ double w1[kStdThreads / 32], w2[kStdThreads / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { w1[warp] = s1; w2[warp] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t1 = 0.0, t2 = 0.0;
        for (int w = 0; w < kStdThreads / 32; ++w) { t1 += w1[w]; t2 += w2[w]; }
        partials[2 * blockIdx.x] = t1;
        partials[2 * blockIdx.x + 1] = t2;
    }
}

// One WARP: lane l adds partials l, l + 32, ... in order, then a fixed-shape butterfly — still one fixed summation order
// for a given block count.  (The first version walked all ~1200 partials with ONE thread: a chain of dependent-latency
// global loads that took ~80 us, five times the HBM pass it follows.)
__global__ void std_combine_kernel(const double* __restrict__ partials, int blocks, double* __restrict__ acc) {
    if (blockIdx.x != 0 || threadIdx.x >= 32) return;
    double t1 = 0.0, t2 = 0.0;
    for (int b = threadIdx.x; b < blocks; b += 32) { t1 += partials[2 * b]; t2 += partials[2 * b + 1]; }
    t1 = warp_sum(t1);
    t2 = warp_sum(t2);
    if (threadIdx.x == 0) {
        acc[0] += t1;
        acc[1] += t2;
    }
}

__global__ void std_finalize_kernel(const double* __restrict__ acc, int64_t n, float* __restrict__ stats,
                                    float* __restrict__ cost_log, const int* __restrict__ step_idx, int add_to_cost) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double dn = (double)n;
    const double mean = acc[0] / dn;
    const double var = (acc[1] - acc[0] * acc[0] / dn) / (dn - 1.0);     // n == 1: 0/0 = NaN, like torch.std()
    const double sd = sqrt(var > 0.0 ? var : (var == var ? 0.0 : var));
    stats[0] = (float)mean;
    stats[1] = (float)sd;
    stats[2] = (float)(1.0 / ((dn - 1.0) * sd));
    if (cost_log) {
        const int s = step_idx ? *step_idx : 0;
        cost_log[s] = add_to_cost ? cost_log[s] + (float)sd : (float)sd;
    }
}

__global__ void __launch_bounds__(kStdThreads)
std_grad_kernel(const float* __restrict__ a, float* __restrict__ grad, int64_t n, const float* __restrict__ stats,
                int relu_mask, int vec) {
    const float mean = stats[0], inv = stats[2];
    const int64_t n4 = vec ? n / 4 : 0;
    const float4* a4 = reinterpret_cast<const float4*>(a);
    float4* g4 = reinterpret_cast<float4*>(grad);
    const int64_t stride = (int64_t)gridDim.x * kStdThreads;
    for (int64_t i = (int64_t)blockIdx.x * kStdThreads + threadIdx.x; i < n4; i += stride) {
        const float4 v = ld_stream(a4 + i);
        float4 g = make_float4((v.x - mean) * inv, (v.y - mean) * inv, (v.z - mean) * inv, (v.w - mean) * inv);
        if (relu_mask) {
            if (!(v.x > 0.f)) g.x = 0.f;
            if (!(v.y > 0.f)) g.y = 0.f;
            if (!(v.z > 0.f)) g.z = 0.f;
            if (!(v.w > 0.f)) g.w = 0.f;
        }
        st_stream(g4 + i, g);
    }
    for (int64_t j = n4 * 4 + (int64_t)blockIdx.x * kStdThreads + threadIdx.x; j < n; j += stride) {
        const float v = a[j];
        grad[j] = (relu_mask && !(v > 0.f)) ? 0.f : (v - mean) * inv;
    }
}

static int std_blocks(int64_t n) {
    const int64_t want = ((n + 3) / 4 + kStdThreads - 1) / kStdThreads;
    const int64_t cap = (int64_t)sm_count() * 8 < kStdMaxBlocks ? (int64_t)sm_count() * 8 : kStdMaxBlocks;
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace i2v

using namespace i2v;

extern "C" int i2v_std_workspace_doubles(void) { return 2 * kStdMaxBlocks; }

extern "C" int i2v_std_accumulate_f32(const float* a, int64_t n, double* workspace, double* acc, i2v_stream_t stream) {
    I2V_REQUIRE(n >= 0, "negative size");
    if (n == 0) return I2V_OK;
    I2V_REQUIRE(a && workspace && acc, "null pointer");
    const int vec = (reinterpret_cast<uintptr_t>(a) & 15) == 0;
    const int blocks = std_blocks(n);
    std_partial_kernel<<<blocks, kStdThreads, 0, as_stream(stream)>>>(a, n, workspace, vec);
    I2V_LAUNCH_CHECK("i2v_std_accumulate_f32 (partials)");
    std_combine_kernel<<<1, 32, 0, as_stream(stream)>>>(workspace, blocks, acc);
    I2V_LAUNCH_CHECK("i2v_std_accumulate_f32 (combine)");
    return I2V_OK;
}

extern "C" int i2v_std_finalize_f32(const double* acc, int64_t n_total, float* stats, float* cost_log, const int* step_idx,
                                    int add_to_cost, i2v_stream_t stream) {
    I2V_REQUIRE(acc && stats, "null pointer");
    I2V_REQUIRE(n_total >= 1, "std of an empty tensor");
    std_finalize_kernel<<<1, 32, 0, as_stream(stream)>>>(acc, n_total, stats, cost_log, step_idx, add_to_cost);
    I2V_LAUNCH_CHECK("i2v_std_finalize_f32");
    return I2V_OK;
}

extern "C" int i2v_std_grad_f32(const float* a, float* grad, int64_t n, const float* stats, int relu_mask,
                                i2v_stream_t stream) {
    I2V_REQUIRE(n >= 0, "negative size");
    if (n == 0) return I2V_OK;
    I2V_REQUIRE(a && grad && stats, "null pointer");
    const int vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(grad)) & 15) == 0;
    const int blocks = std_blocks(n);
    std_grad_kernel<<<blocks, kStdThreads, 0, as_stream(stream)>>>(a, grad, n, stats, relu_mask, vec);
    I2V_LAUNCH_CHECK("i2v_std_grad_f32");
    return I2V_OK;
}
