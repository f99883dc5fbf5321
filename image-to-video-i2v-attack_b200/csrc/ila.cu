// K9: intermediate-level attack loss of ILAF (reference image_attacks.py:586-612).  Per hooked layer, with
// delta = f - f_ori (the step's feature displacement), n = ||delta||_2, n0 = ||delta_0||_2 and d0 = delta_0 / n0 (the
// displacement of the adversarial example being fine-tuned, fixed for the whole run):
//     loss = -(0.5 * n / n0 + <d0, delta / n>)                                   (605-610)
//     d loss / d f = cA * delta + cB * d0,   cA = -(0.5 / (n0 n) - <d0, delta> / n^3),   cB = -1 / n
// The two sums couple the whole tensor, so the layer is reduced like K6: one HBM pass of per-thread float32 partials
// (<= 64 elements) promoted to FP64, block-reduced and combined in a fixed order (12 B / element), a one-thread
// finalize, and one pass that writes the gradient (16 B / element).  Bit-reproducible for a given launch shape.
#include "common.cuh"

namespace i2v {

constexpr int kIlaThreads = 256;
constexpr int kIlaMaxBlocks = 1184;

__global__ void __launch_bounds__(kIlaThreads)
ila_partial_kernel(const float* __restrict__ f, const float* __restrict__ o, const float* __restrict__ d0, int64_t n,
                   double* __restrict__ partials, int vec) {
    const int64_t n4 = vec ? n / 4 : 0;
    const float4* f4 = reinterpret_cast<const float4*>(f);
    const float4* o4 = reinterpret_cast<const float4*>(o);
    const float4* d4 = reinterpret_cast<const float4*>(d0);
    double s1 = 0.0, s2 = 0.0;
    const int64_t stride = (int64_t)gridDim.x * kIlaThreads;
    int64_t i = (int64_t)blockIdx.x * kIlaThreads + threadIdx.x;
    while (i < n4) {
        float p1[4] = {0.f, 0.f, 0.f, 0.f}, p2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (i < n4) {
                    const float4 a = ld_stream(f4 + i), b = ld_stream(o4 + i), c = ld_stream(d4 + i);
                    const float x = a.x - b.x, y = a.y - b.y, z = a.z - b.z, w = a.w - b.w;
                    p1[u] = fmaf(x, x, fmaf(y, y, fmaf(z, z, fmaf(w, w, p1[u]))));
                    p2[u] = fmaf(c.x, x, fmaf(c.y, y, fmaf(c.z, z, fmaf(c.w, w, p2[u]))));
                    i += stride;
                }
            }
        }
        s1 += ((double)p1[0] + (double)p1[1]) + ((double)p1[2] + (double)p1[3]);
        s2 += ((double)p2[0] + (double)p2[1]) + ((double)p2[2] + (double)p2[3]);
    }
    for (int64_t j = n4 * 4 + (int64_t)blockIdx.x * kIlaThreads + threadIdx.x; j < n; j += stride) {
        const double x = (double)(f[j] - o[j]);
        s1 += x * x;
        s2 += (double)d0[j] * x;
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    __shared__ double w1[kIlaThreads / 32], w2[kIlaThreads / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { w1[warp] = s1; w2[warp] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t1 = 0.0, t2 = 0.0;
        for (int w = 0; w < kIlaThreads / 32; ++w) { t1 += w1[w]; t2 += w2[w]; }
        partials[2 * blockIdx.x] = t1;
        partials[2 * blockIdx.x + 1] = t2;
    }
}

// one warp: partials in a fixed order -> (sum delta^2, sum d0*delta) -> loss and the two gradient coefficients
__global__ void ila_finalize_kernel(const double* __restrict__ partials, int blocks, float init_norm, float* __restrict__ stats,
                                    float* __restrict__ cost_log, const int* __restrict__ step_idx, int add_to_cost) {
    if (blockIdx.x != 0 || threadIdx.x >= 32) return;
    double t1 = 0.0, t2 = 0.0;                                   // lane l: partials l, l + 32, ... in order, then a butterfly
    for (int b = threadIdx.x; b < blocks; b += 32) { t1 += partials[2 * b]; t2 += partials[2 * b + 1]; }
    t1 = warp_sum(t1);
    t2 = warp_sum(t2);
    if (threadIdx.x != 0) return;
    const double n = sqrt(t1), n0 = (double)init_norm;
    const double loss = -(0.5 * n / n0 + t2 / n);
    stats[0] = (float)(-(0.5 / (n0 * n) - t2 / (n * n * n)));
    stats[1] = (float)(-1.0 / n);
    stats[2] = (float)loss;
    stats[3] = (float)n;
    if (cost_log) {
        const int s = step_idx ? *step_idx : 0;
        cost_log[s] = add_to_cost ? cost_log[s] + (float)loss : (float)loss;
    }
}

__global__ void __launch_bounds__(kIlaThreads)
ila_grad_kernel(const float* __restrict__ f, const float* __restrict__ o, const float* __restrict__ d0, float* __restrict__ grad,
                int64_t n, const float* __restrict__ stats, int vec) {
    const float cA = stats[0], cB = stats[1];
    const int64_t n4 = vec ? n / 4 : 0;
    const float4* f4 = reinterpret_cast<const float4*>(f);
    const float4* o4 = reinterpret_cast<const float4*>(o);
    const float4* d4 = reinterpret_cast<const float4*>(d0);
    float4* g4 = reinterpret_cast<float4*>(grad);
    const int64_t stride = (int64_t)gridDim.x * kIlaThreads;
    for (int64_t i = (int64_t)blockIdx.x * kIlaThreads + threadIdx.x; i < n4; i += stride) {
        const float4 a = ld_stream(f4 + i), b = ld_stream(o4 + i), c = ld_stream(d4 + i);
        float4 g;
        g.x = fmaf(cA, a.x - b.x, cB * c.x);
        g.y = fmaf(cA, a.y - b.y, cB * c.y);
        g.z = fmaf(cA, a.z - b.z, cB * c.z);
        g.w = fmaf(cA, a.w - b.w, cB * c.w);
        st_stream(g4 + i, g);
    }
    for (int64_t j = n4 * 4 + (int64_t)blockIdx.x * kIlaThreads + threadIdx.x; j < n; j += stride)
        grad[j] = fmaf(cA, f[j] - o[j], cB * d0[j]);
}

static int ila_blocks(int64_t n) {
    const int64_t want = ((n + 3) / 4 + kIlaThreads - 1) / kIlaThreads;
    const int64_t cap = (int64_t)sm_count() * 8 < kIlaMaxBlocks ? (int64_t)sm_count() * 8 : kIlaMaxBlocks;
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace i2v

using namespace i2v;

extern "C" int i2v_ila_workspace_doubles(void) { return 2 * kIlaMaxBlocks; }

extern "C" int i2v_ila_loss_f32(const float* f, const float* f_ori, const float* d0, int64_t n, float init_norm,
                                double* workspace, float* stats, float* cost_log, const int* step_idx, int add_to_cost,
                                i2v_stream_t stream) {
    I2V_REQUIRE(n >= 1, "empty feature map");
    I2V_REQUIRE(f && f_ori && d0 && workspace && stats, "null pointer");
    I2V_REQUIRE(init_norm > 0.f, "the initial displacement has zero norm (the example to fine-tune equals the original)");
    const int vec = ((reinterpret_cast<uintptr_t>(f) | reinterpret_cast<uintptr_t>(f_ori) | reinterpret_cast<uintptr_t>(d0)) & 15) == 0;
    const int blocks = ila_blocks(n);
    ila_partial_kernel<<<blocks, kIlaThreads, 0, as_stream(stream)>>>(f, f_ori, d0, n, workspace, vec);
    I2V_LAUNCH_CHECK("i2v_ila_loss_f32 (partials)");
    ila_finalize_kernel<<<1, 32, 0, as_stream(stream)>>>(workspace, blocks, init_norm, stats, cost_log, step_idx, add_to_cost);
    I2V_LAUNCH_CHECK("i2v_ila_loss_f32 (finalize)");
    return I2V_OK;
}

extern "C" int i2v_ila_grad_f32(const float* f, const float* f_ori, const float* d0, float* grad, int64_t n, const float* stats,
                                i2v_stream_t stream) {
    I2V_REQUIRE(n >= 0, "negative size");
    if (n == 0) return I2V_OK;
    I2V_REQUIRE(f && f_ori && d0 && grad && stats, "null pointer");
    const int vec = ((reinterpret_cast<uintptr_t>(f) | reinterpret_cast<uintptr_t>(f_ori) | reinterpret_cast<uintptr_t>(d0) |
                      reinterpret_cast<uintptr_t>(grad)) & 15) == 0;
    ila_grad_kernel<<<ila_blocks(n), kIlaThreads, 0, as_stream(stream)>>>(f, f_ori, d0, grad, n, stats, vec);
    I2V_LAUNCH_CHECK("i2v_ila_grad_f32");
    return I2V_OK;
}
