// K7: depth-wise "same" stencil over the gradient of a video batch — the smoothing step of the translation-invariant
// attacks (reference base_attacks.py:438-449 `TIFGSM._conv2d_frame`: F.conv2d(grad[:,:,t], kernel[3,1,15,15], groups=3,
// padding=7) frame by frame; 636-648 `TIFGSM3D._conv3d_frame`: F.conv3d(grad, kernel[3,1,15,15,15], groups=3, padding=7)).
// Every channel uses the same kernel, so the tensor is a stack of B*C independent [T,H,W] volumes:
//     out[v,t,h,w] = sum_{a,b,c} k[a,b,c] * in[v, t+a-kt/2, h+b-kh/2, w+c-kw/2]      (zero outside; kt = 1 for the 2-D case)
// Memory-bound on paper (8 B / element) but the 15^2 .. 15^3 taps make it FMA-bound in practice; one output per thread,
// the kernel weights in shared memory, input reads served by L1/L2 (neighbouring threads share 14/15 of their window).
#include "common.cuh"

namespace i2v {

__global__ void __launch_bounds__(256)
depthwise_stencil_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t volumes, int T, int H, int W,
                         const float* __restrict__ k, int kt, int kh, int kw) {
    extern __shared__ float ks[];
    const int taps = kt * kh * kw;
    for (int i = threadIdx.x; i < taps; i += blockDim.x) ks[i] = k[i];
    __syncthreads();
    const int rt = kt / 2, rh = kh / 2, rw = kw / 2;
    const int64_t plane = (int64_t)H * W, vol = plane * T, total = volumes * vol;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int w = (int)(i % W);
        int64_t r = i / W;
        const int h = (int)(r % H); r /= H;
        const int t = (int)(r % T);
        const int64_t v = r / T;
        const float* base = src + v * vol;
        float acc = 0.f;
        for (int a = 0; a < kt; ++a) {
            const int tt = t + a - rt;
            if (tt < 0 || tt >= T) continue;
            for (int b = 0; b < kh; ++b) {
                const int hh = h + b - rh;
                if (hh < 0 || hh >= H) continue;
                const float* row = base + (int64_t)tt * plane + (int64_t)hh * W;
                const float* kr = ks + (a * kh + b) * kw;
                const int c_lo = rw - w > 0 ? rw - w : 0;
                const int c_hi = W - 1 - w + rw < kw - 1 ? W - 1 - w + rw : kw - 1;
                for (int c = c_lo; c <= c_hi; ++c) acc = fmaf(kr[c], __ldg(row + w + c - rw), acc);
            }
        }
        dst[i] = acc;
    }
}

}  // namespace i2v

using namespace i2v;

extern "C" int i2v_depthwise_stencil_f32(const float* src, float* dst, int64_t volumes, int T, int H, int W, const float* k,
                                         int kt, int kh, int kw, i2v_stream_t stream) {
    I2V_REQUIRE(volumes >= 0 && T >= 1 && H >= 1 && W >= 1, "bad sizes");
    I2V_REQUIRE(kt >= 1 && kh >= 1 && kw >= 1 && (kt & 1) && (kh & 1) && (kw & 1) && kt * kh * kw <= 8192,
                "kernel extents must be odd and hold at most 8192 taps");
    if (volumes == 0) return I2V_OK;
    I2V_REQUIRE(src && dst && k && src != dst, "null or aliased pointer");
    const int64_t total = volumes * T * H * W;
    const int64_t want = (total + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    const int blocks = (int)(want < cap ? want : cap);
    depthwise_stencil_kernel<<<blocks, 256, (size_t)kt * kh * kw * sizeof(float), as_stream(stream)>>>(src, dst, volumes, T, H, W, k,
                                                                                                kt, kh, kw);
    I2V_LAUNCH_CHECK("i2v_depthwise_stencil_f32");
    return I2V_OK;
}
