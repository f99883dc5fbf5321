// DenseNet pieces of the truncated backbone (K4/K5), NHWC f32 — the two things a DenseNet has that the other families
// do not (reference image_attacks.py:95-98 only constructs the net; SURVEY.md D3 defines the `denseblock{d}` hook):
//
//   pre-activation   : every dense layer and every transition starts with BatchNorm -> ReLU applied to the CONCATENATED
//                      features, i.e. a different per-channel affine per consumer of the same buffer, with the ReLU between
//                      it and the convolution — it cannot be folded into the convolution weights.  `i2v_bn_relu_f32` reads
//                      the first C channels of the concat buffer (row pitch src_ld) and writes relu(scale*x + shift) densely
//                      ([M, Cp], channels C..Cp-1 zero: the tensor-core kernels want multiples of 64).  Its backward is not
//                      a kernel: the scale is folded into the consumer's data-gradient weights and the 1[t > 0] mask is the
//                      convolution epilogue's.
//   transition       : 2x2 / stride-2 average pooling (forward writes straight into the next block's concat buffer, backward
//                      reads its slice of that block's gradient).
//
// Both are HBM-bound streaming passes: 128-bit accesses, one full wave of grid-stride CTAs.
#include "common.cuh"

namespace i2v {

static int dense_grid(int64_t total, int threads = 256) {
    int64_t want = (total + threads - 1) / threads;
    int64_t cap = (int64_t)sm_count() * 8;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

// dst[m, c] = c < C ? max(0, fma(scale[c], src[m*src_ld + c], shift[c])) : 0      (c < Cp; all counts multiples of 4)
__global__ void __launch_bounds__(256)
bn_relu_kernel(const float* __restrict__ src, int64_t M, int C, int Cp, int src_ld, const float* __restrict__ scale,
               const float* __restrict__ shift, float* __restrict__ dst) {
    const int C4 = Cp >> 2;
    const int64_t total = M * C4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        const int64_t m = i / C4;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c4 * 4 < C) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(src + m * src_ld) + c4);
            const float4 a = __ldg(reinterpret_cast<const float4*>(scale) + c4);
            const float4 b = __ldg(reinterpret_cast<const float4*>(shift) + c4);
            o.x = fmaxf(fmaf(a.x, v.x, b.x), 0.f);
            o.y = fmaxf(fmaf(a.y, v.y, b.y), 0.f);
            o.z = fmaxf(fmaf(a.z, v.z, b.z), 0.f);
            o.w = fmaxf(fmaf(a.w, v.w, b.w), 0.f);
        }
        reinterpret_cast<float4*>(dst)[i] = o;
    }
}

// y[n, p, q, dst_off + c] = 0.25 * (x[n,2p,2q,c] + x[n,2p,2q+1,c] + x[n,2p+1,2q,c] + x[n,2p+1,2q+1,c]), summed in that order
__global__ void __launch_bounds__(256)
avgpool2_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int C, int P, int Q, int dst_ld,
                    int dst_off) {
    const int C4 = C >> 2;
    const int64_t total = (int64_t)N * P * Q * C4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        int64_t m = i / C4;
        const int q = (int)(m % Q); m /= Q;
        const int p = (int)(m % P);
        const int64_t n = m / P;
        const float4* r0 = reinterpret_cast<const float4*>(x + ((n * H + 2 * p) * W + 2 * q) * (int64_t)C) + c4;
        const float4* r1 = reinterpret_cast<const float4*>(x + ((n * H + 2 * p + 1) * W + 2 * q) * (int64_t)C) + c4;
        const float4 a = __ldg(r0), b = __ldg(r0 + C4), c = __ldg(r1), d = __ldg(r1 + C4);
        float4 o;
        o.x = 0.25f * (((a.x + b.x) + c.x) + d.x);
        o.y = 0.25f * (((a.y + b.y) + c.y) + d.y);
        o.z = 0.25f * (((a.z + b.z) + c.z) + d.z);
        o.w = 0.25f * (((a.w + b.w) + c.w) + d.w);
        *(reinterpret_cast<float4*>(y + ((n * P + p) * Q + q) * (int64_t)dst_ld + dst_off) + c4) = o;
    }
}

// dx[n, h, w, c] = 0.25 * dy[n, h/2, w/2, src_off + c] for h < 2P, w < 2Q, else 0 (odd trailing row / column: floor mode)
__global__ void __launch_bounds__(256)
avgpool2_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int N, int H, int W, int C, int P, int Q, int src_ld,
                    int src_off) {
    const int C4 = C >> 2;
    const int64_t total = (int64_t)N * H * W * C4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        int64_t m = i / C4;
        const int w = (int)(m % W); m /= W;
        const int h = (int)(m % H);
        const int64_t n = m / H;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        const int p = h >> 1, q = w >> 1;
        if (p < P && q < Q) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(dy + ((n * P + p) * Q + q) * (int64_t)src_ld + src_off) + c4);
            o = make_float4(0.25f * g.x, 0.25f * g.y, 0.25f * g.z, 0.25f * g.w);
        }
        reinterpret_cast<float4*>(dx)[i] = o;
    }
}

}  // namespace i2v

using namespace i2v;

extern "C" int i2v_bn_relu_f32(const float* src, int64_t M, int C, int Cp, int src_ld, const float* scale, const float* shift,
                               float* dst, i2v_stream_t stream) {
    I2V_REQUIRE(src && scale && shift && dst, "null pointer");
    I2V_REQUIRE(C > 0 && C % 4 == 0 && Cp % 4 == 0 && Cp >= C && src_ld % 4 == 0 && src_ld >= C, "channel counts must be multiples of 4, C <= Cp, C <= src_ld");
    I2V_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(scale) |
                  reinterpret_cast<uintptr_t>(shift)) & 15) == 0, "pointers must be 16-byte aligned");
    if (M == 0) return I2V_OK;
    bn_relu_kernel<<<dense_grid(M * (Cp / 4)), 256, 0, as_stream(stream)>>>(src, M, C, Cp, src_ld, scale, shift, dst);
    I2V_LAUNCH_CHECK("i2v_bn_relu_f32");
    return I2V_OK;
}

extern "C" int i2v_avgpool2_fwd_f32(const float* x, float* y, int N, int H, int W, int C, int dst_ld, int dst_off,
                                    i2v_stream_t stream) {
    I2V_REQUIRE(x && y, "null pointer");
    I2V_REQUIRE(C % 4 == 0 && dst_ld % 4 == 0 && dst_off % 4 == 0 && dst_off + C <= dst_ld, "channel counts / offsets must be multiples of 4");
    I2V_REQUIRE(H >= 2 && W >= 2, "2x2 average pooling needs at least 2x2 pixels");
    if (N == 0) return I2V_OK;
    const int P = H / 2, Q = W / 2;
    avgpool2_fwd_kernel<<<dense_grid((int64_t)N * P * Q * (C / 4)), 256, 0, as_stream(stream)>>>(x, y, N, H, W, C, P, Q, dst_ld, dst_off);
    I2V_LAUNCH_CHECK("i2v_avgpool2_fwd_f32");
    return I2V_OK;
}

extern "C" int i2v_avgpool2_bwd_f32(const float* dy, float* dx, int N, int H, int W, int C, int src_ld, int src_off,
                                    i2v_stream_t stream) {
    I2V_REQUIRE(dy && dx, "null pointer");
    I2V_REQUIRE(C % 4 == 0 && src_ld % 4 == 0 && src_off % 4 == 0 && src_off + C <= src_ld, "channel counts / offsets must be multiples of 4");
    I2V_REQUIRE(H >= 2 && W >= 2, "2x2 average pooling needs at least 2x2 pixels");
    if (N == 0) return I2V_OK;
    avgpool2_bwd_kernel<<<dense_grid((int64_t)N * H * W * (C / 4)), 256, 0, as_stream(stream)>>>(dy, dx, N, H, W, C, H / 2, W / 2, src_ld, src_off);
    I2V_LAUNCH_CHECK("i2v_avgpool2_bwd_f32");
    return I2V_OK;
}
