// K4/K5, first layer: the image-side convolution of every attacked backbone has Cin = 3 and Cout = 64
// (ResNet 7x7/s2/p3, AlexNet 11x11/s4/p2, VGG-16 3x3/s1/p1, SqueezeNet-1.1 3x3/s2/p0).  Three input channels
// are hostile to both the tensor-core path (a 12-byte pixel cannot be a TMA row) and the generic gather-GEMM
// (its 64-wide output tile would carry 3 useful columns in the data gradient: measured, the stem's dgrad alone
// took ~0.3 s of a 0.4 s step).  Two dedicated CUDA-core kernels instead:
//
//   stem forward : reads the [N,3,H,W] image exactly as the update kernel writes it, stages the input patch and
//                  the whole filter bank in shared memory, 2 pixels x 16 channels per thread, fused bias + ReLU,
//                  writes NHWC.
//   stem dgrad   : gather form, one image pixel per thread; a warp owns 32 pixels of one stride-parity class, so
//                  the set of valid filter taps is warp-uniform and filter reads are shared-memory broadcasts;
//                  the dy patch is staged with a padded pixel pitch (68 floats) so 128-bit reads are conflict-free;
//                  writes dcost/dimage as [N,3,H,W] for the Adam kernel.
#include "common.cuh"

namespace i2v {

constexpr int STEM_CO = 64;
constexpr int STEM_TH = 8, STEM_TW = 16;          // forward: output tile 8 x 16 pixels per iteration
constexpr int STEM_DY_PITCH = STEM_CO + 4;        // dgrad: floats per staged dy pixel (bank-conflict padding)

struct StemArgs {
    const float* src; const float* w; const float* bias; float* dst;
    int N, H, W, P, Q, R, stride, pad, relu;
};

// ---------------------------------------------------------------------------------------------------------
// forward: grid (ceil(Q/16), N); each CTA walks down the image in 8-row tiles with the filters resident
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
stem_fwd_kernel(const StemArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int K = 3 * a.R * a.R;
    const int PH = (STEM_TH - 1) * a.stride + a.R;
    int PW = (STEM_TW - 1) * a.stride + a.R;
    PW |= 1;                                          // odd pitch: the two pixel rows of a warp hit different banks
    float* wts = smem;                                // [K][64]
    float* patch = smem + (size_t)K * STEM_CO;        // [3][PH][PW]
    const int tid = threadIdx.x;
    const int n = blockIdx.y;
    const int q0 = blockIdx.x * STEM_TW;

    for (int i = tid; i < K * STEM_CO / 4; i += 256)
        reinterpret_cast<float4*>(wts)[i] = __ldg(reinterpret_cast<const float4*>(a.w) + i);

    const int cg = tid >> 6;                          // 16-channel group
    const int pp = tid & 63;
    const int py = pp >> 3, px = (pp & 7) * 2;        // this thread: pixels (py, px) and (py, px + 1) of the tile
    float bias_r[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) bias_r[j] = a.bias ? __ldg(a.bias + cg * 16 + j) : 0.f;

    const int ix0 = q0 * a.stride - a.pad;
    for (int p0 = 0; p0 < a.P; p0 += STEM_TH) {
        __syncthreads();                              // previous tile's patch fully consumed (and filters loaded)
        const int iy0 = p0 * a.stride - a.pad;
        for (int i = tid; i < 3 * PH * PW; i += 256) {
            const int c = i / (PH * PW);
            const int rem = i - c * PH * PW;
            const int yy = rem / PW, xx = rem - yy * PW;
            const int iy = iy0 + yy, ix = ix0 + xx;
            float v = 0.f;
            if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) v = __ldg(a.src + (((int64_t)n * 3 + c) * a.H + iy) * a.W + ix);
            patch[i] = v;
        }
        __syncthreads();
        float acc0[16], acc1[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { acc0[j] = 0.f; acc1[j] = 0.f; }
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < a.R; ++r) {
                const float* prow = patch + (c * PH + py * a.stride + r) * PW + px * a.stride;
                const float* wrow = wts + (size_t)((c * a.R + r) * a.R) * STEM_CO + cg * 16;
                for (int s = 0; s < a.R; ++s) {
                    const float v0 = prow[s], v1 = prow[s + a.stride];
                    const float4* w4 = reinterpret_cast<const float4*>(wrow + (size_t)s * STEM_CO);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 wv = w4[j];
                        acc0[4 * j + 0] = fmaf(v0, wv.x, acc0[4 * j + 0]); acc1[4 * j + 0] = fmaf(v1, wv.x, acc1[4 * j + 0]);
                        acc0[4 * j + 1] = fmaf(v0, wv.y, acc0[4 * j + 1]); acc1[4 * j + 1] = fmaf(v1, wv.y, acc1[4 * j + 1]);
                        acc0[4 * j + 2] = fmaf(v0, wv.z, acc0[4 * j + 2]); acc1[4 * j + 2] = fmaf(v1, wv.z, acc1[4 * j + 2]);
                        acc0[4 * j + 3] = fmaf(v0, wv.w, acc0[4 * j + 3]); acc1[4 * j + 3] = fmaf(v1, wv.w, acc1[4 * j + 3]);
                    }
                }
            }
        const int p = p0 + py;
        if (p < a.P) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int q = q0 + px + half;
                if (q >= a.Q) continue;
                const float* acc = half ? acc1 : acc0;
                float4* out = reinterpret_cast<float4*>(a.dst + (((int64_t)n * a.P + p) * a.Q + q) * STEM_CO + cg * 16);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float4 v = make_float4(acc[4 * j] + bias_r[4 * j], acc[4 * j + 1] + bias_r[4 * j + 1],
                                           acc[4 * j + 2] + bias_r[4 * j + 2], acc[4 * j + 3] + bias_r[4 * j + 3]);
                    if (a.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                    out[j] = v;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// data gradient: grid (ceil(W/(32*stride)), ceil(H/(8/stride)), N); warp = (tile row, stride class)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
stem_dgrad_kernel(const StemArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int st = a.stride, R = a.R;
    const int RT = 8 / st;                            // image rows per tile
    const int TWc = 32 * st;                          // image columns per tile
    const int h0 = blockIdx.y * RT, w0 = blockIdx.x * TWc, n = blockIdx.z;
    // dy rows / cols that can contribute to this tile:  p = (h + pad - r) / st
    int p_lo = h0 + a.pad - (R - 1); p_lo = p_lo > 0 ? (p_lo + st - 1) / st : 0;
    int p_hi = (h0 + RT - 1 + a.pad) / st; if (p_hi > a.P - 1) p_hi = a.P - 1;
    int q_lo = w0 + a.pad - (R - 1); q_lo = q_lo > 0 ? (q_lo + st - 1) / st : 0;
    int q_hi = (w0 + TWc - 1 + a.pad) / st; if (q_hi > a.Q - 1) q_hi = a.Q - 1;
    const int PR = (RT - 1 + R - 1) / st + 1;         // allocated patch extent (upper bounds, uniform across CTAs)
    const int PC = (TWc - 1 + R - 1) / st + 1;
    float* wts = smem;                                // [(r,s)][c][64]
    float* patch = smem + (size_t)R * R * 3 * STEM_CO;   // [PR][PC][68]
    const int tid = threadIdx.x;

    for (int i = tid; i < R * R * 3 * STEM_CO / 4; i += 256)
        reinterpret_cast<float4*>(wts)[i] = __ldg(reinterpret_cast<const float4*>(a.w) + i);
    // stage the dy patch (zeros outside the tensor), 16 float4 per pixel
    for (int i = tid; i < PR * PC * (STEM_CO / 4); i += 256) {
        const int c4 = i & 15;
        const int pix = i >> 4;
        const int pr = pix / PC, pc = pix - pr * PC;
        const int p = p_lo + pr, q = q_lo + pc;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p <= p_hi && q <= q_hi) v = __ldg(reinterpret_cast<const float4*>(a.src + (((int64_t)n * a.P + p) * a.Q + q) * STEM_CO) + c4);
        *reinterpret_cast<float4*>(patch + (size_t)pix * STEM_DY_PITCH + c4 * 4) = v;
    }
    __syncthreads();

    const int warp = tid >> 5, lane = tid & 31;
    const int h = h0 + warp / st;
    const int w = w0 + st * lane + (warp % st);
    float acc[3] = {0.f, 0.f, 0.f};
    if (h < a.H) {
        // warp-uniform tap sets: r = (h + pad) mod st (+ k*st), s = (w + pad) mod st (+ k*st)
        for (int r = (h + a.pad) % st; r < R; r += st) {
            const int p = (h + a.pad - r) / st;
            if (h + a.pad - r < 0 || p > p_hi) continue;
            for (int s = (w0 + (warp % st) + a.pad) % st; s < R; s += st) {
                const int t = w + a.pad - s;
                const int q = t / st;                              // per lane; out-of-range lanes read staged zeros or skip
                if (t < 0 || q > q_hi || q < q_lo) continue;
                const float4* d4 = reinterpret_cast<const float4*>(patch + (size_t)((p - p_lo) * PC + (q - q_lo)) * STEM_DY_PITCH);
                const float4* w4 = reinterpret_cast<const float4*>(wts + (size_t)(r * R + s) * 3 * STEM_CO);
#pragma unroll 4
                for (int j = 0; j < STEM_CO / 4; ++j) {
                    const float4 d = d4[j];
                    const float4 x0 = w4[j], x1 = w4[16 + j], x2 = w4[32 + j];
                    acc[0] = fmaf(d.x, x0.x, fmaf(d.y, x0.y, fmaf(d.z, x0.z, fmaf(d.w, x0.w, acc[0]))));
                    acc[1] = fmaf(d.x, x1.x, fmaf(d.y, x1.y, fmaf(d.z, x1.z, fmaf(d.w, x1.w, acc[1]))));
                    acc[2] = fmaf(d.x, x2.x, fmaf(d.y, x2.y, fmaf(d.z, x2.z, fmaf(d.w, x2.w, acc[2]))));
                }
            }
        }
        if (w < a.W) {
#pragma unroll
            for (int c = 0; c < 3; ++c) a.dst[(((int64_t)n * 3 + c) * a.H + h) * a.W + w] = acc[c];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// col2im of the tensor-core first-layer data gradient (i2v_conv_stem_dgrad_tc_f32): zt = planes [(c,r,s)][m],
// m = (n,p,q).  One image pixel per thread; a warp owns 32 pixels of ONE stride class of one image row, so its
// tap set is warp-uniform and every plane read is 128 contiguous bytes.  Taps are summed in (r, s) order.
// grid (ceil(W / (32*stride*warps_per_class...)), H, N): block = 8 warps = 8/stride pixel groups x stride classes
// ---------------------------------------------------------------------------------------------------------
// generic tap loops (any R / stride): fallback for filters with more than 4 taps per axis
__global__ void __launch_bounds__(256)
stem_col2im_loop_kernel(const float* __restrict__ zt, float* __restrict__ dx, int N, int H, int W, int P, int Q, int R, int st,
                        int pad, int64_t M) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cls = warp % st, grp = warp / st;
    const int h = blockIdx.y, n = blockIdx.z;
    const int w = (blockIdx.x * (8 / st) + grp) * 32 * st + lane * st + cls;
    if (w >= W) return;
    const int RR = R * R;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
    for (int r = (h + pad) % st; r < R; r += st) {
        const int hp = h + pad - r;
        if (hp < 0) break;
        const int p = hp / st;
        if (p >= P) continue;
        for (int s = (cls + pad) % st; s < R; s += st) {
            const int wq = w + pad - s;
            if (wq < 0) break;
            const int q = wq / st;
            if (q >= Q) continue;
            const float* z = zt + (int64_t)(r * R + s) * M + ((int64_t)n * P + p) * Q + q;
            acc0 += __ldg(z);
            acc1 += __ldg(z + (int64_t)RR * M);
            acc2 += __ldg(z + (int64_t)2 * RR * M);
        }
    }
    const int64_t o = (((int64_t)n * 3) * H + h) * W + w;
    dx[o] = acc0;
    dx[o + (int64_t)H * W] = acc1;
    dx[o + (int64_t)2 * H * W] = acc2;
}

// TAPS = ceil(R / stride) taps per axis at most: all <= 3*TAPS^2 plane reads of a pixel are issued BEFORE the first add
// (the first version looped with data-dependent bounds: three loads in flight per thread, i.e. latency-bound at ~50 % of
// the HBM rate); invalid taps read nothing and add +0, which leaves the (r, s)-ordered sum bit-identical.
template <int TAPS>
__global__ void __launch_bounds__(256)
stem_col2im_kernel(const float* __restrict__ zt, float* __restrict__ dx, int N, int H, int W, int P, int Q, int R, int st,
                   int pad, int64_t M) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cls = warp % st;                                   // w mod stride of this warp's pixels
    const int grp = warp / st;                                   // 8/st groups of 32 pixels per block
    const int h = blockIdx.y, n = blockIdx.z;
    const int w = (blockIdx.x * (8 / st) + grp) * 32 * st + lane * st + cls;
    if (w >= W) return;
    const int RR = R * R;
    const int r0 = (h + pad) % st, s0 = (cls + pad) % st;       // (w + pad) % st == (cls + pad) % st
    // tap (a, b): r = r0 + a*st, s = s0 + b*st, p = p0 - a, q = q0 - b  =>  the plane address is AFFINE in (a, b):
    // z(a, b) = z00 + a*(st*R*M - Q) + b*(st*M - 1)  (two 64-bit adds per tap instead of a multiply chain)
    const int p0 = (h + pad - r0) / st, q0 = (w + pad - s0) / st;         // h + pad - r0 and w + pad - s0 are multiples of st
    const float* __restrict__ z00 = zt + (int64_t)n * P * Q + (int64_t)(r0 * R + s0) * M + (int64_t)p0 * Q + q0;
    const int64_t step_a = (int64_t)st * R * M - Q, step_b = (int64_t)st * M - 1;
    const int64_t plane1 = (int64_t)RR * M, plane2 = 2 * plane1;
    float v[3][TAPS * TAPS];
#pragma unroll
    for (int a = 0; a < TAPS; ++a) {
        const int r = r0 + a * st;
        const int p = p0 - a;
        const bool okr = r < R && p >= 0 && p < P;
        const float* za = z00 + a * step_a;
#pragma unroll
        for (int b = 0; b < TAPS; ++b) {
            const int s = s0 + b * st;
            const int q = q0 - b;
            const bool ok = okr && s < R && q >= 0 && q < Q;
            const float* z = za + b * step_b;
            v[0][a * TAPS + b] = ok ? __ldg(z) : 0.f;
            v[1][a * TAPS + b] = ok ? __ldg(z + plane1) : 0.f;
            v[2][a * TAPS + b] = ok ? __ldg(z + plane2) : 0.f;
        }
    }
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
#pragma unroll
    for (int i = 0; i < TAPS * TAPS; ++i) { acc0 += v[0][i]; acc1 += v[1][i]; acc2 += v[2][i]; }
    const int64_t o = (((int64_t)n * 3) * H + h) * W + w;
    dx[o] = acc0;
    dx[o + (int64_t)H * W] = acc1;
    dx[o + (int64_t)2 * H * W] = acc2;
}

// ---------------------------------------------------------------------------------------------------------
// im2col of the tensor-core first-layer forward (i2v_conv_stem_fwd_tc_f32): x [N,3,H,W] -> col [(n,p,q)][Kp],
// k = (c,r,s), zero beyond 3*R*R and outside the image.  One CTA per output row (n, p): the R input rows of the
// three channels are staged in shared memory (zero-padded), then every thread assembles float4s of the row-major
// patch matrix through a k -> staged-offset table; global writes are fully coalesced (Kp floats per pixel).
// ---------------------------------------------------------------------------------------------------------
// (ncu on the first version — one flat index per float4 with a division and a table lookup per element — showed it
// issue-bound: 84 % issue-active at 45 % of the DRAM rate.  Now a thread owns ONE group of four k's for the whole row, so
// its four staged offsets are loop invariants and the inner loop is 4 LDS + 1 STG.128; taps beyond 3*R*R point at a
// zero row, so there is no select either.)
__global__ void __launch_bounds__(320)
stem_im2col_kernel(const float* __restrict__ x, float* __restrict__ col, int n0, int H, int W, int P, int Q, int R, int st,
                   int pad, int Kp, int qlanes) {
    extern __shared__ __align__(16) float smem[];
    const int pitch = (Q - 1) * st + R;                         // staged columns: image columns -pad .. -pad+pitch-1
    float* rows = smem;                                          // [3*R + 1][pitch]; the last row is zeros (padding taps)
    const int p = blockIdx.x, n = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    const int K = 3 * R * R;
    const int iy0 = p * st - pad;
    for (int cr = warp; cr <= 3 * R; cr += nwarps) {             // one staged row per warp and turn
        const int c = cr / R, r = cr - c * R;
        const int iy = iy0 + r;
        const bool row_ok = cr < 3 * R && iy >= 0 && iy < H;
        const float* __restrict__ src = x + (((int64_t)(n0 + n) * 3 + (row_ok ? c : 0)) * H + (row_ok ? iy : 0)) * W;
        for (int xx = lane; xx < pitch; xx += 32) {
            const int ix = xx - pad;
            rows[cr * pitch + xx] = (row_ok && ix >= 0 && ix < W) ? __ldg(src + ix) : 0.f;
        }
    }
    const int k4n = Kp / 4;
    const int k4 = tid % k4n, ql = tid / k4n;                    // this thread's k group and first pixel
    int off[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int k = k4 * 4 + j;
        int o = 3 * R * pitch;                                   // zero row
        if (k < K) { const int c = k / (R * R), rs = k - c * R * R, r = rs / R, s = rs - r * R; o = (c * R + r) * pitch + s; }
        off[j] = o;
    }
    __syncthreads();
    if (ql >= qlanes) return;
    float4* out = reinterpret_cast<float4*>(col + ((int64_t)n * P + p) * Q * Kp) + k4;
#pragma unroll 2
    for (int q = ql; q < Q; q += qlanes) {
        const int base = q * st;
        out[(int64_t)q * k4n] = make_float4(rows[off[0] + base], rows[off[1] + base], rows[off[2] + base], rows[off[3] + base]);
    }
}

// frames [n0, n0 + n) of x -> col (rows of frame n0 first)
int stem_im2col_launch(const float* x, float* col, int n0, int n, int H, int W, int P, int Q, int R, int stride, int pad, int Kp,
                       cudaStream_t st) {
    const int pitch = (Q - 1) * stride + R;
    const size_t smem = (size_t)(3 * R + 1) * pitch * sizeof(float);
    I2V_REQUIRE(smem <= 200 * 1024, "stem im2col: input rows do not fit in shared memory");
    const int k4n = Kp / 4;
    I2V_REQUIRE(k4n >= 1 && k4n <= 320, "stem im2col: filter too large");
    const int qlanes = 320 / k4n;                                // pixels in flight per turn; threads = k4n * qlanes <= 320
    const int threads = (k4n * qlanes + 31) / 32 * 32;
    cudaError_t e = cudaFuncSetAttribute(stem_im2col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "i2v_conv_stem_fwd_tc_f32 (im2col shared memory)");
    dim3 grid((unsigned)P, (unsigned)n);
    stem_im2col_kernel<<<grid, threads, smem, st>>>(x, col, n0, H, W, P, Q, R, stride, pad, Kp, qlanes);
    I2V_LAUNCH_CHECK("i2v_conv_stem_fwd_tc_f32 (im2col)");
    return I2V_OK;
}

// x [N,3,H,W] -> xp [N,Hp,Wp,4]: zero border of `pad` pixels (more on the right if Wp asks for it), channel 3 = 0.
// One thread per padded pixel: three coalesced plane reads, one 16-byte store.
__global__ void __launch_bounds__(256)
stem_pack_nhwc4_kernel(const float* __restrict__ x, float4* __restrict__ xp, int H, int W, int Hp, int Wp, int pad) {
    const int row = blockIdx.x;                       // n * Hp + hp
    const int n = row / Hp, hp = row - n * Hp;
    const int h = hp - pad;
    const bool row_ok = h >= 0 && h < H;
    const float* __restrict__ src = x + ((int64_t)n * 3 * H + (row_ok ? h : 0)) * W;
    const int64_t plane = (int64_t)H * W;
    for (int wp = threadIdx.x; wp < Wp; wp += 256) {
        const int w = wp - pad;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row_ok && w >= 0 && w < W) { v.x = __ldg(src + w); v.y = __ldg(src + plane + w); v.z = __ldg(src + 2 * plane + w); }
        xp[(int64_t)row * Wp + wp] = v;
    }
}

int stem_pack_nhwc4_launch(const float* x, float* xp, int N, int H, int W, int Hp, int Wp, int pad, cudaStream_t st) {
    I2V_REQUIRE((reinterpret_cast<uintptr_t>(xp) & 15) == 0, "padded image scratch must be 16-byte aligned");
    stem_pack_nhwc4_kernel<<<(unsigned)(N * Hp), 256, 0, st>>>(x, reinterpret_cast<float4*>(xp), H, W, Hp, Wp, pad);
    I2V_LAUNCH_CHECK("i2v_conv_stem_fwd_direct_f32 (pack)");
    return I2V_OK;
}

int stem_col2im_launch(const float* zt, float* dx, int N, int H, int W, int P, int Q, int R, int stride, int pad, cudaStream_t st) {
    I2V_REQUIRE(stride == 1 || stride == 2 || stride == 4 || stride == 8, "col2im: stride must divide 8");
    const int per_block = (8 / stride) * 32 * stride;              // image columns per block
    dim3 grid((unsigned)((W + per_block - 1) / per_block), (unsigned)H, (unsigned)N);
    const int taps = (R + stride - 1) / stride;
    const int64_t M = (int64_t)N * P * Q;
    if (taps > 4) {
        stem_col2im_loop_kernel<<<grid, 256, 0, st>>>(zt, dx, N, H, W, P, Q, R, stride, pad, M);
        I2V_LAUNCH_CHECK("i2v_conv_stem_dgrad_tc_f32 (col2im)");
        return I2V_OK;
    }
    switch (taps) {
        case 1: stem_col2im_kernel<1><<<grid, 256, 0, st>>>(zt, dx, N, H, W, P, Q, R, stride, pad, M); break;
        case 2: stem_col2im_kernel<2><<<grid, 256, 0, st>>>(zt, dx, N, H, W, P, Q, R, stride, pad, M); break;
        case 3: stem_col2im_kernel<3><<<grid, 256, 0, st>>>(zt, dx, N, H, W, P, Q, R, stride, pad, M); break;
        default: stem_col2im_kernel<4><<<grid, 256, 0, st>>>(zt, dx, N, H, W, P, Q, R, stride, pad, M); break;
    }
    I2V_LAUNCH_CHECK("i2v_conv_stem_dgrad_tc_f32 (col2im)");
    return I2V_OK;
}

}  // namespace i2v

using namespace i2v;

static int stem_check(const i2v_conv_desc* d) {
    I2V_REQUIRE(d, "null descriptor");
    I2V_REQUIRE(d->Cin == 3 && d->Cout == STEM_CO && d->R == d->S, "stem kernels take Cin=3, Cout=64, square filters");
    I2V_REQUIRE(d->stride == 1 || d->stride == 2 || d->stride == 4, "stem stride must be 1, 2 or 4");
    I2V_REQUIRE(d->R >= 1 && d->R <= 11 && d->pad >= 0 && d->pad < d->R, "unsupported stem filter");
    return I2V_OK;
}

extern "C" int i2v_conv_stem_supported(const i2v_conv_desc* d) {
    return d && d->Cin == 3 && d->Cout == STEM_CO && d->R == d->S && (d->stride == 1 || d->stride == 2 || d->stride == 4) &&
           d->R <= 11 && d->pad < d->R;
}

// x [N,3,H,W] (NCHW) -> y [N,P,Q,64] (NHWC);  w = [(c,r,s), 64] = weight[co,c,r,s]*bn_scale[co]
extern "C" int i2v_conv_stem_fwd_f32(const i2v_conv_desc* d, const float* x, const float* w, const float* bias, float* y,
                                     int flags, i2v_stream_t stream) {
    if (int r = stem_check(d)) return r;
    I2V_REQUIRE(x && w && y, "null pointer");
    I2V_REQUIRE(((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(y)) & 15) == 0, "w / y must be 16-byte aligned");
    if (d->N == 0) return I2V_OK;
    StemArgs a{x, w, bias, y, d->N, d->H, d->W, d->P, d->Q, d->R, d->stride, d->pad, (flags & I2V_EPI_RELU) ? 1 : 0};
    const int K = 3 * d->R * d->R;
    const int PH = (STEM_TH - 1) * d->stride + d->R;
    const int PW = ((STEM_TW - 1) * d->stride + d->R) | 1;
    const size_t smem = ((size_t)K * STEM_CO + (size_t)3 * PH * PW) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(stem_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "i2v_conv_stem_fwd_f32 (shared memory)");
    dim3 grid((unsigned)((d->Q + STEM_TW - 1) / STEM_TW), (unsigned)d->N);
    stem_fwd_kernel<<<grid, 256, smem, as_stream(stream)>>>(a);
    I2V_LAUNCH_CHECK("i2v_conv_stem_fwd_f32");
    return I2V_OK;
}

// dy [N,P,Q,64] (NHWC) -> dx [N,3,H,W] (NCHW);  w = [(r,s), c, 64] = weight[co,c,r,s]*bn_scale[co]
extern "C" int i2v_conv_stem_dgrad_f32(const i2v_conv_desc* d, const float* dy, const float* w, float* dx, i2v_stream_t stream) {
    if (int r = stem_check(d)) return r;
    I2V_REQUIRE(dy && w && dx, "null pointer");
    I2V_REQUIRE(((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(dy)) & 15) == 0, "w / dy must be 16-byte aligned");
    if (d->N == 0) return I2V_OK;
    StemArgs a{dy, w, nullptr, dx, d->N, d->H, d->W, d->P, d->Q, d->R, d->stride, d->pad, 0};
    const int st = d->stride, RT = 8 / st, TWc = 32 * st;
    const int PR = (RT - 1 + d->R - 1) / st + 1, PC = (TWc - 1 + d->R - 1) / st + 1;
    const size_t smem = ((size_t)d->R * d->R * 3 * STEM_CO + (size_t)PR * PC * STEM_DY_PITCH) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(stem_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "i2v_conv_stem_dgrad_f32 (shared memory)");
    dim3 grid((unsigned)((d->W + TWc - 1) / TWc), (unsigned)((d->H + RT - 1) / RT), (unsigned)d->N);
    stem_dgrad_kernel<<<grid, 256, smem, as_stream(stream)>>>(a);
    I2V_LAUNCH_CHECK("i2v_conv_stem_dgrad_f32");
    return I2V_OK;
}
