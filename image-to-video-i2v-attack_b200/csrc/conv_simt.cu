// K4/K5 (CUDA-core path): generic NHWC convolution forward and data-gradient as ONE gather-GEMM kernel,
// plus max-pooling forward/backward.
//
//   out[m, n] = epilogue( sum_k A(m, k) * B[k, n] )        m = (image, y, x) of the OUTPUT tensor
//                                                           k = (r, s, c)    filter tap and SOURCE channel
//   FPROP : source = x  [N,H,W,C],  A(m,k) = x[img, y*stride - pad + r, x*stride - pad + s, c]
//   DGRAD : source = dy [N,H,W,C],  A(m,k) = dy[img, (y + pad - r)/stride, (x + pad - s)/stride, c]
//                                            (only where both divisions are exact and in range)
//   B is prepared by the host: FPROP B[(r,s,ci), co] = w[co,ci,r,s] * bn_scale[co]
//                              DGRAD B[(r,s,co), ci] = w[co,ci,r,s] * bn_scale[co]
//
// This is the exact-FP32 (FFMA, fp32 accumulate) path: it covers EVERY shape the attacked backbones
// contain (7x7/s2 Cin=3 stem read straight from the [N,3,H,W] image, AlexNet 11x11/s4 and 5x5, SqueezeNet
// 16..64-channel squeezes, strided data-gradients) and is the reference the tcgen05 tensor-core kernels
// (conv_tc.cu) are validated against.  128x64x16 tiles, 256 threads, 8x4 register blocking, float4
// shared-memory reads, register-staged double buffering.
//
// Replaces, per conv layer, what the reference reaches through torchvision + cuDNN:
//   forward  `_ = self.model(true_image)`   image_attacks.py:334 (conv + eval-mode BN + ReLU [+ residual])
//   backward `cost.backward()`              image_attacks.py:352 (data gradient only: the reference also
//            computes weight gradients that nothing reads)
#include "common.cuh"

namespace i2v {

constexpr int BM = 128, BN = 64, BK = 16, CT = 256;
constexpr int APAD = 4;

struct ConvArgs {
    const float* src;        // gather source (x for fprop, dy for dgrad)
    const float* bmat;       // [R*S*C, ldb]
    const float* bias;       // [Kout] or null
    const float* residual;   // [M, Kout] (added before ReLU) or null
    const float* mask_src;   // [M, Kout] forward activation; output is zeroed where it is <= 0 (ReLU backward) or null
    float* dst;              // [M, Kout]
    int N, H, W, C;          // source dims
    int P, Q, Kout, ldb;     // output dims (rows = N*P*Q), real / padded output channels
    int R, S, stride, pad;
    int dgrad, relu, src_nchw, dst_nchw;
};

struct RowInfo { int img, y, x; bool ok; };

template <bool VEC_A>
__global__ void __launch_bounds__(CT)
conv_gather_gemm_kernel(const ConvArgs p) {
    __shared__ __align__(16) float As[2][BK][BM + APAD];
    __shared__ __align__(16) float Bs[2][BK][BN];

    const int tid = threadIdx.x;
    const int64_t M = (int64_t)p.N * p.P * p.Q;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int Kg = p.R * p.S * p.C;
    const int ktiles = (Kg + BK - 1) / BK;

    // --- A loader geometry: two (row, 4-wide k slot) pairs per thread
    const int a_kq = tid & 3;
    RowInfo rows[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int64_t m = m0 + (tid >> 2) + 64 * i;
        rows[i].ok = m < M;
        const int64_t mm = rows[i].ok ? m : 0;
        rows[i].img = (int)(mm / ((int64_t)p.P * p.Q));
        const int rem = (int)(mm - (int64_t)rows[i].img * p.P * p.Q);
        rows[i].y = rem / p.Q;
        rows[i].x = rem - rows[i].y * p.Q;
    }
    // --- B loader geometry
    const int b_k = tid >> 4, b_n = (tid & 15) * 4;

    auto src_coord = [&](const RowInfo& ri, int r, int s, int& iy, int& ix) -> bool {
        if (!p.dgrad) {
            iy = ri.y * p.stride - p.pad + r;
            ix = ri.x * p.stride - p.pad + s;
        } else {
            const int ty = ri.y + p.pad - r, tx = ri.x + p.pad - s;
            if (ty < 0 || tx < 0) return false;
            if (p.stride > 1) {
                if (ty % p.stride || tx % p.stride) return false;
                iy = ty / p.stride; ix = tx / p.stride;
            } else { iy = ty; ix = tx; }
        }
        return ri.ok && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
    };

    auto load_a = [&](int kt, float4 (&reg)[2]) {
        const int k0 = kt * BK + a_kq * 4;
        if (VEC_A) {
            // C % 4 == 0: the four k's share (r,s) and are contiguous channels of one NHWC pixel
            const int rs = k0 / p.C, c0 = k0 - rs * p.C;
            const int r = rs / p.S, s = rs - r * p.S;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                int iy, ix;
                reg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k0 < Kg && src_coord(rows[i], r, s, iy, ix))
                    reg[i] = __ldg(reinterpret_cast<const float4*>(p.src + (((int64_t)rows[i].img * p.H + iy) * p.W + ix) * p.C + c0));
            }
        } else {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int k = k0 + j;
                    v[j] = 0.f;
                    if (k < Kg) {
                        const int rs = k / p.C, c = k - rs * p.C;
                        const int r = rs / p.S, s = rs - r * p.S;
                        int iy, ix;
                        if (src_coord(rows[i], r, s, iy, ix)) {
                            const int64_t off = p.src_nchw ? (((int64_t)rows[i].img * p.C + c) * p.H + iy) * p.W + ix
                                                           : (((int64_t)rows[i].img * p.H + iy) * p.W + ix) * p.C + c;
                            v[j] = __ldg(p.src + off);
                        }
                    }
                }
                reg[i] = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
    };
    auto load_b = [&](int kt, float4& reg) {
        const int k = kt * BK + b_k, n = n0 + b_n;
        reg = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < Kg && n < p.ldb) reg = __ldg(reinterpret_cast<const float4*>(p.bmat + (int64_t)k * p.ldb + n));
    };
    auto store_tiles = [&](int buf, const float4 (&ra)[2], const float4& rb) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int row = (tid >> 2) + 64 * i;
            As[buf][a_kq * 4 + 0][row] = ra[i].x;
            As[buf][a_kq * 4 + 1][row] = ra[i].y;
            As[buf][a_kq * 4 + 2][row] = ra[i].z;
            As[buf][a_kq * 4 + 3][row] = ra[i].w;
        }
        *reinterpret_cast<float4*>(&Bs[buf][b_k][b_n]) = rb;
    };

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int ty = tid >> 4, tx = tid & 15;
    float4 ra[2], rb;
    load_a(0, ra);
    load_b(0, rb);
    store_tiles(0, ra, rb);
    __syncthreads();
    for (int kt = 0; kt < ktiles; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < ktiles) { load_a(kt + 1, ra); load_b(kt + 1, rb); }   // global loads in flight during the FMAs
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < ktiles) {
            store_tiles(buf ^ 1, ra, rb);
            __syncthreads();
        }
    }

    // --- epilogue: bias, residual, ReLU, ReLU-backward mask, store (NHWC float4 or scalar / NCHW)
    const int n = n0 + tx * 4;
    const bool vec_out = !p.dst_nchw && (p.Kout & 3) == 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t m = m0 + ty * 8 + i;
        if (m >= M || n >= p.Kout) continue;
        float v[4] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
        const int nv = (p.Kout - n) < 4 ? (p.Kout - n) : 4;
        if (vec_out) {
            const int64_t off = m * p.Kout + n;
            if (p.bias) { const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + n)); v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w; }
            if (p.residual) { const float4 t = __ldg(reinterpret_cast<const float4*>(p.residual + off)); v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w; }
            if (p.relu) { v[0] = fmaxf(v[0], 0.f); v[1] = fmaxf(v[1], 0.f); v[2] = fmaxf(v[2], 0.f); v[3] = fmaxf(v[3], 0.f); }
            if (p.mask_src) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(p.mask_src + off));
                if (!(t.x > 0.f)) v[0] = 0.f; if (!(t.y > 0.f)) v[1] = 0.f; if (!(t.z > 0.f)) v[2] = 0.f; if (!(t.w > 0.f)) v[3] = 0.f;
            }
            *reinterpret_cast<float4*>(p.dst + off) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
            const int img = (int)(m / ((int64_t)p.P * p.Q));
            const int rem = (int)(m - (int64_t)img * p.P * p.Q);
            for (int j = 0; j < nv; ++j) {
                const int64_t off = p.dst_nchw ? ((int64_t)img * p.Kout + n + j) * p.P * p.Q + rem : m * p.Kout + n + j;
                float t = v[j];
                if (p.bias) t += __ldg(p.bias + n + j);
                if (p.residual) t += __ldg(p.residual + off);
                if (p.relu) t = fmaxf(t, 0.f);
                if (p.mask_src && !(__ldg(p.mask_src + off) > 0.f)) t = 0.f;
                p.dst[off] = t;
            }
        }
    }
}

// --------------------------------------------------------------------------------------------------
// max pooling, NHWC, window k x k, stride s, padding pad (-inf), optional ceil_mode handled by the host
// through (P, Q).  argmax = r*k + s of the FIRST maximum in scan order (torch.nn.MaxPool2d semantics).
// One CTA per (image, output row): all index arithmetic is 32-bit and per-row (the first version spent its time in
// three 64-bit divisions per element: 962 / 334 us per 256 frames of ResNet's stem where HBM needs ~200 / 170 us).
// mark_dead: the pooled tensor is a ReLU output, so a window whose maximum is not > 0 can never pass a gradient on
// (1[x[argmax] > 0] = 1[y > 0]); it is marked argmax = 255 and the backward pass needs no ReLU mask at all.
// --------------------------------------------------------------------------------------------------
constexpr int kPoolThreads = 256;
constexpr unsigned char kPoolNoWinner = 255;

// idx / d for idx * d < 2^32 (one IMAD.HI): magic = ceil(2^32 / d); magic = 0 stands for d = 1
__device__ __forceinline__ int pool_div(int idx, unsigned magic) { return magic ? (int)__umulhi((unsigned)idx, magic) : idx; }

// KT, ST > 0: window size / stride known at compile time (3,2 and 2,2 cover every pooling of the supported backbones):
// the window loops unroll with predicates, so all K*K loads of an output are in flight at once and the index
// arithmetic is shifts.  KT = 0: run-time k / stride.  (ncu, per-row version with run-time loops: backward issue-bound at
// 85 % issue-active / 23 % DRAM, forward latency-bound at 53 % DRAM.)
template <int KT, int ST>
__global__ void __launch_bounds__(kPoolThreads)
maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, uint8_t* __restrict__ argmax, int H, int W,
                   int C4, unsigned c4_magic, int P, int Q, int k_rt, int stride_rt, int pad, int mark_dead) {
    const int k = KT > 0 ? KT : k_rt, stride = KT > 0 ? ST : stride_rt;
    const int row = blockIdx.x;                      // img * P + pp
    const int img = row / P, pp = row - img * P;
    const int iy0 = pp * stride - pad;
    const float4* __restrict__ x4 = reinterpret_cast<const float4*>(x) + (int64_t)img * H * W * C4;
    float4* __restrict__ y4 = reinterpret_cast<float4*>(y) + (int64_t)row * Q * C4;
    uchar4* __restrict__ a4 = reinterpret_cast<uchar4*>(argmax) + (int64_t)row * Q * C4;
    const int items = Q * C4;
    for (int idx = threadIdx.x; idx < items; idx += kPoolThreads) {
        const int q = pool_div(idx, c4_magic), c4 = idx - q * C4;
        const int ix0 = q * stride - pad;
        float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        int arg[4] = {0, 0, 0, 0};
        bool any = false;
        if (KT > 0) {
            float4 v[KT > 0 ? KT * KT : 1];
#pragma unroll
            for (int r = 0; r < KT; ++r)
#pragma unroll
                for (int s = 0; s < KT; ++s) {
                    const int iy = iy0 + r, ix = ix0 + s;
                    const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
                    v[r * KT + s] = ok ? __ldg(x4 + ((int64_t)iy * W + ix) * C4 + c4) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
                }
#pragma unroll
            for (int t = 0; t < KT * KT; ++t) {
                constexpr int KD = KT > 0 ? KT : 1;
                const int iy = iy0 + t / KD, ix = ix0 + t % KD;
                const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
                const float vv[4] = {v[t].x, v[t].y, v[t].z, v[t].w};
                if (ok) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (!any || vv[j] > best[j]) { best[j] = vv[j]; arg[j] = t; }
                    any = true;
                }
            }
        } else {
            for (int r = 0; r < k; ++r) {
                const int iy = iy0 + r;
                if (iy < 0 || iy >= H) continue;
                for (int s = 0; s < k; ++s) {
                    const int ix = ix0 + s;
                    if (ix < 0 || ix >= W) continue;
                    const float4 v = __ldg(x4 + ((int64_t)iy * W + ix) * C4 + c4);
                    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (!any || vv[j] > best[j]) { best[j] = vv[j]; arg[j] = r * k + s; }
                    any = true;
                }
            }
        }
        if (mark_dead) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (!(best[j] > 0.f)) arg[j] = kPoolNoWinner;
        }
        y4[idx] = make_float4(best[0], best[1], best[2], best[3]);
        a4[idx] = make_uchar4((unsigned char)arg[0], (unsigned char)arg[1], (unsigned char)arg[2], (unsigned char)arg[3]);
    }
}

// dx[img,h,w,c] = sum over the windows that contain (h,w) and whose argmax IS (h,w) of dy; optionally
// multiplied by the ReLU-backward mask of the pooled tensor's producer: mask_src = that forward activation x
// (mask_pooled = 0), or the POOLED output y (mask_pooled = 1) — the winner of a window IS y, so 1[x[argmax] > 0] =
// 1[y > 0], and y is k*k/stride^2 times smaller than x.  With argmax written by the forward pass's mark_dead mode no mask
// is needed (dead windows have no winner).  One CTA per (image, input row); the sum runs over windows in (p, q) order.
// MASK: 0 = none (the forward pass marked dead windows), 1 = f32 activation of the pooled tensor's producer, 2 = pooled
// output — a template parameter so that the hot (mask-free) instantiation carries none of the mask code.
template <int KT, int ST, int MASK>
__global__ void __launch_bounds__(kPoolThreads)
maxpool_bwd_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ argmax, const float* __restrict__ mask_src,
                   float* __restrict__ dx, int H, int W, int C4, unsigned c4_magic, int P, int Q, int k_rt, int stride_rt, int pad,
                   int accumulate) {
    const int k = KT > 0 ? KT : k_rt, stride = KT > 0 ? ST : stride_rt;
    constexpr int NW = KT > 0 ? (KT + ST - 1) / ST : 1;         // windows per axis that can contain a pixel
    const int row = blockIdx.x;                      // img * H + h
    const int img = row / H, h = row - img * H;
    // windows p with p*stride - pad <= h <= p*stride - pad + k - 1
    int p_lo = (h + pad - k + 1 + stride - 1) / stride; if (h + pad - k + 1 < 0) p_lo = 0;
    int p_hi = (h + pad) / stride; if (p_hi > P - 1) p_hi = P - 1;
    const float4* __restrict__ dy4 = reinterpret_cast<const float4*>(dy) + (int64_t)img * P * Q * C4;
    const uchar4* __restrict__ am4 = reinterpret_cast<const uchar4*>(argmax) + (int64_t)img * P * Q * C4;
    const float4* __restrict__ yk4 = MASK == 2 ? reinterpret_cast<const float4*>(mask_src) + (int64_t)img * P * Q * C4 : nullptr;
    const float4* __restrict__ mk4 = MASK == 1 ? reinterpret_cast<const float4*>(mask_src) + (int64_t)row * W * C4 : nullptr;
    float4* __restrict__ dx4 = reinterpret_cast<float4*>(dx) + (int64_t)row * W * C4;
    const int items = W * C4;
    for (int idx = threadIdx.x; idx < items; idx += kPoolThreads) {
        const int w = pool_div(idx, c4_magic), c4 = idx - w * C4;
        int q_lo = (w + pad - k + 1 + stride - 1) / stride; if (w + pad - k + 1 < 0) q_lo = 0;
        int q_hi = (w + pad) / stride; if (q_hi > Q - 1) q_hi = Q - 1;
        float g[4] = {0.f, 0.f, 0.f, 0.f};
        if (KT > 0) {
            uchar4 a[NW * NW];
            float4 d[NW * NW];
#pragma unroll
            for (int i = 0; i < NW; ++i)
#pragma unroll
                for (int j = 0; j < NW; ++j) {
                    const int pp = p_lo + i, q = q_lo + j;
                    const bool ok = pp <= p_hi && q <= q_hi;
                    const int o = (pp * Q + q) * C4 + c4;
                    a[i * NW + j] = ok ? __ldg(am4 + o) : make_uchar4(kPoolNoWinner, kPoolNoWinner, kPoolNoWinner, kPoolNoWinner);
                    d[i * NW + j] = ok ? __ldg(dy4 + o) : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (MASK == 2 && ok) {
                        const float4 yv = __ldg(yk4 + o);
                        float4& dd = d[i * NW + j];
                        if (!(yv.x > 0.f)) dd.x = 0.f; if (!(yv.y > 0.f)) dd.y = 0.f; if (!(yv.z > 0.f)) dd.z = 0.f; if (!(yv.w > 0.f)) dd.w = 0.f;
                    }
                }
#pragma unroll
            for (int i = 0; i < NW; ++i)
#pragma unroll
                for (int j = 0; j < NW; ++j) {
                    const int pp = p_lo + i, q = q_lo + j;
                    const int want = (h - (pp * ST - pad)) * KT + (w - (q * ST - pad));
                    const uchar4 aa = a[i * NW + j];
                    const float4 dd = d[i * NW + j];
                    if (aa.x == want) g[0] += dd.x;
                    if (aa.y == want) g[1] += dd.y;
                    if (aa.z == want) g[2] += dd.z;
                    if (aa.w == want) g[3] += dd.w;
                }
        } else {
            for (int pp = p_lo; pp <= p_hi; ++pp) {
                const int r = h - (pp * stride - pad);
                for (int q = q_lo; q <= q_hi; ++q) {
                    const int want = r * k + (w - (q * stride - pad));
                    const int o = (pp * Q + q) * C4 + c4;
                    const uchar4 a = __ldg(am4 + o);
                    float4 d = __ldg(dy4 + o);
                    if (MASK == 2) {
                        const float4 yv = __ldg(yk4 + o);
                        if (!(yv.x > 0.f)) d.x = 0.f; if (!(yv.y > 0.f)) d.y = 0.f; if (!(yv.z > 0.f)) d.z = 0.f; if (!(yv.w > 0.f)) d.w = 0.f;
                    }
                    if (a.x == want) g[0] += d.x;
                    if (a.y == want) g[1] += d.y;
                    if (a.z == want) g[2] += d.z;
                    if (a.w == want) g[3] += d.w;
                }
            }
        }
        if (MASK == 1) {
            const float4 mk = __ldg(mk4 + idx);
            if (!(mk.x > 0.f)) g[0] = 0.f; if (!(mk.y > 0.f)) g[1] = 0.f; if (!(mk.z > 0.f)) g[2] = 0.f; if (!(mk.w > 0.f)) g[3] = 0.f;
        }
        if (accumulate) {   // the pooled tensor's input has another consumer (a hooked layer: K1 wrote its gradient first)
            const float4 o = dx4[idx];
            g[0] += o.x; g[1] += o.y; g[2] += o.z; g[3] += o.w;
        }
        dx4[idx] = make_float4(g[0], g[1], g[2], g[3]);
    }
}

// channel concat / split helpers for Fire modules: dst[m, off : off+C] = src[m, :]  (and the reverse)
__global__ void __launch_bounds__(256)
copy_channels_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t M, int Csrc, int src_off, int Cdst,
                     int dst_off, int Ccopy, int accumulate) {
    const int C4 = Ccopy >> 2;
    const int64_t total = M * C4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        const int64_t m = i / C4;
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + m * Csrc + src_off) + c4);
        float4* d = reinterpret_cast<float4*>(dst + m * Cdst + dst_off) + c4;
        if (accumulate) { float4 o = *d; o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w; *d = o; }
        else *d = v;
    }
}

static int grid1d(int64_t total, int threads = 256) {
    int64_t want = (total + threads - 1) / threads;
    int64_t cap = (int64_t)sm_count() * 8;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

}  // namespace i2v

using namespace i2v;

static int conv_launch(const i2v_conv_desc* d, int dgrad, const float* src, const float* bmat, const float* bias,
                       const float* residual, const float* mask_src, float* dst, int flags, cudaStream_t st) {
    I2V_REQUIRE(d && src && bmat && dst, "null pointer");
    I2V_REQUIRE(d->N >= 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0 && d->R > 0 && d->S > 0 && d->stride > 0 &&
                d->pad >= 0 && d->P > 0 && d->Q > 0, "bad convolution descriptor");
    if (d->N == 0) return I2V_OK;
    ConvArgs a{};
    a.src = src; a.bmat = bmat; a.bias = bias; a.residual = residual; a.mask_src = mask_src; a.dst = dst;
    a.N = d->N; a.R = d->R; a.S = d->S; a.stride = d->stride; a.pad = d->pad;
    a.dgrad = dgrad;
    a.relu = (flags & I2V_EPI_RELU) ? 1 : 0;
    if (!dgrad) {   // source = x [N,H,W,Cin], output = y [N,P,Q,Cout]
        a.H = d->H; a.W = d->W; a.C = d->Cin; a.P = d->P; a.Q = d->Q; a.Kout = d->Cout;
        a.src_nchw = (flags & I2V_LAYOUT_X_NCHW) ? 1 : 0; a.dst_nchw = 0;
    } else {        // source = dy [N,P,Q,Cout], output = dx [N,H,W,Cin]
        a.H = d->P; a.W = d->Q; a.C = d->Cout; a.P = d->H; a.Q = d->W; a.Kout = d->Cin;
        a.src_nchw = 0; a.dst_nchw = (flags & I2V_LAYOUT_X_NCHW) ? 1 : 0;
    }
    a.ldb = (a.Kout + 3) & ~3;
    I2V_REQUIRE((reinterpret_cast<uintptr_t>(bmat) & 15) == 0, "weight matrix must be 16-byte aligned");
    const bool vec_a = (a.C % 4 == 0) && !a.src_nchw && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    const bool vec_out = !a.dst_nchw && (a.Kout % 4 == 0);
    if (vec_out) {
        I2V_REQUIRE((reinterpret_cast<uintptr_t>(dst) & 15) == 0, "output must be 16-byte aligned");
        I2V_REQUIRE(!bias || (reinterpret_cast<uintptr_t>(bias) & 15) == 0, "bias must be 16-byte aligned");
        I2V_REQUIRE(!residual || (reinterpret_cast<uintptr_t>(residual) & 15) == 0, "residual must be 16-byte aligned");
        I2V_REQUIRE(!mask_src || (reinterpret_cast<uintptr_t>(mask_src) & 15) == 0, "mask source must be 16-byte aligned");
    }
    const int64_t M = (int64_t)a.N * a.P * a.Q;
    dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((a.Kout + BN - 1) / BN));
    if (vec_a) conv_gather_gemm_kernel<true><<<grid, CT, 0, st>>>(a);
    else conv_gather_gemm_kernel<false><<<grid, CT, 0, st>>>(a);
    I2V_LAUNCH_CHECK(dgrad ? "i2v_conv_dgrad_simt_f32" : "i2v_conv_fwd_simt_f32");
    return I2V_OK;
}

extern "C" int i2v_conv_fwd_simt_f32(const i2v_conv_desc* d, const float* x, const float* bmat, const float* bias,
                                     const float* residual, float* y, int flags, i2v_stream_t stream) {
    return conv_launch(d, 0, x, bmat, bias, residual, nullptr, y, flags, as_stream(stream));
}

extern "C" int i2v_conv_dgrad_simt_f32(const i2v_conv_desc* d, const float* dy, const float* bmat, const float* addend,
                                       const float* mask_src, float* dx, int flags, i2v_stream_t stream) {
    return conv_launch(d, 1, dy, bmat, nullptr, addend, mask_src, dx, flags & ~I2V_EPI_RELU, as_stream(stream));
}

extern "C" int i2v_maxpool_fwd_flags_f32(const float* x, float* y, uint8_t* argmax, int N, int H, int W, int C, int P, int Q,
                                         int k, int stride, int pad, int flags, i2v_stream_t stream) {
    I2V_REQUIRE(x && y && argmax, "null pointer");
    I2V_REQUIRE(C % 4 == 0 && k >= 1 && k <= 15 && stride >= 1 && pad >= 0 && pad < k, "unsupported pooling shape");
    I2V_REQUIRE((int64_t)P * Q * (C / 4) < 0x7fffffff && (int64_t)N * P < 0x7fffffff, "pooled plane too large");
    if (N == 0 || P == 0 || Q == 0) return I2V_OK;
    I2V_REQUIRE((int64_t)(Q > W ? Q : W) * (C / 4) * (C / 4) < 0xffffffffLL, "row too long for the 32-bit index split");
    const int C4 = C / 4, md = (flags & 4) ? 1 : 0;
    const unsigned magic = C4 == 1 ? 0u : (unsigned)((0x100000000ULL + C4 - 1) / C4);
    const unsigned grid = (unsigned)(N * P);
    if (k == 3 && stride == 2)
        maxpool_fwd_kernel<3, 2><<<grid, kPoolThreads, 0, as_stream(stream)>>>(x, y, argmax, H, W, C4, magic, P, Q, k, stride, pad, md);
    else if (k == 2 && stride == 2)
        maxpool_fwd_kernel<2, 2><<<grid, kPoolThreads, 0, as_stream(stream)>>>(x, y, argmax, H, W, C4, magic, P, Q, k, stride, pad, md);
    else
        maxpool_fwd_kernel<0, 0><<<grid, kPoolThreads, 0, as_stream(stream)>>>(x, y, argmax, H, W, C4, magic, P, Q, k, stride, pad, md);
    I2V_LAUNCH_CHECK("i2v_maxpool_fwd_f32");
    return I2V_OK;
}

extern "C" int i2v_maxpool_fwd_f32(const float* x, float* y, uint8_t* argmax, int N, int H, int W, int C, int P, int Q,
                                   int k, int stride, int pad, i2v_stream_t stream) {
    return i2v_maxpool_fwd_flags_f32(x, y, argmax, N, H, W, C, P, Q, k, stride, pad, 0, stream);
}

extern "C" int i2v_maxpool_bwd_f32(const float* dy, const uint8_t* argmax, const float* mask_src, float* dx, int N, int H,
                                   int W, int C, int P, int Q, int k, int stride, int pad, int flags,
                                   i2v_stream_t stream) {
    I2V_REQUIRE(dy && argmax && dx, "null pointer");
    I2V_REQUIRE(C % 4 == 0 && k >= 1 && k <= 15 && stride >= 1 && pad >= 0 && pad < k, "unsupported pooling shape");
    I2V_REQUIRE(!(flags & 2) || mask_src, "I2V_POOL_MASK_POOLED needs mask_src = the pooled output");
    I2V_REQUIRE((int64_t)P * Q * (C / 4) < 0x7fffffff && (int64_t)N * H < 0x7fffffff, "pooled plane too large");
    if (N == 0 || H == 0 || W == 0) return I2V_OK;
    I2V_REQUIRE((int64_t)(Q > W ? Q : W) * (C / 4) * (C / 4) < 0xffffffffLL, "row too long for the 32-bit index split");
    const int C4 = C / 4, acc = flags & 1, mp = (flags >> 1) & 1;
    const unsigned magic = C4 == 1 ? 0u : (unsigned)((0x100000000ULL + C4 - 1) / C4);
    const unsigned grid = (unsigned)(N * H);
    const int mode = mask_src ? (mp ? 2 : 1) : 0;
#define I2V_POOL_BWD(KT_, ST_)                                                                                                  \
    do {                                                                                                                        \
        if (mode == 0) maxpool_bwd_kernel<KT_, ST_, 0><<<grid, kPoolThreads, 0, as_stream(stream)>>>(dy, argmax, mask_src, dx, H, W, C4, magic, P, Q, k, stride, pad, acc);       \
        else if (mode == 1) maxpool_bwd_kernel<KT_, ST_, 1><<<grid, kPoolThreads, 0, as_stream(stream)>>>(dy, argmax, mask_src, dx, H, W, C4, magic, P, Q, k, stride, pad, acc);  \
        else maxpool_bwd_kernel<KT_, ST_, 2><<<grid, kPoolThreads, 0, as_stream(stream)>>>(dy, argmax, mask_src, dx, H, W, C4, magic, P, Q, k, stride, pad, acc);                 \
    } while (0)
    if (k == 3 && stride == 2) I2V_POOL_BWD(3, 2);
    else if (k == 2 && stride == 2) I2V_POOL_BWD(2, 2);
    else I2V_POOL_BWD(0, 0);
#undef I2V_POOL_BWD
    I2V_LAUNCH_CHECK("i2v_maxpool_bwd_f32");
    return I2V_OK;
}

extern "C" int i2v_copy_channels_f32(const float* src, float* dst, int64_t M, int Csrc, int src_off, int Cdst, int dst_off,
                                     int Ccopy, int accumulate, i2v_stream_t stream) {
    I2V_REQUIRE(src && dst, "null pointer");
    I2V_REQUIRE(Ccopy % 4 == 0 && Csrc % 4 == 0 && Cdst % 4 == 0 && src_off % 4 == 0 && dst_off % 4 == 0, "channel counts / offsets must be multiples of 4");
    I2V_REQUIRE(src_off + Ccopy <= Csrc && dst_off + Ccopy <= Cdst, "channel slice out of range");
    if (M == 0 || Ccopy == 0) return I2V_OK;
    copy_channels_kernel<<<grid1d(M * (Ccopy / 4)), 256, 0, as_stream(stream)>>>(src, dst, M, Csrc, src_off, Cdst, dst_off, Ccopy, accumulate);
    I2V_LAUNCH_CHECK("i2v_copy_channels_f32");
    return I2V_OK;
}
