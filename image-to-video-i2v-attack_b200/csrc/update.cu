// K3 family: fused per-pixel update kernels of the attack loop (HBM-bound, 128-bit streaming).
//
//   K3a  i2v_adam_compose_f32        Adam on the modifier + compose/normalise backward and forward
//                                     (reference image_attacks.py:331-332, 351-353)
//   K3b  i2v_sign_step_project_f32   BIM/FGSM sign step, eps-projection, [0,1] clamp, re-normalise
//                                     (reference base_attacks.py:289-293, 254-257)
//   K3c  i2v_frame_absmean_f32 + i2v_mi_sign_step_project_f32   MI-FGSM (base_attacks.py:328-338,
//                                     utils.py:58-67)
//
// Rounding contract: every arithmetic op below is an explicit round-to-nearest intrinsic
// (__fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn/__fsqrt_rn, __fmaf_rn only where the oracle says `fma`),
// so the compiler can neither contract nor reassociate; results are bit-identical to
// oracle/i2v_oracle.c for any grid size.
#include "common.cuh"
#include <stdlib.h>

namespace i2v {

constexpr int kThreads = 256;

// Grid = SM count x CTAs that are actually co-resident per SM for THIS kernel (registers decide), so the
// grid-stride loop runs as exactly one full wave: a grid of 8 CTAs/SM when only 6 fit leaves a
// one-third-full second wave (measured: 1.33 waves, 61 % warps active on K3a).
template <typename K>
static inline int grid_for(K kernel, int64_t nvec) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kThreads, 0) != cudaSuccess || occ < 1) {
        cudaGetLastError();
        occ = 4;
    }
    int64_t want = (nvec + kThreads - 1) / kThreads;
    int64_t cap = (int64_t)sm_count() * occ;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

// Channel of scalar element i for layout (inner, C).
template <typename IdxT>
__device__ __forceinline__ int chan_of(IdxT i, IdxT inner, int C) {
    return (int)((i / inner) % (IdxT)C);
}

// Generic driver: F is called as f(lane_channel, element_index_in_vector, vector_index) on float4 data.
// UNIFORM: inner % 4 == 0, so the four lanes of a vector share a channel.
template <typename IdxT, bool UNIFORM>
struct ChanIter {
    IdxT inner; int C;
    __device__ __forceinline__ void channels(IdxT iv, int c[4]) const {
        IdxT i0 = iv * 4;
        if (UNIFORM) {
            int cc = chan_of<IdxT>(i0, inner, C);
            c[0] = c[1] = c[2] = c[3] = cc;
        } else if (inner == 1) {
            int cc = (int)(i0 % (IdxT)C);
#pragma unroll
            for (int j = 0; j < 4; ++j) { c[j] = cc; cc = (cc + 1 == C) ? 0 : cc + 1; }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) c[j] = chan_of<IdxT>(i0 + j, inner, C);
        }
    }
};

__device__ __forceinline__ float& lane(float4& v, int j) { return (&v.x)[j]; }
__device__ __forceinline__ const float& lane(const float4& v, int j) { return (&v.x)[j]; }

// ---------------------------------------------------------------------------------------------
// element-level arithmetic (shared by the vector body and the scalar tail)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float denorm1(float inp, int c) {
    return __fadd_rn(__fmul_rn(inp, chan_std(c)), chan_mean(c));
}
__device__ __forceinline__ float norm1(float t, int c) {
    return __fdiv_rn(__fsub_rn(t, chan_mean(c)), chan_std(c));
}
__device__ __forceinline__ float compose1(float x, float mod, float eps, int c) {
    float mc = clampf(mod, -eps, eps);
    float t = clampf(__fadd_rn(x, mc), 0.0f, 1.0f);
    return norm1(t, c);
}

struct AdamScalars {
    float w1;        // f32(1 - beta1)
    float beta2;     // f32(beta2)
    float a2;        // f32(1 - beta2)
    float adam_eps;  // f32(1e-8)
    float bc2_sqrt;  // f32(sqrt(1 - beta2^t))
    float neg_ss;    // f32(-lr / (1 - beta1^t))
    int cuda_arith;  // 1: torch's CUDA (foreach) Adam kernels, 0: torch's CPU Adam kernels (i2v_set_adam_arithmetic)
};

__device__ __forceinline__ void adam1(float g, float& m, float& v, float& mod, float x, float& out,
                                      float eps, const AdamScalars& s, int c, bool pad_lane) {
    if (pad_lane) { out = 0.0f; return; }   // NHWC4 padding channel: state stays 0
    const float sd = chan_std(c);
    float mc = clampf(mod, -eps, eps);
    float sum = __fadd_rn(x, mc);
    bool inside = (sum >= 0.0f) && (sum <= 1.0f) && (mod >= -eps) && (mod <= eps);
    float gm = __fmul_rn(__fdiv_rn(g, sd), inside ? 1.0f : 0.0f);
    float m2 = __fmaf_rn(s.w1, __fsub_rn(gm, m), m);
    // The two torch back ends group addcmul_ / addcdiv_ differently (probed bit for bit: tools/adam_cuda_probe.py,
    // profiles/r02_adam_cuda_probe.json).  CUDA (what the reference runs, `.cuda()` is hard-coded at image_attacks.py:304):
    //   v = fma(a2, g*g, v*b2);  mod = fma(-ss, m / den, mod)     CPU:  v = fma(a2*g, g, v*b2);  mod = mod + (-ss*m) / den
    float v2, mod2;
    if (s.cuda_arith) {
        v2 = __fmaf_rn(s.a2, __fmul_rn(gm, gm), __fmul_rn(v, s.beta2));
        float den = __fadd_rn(__fdiv_rn(__fsqrt_rn(v2), s.bc2_sqrt), s.adam_eps);
        mod2 = __fmaf_rn(s.neg_ss, __fdiv_rn(m2, den), mod);
    } else {
        v2 = __fmaf_rn(__fmul_rn(s.a2, gm), gm, __fmul_rn(v, s.beta2));
        float den = __fadd_rn(__fdiv_rn(__fsqrt_rn(v2), s.bc2_sqrt), s.adam_eps);
        mod2 = __fadd_rn(mod, __fdiv_rn(__fmul_rn(s.neg_ss, m2), den));
    }
    m = m2; v = v2; mod = mod2;
    out = compose1(x, mod2, eps, c);
}

__device__ __forceinline__ float signf(float g) { return g > 0.0f ? 1.0f : (g < 0.0f ? -1.0f : 0.0f); }

__device__ __forceinline__ float sign_step1(float adv, float sg, float x, float step, float eps, int project, int c) {
    float a = denorm1(adv, c);
    a = __fadd_rn(a, __fmul_rn(step, sg));
    if (project) {
        float d = clampf(__fsub_rn(a, x), -eps, eps);
        a = clampf(__fadd_rn(x, d), 0.0f, 1.0f);
    } else {
        a = clampf(a, 0.0f, 1.0f);
    }
    return norm1(a, c);
}

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
template <typename IdxT, bool UNIFORM>
__global__ void __launch_bounds__(kThreads) denorm_kernel(const float* __restrict__ inp, float* __restrict__ x,
                                                          IdxT n, IdxT inner, int C) {
    ChanIter<IdxT, UNIFORM> it{inner, C};
    const IdxT nv = n / 4;
    const float4* in4 = reinterpret_cast<const float4*>(inp);
    float4* x4 = reinterpret_cast<float4*>(x);
    for (IdxT iv = (IdxT)blockIdx.x * kThreads + threadIdx.x; iv < nv; iv += (IdxT)gridDim.x * kThreads) {
        int c[4]; it.channels(iv, c);
        float4 a = ld_stream(in4 + iv), r;
#pragma unroll
        for (int j = 0; j < 4; ++j) lane(r, j) = (C == 4 && c[j] == 3) ? 0.0f : denorm1(lane(a, j), c[j]);
        st_stream(x4 + iv, r);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (IdxT i = nv * 4; i < n; ++i) x[i] = denorm1(inp[i], chan_of<IdxT>(i, inner, C));
}

template <typename IdxT, bool UNIFORM>
__global__ void __launch_bounds__(kThreads) normalize_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                             IdxT n, IdxT inner, int C) {
    ChanIter<IdxT, UNIFORM> it{inner, C};
    const IdxT nv = n / 4;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* o4 = reinterpret_cast<float4*>(out);
    for (IdxT iv = (IdxT)blockIdx.x * kThreads + threadIdx.x; iv < nv; iv += (IdxT)gridDim.x * kThreads) {
        int c[4]; it.channels(iv, c);
        float4 a = ld_stream(x4 + iv), r;
#pragma unroll
        for (int j = 0; j < 4; ++j) lane(r, j) = (C == 4 && c[j] == 3) ? 0.0f : norm1(lane(a, j), c[j]);
        st_stream(o4 + iv, r);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (IdxT i = nv * 4; i < n; ++i) out[i] = norm1(x[i], chan_of<IdxT>(i, inner, C));
}

template <typename IdxT, bool UNIFORM>
__global__ void __launch_bounds__(kThreads) compose_kernel(const float* __restrict__ x, const float* __restrict__ mod,
                                                           float* __restrict__ out, IdxT n, IdxT inner, int C, float eps) {
    ChanIter<IdxT, UNIFORM> it{inner, C};
    const IdxT nv = n / 4;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const float4* m4 = reinterpret_cast<const float4*>(mod);
    float4* o4 = reinterpret_cast<float4*>(out);
    for (IdxT iv = (IdxT)blockIdx.x * kThreads + threadIdx.x; iv < nv; iv += (IdxT)gridDim.x * kThreads) {
        int c[4]; it.channels(iv, c);
        float4 a = ld_stream(x4 + iv), b = ld_stream(m4 + iv), r;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            lane(r, j) = (C == 4 && c[j] == 3) ? 0.0f : compose1(lane(a, j), lane(b, j), eps, c[j]);
        st_stream(o4 + iv, r);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (IdxT i = nv * 4; i < n; ++i) out[i] = compose1(x[i], mod[i], eps, chan_of<IdxT>(i, inner, C));
}

__global__ void __launch_bounds__(kThreads) fill_kernel(float* __restrict__ p, float value, int64_t n) {
    const int64_t nv = n / 4;
    float4* p4 = reinterpret_cast<float4*>(p);
    float4 v = make_float4(value, value, value, value);
    for (int64_t iv = (int64_t)blockIdx.x * kThreads + threadIdx.x; iv < nv; iv += (int64_t)gridDim.x * kThreads)
        st_stream(p4 + iv, v);
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (int64_t i = nv * 4; i < n; ++i) p[i] = value;
}

// K3a.  TABLE: step scalars come from step_table[2 * *step_idx + {0,1}] (CUDA-graph replay).
template <typename IdxT, bool UNIFORM, bool TABLE>
__global__ void __launch_bounds__(kThreads)
adam_compose_kernel(const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                    float* __restrict__ mod, const float* __restrict__ x, float* __restrict__ out,
                    IdxT n, IdxT inner, int C, float eps, AdamScalars s,
                    const float* __restrict__ step_table, const int* __restrict__ step_idx) {
    if (TABLE) {
        int k = *step_idx;
        s.bc2_sqrt = step_table[2 * k];
        s.neg_ss = step_table[2 * k + 1];
    }
    ChanIter<IdxT, UNIFORM> it{inner, C};
    const IdxT nv = n / 4;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    float4* d4 = reinterpret_cast<float4*>(mod);
    float4* o4 = reinterpret_cast<float4*>(out);
    for (IdxT iv = (IdxT)blockIdx.x * kThreads + threadIdx.x; iv < nv; iv += (IdxT)gridDim.x * kThreads) {
        int c[4]; it.channels(iv, c);
        // five independent 16-byte loads in flight per thread before the first use
        float4 gg = ld_stream(g4 + iv);
        float4 xx = ld_stream(x4 + iv);
        float4 mm = ld_plain(m4 + iv);
        float4 vv = ld_plain(v4 + iv);
        float4 dd = ld_plain(d4 + iv);
        float4 oo;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            adam1(lane(gg, j), lane(mm, j), lane(vv, j), lane(dd, j), lane(xx, j), lane(oo, j), eps, s, c[j],
                  C == 4 && c[j] == 3);
        m4[iv] = mm; v4[iv] = vv; d4[iv] = dd;
        st_stream(o4 + iv, oo);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (IdxT i = nv * 4; i < n; ++i)
            adam1(g[i], m[i], v[i], mod[i], x[i], out[i], eps, s, chan_of<IdxT>(i, inner, C), false);
}

// K3d: ILAF's update (image_attacks.py:615-617 + 582-585 of the next iteration): sign-gradient DESCENT on the modifier,
// fused with the backward of the compose/normalise block in front of it and the compose/normalise behind it.
__device__ __forceinline__ void sign_descent1(float g, float& mod, float x, float& out, float eps, float step, int c) {
    const float sd = chan_std(c);
    float mc = clampf(mod, -eps, eps);
    float sum = __fadd_rn(x, mc);
    bool inside = (sum >= 0.0f) && (sum <= 1.0f) && (mod >= -eps) && (mod <= eps);
    float gm = __fmul_rn(__fdiv_rn(g, sd), inside ? 1.0f : 0.0f);
    float mod2 = __fsub_rn(mod, __fmul_rn(step, signf(gm)));
    mod = mod2;
    out = compose1(x, mod2, eps, c);
}

template <typename IdxT, bool UNIFORM>
__global__ void __launch_bounds__(kThreads)
sign_descent_compose_kernel(const float* __restrict__ g, float* __restrict__ mod, const float* __restrict__ x,
                            float* __restrict__ out, IdxT n, IdxT inner, int C, float eps, float step) {
    ChanIter<IdxT, UNIFORM> it{inner, C};
    const IdxT nv = n / 4;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* d4 = reinterpret_cast<float4*>(mod);
    float4* o4 = reinterpret_cast<float4*>(out);
    for (IdxT iv = (IdxT)blockIdx.x * kThreads + threadIdx.x; iv < nv; iv += (IdxT)gridDim.x * kThreads) {
        int c[4]; it.channels(iv, c);
        float4 gg = ld_stream(g4 + iv);
        float4 xx = ld_stream(x4 + iv);
        float4 dd = ld_plain(d4 + iv);
        float4 oo;
#pragma unroll
        for (int j = 0; j < 4; ++j) sign_descent1(lane(gg, j), lane(dd, j), lane(xx, j), lane(oo, j), eps, step, c[j]);
        d4[iv] = dd;
        st_stream(o4 + iv, oo);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (IdxT i = nv * 4; i < n; ++i) sign_descent1(g[i], mod[i], x[i], out[i], eps, step, chan_of<IdxT>(i, inner, C));
}

// K3b
template <typename IdxT, bool UNIFORM>
__global__ void __launch_bounds__(kThreads)
sign_step_kernel(float* __restrict__ adv, const float* __restrict__ g, const float* __restrict__ x, IdxT n,
                 IdxT inner, int C, float step, float eps, int project) {
    ChanIter<IdxT, UNIFORM> it{inner, C};
    const IdxT nv = n / 4;
    float4* a4 = reinterpret_cast<float4*>(adv);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (IdxT iv = (IdxT)blockIdx.x * kThreads + threadIdx.x; iv < nv; iv += (IdxT)gridDim.x * kThreads) {
        int c[4]; it.channels(iv, c);
        float4 aa = ld_plain(a4 + iv);
        float4 gg = ld_stream(g4 + iv);
        float4 xx = project ? ld_stream(x4 + iv) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 r;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            lane(r, j) = sign_step1(lane(aa, j), signf(lane(gg, j)), lane(xx, j), step, eps, project, c[j]);
        a4[iv] = r;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (IdxT i = nv * 4; i < n; ++i)
            adv[i] = sign_step1(adv[i], signf(g[i]), project ? x[i] : 0.f, step, eps, project, chan_of<IdxT>(i, inner, C));
}

// K3c part 1: one CTA per (b,t) [frame level] or per b [clip level]; FP64 accumulation, fixed order.
__global__ void __launch_bounds__(512)
frame_absmean_kernel(const float* __restrict__ g, float* __restrict__ norm, int C, int T, int64_t HW, int clip_level) {
    const int b = clip_level ? blockIdx.x : blockIdx.x / T;
    const int t0 = clip_level ? 0 : blockIdx.x % T;
    const int nt = clip_level ? T : 1;
    double acc = 0.0;
    for (int c = 0; c < C; ++c)
        for (int t = t0; t < t0 + nt; ++t) {
            const float* p = g + (((int64_t)b * C + c) * T + t) * HW;
            if ((HW & 3) == 0) {
                const float4* p4 = reinterpret_cast<const float4*>(p);
                for (int64_t i = threadIdx.x; i < HW / 4; i += blockDim.x) {
                    float4 q = ld_plain(p4 + i);   // re-read by the step kernel right after: keep in L1/L2
                    acc += (double)fabsf(q.x) + (double)fabsf(q.y) + (double)fabsf(q.z) + (double)fabsf(q.w);
                }
            } else {
                for (int64_t i = threadIdx.x; i < HW; i += blockDim.x) acc += (double)fabsf(p[i]);
            }
        }
    __shared__ double part[16];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += part[w];
        norm[blockIdx.x] = (float)(s / (double)((int64_t)C * nt * HW));
    }
}

// K3c part 2
__global__ void __launch_bounds__(kThreads)
mi_step_kernel(float* __restrict__ adv, const float* __restrict__ g, float* __restrict__ mom,
               const float* __restrict__ norm, const float* __restrict__ x, int C, int T, int64_t HW,
               int clip_level, float decay, float step, float eps, int64_t n) {
    const int64_t inner = (int64_t)T * HW;
    const bool vec = (HW & 3) == 0;
    const int64_t nv = vec ? n / 4 : 0;
    if (vec) {
        float4* a4 = reinterpret_cast<float4*>(adv);
        const float4* g4 = reinterpret_cast<const float4*>(g);
        float4* m4 = reinterpret_cast<float4*>(mom);
        const float4* x4 = reinterpret_cast<const float4*>(x);
        for (int64_t iv = (int64_t)blockIdx.x * kThreads + threadIdx.x; iv < nv; iv += (int64_t)gridDim.x * kThreads) {
            int64_t i0 = iv * 4;
            int c = (int)((i0 / inner) % C);
            int64_t b = i0 / (inner * C);
            int t = (int)((i0 / HW) % T);
            float nrm = norm[clip_level ? b : b * T + t];
            float4 aa = ld_plain(a4 + iv), gg = ld_stream(g4 + iv), mm = ld_plain(m4 + iv), xx = ld_stream(x4 + iv), r;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float gn = __fdiv_rn(lane(gg, j), nrm);
                gn = __fadd_rn(gn, __fmul_rn(lane(mm, j), decay));
                lane(mm, j) = gn;
                lane(r, j) = sign_step1(lane(aa, j), signf(gn), lane(xx, j), step, eps, 1, c);
            }
            m4[iv] = mm; a4[iv] = r;
        }
    } else {
        for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
            int c = (int)((i / inner) % C);
            int64_t b = i / (inner * C);
            int t = (int)((i / HW) % T);
            float nrm = norm[clip_level ? b : b * T + t];
            float gn = __fadd_rn(__fdiv_rn(g[i], nrm), __fmul_rn(mom[i], decay));
            mom[i] = gn;
            adv[i] = sign_step1(adv[i], signf(gn), x[i], step, eps, 1, c);
        }
    }
}

__global__ void step_advance_kernel(int* step_idx) { *step_idx += 1; }

// ---------------------------------------------------------------------------------------------
// host-side dispatch
// ---------------------------------------------------------------------------------------------
static int check_layout(const void* p0, int64_t n, int64_t inner, int channels) {
    if (n == 0) return I2V_OK;   // empty batch: nothing to check, nothing to launch (pointers may be null)
    I2V_REQUIRE(p0 != nullptr, "null tensor pointer");
    I2V_REQUIRE(n >= 0 && inner >= 1, "bad sizes n=%lld inner=%lld", (long long)n, (long long)inner);
    I2V_REQUIRE(channels == 3 || channels == 4, "channels must be 3 or 4 (got %d)", channels);
    I2V_REQUIRE(channels == 3 || inner == 1, "channels=4 is the NHWC4 layout and needs inner=1");
    I2V_REQUIRE(n % (inner * channels) == 0, "n=%lld is not a multiple of inner*channels", (long long)n);
    I2V_REQUIRE((reinterpret_cast<uintptr_t>(p0) & 15) == 0, "tensor pointer must be 16-byte aligned");
    return I2V_OK;
}
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace i2v

using namespace i2v;

extern "C" int i2v_denorm_f32(const float* inp, float* x, int64_t n, int64_t inner, int channels, i2v_stream_t stream) {
    if (int r = check_layout(inp, n, inner, channels)) return r;
    if (n == 0) return I2V_OK;
    I2V_REQUIRE(x && aligned16(x), "x must be a 16-byte aligned device pointer");
    cudaStream_t st = as_stream(stream);
    if (n < (int64_t)0x7fffffff) {
        if (inner % 4 == 0) denorm_kernel<uint32_t, true><<<grid_for(denorm_kernel<uint32_t, true>, n / 4), kThreads, 0, st>>>(inp, x, (uint32_t)n, (uint32_t)inner, channels);
        else denorm_kernel<uint32_t, false><<<grid_for(denorm_kernel<uint32_t, false>, n / 4), kThreads, 0, st>>>(inp, x, (uint32_t)n, (uint32_t)inner, channels);
    } else {
        if (inner % 4 == 0) denorm_kernel<uint64_t, true><<<grid_for(denorm_kernel<uint64_t, true>, n / 4), kThreads, 0, st>>>(inp, x, (uint64_t)n, (uint64_t)inner, channels);
        else denorm_kernel<uint64_t, false><<<grid_for(denorm_kernel<uint64_t, false>, n / 4), kThreads, 0, st>>>(inp, x, (uint64_t)n, (uint64_t)inner, channels);
    }
    I2V_LAUNCH_CHECK("i2v_denorm_f32");
    return I2V_OK;
}

extern "C" int i2v_normalize_f32(const float* x, float* out, int64_t n, int64_t inner, int channels, i2v_stream_t stream) {
    if (int r = check_layout(x, n, inner, channels)) return r;
    if (n == 0) return I2V_OK;
    I2V_REQUIRE(out && aligned16(out), "out must be a 16-byte aligned device pointer");
    cudaStream_t st = as_stream(stream);
    if (n < (int64_t)0x7fffffff) {
        if (inner % 4 == 0) normalize_kernel<uint32_t, true><<<grid_for(normalize_kernel<uint32_t, true>, n / 4), kThreads, 0, st>>>(x, out, (uint32_t)n, (uint32_t)inner, channels);
        else normalize_kernel<uint32_t, false><<<grid_for(normalize_kernel<uint32_t, false>, n / 4), kThreads, 0, st>>>(x, out, (uint32_t)n, (uint32_t)inner, channels);
    } else {
        if (inner % 4 == 0) normalize_kernel<uint64_t, true><<<grid_for(normalize_kernel<uint64_t, true>, n / 4), kThreads, 0, st>>>(x, out, (uint64_t)n, (uint64_t)inner, channels);
        else normalize_kernel<uint64_t, false><<<grid_for(normalize_kernel<uint64_t, false>, n / 4), kThreads, 0, st>>>(x, out, (uint64_t)n, (uint64_t)inner, channels);
    }
    I2V_LAUNCH_CHECK("i2v_normalize_f32");
    return I2V_OK;
}

extern "C" int i2v_compose_norm_f32(const float* x, const float* mod, float* out, int64_t n, int64_t inner,
                                    int channels, float eps, i2v_stream_t stream) {
    if (int r = check_layout(x, n, inner, channels)) return r;
    if (n == 0) return I2V_OK;
    I2V_REQUIRE(mod && out && aligned16(mod) && aligned16(out), "mod/out must be 16-byte aligned device pointers");
    cudaStream_t st = as_stream(stream);
    if (n < (int64_t)0x7fffffff) {
        if (inner % 4 == 0) compose_kernel<uint32_t, true><<<grid_for(compose_kernel<uint32_t, true>, n / 4), kThreads, 0, st>>>(x, mod, out, (uint32_t)n, (uint32_t)inner, channels, eps);
        else compose_kernel<uint32_t, false><<<grid_for(compose_kernel<uint32_t, false>, n / 4), kThreads, 0, st>>>(x, mod, out, (uint32_t)n, (uint32_t)inner, channels, eps);
    } else {
        if (inner % 4 == 0) compose_kernel<uint64_t, true><<<grid_for(compose_kernel<uint64_t, true>, n / 4), kThreads, 0, st>>>(x, mod, out, (uint64_t)n, (uint64_t)inner, channels, eps);
        else compose_kernel<uint64_t, false><<<grid_for(compose_kernel<uint64_t, false>, n / 4), kThreads, 0, st>>>(x, mod, out, (uint64_t)n, (uint64_t)inner, channels, eps);
    }
    I2V_LAUNCH_CHECK("i2v_compose_norm_f32");
    return I2V_OK;
}

extern "C" int i2v_fill_f32(float* p, float value, int64_t n, i2v_stream_t stream) {
    I2V_REQUIRE(p && aligned16(p) && n >= 0, "bad fill arguments");
    if (n == 0) return I2V_OK;
    fill_kernel<<<grid_for(fill_kernel, n / 4), kThreads, 0, as_stream(stream)>>>(p, value, n);
    I2V_LAUNCH_CHECK("i2v_fill_f32");
    return I2V_OK;
}

// Which of torch's two Adam arithmetics K3a reproduces bit for bit: 1 (default) = the CUDA foreach kernels the reference
// hits, 0 = the CPU kernels the committed fixtures were generated with.  $I2V_ADAM_ARITH=cpu|cuda sets the initial value.
static int g_adam_cuda_arith = -1;
static int adam_cuda_arith() {
    if (g_adam_cuda_arith < 0) {
        const char* e = getenv("I2V_ADAM_ARITH");
        g_adam_cuda_arith = (e && e[0] == 'c' && e[1] == 'p') ? 0 : 1;
    }
    return g_adam_cuda_arith;
}
extern "C" int i2v_set_adam_arithmetic(int cuda_arith) {
    I2V_REQUIRE(cuda_arith == 0 || cuda_arith == 1, "adam arithmetic must be 0 (torch CPU kernels) or 1 (torch CUDA kernels)");
    g_adam_cuda_arith = cuda_arith;
    return I2V_OK;
}
extern "C" int i2v_get_adam_arithmetic(void) { return adam_cuda_arith(); }

// The step scalars are formed exactly as torch.optim.adam._single_tensor_adam does for python-float
// hyper-parameters: double arithmetic on the host, rounded to f32 when they meet the f32 tensor.
static void adam_step_scalars(double lr, double beta1, double beta2, int step, float* bc2_sqrt, float* neg_ss) {
    double bc1 = 1.0 - pow(beta1, (double)step);
    double bc2 = 1.0 - pow(beta2, (double)step);
    *bc2_sqrt = (float)pow(bc2, 0.5);   // `bias_correction2 ** 0.5` in torch/optim/adam.py
    *neg_ss = (float)(-(lr / bc1));
}

extern "C" int i2v_adam_step_table(float* host_table, int steps, double lr, double beta1, double beta2) {
    I2V_REQUIRE(host_table && steps >= 0, "bad step table arguments");
    for (int k = 0; k < steps; ++k) adam_step_scalars(lr, beta1, beta2, k + 1, host_table + 2 * k, host_table + 2 * k + 1);
    return I2V_OK;
}

template <bool TABLE>
static int adam_launch(const float* g, float* m, float* v, float* mod, const float* x, float* out, int64_t n,
                       int64_t inner, int channels, float eps, AdamScalars s, const float* table, const int* idx,
                       cudaStream_t st) {
    const int grid = grid_for(adam_compose_kernel<uint32_t, true, TABLE>, n / 4);
    if (n < (int64_t)0x7fffffff) {
        if (inner % 4 == 0) adam_compose_kernel<uint32_t, true, TABLE><<<grid, kThreads, 0, st>>>(g, m, v, mod, x, out, (uint32_t)n, (uint32_t)inner, channels, eps, s, table, idx);
        else adam_compose_kernel<uint32_t, false, TABLE><<<grid, kThreads, 0, st>>>(g, m, v, mod, x, out, (uint32_t)n, (uint32_t)inner, channels, eps, s, table, idx);
    } else {
        if (inner % 4 == 0) adam_compose_kernel<uint64_t, true, TABLE><<<grid, kThreads, 0, st>>>(g, m, v, mod, x, out, (uint64_t)n, (uint64_t)inner, channels, eps, s, table, idx);
        else adam_compose_kernel<uint64_t, false, TABLE><<<grid, kThreads, 0, st>>>(g, m, v, mod, x, out, (uint64_t)n, (uint64_t)inner, channels, eps, s, table, idx);
    }
    I2V_LAUNCH_CHECK("i2v_adam_compose_f32");
    return I2V_OK;
}

extern "C" int i2v_adam_compose_f32(const float* g, float* m, float* v, float* mod, const float* x, float* next_img,
                                    int64_t n, int64_t inner, int channels, float eps, double lr, double beta1,
                                    double beta2, double adam_eps, int step, i2v_stream_t stream) {
    if (int r = check_layout(g, n, inner, channels)) return r;
    if (n == 0) return I2V_OK;
    I2V_REQUIRE(m && v && mod && x && next_img, "null state pointer");
    I2V_REQUIRE(aligned16(m) && aligned16(v) && aligned16(mod) && aligned16(x) && aligned16(next_img), "state pointers must be 16-byte aligned");
    I2V_REQUIRE(step >= 1, "Adam step is 1-based (got %d)", step);
    AdamScalars s;
    s.w1 = (float)(1.0 - beta1);
    s.beta2 = (float)beta2;
    s.a2 = (float)(1.0 - beta2);
    s.adam_eps = (float)adam_eps;
    s.cuda_arith = adam_cuda_arith();
    adam_step_scalars(lr, beta1, beta2, step, &s.bc2_sqrt, &s.neg_ss);
    return adam_launch<false>(g, m, v, mod, x, next_img, n, inner, channels, eps, s, nullptr, nullptr, as_stream(stream));
}

extern "C" int i2v_adam_compose_table_f32(const float* g, float* m, float* v, float* mod, const float* x,
                                          float* next_img, int64_t n, int64_t inner, int channels, float eps, float w1,
                                          float beta2, float a2, float adam_eps, const float* step_table,
                                          const int* step_idx, i2v_stream_t stream) {
    if (int r = check_layout(g, n, inner, channels)) return r;
    if (n == 0) return I2V_OK;
    I2V_REQUIRE(m && v && mod && x && next_img && step_table && step_idx, "null state pointer");
    I2V_REQUIRE(aligned16(m) && aligned16(v) && aligned16(mod) && aligned16(x) && aligned16(next_img), "state pointers must be 16-byte aligned");
    AdamScalars s{w1, beta2, a2, adam_eps, 0.f, 0.f, adam_cuda_arith()};
    return adam_launch<true>(g, m, v, mod, x, next_img, n, inner, channels, eps, s, step_table, step_idx, as_stream(stream));
}

extern "C" int i2v_step_advance(int* step_idx, i2v_stream_t stream) {
    I2V_REQUIRE(step_idx, "null step counter");
    step_advance_kernel<<<1, 1, 0, as_stream(stream)>>>(step_idx);
    I2V_LAUNCH_CHECK("i2v_step_advance");
    return I2V_OK;
}

extern "C" int i2v_sign_step_project_f32(float* adv, const float* g, const float* x, int64_t n, int64_t inner,
                                         int channels, float step_size, float eps, int project, i2v_stream_t stream) {
    if (int r = check_layout(adv, n, inner, channels)) return r;
    if (n == 0) return I2V_OK;
    I2V_REQUIRE(channels == 3, "sign-step kernels take the reference's 3-channel layouts only");
    I2V_REQUIRE(g && aligned16(g), "g must be a 16-byte aligned device pointer");
    I2V_REQUIRE(!project || (x && aligned16(x)), "x is required (16-byte aligned) when project != 0");
    cudaStream_t st = as_stream(stream);
    const int grid = grid_for(sign_step_kernel<uint32_t, true>, n / 4);
    if (n < (int64_t)0x7fffffff) {
        if (inner % 4 == 0) sign_step_kernel<uint32_t, true><<<grid, kThreads, 0, st>>>(adv, g, x, (uint32_t)n, (uint32_t)inner, channels, step_size, eps, project);
        else sign_step_kernel<uint32_t, false><<<grid, kThreads, 0, st>>>(adv, g, x, (uint32_t)n, (uint32_t)inner, channels, step_size, eps, project);
    } else {
        if (inner % 4 == 0) sign_step_kernel<uint64_t, true><<<grid, kThreads, 0, st>>>(adv, g, x, (uint64_t)n, (uint64_t)inner, channels, step_size, eps, project);
        else sign_step_kernel<uint64_t, false><<<grid, kThreads, 0, st>>>(adv, g, x, (uint64_t)n, (uint64_t)inner, channels, step_size, eps, project);
    }
    I2V_LAUNCH_CHECK("i2v_sign_step_project_f32");
    return I2V_OK;
}

extern "C" int i2v_sign_descent_compose_f32(const float* g, float* mod, const float* x, float* next_img, int64_t n,
                                           int64_t inner, int channels, float eps, float step_size, i2v_stream_t stream) {
    if (int r = check_layout(g, n, inner, channels)) return r;
    if (n == 0) return I2V_OK;
    I2V_REQUIRE(channels == 3, "sign-step kernels take the reference's 3-channel layouts only");
    I2V_REQUIRE(mod && x && next_img, "null state pointer");
    I2V_REQUIRE(aligned16(mod) && aligned16(x) && aligned16(next_img), "state pointers must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    const int grid = grid_for(sign_descent_compose_kernel<uint32_t, true>, n / 4);
    if (n < (int64_t)0x7fffffff) {
        if (inner % 4 == 0) sign_descent_compose_kernel<uint32_t, true><<<grid, kThreads, 0, st>>>(g, mod, x, next_img, (uint32_t)n, (uint32_t)inner, channels, eps, step_size);
        else sign_descent_compose_kernel<uint32_t, false><<<grid, kThreads, 0, st>>>(g, mod, x, next_img, (uint32_t)n, (uint32_t)inner, channels, eps, step_size);
    } else {
        if (inner % 4 == 0) sign_descent_compose_kernel<uint64_t, true><<<grid, kThreads, 0, st>>>(g, mod, x, next_img, (uint64_t)n, (uint64_t)inner, channels, eps, step_size);
        else sign_descent_compose_kernel<uint64_t, false><<<grid, kThreads, 0, st>>>(g, mod, x, next_img, (uint64_t)n, (uint64_t)inner, channels, eps, step_size);
    }
    I2V_LAUNCH_CHECK("i2v_sign_descent_compose_f32");
    return I2V_OK;
}

extern "C" int i2v_frame_absmean_f32(const float* g, float* norm, int B, int C, int T, int64_t HW, int clip_level,
                                     i2v_stream_t stream) {
    I2V_REQUIRE(g && norm, "null pointer");
    I2V_REQUIRE(B >= 0 && C >= 1 && T >= 1 && HW >= 1, "bad shape B=%d C=%d T=%d HW=%lld", B, C, T, (long long)HW);
    I2V_REQUIRE((HW & 3) != 0 || aligned16(g), "g must be 16-byte aligned");
    if (B == 0) return I2V_OK;
    const int blocks = clip_level ? B : B * T;
    frame_absmean_kernel<<<blocks, 512, 0, as_stream(stream)>>>(g, norm, C, T, HW, clip_level);
    I2V_LAUNCH_CHECK("i2v_frame_absmean_f32");
    return I2V_OK;
}

extern "C" int i2v_mi_sign_step_project_f32(float* adv, const float* g, float* momentum, const float* norm,
                                            const float* x, int B, int C, int T, int64_t HW, int clip_level, float decay,
                                            float step_size, float eps, i2v_stream_t stream) {
    I2V_REQUIRE(adv && g && momentum && norm && x, "null pointer");
    I2V_REQUIRE(C == 3, "MI update takes the reference's [B,3,T,H,W] layout (C=%d)", C);
    I2V_REQUIRE(B >= 0 && T >= 1 && HW >= 1, "bad shape");
    I2V_REQUIRE((HW & 3) != 0 || (aligned16(adv) && aligned16(g) && aligned16(momentum) && aligned16(x)), "tensors must be 16-byte aligned");
    const int64_t n = (int64_t)B * C * T * HW;
    if (n == 0) return I2V_OK;
    mi_step_kernel<<<grid_for(mi_step_kernel, (HW & 3) == 0 ? n / 4 : n), kThreads, 0, as_stream(stream)>>>(
        adv, g, momentum, norm, x, C, T, HW, clip_level, decay, step_size, eps, n);
    I2V_LAUNCH_CHECK("i2v_mi_sign_step_project_f32");
    return I2V_OK;
}
