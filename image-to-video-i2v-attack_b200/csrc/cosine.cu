// K1: per-frame cosine feature loss + analytic gradient; K2: adaptive layer re-weighting.
//
// K1 maps one thread-block CLUSTER to one frame (reference image_attacks.py:341-343 computes
// F.cosine_similarity over the flattened [N, D] feature map, one cosine per frame).  Every CTA of the
// cluster streams its slice of a and b once from HBM accumulating <a,b>, <a,a>, <b,b> in FP64, the
// three partials are exchanged through distributed shared memory (no global atomics, fixed order =>
// bit-reproducible), and the gradient is written in a second pass that never goes back to HBM:
// while streaming, each CTA stashes as much of its slice as fits in its ~200 KB of shared memory
// (with a 16-CTA cluster a whole ResNet layer2 frame, 2 x 1.6 MB, lives on-chip); what does not fit
// is the TAIL of the slice, which is walked backwards so the most recently read lines are still in
// the 126 MB L2 (one CTA per SM => at most 148/CLUSTER frames in flight).
// Algorithmic HBM bytes: 12*D per frame (read a, read b, write grad).
//
// Numerics (SURVEY.md 7.3 "step-1 numerics"): at step 1 cos ~ 1 and the gradient
//     g = b/(|a||b|) - cos * a/|a|^2
// is a catastrophic cancellation (|g| ~ 1e-8 of its terms).  Sums are FP64, the two scalars stay in
// FP64 and each gradient element is formed in FP64 and rounded once, so the result is the correctly
// rounded analytic gradient, i.e. at least as close to the float64 reference as torch-f32 autograd is.
#include <cooperative_groups.h>
#include <stdlib.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace i2v {

constexpr int kCosThreads = 512;
constexpr int kUnroll = 4;
constexpr double kCosEps = 1e-8;   // F.cosine_similarity default eps (image_attacks.py:343)

struct CosPartial { double dot, aa, bb; };
struct CosCoef { float ah, al, bh, bl; };   // alpha, beta as unevaluated float pairs (hi + lo)

// One gradient element, float32 only (no F2F conversions: the XU pipe was 34 % busy with them):
//   g = alpha*b - beta*a   with alpha = ah+al, beta = bh+bl (each pair carries ~48 bits)
// t + e == bh*a exactly (e is the rounding error of the product, recovered by an FMA), so the
// cancellation alpha*b - beta*a happens inside one FMA on exact operands; the result is within
// ~2 ulp of the correctly rounded float64 value even at step 1 where |g| ~ 1e-8 |alpha*b|.
__device__ __forceinline__ float cos_grad1(float a, float b, const CosCoef& c) {
    const float t = __fmul_rn(c.bh, a);
    const float e = __fmaf_rn(c.bh, a, -t);
    const float g1 = __fmaf_rn(c.ah, b, -t);
    const float lo = __fmaf_rn(c.al, b, -__fmul_rn(c.bl, a));
    return __fadd_rn(__fsub_rn(g1, e), lo);
}

// THREADS = 512 with two CTAs per SM (each stashes half an SM's shared memory) or 1024 with ONE CTA per SM that owns the
// whole 227 KB: same threads and loads in flight per SM, twice the stash per frame slice, half the frames in flight.
template <bool VEC, int THREADS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS)
cosine_loss_grad_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ grad,
                        float* __restrict__ cos_out, int64_t D, const float* __restrict__ w_dev, float w_host,
                        int relu_mask, int64_t cap, int64_t N, int hint_mode) {
    extern __shared__ float4 stash[];   // [2][cap]: this CTA's slice of a and of b (as much as fits)
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned S = cluster.num_blocks();
    const unsigned rank = cluster.block_rank();
    __shared__ CosPartial warp_part[THREADS / 32];
    __shared__ CosPartial cta_part;
    __shared__ CosCoef coef;
    // L2 hints: what the gradient pass will re-read from L2 (the un-stashed tail of the slice) is loaded evict_last,
    // everything else — stashed loads, the re-reads themselves (last use), the gradient stores — evict_first.  ncu on
    // the un-hinted kernel: 1.47 GB of DRAM traffic for 1.23 GB of algorithmic bytes, i.e. most tail re-reads missed;
    // with the hints 384 -> 336 us per 256 frames inside the attack step (49 % -> 56 % of the copy peak), 356 -> 309 us alone.
    const uint64_t pol_stream = l2_policy_evict_first();
    const uint64_t pol_keep = grad ? l2_policy_evict_last() : pol_stream;     // loss only: nothing is read twice
    const bool hints = hint_mode != 0;
    // Persistent clusters: the grid holds as many clusters as are co-resident and each walks over frames
    // with that stride, so no SM slot idles waiting for 16 free slots in one GPC between frames.
    // Split cluster barrier: after the coefficients are published CTA-wide the threads only ARRIVE at the cluster barrier
    // (their remote reads of the other CTAs' partials are done) and go straight into the gradient pass; the matching
    // WAIT comes right before this CTA overwrites its partial for the next frame (or exits).  One cluster-wide barrier
    // latency per frame leaves the critical path.
    bool pending = false;
    for (int64_t frame = blockIdx.x / S; frame < N; frame += gridDim.x / S) {
    const float* af = a + frame * D;
    const float* bf = b + frame * D;

    // slice of this CTA, in units of float4 (VEC) or float
    const int64_t units = VEC ? D / 4 : D;
    const int64_t lo = units * rank / S;
    const int64_t hi = units * (rank + 1) / S;

    // Per-thread partial sums stay in float32: a thread sees only slice/512 (~50-100) elements per lane
    // group, four independent accumulators per quantity; the ~1e-7 relative rounding of each partial is
    // random across the 8K threads of a frame, so the frame sums (combined in FP64 below) are good to
    // ~2e-9 relative — two orders better than the 1e-5 the loss needs.
    float d4[4] = {0.f, 0.f, 0.f, 0.f}, p4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
    if (VEC) {
        const float4* a4 = reinterpret_cast<const float4*>(af);
        const float4* b4 = reinterpret_cast<const float4*>(bf);
        float4* sa = stash;
        float4* sb = stash + cap;
        int64_t i = lo + threadIdx.x;
        // kUnroll independent 16-byte loads per tensor in flight per thread
        for (; i + (int64_t)(kUnroll - 1) * THREADS < hi; i += (int64_t)kUnroll * THREADS) {
            float4 av[kUnroll], bv[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const uint64_t pol = (i + u * THREADS - lo < cap) ? pol_stream : pol_keep;
                av[u] = hints ? ld_hint(a4 + i + u * THREADS, pol) : ld_stream(a4 + i + u * THREADS);
                bv[u] = hints ? ld_hint(b4 + i + u * THREADS, pol) : ld_stream(b4 + i + u * THREADS);
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int64_t k = i + u * THREADS - lo;
                if (k < cap) { sa[k] = av[u]; sb[k] = bv[u]; }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float x = (&av[u].x)[j], y = (&bv[u].x)[j];
                    d4[j] = __fmaf_rn(x, y, d4[j]); p4[j] = __fmaf_rn(x, x, p4[j]); q4[j] = __fmaf_rn(y, y, q4[j]);
                }
            }
        }
        for (; i < hi; i += THREADS) {
            const int64_t k = i - lo;
            const uint64_t pol = (k < cap) ? pol_stream : pol_keep;
            float4 av = hints ? ld_hint(a4 + i, pol) : ld_stream(a4 + i), bv = hints ? ld_hint(b4 + i, pol) : ld_stream(b4 + i);
            if (k < cap) { sa[k] = av; sb[k] = bv; }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float x = (&av.x)[j], y = (&bv.x)[j];
                d4[j] = __fmaf_rn(x, y, d4[j]); p4[j] = __fmaf_rn(x, x, p4[j]); q4[j] = __fmaf_rn(y, y, q4[j]);
            }
        }
    } else {
        for (int64_t i = lo + threadIdx.x; i < hi; i += THREADS) {
            const float x = af[i], y = bf[i];
            d4[0] = __fmaf_rn(x, y, d4[0]); p4[0] = __fmaf_rn(x, x, p4[0]); q4[0] = __fmaf_rn(y, y, q4[0]);
        }
    }
    double dot = ((double)d4[0] + (double)d4[1]) + ((double)d4[2] + (double)d4[3]);
    double aa = ((double)p4[0] + (double)p4[1]) + ((double)p4[2] + (double)p4[3]);
    double bb = ((double)q4[0] + (double)q4[1]) + ((double)q4[2] + (double)q4[3]);

    // CTA reduction: warp shuffles, then the 16 warp partials by warp 0 with the same fixed-shape butterfly; the cluster
    // partials likewise — every lane fetches one rank's partial through DSMEM in parallel (the first version had thread 0
    // walk the ranks serially while 8K threads waited at the barrier: 30 % of the stall samples were barrier waits).
    // The butterfly's grouping is the same in every lane and every CTA, so the result is bit-reproducible.
    dot = warp_sum(dot); aa = warp_sum(aa); bb = warp_sum(bb);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = CosPartial{dot, aa, bb};
    __syncthreads();
    if (pending) { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); pending = false; }
    if (threadIdx.x < 32) {
        CosPartial s{0.0, 0.0, 0.0};
        if (threadIdx.x < THREADS / 32) s = warp_part[threadIdx.x];
        s.dot = warp_sum(s.dot); s.aa = warp_sum(s.aa); s.bb = warp_sum(s.bb);
        if (threadIdx.x == 0) cta_part = s;
    }
    cluster.sync();   // every CTA's partial is visible cluster-wide

    if (threadIdx.x < 32) {
        CosPartial t{0.0, 0.0, 0.0};
        if (threadIdx.x < S) t = *cluster.map_shared_rank(&cta_part, threadIdx.x);
        t.dot = warp_sum(t.dot); t.aa = warp_sum(t.aa); t.bb = warp_sum(t.bb);
        if (threadIdx.x == 0) {
            const double na_raw = sqrt(t.aa), nb_raw = sqrt(t.bb);
            const double na = fmax(na_raw, kCosEps), nb = fmax(nb_raw, kCosEps);
            const double inv = 1.0 / (na * nb);
            const double cosv = t.dot * inv;
            const double w = (double)(w_dev ? *w_dev : w_host);
            const double alpha = w * inv;                                             // multiplies b
            const double beta = (na_raw > kCosEps) ? w * cosv / (na * na) : 0.0;      // multiplies a (0 when |a| is clamped)
            CosCoef c;
            c.ah = (float)alpha; c.al = (float)(alpha - (double)c.ah);
            c.bh = (float)beta;  c.bl = (float)(beta - (double)c.bh);
            coef = c;
            if (rank == 0 && cos_out) cos_out[frame] = (float)cosv;
        }
    }
    __syncthreads();  // publishes coef; warp 0's remote reads precede its arrival below
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    pending = true;
    if (grad == nullptr) continue;

    const CosCoef c = coef;
    float* gf = grad + frame * D;
    if (VEC) {
        const float4* a4 = reinterpret_cast<const float4*>(af);
        const float4* b4 = reinterpret_cast<const float4*>(bf);
        float4* g4 = reinterpret_cast<float4*>(gf);
        const float4* sa = stash;
        const float4* sb = stash + cap;
        // backwards: the un-stashed tail of the slice was read last in pass 1 and is the hottest in L2;
        // each thread revisits exactly the vectors it stashed itself (same i mod 512), so the stash
        // needs no barrier of its own.
        const int64_t last = lo + ((hi - 1 - lo - threadIdx.x) / THREADS) * THREADS + threadIdx.x;
#pragma unroll 4
        for (int64_t i = (hi - lo > (int64_t)threadIdx.x) ? last : lo - 1; i >= lo; i -= THREADS) {
            const int64_t k = i - lo;
            float4 av, bv, r;
            if (k < cap) { av = sa[k]; bv = sb[k]; }
            else if (hints) { av = ld_hint(a4 + i, pol_stream); bv = ld_hint(b4 + i, pol_stream); }
            else { av = ld_plain(a4 + i); bv = ld_plain(b4 + i); }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float x = (&av.x)[j];
                const float gval = cos_grad1(x, (&bv.x)[j], c);
                (&r.x)[j] = (relu_mask && !(x > 0.0f)) ? 0.0f : gval;
            }
            if (hints) st_hint(g4 + i, r, pol_stream); else st_stream(g4 + i, r);
        }
    } else {
        for (int64_t i = hi - 1 - threadIdx.x; i >= lo; i -= THREADS) {
            const float x = af[i];
            const float gval = cos_grad1(x, bf[i], c);
            gf[i] = (relu_mask && !(x > 0.0f)) ? 0.0f : gval;
        }
    }
    }   // frames
    if (pending) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");   // nobody reads an exited CTA's partial
}

// K2: one warp.  coeffs <- softmax(softmax(prev) + momentum*coeffs)   (TPAMI_attack.py:265)
__global__ void layer_reweight_kernel(float* __restrict__ coeffs, const float* __restrict__ prev, int L,
                                      float momentum, float* __restrict__ w_out, float* __restrict__ weights_log,
                                      const int* __restrict__ step_idx) {
    const int l = threadIdx.x;
    const bool on = l < L;
    // inner softmax over prev
    double p = on ? (double)prev[l] : -INFINITY;
    double mx = p;
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    double e = on ? exp(p - mx) : 0.0;
    double s1 = (float)(e / warp_sum(e));          // torch materialises the inner softmax as f32
    // outer softmax over s1 + momentum*coeffs  (momentum*coeffs and the add are f32 ops in torch)
    float mc = on ? __fmul_rn(momentum, coeffs[l]) : 0.0f;
    double q = on ? (double)__fadd_rn((float)s1, mc) : -INFINITY;
    double mx2 = q;
    for (int o = 16; o > 0; o >>= 1) mx2 = fmax(mx2, __shfl_xor_sync(0xffffffffu, mx2, o));
    double e2 = on ? exp(q - mx2) : 0.0;
    float c = (float)(e2 / warp_sum(e2));
    if (on) {
        coeffs[l] = c;
        if (w_out) w_out[l] = __fmul_rn(__fdiv_rn(1.0f, (float)L), c);   // d mean_l / d each_l = 1/L, times coeffs[l]
        if (weights_log) weights_log[(int64_t)(step_idx ? *step_idx : 0) * L + l] = c;
    }
}

// Layer sums / cost / prev update.  One CTA; row sums in FP64 in a fixed order.
__global__ void __launch_bounds__(256)
layer_sums_kernel(const float* __restrict__ cosv, const float* __restrict__ coeffs, float* __restrict__ prev,
                  float* __restrict__ cost_log, const int* __restrict__ step_idx, int L, int64_t N, int mode, int coef_CE) {
    __shared__ double part[8];
    __shared__ double rows[32];
    for (int l = 0; l < L; ++l) {
        double acc = 0.0;
        for (int64_t n = threadIdx.x; n < N; n += blockDim.x) acc += (double)cosv[(int64_t)l * N + n];
        acc = warp_sum(acc);
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int w = 0; w < 8; ++w) s += part[w];
            rows[l] = s;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double cost = 0.0;
        for (int l = 0; l < L; ++l) {
            float s = (float)rows[l];
            if (mode == 0) {
                cost += (double)s;
            } else {
                float weighted = __fmul_rn(coeffs[l], s);   // each_features_loss[l] (TPAMI_attack.py:290)
                cost += (double)weighted;
                if (prev) prev[l] = coef_CE ? weighted : s;  // TPAMI_attack.py:293-297
            }
        }
        if (mode == 1) cost /= (double)L;
        if (cost_log) cost_log[step_idx ? *step_idx : 0] = (float)cost;
    }
}

}  // namespace i2v

using namespace i2v;

// Cluster size / stash policy (host).  I2V_COS_CLUSTER=<1|2|4|8|16> overrides for experiments.
static int pick_cluster(int64_t N, int64_t units) {
    if (const char* e = getenv("I2V_COS_CLUSTER")) {
        int v = atoi(e);
        if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16) return v;
    }
    // 8 CTAs per frame while every CTA keeps >= 16K elements to stream.  Measured with the L2 hints (256 frames of
    // ResNet layer2, timed alone): 289 us at 8, 311 us at 16, 337 us at 4 — 16 keeps more of a frame on-chip but pays a
    // cluster-wide reduction per 200 KB streamed and fits the GPCs worse; the evict_last tails of 37 frames in flight
    // (85 MB) still sit in the 126 MB L2.
    int S = 8;
    while (S > 1 && units / S < 4096) S >>= 1;
    (void)N;
    return S;
}

template <bool VEC, int THREADS>
static int cosine_launch_t(const float* a, const float* b, float* grad_a, float* cos_out, int64_t N, int64_t D,
                         const float* w_dev, float w_host, int relu_mask, cudaStream_t st) {
    static int smem_optin = -1;
    static bool attr_done = false;
    auto kern = cosine_loss_grad_kernel<VEC, THREADS>;
    if (smem_optin < 0) {
        int dev = 0, v = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || v <= 0) v = 48 * 1024;
        smem_optin = v;
    }
    const int64_t units = VEC ? D / 4 : D;
    // Two CTAs share an SM so that one CTA's load phase overlaps the other's reduce/store phase; each
    // gets half of the shared memory for its stash.  I2V_COS_CTAS_PER_SM=1 gives one CTA the whole SM.
    const int per_sm = 1024 / THREADS;
    const int64_t cap_max = VEC ? ((smem_optin - 2048) / per_sm - 1024) / 32 : 0;   // float4 pairs per CTA
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - 2048);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return cuda_fail(e, "i2v_cosine_loss_grad_f32 (attributes)");
        attr_done = true;
    }
    int S = pick_cluster(N, units);

    for (;;) {
        const int64_t slice = (units + S - 1) / S;
        const int64_t cap = !VEC || grad_a == nullptr ? 0 : (slice <= cap_max ? slice : cap_max);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(N * S));   // upper bound; trimmed to the co-resident cluster count below
        cfg.blockDim = dim3(THREADS);
        cfg.dynamicSmemBytes = (size_t)cap * 32;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)S;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int nclusters = 0;
        if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) != cudaSuccess || nclusters < 1) {
            cudaGetLastError();
            if (S > 1) { S >>= 1; continue; }   // e.g. a non-portable size this device cannot co-schedule
            nclusters = 1;
        }
        if (getenv("I2V_COS_NONPERSISTENT") == nullptr && (int64_t)nclusters < N) cfg.gridDim = dim3((unsigned)(nclusters * S));
        static const int hint_mode = getenv("I2V_COS_L2_HINTS") ? atoi(getenv("I2V_COS_L2_HINTS")) : 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a, b, grad_a, cos_out, D, w_dev, w_host, relu_mask, cap, N, hint_mode);
        if (e != cudaSuccess) return cuda_fail(e, "i2v_cosine_loss_grad_f32");
        return I2V_OK;
    }
}

// $I2V_COS_CTAS_PER_SM = 2 (default): 512-thread CTAs, two per SM; 1: 1024-thread CTAs, one per SM
template <bool VEC>
static int cosine_launch(const float* a, const float* b, float* grad_a, float* cos_out, int64_t N, int64_t D,
                         const float* w_dev, float w_host, int relu_mask, cudaStream_t st) {
    static const int per_sm = getenv("I2V_COS_CTAS_PER_SM") ? atoi(getenv("I2V_COS_CTAS_PER_SM")) : 2;
    if (per_sm == 1) return cosine_launch_t<VEC, 1024>(a, b, grad_a, cos_out, N, D, w_dev, w_host, relu_mask, st);
    return cosine_launch_t<VEC, 512>(a, b, grad_a, cos_out, N, D, w_dev, w_host, relu_mask, st);
}

extern "C" int i2v_cosine_loss_grad_f32(const float* a, const float* b, float* grad_a, float* cos_out, int64_t N,
                                        int64_t D, const float* w_dev, float w_host, int relu_mask,
                                        i2v_stream_t stream) {
    if (N == 0) return I2V_OK;   // empty batch (pointers may be null, D is whatever the caller computed)
    I2V_REQUIRE(N > 0 && D >= 1, "bad sizes N=%lld D=%lld", (long long)N, (long long)D);
    I2V_REQUIRE(a && b, "null feature pointer");
    I2V_REQUIRE(N <= 0x7fffffff / 16, "too many frames in one call");
    const bool vec = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
                                       reinterpret_cast<uintptr_t>(grad_a)) & 15) == 0;
    if (vec) return cosine_launch<true>(a, b, grad_a, cos_out, N, D, w_dev, w_host, relu_mask, as_stream(stream));
    return cosine_launch<false>(a, b, grad_a, cos_out, N, D, w_dev, w_host, relu_mask, as_stream(stream));
}

extern "C" int i2v_layer_reweight_f32(float* coeffs, const float* prev, int L, float momentum, float* w_out,
                                      float* weights_log, const int* step_idx, i2v_stream_t stream) {
    I2V_REQUIRE(coeffs && prev, "null pointer");
    I2V_REQUIRE(L >= 1 && L <= 32, "L must be in [1,32] (got %d)", L);
    layer_reweight_kernel<<<1, 32, 0, as_stream(stream)>>>(coeffs, prev, L, momentum, w_out, weights_log, step_idx);
    I2V_LAUNCH_CHECK("i2v_layer_reweight_f32");
    return I2V_OK;
}

extern "C" int i2v_layer_sums_f32(const float* cosv, const float* coeffs, float* prev, float* cost_log,
                                  const int* step_idx, int L, int64_t N, int mode, int coef_CE, i2v_stream_t stream) {
    I2V_REQUIRE(cosv, "null pointer");
    I2V_REQUIRE(L >= 1 && L <= 32 && N >= 0, "bad sizes L=%d N=%lld", L, (long long)N);
    I2V_REQUIRE(mode == 0 || (mode == 1 && coeffs), "mode 1 needs coeffs");
    layer_sums_kernel<<<1, 256, 0, as_stream(stream)>>>(cosv, coeffs, prev, cost_log, step_idx, L, N, mode, coef_CE);
    I2V_LAUNCH_CHECK("i2v_layer_sums_f32");
    return I2V_OK;
}
