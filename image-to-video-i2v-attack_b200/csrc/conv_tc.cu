// K4/K5 (tensor-core path): NHWC convolution forward / stride-1 data-gradient as an implicit GEMM on the
// 5th-generation tensor cores — tcgen05.mma.kind::tf32 with the accumulator in TMEM, operands staged in
// shared memory by TMA (tiled 2-D loads for 1x1/s1, im2col-mode loads for everything else), mbarrier
// producer/consumer pipeline, one output tile of 128 pixels x BN channels per CTA.
//
//     D[m, n] = sum_{tap=(r,s)} sum_{c} A[pixel(m) + tap, c] * B[n, (tap, c)]
//       A : activations [N,H,W,C] f32 (C % 32 == 0); one pipeline stage = one tap x 32 channels
//           = a 128-row x 128-byte tile, SWIZZLE_128B, written by ONE TMA instruction
//       B : weights [Cout, R*S*C] f32, K-major, eval-mode BatchNorm scale folded in (host)
//       D : TMEM, 128 lanes x BN f32 columns
//
// FP32-parity mode (X3 = true): the tensor core reads f32 bit patterns as TF32, i.e. it ignores the low
// 13 mantissa bits.  Each operand is split exactly into hi = trunc_tf32(v) and lo = v - hi, and
//     a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo          (the dropped a_lo*b_lo term is < 2^-22 |a*b|)
// is accumulated in FP32 in TMEM: three MMAs per K-slice.  The tensor core adds into its FP32 accumulator
// with truncation (round toward zero), one truncation per MMA instruction, which shows up as a
// systematic relative bias of ~(#instructions/2) ulp (measured 4e-5 at K = 576 with a single accumulator).
// The two small cross terms therefore go to a SECOND TMEM accumulator (their magnitude is 2^-11 of the
// main one, so its truncation error is negligible) and the two are added in the epilogue with a
// round-to-nearest FADD; the main accumulator then sees K/8 truncations instead of 3K/8.
// B_hi/B_lo are precomputed once per model;
// A_hi is the raw tile (hardware truncation), A_lo is produced on the fly by four "split" warps that
// rewrite each landed A tile into a second shared-memory buffer (element-wise, so the swizzle pattern is
// preserved).  X3 = false is plain TF32 (one MMA per K-slice), the mode cuDNN uses with allow_tf32.
//
// Kernels in this file (conv3x3_halo_kernel, the CTA-pair kernel and the first-layer kernels are described next to their code):
//   conv_tc_kernel          (v1) one tile per CTA, 256 threads: warp 0 TMA producer, warp 1 MMA issuer + TMEM allocator,
//                           warps 4..7 A-split during the main loop, then the epilogue (TMEM -> registers -> 128-bit stores).
//                           Kept as the simple reference implementation ($I2V_TC_PERSISTENT=0).
//   conv_tc_persist_kernel  (v2, the one the engine runs) persistent, 512 threads, one CTA per SM: producer, one or TWO MMA
//                           issuers (A_lo in tensor memory), epilogue-TMA issuer(s), four split warps, eight epilogue warps;
//                           accumulators multi-buffered in TMEM; TMA epilogue through swizzled staging slots with bit-packed
//                           ReLU masks, or a register epilogue for scattered (strided-class) output rows.  Its variants,
//                           what was measured for each and what was tried and dropped are described next to the code.
// Host side: tensor-map cache, shape dispatch (tc_run), the first-layer entry points (im2col + GEMM, GEMM + col2im, and the
// experimental patch-matrix-free forward) and the strided data-gradient classes.
// Every mbarrier wait has a clock-based watchdog that traps instead of hanging the GPU.
#include <cuda.h>
#include <stdlib.h>
#include <mutex>
#include <unordered_map>
#include <type_traits>
#include "common.cuh"

namespace i2v {

constexpr int TC_BM = 128;        // pixels per tile (UMMA M)
constexpr int TC_BK = 32;         // f32 per stage row = 128 bytes = one swizzle span
constexpr int TC_THREADS = 256;
constexpr uint32_t TC_A_BYTES = TC_BM * TC_BK * 4;   // 16 KB

struct TcArgs {
    const float* bias;       // [Cout] or null
    const float* residual;   // [M, Cout] or null
    const float* mask_src;   // [M, Cout] or null
    float* dst;              // [M, Cout]
    const uint32_t* mask_bits;   // [Cout/32][M] bit j of word (w, m) = 1[mask source (m, 32w+j) > 0], or null   (TMA epilogue)
    uint32_t* bits_out;          // [Cout/32][M] activity bits of dst written by the epilogue, or null           (TMA epilogue)
    int out_transposed;          // TMA epilogue: dst is [Cout][M] (column planes) instead of [M][Cout]; no residual / bits
    int store_cols;              // TMA epilogue: 32-column sub-tiles starting at or beyond this column are not stored
    int stem4d;                  // direct first-layer forward: A tiles are 16 x 8 pixel boxes of a padded NHWC4 image (see
    int stem_tq, stem_tp;        // i2v_conv_stem_fwd_direct_f32); tiles per image row / column
    int prefetch_tiles;          // L2 prefetch distance in tiles of this CTA (A of 1x1 convolutions, residual sub-tiles); 0 = off
    int64_t M;               // N*P*Q GEMM rows
    int Cout;
    int P, Q;                // row grid: m = (img, p, q)
    int stride;              // im2col traversal stride of the window corner
    int lower_h, lower_w;    // window corner of (p,q) = (0,0) in source coordinates (= -pad for a forward conv)
    int taps_h, taps_w;      // K = (tap_h, tap_w, channel); TMA im2col offsets = (tap_w, tap_h)
    int cblocks;             // C / 32 (both sources together when a second one is attached)
    int a2_cb0;              // > 0: k-steps cb >= a2_cb0 read their A tile from a SECOND, dense [M, C2] tensor through tmRes (1x1
                             // taps only, no residual): two convolutions that are summed anyway — a bottleneck's downsample
                             // branch and its last 1x1 — run as ONE GEMM over the concatenated K
    int relu;
    // output row of GEMM row (img,p,q): ((img*out_H + p*out_s + out_h0)*out_W + q*out_s + out_w0); identity when out_s == 0
    int out_s, out_h0, out_w0, out_H, out_W;
    int64_t M_out;           // rows of dst (= M unless out_s != 0): plane pitch of mask_bits
    // pipeline trace (debug, i2v_conv_tc_set_trace): CTA 0 stamps clock64() at 8 points of each of its first
    // trace_tiles tiles — [tile][0] producer starts the tile, [1] producer issued its last load, [2] MMA warp owns
    // an accumulator, [3] first operands landed (and split), [4] last MMA issued, [5] epilogue sees the
    // accumulator, [6] epilogue done, [7] split warps done with the tile
    unsigned long long* trace;
    int trace_tiles;
    int b_resident;              // persistent kernel: the n-tile's whole weight block [kiters][b_hi | b_lo] is loaded ONCE per CTA and stays in
                                 // shared memory (every tile of a CTA has the same n-tile when gridDim % num_n_tiles == 0); the ring carries A only
    int dbg;                     // pair kernel timing experiments ($I2V_TC_PAIR_DBG): bit 0 skips MMA1, bit 1 skips MMA2, bit 2 skips the A split, bit 3 skips the weight loads (wrong results)
};
#define TC_TRACE(slot, tile_no)                                                                          \
    do {                                                                                                   \
        if (args.trace && blockIdx.x == 0 && (tile_no) < args.trace_tiles)                                 \
            args.trace[(size_t)(tile_no) * 8 + (slot)] = (unsigned long long)clock64();                    \
    } while (0)

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Non-blocking poll: try_wait may SUSPEND the thread up to a system-dependent time limit when the phase is not complete, which
// a thread that has other work to issue (the epilogue-TMA issuer polling two groups, the producer polling a patch slot between
// weight stages) cannot afford.
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Watchdog: ~4 s at 2 GHz.  A protocol bug traps (CUDA error on the host) instead of hanging the box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 8000000000LL) { printf("i2v conv_tc: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
    }
}
// Role selection.  `if (lane == 0)` leaves the compiler unable to prove that the operands of the uniform-datapath instructions
// (UTCHMMA, UTMALDG, UTCBAR, SYNCS) are warp-uniform, so it wraps EVERY one of them in a waterfall loop (ELECT + up to seven
// R2UR.BROADCAST + BRA.U.ANY): ~200 cycles per tcgen05.mma on the issuing thread where an N = 64 / 128 instruction occupies the
// tensor pipe for 32 / 64.  Behind `elect.sync` the same code is straight-line UTCHMMA (cuobjdump -sass: no BRA.U.ANY).
// The whole warp must reach elect_one() converged; the warp index goes through a shuffle so that it is uniform by construction.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
// advance a ring position: `it % stages` / `it / stages` with a run-time divisor cost two emulated integer divisions (~200
// dependent cycles) per k-step on the single thread that paces the pipeline
__device__ __forceinline__ void ring_next(int& st, uint32_t& ph, int n) { if (++st == n) { st = 0; ph ^= 1u; } }
// tile / num_n_tiles for the n-tile counts that occur (1, 2, 4, 8: shifts; anything else: the emulated division) — every role
// derives its tile coordinates once per tile or sub-tile, the epilogue threads among them
__device__ __forceinline__ int div_ntiles(int tile, int nn) {
    if (nn == 1) return tile;
    if ((nn & (nn - 1)) == 0) return tile >> (31 - __clz(nn));
    return tile / nn;
}
__device__ __forceinline__ int uniform_warp_idx() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c, int w, int h,
                                                   int n, uint16_t off_w, uint16_t off_h) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
                 " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 :: "l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// im2col-mode STORE (UTMASTG.4D.IM2COL): the 128 pixel rows of the shared-memory tile go to the pixels the map's traversal visits
// from (w, h, n) on — every elementStride-th pixel of the bounding box, wrapping over rows and images: the scatter of a strided
// data-gradient class
__device__ __forceinline__ void tma_store_im2col_4d(const CUtensorMap* map, const void* src, int c, int w, int h, int n) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.im2col_no_offs.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 :: "l"(map), "r"(smem_u32(src)), "r"(c), "r"(w), "r"(h), "r"(n) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 :: "l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// L2 prefetch of a 2-D tile (no shared memory, no completion tracking): turns the later TMA load's DRAM latency into an
// L2 hit, which is what a pipeline with only 2-3 shared-memory stages of 48-64 KB needs
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" :: "l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(map) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4,
// LBO = 1 (ignored for swizzled K-major), SBO = 1024 B (8 rows x 128 B), version 1 (Blackwell), layout 2.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A = B = tf32, both K-major, M = 128, N = BN.
__device__ __forceinline__ constexpr uint32_t umma_idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// one TMEM lane (= tile row) per thread, 32 consecutive columns; no wait: callers batch loads per wait
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
}

// registers -> one TMEM lane per thread, 32 consecutive columns
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                 "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                 "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
                    "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
                    "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
                 : "memory");
}
// D[tmem] (+)= A[tmem] x B[smem]: A operand in tensor memory (lane = row m, column = k element; cute SM100_MMA_TF32_TS)
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// 16 TMEM lanes x 16 columns per warp: thread t holds row t/4 (regs 4j+e) and row t/4+8 (regs 4j+2+e) of the lane
// window, columns 8j + 2(t%4) + e (cute SM100_TMEM_LOAD_16dp256b2x) -- a quad owns 32 contiguous bytes of a row, so the
// epilogue's global loads / stores are whole 32-byte sectors.  No wait: callers batch several loads per wait.
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int BN, bool X3>
struct TcSmem {
    static constexpr uint32_t B_BYTES = BN * TC_BK * 4;
    static constexpr uint32_t STAGE_BYTES = TC_A_BYTES * (X3 ? 2 : 1) + B_BYTES * (X3 ? 2 : 1);
};

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
template <int BN, bool X3, bool IM2COL>
__global__ void __launch_bounds__(TC_THREADS)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
               const __grid_constant__ CUtensorMap tmBlo, const TcArgs args, const int stages) {
    using L = TcSmem<BN, X3>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment is required by SWIZZLE_128B (TMA destination and UMMA descriptors)
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* tiles = smem;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)stages * L::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + stages;
    uint64_t* split_bar = empty_bar + stages;
    uint64_t* accum_bar = split_bar + stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
    float* bias_s = reinterpret_cast<float*>(tmem_slot + 2);

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int64_t m0 = (int64_t)blockIdx.x * TC_BM;
    const int n0 = blockIdx.y * BN;
    const int kiters = args.taps_h * args.taps_w * args.cblocks;

    auto stage_a = [&](int s) { return tiles + (size_t)s * L::STAGE_BYTES; };
    auto stage_alo = [&](int s) { return tiles + (size_t)s * L::STAGE_BYTES + TC_A_BYTES; };
    auto stage_bhi = [&](int s) { return tiles + (size_t)s * L::STAGE_BYTES + TC_A_BYTES * (X3 ? 2 : 1); };
    auto stage_blo = [&](int s) { return tiles + (size_t)s * L::STAGE_BYTES + TC_A_BYTES * 2 + L::B_BYTES; };

    // ---- one-time setup ----------------------------------------------------------------------
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA); prefetch_tmap(&tmBhi);
        if (X3) prefetch_tmap(&tmBlo);
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
            mbar_init(&split_bar[s], 128);
        }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    constexpr uint32_t kTmemCols = X3 ? 2 * BN : BN;   // main accumulator [+ cross-term accumulator]
    if (warp == 1) {   // TMEM: 128 lanes x kTmemCols f32 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x >= 128 && threadIdx.x - 128 < BN)
        bias_s[threadIdx.x - 128] = args.bias ? args.bias[n0 + threadIdx.x - 128] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====================================================================
        if (elect_one()) {
            int img = 0, base_h = 0, base_w = 0;
            if (IM2COL) {
                const int64_t pq = (int64_t)args.P * args.Q;
                img = (int)(m0 / pq);
                const int rem = (int)(m0 - (int64_t)img * pq);
                const int p = rem / args.Q, q = rem - p * args.Q;
                base_h = p * args.stride + args.lower_h;   // coordinate of the filter window's corner (tap 0,0)
                base_w = q * args.stride + args.lower_w;
            }
            int it = 0;
            int st = 0; uint32_t ph = 0;               // ring position of k-step `it` (no division in the loop)
            for (int r = 0; r < args.taps_h; ++r)
                for (int s = 0; s < args.taps_w; ++s)
                    for (int cb = 0; cb < args.cblocks; ++cb, ++it, ring_next(st, ph, stages)) {
                        mbar_wait(&empty_bar[st], ph ^ 1);
                        mbar_arrive_expect_tx(&full_bar[st], TC_A_BYTES + L::B_BYTES * (X3 ? 2 : 1));
                        if (IM2COL) tma_load_im2col_4d(&tmA, &full_bar[st], stage_a(st), cb * TC_BK, base_w, base_h, img, (uint16_t)s, (uint16_t)r);
                        else        tma_load_2d(&tmA, &full_bar[st], stage_a(st), cb * TC_BK, (int)m0);
                        const int kcol = ((r * args.taps_w + s) * args.cblocks + cb) * TC_BK;
                        tma_load_2d(&tmBhi, &full_bar[st], stage_bhi(st), kcol, n0);
                        if (X3) tma_load_2d(&tmBlo, &full_bar[st], stage_blo(st), kcol, n0);
                    }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =======================================================================
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_tf32(BN);
            int st = 0; uint32_t ph = 0;
            for (int it = 0; it < kiters; ++it, ring_next(st, ph, stages)) {
                mbar_wait(&full_bar[st], ph);
                if (X3) mbar_wait(&split_bar[st], ph);
                tc_fence_after();
                const uint64_t da = umma_desc_sw128(smem_u32(stage_a(st)));
                const uint64_t dbh = umma_desc_sw128(smem_u32(stage_bhi(st)));
                uint64_t dal = 0, dbl = 0;
                if (X3) { dal = umma_desc_sw128(smem_u32(stage_alo(st))); dbl = umma_desc_sw128(smem_u32(stage_blo(st))); }
#pragma unroll
                for (int kk = 0; kk < TC_BK / 8; ++kk) {              // UMMA K = 8 tf32 = 32 bytes: +2 in the >>4 address field
                    umma_tf32(tmem_base, da + 2 * kk, dbh + 2 * kk, idesc, (it | kk) ? 1u : 0u);
                    if (X3) {   // cross terms into the second accumulator (columns BN..2BN)
                        umma_tf32(tmem_base + BN, dal + 2 * kk, dbh + 2 * kk, idesc, (it | kk) ? 1u : 0u);
                        umma_tf32(tmem_base + BN, da + 2 * kk, dbl + 2 * kk, idesc, 1u);
                    }
                }
                umma_commit(&empty_bar[st]);       // frees the stage when these MMAs have read it
            }
            umma_commit(accum_bar);                // accumulator complete
        }
    } else if (warp >= 4) {
        const int t = threadIdx.x - 128;           // 0..127
        if (X3) {
            // ===== A split: A_lo = a - trunc_tf32(a), element-wise on the swizzled tile ==============
            int st = 0; uint32_t ph = 0;
            for (int it = 0; it < kiters; ++it, ring_next(st, ph, stages)) {
                mbar_wait(&full_bar[st], ph);
                const float4* src = reinterpret_cast<const float4*>(stage_a(st));
                float4* dst = reinterpret_cast<float4*>(stage_alo(st));
#pragma unroll
                for (int j = 0; j < (int)(TC_A_BYTES / 16 / 128); ++j) {
                    float4 v = src[t + 128 * j], o;
                    o.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
                    o.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
                    o.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
                    o.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
                    dst[t + 128 * j] = o;
                }
                fence_proxy_async();               // generic-proxy stores -> visible to the tensor core (async proxy)
                mbar_arrive(&split_bar[st]);
            }
        }
        // ===== epilogue ===========================================================================
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const int64_t m = m0 + t;                  // thread t owns TMEM lane t = GEMM row m
        int64_t orow = m;
        if (args.out_s != 0 && m < args.M) {       // strided data-gradient class: scatter rows into the full-resolution tensor
            const int64_t pq = (int64_t)args.P * args.Q;
            const int img = (int)(m / pq);
            const int rem = (int)(m - (int64_t)img * pq);
            const int p = rem / args.Q, q = rem - p * args.Q;
            orow = ((int64_t)img * args.out_H + (p * args.out_s + args.out_h0)) * args.out_W + (q * args.out_s + args.out_w0);
        }
        const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(lane_base + (uint32_t)c0, r);   // warp-collective: every lane participates
            if (X3) {
                uint32_t r2[32];
                tmem_ld32(lane_base + (uint32_t)(BN + c0), r2);
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__fadd_rn(__uint_as_float(r[j]), __uint_as_float(r2[j])));
            }
            if (m < args.M) {
                const int64_t off = orow * args.Cout + n0 + c0;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 v = make_float4(__uint_as_float(r[j]) + bias_s[c0 + j], __uint_as_float(r[j + 1]) + bias_s[c0 + j + 1],
                                           __uint_as_float(r[j + 2]) + bias_s[c0 + j + 2], __uint_as_float(r[j + 3]) + bias_s[c0 + j + 3]);
                    if (args.residual) {
                        const float4 q = __ldg(reinterpret_cast<const float4*>(args.residual + off + j));
                        v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
                    }
                    if (args.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                    if (args.mask_src) {
                        const float4 q = __ldg(reinterpret_cast<const float4*>(args.mask_src + off + j));
                        if (!(q.x > 0.f)) v.x = 0.f; if (!(q.y > 0.f)) v.y = 0.f; if (!(q.z > 0.f)) v.z = 0.f; if (!(q.w > 0.f)) v.w = 0.f;
                    }
                    *reinterpret_cast<float4*>(args.dst + off + j) = v;
                }
            }
        }
    }

    // ---- teardown --------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// v2: persistent, fully warp-specialised variant.
// One CTA per SM walks over (m-tile, n-tile) pairs; the TMA producer and the split warps run ahead across
// tile boundaries, the accumulator is multi-buffered in TMEM so tile i+1's MMAs overlap tile i's epilogue,
// which has its own four warps.  Removes the per-tile launch / TMEM-allocation / pipeline-fill bubbles that
// dominate v1 on the many small-K 1x1 convolutions (measured v1: 2.2 us per 128x64 tile at K = 64).
//   warp 0: TMA producer   warp 1: MMA issuer   warp 2: TMEM allocator   warps 4-7: A split (3xTF32)
//   warps 8-15: epilogue.  Warp w owns TMEM lanes 32*(w%4)..+31 and the 16-column units u = (w-8)/4, +2, ...
//   of the tile.  TMEM is read with the 16x256b shape (a quad of threads = 8 consecutive columns of one row), so
//   residual / mask loads and the output stores are float2 accesses that fill whole 32-byte sectors; the loads of
//   unit u+1 are issued before unit u is processed (and the first unit's before the accumulator is even
//   complete), which is what keeps the HBM-bound 1x1 convolutions with a residual from serialising on latency
//   (measured with the one-row-per-thread epilogue: 141 us for layer1.conv3 on 32 frames, 6.8 % tensor-pipe active).
// 3xTF32 issue order per 8-wide k slice: one N = 2*BN instruction a_hi x [b_hi | b_lo] (B_hi and B_lo are adjacent
// in the stage, main and cross accumulators are adjacent in TMEM), then a_lo x b_hi into the cross accumulator.
// ---------------------------------------------------------------------------------------------
constexpr int TC2_THREADS = 512;
constexpr int TC2_EPI_THREADS = 256;

// Epilogue v3 (EPI_TMA): no per-thread global access at all.  The two epilogue groups (warps 8-11, 12-15: one warp per
// TMEM lane quarter) work on 128-row x 32-column sub-tiles: tcgen05.ld 32x32b (thread = tile row), + bias, + residual read
// from a SWIZZLE_128B staging slot that a TMA load filled, ReLU, ReLU-backward mask from one bit-mask word per thread,
// result written back to the same slot (conflict-free 128-bit accesses thanks to the swizzle) and shipped by ONE TMA
// store; warps 2 / 3 (one elected lane each) issue a group's stores and residual loads, decoupled by mbarriers
// (out_ready: slot written; slot_ready: slot free again / next residual landed).  Measured reason: with the v2 epilogue a
// warp-level float2 access touches 8 cache lines, and the LSU wavefronts of output + residual + mask (3 x 2048 cycles per
// 128x128 tile) made every 1x1 convolution epilogue-bound (2.5-4.5 us per tile against 1-1.5 us of main loop).
constexpr uint32_t EPI_SLOT_BYTES = TC_BM * 32 * 4;            // 16 KB: 128 rows x 128 bytes
constexpr uint32_t EPI_STAGING_BYTES = 4 * EPI_SLOT_BYTES;     // two slots per epilogue group

template <int BN, bool X3, bool IM2COL, bool EPI_TMA, bool ALO_TMEM>
__global__ void __launch_bounds__(TC2_THREADS, 1)
conv_tc_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
                       const __grid_constant__ CUtensorMap tmBlo, const __grid_constant__ CUtensorMap tmOut,
                       const __grid_constant__ CUtensorMap tmRes, const TcArgs args, const int stages,
                       const int num_m_tiles, const int num_n_tiles, const int epi_slots) {
    using L = TcSmem<BN, X3>;
    static_assert(!ALO_TMEM || X3, "A_lo in tensor memory is a 3xTF32 variant");
    // ALO_TMEM = the DUAL-ISSUER variant.  Measured (tools/mma_probe.py, profiles/r01_mma_issue_probe.txt): ONE thread
    // issues a tcgen05.mma every ~200 cycles whatever its width (N = 64..256, floor N/2 cycles), while several issuing
    // warps proceed concurrently at that same rate each — the 3xTF32 main loop (8 MMAs per k-step) was issue-bound.
    // (End of round 2: those 200 cycles were the compiler's waterfall loop around every UTCHMMA of an `if (lane == 0)` role and
    // the ring's integer divisions — see elect_one() / ring_next() above.  With both gone two issuers still win, because each
    // keeps its own accumulators and the two instruction streams overlap in the tensor pipe; the A/B is in
    // profiles/r02_dispatch_ab_per_shape.txt.)
    // So the two MMAs of a k-slice go to two issuers with DISJOINT accumulators: warp 1 issues a_hi x [b_hi | b_lo]
    // (main | cross), warp 2 issues a_lo x b_hi with A_lo read from tensor memory into a third accumulator (cross2);
    // the epilogue adds the three.  TMEM: kAcc stages of [main | cross | cross2] in 384 columns + a 4-slot A_lo ring.
    // (Tried and measured NOT to help: a THIRD issuing warp for the K-heavy BN = 64 layers — k-slices 0..2 on warps 1 / 2,
    // both instructions of k-slice 3 on a 17th warp into its own accumulators, three instructions per issuer and k-step
    // instead of four.  3x3 64->64 at 256 frames: 474 us against 454 us with two issuers in the same run.  So once two
    // threads issue, these layers are no longer bound by the issue rate; what remains is operand delivery — every SM
    // pulls the same 16 KB weight block and a 16 KB im2col tile per k-step through L2, 8 TB/s in aggregate — which is
    // what TMA multicast across a CTA pair / cta_group::2 would halve.  Two more layouts of the same kernel were measured on
    // that layer and dropped: ONE accumulator stage + an 8-slot A_lo ring so that 6 shared-memory stages are in flight
    // instead of 4 (447.8 against 447.4 us: not bound by pipeline depth either), and BOTH A operands in tensor memory —
    // the split warps also copy the raw tile, both MMAs read A from TMEM, 16 KB less shared-memory traffic per k-step
    // (453.7 against 429.9 us: not shared-memory bandwidth).  Plain TF32 without any split runs the layer in 321 us.)
    constexpr uint32_t kAccCols = X3 ? (ALO_TMEM ? 3 * BN : 2 * BN) : BN;   // TMEM columns per accumulator stage
    constexpr int kAcc = ALO_TMEM ? (int)(384 / kAccCols) : ((512 / kAccCols) > 4 ? 4 : (512 / kAccCols));
    constexpr uint32_t kAloBase = 384, kAloSlots = 4;
    constexpr uint32_t kTmemCols = ALO_TMEM ? 512 : kAccCols * kAcc;   // 256 or 512: a power of two
    constexpr uint32_t kStageBytes = ALO_TMEM ? L::STAGE_BYTES - TC_A_BYTES : L::STAGE_BYTES;   // no A_lo tile in smem
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // resident weights (args.b_resident): the ring stages shrink to the A tile(s), the n-tile's weights follow the ring
    const bool bres_on = args.b_resident != 0;
    const uint32_t stage_bytes = bres_on ? kStageBytes - (X3 ? 2u : 1u) * L::B_BYTES : kStageBytes;
    const uint32_t bres_kb = (X3 ? 2u : 1u) * L::B_BYTES;       // per k-step: [b_hi | b_lo]
    const uint32_t bres_bytes = bres_on ? (uint32_t)(args.taps_h * args.taps_w * args.cblocks) * bres_kb : 0u;
    uint8_t* tiles = smem;
    uint8_t* bres = smem + (size_t)stages * stage_bytes;
    uint8_t* staging = bres + bres_bytes;                       // stage sizes are multiples of 1024: stays swizzle-aligned
    // staging: epi_slots (1 or 2) slots of 16 KB per epilogue group.  K-heavy layers without a residual take ONE slot
    // per group so that a third 64 KB pipeline stage fits (3xTF32, BN = 128): with two stages the tensor pipe idled
    // ~45 % of the time waiting for the next k-step's operands (measured 53 % active on the 3x3 128->128 layers)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + (EPI_TMA ? (size_t)epi_slots * 2 * EPI_SLOT_BYTES : 0));
    uint64_t* empty_bar = full_bar + stages;
    uint64_t* split_bar = empty_bar + stages;
    uint64_t* tfull_bar = split_bar + stages;
    uint64_t* tempty_bar = tfull_bar + kAcc;
    uint64_t* out_ready = tempty_bar + kAcc;                    // [group][slot]
    uint64_t* slot_ready = out_ready + 4;                       // [group][slot]
    uint64_t* bres_full = slot_ready + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres_full + 1);
    float* bias_s = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 2) + 15) & ~(uintptr_t)15);   // [Cout], 16-byte aligned

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int kiters = args.taps_h * args.taps_w * args.cblocks;
    const int num_tiles = num_m_tiles * num_n_tiles;

    auto stage_a = [&](int s) { return tiles + (size_t)s * stage_bytes; };
    auto stage_alo = [&](int s) { return tiles + (size_t)s * stage_bytes + TC_A_BYTES; };
    auto stage_bhi = [&](int s) { return tiles + (size_t)s * stage_bytes + TC_A_BYTES * ((X3 && !ALO_TMEM) ? 2 : 1); };
    auto stage_blo = [&](int s) { return stage_bhi(s) + L::B_BYTES; };
    // weights of k-step kb: the ring stage, or the resident block
    auto b_of = [&](int s, int kb) { return bres_on ? bres + (size_t)kb * bres_kb : stage_bhi(s); };

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA); prefetch_tmap(&tmBhi);
        if (X3) prefetch_tmap(&tmBlo);
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], ALO_TMEM ? 2 : 1);        // dual issuer: both MMA warps release a stage
            mbar_init(&split_bar[s], 128);
        }
        for (int a = 0; a < kAcc; ++a) {
            mbar_init(&tfull_bar[a], ALO_TMEM ? 2 : 1);        // ... and both complete an accumulator stage
            mbar_init(&tempty_bar[a], TC2_EPI_THREADS);
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(&out_ready[i], TC2_EPI_THREADS / 2);
            mbar_init(&slot_ready[i], 1);
        }
        mbar_init(bres_full, 1);
        if (EPI_TMA) { prefetch_tmap(&tmOut); prefetch_tmap(&tmRes); }
        fence_barrier_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (!EPI_TMA)      // the TMA epilogue reads the bias through the read-only path (warp-uniform addresses): no smem
        for (int i = threadIdx.x; i < args.Cout; i += TC2_THREADS) bias_s[i] = args.bias ? args.bias[i] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer: runs ahead across tiles, bounded only by the smem ring ===================
        if (elect_one()) {
            int it = 0, tno = 0;
            int st = 0; uint32_t ph = 0;               // ring position of k-step `it` (no division in the loop)
            if (bres_on && (int)blockIdx.x < num_tiles) {
                const int n0r = ((int)blockIdx.x % num_n_tiles) * BN;      // the n-tile of every tile of this CTA
                mbar_arrive_expect_tx(bres_full, bres_bytes);
                for (int kb = 0; kb < kiters; ++kb) {
                    tma_load_2d(&tmBhi, bres_full, bres + (size_t)kb * bres_kb, kb * TC_BK, n0r);
                    if (X3) tma_load_2d(&tmBlo, bres_full, bres + (size_t)kb * bres_kb + L::B_BYTES, kb * TC_BK, n0r);
                }
            }
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tno) {
                const int m_tile = div_ntiles(tile, num_n_tiles), n_tile = tile - m_tile * num_n_tiles;
                const int64_t m0 = (int64_t)m_tile * TC_BM;
                const int n0 = n_tile * BN;
                int img = 0, base_h = 0, base_w = 0;
                if (IM2COL) {
                    const int64_t pq = (int64_t)args.P * args.Q;
                    img = (int)(m0 / pq);
                    const int rem = (int)(m0 - (int64_t)img * pq);
                    const int p = rem / args.Q, q = rem - p * args.Q;
                    base_h = p * args.stride + args.lower_h;
                    base_w = q * args.stride + args.lower_w;
                } else if (args.prefetch_tiles > 0 && args.a2_cb0 == 0) {
                    // the first tiles' prefetches are issued up front, afterwards one tile per tile
                    for (int d = (tno == 0 ? 1 : args.prefetch_tiles); d <= args.prefetch_tiles; ++d) {
                        const int ptile = tile + d * (int)gridDim.x;
                        if (ptile >= num_tiles) break;
                        const int pm = div_ntiles(ptile, num_n_tiles);
                        for (int cb = 0; cb < args.cblocks; ++cb) tma_prefetch_l2_2d(&tmA, cb * TC_BK, pm * TC_BM);
                    }
                }
                for (int r = 0; r < args.taps_h; ++r)
                    for (int s = 0; s < args.taps_w; ++s)
                        for (int cb = 0; cb < args.cblocks; ++cb, ++it, ring_next(st, ph, stages)) {
                            mbar_wait(&empty_bar[st], ph ^ 1);
                            if ((r | s | cb) == 0) TC_TRACE(0, tno);
                            // (measured: deriving B_lo on the fly in the split warps instead of loading it — 25-33 % fewer TMA
                            // rows per k-step — made every layer 5-15 % SLOWER: the four split warps are the tighter resource)
                            const bool no_b = (args.dbg & 8) != 0 || bres_on;   // resident weights (or the timing experiment): no weight loads
                            mbar_arrive_expect_tx(&full_bar[st], TC_A_BYTES + (no_b ? 0u : L::B_BYTES * (X3 ? 2 : 1)));
                            if (args.a2_cb0 > 0 && cb >= args.a2_cb0) tma_load_2d(&tmRes, &full_bar[st], stage_a(st), (cb - args.a2_cb0) * TC_BK, (int)m0);
                            else if (IM2COL) tma_load_im2col_4d(&tmA, &full_bar[st], stage_a(st), cb * TC_BK, base_w, base_h, img, (uint16_t)s, (uint16_t)r);
                            else if (args.stem4d) {
                                // filter row r of a 16 x 8 pixel box: 32 floats (8 padded pixels x 4) per output pixel, rows of
                                // parity r % 2 (tmA even / tmRes odd: the residual map is free, a first layer has none)
                                const int per = args.stem_tq * args.stem_tp;
                                const int im = m_tile / per, rem = m_tile - im * per;
                                const int pt = rem / args.stem_tq, qt = rem - pt * args.stem_tq;
                                tma_load_4d((r & 1) ? &tmRes : &tmA, &full_bar[st], stage_a(st), 0, qt * 16, pt * 8 + (r >> 1), im);
                            }
                            else        tma_load_2d(&tmA, &full_bar[st], stage_a(st), cb * TC_BK, (int)m0);
                            const int kcol = ((r * args.taps_w + s) * args.cblocks + cb) * TC_BK;
                            if (!no_b) {
                                tma_load_2d(&tmBhi, &full_bar[st], stage_bhi(st), kcol, n0);
                                if (X3) tma_load_2d(&tmBlo, &full_bar[st], stage_blo(st), kcol, n0);
                            }
                        }
                TC_TRACE(1, tno);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer ==============================================================================
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_tf32(BN);
            constexpr uint32_t idesc2 = umma_idesc_tf32(2 * BN);
            int it = 0, t = 0;
            int st = 0; uint32_t ph = 0;               // ring position of k-step `it` (no division in the loop)
            if (bres_on && (int)blockIdx.x < num_tiles) { mbar_wait(bres_full, 0); tc_fence_after(); }
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
                const int acc = t % kAcc;
                const uint32_t aph = (uint32_t)(t / kAcc) & 1;
                mbar_wait(&tempty_bar[acc], aph ^ 1);          // epilogue has drained this accumulator stage
                tc_fence_after();
                TC_TRACE(2, t);
                const uint32_t d0 = tmem_base + (uint32_t)acc * kAccCols;
                for (int kb = 0; kb < kiters; ++kb, ++it, ring_next(st, ph, stages)) {
                    mbar_wait(&full_bar[st], ph);
                    if (X3 && !ALO_TMEM) mbar_wait(&split_bar[st], ph);
                    tc_fence_after();
                    if (kb == 0) TC_TRACE(3, t);
                    const uint64_t da = umma_desc_sw128(smem_u32(stage_a(st)));
                    const uint64_t dbh = umma_desc_sw128(smem_u32(b_of(st, kb)));
                    uint64_t dal = 0;
                    if (X3 && !ALO_TMEM) dal = umma_desc_sw128(smem_u32(stage_alo(st)));
#pragma unroll
                    for (int kk = 0; kk < TC_BK / 8; ++kk) {
                        if (args.dbg & 1) continue;                                  // timing experiment: no MMA
                        if (X3) {
                            // [main | cross] += a_hi x [b_hi | b_lo]  (N = 2*BN), then cross += a_lo x b_hi
                            umma_tf32(d0, da + 2 * kk, dbh + 2 * kk, idesc2, (kb | kk) ? 1u : 0u);
                            if (!ALO_TMEM) umma_tf32(d0 + BN, dal + 2 * kk, dbh + 2 * kk, idesc, 1u);
                        } else {
                            umma_tf32(d0, da + 2 * kk, dbh + 2 * kk, idesc, (kb | kk) ? 1u : 0u);
                        }
                    }
                    umma_commit(&empty_bar[st]);
                }
                umma_commit(&tfull_bar[acc]);
                TC_TRACE(4, t);
            }
        }
    } else if (ALO_TMEM && warp == 2) {
        // ===== second MMA issuer: cross2 += a_lo (tensor memory) x b_hi ===================================
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_tf32(BN);
            int it = 0, t = 0;
            int st = 0; uint32_t ph = 0;               // ring position of k-step `it` (no division in the loop)
            if (bres_on && (int)blockIdx.x < num_tiles) { mbar_wait(bres_full, 0); tc_fence_after(); }
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
                const int acc = t % kAcc;
                const uint32_t aph = (uint32_t)(t / kAcc) & 1;
                mbar_wait(&tempty_bar[acc], aph ^ 1);
                tc_fence_after();
                const uint32_t d2 = tmem_base + (uint32_t)acc * kAccCols + 2u * BN;
                for (int kb = 0; kb < kiters; ++kb, ++it, ring_next(st, ph, stages)) {
                    mbar_wait(&full_bar[st], ph);               // b_hi landed
                    mbar_wait(&split_bar[st], ph);              // a_lo of this k-step is in its ring slot
                    tc_fence_after();
                    const uint64_t dbh = umma_desc_sw128(smem_u32(b_of(st, kb)));
                    const uint32_t talo = tmem_base + kAloBase + (uint32_t)(it % kAloSlots) * 32u;
#pragma unroll
                    for (int kk = 0; kk < TC_BK / 8; ++kk)
                        if (!(args.dbg & 2)) umma_tf32_ts(d2, talo + 8u * kk, dbh + 2 * kk, idesc, (kb | kk) ? 1u : 0u);
                    umma_commit(&empty_bar[st]);
                }
                umma_commit(&tfull_bar[acc]);
            }
        }
    } else if (ALO_TMEM && EPI_TMA && warp == 3) {
        // ===== epilogue TMA issuer of BOTH groups (warp 2 issues MMAs here): polls the groups' out_ready barriers ====
        if (elect_one()) {
            constexpr uint32_t SUBS = BN / 64;
            const bool has_res = args.residual != nullptr;
            const uint32_t ns = (uint32_t)epi_slots;
            const uint32_t my_tiles = (uint32_t)((num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x);
            const uint32_t total = my_tiles * SUBS;
            auto coords = [&](int g, uint32_t k, int& col, int& row) {
                const int tile = (int)blockIdx.x + (int)(k / SUBS) * (int)gridDim.x;
                const int m_tile = div_ntiles(tile, num_n_tiles), n_tile = tile - m_tile * num_n_tiles;
                col = n_tile * BN + (g + 2 * (int)(k % SUBS)) * 32;
                row = m_tile * TC_BM;
            };
            // strided data-gradient class: tile row m = (img, i, j) of the class grid lives at pixel (i out_s, j out_s) of the
            // class VIEW of dst (base shifted to the class's first pixel); im2col-mode TMA walks the view with that stride
            auto view_coords = [&](int row, int& w, int& h, int& n) {
                const int pq = args.P * args.Q;
                n = row / pq;
                const int rem = row - n * pq, i = rem / args.Q;
                h = i * args.out_s; w = (rem - i * args.Q) * args.out_s;
            };
            auto refill = [&](int g, uint32_t k) {              // slot k % ns of group g is free: next residual or a plain arrive
                const uint32_t s = k & (ns - 1);
                if (has_res) {
                    int col, row;
                    coords(g, k, col, row);
                    mbar_arrive_expect_tx(&slot_ready[g * 2 + s], EPI_SLOT_BYTES);
                    if (args.out_s != 0) {
                        int w, h, n;
                        view_coords(row, w, h, n);
                        tma_load_im2col_4d(&tmRes, &slot_ready[g * 2 + s], staging + ((size_t)g * ns + s) * EPI_SLOT_BYTES, col, w, h, n, 0, 0);
                    } else
                    tma_load_2d(&tmRes, &slot_ready[g * 2 + s], staging + ((size_t)g * ns + s) * EPI_SLOT_BYTES, col, row);
                } else {
                    mbar_arrive(&slot_ready[g * 2 + s]);
                }
            };
            for (int g = 0; g < 2; ++g)
                for (uint32_t k = 0; k < ns && k < total; ++k) refill(g, k);
            uint32_t kdone[2] = {0, 0};
            const long long t0 = clock64();
            while (kdone[0] < total || kdone[1] < total) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const uint32_t k = kdone[g];
                    if (k >= total) continue;
                    const uint32_t s = k & (ns - 1), ph = (k >> (ns - 1)) & 1;      // ns is 1 or 2: no division
                    if (!mbar_test_wait(&out_ready[g * 2 + s], ph)) continue;
                    int col, row;
                    coords(g, k, col, row);
                    const uint8_t* slot = staging + ((size_t)g * ns + s) * EPI_SLOT_BYTES;
                    if (col < args.store_cols) {
                        if (args.stem4d) {
                            const int m_tile = row / TC_BM, per = args.stem_tq * args.stem_tp;
                            const int im = m_tile / per, rem = m_tile - im * per;
                            const int pt = rem / args.stem_tq, qt = rem - pt * args.stem_tq;
                            tma_store_4d(&tmOut, slot, col, qt * 16, pt * 8, im);
                        }
                        else if (args.out_s != 0) {
                            int w, h, n;
                            view_coords(row, w, h, n);
                            tma_store_im2col_4d(&tmOut, slot, col, w, h, n);
                        }
                        else if (args.out_transposed) tma_store_2d(&tmOut, slot, row, col);
                        else                          tma_store_2d(&tmOut, slot, col, row);
                    }
                    bulk_commit();
                    if (k + ns < total) {
                        bulk_wait_read0();
                        refill(g, k + ns);
                    }
                    kdone[g] = k + 1;
                }
                if (clock64() - t0 > 40000000000LL) { printf("i2v conv_tc: epilogue issuer timeout (block %d)\n", blockIdx.x); __trap(); }
            }
            bulk_wait_all();
        }
    } else if (warp >= 4 && warp < 8) {
        // ===== A split (3xTF32 only) ====================================================================
        if (X3) {
            const int t128 = threadIdx.x - 128;
            int it = 0, tno = 0;
            int st = 0; uint32_t ph = 0;               // ring position of k-step `it` (no division in the loop)
            if (ALO_TMEM) {
                // A_lo goes to TENSOR MEMORY instead of shared memory: thread = tile row (TMEM lane), 32 k-values of its
                // row read from the swizzled A tile, lo parts written with tcgen05.st into a ring of 32-column slots that
                // the second MMA of the k-step reads as its A operand.  Saves, per k-step, the 16 KB shared-memory write
                // of A_lo and the 16 KB the tensor core would read back — shared-memory bandwidth (128 B/clk) is what
                // bounds the BN = 64 layers: 120 KB per k-step against 640 cycles of MMA.
                const int row = t128;                              // warp w -> TMEM lanes 32*(w%4)..+31 (warps 4-7)
                const uint32_t swz = (uint32_t)(row & 7);
                const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kAloBase;
                for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tno) {
                    for (int kb = 0; kb < kiters; ++kb, ++it, ring_next(st, ph, stages)) {
                        mbar_wait(&full_bar[st], ph);
                        if (args.dbg & 4) { mbar_arrive(&split_bar[st]); continue; }     // timing experiment: no split work
                        const uint8_t* arow = stage_a(st) + (size_t)row * 128;
                        uint32_t lo[32];
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float4 v = *reinterpret_cast<const float4*>(arow + (((uint32_t)c ^ swz) << 4));
                            lo[4 * c + 0] = __float_as_uint(v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u));
                            lo[4 * c + 1] = __float_as_uint(v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u));
                            lo[4 * c + 2] = __float_as_uint(v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u));
                            lo[4 * c + 3] = __float_as_uint(v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u));
                        }
                        tmem_st32(lane_addr + (uint32_t)(it % kAloSlots) * 32u, lo);
                        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                        tc_fence_before();
                        mbar_arrive(&split_bar[st]);
                    }
                    if (t128 == 0) TC_TRACE(7, tno);
                }
            } else {
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tno) {
                for (int kb = 0; kb < kiters; ++kb, ++it, ring_next(st, ph, stages)) {
                    mbar_wait(&full_bar[st], ph);
                    const float4* src = reinterpret_cast<const float4*>(stage_a(st));
                    float4* dst = reinterpret_cast<float4*>(stage_alo(st));
#pragma unroll
                    for (int j = 0; j < (int)(TC_A_BYTES / 16 / 128); ++j) {
                        float4 v = src[t128 + 128 * j], o;
                        o.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
                        o.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
                        o.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
                        o.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
                        dst[t128 + 128 * j] = o;
                    }
                    fence_proxy_async();
                    mbar_arrive(&split_bar[st]);
                }
                if (t128 == 0) TC_TRACE(7, tno);
            }
            }
        }
    } else if (EPI_TMA && !ALO_TMEM && (warp == 2 || warp == 3)) {
        // ===== epilogue TMA issuer of group g = warp - 2: output stores and residual loads ================
        if (elect_one()) {
            const int g = warp - 2;
            constexpr uint32_t SUBS = BN / 64;                  // sub-tiles per tile and group
            const bool has_res = args.residual != nullptr;
            const uint32_t ns = (uint32_t)epi_slots;
            uint8_t* sbase = staging + (size_t)g * ns * EPI_SLOT_BYTES;
            const uint32_t my_tiles = (uint32_t)((num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x);
            const uint32_t total = my_tiles * SUBS;
            auto coords = [&](uint32_t k, int& col, int& row) {
                const int tile = (int)blockIdx.x + (int)(k / SUBS) * (int)gridDim.x;
                const int m_tile = div_ntiles(tile, num_n_tiles), n_tile = tile - m_tile * num_n_tiles;
                col = n_tile * BN + (g + 2 * (int)(k % SUBS)) * 32;
                row = m_tile * TC_BM;
            };
            int col, row;
            auto view_coords = [&](int row_, int& w, int& h, int& n) {       // see the dual-issuer variant's issuer warp
                const int pq = args.P * args.Q;
                n = row_ / pq;
                const int rem = row_ - n * pq, i = rem / args.Q;
                h = i * args.out_s; w = (rem - i * args.Q) * args.out_s;
            };
            auto load_res = [&](uint32_t slot_idx, uint8_t* dst_) {
                mbar_arrive_expect_tx(&slot_ready[g * 2 + slot_idx], EPI_SLOT_BYTES);
                if (args.out_s != 0) {
                    int w, h, n;
                    view_coords(row, w, h, n);
                    tma_load_im2col_4d(&tmRes, &slot_ready[g * 2 + slot_idx], dst_, col, w, h, n, 0, 0);
                } else {
                    tma_load_2d(&tmRes, &slot_ready[g * 2 + slot_idx], dst_, col, row);
                }
            };
            const uint32_t pf = (has_res && args.out_s == 0) ? (uint32_t)args.prefetch_tiles * SUBS : 0u;   // L2 prefetch distance in sub-tiles
            for (uint32_t k = ns; k < ns + pf && k < total; ++k) { coords(k, col, row); tma_prefetch_l2_2d(&tmRes, col, row); }
            for (uint32_t k = 0; k < ns && k < total; ++k) {    // prime the slots
                if (has_res) {
                    coords(k, col, row);
                    load_res(k, sbase + (size_t)k * EPI_SLOT_BYTES);
                } else {
                    mbar_arrive(&slot_ready[g * 2 + k]);
                }
            }
            for (uint32_t k = 0; k < total; ++k) {
                const uint32_t s = k & (ns - 1), ph = (k >> (ns - 1)) & 1;      // ns is 1 or 2: no division
                mbar_wait(&out_ready[g * 2 + s], ph);           // the group's 128 threads wrote the slot (and fenced)
                coords(k, col, row);
                if (col < args.store_cols) {
                    if (args.out_s != 0) {
                        int w, h, n;
                        view_coords(row, w, h, n);
                        tma_store_im2col_4d(&tmOut, sbase + (size_t)s * EPI_SLOT_BYTES, col, w, h, n);
                    }
                    else if (args.out_transposed) tma_store_2d(&tmOut, sbase + (size_t)s * EPI_SLOT_BYTES, row, col);
                    else                          tma_store_2d(&tmOut, sbase + (size_t)s * EPI_SLOT_BYTES, col, row);
                }
                bulk_commit();
                if (k + ns < total) {
                    bulk_wait_read0();                           // the store has read the slot: it may be refilled
                    if (has_res) {
                        if (pf && k + ns + pf < total) { coords(k + ns + pf, col, row); tma_prefetch_l2_2d(&tmRes, col, row); }
                        coords(k + ns, col, row);
                        load_res(s, sbase + (size_t)s * EPI_SLOT_BYTES);
                    } else {
                        mbar_arrive(&slot_ready[g * 2 + s]);
                    }
                }
            }
            bulk_wait_all();
        }
    } else if (EPI_TMA && warp >= 8) {
        // ===== epilogue v3: thread = tile row, 32-column sub-tiles through swizzled staging slots =========
        constexpr int SUBS = BN / 64;
        const int g = (warp - 8) >> 2;
        const int row = (warp & 3) * 32 + lane;                 // tile row = TMEM lane
        const uint32_t swz = (uint32_t)(row & 7);
        const uint32_t ns = (uint32_t)epi_slots;
        uint8_t* sbase = staging + (size_t)g * ns * EPI_SLOT_BYTES + (size_t)row * 128;
        const bool has_res = args.residual != nullptr;
        const float* __restrict__ gbias = args.bias;
        const uint32_t* __restrict__ mbits = args.mask_bits;
        uint32_t* __restrict__ obits = args.bits_out;
        uint32_t k = 0;
        int t = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
            const int m_tile = div_ntiles(tile, num_n_tiles), n_tile = tile - m_tile * num_n_tiles;
            const int n0 = n_tile * BN;
            const int acc = t % kAcc;
            const uint32_t aph = (uint32_t)(t / kAcc) & 1;
            const int64_t m = (int64_t)m_tile * TC_BM + row;
            const bool valid = m < args.M;
            int64_t mo = m;                                      // row of dst: the scattered pixel of a strided class
            if (args.out_s != 0 && valid && mbits) {
                const int64_t pq = (int64_t)args.P * args.Q;
                const int img = (int)(m / pq);
                const int rem = (int)(m - (int64_t)img * pq);
                const int p = rem / args.Q, q = rem - p * args.Q;
                mo = ((int64_t)img * args.out_H + (p * args.out_s + args.out_h0)) * args.out_W + (q * args.out_s + args.out_w0);
            }
            uint32_t mw[SUBS];
#pragma unroll
            for (int j = 0; j < SUBS; ++j)
                mw[j] = (mbits && valid) ? __ldg(mbits + (int64_t)(n0 / 32 + g + 2 * j) * args.M_out + mo) : 0xFFFFFFFFu;
            mbar_wait(&tfull_bar[acc], aph);
            tc_fence_after();
            if (threadIdx.x == 256) TC_TRACE(5, t);
            const uint32_t tacc = tmem_base + (uint32_t)acc * kAccCols + ((uint32_t)((warp & 3) * 32) << 16);
            // drain this thread's part of the accumulator stage into registers FIRST and hand the stage back (the
            // dual-issuer variant has a single stage at BN = 128: the MMA warps wait for exactly this)
            uint32_t vals[SUBS][32];
#pragma unroll
            for (int j = 0; j < SUBS; ++j) {
                const int c0 = (g + 2 * j) * 32;
                tmem_ld32_nowait(tacc + (uint32_t)c0, vals[j]);
                if (X3) {
                    uint32_t b[32];
                    tmem_ld32_nowait(tacc + (uint32_t)(BN + c0), b);
                    tmem_ld_wait();
                    if (ALO_TMEM) {                              // cross + cross2 first (both ~2^-11 of main), then + main
                        uint32_t c2[32];
                        tmem_ld32_nowait(tacc + (uint32_t)(2 * BN + c0), c2);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) b[i] = __float_as_uint(__fadd_rn(__uint_as_float(b[i]), __uint_as_float(c2[i])));
                    }
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        vals[j][i] = __float_as_uint(__fadd_rn(__uint_as_float(vals[j][i]), __uint_as_float(b[i])));
                } else {
                    tmem_ld_wait();
                }
            }
            tc_fence_before();
            mbar_arrive(&tempty_bar[acc]);                       // accumulator drained: the MMA warps may reuse it
#pragma unroll
            for (int j = 0; j < SUBS; ++j, ++k) {
                const int c0 = (g + 2 * j) * 32;
                uint32_t (&a)[32] = vals[j];
                const uint32_t s = k & (ns - 1), ph = (k >> (ns - 1)) & 1;      // ns is 1 or 2: no division
                mbar_wait(&slot_ready[g * 2 + s], ph);          // residual landed / previous store has read the slot
                uint8_t* srow = sbase + (size_t)s * EPI_SLOT_BYTES;
                const float4* bs4 = reinterpret_cast<const float4*>(gbias + n0 + c0);   // warp-uniform: one broadcast per load
                if (args.out_transposed) {                       // slot = [32 columns][128 rows]: plain bias add, no swizzle
                    float* tcol = reinterpret_cast<float*>(staging + ((size_t)g * ns + s) * EPI_SLOT_BYTES) + row;
#pragma unroll
                    for (int i = 0; i < 32; ++i) tcol[i * TC_BM] = __uint_as_float(a[i]) + (gbias ? __ldg(gbias + n0 + c0 + i) : 0.f);
                    fence_proxy_async();
                    mbar_arrive(&out_ready[g * 2 + s]);
                    continue;
                }
                const uint32_t mword = mw[j];
                uint32_t oword = 0;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float4* p = reinterpret_cast<float4*>(srow + (((uint32_t)c ^ swz) << 4));
                    const float4 bv = gbias ? __ldg(bs4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                    float4 v = make_float4(__uint_as_float(a[4 * c]) + bv.x, __uint_as_float(a[4 * c + 1]) + bv.y,
                                           __uint_as_float(a[4 * c + 2]) + bv.z, __uint_as_float(a[4 * c + 3]) + bv.w);
                    if (has_res) { const float4 r = *p; v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
                    if (args.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                    if (!((mword >> (4 * c)) & 1u)) v.x = 0.f;
                    if (!((mword >> (4 * c + 1)) & 1u)) v.y = 0.f;
                    if (!((mword >> (4 * c + 2)) & 1u)) v.z = 0.f;
                    if (!((mword >> (4 * c + 3)) & 1u)) v.w = 0.f;
                    oword |= (v.x > 0.f ? 1u : 0u) << (4 * c) | (v.y > 0.f ? 1u : 0u) << (4 * c + 1) |
                             (v.z > 0.f ? 1u : 0u) << (4 * c + 2) | (v.w > 0.f ? 1u : 0u) << (4 * c + 3);
                    *p = v;
                }
                if (obits && valid) obits[(int64_t)((n0 + c0) / 32) * args.M + m] = oword;
                fence_proxy_async();                             // generic-proxy writes -> visible to the TMA store
                mbar_arrive(&out_ready[g * 2 + s]);
            }
            if (threadIdx.x == 256) TC_TRACE(6, t);
        }
    } else if (!EPI_TMA && warp >= 8) {
        // ===== epilogue v2 (strided scatter / f32 mask source) ============================================
        constexpr int UNITS = BN / 32;                          // 16-column units per warp (two warps per lane quarter)
        const int quarter = warp & 3;                           // TMEM lanes 32*quarter .. +31
        const int uhalf = (warp - 8) >> 2;                      // this warp's units: uhalf, uhalf + 2, ...
        const int lr = lane >> 2, lc = (lane & 3) * 2;
        const float* __restrict__ resid = args.residual;
        const float* __restrict__ masks = args.mask_src;
        int t = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
            const int m_tile = div_ntiles(tile, num_n_tiles), n_tile = tile - m_tile * num_n_tiles;
            const int n0 = n_tile * BN;
            const int acc = t % kAcc;
            const uint32_t aph = (uint32_t)(t / kAcc) & 1;
            // this thread's four rows: (i, h) -> tile row 32*quarter + 16*i + 8*h + lr
            int64_t off[4];
            bool valid[4];
#pragma unroll
            for (int ih = 0; ih < 4; ++ih) {
                const int64_t m = (int64_t)m_tile * TC_BM + quarter * 32 + (ih >> 1) * 16 + (ih & 1) * 8 + lr;
                valid[ih] = m < args.M;
                int64_t orow = m;
                if (args.out_s != 0 && valid[ih]) {
                    const int64_t pq = (int64_t)args.P * args.Q;
                    const int img = (int)(m / pq);
                    const int rem = (int)(m - (int64_t)img * pq);
                    const int p = rem / args.Q, q = rem - p * args.Q;
                    orow = ((int64_t)img * args.out_H + (p * args.out_s + args.out_h0)) * args.out_W + (q * args.out_s + args.out_w0);
                }
                off[ih] = orow * args.Cout + n0 + lc;
            }
            float2 res[2][8], mk[2][8];                         // [buffer][(i,h) x j]
            auto prefetch = [&](int buf, int c0) {
#pragma unroll
                for (int ih = 0; ih < 4; ++ih)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        float2 r = make_float2(0.f, 0.f), q = make_float2(1.f, 1.f);
                        if (valid[ih]) {
                            if (resid) r = __ldg(reinterpret_cast<const float2*>(resid + off[ih] + c0 + 8 * j));
                            if (masks) q = __ldg(reinterpret_cast<const float2*>(masks + off[ih] + c0 + 8 * j));
                        }
                        res[buf][ih * 2 + j] = r;
                        mk[buf][ih * 2 + j] = q;
                    }
            };
            prefetch(0, uhalf * 16);                            // independent of the accumulator: issued before the wait
            mbar_wait(&tfull_bar[acc], aph);
            tc_fence_after();
            if (threadIdx.x == 256) TC_TRACE(5, t);
            const uint32_t tacc = tmem_base + (uint32_t)acc * kAccCols + ((uint32_t)(quarter * 32) << 16);
#pragma unroll
            for (int uu = 0; uu < UNITS; ++uu) {
                const int c0 = (uhalf + 2 * uu) * 16;
                if (uu + 1 < UNITS) prefetch((uu + 1) & 1, c0 + 32);
                uint32_t a0[8], a1[8];                          // rows +0..15 and +16..31 of the quarter
                tmem_ld_16x256b_x2(tacc + (uint32_t)c0, a0);
                tmem_ld_16x256b_x2(tacc + (16u << 16) + (uint32_t)c0, a1);
                if (X3) {
                    uint32_t b0[8], b1[8];
                    tmem_ld_16x256b_x2(tacc + (uint32_t)(BN + c0), b0);
                    tmem_ld_16x256b_x2(tacc + (16u << 16) + (uint32_t)(BN + c0), b1);
                    if (ALO_TMEM) {                              // dual issuer: cross + cross2 first, then + main
                        uint32_t e0[8], e1[8];
                        tmem_ld_16x256b_x2(tacc + (uint32_t)(2 * BN + c0), e0);
                        tmem_ld_16x256b_x2(tacc + (16u << 16) + (uint32_t)(2 * BN + c0), e1);
                        tmem_ld_wait();
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            b0[k] = __float_as_uint(__fadd_rn(__uint_as_float(b0[k]), __uint_as_float(e0[k])));
                            b1[k] = __float_as_uint(__fadd_rn(__uint_as_float(b1[k]), __uint_as_float(e1[k])));
                        }
                    } else {
                        tmem_ld_wait();
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        a0[k] = __float_as_uint(__fadd_rn(__uint_as_float(a0[k]), __uint_as_float(b0[k])));
                        a1[k] = __float_as_uint(__fadd_rn(__uint_as_float(a1[k]), __uint_as_float(b1[k])));
                    }
                } else {
                    tmem_ld_wait();
                }
                if (uu == UNITS - 1) {                           // last TMEM read of the tile: hand the stage back before the
                    tc_fence_before();                           // (slow) global stores — the dual-issuer kernel has ONE stage
                    mbar_arrive(&tempty_bar[acc]);               // at BN = 128 and its MMA warps wait for exactly this
                }
                const float* bs = bias_s + n0 + c0 + lc;
#pragma unroll
                for (int ih = 0; ih < 4; ++ih)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const uint32_t* a = (ih >> 1) ? a1 : a0;
                        const int reg = 4 * j + 2 * (ih & 1);
                        float2 v = make_float2(__uint_as_float(a[reg]) + bs[8 * j], __uint_as_float(a[reg + 1]) + bs[8 * j + 1]);
                        const float2 r = res[uu & 1][ih * 2 + j], q = mk[uu & 1][ih * 2 + j];
                        v.x += r.x; v.y += r.y;
                        if (args.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); }
                        if (!(q.x > 0.f)) v.x = 0.f;
                        if (!(q.y > 0.f)) v.y = 0.f;
                        if (valid[ih]) *reinterpret_cast<float2*>(args.dst + off[ih] + c0 + 8 * j) = v;
                    }
            }
            if (threadIdx.x == 256) TC_TRACE(6, t);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair kernel: tcgen05.mma.cta_group::2 (M = 256: two SMs, 128 pixels each, ONE instruction for both) for the
// 3xTF32 / dual-issuer / TMA-epilogue configuration of the persistent kernel above.
//
// Why: a tcgen05.mma.kind::tf32 is accepted every ~237 cycles per issuing thread whatever its shape (profiles/
// r01_mma_issue_probe.txt), and every SM pulls the complete weight block of every k-step through L2.  A pair halves both:
// the leader CTA's two issuers drive both SMs, and each CTA loads only HALF of the weight rows of a k-step — the tensor core
// reads B from both CTAs' shared memory (N is split across the pair).
//
// Layout of one pipeline stage in EACH CTA (rank r of the cluster, h = BN/2):
//     A   [128 pixels x 32 ch]                this CTA's own pixel tile (im2col / tiled TMA), signals the LOCAL barrier afull
//     B   [h rows of B_hi | h rows of B_lo]   weight rows n0 + r*h .. +h-1, loaded with cta_group::2 TMA: the bytes of both
//                                             CTAs complete on the LEADER's barrier `full`
//   MMA1 (leader warp 1)  [d0 .. d0+2BN)  += a_hi x B  with N = 2*BN: each CTA contributes its BN rows, so the accumulator
//        columns are [main(0:h) | cross(0:h) | main(h:BN) | cross(h:BN)]  (main = a_hi b_hi, cross = a_hi b_lo)
//   MMA2 (leader warp 2)  [d0+2BN .. +BN) += a_lo (tensor memory, written by each CTA's split warps) x B with N = BN: each CTA
//        contributes its first h rows = its B_hi half -> cross2 = a_lo b_hi in natural column order.
// Barriers: afull / empty / tfull / out_ready / slot_ready are per CTA (tcgen05.commit multicasts to both); full / split /
// tempty live in the leader and receive remote arrivals (release.cluster) from the peer's producer, split and epilogue
// threads — the peer's split threads arrive only after they have seen their own A tile land, which is also what tells the
// leader's first issuer that the peer's a_hi is in place.  Everything behind the accumulator (epilogue warps, TMA stores,
// residual loads, mask bits) is the per-CTA code of the kernel above with the column map of the pair layout.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;       // shared::cluster address of the same offset in the even (leader) CTA

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    // default semantics (release at CTA scope), as cutlass::arch::ClusterBarrier::arrive(cta_id): a .release.cluster here
    // compiles to MEMBAR.ALL.GPU + CCTL.IVALL per arrive — measured ~1 us per k-step on every split thread.  What crosses
    // the pair is produced and consumed by the async / tensor-core proxies (TMA writes, tcgen05.st, tcgen05.mma reads),
    // ordered by the tcgen05 fences around the barrier.
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" :: "r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_leader(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;"
                 :: "r"(smem_u32(bar) & kPeerBitMask), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (clock64() - t0 > 8000000000LL) { printf("i2v conv_tc pair: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
    }
}
// cta_group::2 TMA: the transaction bytes complete on the LEADER CTA's barrier at the same offset
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma2_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma2_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
// arrives once on the barrier at this offset in BOTH CTAs of the pair when all prior MMAs of this thread are complete
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ constexpr uint32_t umma2_idesc_tf32(int n) {      // M = 256 across the pair
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

template <int BN, bool IM2COL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC2_THREADS, 1)
conv_tc_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
                    const __grid_constant__ CUtensorMap tmBlo, const __grid_constant__ CUtensorMap tmOut,
                    const __grid_constant__ CUtensorMap tmRes, const TcArgs args, const int stages,
                    const int num_m_tiles, const int num_n_tiles, const int epi_slots) {
    constexpr int H = BN / 2;                                   // weight rows of a k-step held by each CTA (per hi / lo)
    constexpr uint32_t kBBytes = (uint32_t)BN * TC_BK * 4;      // [h x b_hi | h x b_lo]
    constexpr uint32_t kStageBytes = TC_A_BYTES + kBBytes;
    constexpr uint32_t kAccCols = 3 * BN;                       // [MMA1: 2*BN | cross2: BN]
    constexpr int kAcc = (int)(384 / kAccCols);
    constexpr uint32_t kAloBase = 384, kAloSlots = 4, kTmemCols = 512;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* tiles = smem;
    uint8_t* staging = smem + (size_t)stages * kStageBytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + (size_t)epi_slots * 2 * EPI_SLOT_BYTES);   // leader's is used
    uint64_t* afull_bar = full_bar + stages;                    // local: this CTA's A tile landed
    uint64_t* empty_bar = afull_bar + stages;                   // local: both issuers are done with the stage (multicast commit)
    uint64_t* split_bar = empty_bar + stages;                   // leader's: 2 x 128 split threads
    uint64_t* tfull_bar = split_bar + stages;                   // local (multicast commit)
    uint64_t* tempty_bar = tfull_bar + kAcc;                    // leader's: 2 x 256 epilogue threads
    uint64_t* out_ready = tempty_bar + kAcc;
    uint64_t* slot_ready = out_ready + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(slot_ready + 4);

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int kiters = args.taps_h * args.taps_w * args.cblocks;
    const int pid = (int)(blockIdx.x >> 1), npairs = (int)(gridDim.x >> 1);
    const int num_ptiles = ((num_m_tiles + 1) >> 1) * num_n_tiles;   // a pair tile = two consecutive m-tiles x one n-tile

    auto stage_a = [&](int s) { return tiles + (size_t)s * kStageBytes; };
    auto stage_b = [&](int s) { return tiles + (size_t)s * kStageBytes + TC_A_BYTES; };

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA); prefetch_tmap(&tmBhi); prefetch_tmap(&tmBlo); prefetch_tmap(&tmOut); prefetch_tmap(&tmRes);
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full_bar[s], 2);                         // the two producers' arrive.expect_tx
            mbar_init(&afull_bar[s], 1);
            mbar_init(&empty_bar[s], 2);                        // both issuers release a stage
            mbar_init(&split_bar[s], 256);
        }
        for (int a = 0; a < kAcc; ++a) {
            mbar_init(&tfull_bar[a], 2);
            mbar_init(&tempty_bar[a], 2 * TC2_EPI_THREADS);
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(&out_ready[i], TC2_EPI_THREADS / 2);
            mbar_init(&slot_ready[i], 1);
        }
        fence_barrier_init();
    }
    cluster_sync_all();                                          // both CTAs' barriers exist before anyone arrives remotely
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer (both CTAs): own A tile -> afull (local); own half of the weight rows -> full (leader) =========
        if (elect_one()) {
            int it = 0, tno = 0;
            int st = 0; uint32_t ph = 0;               // ring position of k-step `it` (no division in the loop)
            for (int pt = pid; pt < num_ptiles; pt += npairs, ++tno) {
                const int pm = div_ntiles(pt, num_n_tiles), n_tile = pt - pm * num_n_tiles;
                const int m_tile = 2 * pm + (int)rank;
                const int64_t m0 = (int64_t)m_tile * TC_BM;
                const int n0 = n_tile * BN + (int)rank * H;
                int img = 0, base_h = 0, base_w = 0;
                if (IM2COL) {
                    const int64_t pq = (int64_t)args.P * args.Q;
                    img = (int)(m0 / pq);
                    const int rem = (int)(m0 - (int64_t)img * pq);
                    const int p = rem / args.Q, q = rem - p * args.Q;
                    base_h = p * args.stride + args.lower_h;
                    base_w = q * args.stride + args.lower_w;
                }
                for (int r = 0; r < args.taps_h; ++r)
                    for (int s = 0; s < args.taps_w; ++s)
                        for (int cb = 0; cb < args.cblocks; ++cb, ++it, ring_next(st, ph, stages)) {
                            mbar_wait(&empty_bar[st], ph ^ 1);
                            if ((r | s | cb) == 0) TC_TRACE(0, tno);
                            mbar_arrive_expect_tx(&afull_bar[st], TC_A_BYTES);
                            if (IM2COL) tma_load_im2col_4d(&tmA, &afull_bar[st], stage_a(st), cb * TC_BK, base_w, base_h, img, (uint16_t)s, (uint16_t)r);
                            else        tma_load_2d(&tmA, &afull_bar[st], stage_a(st), cb * TC_BK, (int)m0);
                            const int kcol = ((r * args.taps_w + s) * args.cblocks + cb) * TC_BK;
                            mbar_arrive_expect_tx_leader(&full_bar[st], kBBytes);
                            tma_load_2d_2sm(&tmBhi, &full_bar[st], stage_b(st), kcol, n0);
                            tma_load_2d_2sm(&tmBlo, &full_bar[st], stage_b(st) + (size_t)H * TC_BK * 4, kcol, n0);
                        }
                TC_TRACE(1, tno);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer 1 (leader): [main | cross] halves of both CTAs += a_hi x [b_hi half | b_lo half] ==================
        if (leader && elect_one()) {
            constexpr uint32_t idesc2 = umma2_idesc_tf32(2 * BN);
            int it = 0, t = 0;
            int st = 0; uint32_t ph = 0;               // ring position of k-step `it` (no division in the loop)
            for (int pt = pid; pt < num_ptiles; pt += npairs, ++t) {
                const int acc = t % kAcc;
                const uint32_t aph = (uint32_t)(t / kAcc) & 1;
                mbar_wait_cluster(&tempty_bar[acc], aph ^ 1);
                tc_fence_after();
                TC_TRACE(2, t);
                const uint32_t d0 = tmem_base + (uint32_t)acc * kAccCols;
                for (int kb = 0; kb < kiters; ++kb, ++it, ring_next(st, ph, stages)) {
                    mbar_wait_cluster(&full_bar[st], ph);        // the weight rows of both CTAs
                    mbar_wait(&afull_bar[st], ph);               // this CTA's a_hi
                    mbar_wait_cluster(&split_bar[st], ph);       // the peer's a_hi (its split threads saw it land)
                    tc_fence_after();
                    if (kb == 0) TC_TRACE(3, t);
                    const uint64_t da = umma_desc_sw128(smem_u32(stage_a(st)));
                    const uint64_t db = umma_desc_sw128(smem_u32(stage_b(st)));
#pragma unroll
                    for (int kk = 0; kk < TC_BK / 8; ++kk)
                        if (!(args.dbg & 1)) umma2_tf32(d0, da + 2 * kk, db + 2 * kk, idesc2, (kb | kk) ? 1u : 0u);
                    umma2_commit(&empty_bar[st]);
                }
                umma2_commit(&tfull_bar[acc]);
                TC_TRACE(4, t);
            }
        }
    } else if (warp == 2) {
        // ===== MMA issuer 2 (leader): cross2 += a_lo (tensor memory of each CTA) x b_hi halves ===============================
        if (leader && elect_one()) {
            constexpr uint32_t idesc = umma2_idesc_tf32(BN);
            int it = 0, t = 0;
            int st = 0; uint32_t ph = 0;               // ring position of k-step `it` (no division in the loop)
            for (int pt = pid; pt < num_ptiles; pt += npairs, ++t) {
                const int acc = t % kAcc;
                const uint32_t aph = (uint32_t)(t / kAcc) & 1;
                mbar_wait_cluster(&tempty_bar[acc], aph ^ 1);
                tc_fence_after();
                const uint32_t d2 = tmem_base + (uint32_t)acc * kAccCols + 2u * BN;
                for (int kb = 0; kb < kiters; ++kb, ++it, ring_next(st, ph, stages)) {
                    mbar_wait_cluster(&full_bar[st], ph);
                    mbar_wait_cluster(&split_bar[st], ph);       // a_lo of both CTAs is in its ring slot
                    tc_fence_after();
                    const uint64_t db = umma_desc_sw128(smem_u32(stage_b(st)));
                    const uint32_t talo = tmem_base + kAloBase + (uint32_t)(it % kAloSlots) * 32u;
#pragma unroll
                    for (int kk = 0; kk < TC_BK / 8; ++kk)
                        if (!(args.dbg & 2)) umma2_tf32_ts(d2, talo + 8u * kk, db + 2 * kk, idesc, (kb | kk) ? 1u : 0u);
                    umma2_commit(&empty_bar[st]);
                }
                umma2_commit(&tfull_bar[acc]);
            }
        }
    } else if (warp == 3) {
        // ===== epilogue TMA issuer of both groups (per CTA): output stores, residual loads ==================================
        if (elect_one()) {
            constexpr uint32_t SUBS = BN / 64;
            const bool has_res = args.residual != nullptr;
            const uint32_t ns = (uint32_t)epi_slots;
            const uint32_t my_tiles = (uint32_t)((num_ptiles - pid + npairs - 1) / npairs);
            const uint32_t total = my_tiles * SUBS;
            auto coords = [&](int g, uint32_t k, int& col, int& row) {
                const int pt = pid + (int)(k / SUBS) * npairs;
                const int pm = div_ntiles(pt, num_n_tiles), n_tile = pt - pm * num_n_tiles;
                col = n_tile * BN + (g + 2 * (int)(k % SUBS)) * 32;
                row = (2 * pm + (int)rank) * TC_BM;
            };
            auto refill = [&](int g, uint32_t k) {
                const uint32_t s = k & (ns - 1);
                if (has_res) {
                    int col, row;
                    coords(g, k, col, row);
                    mbar_arrive_expect_tx(&slot_ready[g * 2 + s], EPI_SLOT_BYTES);
                    tma_load_2d(&tmRes, &slot_ready[g * 2 + s], staging + ((size_t)g * ns + s) * EPI_SLOT_BYTES, col, row);
                } else {
                    mbar_arrive(&slot_ready[g * 2 + s]);
                }
            };
            for (int g = 0; g < 2; ++g)
                for (uint32_t k = 0; k < ns && k < total; ++k) refill(g, k);
            uint32_t kdone[2] = {0, 0};
            const long long t0 = clock64();
            while (kdone[0] < total || kdone[1] < total) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const uint32_t k = kdone[g];
                    if (k >= total) continue;
                    const uint32_t s = k & (ns - 1), ph = (k >> (ns - 1)) & 1;      // ns is 1 or 2: no division
                    if (!mbar_test_wait(&out_ready[g * 2 + s], ph)) continue;
                    int col, row;
                    coords(g, k, col, row);
                    const uint8_t* slot = staging + ((size_t)g * ns + s) * EPI_SLOT_BYTES;
                    if (col < args.store_cols && row < (int64_t)num_m_tiles * TC_BM) tma_store_2d(&tmOut, slot, col, row);
                    bulk_commit();
                    if (k + ns < total) {
                        bulk_wait_read0();
                        refill(g, k + ns);
                    }
                    kdone[g] = k + 1;
                }
                if (clock64() - t0 > 40000000000LL) { printf("i2v conv_tc pair: epilogue issuer timeout (block %d)\n", blockIdx.x); __trap(); }
            }
            bulk_wait_all();
        }
    } else if (warp >= 4 && warp < 8) {
        // ===== A split (both CTAs): a_lo of the own tile -> own tensor memory; arrive on the LEADER's split barrier ==========
        const int row = threadIdx.x - 128;
        const uint32_t swz = (uint32_t)(row & 7);
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kAloBase;
        int it = 0, tno = 0;
        int st = 0; uint32_t ph = 0;               // ring position of k-step `it` (no division in the loop)
        for (int pt = pid; pt < num_ptiles; pt += npairs, ++tno) {
            for (int kb = 0; kb < kiters; ++kb, ++it, ring_next(st, ph, stages)) {
                mbar_wait(&afull_bar[st], ph);
                if (args.dbg & 4) { mbar_arrive_leader(&split_bar[st]); continue; }      // timing experiment: no split work
                const uint8_t* arow = stage_a(st) + (size_t)row * 128;
                uint32_t lo[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 v = *reinterpret_cast<const float4*>(arow + (((uint32_t)c ^ swz) << 4));
                    lo[4 * c + 0] = __float_as_uint(v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u));
                    lo[4 * c + 1] = __float_as_uint(v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u));
                    lo[4 * c + 2] = __float_as_uint(v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u));
                    lo[4 * c + 3] = __float_as_uint(v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u));
                }
                tmem_st32(lane_addr + (uint32_t)(it % kAloSlots) * 32u, lo);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                mbar_arrive_leader(&split_bar[st]);
            }
            if (row == 0) TC_TRACE(7, tno);
        }
    } else if (warp >= 8) {
        // ===== epilogue v3 (per CTA), pair column map ======================================================================
        constexpr int SUBS = BN / 64;
        const int g = (warp - 8) >> 2;
        const int row = (warp & 3) * 32 + lane;
        const uint32_t swz = (uint32_t)(row & 7);
        const uint32_t ns = (uint32_t)epi_slots;
        uint8_t* sbase = staging + (size_t)g * ns * EPI_SLOT_BYTES + (size_t)row * 128;
        const bool has_res = args.residual != nullptr;
        const float* __restrict__ gbias = args.bias;
        const uint32_t* __restrict__ mbits = args.mask_bits;
        uint32_t* __restrict__ obits = args.bits_out;
        uint32_t k = 0;
        int t = 0;
        for (int pt = pid; pt < num_ptiles; pt += npairs, ++t) {
            const int pm = div_ntiles(pt, num_n_tiles), n_tile = pt - pm * num_n_tiles;
            const int m_tile = 2 * pm + (int)rank;
            const int n0 = n_tile * BN;
            const int acc = t % kAcc;
            const uint32_t aph = (uint32_t)(t / kAcc) & 1;
            const int64_t m = (int64_t)m_tile * TC_BM + row;
            const bool valid = m < args.M;
            uint32_t mw[SUBS];
#pragma unroll
            for (int j = 0; j < SUBS; ++j)
                mw[j] = (mbits && valid) ? __ldg(mbits + (int64_t)(n0 / 32 + g + 2 * j) * args.M + m) : 0xFFFFFFFFu;
            mbar_wait(&tfull_bar[acc], aph);
            tc_fence_after();
            if (threadIdx.x == 256) TC_TRACE(5, t);
            const uint32_t tacc = tmem_base + (uint32_t)acc * kAccCols + ((uint32_t)((warp & 3) * 32) << 16);
            uint32_t vals[SUBS][32];
#pragma unroll
            for (int j = 0; j < SUBS; ++j) {
                const int c0 = (g + 2 * j) * 32;                                   // output column block of this sub-tile
                const int cm = c0 < H ? c0 : BN + (c0 - H);                        // where MMA1 put main(c0 ..): see the layout
                uint32_t b[32];
                {
                    uint32_t c2[32];
                    tmem_ld32_nowait(tacc + (uint32_t)(cm + H), b);                // cross
                    tmem_ld32_nowait(tacc + (uint32_t)(2 * BN + c0), c2);          // cross2
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) b[i] = __float_as_uint(__fadd_rn(__uint_as_float(b[i]), __uint_as_float(c2[i])));
                }
                tmem_ld32_nowait(tacc + (uint32_t)cm, vals[j]);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    vals[j][i] = __float_as_uint(__fadd_rn(__uint_as_float(vals[j][i]), __uint_as_float(b[i])));
            }
            tc_fence_before();
            mbar_arrive_leader(&tempty_bar[acc]);
#pragma unroll
            for (int j = 0; j < SUBS; ++j, ++k) {
                const int c0 = (g + 2 * j) * 32;
                uint32_t (&a)[32] = vals[j];
                const uint32_t s = k & (ns - 1), ph = (k >> (ns - 1)) & 1;      // ns is 1 or 2: no division
                mbar_wait(&slot_ready[g * 2 + s], ph);
                uint8_t* srow = sbase + (size_t)s * EPI_SLOT_BYTES;
                const float4* bs4 = reinterpret_cast<const float4*>(gbias + n0 + c0);
                const uint32_t mword = mw[j];
                uint32_t oword = 0;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float4* p = reinterpret_cast<float4*>(srow + (((uint32_t)c ^ swz) << 4));
                    const float4 bv = gbias ? __ldg(bs4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                    float4 v = make_float4(__uint_as_float(a[4 * c]) + bv.x, __uint_as_float(a[4 * c + 1]) + bv.y,
                                           __uint_as_float(a[4 * c + 2]) + bv.z, __uint_as_float(a[4 * c + 3]) + bv.w);
                    if (has_res) { const float4 r = *p; v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
                    if (args.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                    if (!((mword >> (4 * c)) & 1u)) v.x = 0.f;
                    if (!((mword >> (4 * c + 1)) & 1u)) v.y = 0.f;
                    if (!((mword >> (4 * c + 2)) & 1u)) v.z = 0.f;
                    if (!((mword >> (4 * c + 3)) & 1u)) v.w = 0.f;
                    oword |= (v.x > 0.f ? 1u : 0u) << (4 * c) | (v.y > 0.f ? 1u : 0u) << (4 * c + 1) |
                             (v.z > 0.f ? 1u : 0u) << (4 * c + 2) | (v.w > 0.f ? 1u : 0u) << (4 * c + 3);
                    *p = v;
                }
                if (obits && valid) obits[(int64_t)((n0 + c0) / 32) * args.M + m] = oword;
                fence_proxy_async();
                mbar_arrive(&out_ready[g * 2 + s]);
            }
            if (threadIdx.x == 256) TC_TRACE(6, t);
        }
    }

    tc_fence_before();
    cluster_sync_all();                  // nobody leaves (or frees tensor memory) while the peer may still signal / read here
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// First-layer data gradient WITHOUT the Z^T scratch (i2v_conv_stem_dgrad_direct_f32; 7x7 / stride 2 / pad 3, 64 output
// channels, Q <= 128: ResNet's and DenseNet's stem).  The GEMM + col2im pair above moves 4.85 GB through HBM for 0.98 GB of
// tensors (ncu: GEMM 823 MB in, 1 998 MB of Z^T out; col2im 1 875 MB in): 1.24 ms per 256 frames.  Here the col2im happens on
// chip.  A tile is ONE dy row (n, p): 128 TMEM lanes = the row's Q pixels (lanes >= Q are junk and zeroed).  The contraction
// over the 64 output channels gives Z[q][(c,r,s)] in tensor memory (two N = 80 halves of the 147 padded taps, each half
// double-buffered against the epilogue; 3xTF32: main += a_hi b_hi, cross += a_hi b_lo + a_lo b_hi, A_lo in a TMEM ring).
// Epilogue thread q owns image columns 2q and 2q+1: column w receives Z[q'][c][r][s] with 2q' - 3 + s = w, i.e. its own
// s = 3 / 4, lane q+1's s = 1 / 2, lane q+2's s = 0 and lane q-1's s = 5 / 6 — warp shuffles, plus a 6-float-per-group
// exchange through shared memory at warp edges.  The sums go into a REGISTER window of the 7 image rows 2p-3 .. 2p+3 that
// dy row p touches (3 channels x 7 rows x 2 columns per thread); after row p the rows 2p-3 and 2p-2 are complete, are
// written to dcost/dimage [N,3,H,W] with coalesced float2 stores and the window shifts down by two.  A CTA walks a strip of
// consecutive dy rows (half an image: 2N strips over the SMs; the second half re-computes three rows), so every dx element
// is summed in a fixed order (dy rows ascending, taps s ascending): bit-reproducible, no atomics, no scratch.
// HBM traffic = dy once + dx once.
// ---------------------------------------------------------------------------------------------
constexpr int SD_THREADS = 640;          // warp 0 producer, 1 MMA issuer, 2 TMEM allocator, 4-7 A split, 8-19 epilogue (3 channels x 4 lane quarters)
constexpr int SD_STAGES = 3;             // A stages (one dy row each: two 16 KB k-blocks); the A_lo ring has 2 slots per stage
constexpr int SD_NZ = 160;               // UMMA N: the 147 taps k = (c,r,s) + 13 zero rows; 2 accumulator stages x 160 + 6 x 32 ring columns = 512
constexpr int SD_NC = 49;                // taps (columns) per channel: channel c at columns 49c .. 49c+48
constexpr uint32_t SD_B_TILE = SD_NZ * TC_BK * 4;      // 20 KB: 160 taps x 32 channels

struct StemDirectArgs {
    float* dx;
    int N, H, W, P, Q;
    int units;                           // strips: 2 per image (1 when the image is small)
    int strips_per_image;
    int dbg;                             // timing experiments ($I2V_STEM_DBG; wrong results): 1 no MMA, 2 no col2im math, 4 no split, 8 no TMEM loads
    int P2, Q2;                          // fused pooling variant: size of the pooled map (3x3 / stride 2 / pad 1 over P x Q)
};

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}

// strip u -> image, dy rows [pa, pb] it needs and image rows [ha, hb) it writes
__device__ __forceinline__ void sd_strip(const StemDirectArgs& a, int u, int& img, int& pa, int& pb, int& ha, int& hb) {
    img = u / a.strips_per_image;
    const int part = u - img * a.strips_per_image;
    const int hs = a.strips_per_image == 2 ? ((a.H / 2) & ~1) : a.H;      // even split row
    ha = part == 0 ? 0 : hs;
    hb = (part == 0 && a.strips_per_image == 2) ? hs : a.H;
    pa = ha - 3 > 0 ? (ha - 3 + 1) / 2 : 0;                              // dy rows p with 2p-3 <= h <= 2p+3 for some h in [ha, hb)
    pb = (hb - 1 + 3) / 2; if (pb > a.P - 1) pb = a.P - 1;
}

// Measured history of this kernel (256 frames, against 1.11-1.27 ms for the GEMM + col2im pair it replaces):
//   v1  one issuer, 48 N = 80 instructions per dy row, 4 epilogue warps, per-lane branches around every edge case  2.06 ms
//   v2  three issuers                                                                                               2.17 ms
//       (switching every MMA off saved 0.1 ms: never the problem; the epilogue math cost 1.28 ms — ~4000 instructions per row)
//   v3  branch-free epilogue (broadcast edge loads + selects), two accumulators                                     0.93 ms
//       (0.40 ms of it still the col2im math of 4 warps at ~0.5 IPC, 0.48 ms everything else)
//   v4  TWELVE epilogue warps — three per lane quarter, one channel each — share the math; A_lo back in tensor memory       0.64 ms
//       (math 0.07, TMEM loads 0.10, MMAs 0.08; 0.31 ms is the bare skeleton: with two stages a dy row's TMA round trip is
//       exposed every other row, and an L2 prefetch six rows ahead changed nothing)
//   v5  N = 160 (the channels' 49 taps back to back) frees tensor memory for a third A stage                               0.61 ms
//       (0.42 ms skeleton: with main + cross accumulators of 160 columns each only ONE accumulator stage fits, so every dy
//       row pays MMA -> commit -> epilogue load -> release -> next MMA in series)
//   v6  this one: the two cross products accumulate into the SAME accumulator as the main product (K = 64: 24 truncating
//       accumulations instead of 8, a bias of ~4e-6 relative instead of ~1.5e-6 — both far inside the 2e-5 the parity tests
//       allow and the ~2e-6 of an FP32 FMA chain), which makes room for TWO accumulator stages: MMAs of row t+1 overlap the
//       epilogue of row t.  One issuer, fixed order (bit-reproducible).
__global__ void __launch_bounds__(SD_THREADS, 1)
stem_dgrad_direct_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
                         const __grid_constant__ CUtensorMap tmBlo, const StemDirectArgs args) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* bhi = smem;                                       // [kb] tiles of 20 KB (160 taps x 32 channels)
    uint8_t* blo = smem + 2 * SD_B_TILE;
    uint8_t* atiles = smem + 4 * SD_B_TILE;                    // 80 KB = 1024-aligned; per stage: kb0, kb1
    float* edge = reinterpret_cast<float*>(atiles + (size_t)SD_STAGES * 2 * TC_A_BYTES);   // [3 channels][2 buffers][6 slots][7 groups][6]
    constexpr int kEdgeSlot = 7 * 6, kEdgeBuf = 6 * kEdgeSlot;
    uint64_t* bars = reinterpret_cast<uint64_t*>(edge + 3 * 2 * kEdgeBuf);
    uint64_t* bfull = bars;
    uint64_t* afull = bfull + 1;
    uint64_t* aempty = afull + SD_STAGES;
    uint64_t* splitb = aempty + SD_STAGES;
    uint64_t* tfull = splitb + SD_STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    auto a_tile = [&](int st, int kb) { return atiles + ((size_t)st * 2 + kb) * TC_A_BYTES; };

    for (int i = threadIdx.x; i < 3 * 2 * kEdgeBuf; i += SD_THREADS) edge[i] = 0.f;     // slots 0 and 5 are never written again
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA); prefetch_tmap(&tmBhi); prefetch_tmap(&tmBlo);
        mbar_init(bfull, 1);
        for (int s = 0; s < SD_STAGES; ++s) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], 1); mbar_init(&splitb[s], 4); }       // one arrival per warp
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 12); }
        fence_barrier_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;       // columns: accumulator stage a at 160 a, A_lo ring 320 + 32 (2 stage + kb)
    constexpr uint32_t kRing = 320;

    if (warp == 0) {
        if (elect_one()) {
            mbar_arrive_expect_tx(bfull, 4 * SD_B_TILE);          // the weights stay resident
            for (int kb = 0; kb < 2; ++kb) {
                tma_load_2d(&tmBhi, bfull, bhi + (size_t)kb * SD_B_TILE, kb * TC_BK, 0);
                tma_load_2d(&tmBlo, bfull, blo + (size_t)kb * SD_B_TILE, kb * TC_BK, 0);
            }
            int t = 0;
            for (int u = blockIdx.x; u < args.units; u += gridDim.x) {
                int img, pa, pb, ha, hb;
                sd_strip(args, u, img, pa, pb, ha, hb);
                // Two shared-memory stages cannot hide an HBM round trip per dy row (the A_lo ring in tensor memory caps the
                // stages at two), so the rows are pulled into L2 a few tiles ahead: the TMA loads below then hit L2.
                constexpr int kAhead = 6;
                for (int d = 0; d < kAhead && pa + d <= pb; ++d) {
                    const int mf = (img * args.P + pa + d) * args.Q;
                    tma_prefetch_l2_2d(&tmA, 0, mf); tma_prefetch_l2_2d(&tmA, TC_BK, mf);
                }
                for (int p = pa; p <= pb; ++p, ++t) {
                    const int st = t % SD_STAGES;
                    const uint32_t ph = (uint32_t)(t / SD_STAGES) & 1;
                    if (p + kAhead <= pb) {
                        const int mf = (img * args.P + p + kAhead) * args.Q;
                        tma_prefetch_l2_2d(&tmA, 0, mf); tma_prefetch_l2_2d(&tmA, TC_BK, mf);
                    }
                    mbar_wait(&aempty[st], ph ^ 1);
                    mbar_arrive_expect_tx(&afull[st], 2 * TC_A_BYTES);
                    const int m0 = (img * args.P + p) * args.Q;
                    tma_load_2d(&tmA, &afull[st], a_tile(st, 0), 0, m0);
                    tma_load_2d(&tmA, &afull[st], a_tile(st, 1), TC_BK, m0);
                }
            }
        }
    } else if (warp == 1) {
        // the MMA issuer: acc (+)= a_hi b_hi + a_hi b_lo + a_lo (tensor memory) b_hi, 24 instructions of N = 160 per dy row
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_tf32(SD_NZ);
            mbar_wait(bfull, 0);
            int t = 0;
            for (int u = blockIdx.x; u < args.units; u += gridDim.x) {
                int img, pa, pb, ha, hb;
                sd_strip(args, u, img, pa, pb, ha, hb);
                for (int p = pa; p <= pb; ++p, ++t) {
                    const int st = t % SD_STAGES;
                    const uint32_t ph = (uint32_t)(t / SD_STAGES) & 1;
                    const int acc = t & 1;
                    mbar_wait(&afull[st], ph);
                    mbar_wait(&splitb[st], ph);
                    mbar_wait(&tempty[acc], ((uint32_t)(t >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t dacc = tmem_base + (uint32_t)acc * SD_NZ;
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb) {
                        const uint64_t da = umma_desc_sw128(smem_u32(a_tile(st, kb)));
                        const uint64_t dbh = umma_desc_sw128(smem_u32(bhi + (size_t)kb * SD_B_TILE));
                        const uint64_t dbl = umma_desc_sw128(smem_u32(blo + (size_t)kb * SD_B_TILE));
                        const uint32_t talo = tmem_base + kRing + 32u * (uint32_t)(2 * st + kb);
#pragma unroll
                        for (int kk = 0; kk < TC_BK / 8; ++kk) {
                            if (args.dbg & 1) continue;
                            umma_tf32(dacc, da + 2 * kk, dbh + 2 * kk, idesc, (kb | kk) ? 1u : 0u);
                            umma_tf32(dacc, da + 2 * kk, dbl + 2 * kk, idesc, 1u);
                            umma_tf32_ts(dacc, talo + 8u * kk, dbh + 2 * kk, idesc, 1u);
                        }
                    }
                    umma_commit(&tfull[acc]);
                    umma_commit(&aempty[st]);
                }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // A split: a_lo of this thread's dy pixel (row of the tile), both k-blocks, into the stage's two ring slots
        const int row = threadIdx.x - 128;
        const uint32_t swz = (uint32_t)(row & 7);
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kRing;
        int t = 0;
        for (int u = blockIdx.x; u < args.units; u += gridDim.x) {
            int img, pa, pb, ha, hb;
            sd_strip(args, u, img, pa, pb, ha, hb);
            for (int p = pa; p <= pb; ++p, ++t) {
                const int st = t % SD_STAGES;
                const uint32_t ph = (uint32_t)(t / SD_STAGES) & 1;
                mbar_wait(&afull[st], ph);
                if (args.dbg & 4) { __syncwarp(); if (lane == 0) mbar_arrive(&splitb[st]); continue; }
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
                    const uint8_t* arow = a_tile(st, kb) + (size_t)row * 128;
                    uint32_t lo[32];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 v = *reinterpret_cast<const float4*>(arow + (((uint32_t)c ^ swz) << 4));
                        lo[4 * c + 0] = __float_as_uint(v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u));
                        lo[4 * c + 1] = __float_as_uint(v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u));
                        lo[4 * c + 2] = __float_as_uint(v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u));
                        lo[4 * c + 3] = __float_as_uint(v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u));
                    }
                    tmem_st32(lane_addr + 32u * (uint32_t)(2 * st + kb), lo);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&splitb[st]);           // one arrival per warp: 128 arrivals per row serialise on the barrier word
            }
        }
    } else if (warp >= 8) {
        // ===== epilogue: on-chip col2im; warp = (channel c, lane quarter); thread = dy pixel q = image columns 2q, 2q+1 =====
        const int c = (warp - 8) >> 2;
        const int ew = warp & 3;                                  // TMEM lane quarter of this warp
        const int q = ew * 32 + lane;
        const bool v1ok = q + 1 < args.Q, v0ok = q + 2 < args.Q;
        const bool l31 = lane == 31, l30 = lane == 30, l0 = lane == 0;
        const uint32_t tlane = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(c * SD_NC);
        float* ebase = edge + (size_t)c * 2 * kEdgeBuf;
        float win[7][2];                                          // [image row 2p-3+i][column 2q+e] of channel c
        int t = 0;
        for (int u = blockIdx.x; u < args.units; u += gridDim.x) {
            int img, pa, pb, ha, hb;
            sd_strip(args, u, img, pa, pb, ha, hb);
#pragma unroll
            for (int i = 0; i < 7; ++i) { win[i][0] = 0.f; win[i][1] = 0.f; }
            float* dxc = args.dx + ((int64_t)img * 3 + c) * args.H * args.W + 2 * q;
            auto emit = [&](int h, const float (&w0)[2]) {
                if (h < ha || h >= hb) return;
                float* o = dxc + (int64_t)h * args.W;
                if (2 * q + 1 < args.W) {
                    if ((args.W & 1) == 0) *reinterpret_cast<float2*>(o) = make_float2(w0[0], w0[1]);
                    else { o[0] = w0[0]; o[1] = w0[1]; }
                } else if (2 * q < args.W) o[0] = w0[0];
            };
            for (int p = pa; p <= pb; ++p, ++t) {
                const int acc = t & 1;
                mbar_wait(&tfull[acc], (uint32_t)(t >> 1) & 1);
                tc_fence_after();
                const uint32_t tacc = tlane + (uint32_t)acc * SD_NZ;
                float z[SD_NC];
                if (args.dbg & 8) {
#pragma unroll
                    for (int i = 0; i < SD_NC; ++i) z[i] = 1.f;
                } else {
                    uint32_t m0[16], m1[16], m2[16], m3[16];       // taps 0..47 and (last column of a load starting at 33) tap 48
                    tmem_ld16_nowait(tacc, m0);
                    tmem_ld16_nowait(tacc + 16u, m1);
                    tmem_ld16_nowait(tacc + 32u, m2);
                    tmem_ld16_nowait(tacc + 33u, m3);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        z[i] = __uint_as_float(m0[i]); z[16 + i] = __uint_as_float(m1[i]); z[32 + i] = __uint_as_float(m2[i]);
                    }
                    z[48] = __uint_as_float(m3[15]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);          // this warp's share of the accumulator stage is in registers
                if (!(args.dbg & 2)) {
                    float* eb = ebase + (size_t)(t & 1) * kEdgeBuf;
                    float* mine = eb + (size_t)(ew + 1) * kEdgeSlot;
#pragma unroll
                    for (int r = 0; r < 7; ++r) {                  // publish: lane 0 -> s = 0,1,2; lane 1 -> s = 0; lane 31 -> s = 5,6
                        if (l0) { mine[r * 6 + 0] = z[r * 7 + 0]; mine[r * 6 + 1] = z[r * 7 + 1]; mine[r * 6 + 2] = z[r * 7 + 2]; }
                        if (lane == 1) mine[r * 6 + 3] = z[r * 7 + 0];
                        if (l31) { mine[r * 6 + 4] = z[r * 7 + 5]; mine[r * 6 + 5] = z[r * 7 + 6]; }
                    }
                    asm volatile("bar.sync %0, 128;" :: "r"(1 + c) : "memory");
                    const float* en = eb + (size_t)(ew + 2) * kEdgeSlot;     // next warp (slot 5 = zeros for the last one)
                    const float* ep = eb + (size_t)ew * kEdgeSlot;           // previous warp (slot 0 = zeros for the first one)
#pragma unroll
                    for (int r = 0; r < 7; ++r) {
                        float v1 = __shfl_down_sync(0xffffffffu, z[r * 7 + 1], 1);
                        float v2 = __shfl_down_sync(0xffffffffu, z[r * 7 + 2], 1);
                        float v0 = __shfl_down_sync(0xffffffffu, z[r * 7 + 0], 2);
                        float v5 = __shfl_up_sync(0xffffffffu, z[r * 7 + 5], 1);
                        float v6 = __shfl_up_sync(0xffffffffu, z[r * 7 + 6], 1);
                        const float n0 = en[r * 6 + 0], n1 = en[r * 6 + 1], n2 = en[r * 6 + 2], n3 = en[r * 6 + 3];   // broadcast loads
                        const float p5 = ep[r * 6 + 4], p6 = ep[r * 6 + 5];
                        v1 = l31 ? n1 : v1; v2 = l31 ? n2 : v2;
                        v0 = l31 ? n3 : (l30 ? n0 : v0);
                        v5 = l0 ? p5 : v5; v6 = l0 ? p6 : v6;
                        v1 = v1ok ? v1 : 0.f; v2 = v1ok ? v2 : 0.f; v0 = v0ok ? v0 : 0.f;       // q+1 / q+2 beyond the dy row
                        // column 2q: s = 1 (q+1), 3 (q), 5 (q-1); column 2q+1: s = 0 (q+2), 2 (q+1), 4 (q), 6 (q-1)
                        win[r][0] = __fadd_rn(win[r][0], __fadd_rn(__fadd_rn(v1, z[r * 7 + 3]), v5));
                        win[r][1] = __fadd_rn(win[r][1], __fadd_rn(__fadd_rn(__fadd_rn(v0, v2), z[r * 7 + 4]), v6));
                    }
                } else {
                    win[0][0] += z[0] + z[48];
                }
                emit(2 * p - 3, win[0]);                           // image rows 2p-3 and 2p-2 are complete
                emit(2 * p - 2, win[1]);
#pragma unroll
                for (int i = 0; i < 5; ++i) { win[i][0] = win[i + 2][0]; win[i][1] = win[i + 2][1]; }
                win[5][0] = win[5][1] = win[6][0] = win[6][1] = 0.f;
            }
            // strip done: the window holds image rows 2 pb - 1 .. 2 pb + 3
#pragma unroll
            for (int i = 0; i < 5; ++i) emit(2 * pb - 1 + i, win[i]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
    }
}

__device__ __forceinline__ void lds128(uint32_t addr, uint32_t (&v)[4]) {
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(addr));
}
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// registers -> one TMEM lane per thread, 4 consecutive columns
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
constexpr int SP_AST = 2, SP_PST = 2;      // fused pooling variant: A stages / pooled-row stages
__host__ __device__ __forceinline__ uint32_t sp_gtile_bytes(int Q2) { return ((uint32_t)Q2 * 128u + 1023u) & ~1023u; }   // SWIZZLE_128B atom
__host__ __device__ __forceinline__ uint32_t sp_atile_bytes(int Q2) { return ((uint32_t)Q2 * 64u + 1023u) & ~1023u; }    // keeps every tile 1024-aligned

// The same kernel with the max pooling's backward pass fused in front of it (i2v_conv_stem_dgrad_pool_f32): the 822 MB
// gradient of the stem activation is never written or read.  The producer stages the one or two POOLED gradient rows (and
// their argmax rows) whose 3x3 / stride-2 / pad-1 windows cover stem row p; the four assemble warps (thread = stem pixel) gather
// g[c] = sum over the <= 4 covering windows with argmax == this position of the pooled gradient, in the window order of
// i2v_maxpool_bwd_f32 (bit-identical to running that kernel first), write a_hi into the swizzled A tile and a_lo into the
// tensor-memory ring.  Everything behind the A tile is the kernel above.
__global__ void __launch_bounds__(SD_THREADS, 1)
stem_dgrad_pool_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmAm,
                       const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo, const StemDirectArgs args) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* bhi = smem;                                       // [kb] tiles of 20 KB (160 taps x 32 channels)
    uint8_t* blo = smem + 2 * SD_B_TILE;
    uint8_t* atiles = smem + 4 * SD_B_TILE;                    // 80 KB = 1024-aligned; per stage: kb0, kb1
    constexpr int AST = SP_AST, PST = SP_PST;                  // A stages / pooled-row stages of this variant
    // pooled rows: per stage and pooled row i in {0, 1}: two SWIZZLE_128B tiles [Q2 pixels][32 channels] f32 (kb = 0, 1), then
    // per stage and i one SWIZZLE_64B tile [Q2 pixels][64 channels] u8.  A thread (stem pixel q) reads pooled pixels q/2 and
    // (q+1)/2: 16 different pixels per warp at the same channels — with dense rows that is a 16-way bank conflict; with the
    // TMA swizzles 8 consecutive pixels land in 8 different 16-byte bank groups (2 wavefronts for 256 bytes: conflict-free).
    uint8_t* prows = atiles + (size_t)AST * 2 * TC_A_BYTES;
    const uint32_t gtile = sp_gtile_bytes(args.Q2), atile = sp_atile_bytes(args.Q2);
    const uint32_t pstage = 4u * gtile + 2u * atile;           // [i][kb] gradient tiles, then [i] argmax tiles
    float* edge = reinterpret_cast<float*>(prows + (size_t)PST * pstage);   // [3 channels][2 buffers][6 slots][7 groups][6]
    constexpr int kEdgeSlot = 7 * 6, kEdgeBuf = 6 * kEdgeSlot;
    uint64_t* bars = reinterpret_cast<uint64_t*>(edge + 3 * 2 * kEdgeBuf);
    uint64_t* bfull = bars;
    uint64_t* aempty = bfull + 1;
    uint64_t* splitb = aempty + AST;
    uint64_t* tfull = splitb + AST;
    uint64_t* tempty = tfull + 2;
    uint64_t* pfull = tempty + 2;                               // pooled rows landed
    uint64_t* pempty = pfull + PST;                             // ... consumed by the 4 assemble warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pempty + PST);
    // 64 bytes of argmax = 255 ("dead window"): where the windows that do not exist for a pixel point their argmax loads
    uint8_t* dead64 = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tmem_slot + 1) + 63) & ~(uintptr_t)63);

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    auto a_tile = [&](int st, int kb) { return atiles + ((size_t)st * 2 + kb) * TC_A_BYTES; };

    for (int i = threadIdx.x; i < 3 * 2 * kEdgeBuf; i += SD_THREADS) edge[i] = 0.f;     // slots 0 and 5 are never written again
    if (threadIdx.x < 16) reinterpret_cast<uint32_t*>(dead64)[threadIdx.x] = 0xFFFFFFFFu;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmG); prefetch_tmap(&tmAm); prefetch_tmap(&tmBhi); prefetch_tmap(&tmBlo);
        for (int s2 = 0; s2 < PST; ++s2) { mbar_init(&pfull[s2], 1); mbar_init(&pempty[s2], 4); }
        mbar_init(bfull, 1);
        for (int s = 0; s < AST; ++s) { mbar_init(&aempty[s], 1); mbar_init(&splitb[s], 4); }       // one arrival per warp
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 12); }
        fence_barrier_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;       // columns: accumulator stage a at 160 a, A_lo ring 320 + 32 (2 stage + kb)
    constexpr uint32_t kRing = 320;

    if (warp == 0) {
        if (elect_one()) {
            mbar_arrive_expect_tx(bfull, 4 * SD_B_TILE);          // the weights stay resident
            for (int kb = 0; kb < 2; ++kb) {
                tma_load_2d(&tmBhi, bfull, bhi + (size_t)kb * SD_B_TILE, kb * TC_BK, 0);
                tma_load_2d(&tmBlo, bfull, blo + (size_t)kb * SD_B_TILE, kb * TC_BK, 0);
            }
            const uint32_t row_tx = (uint32_t)args.Q2 * (64u * 4u + 64u);      // bytes of one pooled row: gradient + argmax
            int t = 0;
            for (int u = blockIdx.x; u < args.units; u += gridDim.x) {
                int img, pa, pb, ha, hb;
                sd_strip(args, u, img, pa, pb, ha, hb);
                for (int p = pa; p <= pb; ++p, ++t) {
                    // stem row p is covered by the pooling windows of pooled rows p/2 (p even) or (p-1)/2 and (p+1)/2 (p odd)
                    const int ps = t % PST;
                    const uint32_t ph = (uint32_t)(t / PST) & 1;
                    const int ra = p >> 1, rb = (p + 1) >> 1;
                    const int nrows = (rb != ra && rb < args.P2) ? 2 : 1;
                    mbar_wait(&pempty[ps], ph ^ 1);
                    mbar_arrive_expect_tx(&pfull[ps], (uint32_t)nrows * row_tx);
                    uint8_t* gdst = prows + (size_t)ps * pstage;
                    uint8_t* adst = gdst + 4 * gtile;
                    for (int i = 0; i < nrows; ++i) {
                        const int m0 = (img * args.P2 + (i ? rb : ra)) * args.Q2;
                        tma_load_2d(&tmG, &pfull[ps], gdst + (size_t)(2 * i) * gtile, 0, m0);
                        tma_load_2d(&tmG, &pfull[ps], gdst + (size_t)(2 * i + 1) * gtile, TC_BK, m0);
                        tma_load_2d(&tmAm, &pfull[ps], adst + (size_t)i * atile, 0, m0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // the MMA issuer: acc (+)= a_hi b_hi + a_hi b_lo + a_lo (tensor memory) b_hi, 24 instructions of N = 160 per dy row
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_tf32(SD_NZ);
            mbar_wait(bfull, 0);
            int t = 0;
            for (int u = blockIdx.x; u < args.units; u += gridDim.x) {
                int img, pa, pb, ha, hb;
                sd_strip(args, u, img, pa, pb, ha, hb);
                for (int p = pa; p <= pb; ++p, ++t) {
                    const int st = t % AST;
                    const uint32_t ph = (uint32_t)(t / AST) & 1;
                    const int acc = t & 1;
                    mbar_wait(&splitb[st], ph);                  // the assemble warps have built a_hi (shared memory) and a_lo (tensor memory)
                    mbar_wait(&tempty[acc], ((uint32_t)(t >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t dacc = tmem_base + (uint32_t)acc * SD_NZ;
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb) {
                        const uint64_t da = umma_desc_sw128(smem_u32(a_tile(st, kb)));
                        const uint64_t dbh = umma_desc_sw128(smem_u32(bhi + (size_t)kb * SD_B_TILE));
                        const uint64_t dbl = umma_desc_sw128(smem_u32(blo + (size_t)kb * SD_B_TILE));
                        const uint32_t talo = tmem_base + kRing + 32u * (uint32_t)(2 * st + kb);
#pragma unroll
                        for (int kk = 0; kk < TC_BK / 8; ++kk) {
                            if (args.dbg & 1) continue;
                            umma_tf32(dacc, da + 2 * kk, dbh + 2 * kk, idesc, (kb | kk) ? 1u : 0u);
                            umma_tf32(dacc, da + 2 * kk, dbl + 2 * kk, idesc, 1u);
                            umma_tf32_ts(dacc, talo + 8u * kk, dbh + 2 * kk, idesc, 1u);
                        }
                    }
                    umma_commit(&tfull[acc]);
                    umma_commit(&aempty[st]);
                }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ===== assemble: thread = stem pixel q; its 64-channel gradient = max-pool backward of the pooled gradient (gather form,
        // windows in (P', Q') order exactly as i2v_maxpool_bwd_f32; argmax 255 = the forward pass marked the window dead) ======
        const int row = threadIdx.x - 128;
        const uint32_t swz = (uint32_t)(row & 7);
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kRing;
        const bool qv = row < args.Q;
        // pooled columns Q' with 2Q'-1 <= q <= 2Q'+1: window j = 0 is pooled pixel q/2, window j = 1 pooled pixel (q+1)/2 (odd q)
        const int qa = row >> 1, qb = (row + 1) >> 1;
        const bool onj[2] = {qv, qv && qb != qa && qb < args.Q2};
        // Addresses are XOR-composed: a tile base is 1024-aligned and a pixel's record 128 (gradient) / 64 (argmax) bytes, so
        // base + swizzled chunk + word = (base ^ swizzle bits) ^ (loop bits) — one LOP3 per load inside the loop.  Windows that
        // do not exist read the 64-byte block of 255s (never a winner) and the first gradient row (never added).
        uint32_t goff[2], aoff[2];                                  // relative to the stage's gradient / argmax tile of pooled row i
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int qc = j ? qb : qa;
            goff[j] = onj[j] ? ((uint32_t)qc * 128u) ^ ((uint32_t)(qc & 7) << 4) : 0u;
            aoff[j] = onj[j] ? ((uint32_t)qc * 64u) ^ ((uint32_t)((qc >> 1) & 3) << 4) : 0u;
        }
        const uint32_t dead_addr = smem_u32(dead64);
        const int wq[2] = {row - (2 * qa - 1), row - (2 * qb - 1)};          // column of this pixel inside the window of qa / qb
        int t = 0;
        for (int u = blockIdx.x; u < args.units; u += gridDim.x) {
            int img, pa, pb, ha, hb;
            sd_strip(args, u, img, pa, pb, ha, hb);
            for (int p = pa; p <= pb; ++p, ++t) {
                const int st = t % AST, ps = t % PST;
                const uint32_t ph = (uint32_t)(t / AST) & 1, pph = (uint32_t)(t / PST) & 1;
                const int ra = p >> 1, rb = (p + 1) >> 1;
                const bool two_p = rb != ra && rb < args.P2;
                mbar_wait(&pfull[ps], pph);
                mbar_wait(&aempty[st], ph ^ 1);
                tc_fence_after();
                if (args.dbg & 4) { __syncwarp(); if (lane == 0) { mbar_arrive(&splitb[st]); mbar_arrive(&pempty[ps]); } continue; }
                const uint32_t gsm = smem_u32(prows + (size_t)ps * pstage), asm_ = gsm + 4 * gtile;
                // NI = pooled rows whose windows cover stem row p (warp-uniform): a compile-time constant per instantiation
                auto gather = [&](auto ni_tag) {
                    constexpr int NI = decltype(ni_tag)::value;
                    uint32_t want[NI][2], aB[NI][2];
#pragma unroll
                    for (int i = 0; i < NI; ++i)
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            // the window position this pixel has in window (i, j), replicated into the four bytes of a word
                            want[i][j] = (uint32_t)((p - (2 * (i ? rb : ra) - 1)) * 3 + wq[j]) * 0x01010101u;
                            aB[i][j] = onj[j] ? asm_ + (uint32_t)i * atile + aoff[j] : dead_addr;
                        }
                    // REAL loops over the 16 four-channel chunks of the pixel (2 k-blocks x 8): fully unrolled the body was
                    // 52 KB of straight-line code run once per row, and 42 % of the assemble warps' stall samples were
                    // instruction-cache misses (ncu source page); a_lo goes to tensor memory four columns at a time
#pragma unroll 1
                    for (int kb = 0; kb < 2; ++kb) {
                        uint32_t gB[NI][2];
#pragma unroll
                        for (int i = 0; i < NI; ++i)
#pragma unroll
                            for (int j = 0; j < 2; ++j) gB[i][j] = gsm + (uint32_t)(2 * i + kb) * gtile + goff[j];
                        const uint32_t arow = smem_u32(a_tile(st, kb)) + (uint32_t)row * 128u;
                        const uint32_t tcol = lane_addr + 32u * (uint32_t)(2 * st + kb);
#pragma unroll 2
                        for (int c4 = 0; c4 < 8; ++c4) {
                            float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                            for (int i = 0; i < NI; ++i)
#pragma unroll
                                for (int j = 0; j < 2; ++j) {
                                    // x has a zero byte where argmax == this pixel's position; the adds happen in window
                                    // order under predicates (exact selection; what a window that is off loads is never added)
                                    const uint32_t x = lds32(aB[i][j] ^ (uint32_t)((kb * 8 + c4) << 2)) ^ want[i][j];
                                    uint32_t gv[4];
                                    lds128(gB[i][j] ^ ((uint32_t)c4 << 4), gv);
                                    if (!(x & 0x000000ffu)) g[0] += __uint_as_float(gv[0]);
                                    if (!(x & 0x0000ff00u)) g[1] += __uint_as_float(gv[1]);
                                    if (!(x & 0x00ff0000u)) g[2] += __uint_as_float(gv[2]);
                                    if (!(x & 0xff000000u)) g[3] += __uint_as_float(gv[3]);
                                }
                            uint32_t lo4[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                lo4[j] = __float_as_uint(g[j] - __uint_as_float(__float_as_uint(g[j]) & 0xFFFFE000u));
                            sts128(arow + (((uint32_t)c4 ^ swz) << 4), g[0], g[1], g[2], g[3]);
                            tmem_st4(tcol + 4u * (uint32_t)c4, lo4);
                        }
                    }
                };
                if (two_p) gather(std::integral_constant<int, 2>{}); else gather(std::integral_constant<int, 1>{});
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                if (!(args.dbg & 64)) fence_proxy_async();         // generic-proxy writes of a_hi -> visible to the tensor core
                __syncwarp();
                if (lane == 0) { mbar_arrive(&splitb[st]); mbar_arrive(&pempty[ps]); }
            }
        }
    } else if (warp >= 8) {
        // ===== epilogue: on-chip col2im; warp = (channel c, lane quarter); thread = dy pixel q = image columns 2q, 2q+1 =====
        const int c = (warp - 8) >> 2;
        const int ew = warp & 3;                                  // TMEM lane quarter of this warp
        const int q = ew * 32 + lane;
        const bool v1ok = q + 1 < args.Q, v0ok = q + 2 < args.Q;
        const bool l31 = lane == 31, l30 = lane == 30, l0 = lane == 0;
        const uint32_t tlane = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(c * SD_NC);
        float* ebase = edge + (size_t)c * 2 * kEdgeBuf;
        float win[7][2];                                          // [image row 2p-3+i][column 2q+e] of channel c
        int t = 0;
        for (int u = blockIdx.x; u < args.units; u += gridDim.x) {
            int img, pa, pb, ha, hb;
            sd_strip(args, u, img, pa, pb, ha, hb);
#pragma unroll
            for (int i = 0; i < 7; ++i) { win[i][0] = 0.f; win[i][1] = 0.f; }
            float* dxc = args.dx + ((int64_t)img * 3 + c) * args.H * args.W + 2 * q;
            auto emit = [&](int h, const float (&w0)[2]) {
                if (h < ha || h >= hb) return;
                float* o = dxc + (int64_t)h * args.W;
                if (2 * q + 1 < args.W) {
                    if ((args.W & 1) == 0) *reinterpret_cast<float2*>(o) = make_float2(w0[0], w0[1]);
                    else { o[0] = w0[0]; o[1] = w0[1]; }
                } else if (2 * q < args.W) o[0] = w0[0];
            };
            for (int p = pa; p <= pb; ++p, ++t) {
                const int acc = t & 1;
                mbar_wait(&tfull[acc], (uint32_t)(t >> 1) & 1);
                tc_fence_after();
                const uint32_t tacc = tlane + (uint32_t)acc * SD_NZ;
                float z[SD_NC];
                if (args.dbg & 8) {
#pragma unroll
                    for (int i = 0; i < SD_NC; ++i) z[i] = 1.f;
                } else {
                    uint32_t m0[16], m1[16], m2[16], m3[16];       // taps 0..47 and (last column of a load starting at 33) tap 48
                    tmem_ld16_nowait(tacc, m0);
                    tmem_ld16_nowait(tacc + 16u, m1);
                    tmem_ld16_nowait(tacc + 32u, m2);
                    tmem_ld16_nowait(tacc + 33u, m3);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        z[i] = __uint_as_float(m0[i]); z[16 + i] = __uint_as_float(m1[i]); z[32 + i] = __uint_as_float(m2[i]);
                    }
                    z[48] = __uint_as_float(m3[15]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);          // this warp's share of the accumulator stage is in registers
                if (!(args.dbg & 2)) {
                    float* eb = ebase + (size_t)(t & 1) * kEdgeBuf;
                    float* mine = eb + (size_t)(ew + 1) * kEdgeSlot;
#pragma unroll
                    for (int r = 0; r < 7; ++r) {                  // publish: lane 0 -> s = 0,1,2; lane 1 -> s = 0; lane 31 -> s = 5,6
                        if (l0) { mine[r * 6 + 0] = z[r * 7 + 0]; mine[r * 6 + 1] = z[r * 7 + 1]; mine[r * 6 + 2] = z[r * 7 + 2]; }
                        if (lane == 1) mine[r * 6 + 3] = z[r * 7 + 0];
                        if (l31) { mine[r * 6 + 4] = z[r * 7 + 5]; mine[r * 6 + 5] = z[r * 7 + 6]; }
                    }
                    asm volatile("bar.sync %0, 128;" :: "r"(1 + c) : "memory");
                    const float* en = eb + (size_t)(ew + 2) * kEdgeSlot;     // next warp (slot 5 = zeros for the last one)
                    const float* ep = eb + (size_t)ew * kEdgeSlot;           // previous warp (slot 0 = zeros for the first one)
#pragma unroll
                    for (int r = 0; r < 7; ++r) {
                        float v1 = __shfl_down_sync(0xffffffffu, z[r * 7 + 1], 1);
                        float v2 = __shfl_down_sync(0xffffffffu, z[r * 7 + 2], 1);
                        float v0 = __shfl_down_sync(0xffffffffu, z[r * 7 + 0], 2);
                        float v5 = __shfl_up_sync(0xffffffffu, z[r * 7 + 5], 1);
                        float v6 = __shfl_up_sync(0xffffffffu, z[r * 7 + 6], 1);
                        const float n0 = en[r * 6 + 0], n1 = en[r * 6 + 1], n2 = en[r * 6 + 2], n3 = en[r * 6 + 3];   // broadcast loads
                        const float p5 = ep[r * 6 + 4], p6 = ep[r * 6 + 5];
                        v1 = l31 ? n1 : v1; v2 = l31 ? n2 : v2;
                        v0 = l31 ? n3 : (l30 ? n0 : v0);
                        v5 = l0 ? p5 : v5; v6 = l0 ? p6 : v6;
                        v1 = v1ok ? v1 : 0.f; v2 = v1ok ? v2 : 0.f; v0 = v0ok ? v0 : 0.f;       // q+1 / q+2 beyond the dy row
                        // column 2q: s = 1 (q+1), 3 (q), 5 (q-1); column 2q+1: s = 0 (q+2), 2 (q+1), 4 (q), 6 (q-1)
                        win[r][0] = __fadd_rn(win[r][0], __fadd_rn(__fadd_rn(v1, z[r * 7 + 3]), v5));
                        win[r][1] = __fadd_rn(win[r][1], __fadd_rn(__fadd_rn(__fadd_rn(v0, v2), z[r * 7 + 4]), v6));
                    }
                } else {
                    win[0][0] += z[0] + z[48];
                }
                emit(2 * p - 3, win[0]);                           // image rows 2p-3 and 2p-2 are complete
                emit(2 * p - 2, win[1]);
#pragma unroll
                for (int i = 0; i < 5; ++i) { win[i][0] = win[i + 2][0]; win[i][1] = win[i + 2][1]; }
                win[5][0] = win[5][1] = win[6][0] = win[6][1] = 0.f;
            }
            // strip done: the window holds image rows 2 pb - 1 .. 2 pb + 3
#pragma unroll
            for (int i = 0; i < 5; ++i) emit(2 * pb - 1 + i, win[i]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// 3x3 / stride 1 / pad 1 convolutions with every input patch delivered ONCE ("halo" kernel; 3xTF32, TMA epilogue).
// The im2col-mode main loop above pulls a 128-pixel x 32-channel tile through L2 for each of the nine filter taps: the same
// input pixels nine times (3.7 GB of L2 -> shared-memory traffic per 256-frame launch of a 64 -> 64 layer for 0.4 GB of
// tensors).  Here a tile is R whole output rows of one image laid out on the ZERO-PADDED raster, position j = r (W+2) + q
// (R (W+2) <= 128: the two padding columns of every row are junk GEMM rows, 3-7 % of the tile), and the activations a
// 32-channel block of it needs are ONE tiled TMA box {32 ch, W+2, R+2 rows} starting at (-1, p0-1) — out-of-bounds pixels
// arrive as zeros, which IS the padding.  In that patch the operand of filter tap (r, s) is the SAME 128 rows shifted by
// r (W+2) + s rows: the tensor core applies the 128-byte swizzle to absolute shared-memory address bits, so a descriptor
// whose start address is shifted by any number of 128-byte rows reads exactly those rows (tools/mma_shift_probe.py, exact on
// every shift).  Per 32-channel block: one 30 KB load instead of nine 16 KB ones, and the 3xTF32 split (a_lo = a - tf32(a))
// is computed once per patch instead of once per tap, into a second patch that the a_lo x b_hi instruction reads.
//   warp 0: producer — the weight ring (per tap and block: [b_hi | b_lo]) and the two patch slots; the next patch is requested
//           between weight stages the moment its slot frees
//   warp 1: a_hi x [b_hi | b_lo] -> [main | cross]                     warp 2: a_lo x b_hi -> cross2
//   warp 3: TMA stores of the epilogue groups                          warps 4-7: split   warps 8-15: epilogue
// Epilogue: thread = padded-raster position; valid positions write their 32-column slice into the staging slot at the DENSE
// row (r W + q), so that one 2-D TMA store of R W rows ships a tile (a second map with fewer rows for an image's last tile).
// ---------------------------------------------------------------------------------------------
constexpr int HL_THREADS = 512;

struct HaloArgs {
    const float* bias;
    const uint32_t* mask_bits;   // [Cout/32][M] or null (data gradient: ReLU-backward mask of dst)
    uint32_t* bits_out;          // [Cout/32][M] or null (forward: activity bits of dst)
    int N, H, W, PW, R, TI;      // padded width W + 2, output rows per tile, tiles per image
    int CB;                      // Cin / 32
    int relu;
    int rows_load;               // (R + 2) * PW rows per patch load
    uint32_t patch_bytes;        // one patch (raw or lo), multiple of 1024, >= (2 PW + 2 + 128) rows
    int bstages;
    int tail_rows;               // H % R: rows of an image's last tile when it is short (0: all tiles full)
    int dbg;                     // timing experiments ($I2V_TC_HALO_DBG, wrong results): bit 0 skips the main MMAs, bit 1 the a_lo MMAs, bit 2 the split, bit 3 the weight loads
    unsigned long long* trace;   // i2v_conv_tc_set_trace: [tile][0] first patch requested, [1] last weight stage requested, [2] issuer owns the
    int trace_tiles;             // accumulator, [3] first patch landed, [4] last MMA issued, [5] epilogue sees the accumulator, [6] epilogue done, [7] split done
    int64_t M;                   // N * H * W
};

template <int BN, bool DUAL>
__global__ void __launch_bounds__(HL_THREADS, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmBhi,
                    const __grid_constant__ CUtensorMap tmBlo, const __grid_constant__ CUtensorMap tmOut,
                    const __grid_constant__ CUtensorMap tmOutTail, const HaloArgs args, const int num_tiles, const int num_n_tiles) {
    // DUAL: two issuers, [main | cross | cross2], 2 accumulator stages at BN = 64 but only 1 at BN = 128.
    // !DUAL: ONE issuer (behind elect.sync an issuer keeps the tensor pipe fed on its own) adds a_lo x b_hi into `cross` after
    // a_hi x [b_hi | b_lo]: [main | cross], 2 stages at either width (512 columns at BN = 128).
    constexpr uint32_t kAccCols = (DUAL ? 3 : 2) * BN;
    constexpr int kAcc = DUAL ? 384 / kAccCols : 2;
    constexpr uint32_t kIssuers = DUAL ? 2 : 1;
    constexpr uint32_t kBStage = 2u * BN * 128u;                // [b_hi | b_lo], 32 k-values per row
    constexpr int SUBS = BN / 64;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t pbytes = args.patch_bytes;
    const int bstages = args.bstages;
    uint8_t* patches = smem;                                    // [2 slots][raw | lo]
    uint8_t* bring = patches + 4 * (size_t)pbytes;
    uint8_t* staging = bring + (size_t)bstages * kBStage;       // one 16 KB slot per epilogue group
    uint64_t* pfull = reinterpret_cast<uint64_t*>(staging + 2 * EPI_SLOT_BYTES);
    uint64_t* pempty = pfull + 2;
    uint64_t* psplit = pempty + 2;
    uint64_t* bfull = psplit + 2;
    uint64_t* bempty = bfull + 8;
    uint64_t* tfull = bempty + 8;
    uint64_t* tempty = tfull + 2;
    uint64_t* out_ready = tempty + 2;
    uint64_t* slot_ready = out_ready + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(slot_ready + 2);

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int CB = args.CB, PW = args.PW;
    auto patch_raw = [&](int ps) { return patches + (size_t)ps * 2 * pbytes; };
    auto patch_lo = [&](int ps) { return patches + (size_t)ps * 2 * pbytes + pbytes; };
    auto bstage = [&](int st) { return bring + (size_t)st * kBStage; };

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmX); prefetch_tmap(&tmBhi); prefetch_tmap(&tmBlo); prefetch_tmap(&tmOut); prefetch_tmap(&tmOutTail);
        for (int i = 0; i < 2; ++i) { mbar_init(&pfull[i], 1); mbar_init(&pempty[i], kIssuers); mbar_init(&psplit[i], 128); }
        for (int i = 0; i < bstages; ++i) { mbar_init(&bfull[i], 1); mbar_init(&bempty[i], kIssuers); }
        for (int i = 0; i < kAcc; ++i) { mbar_init(&tfull[i], kIssuers); mbar_init(&tempty[i], TC2_EPI_THREADS); }
        for (int i = 0; i < 2; ++i) { mbar_init(&out_ready[i], TC2_EPI_THREADS / 2); mbar_init(&slot_ready[i], 1); }
        fence_barrier_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // tile -> (image, first output row, first output channel)
    auto tile_of = [&](int tile, int& img, int& p0, int& n0, bool& tail) {
        const int mt = div_ntiles(tile, num_n_tiles);
        n0 = (tile - mt * num_n_tiles) * BN;
        img = mt / args.TI;
        const int ti = mt - img * args.TI;
        p0 = ti * args.R;
        tail = args.tail_rows != 0 && ti == args.TI - 1;
    };

    if (warp == 0) {
        // ===== producer (one elected thread): the weight ring, and the patches (one box per 32-channel block, two slots).  The
        // patch of block k is requested, blocking, before the weights of block k (the issuers cannot pass block k without it, and its
        // slot frees on the MMAs of block k - 2, whose weights are already in flight: no deadlock); the patch of block k + 1 is
        // requested between the weight stages of block k the moment its slot frees (try_wait).
        if (elect_one()) {
            int bi = 0, wi = 0, st = 0;
            uint32_t ph = 0;
            int ptile = blockIdx.x, pcb = 0, pi = 0;             // the next patch to request
            auto issue_patch = [&]() {
                int img, p0, n0; bool tail;
                tile_of(ptile, img, p0, n0, tail);
                const int ps = pi & 1;
                if (pcb == 0) TC_TRACE(0, (ptile - (int)blockIdx.x) / (int)gridDim.x);
                mbar_arrive_expect_tx(&pfull[ps], (uint32_t)args.rows_load * 128u);
                tma_load_4d(&tmX, &pfull[ps], patch_raw(ps), pcb * TC_BK, -1, p0 - 1, img);
                ++pi;
                if (++pcb == CB) { pcb = 0; ptile += gridDim.x; }
            };
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int img, p0, n0; bool tail;
                tile_of(tile, img, p0, n0, tail);
                for (int cb = 0; cb < CB; ++cb, ++wi) {
                    while (pi <= wi && ptile < num_tiles) {
                        mbar_wait(&pempty[pi & 1], ((uint32_t)(pi >> 1) & 1) ^ 1);
                        issue_patch();
                    }
                    for (int tap = 0; tap < 9; ++tap, ++bi, ring_next(st, ph, bstages)) {
                        if (pi == wi + 1 && ptile < num_tiles && mbar_test_wait(&pempty[pi & 1], ((uint32_t)(pi >> 1) & 1) ^ 1)) issue_patch();
                        mbar_wait(&bempty[st], ph ^ 1);
                        if (args.dbg & 8) { mbar_arrive(&bfull[st]); continue; }
                        mbar_arrive_expect_tx(&bfull[st], kBStage);
                        const int kcol = (tap * CB + cb) * TC_BK;
                        tma_load_2d(&tmBhi, &bfull[st], bstage(st), kcol, n0);
                        tma_load_2d(&tmBlo, &bfull[st], bstage(st) + (size_t)BN * 128, kcol, n0);
                    }
                }
                TC_TRACE(1, (tile - (int)blockIdx.x) / (int)gridDim.x);
            }
        }
    } else if (!DUAL && warp == 1) {
        // ===== the MMA issuer: [main | cross] += a x [b_hi | b_lo], then cross += a_lo x b_hi =================
        if (elect_one()) {
            constexpr uint32_t idesc2 = umma_idesc_tf32(2 * BN), idesc1 = umma_idesc_tf32(BN);
            int pi = 0, bi = 0, t = 0, st = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
                const int acc = t % kAcc;
                mbar_wait(&tempty[acc], ((uint32_t)(t / kAcc) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)acc * kAccCols;
                TC_TRACE(2, t);
                for (int cb = 0; cb < CB; ++cb, ++pi) {
                    const int ps = pi & 1;
                    const uint32_t pph = (uint32_t)(pi >> 1) & 1;
                    mbar_wait(&pfull[ps], pph);
                    tc_fence_after();
                    if (cb == 0) TC_TRACE(3, t);
                    const uint32_t araw = smem_u32(patch_raw(ps)), alo = smem_u32(patch_lo(ps));
                    for (int tap = 0; tap < 9; ++tap, ++bi, ring_next(st, ph, bstages)) {
                        mbar_wait(&bfull[st], ph);
                        tc_fence_after();
                        const int r = tap / 3, sx = tap - 3 * r;
                        const uint32_t shift = (uint32_t)(r * PW + sx) * 128u;                          // the tap = a row shift
                        const uint64_t da = umma_desc_sw128(araw + shift), dl = umma_desc_sw128(alo + shift);
                        const uint64_t db = umma_desc_sw128(smem_u32(bstage(st)));
                        if (!(args.dbg & 1)) {
#pragma unroll
                            for (int kk = 0; kk < TC_BK / 8; ++kk)
                                umma_tf32(d, da + 2 * kk, db + 2 * kk, idesc2, (cb | tap | kk) ? 1u : 0u);
                        }
                        if (tap == 0) { mbar_wait(&psplit[ps], pph); tc_fence_after(); }
                        if (!(args.dbg & 2)) {
#pragma unroll
                            for (int kk = 0; kk < TC_BK / 8; ++kk)
                                umma_tf32(d + BN, dl + 2 * kk, db + 2 * kk, idesc1, 1u);
                        }
                        if (args.dbg & 16) mbar_arrive(&bempty[st]); else umma_commit(&bempty[st]);
                    }
                    umma_commit(&pempty[ps]);
                }
                TC_TRACE(4, t);
                umma_commit(&tfull[acc]);
            }
        }
    } else if (DUAL && (warp == 1 || warp == 2)) {
        // ===== MMA issuers: warp 1  [main | cross] += a x [b_hi | b_lo];  warp 2  cross2 += a_lo x b_hi ============
        if (elect_one()) {
            const bool lo_issuer = warp == 2;
            const uint32_t idesc = lo_issuer ? umma_idesc_tf32(BN) : umma_idesc_tf32(2 * BN);
            int pi = 0, bi = 0, t = 0, st = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
                const int acc = t % kAcc;
                mbar_wait(&tempty[acc], ((uint32_t)(t / kAcc) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)acc * kAccCols + (lo_issuer ? 2u * BN : 0u);
                if (!lo_issuer) TC_TRACE(2, t);
                for (int cb = 0; cb < CB; ++cb, ++pi) {
                    const int ps = pi & 1;
                    const uint32_t pph = (uint32_t)(pi >> 1) & 1;
                    mbar_wait(&pfull[ps], pph);
                    if (lo_issuer) mbar_wait(&psplit[ps], pph);
                    tc_fence_after();
                    if (!lo_issuer && cb == 0) TC_TRACE(3, t);
                    const uint32_t abase = smem_u32(lo_issuer ? patch_lo(ps) : patch_raw(ps));
                    for (int tap = 0; tap < 9; ++tap, ++bi, ring_next(st, ph, bstages)) {
                        mbar_wait(&bfull[st], ph);
                        tc_fence_after();
                        const int r = tap / 3, sx = tap - 3 * r;
                        const uint64_t da = umma_desc_sw128(abase + (uint32_t)(r * PW + sx) * 128u);    // the tap = a row shift
                        const uint64_t db = umma_desc_sw128(smem_u32(bstage(st)));
                        if (!(args.dbg & (lo_issuer ? 2 : 1))) {
#pragma unroll
                            for (int kk = 0; kk < TC_BK / 8; ++kk)
                                umma_tf32(d, da + 2 * kk, db + 2 * kk, idesc, (cb | tap | kk) ? 1u : 0u);
                        }
                        umma_commit(&bempty[st]);
                    }
                    umma_commit(&pempty[ps]);
                }
                if (!lo_issuer) TC_TRACE(4, t);
                umma_commit(&tfull[acc]);
            }
        }
    } else if (warp == 3) {
        // ===== epilogue TMA issuer of both groups ===========================================================
        if (elect_one()) {
            const uint32_t my_tiles = (uint32_t)((num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x);
            const uint32_t total = my_tiles * SUBS;
            if (total > 0) { mbar_arrive(&slot_ready[0]); mbar_arrive(&slot_ready[1]); }
            uint32_t kdone[2] = {0, 0};
            const long long t0 = clock64();
            while (kdone[0] < total || kdone[1] < total) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const uint32_t k = kdone[g];
                    if (k >= total) continue;
                    if (!((args.dbg & 32) ? mbar_try_wait(&out_ready[g], k & 1) : mbar_test_wait(&out_ready[g], k & 1))) continue;
                    const int tile = (int)blockIdx.x + (int)(k / SUBS) * (int)gridDim.x;
                    int img, p0, n0; bool tail;
                    tile_of(tile, img, p0, n0, tail);
                    const int col = n0 + (g + 2 * (int)(k % SUBS)) * 32;
                    const int row = (img * args.H + p0) * args.W;
                    tma_store_2d(tail ? &tmOutTail : &tmOut, staging + (size_t)g * EPI_SLOT_BYTES, col, row);
                    bulk_commit();
                    if (k + 1 < total) { bulk_wait_read0(); mbar_arrive(&slot_ready[g]); }
                    kdone[g] = k + 1;
                }
                if (clock64() - t0 > 40000000000LL) { printf("i2v conv3x3_halo: epilogue issuer timeout (block %d)\n", blockIdx.x); __trap(); }
            }
            bulk_wait_all();
        }
    } else if (warp >= 4 && warp < 8) {
        // ===== split: a_lo patch = a - tf32(a), element-wise at the same (swizzled) offsets, once per patch ============
        const int t128 = threadIdx.x - 128;
        const int n4 = args.rows_load * 8;
        int pi = 0, t = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t)
            for (int cb = 0; cb < CB; ++cb, ++pi) {
                const int ps = pi & 1;
                mbar_wait(&pfull[ps], (uint32_t)(pi >> 1) & 1);
                const float4* src = reinterpret_cast<const float4*>(patch_raw(ps));
                float4* dst = reinterpret_cast<float4*>(patch_lo(ps));
#pragma unroll 4
                for (int i = (args.dbg & 4) ? n4 : t128; i < n4; i += 128) {
                    const float4 v = src[i];
                    float4 o;
                    o.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
                    o.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
                    o.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
                    o.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
                    dst[i] = o;
                }
                fence_proxy_async();
                mbar_arrive(&psplit[ps]);
                if (t128 == 0 && cb == CB - 1) TC_TRACE(7, t);
            }
    } else if (warp >= 8) {
        // ===== epilogue: thread = padded-raster position of the tile =========================================
        const int g = (warp - 8) >> 2;
        const int j = (warp & 3) * 32 + lane;
        const int rr = j / PW, q = j - rr * PW;
        const float* __restrict__ gbias = args.bias;
        const uint32_t* __restrict__ mbits = args.mask_bits;
        uint32_t* __restrict__ obits = args.bits_out;
        const int jd = rr * args.W + q;                           // dense row of the staging slot
        uint8_t* srow = staging + (size_t)g * EPI_SLOT_BYTES + (size_t)jd * 128;
        const uint32_t swz = (uint32_t)(jd & 7);
        uint32_t k = 0;
        int t = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
            int img, p0, n0; bool tail;
            tile_of(tile, img, p0, n0, tail);
            const int acc = t % kAcc;
            const int p = p0 + rr;
            const bool valid = rr < args.R && q < args.W && p < args.H;
            const int64_t m = ((int64_t)img * args.H + p) * args.W + q;
            uint32_t mw[SUBS];
#pragma unroll
            for (int u = 0; u < SUBS; ++u)
                mw[u] = (mbits && valid) ? __ldg(mbits + (int64_t)(n0 / 32 + g + 2 * u) * args.M + m) : 0xFFFFFFFFu;
            mbar_wait(&tfull[acc], (uint32_t)(t / kAcc) & 1);
            tc_fence_after();
            if (threadIdx.x == 256) TC_TRACE(5, t);
            const uint32_t tacc = tmem_base + (uint32_t)acc * kAccCols + ((uint32_t)((warp & 3) * 32) << 16);
            uint32_t vals[SUBS][32];
#pragma unroll
            for (int u = 0; u < SUBS; ++u) {
                const int c0 = (g + 2 * u) * 32;
                uint32_t b[32];
                tmem_ld32_nowait(tacc + (uint32_t)c0, vals[u]);
                tmem_ld32_nowait(tacc + (uint32_t)(BN + c0), b);
                if (DUAL) {
                    uint32_t c2[32];
                    tmem_ld32_nowait(tacc + (uint32_t)(2 * BN + c0), c2);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) {               // cross + cross2 first (both ~2^-11 of main), then + main
                        const float x = __fadd_rn(__uint_as_float(b[i]), __uint_as_float(c2[i]));
                        vals[u][i] = __float_as_uint(__fadd_rn(__uint_as_float(vals[u][i]), x));
                    }
                } else {
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        vals[u][i] = __float_as_uint(__fadd_rn(__uint_as_float(vals[u][i]), __uint_as_float(b[i])));
                }
            }
            tc_fence_before();
            mbar_arrive(&tempty[acc]);                           // accumulator drained: the MMA warps may reuse it
#pragma unroll
            for (int u = 0; u < SUBS; ++u, ++k) {
                const int c0 = (g + 2 * u) * 32;
                uint32_t (&a)[32] = vals[u];
                mbar_wait(&slot_ready[g], k & 1);                // the previous store has read the slot
                if (valid) {
                    const float4* bs4 = reinterpret_cast<const float4*>(gbias + n0 + c0);
                    const uint32_t mword = mw[u];
                    uint32_t oword = 0;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 bv = gbias ? __ldg(bs4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                        float4 v = make_float4(__uint_as_float(a[4 * c]) + bv.x, __uint_as_float(a[4 * c + 1]) + bv.y,
                                               __uint_as_float(a[4 * c + 2]) + bv.z, __uint_as_float(a[4 * c + 3]) + bv.w);
                        if (args.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                        if (!((mword >> (4 * c)) & 1u)) v.x = 0.f;
                        if (!((mword >> (4 * c + 1)) & 1u)) v.y = 0.f;
                        if (!((mword >> (4 * c + 2)) & 1u)) v.z = 0.f;
                        if (!((mword >> (4 * c + 3)) & 1u)) v.w = 0.f;
                        oword |= (v.x > 0.f ? 1u : 0u) << (4 * c) | (v.y > 0.f ? 1u : 0u) << (4 * c + 1) |
                                 (v.z > 0.f ? 1u : 0u) << (4 * c + 2) | (v.w > 0.f ? 1u : 0u) << (4 * c + 3);
                        *reinterpret_cast<float4*>(srow + (((uint32_t)c ^ swz) << 4)) = v;
                    }
                    if (obits) obits[(int64_t)((n0 + c0) / 32) * args.M + m] = oword;
                }
                fence_proxy_async();                             // generic-proxy writes -> visible to the TMA store
                mbar_arrive(&out_ready[g]);
            }
            if (threadIdx.x == 256) TC_TRACE(6, t);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// First-layer FORWARD without the patch matrix (i2v_conv_stem_fwd_rows_f32; 7x7 / stride 2 / pad 3, 64 output channels,
// Q <= 128, W % 4 == 0).  The im2col + GEMM pair writes a 2 GB patch matrix and reads it back (1.1 ms per 256 frames for
// 0.98 GB of tensors).  Here a tile is ONE output row (n, p): TMA stages the 7 input rows x 3 channels it needs (zero filled
// beyond the image: coordinates start at -3), four "assemble" warps (thread = output pixel q) build the 128 x 160 patch tile
// in shared memory k-block by k-block — A_hi in the swizzled layout the tensor core reads, A_lo into a tensor-memory ring —
// and the weights [64, 160] (hi | lo) stay resident.  K = 147 taps (c,r,s) + 13 zeros = 5 k-blocks.  [main | cross] +=
// a_hi [b_hi | b_lo] (N = 128) and cross += a_lo b_hi (N = 64) from one issuer; two accumulator stages; bias + ReLU in the
// epilogue, output through two swizzled staging slots and TMA stores whose box is exactly the Q rows of the tile.
// HBM traffic = the image once + the output once.
// ---------------------------------------------------------------------------------------------
constexpr int SF_THREADS = 512;          // warp 0 producer, 1 MMA issuer, 2 TMEM allocator, 4-7 and 12-15 assemble (alternate k-blocks), 8-11 epilogue
constexpr int SF_KB = 5;                 // k-blocks of 32 taps
constexpr int SF_ASLOTS = 4;             // ring of 16 KB A k-block slots (and of 32-column A_lo slots in tensor memory)
constexpr int SF_ISTAGES = 2;            // staged input rows
constexpr uint32_t SF_B_KB = 2 * 64 * TC_BK * 4;      // 16 KB per k-block: [64 x b_hi | 64 x b_lo]

struct StemFwdArgs {
    const float* bias;
    int N, H, W, P, Q, relu;
    int pitch;                           // staged row pitch in floats = TMA box width (>= W + 7: columns -4 .. W+2, multiple of 4)
    int cstride;                         // floats between the staged channels (7 * pitch rounded up to 128 bytes: TMA destination alignment)
    int tiles;                           // N * P
    // fused 3x3 / stride-2 / pad-1 max pooling (stem_fwd_rows_kernel<true>): the CTA walks UNITS = (image, strip of S pooled rows),
    // conv rows 2 i0 - 1 .. 2 i1 - 1 of a strip in order, and the stem activation never leaves the chip
    float* pooled;                       // [N, P2, Q2, 64]
    uint8_t* argmax;                     // [N, P2, Q2, 64] window position r * 3 + s of the winner, 255 = no winner (mark_dead)
    int P2, Q2, S, U, units, mark_dead;
};

// the rows (image n, conv row p) a CTA computes, in order; t counts them
struct SfIter { int t, n, p, i0, i1, u, pend; };
__device__ __forceinline__ void sf_unit(const StemFwdArgs& a, SfIter& it) {
    it.n = it.u / a.U;
    const int k = it.u - it.n * a.U;
    it.i0 = k * a.S;
    it.i1 = it.i0 + a.S < a.P2 ? it.i0 + a.S : a.P2;
    it.p = it.i0 > 0 ? 2 * it.i0 - 1 : 0;          // the lead-in row 2 i0 - 1 is row r = 0 of pooled row i0
    it.pend = 2 * it.i1 - 1;
}
template <bool POOL>
__device__ __forceinline__ bool sf_begin(const StemFwdArgs& a, SfIter& it) {
    it.t = 0; it.u = blockIdx.x; it.i0 = it.i1 = 0; it.pend = 0;
    if (!POOL) { if (it.u >= a.tiles) return false; it.n = it.u / a.P; it.p = it.u - it.n * a.P; return true; }
    if (it.u >= a.units) return false;
    sf_unit(a, it);
    return true;
}
template <bool POOL>
__device__ __forceinline__ bool sf_next(const StemFwdArgs& a, SfIter& it) {
    ++it.t;
    if (!POOL) { it.u += gridDim.x; if (it.u >= a.tiles) return false; it.n = it.u / a.P; it.p = it.u - it.n * a.P; return true; }
    if (it.p < it.pend) { ++it.p; return true; }
    it.u += gridDim.x;
    if (it.u >= a.units) return false;
    sf_unit(a, it);
    return true;
}

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

template <bool POOL>
__global__ void __launch_bounds__(SF_THREADS, 1)
stem_fwd_rows_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmBhi,
                     const __grid_constant__ CUtensorMap tmBlo, const __grid_constant__ CUtensorMap tmOut, const StemFwdArgs args) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* bt = smem;                                         // [kb]: b_hi rows then b_lo rows, 16 KB each
    uint8_t* aslots = smem + SF_KB * SF_B_KB;                   // 80 KB in: 1024-aligned
    uint8_t* staging = aslots + (size_t)SF_ASLOTS * TC_A_BYTES; // two 16 KB output slots (32 channels each)
    const uint32_t in_bytes = (uint32_t)(3 * 7 * args.pitch * 4);
    const uint32_t in_stride = (uint32_t)(3 * args.cstride * 4);
    float* inrows = reinterpret_cast<float*>(staging + 2 * EPI_SLOT_BYTES);   // [stage][c][r][pitch]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(inrows) + (size_t)SF_ISTAGES * in_stride);
    uint64_t* bfull = bars;
    uint64_t* ifull = bfull + 1;             // input rows landed
    uint64_t* iempty = ifull + SF_ISTAGES;   // the 4 assemble warps are done with them
    uint64_t* afull = iempty + SF_ISTAGES;   // A k-block assembled (4 warps)
    uint64_t* aempty = afull + SF_ASLOTS;    // ... consumed by the MMAs
    uint64_t* tfull = aempty + SF_ASLOTS;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmX); prefetch_tmap(&tmBhi); prefetch_tmap(&tmBlo); prefetch_tmap(&tmOut);
        mbar_init(bfull, 1);
        for (int s = 0; s < SF_ISTAGES; ++s) { mbar_init(&ifull[s], 1); mbar_init(&iempty[s], 8); }
        for (int s = 0; s < SF_ASLOTS; ++s) { mbar_init(&afull[s], 4); mbar_init(&aempty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
        fence_barrier_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;   // accumulator stage a: [main 64 | cross 64] at 128 a; A_lo ring at 256 + 32 slot
    constexpr uint32_t kRing = 256;

    if (warp == 0) {
        if (elect_one()) {
            mbar_arrive_expect_tx(bfull, SF_KB * SF_B_KB);
            for (int kb = 0; kb < SF_KB; ++kb) {
                tma_load_2d(&tmBhi, bfull, bt + (size_t)kb * SF_B_KB, kb * TC_BK, 0);
                tma_load_2d(&tmBlo, bfull, bt + (size_t)kb * SF_B_KB + 64 * TC_BK * 4, kb * TC_BK, 0);
            }
            SfIter ri;
            for (bool more = sf_begin<POOL>(args, ri); more; more = sf_next<POOL>(args, ri)) {
                const int n = ri.n, p = ri.p, t = ri.t;
                const int st = t % SF_ISTAGES;
                const uint32_t ph = (uint32_t)(t / SF_ISTAGES) & 1;
                mbar_wait(&iempty[st], ph ^ 1);
                mbar_arrive_expect_tx(&ifull[st], in_bytes);
                uint8_t* dst = reinterpret_cast<uint8_t*>(inrows) + (size_t)st * in_stride;
                for (int c = 0; c < 3; ++c)
                    // the innermost start coordinate must be 16-byte aligned (-3 floats is an illegal instruction, measured):
                    // the box starts at column -4, so image column w sits at staged column w + 4
                    tma_load_3d(&tmX, &ifull[st], dst + (size_t)c * args.cstride * 4, -4, 2 * p - 3, n * 3 + c);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_tf32(64), idesc2 = umma_idesc_tf32(128);
            mbar_wait(bfull, 0);
            int it = 0;
            SfIter ri;
            for (bool more = sf_begin<POOL>(args, ri); more; more = sf_next<POOL>(args, ri)) {
                const int t = ri.t;
                const int acc = t & 1;
                mbar_wait(&tempty[acc], ((uint32_t)(t >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d0 = tmem_base + (uint32_t)acc * 128u;
                for (int kb = 0; kb < SF_KB; ++kb, ++it) {
                    const int sl = it % SF_ASLOTS;
                    const uint32_t ph = (uint32_t)(it / SF_ASLOTS) & 1;
                    mbar_wait(&afull[sl], ph);
                    tc_fence_after();
                    const uint64_t da = umma_desc_sw128(smem_u32(aslots + (size_t)sl * TC_A_BYTES));
                    const uint64_t db = umma_desc_sw128(smem_u32(bt + (size_t)kb * SF_B_KB));
                    const uint32_t talo = tmem_base + kRing + 32u * (uint32_t)sl;
#pragma unroll
                    for (int kk = 0; kk < TC_BK / 8; ++kk) {
                        umma_tf32(d0, da + 2 * kk, db + 2 * kk, idesc2, (kb | kk) ? 1u : 0u);          // [main | cross] += a_hi [b_hi | b_lo]
                        umma_tf32_ts(d0 + 64u, talo + 8u * kk, db + 2 * kk, idesc, 1u);               // cross += a_lo b_hi
                    }
                    umma_commit(&aempty[sl]);
                }
                umma_commit(&tfull[acc]);
            }
        }
    } else if ((warp >= 4 && warp < 8) || warp >= 12) {
        // ===== assemble: thread = output pixel q; patch value k = (c,r,s) is staged[c][r][2q + s] ===========================
        // Two groups of four warps (one warp per TMEM lane quarter each) take alternate k-blocks: an assemble warp's k-block is a
        // chain of dependent latencies (32 LDS, the split, 8 STS.128, tcgen05.st + wait, two fences, the arrive), and with one
        // group that chain, not the tensor pipe, set the 2.8 us per output row.
        const int group = warp >= 12 ? 1 : 0;
        const int row = (warp & 3) * 32 + lane;
        const int qq = row < args.Q ? row : args.Q - 1;           // junk rows repeat the last pixel (never stored)
        const uint32_t swz = (uint32_t)(row & 7);
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kRing;
        int it = 0;
        SfIter ri;
        for (bool more = sf_begin<POOL>(args, ri); more; more = sf_next<POOL>(args, ri)) {
            const int t = ri.t;
            const int st = t % SF_ISTAGES;
            const uint32_t iph = (uint32_t)(t / SF_ISTAGES) & 1;
            mbar_wait(&ifull[st], iph);
            const float* in = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(inrows) + (size_t)st * in_stride) + 2 * qq + 1;   // column 2q - 3 + s -> staged 2q + s + 1
#pragma unroll
            for (int kb = 0; kb < SF_KB; ++kb, ++it) {
                if ((it & 1) != group) continue;
                const int sl = it % SF_ASLOTS;
                const uint32_t ph = (uint32_t)(it / SF_ASLOTS) & 1;
                mbar_wait(&aempty[sl], ph ^ 1);
                uint8_t* arow = aslots + (size_t)sl * TC_A_BYTES + (size_t)row * 128;
                uint32_t lo[32];
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    float v[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int k = kb * 32 + c4 * 4 + j;        // compile-time after unrolling
                        if (k < 147) {
                            const int c = k / 49, rs = k - c * 49, r = rs / 7, s = rs - r * 7;
                            v[j] = in[c * args.cstride + r * args.pitch + s];
                        } else v[j] = 0.f;
                        lo[c4 * 4 + j] = __float_as_uint(v[j] - __uint_as_float(__float_as_uint(v[j]) & 0xFFFFE000u));
                    }
                    *reinterpret_cast<float4*>(arow + (((uint32_t)c4 ^ swz) << 4)) = make_float4(v[0], v[1], v[2], v[3]);
                }
                tmem_st32(lane_addr + 32u * (uint32_t)sl, lo);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                fence_proxy_async();                               // generic-proxy writes of A_hi -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(&afull[sl]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&iempty[st]);
        }
    } else if (warp >= 8) {
        // ===== epilogue: thread = output pixel; + bias, ReLU, two swizzled 32-channel slots, TMA store of the Q valid rows =====
        const int ew = warp - 8;
        const int row = ew * 32 + lane;
        const uint32_t swz = (uint32_t)(row & 7);
        const uint32_t tlane = tmem_base + ((uint32_t)(ew * 32) << 16);
        const float* __restrict__ gbias = args.bias;
        // fused pooling: thread = (pooled column pj, channel half ph_) keeps the running window maximum of 32 channels in registers
        const int et = threadIdx.x - 256;
        const int pj = et >> 1, ph_ = et & 1;
        float pacc[32];
        uint32_t pidx[8];                                           // window positions, one byte per channel
#pragma unroll
        for (int c = 0; c < 32; ++c) pacc[c] = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) pidx[c] = 0u;
        SfIter ri;
        for (bool more = sf_begin<POOL>(args, ri); more; more = sf_next<POOL>(args, ri)) {
            const int t = ri.t;
            const int acc = t & 1;
            mbar_wait(&tfull[acc], (uint32_t)(t >> 1) & 1);
            tc_fence_after();
            const uint32_t tacc = tlane + (uint32_t)acc * 128u;
            float o[64];
#pragma unroll
            for (int off = 0; off < 64; off += 16) {
                uint32_t m0[16], c0[16];
                tmem_ld16_nowait(tacc + (uint32_t)off, m0);
                tmem_ld16_nowait(tacc + 64u + (uint32_t)off, c0);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) o[off + i] = __fadd_rn(__uint_as_float(m0[i]), __uint_as_float(c0[i]));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            asm volatile("bar.sync 1, 128;" ::: "memory");          // the previous tile's stores have read the slots (thread 256 waited)
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint8_t* srow = staging + (size_t)half * EPI_SLOT_BYTES + (size_t)row * 128;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 bv = gbias ? __ldg(reinterpret_cast<const float4*>(gbias + half * 32) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                    float4 v = make_float4(o[half * 32 + 4 * c] + bv.x, o[half * 32 + 4 * c + 1] + bv.y,
                                           o[half * 32 + 4 * c + 2] + bv.z, o[half * 32 + 4 * c + 3] + bv.w);
                    if (args.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                    *reinterpret_cast<float4*>(srow + (((uint32_t)c ^ swz) << 4)) = v;
                }
            }
            if (!POOL) fence_proxy_async();
            asm volatile("bar.sync 2, 128;" ::: "memory");
            if (!POOL) {
                if (threadIdx.x == 256) {
                    const int m0 = (ri.n * args.P + ri.p) * args.Q;    // rows (n, p, 0..Q-1) of y [N*P*Q, 64]
                    tma_store_2d(&tmOut, staging, 0, m0);
                    tma_store_2d(&tmOut, staging + EPI_SLOT_BYTES, 32, m0);
                    bulk_commit();
                    bulk_wait_read0();
                }
            } else if (pj < args.Q2) {
                // ---- max pooling of the row that now sits in the staging slots, separably and in i2v_maxpool_fwd_f32's window order
                // (first maximum in (r, s) order wins: strict > for every later candidate; columns / rows outside the image are
                // skipped).  Horizontal: pixels 2 pj - 1, 2 pj, 2 pj + 1 of this row.
                const int p = ri.p;
                const uint8_t* slot = staging + (size_t)ph_ * EPI_SLOT_BYTES;
                const int x0 = 2 * pj - 1, x1 = 2 * pj, x2 = 2 * pj + 1;
                const bool v0 = x0 >= 0, v2 = x2 < args.Q;
                const uint8_t* r0 = slot + (size_t)(v0 ? x0 : x1) * 128;
                const uint8_t* r1 = slot + (size_t)x1 * 128;
                const uint8_t* r2 = slot + (size_t)(v2 ? x2 : x1) * 128;
                const uint32_t z0 = (uint32_t)((v0 ? x0 : x1) & 7), z1 = (uint32_t)(x1 & 7), z2 = (uint32_t)((v2 ? x2 : x1) & 7);
                // Vertical: an even row p = 2i is row r = 1 of pooled row i; an odd row p = 2i + 1 is row r = 2 of pooled row i (which
                // it completes) and row r = 0 of pooled row i + 1.  Horizontal and vertical steps are fused per group of four
                // channels, so only the running maxima (32 registers) and their packed positions (8) live across rows.
                const bool odd = (p & 1) != 0;
                const bool first = p == 0;                           // row -1 does not exist: row 0 opens the window
                const int i = (p - 1) >> 1;
                const bool emit = odd && i >= ri.i0;                 // (a strip's lead-in row completes a pooled row of another unit)
                const uint32_t rbase = odd ? 6u : 3u;
                const int64_t o0 = ((((int64_t)ri.n * args.P2 + (emit ? i : 0)) * args.Q2 + pj) * 64 + ph_ * 32);
                float4* py = reinterpret_cast<float4*>(args.pooled + o0);
                uint32_t ow[8];
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    const float4 a0 = *reinterpret_cast<const float4*>(r0 + (((uint32_t)c4 ^ z0) << 4));
                    const float4 a1 = *reinterpret_cast<const float4*>(r1 + (((uint32_t)c4 ^ z1) << 4));
                    const float4 a2 = *reinterpret_cast<const float4*>(r2 + (((uint32_t)c4 ^ z2) << 4));
                    const float w0[4] = {a0.x, a0.y, a0.z, a0.w}, w1[4] = {a1.x, a1.y, a1.z, a1.w}, w2[4] = {a2.x, a2.y, a2.z, a2.w};
                    const uint32_t oldw = pidx[c4];
                    uint32_t keepw = 0u, outw = 0u;
                    float o4[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int c = 4 * c4 + k;
                        // horizontal, first maximum in s order: s = 0 keeps ties against s = 1, s = 2 needs strictly more
                        float b = w1[k];
                        uint32_t sidx = 1u;
                        const bool t0 = v0 && !(b > w0[k]);
                        b = t0 ? w0[k] : b; sidx = t0 ? 0u : sidx;
                        const bool t2 = v2 && (w2[k] > b);
                        b = t2 ? w2[k] : b; sidx = t2 ? 2u : sidx;
                        const uint32_t old = (oldw >> (8 * k)) & 0xffu;
                        // vertical against the running maximum of rows r < this one
                        const bool take = first || b > pacc[c];
                        const float best = (take && !(odd && !emit)) ? b : pacc[c];
                        uint32_t bi = take ? rbase + sidx : old;
                        if (args.mark_dead && !(best > 0.f)) bi = 255u;
                        o4[k] = best;
                        outw |= bi << (8 * k);
                        // state for the next row: an odd row restarts the window with itself as r = 0
                        pacc[c] = odd ? b : (take ? b : pacc[c]);
                        keepw |= (odd ? sidx : (take ? 3u + sidx : old)) << (8 * k);
                    }
                    pidx[c4] = keepw;
                    ow[c4] = outw;
                    if (emit) py[c4] = make_float4(o4[0], o4[1], o4[2], o4[3]);
                }
                if (emit) {
                    uint4* pa = reinterpret_cast<uint4*>(args.argmax + o0);
                    pa[0] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                    pa[1] = make_uint4(ow[4], ow[5], ow[6], ow[7]);
                }
            }
        }
        if (!POOL && threadIdx.x == 256) bulk_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// host: tensor maps (driver entry points resolved at run time: no link-time libcuda dependency)
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode_tiled = nullptr;
static EncodeIm2colFn g_encode_im2col = nullptr;

static int resolve_driver() {
    if (g_encode_tiled && g_encode_im2col) return I2V_OK;
    cudaDriverEntryPointQueryResult q;
    void* f = nullptr;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
    if (e != cudaSuccess || !f) return cuda_fail(e == cudaSuccess ? cudaErrorUnknown : e, "cuTensorMapEncodeTiled entry point");
    g_encode_tiled = (EncodeTiledFn)f;
    f = nullptr;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f, cudaEnableDefault, &q);
    if (e != cudaSuccess || !f) return cuda_fail(e == cudaSuccess ? cudaErrorUnknown : e, "cuTensorMapEncodeIm2col entry point");
    g_encode_im2col = (EncodeIm2colFn)f;
    return I2V_OK;
}

// 2-D row-major [rows, cols] f32, box [box_rows, 32], SWIZZLE_128B
static int make_map_2d(CUtensorMap* map, const float* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * sizeof(float)};
    cuuint32_t box[2] = {TC_BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu", (int)r, (unsigned long long)rows, (unsigned long long)cols); return I2V_ECUDA; }
    return I2V_OK;
}

// 2-D row-major [rows, cols] f32, box [box_rows, box_cols], no swizzle (transposed epilogue store: rows = channels,
// cols = GEMM rows)
static int make_map_2d_plain(CUtensorMap* map, const float* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols) {
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * sizeof(float)};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (plain) failed (%d) rows=%llu cols=%llu", (int)r, (unsigned long long)rows, (unsigned long long)cols); return I2V_ECUDA; }
    return I2V_OK;
}

// 2-D row-major [rows, 64] u8 (the pooling's argmax plane), box [box_rows, 64], SWIZZLE_64B: the 16-byte chunk of a 64-byte
// row is XORed with (row / 2) % 4, so 8 consecutive rows occupy 8 different bank groups
static int make_map_u8_rows64(CUtensorMap* map, const uint8_t* base, uint64_t rows, uint32_t box_rows) {
    cuuint64_t dims[2] = {64, rows};
    cuuint64_t strides[1] = {64};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(base), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (u8 rows) failed (%d) rows=%llu box=%u", (int)r, (unsigned long long)rows, box_rows); return I2V_ECUDA; }
    return I2V_OK;
}

// im2col-mode map over an NHWC activation tensor: 32 channels x 128 pixels per load
// Bounding box of the window corner in source coordinates: [lower, dim - 1 + upper] per axis (w, h).
// Forward conv: lower = -pad, upper = pad - (filter - 1) (cutlass/conv/collective/detail.hpp
// compute_{lower,upper}_corner_whd); a strided data-gradient class uses its own box (see tc_dgrad_class).
static int make_map_im2col(CUtensorMap* map, const float* base, int N, int H, int W, int C, int lower_w, int lower_h,
                           int upper_w, int upper_h, int stride) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    int lower[2] = {lower_w, lower_h};
    int upper[2] = {upper_w, upper_h};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = g_encode_im2col(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, lower, upper,
                                 TC_BK, TC_BM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeIm2col failed (%d) N=%d H=%d W=%d C=%d lower=(%d,%d) upper=(%d,%d) stride=%d", (int)r, N, H, W, C, lower_w, lower_h, upper_w, upper_h, stride); return I2V_ECUDA; }
    return I2V_OK;
}

// im2col-mode map over a VIEW of an NHWC tensor: Hv x Wv pixels starting at `base` (already shifted to the view's first
// pixel) inside images of Hfull x Wfull pixels; traversal stride `stride`, bounding box = the view.  The epilogue of a strided
// data-gradient class stores (and reads its addend) through it: tile row (img, i, j) <-> view pixel (i stride, j stride).
static int make_map_im2col_view(CUtensorMap* map, const float* base, int N, int Hv, int Wv, int C, int Hfull, int Wfull, int stride) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wv, (cuuint64_t)Hv, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)Wfull * C * 4, (cuuint64_t)Hfull * Wfull * C * 4};
    int lower[2] = {0, 0};
    int upper[2] = {0, 0};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = g_encode_im2col(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, lower, upper,
                                 TC_BK, TC_BM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeIm2col (view) failed (%d) N=%d Hv=%d Wv=%d C=%d H=%d W=%d stride=%d", (int)r, N, Hv, Wv, C, Hfull, Wfull, stride); return I2V_ECUDA; }
    return I2V_OK;
}

// 4-D tiled map with explicit byte strides (dims / strides innermost first; box = {32, 16, 8, 1} floats x pixels x rows x
// images, SWIZZLE_128B): the 16 x 8 pixel boxes of the direct first-layer forward.  Strides may OVERLAP the inner extent
// (the 8-pixel window of output pixel q+1 starts two pixels after that of q).
static int make_map_4d(CUtensorMap* map, const float* base, const uint64_t dims[4], const uint64_t strides_bytes[3]) {
    cuuint64_t d[4] = {dims[0], dims[1], dims[2], dims[3]};
    cuuint64_t st[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
    cuuint32_t box[4] = {TC_BK, 16, 8, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), d, st, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (4-D) failed (%d) dims=(%llu,%llu,%llu,%llu) strides=(%llu,%llu,%llu)", (int)r,
                  (unsigned long long)d[0], (unsigned long long)d[1], (unsigned long long)d[2], (unsigned long long)d[3],
                  (unsigned long long)st[0], (unsigned long long)st[1], (unsigned long long)st[2]);
        return I2V_ECUDA;
    }
    return I2V_OK;
}

// Tensor maps are keyed by (pointer, geometry): the engine reuses its activation buffers every step, so
// after the first step no descriptor is encoded on the hot path.
struct MapKey {
    const void* p; int a, b, c, d, e, f, g, h;
    bool operator==(const MapKey& o) const { return p == o.p && a == o.a && b == o.b && c == o.c && d == o.d && e == o.e && f == o.f && g == o.g && h == o.h; }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t x = reinterpret_cast<size_t>(k.p);
        for (int v : {k.a, k.b, k.c, k.d, k.e, k.f, k.g, k.h}) x = x * 1000003u ^ (size_t)(unsigned)v;
        return x;
    }
};
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;
static std::mutex g_maps_mu;

static int get_map_2d(CUtensorMap* out, const float* base, int rows, int cols, int box_rows) {
    MapKey key{base, rows, cols, box_rows, 0, 0, 0, 0, 1};
    std::lock_guard<std::mutex> lk(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return I2V_OK; }
    if (int r = make_map_2d(out, base, (uint64_t)rows, (uint64_t)cols, (uint32_t)box_rows)) return r;
    if (g_maps.size() > 4096) g_maps.clear();
    g_maps.emplace(key, *out);
    return I2V_OK;
}
static int get_map_2d_plain(CUtensorMap* out, const float* base, int rows, int cols, int box_rows, int box_cols) {
    MapKey key{base, rows, cols, box_rows, box_cols, 0, 0, 0, 3};
    std::lock_guard<std::mutex> lk(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return I2V_OK; }
    if (int r = make_map_2d_plain(out, base, (uint64_t)rows, (uint64_t)cols, (uint32_t)box_rows, (uint32_t)box_cols)) return r;
    if (g_maps.size() > 4096) g_maps.clear();
    g_maps.emplace(key, *out);
    return I2V_OK;
}
static int get_map_u8_rows64(CUtensorMap* out, const uint8_t* base, int rows, int box_rows) {
    MapKey key{base, rows, box_rows, 0, 0, 0, 0, 0, 6};
    std::lock_guard<std::mutex> lk(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return I2V_OK; }
    if (int r = make_map_u8_rows64(out, base, (uint64_t)rows, (uint32_t)box_rows)) return r;
    if (g_maps.size() > 4096) g_maps.clear();
    g_maps.emplace(key, *out);
    return I2V_OK;
}
static int get_map_im2col_view(CUtensorMap* out, const float* base, int N, int Hv, int Wv, int C, int Hfull, int Wfull, int stride) {
    MapKey key{base, N, Hv, Wv, C, Hfull, Wfull, stride, 7};
    std::lock_guard<std::mutex> lk(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return I2V_OK; }
    if (int r = make_map_im2col_view(out, base, N, Hv, Wv, C, Hfull, Wfull, stride)) return r;
    if (g_maps.size() > 4096) g_maps.clear();
    g_maps.emplace(key, *out);
    return I2V_OK;
}
// 4-D tiled map over an NHWC tensor, box {32 channels, box_w pixels, box_h rows, 1 image}, SWIZZLE_128B, zeros outside the
// image (negative start coordinates = the convolution's padding): the input patch of the halo kernel
static int get_map_nhwc_box(CUtensorMap* out, const float* base, int N, int H, int W, int C, int box_w, int box_h) {
    MapKey key{base, N, H, W, C, box_w, box_h, 0, 8};
    std::lock_guard<std::mutex> lk(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return I2V_OK; }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    cuuint32_t box[4] = {TC_BK, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode_tiled(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (NHWC box) failed (%d) N=%d H=%d W=%d C=%d box=%dx%d", (int)r, N, H, W, C, box_w, box_h); return I2V_ECUDA; }
    if (g_maps.size() > 4096) g_maps.clear();
    g_maps.emplace(key, *out);
    return I2V_OK;
}
static int get_map_4d(CUtensorMap* out, const float* base, const uint64_t dims[4], const uint64_t strides_bytes[3]) {
    MapKey key{base, (int)dims[1], (int)dims[2], (int)dims[3], (int)(strides_bytes[0]), (int)(strides_bytes[1] & 0x7fffffff),
               (int)(strides_bytes[2] & 0x7fffffff), (int)(((strides_bytes[2] >> 31) << 12) | (dims[0] & 0xfff)), 4};
    std::lock_guard<std::mutex> lk(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return I2V_OK; }
    if (int r = make_map_4d(out, base, dims, strides_bytes)) return r;
    if (g_maps.size() > 4096) g_maps.clear();
    g_maps.emplace(key, *out);
    return I2V_OK;
}
static int get_map_im2col(CUtensorMap* out, const float* base, int N, int H, int W, int C, int lower_w, int lower_h,
                          int upper_w, int upper_h, int stride) {
    MapKey key{base, N, H, W, C, (lower_w + 64) * 256 + (lower_h + 64), (upper_w + 64) * 256 + (upper_h + 64), stride, 2};
    std::lock_guard<std::mutex> lk(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return I2V_OK; }
    if (int r = make_map_im2col(out, base, N, H, W, C, lower_w, lower_h, upper_w, upper_h, stride)) return r;
    if (g_maps.size() > 4096) g_maps.clear();
    g_maps.emplace(key, *out);
    return I2V_OK;
}


// 3-D [planes, H, W] f32 view of an NCHW image batch, box [1, 7, box_w], no swizzle, zero fill outside (negative start
// coordinates are the convolution's padding)
static int make_map_planes(CUtensorMap* map, const float* base, uint64_t planes, uint64_t H, uint64_t W, uint32_t box_w, uint32_t box_h) {
    cuuint64_t dims[3] = {W, H, planes};
    cuuint64_t strides[2] = {W * sizeof(float), H * W * sizeof(float)};
    cuuint32_t box[3] = {box_w, box_h, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (planes) failed (%d) planes=%llu H=%llu W=%llu box=%u", (int)r, (unsigned long long)planes, (unsigned long long)H, (unsigned long long)W, box_w); return I2V_ECUDA; }
    return I2V_OK;
}
static int get_map_planes(CUtensorMap* out, const float* base, int planes, int H, int W, int box_w, int box_h) {
    MapKey key{base, planes, H, W, box_w, box_h, 0, 0, 5};
    std::lock_guard<std::mutex> lk(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return I2V_OK; }
    if (int r = make_map_planes(out, base, (uint64_t)planes, (uint64_t)H, (uint64_t)W, (uint32_t)box_w, (uint32_t)box_h)) return r;
    if (g_maps.size() > 4096) g_maps.clear();
    g_maps.emplace(key, *out);
    return I2V_OK;
}

template <int BN, bool X3, bool IM2COL>
static int tc_launch(const CUtensorMap& tmA, const CUtensorMap& tmBhi, const CUtensorMap& tmBlo, const TcArgs& args, cudaStream_t st) {
    using L = TcSmem<BN, X3>;
    auto kern = conv_tc_kernel<BN, X3, IM2COL>;
    // two CTAs per SM so that one tile's epilogue overlaps another tile's main loop
    static int stages = 0;
    static size_t smem = 0;
    if (stages == 0) {
        int dev = 0, optin = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        const size_t budget = (size_t)(optin > 0 ? optin : 227 * 1024) / 2 - 2048;
        const size_t fixed = 1024 /*align slack*/ + 512 /*barriers*/ + BN * 4;
        int s = (int)((budget - fixed) / L::STAGE_BYTES);
        if (s < 2) s = 2;        // fall back to one CTA per SM if two stages do not fit in half an SM
        if (s > 6) s = 6;
        smem = fixed + (size_t)s * L::STAGE_BYTES;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "conv_tc: shared memory attribute");
        stages = s;
    }
    dim3 grid((unsigned)((args.M + TC_BM - 1) / TC_BM), (unsigned)(args.Cout / BN));
    kern<<<grid, TC_THREADS, smem, st>>>(tmA, tmBhi, tmBlo, args, stages);
    I2V_LAUNCH_CHECK("i2v_conv_tc_f32");
    return I2V_OK;
}

template <int BN, bool X3, bool IM2COL, bool EPI_TMA, bool ALO_TMEM>
static int tc_launch_persist(const CUtensorMap& tmA, const CUtensorMap& tmBhi, const CUtensorMap& tmBlo, const CUtensorMap& tmOut,
                             const CUtensorMap& tmRes, const TcArgs& args, cudaStream_t st) {
    using L = TcSmem<BN, X3>;
    auto kern = conv_tc_persist_kernel<BN, X3, IM2COL, EPI_TMA, ALO_TMEM>;
    constexpr size_t kStageBytes = ALO_TMEM ? L::STAGE_BYTES - TC_A_BYTES : L::STAGE_BYTES;
    static size_t budget = 0;
    if (budget == 0) {
        int dev = 0, optin = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        const size_t b = (size_t)(optin > 0 ? optin : 227 * 1024);
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b);
        if (e != cudaSuccess) return cuda_fail(e, "conv_tc (persistent): shared memory attribute");
        budget = b;
    }
    I2V_REQUIRE(args.Cout <= 2048, "Cout > 2048 not supported by the persistent tensor-core kernel");
    // epilogue staging: two 16 KB slots per group when the tile is short (the epilogue is then on the critical path);
    // one slot per group for K-heavy layers, which frees room for one more pipeline stage ($I2V_TC_EPI_SLOTS=1|2 overrides)
    static const int slots_env = getenv("I2V_TC_EPI_SLOTS") ? atoi(getenv("I2V_TC_EPI_SLOTS")) : 0;
    const int kiters = args.taps_h * args.taps_w * args.cblocks;
    int slots = kiters < 8 ? 2 : 1;       // measured: with >= 8 k-steps per tile one slot wins even when a residual streams through
    if (slots_env == 1 || slots_env == 2) slots = slots_env;
    const size_t fixed = 1008 /*align slack*/ + 512 /*barriers*/ +
                         (EPI_TMA ? (size_t)slots * 2 * EPI_SLOT_BYTES : (size_t)2048 * 4 /*bias, Cout <= 2048*/);
    int stages = (int)((budget - fixed) / kStageBytes);
    if (stages > 8) stages = 8;
    if (ALO_TMEM && stages > 4) stages = 4;            // the A_lo ring in tensor memory has 4 slots
    if (stages < 2) { set_error("conv_tc (persistent): %d pipeline stages fit in shared memory", stages); return I2V_ECUDA; }
    I2V_REQUIRE(!EPI_TMA || args.bias == nullptr || (reinterpret_cast<uintptr_t>(args.bias) & 15) == 0, "bias must be 16-byte aligned");
    const int num_m_tiles = (int)((args.M + TC_BM - 1) / TC_BM), num_n_tiles = args.Cout / BN;
    const int64_t tiles = (int64_t)num_m_tiles * num_n_tiles;
    const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
    // Resident weights ($I2V_TC_BRES=0: off): when the n-tile's whole weight block fits beside >= 3 A stages and every tile of a
    // CTA has the same n-tile (grid % num_n_tiles == 0), it is loaded once per CTA instead of once per tile — the 1x1 layers with
    // K <= 128 (BN = 128) / K <= 256 (BN = 64) re-fetched 2-3x the bytes of their A tiles from L2 as weights.
    static const bool bres_env = !(getenv("I2V_TC_BRES") && atoi(getenv("I2V_TC_BRES")) == 0);
    TcArgs a2 = args;
    size_t smem_bytes = fixed + (size_t)stages * kStageBytes;
    // (measured in the attack step: 1-2 % on each of these layers — 56x56 256->64 +res 358 -> 353 us, 64->256 207 -> 203 us —
    // except the two-k-step BN = 64 launches, 100 -> 107 us, which keep the ring)
    if (bres_env && X3 && !IM2COL && grid % num_n_tiles == 0 && tiles >= 4 * (int64_t)grid && (BN == 128 || kiters >= 3)) {
        constexpr size_t kStageA = kStageBytes - 2 * (size_t)L::B_BYTES;
        const size_t bres = (size_t)kiters * 2 * L::B_BYTES;
        if (fixed + bres + 3 * kStageA <= budget) {
            int sa = (int)((budget - fixed - bres) / kStageA);
            if (sa > 8) sa = 8;
            if (ALO_TMEM && sa > 4) sa = 4;
            a2.b_resident = 1;
            stages = sa;
            smem_bytes = fixed + bres + (size_t)sa * kStageA;
        }
    }
    kern<<<grid, TC2_THREADS, smem_bytes, st>>>(tmA, tmBhi, tmBlo, tmOut, tmRes, a2, stages, num_m_tiles, num_n_tiles, slots);
    I2V_LAUNCH_CHECK("i2v_conv_tc_f32 (persistent)");
    return I2V_OK;
}


// CTA-pair launch (conv_tc_pair_kernel): clusters of 2, one cluster per SM pair; tmBhi / tmBlo are HALF-height boxes (BN/2 rows)
template <int BN, bool IM2COL>
static int tc_launch_pair(const CUtensorMap& tmA, const CUtensorMap& tmBhi, const CUtensorMap& tmBlo, const CUtensorMap& tmOut,
                          const CUtensorMap& tmRes, const TcArgs& args, cudaStream_t st) {
    auto kern = conv_tc_pair_kernel<BN, IM2COL>;
    constexpr size_t kStageBytes = TC_A_BYTES + (size_t)BN * TC_BK * 4;
    static size_t budget = 0;
    if (budget == 0) {
        int dev = 0, optin = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        const size_t b = (size_t)(optin > 0 ? optin : 227 * 1024);
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b);
        if (e != cudaSuccess) return cuda_fail(e, "conv_tc (pair): shared memory attribute");
        budget = b;
    }
    static const int slots_env = getenv("I2V_TC_EPI_SLOTS") ? atoi(getenv("I2V_TC_EPI_SLOTS")) : 0;
    int slots = 2;                          // a pair stage is 24-32 KB: four stages and two staging slots per group fit
    if (slots_env == 1 || slots_env == 2) slots = slots_env;
    const size_t fixed = 1008 + 1024 + (size_t)slots * 2 * EPI_SLOT_BYTES;
    int stages = (int)((budget - fixed) / kStageBytes);
    if (stages > 4) stages = 4;             // the A_lo ring in tensor memory has 4 slots
    static const int stages_env = getenv("I2V_TC_PAIR_STAGES") ? atoi(getenv("I2V_TC_PAIR_STAGES")) : 0;
    if (stages_env >= 2 && stages_env < stages) stages = stages_env;
    if (stages < 2) { set_error("conv_tc (pair): %d pipeline stages fit in shared memory", stages); return I2V_ECUDA; }
    I2V_REQUIRE(args.bias == nullptr || (reinterpret_cast<uintptr_t>(args.bias) & 15) == 0, "bias must be 16-byte aligned");
    const int num_m_tiles = (int)((args.M + TC_BM - 1) / TC_BM), num_n_tiles = args.Cout / BN;
    const int64_t ptiles = (int64_t)((num_m_tiles + 1) / 2) * num_n_tiles;
    const int pairs = (int)(ptiles < sm_count() / 2 ? ptiles : sm_count() / 2);
    kern<<<2 * pairs, TC2_THREADS, fixed + (size_t)stages * kStageBytes, st>>>(tmA, tmBhi, tmBlo, tmOut, tmRes, args, stages,
                                                                              num_m_tiles, num_n_tiles, slots);
    I2V_LAUNCH_CHECK("i2v_conv_tc_f32 (pair)");
    return I2V_OK;
}

static unsigned long long* g_trace = nullptr;
static int g_trace_tiles = 0;
// Halo kernel launch (conv3x3_halo_kernel).  Returns I2V_OK after launching, or -1 when the shape does not fit (the caller
// falls back to the im2col-mode kernel).
template <int BN, bool DUAL>
static int halo_launch(const float* src, int N, int H, int W, int C, const float* w_hi, const float* w_lo, int Cout,
                       const float* bias, const uint32_t* mask_bits, uint32_t* bits_out, float* dst, int relu, cudaStream_t st) {
    const int PW = W + 2;
    if (PW > TC_BM || C % 32 != 0 || C / 32 > 16 || Cout % BN != 0) return -1;
    const int R = TC_BM / PW;
    const int64_t M = (int64_t)N * H * W;
    HaloArgs a{};
    a.bias = bias; a.mask_bits = mask_bits; a.bits_out = bits_out;
    a.N = N; a.H = H; a.W = W; a.PW = PW; a.R = R; a.TI = (H + R - 1) / R; a.CB = C / 32; a.relu = relu;
    a.rows_load = (R + 2) * PW;
    a.patch_bytes = (uint32_t)(((2 * PW + 2 + TC_BM) * 128 + 1023) / 1024 * 1024);
    a.tail_rows = H % R;
    static const int halo_dbg = getenv("I2V_TC_HALO_DBG") ? atoi(getenv("I2V_TC_HALO_DBG")) : 0;
    a.dbg = halo_dbg;
    a.trace = g_trace; a.trace_tiles = g_trace_tiles;
    a.M = M;
    static size_t budget = 0;
    auto kern = conv3x3_halo_kernel<BN, DUAL>;
    if (budget == 0) {
        int dev = 0, optin = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        const size_t b = (size_t)(optin > 0 ? optin : 227 * 1024);
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b);
        if (e != cudaSuccess) return cuda_fail(e, "conv3x3 halo: shared memory attribute");
        budget = b;
    }
    const size_t bstage = (size_t)2 * BN * 128;
    const size_t fixed = 1024 + 4 * (size_t)a.patch_bytes + 2 * EPI_SLOT_BYTES + 512;
    if (fixed + 2 * bstage > budget) return -1;
    int bst = (int)((budget - fixed) / bstage);
    if (bst > 8) bst = 8;
    static const int bst_env = getenv("I2V_TC_HALO_BSTAGES") ? atoi(getenv("I2V_TC_HALO_BSTAGES")) : 0;     // ring-depth experiments
    if (bst_env >= 1 && bst_env < bst) bst = bst_env;
    a.bstages = bst;
    CUtensorMap tmX, tmBhi, tmBlo, tmOut, tmTail;
    if (int r = get_map_nhwc_box(&tmX, src, N, H, W, C, PW, R + 2)) return r;
    if (int r = get_map_2d(&tmBhi, w_hi, Cout, 9 * C, BN)) return r;
    if (int r = get_map_2d(&tmBlo, w_lo, Cout, 9 * C, BN)) return r;
    const int full_rows = (H < R ? H : R) * W;
    if (int r = get_map_2d(&tmOut, dst, (int)M, Cout, full_rows)) return r;
    tmTail = tmOut;
    if (a.tail_rows) { if (int r = get_map_2d(&tmTail, dst, (int)M, Cout, a.tail_rows * W)) return r; }
    const int num_n_tiles = Cout / BN;
    const int64_t tiles = (int64_t)N * a.TI * num_n_tiles;
    if (tiles >= (int64_t)0x7fffffff) return -1;
    const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
    kern<<<grid, HL_THREADS, fixed + (size_t)bst * bstage, st>>>(tmX, tmBhi, tmBlo, tmOut, tmTail, a, (int)tiles, num_n_tiles);
    I2V_LAUNCH_CHECK("i2v_conv_tc_f32 (3x3 halo)");
    return I2V_OK;
}

// CTA-pair kernel selection ($I2V_TC_PAIR, i2v_conv_tc_set_pair_minkit): n > 0 = every tile of >= n k-steps, 0 = never,
// -1 (default) = where it was measured to win in the attack step (profiles/r02_pair_kernel.md): convolutions that stream a
// residual / addend through the epilogue, from 4 k-steps per tile
static int g_pair_minkit = -2;
static int g_halo_mode = -2;        // $I2V_TC_HALO / i2v_conv_tc_set_halo_mode; see tc_run
static int pair_min_ksteps() {
    if (g_pair_minkit == -2) g_pair_minkit = getenv("I2V_TC_PAIR") ? atoi(getenv("I2V_TC_PAIR")) : -1;
    return g_pair_minkit;
}


// Common driver: `src` [N,H,W,C] is the gathered tensor, GEMM rows are the (img,p,q) grid, K = taps x C.
struct TcProblem {
    const float* src; int N, H, W, C;
    int P, Q, stride, lower_h, lower_w, upper_h, upper_w, taps_h, taps_w;
    const float* w_hi; const float* w_lo; int Cout;
    const float* bias; const float* residual; const float* mask_src; float* dst; int relu;
    int out_s, out_h0, out_w0, out_H, out_W;
    const uint32_t* mask_bits; uint32_t* bits_out;
    int out_transposed, store_cols;     // TMA epilogue: dst = [Cout][M]; only columns < store_cols are written (0 = all)
    int force_dual;                     // take the dual-issuer kernel whatever the tile length (first-layer dgrad GEMM)
    const float* src2; int C2;          // second A source [M, C2] (K = taps x C followed by C2); see TcArgs::a2_cb0
};

static int tc_run(const TcProblem& pr, cudaStream_t st) {
    if (int r = resolve_driver()) return r;
    const bool x3 = pr.w_lo != nullptr;
    const bool im2col = !(pr.taps_h == 1 && pr.taps_w == 1 && pr.stride == 1 && pr.lower_h == 0 && pr.lower_w == 0 &&
                          pr.P == pr.H && pr.Q == pr.W);
    I2V_REQUIRE(((reinterpret_cast<uintptr_t>(pr.src) | reinterpret_cast<uintptr_t>(pr.dst) | reinterpret_cast<uintptr_t>(pr.w_hi) |
                  reinterpret_cast<uintptr_t>(pr.w_lo) | reinterpret_cast<uintptr_t>(pr.residual) | reinterpret_cast<uintptr_t>(pr.mask_src) |
                  reinterpret_cast<uintptr_t>(pr.bias)) & 15) == 0, "all tensors must be 16-byte aligned");
    const int64_t M = (int64_t)pr.N * pr.P * pr.Q;
    I2V_REQUIRE(M < (int64_t)0x7fffffff, "too many output pixels for one launch");
    if (M == 0) return I2V_OK;
    // X3 needs 48 KB (BN=64) or 64 KB (BN=128) per stage: BN=64 keeps two stages in half an SM so that two
    // CTAs are co-resident; plain TF32 has room for BN=128.  I2V_TC_BN=64|128 overrides for experiments.
    static const bool persistent = !(getenv("I2V_TC_PERSISTENT") && atoi(getenv("I2V_TC_PERSISTENT")) == 0);
    int BN = (pr.Cout % 128 == 0 && (!x3 || persistent)) ? 128 : 64;
    // dual-source launches of <= $I2V_TC_DUAL_BN64 (default 4) k-steps: 64-channel tiles keep TWO accumulator stages with two
    // issuers (three accumulators of 128 columns leave one), so the epilogue overlaps the next tile's main loop
    // (layer1.0, 64 + 64 -> 256: 393 -> 352 us per 256 frames)
    static const int dual_bn64 = getenv("I2V_TC_DUAL_BN64") ? atoi(getenv("I2V_TC_DUAL_BN64")) : 4;
    if (pr.src2 && x3 && persistent && pr.taps_h * pr.taps_w * (pr.C / 32) + pr.C2 / 32 <= dual_bn64) BN = 64;
    if (const char* e = getenv("I2V_TC_BN")) { int v = atoi(e); if ((v == 64 || v == 128) && pr.Cout % v == 0) BN = v; }
    const int Ktot = pr.taps_h * pr.taps_w * pr.C + (pr.src2 ? pr.C2 : 0);

    // TMA epilogue (v3) whenever the output rows are dense and no f32 mask source is involved; the strided
    // data-gradient classes and callers that pass an f32 mask keep the register/LSU epilogue (v2)
    static const bool epi_tma_on = !(getenv("I2V_TC_EPI_TMA") && atoi(getenv("I2V_TC_EPI_TMA")) == 0);
    // ... since round 2 the strided classes too: their rows scatter through an im2col-mode TMA store ($I2V_TC_CLASS_TMA=0: v2)
    static const bool class_tma_on = !(getenv("I2V_TC_CLASS_TMA") && atoi(getenv("I2V_TC_CLASS_TMA")) == 0);
    const bool epi_tma = persistent && epi_tma_on && (pr.out_s == 0 || class_tma_on) && pr.mask_src == nullptr;
    I2V_REQUIRE(epi_tma || (pr.mask_bits == nullptr && pr.bits_out == nullptr),
                "bit masks need the TMA epilogue (dense output rows, no f32 mask source)");

    // 3x3 / stride 1 / pad 1 in FP32-parity mode: every input patch delivered once (conv3x3_halo_kernel).  $I2V_TC_HALO: 1 = wherever
    // it fits, 0 = never, default = the 64-channel tiles only, where it was measured to win in the attack step (56x56 64 -> 64:
    // 325 us against 382 us per 256 frames = 88 % of the measured TF32 GEMM rate counting issued MMAs; at BN = 128 its single
    // accumulator stage and two-stage weight ring lose, 255 us against 233 us)
    if (g_halo_mode == -2) g_halo_mode = getenv("I2V_TC_HALO") ? atoi(getenv("I2V_TC_HALO")) : -1;
    const int halo_mode = g_halo_mode;
    const bool halo_on = halo_mode > 0 || (halo_mode < 0 && BN == 64);   // 2 = wherever it fits, always with 64-channel tiles
    if (halo_on && x3 && epi_tma && pr.taps_h == 3 && pr.taps_w == 3 && pr.stride == 1 && pr.lower_h == -1 && pr.lower_w == -1 &&
        pr.P == pr.H && pr.Q == pr.W && !pr.residual && !pr.out_transposed && pr.out_s == 0 && !pr.src2 && pr.store_cols == 0 &&
        (reinterpret_cast<uintptr_t>(pr.bias) & 15) == 0) {
        // two issuers and three accumulators (1 stage at BN = 128); $I2V_TC_HALO_DUAL=0: one issuer, two stages (measured slower:
        // 339 against 294 us at 56x56 64->64)
        static const bool halo_dual = !(getenv("I2V_TC_HALO_DUAL") && atoi(getenv("I2V_TC_HALO_DUAL")) == 0);
        const bool wide = BN == 128 && halo_mode != 2;
#define I2V_HALO_CALL(bn, dual) halo_launch<bn, dual>(pr.src, pr.N, pr.H, pr.W, pr.C, pr.w_hi, pr.w_lo, pr.Cout, pr.bias, pr.mask_bits, \
                                                       pr.bits_out, pr.dst, pr.relu, st)
        const int r = wide ? (halo_dual ? I2V_HALO_CALL(128, true) : I2V_HALO_CALL(128, false))
                           : (halo_dual ? I2V_HALO_CALL(64, true) : I2V_HALO_CALL(64, false));
#undef I2V_HALO_CALL
        if (r != -1) return r;
    }

    CUtensorMap tmA, tmBhi, tmBlo, tmOut, tmRes;
    if (im2col) { if (int r = get_map_im2col(&tmA, pr.src, pr.N, pr.H, pr.W, pr.C, pr.lower_w, pr.lower_h, pr.upper_w, pr.upper_h, pr.stride)) return r; }
    else        { if (int r = get_map_2d(&tmA, pr.src, (int)M, pr.C, TC_BM)) return r; }
    if (int r = get_map_2d(&tmBhi, pr.w_hi, pr.Cout, Ktot, BN)) return r;
    if (x3) { if (int r = get_map_2d(&tmBlo, pr.w_lo, pr.Cout, Ktot, BN)) return r; }
    else tmBlo = tmBhi;
    tmOut = tmBhi; tmRes = tmBhi;
    I2V_REQUIRE(!pr.out_transposed || (epi_tma && !pr.residual && !pr.mask_bits && !pr.bits_out && !pr.relu && M % 4 == 0),
                "transposed output needs the TMA epilogue without residual / masks / ReLU and M % 4 == 0");
    if (epi_tma && pr.out_s != 0) {
        // the class's view of dst: first pixel (out_h0, out_w0), every out_s-th pixel from there
        I2V_REQUIRE(!pr.out_transposed && !pr.bits_out && pr.out_h0 < pr.out_H && pr.out_w0 < pr.out_W, "bad strided output");
        const int64_t first = ((int64_t)pr.out_h0 * pr.out_W + pr.out_w0) * pr.Cout;
        if (int r = get_map_im2col_view(&tmOut, pr.dst + first, pr.N, pr.out_H - pr.out_h0, pr.out_W - pr.out_w0, pr.Cout, pr.out_H,
                                        pr.out_W, pr.out_s)) return r;
        if (pr.residual) {
            if (int r = get_map_im2col_view(&tmRes, pr.residual + first, pr.N, pr.out_H - pr.out_h0, pr.out_W - pr.out_w0, pr.Cout,
                                            pr.out_H, pr.out_W, pr.out_s)) return r;
        }
    } else if (epi_tma) {
        if (pr.out_transposed) { if (int r = get_map_2d_plain(&tmOut, pr.dst, pr.Cout, (int)M, 32, TC_BM)) return r; }
        else                   { if (int r = get_map_2d(&tmOut, pr.dst, (int)M, pr.Cout, TC_BM)) return r; }
        if (pr.residual) { if (int r = get_map_2d(&tmRes, pr.residual, (int)M, pr.Cout, TC_BM)) return r; }
    }
    if (pr.src2) {
        I2V_REQUIRE(persistent && epi_tma && pr.taps_h == 1 && pr.taps_w == 1 && !pr.residual && !pr.out_transposed && pr.C2 > 0 &&
                    pr.C2 % 32 == 0 && (reinterpret_cast<uintptr_t>(pr.src2) & 15) == 0,
                    "a second A source needs 1x1 taps, the TMA epilogue, no residual and C2 % 32 == 0");
        if (int r = get_map_2d(&tmRes, pr.src2, (int)M, pr.C2, TC_BM)) return r;      // an A-operand map in the residual's seat
    }

    TcArgs a{};
    a.bias = pr.bias; a.residual = pr.residual; a.mask_src = pr.mask_src; a.dst = pr.dst;
    a.mask_bits = pr.mask_bits; a.bits_out = pr.bits_out;
    a.out_transposed = pr.out_transposed; a.store_cols = pr.store_cols > 0 ? pr.store_cols : pr.Cout;
    // measured (profiles/): no gain in the attack pipeline, where a layer's input was written by the previous launch and
    // is largely L2-resident already -> off by default
    static const int prefetch_tiles = getenv("I2V_TC_PREFETCH") ? atoi(getenv("I2V_TC_PREFETCH")) : 0;
    a.prefetch_tiles = prefetch_tiles > 0 ? (prefetch_tiles < 8 ? prefetch_tiles : 8) : 0;
    a.M = M; a.Cout = pr.Cout; a.P = pr.P; a.Q = pr.Q; a.stride = pr.stride; a.lower_h = pr.lower_h; a.lower_w = pr.lower_w;
    a.taps_h = pr.taps_h; a.taps_w = pr.taps_w; a.cblocks = pr.C / 32; a.relu = pr.relu;
    if (pr.src2) { a.a2_cb0 = pr.C / 32; a.cblocks = (pr.C + pr.C2) / 32; }
    a.out_s = pr.out_s; a.out_h0 = pr.out_h0; a.out_w0 = pr.out_w0; a.out_H = pr.out_H; a.out_W = pr.out_W;
    a.M_out = pr.out_s != 0 ? (int64_t)pr.N * pr.out_H * pr.out_W : M;
    a.trace = g_trace; a.trace_tiles = g_trace_tiles;
    static const int pair_dbg = getenv("I2V_TC_PAIR_DBG") ? atoi(getenv("I2V_TC_PAIR_DBG")) : 0;
    a.dbg = pair_dbg;
#define I2V_TC_DISPATCH_P(BN_, EPI_, ALO_)                                                          \
    do {                                                                                            \
        if (x3) return im2col ? tc_launch_persist<BN_, true, true, EPI_, ALO_>(tmA, tmBhi, tmBlo, tmOut, tmRes, a, st)      \
                              : tc_launch_persist<BN_, true, false, EPI_, ALO_>(tmA, tmBhi, tmBlo, tmOut, tmRes, a, st);    \
        return im2col ? tc_launch_persist<BN_, false, true, EPI_, false>(tmA, tmBhi, tmBlo, tmOut, tmRes, a, st)            \
                      : tc_launch_persist<BN_, false, false, EPI_, false>(tmA, tmBhi, tmBlo, tmOut, tmRes, a, st);          \
    } while (0)
    if (persistent) {
        if (epi_tma) {
            // dual-issuer variant (A_lo in tensor memory): BN = 64 always (two accumulator stages remain), BN = 128 for
            // tiles of >= 4 k-steps only — with its single accumulator stage the shortest tiles (K = 64) of the HBM-bound 1x1
            // layers lose the main-loop / epilogue overlap (measured per 256 frames: 64->256 +res 418 vs 360 us at threshold
            // 2; 256->128 dgrad (K = 128) 357 vs 429 us at threshold 4 against 8).  $I2V_TC_ALO_MINKIT overrides.
            // $I2V_TC_ALO_TMEM=0 selects the single-issuer kernel everywhere, =2 the dual-issuer kernel everywhere
            static const int alo_env = getenv("I2V_TC_ALO_TMEM") ? atoi(getenv("I2V_TC_ALO_TMEM")) : 1;
            const int kit = pr.taps_h * pr.taps_w * (pr.C / 32) + (pr.src2 ? pr.C2 / 32 : 0);
            static const int alo_minkit = getenv("I2V_TC_ALO_MINKIT") ? atoi(getenv("I2V_TC_ALO_MINKIT")) : 4;
            // (a dual-source launch of 4 k-steps — layer1.0's downsample + conv3 — runs 430 us on the single-issuer kernel
            // with its two accumulator stages against 475 us: threshold 8 there)
            const bool alo = x3 && alo_env != 0 && (BN == 64 || kit >= (pr.src2 ? 2 * alo_minkit : alo_minkit) || alo_env == 2 || pr.force_dual);
            // CTA pairs (tcgen05.mma.cta_group::2, half of the weight rows per CTA): $I2V_TC_PAIR = minimum k-steps per tile
            // from which the pair kernel takes over (0 = never); needs at least two m-tiles
            const int pair_minkit = pair_min_ksteps();
            // default rule (-1), per-shape A/B in situ (gpurun_out/r3h_*): a residual streaming through the epilogue, or a
            // plain 1x1 with BN = 128 tiles of 4-8 k-steps and at most two n-tiles (56x56 256->128: fwd 320 -> 298 us, dgrad
            // 432 -> 338 us); the im2col layers and the BN = 64 ones lose on the pair
            const bool pair_on = pair_minkit > 0 ? kit >= pair_minkit
                               : (pair_minkit == -1 && kit >= 4 &&
                                  (pr.residual != nullptr || (!im2col && BN == 128 && kit <= 8 && pr.Cout / BN <= 2)));
            if (alo && pair_on && !pr.out_transposed && M > TC_BM && !pr.src2 && pr.out_s == 0) {
                CUtensorMap hBhi, hBlo;
                if (int r = get_map_2d(&hBhi, pr.w_hi, pr.Cout, Ktot, BN / 2)) return r;
                if (int r = get_map_2d(&hBlo, pr.w_lo, pr.Cout, Ktot, BN / 2)) return r;
                if (BN == 128) return im2col ? tc_launch_pair<128, true>(tmA, hBhi, hBlo, tmOut, tmRes, a, st)
                                             : tc_launch_pair<128, false>(tmA, hBhi, hBlo, tmOut, tmRes, a, st);
                return im2col ? tc_launch_pair<64, true>(tmA, hBhi, hBlo, tmOut, tmRes, a, st)
                              : tc_launch_pair<64, false>(tmA, hBhi, hBlo, tmOut, tmRes, a, st);
            }
            if (alo) {
                if (BN == 128) I2V_TC_DISPATCH_P(128, true, true);
                I2V_TC_DISPATCH_P(64, true, true);
            }
            if (BN == 128) I2V_TC_DISPATCH_P(128, true, false);
            I2V_TC_DISPATCH_P(64, true, false);
        }
        // register epilogue (strided data-gradient classes, f32 mask sources): dual issuers are available behind
        // $I2V_TC_ALO_REGEPI=1 but OFF by default — measured on the stride-2 classes of ResNet's layer2.0 (256 frames):
        // 234 us per launch with them against 171 us without; the per-thread global stores of this epilogue are slow
        // enough that the single accumulator stage of the BN = 128 dual-issuer layout costs more than the issue rate gains
        {
            static const int alo_env2 = getenv("I2V_TC_ALO_REGEPI") ? atoi(getenv("I2V_TC_ALO_REGEPI")) : 0;
            static const int alo_minkit2 = getenv("I2V_TC_ALO_MINKIT") ? atoi(getenv("I2V_TC_ALO_MINKIT")) : 4;
            const int kit2 = pr.taps_h * pr.taps_w * (pr.C / 32);
            if (x3 && alo_env2 != 0 && (BN == 64 || kit2 >= alo_minkit2)) {
                if (BN == 128) I2V_TC_DISPATCH_P(128, false, true);
                I2V_TC_DISPATCH_P(64, false, true);
            }
        }
        if (BN == 128) I2V_TC_DISPATCH_P(128, false, false);
        I2V_TC_DISPATCH_P(64, false, false);
    }
#undef I2V_TC_DISPATCH_P
#define I2V_TC_DISPATCH(BN_)                                                                        \
    do {                                                                                            \
        if (x3) return im2col ? tc_launch<BN_, true, true>(tmA, tmBhi, tmBlo, a, st) : tc_launch<BN_, true, false>(tmA, tmBhi, tmBlo, a, st);   \
        return im2col ? tc_launch<BN_, false, true>(tmA, tmBhi, tmBlo, a, st) : tc_launch<BN_, false, false>(tmA, tmBhi, tmBlo, a, st);          \
    } while (0)
    if (BN == 128) I2V_TC_DISPATCH(128);
    I2V_TC_DISPATCH(64);
#undef I2V_TC_DISPATCH
}

}  // namespace i2v

using namespace i2v;

extern "C" int i2v_conv_tc_set_pair_minkit(int min_ksteps) {
    g_pair_minkit = min_ksteps < -1 ? -1 : min_ksteps;
    return I2V_OK;
}

extern "C" int i2v_conv_tc_set_halo_mode(int mode) {
    g_halo_mode = mode < -1 ? -1 : mode;
    return I2V_OK;
}

extern "C" int i2v_conv_tc_set_trace(unsigned long long* device_buf, int tiles) {
    g_trace = device_buf;
    g_trace_tiles = device_buf ? tiles : 0;
    return I2V_OK;
}

extern "C" int i2v_conv_tc_supported(const i2v_conv_desc* d, int dgrad) {
    if (!d) return 0;
    if (d->R != d->S || d->pad >= d->R) return 0;
    if (!dgrad) return (d->Cin % 32 == 0) && (d->Cout % 64 == 0);
    // stride 1: a forward-style implicit GEMM over dy with the flipped filter (i2v_conv_tc_f32, dgrad = 1);
    // stride > 1: one launch per stride-parity class (i2v_conv_tc_dgrad_class_f32)
    return (d->Cout % 32 == 0) && (d->Cin % 64 == 0) && d->stride >= 1 && d->stride <= 4;
}

// Forward:  src = x  [N,H,W,Cin],  dst = y  [N,P,Q,Cout], w_* = [Cout, R*S*Cin]  K-major (tap-major, channel-minor)
// Dgrad  :  src = dy [N,P,Q,Cout], dst = dx [N,H,W,Cin],  w_* = [Cin, R*S*Cout] with the filter flipped (host); stride 1
// mask_bits: [C_dst/32][M] words, M = rows of dst; forward: OUTPUT (activity bits of dst, optional); dgrad: INPUT mask
extern "C" int i2v_conv_tc_bits_f32(const i2v_conv_desc* d, int dgrad, const float* src, const float* w_hi, const float* w_lo,
                                    const float* bias, const float* residual, const float* mask_src, uint32_t* mask_bits,
                                    float* dst, int flags, i2v_stream_t stream) {
    I2V_REQUIRE(d && src && w_hi && dst, "null pointer");
    I2V_REQUIRE(i2v_conv_tc_supported(d, dgrad) && (!dgrad || d->stride == 1), "shape not supported by this tensor-core entry point");
    I2V_REQUIRE(!(mask_bits && mask_src), "pass either an f32 mask source or a bit mask, not both");
    if (d->N == 0) return I2V_OK;
    TcProblem pr{};
    pr.src = src; pr.N = d->N; pr.w_hi = w_hi; pr.w_lo = w_lo; pr.bias = bias; pr.residual = residual; pr.mask_src = mask_src;
    pr.dst = dst; pr.relu = (flags & I2V_EPI_RELU) ? 1 : 0;
    pr.taps_h = d->R; pr.taps_w = d->S;
    if (!dgrad) {
        pr.H = d->H; pr.W = d->W; pr.C = d->Cin; pr.P = d->P; pr.Q = d->Q; pr.Cout = d->Cout; pr.stride = d->stride;
        pr.lower_h = pr.lower_w = -d->pad; pr.upper_h = d->pad - (d->R - 1); pr.upper_w = d->pad - (d->S - 1);
        pr.bits_out = mask_bits;
    } else {
        const int padp = d->R - 1 - d->pad;
        pr.H = d->P; pr.W = d->Q; pr.C = d->Cout; pr.P = d->H; pr.Q = d->W; pr.Cout = d->Cin; pr.stride = 1;
        pr.lower_h = pr.lower_w = -padp; pr.upper_h = padp - (d->R - 1); pr.upper_w = padp - (d->S - 1);
        pr.mask_bits = mask_bits;
    }
    return tc_run(pr, as_stream(stream));
}

// Two 1x1 convolutions whose outputs are added — a bottleneck's downsample branch (d: 1x1, stride s, over x) and its last
// convolution (1x1 / stride 1 over t [N,P,Q,C2]) — as ONE GEMM: K = Cin followed by C2, w_* = [Cout, Cin + C2] K-major (the two
// folded weight matrices side by side), bias = the sum of the two.  The downsample output (4 bytes per output element written
// and read back as the residual) never exists.  Forward only; bits_out as for i2v_conv_tc_bits_f32.
extern "C" int i2v_conv_tc_dual_f32(const i2v_conv_desc* d, const float* x, int C2, const float* t, const float* w_hi,
                                    const float* w_lo, const float* bias, uint32_t* bits_out, float* dst, int flags,
                                    i2v_stream_t stream) {
    I2V_REQUIRE(d && x && t && w_hi && dst, "null pointer");
    I2V_REQUIRE(i2v_conv_tc_supported(d, 0) && d->R == 1 && d->S == 1 && d->pad == 0, "the first convolution must be a supported 1x1");
    I2V_REQUIRE(C2 > 0 && C2 % 32 == 0, "C2 must be a positive multiple of 32");
    if (d->N == 0) return I2V_OK;
    TcProblem pr{};
    pr.src = x; pr.N = d->N; pr.w_hi = w_hi; pr.w_lo = w_lo; pr.bias = bias; pr.dst = dst; pr.relu = (flags & I2V_EPI_RELU) ? 1 : 0;
    pr.taps_h = pr.taps_w = 1;
    pr.H = d->H; pr.W = d->W; pr.C = d->Cin; pr.P = d->P; pr.Q = d->Q; pr.Cout = d->Cout; pr.stride = d->stride;
    pr.lower_h = pr.lower_w = 0; pr.upper_h = pr.upper_w = 0;
    pr.bits_out = bits_out;
    pr.src2 = t; pr.C2 = C2;
    return tc_run(pr, as_stream(stream));
}

extern "C" int i2v_conv_tc_f32(const i2v_conv_desc* d, int dgrad, const float* src, const float* w_hi, const float* w_lo,
                               const float* bias, const float* residual, const float* mask_src, float* dst, int flags,
                               i2v_stream_t stream) {
    return i2v_conv_tc_bits_f32(d, dgrad, src, w_hi, w_lo, bias, residual, mask_src, nullptr, dst, flags, stream);
}

// First-layer data gradient on the tensor cores.  dcost/dimage[n,c,h,w] = sum_{r,s,co} dy[n,p,q,co] * W[co,c,r,s] with
// h = stride*p - pad + r (w likewise) has only three output channels: as an implicit GEMM over image pixels it would
// waste the tensor core on N = 3.  Instead the contraction over co is done FIRST, as a plain GEMM
//     Z[(n,p,q), (c,r,s)] = sum_co dy[(n,p,q), co] * W[co, (c,r,s)]          (M = N*P*Q, N = 3*R*S padded, K = Cout)
// whose result is stored transposed (planes Z^T[(c,r,s)][m]) by the TMA epilogue, and a col2im pass then gathers
// the <= ceil(R/stride)^2 taps of every image pixel from those planes with fully coalesced reads (each Z element
// is read exactly once).  wz_* = [NZ, Cout] K-major = w_stem rows (c,r,s) zero-padded to NZ = ceil(3*R*S / 64) * 64.
// bytes of scratch a frame group may take ($I2V_STEM_GROUP_MB).  Measured (profiles/r01_bench_engine_chunk_sweep.json):
// groups small enough to keep the scratch L2-resident (48 MB = 6 frames at 224^2) lose more to the ~15 us fixed cost of
// every extra launch pair than they save in HBM traffic — 256 frames: 2.10 / 1.81 ms grouped vs 1.49 / 1.30 ms in one
// pass (dgrad / fwd) — so the default only bounds the scratch allocation (4 GB).
// Rows (c,r,s) of the first-layer dgrad weight matrix are zero-padded to a multiple of 64 (BN = 64 tiles) or, with
// $I2V_STEM_NZ=128, of 128: BN = 128 tiles need a third fewer MMA instructions per pixel tile (the cost of an instruction
// does not depend on its N) at the price of a single accumulator stage.  Measured (256 frames of ResNet's stem): 1.71 ms
// with 128 against 1.37 ms with 64 — the two-k-step tiles of this GEMM need the main-loop / epilogue overlap more than they
// need fewer instructions, so 64 stays the default.
static int stem_nz_multiple() {
    static const int m = getenv("I2V_STEM_NZ") ? atoi(getenv("I2V_STEM_NZ")) : 64;
    return m == 128 ? 128 : 64;
}

extern "C" int i2v_conv_stem_dgrad_tc_rows(int cols) {
    const int m = stem_nz_multiple();
    return (cols + m - 1) / m * m;
}

static int64_t stem_group_bytes() {
    static const int64_t mb = getenv("I2V_STEM_GROUP_MB") ? atoll(getenv("I2V_STEM_GROUP_MB")) : 4096;
    return (mb > 0 ? mb : 4096) << 20;
}

extern "C" int i2v_conv_stem_dgrad_tc_group(const i2v_conv_desc* d) {
    if (!d) return 0;
    const int64_t per_frame = (int64_t)d->P * d->Q * ((3 * d->R * d->S + 31) / 32 * 32) * 4;
    int64_t g = stem_group_bytes() / (per_frame > 0 ? per_frame : 1);
    if (g < 1) g = 1;
    if (g > d->N) g = d->N;
    while (g > 1 && (g * d->P * d->Q) % 4 != 0) --g;             // transposed TMA store: plane pitch multiple of 16 bytes
    return (int)g;
}

extern "C" int i2v_conv_stem_dgrad_tc_f32(const i2v_conv_desc* d, const float* dy, const float* wz_hi, const float* wz_lo,
                                          float* z_scratch, float* dx, i2v_stream_t stream) {
    I2V_REQUIRE(d && dy && wz_hi && z_scratch && dx, "null pointer");
    I2V_REQUIRE(d->Cin == 3 && d->R == d->S && d->Cout % 32 == 0 && d->stride >= 1 && d->pad < d->R, "not a first-layer shape");
    if (d->N == 0) return I2V_OK;
    const int cols = 3 * d->R * d->S;
    const int NZ = i2v_conv_stem_dgrad_tc_rows(cols);
    // frames in groups that bound the scratch (see stem_group_bytes)
    const int G = i2v_conv_stem_dgrad_tc_group(d);
    for (int n0 = 0; n0 < d->N; n0 += G) {
        const int n = d->N - n0 < G ? d->N - n0 : G;
        I2V_REQUIRE(((int64_t)n * d->P * d->Q) % 4 == 0, "frames-per-group * P * Q must be a multiple of 4");
        TcProblem pr{};
        pr.src = dy + (int64_t)n0 * d->P * d->Q * d->Cout; pr.N = n; pr.H = d->P; pr.W = d->Q; pr.C = d->Cout;
        pr.P = d->P; pr.Q = d->Q; pr.stride = 1;
        pr.taps_h = pr.taps_w = 1;
        pr.w_hi = wz_hi; pr.w_lo = wz_lo; pr.Cout = NZ; pr.dst = z_scratch;
        pr.out_transposed = 1; pr.store_cols = (cols + 31) / 32 * 32;
        pr.force_dual = NZ % 128 == 0;
        if (int r = tc_run(pr, as_stream(stream))) return r;
        if (int r = stem_col2im_launch(z_scratch, dx + (int64_t)n0 * 3 * d->H * d->W, n, d->H, d->W, d->P, d->Q, d->R, d->stride,
                                       d->pad, as_stream(stream))) return r;
    }
    return I2V_OK;
}

// First-layer data gradient without scratch (stem_dgrad_direct_kernel).  wd_hi / wd_lo = [160, 64] K-major: the 147 taps
// k = (c, r, s) of w_stem followed by 13 zero rows — split into hi = w (the tensor core truncates) and lo = w - trunc_tf32(w).
extern "C" int i2v_conv_stem_dgrad_direct_supported(const i2v_conv_desc* d) {
    return d && d->Cin == 3 && d->Cout == 64 && d->R == 7 && d->S == 7 && d->stride == 2 && d->pad == 3 && d->Q >= 1 &&
           d->Q <= TC_BM && d->P >= 1 && d->W <= 2 * d->Q && d->H <= 2 * d->P;
}

extern "C" int i2v_conv_stem_dgrad_direct_f32(const i2v_conv_desc* d, const float* dy, const float* wd_hi, const float* wd_lo,
                                              float* dx, i2v_stream_t stream) {
    I2V_REQUIRE(d && dy && wd_hi && wd_lo && dx, "null pointer");
    I2V_REQUIRE(i2v_conv_stem_dgrad_direct_supported(d), "shape not supported by the direct first-layer data gradient");
    I2V_REQUIRE(((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(wd_hi) | reinterpret_cast<uintptr_t>(wd_lo) |
                  reinterpret_cast<uintptr_t>(dx)) & 15) == 0, "all tensors must be 16-byte aligned");
    if (d->N == 0) return I2V_OK;
    if (int r = resolve_driver()) return r;
    const int64_t M = (int64_t)d->N * d->P * d->Q;
    I2V_REQUIRE(M < (int64_t)0x7fffffff, "too many pixels for one launch");
    CUtensorMap tmA, tmBhi, tmBlo;
    if (int r = get_map_2d(&tmA, dy, (int)M, 64, TC_BM)) return r;
    if (int r = get_map_2d(&tmBhi, wd_hi, SD_NZ, 64, SD_NZ)) return r;
    if (int r = get_map_2d(&tmBlo, wd_lo, SD_NZ, 64, SD_NZ)) return r;
    StemDirectArgs a{};
    a.dx = dx; a.N = d->N; a.H = d->H; a.W = d->W; a.P = d->P; a.Q = d->Q;
    static const int sd_dbg = getenv("I2V_STEM_DBG") ? atoi(getenv("I2V_STEM_DBG")) : 0;
    a.dbg = sd_dbg;
    a.strips_per_image = d->H >= 32 ? 2 : 1;
    a.units = d->N * a.strips_per_image;
    const size_t smem = 1024 + 4 * SD_B_TILE + (size_t)SD_STAGES * 2 * TC_A_BYTES + 3 * 2 * 6 * 42 * sizeof(float) + 256;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(stem_dgrad_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "i2v_conv_stem_dgrad_direct_f32 (shared memory)");
        attr_done = true;
    }
    const int grid = a.units < sm_count() ? a.units : sm_count();
    stem_dgrad_direct_kernel<<<grid, SD_THREADS, smem, as_stream(stream)>>>(tmA, tmBhi, tmBlo, a);
    I2V_LAUNCH_CHECK("i2v_conv_stem_dgrad_direct_f32");
    return I2V_OK;
}

// The same with the 3x3 / stride-2 / pad-1 max pooling's backward pass in front (stem_dgrad_pool_kernel): dy_pooled / argmax =
// [N, P2, Q2, 64] gradient of the POOLED map and the argmax plane i2v_maxpool_fwd_flags_f32 wrote (mark_dead mode: the stem's
// ReLU mask is folded into it).  Bit-identical to i2v_maxpool_bwd_f32 followed by i2v_conv_stem_dgrad_direct_f32; the gradient
// of the stem activation (4x the pooled bytes) never exists.
static size_t stem_dgrad_pool_smem(int Q2) {
    return 1024 + 4 * SD_B_TILE + (size_t)SP_AST * 2 * TC_A_BYTES + (size_t)SP_PST * (4 * sp_gtile_bytes(Q2) + 2 * sp_atile_bytes(Q2)) +
           3 * 2 * 6 * 42 * sizeof(float) + 384;
}

extern "C" int i2v_conv_stem_dgrad_pool_supported(const i2v_conv_desc* d, int P2, int Q2) {
    if (!i2v_conv_stem_dgrad_direct_supported(d)) return 0;
    if (P2 != (d->P - 1) / 2 + 1 || Q2 != (d->Q - 1) / 2 + 1) return 0;          // 3x3 / stride 2 / pad 1, floor mode
    int dev = 0, optin = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    return stem_dgrad_pool_smem(Q2) <= (size_t)optin;
}

extern "C" int i2v_conv_stem_dgrad_pool_f32(const i2v_conv_desc* d, int P2, int Q2, const float* dy_pooled, const uint8_t* argmax,
                                            const float* wd_hi, const float* wd_lo, float* dx, i2v_stream_t stream) {
    I2V_REQUIRE(d && dy_pooled && argmax && wd_hi && wd_lo && dx, "null pointer");
    I2V_REQUIRE(i2v_conv_stem_dgrad_pool_supported(d, P2, Q2), "shape not supported by the pooled first-layer data gradient");
    I2V_REQUIRE(((reinterpret_cast<uintptr_t>(dy_pooled) | reinterpret_cast<uintptr_t>(argmax) | reinterpret_cast<uintptr_t>(wd_hi) |
                  reinterpret_cast<uintptr_t>(wd_lo) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0, "all tensors must be 16-byte aligned");
    if (d->N == 0) return I2V_OK;
    if (int r = resolve_driver()) return r;
    const int64_t M2 = (int64_t)d->N * P2 * Q2;
    I2V_REQUIRE(M2 < (int64_t)0x7fffffff, "too many pixels for one launch");
    CUtensorMap tmG, tmAm, tmBhi, tmBlo;
    if (int r = get_map_2d(&tmG, dy_pooled, (int)M2, 64, Q2)) return r;
    if (int r = get_map_u8_rows64(&tmAm, argmax, (int)M2, Q2)) return r;
    if (int r = get_map_2d(&tmBhi, wd_hi, SD_NZ, 64, SD_NZ)) return r;
    if (int r = get_map_2d(&tmBlo, wd_lo, SD_NZ, 64, SD_NZ)) return r;
    StemDirectArgs a{};
    a.dx = dx; a.N = d->N; a.H = d->H; a.W = d->W; a.P = d->P; a.Q = d->Q; a.P2 = P2; a.Q2 = Q2;
    a.dbg = getenv("I2V_STEM_DBG") ? atoi(getenv("I2V_STEM_DBG")) : 0;          // timing experiments; re-read per call
    a.strips_per_image = d->H >= 32 ? 2 : 1;
    a.units = d->N * a.strips_per_image;
    const size_t smem = stem_dgrad_pool_smem(Q2);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(stem_dgrad_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "i2v_conv_stem_dgrad_pool_f32 (shared memory)");
        smem_set = smem;
    }
    const int grid = a.units < sm_count() ? a.units : sm_count();
    stem_dgrad_pool_kernel<<<grid, SD_THREADS, smem, as_stream(stream)>>>(tmG, tmAm, tmBhi, tmBlo, a);
    I2V_LAUNCH_CHECK("i2v_conv_stem_dgrad_pool_f32");
    return I2V_OK;
}

// First-layer forward without the patch matrix (stem_fwd_rows_kernel).  wk_hi / wk_lo = [64, 160] K-major, k = (c,r,s) + 13
// zero columns: exactly the operands of i2v_conv_stem_fwd_tc_f32.
extern "C" int i2v_conv_stem_fwd_rows_supported(const i2v_conv_desc* d) {
    return d && d->Cin == 3 && d->Cout == 64 && d->R == 7 && d->S == 7 && d->stride == 2 && d->pad == 3 && d->Q >= 1 &&
           d->Q <= TC_BM && d->P >= 1 && d->W % 4 == 0 && d->W + 10 <= 256 && d->H >= 1;
}

extern "C" int i2v_conv_stem_fwd_rows_f32(const i2v_conv_desc* d, const float* x, const float* wk_hi, const float* wk_lo,
                                          const float* bias, float* y, int flags, i2v_stream_t stream) {
    I2V_REQUIRE(d && x && wk_hi && wk_lo && y, "null pointer");
    I2V_REQUIRE(i2v_conv_stem_fwd_rows_supported(d), "shape not supported by the row-tile first-layer forward");
    I2V_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(wk_hi) | reinterpret_cast<uintptr_t>(wk_lo) |
                  reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(bias)) & 15) == 0, "all tensors must be 16-byte aligned");
    if (d->N == 0) return I2V_OK;
    if (int r = resolve_driver()) return r;
    const int64_t M = (int64_t)d->N * d->P * d->Q;
    I2V_REQUIRE(M < (int64_t)0x7fffffff, "too many pixels for one launch");
    const int pitch = (d->W + 7 + 3) / 4 * 4;
    CUtensorMap tmX, tmBhi, tmBlo, tmOut;
    if (int r = get_map_planes(&tmX, x, d->N * 3, d->H, d->W, pitch, 7)) return r;
    if (int r = get_map_2d(&tmBhi, wk_hi, 64, SF_KB * TC_BK, 64)) return r;
    if (int r = get_map_2d(&tmBlo, wk_lo, 64, SF_KB * TC_BK, 64)) return r;
    if (int r = get_map_2d(&tmOut, y, (int)M, 64, d->Q)) return r;
    StemFwdArgs a{};
    a.bias = bias; a.N = d->N; a.H = d->H; a.W = d->W; a.P = d->P; a.Q = d->Q; a.relu = (flags & I2V_EPI_RELU) ? 1 : 0;
    a.pitch = pitch; a.tiles = d->N * d->P;
    a.cstride = (7 * pitch * 4 + 127) / 128 * 32;
    const size_t in_stride = (size_t)3 * a.cstride * 4;
    const size_t smem = 1024 + SF_KB * SF_B_KB + (size_t)SF_ASLOTS * TC_A_BYTES + 2 * EPI_SLOT_BYTES + SF_ISTAGES * in_stride + 512;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(stem_fwd_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "i2v_conv_stem_fwd_rows_f32 (shared memory)");
        smem_set = smem;
    }
    const int grid = a.tiles < sm_count() ? a.tiles : sm_count();
    stem_fwd_rows_kernel<false><<<grid, SF_THREADS, smem, as_stream(stream)>>>(tmX, tmBhi, tmBlo, tmOut, a);
    I2V_LAUNCH_CHECK("i2v_conv_stem_fwd_rows_f32");
    return I2V_OK;
}

// The same with the 3x3 / stride-2 / pad-1 max pooling behind it fused into the epilogue (stem_fwd_rows_kernel<true>): pooled
// [N, P2, Q2, 64] and its argmax plane come out, the 4x larger stem activation is never written or read back.  A CTA walks
// strips of S pooled rows of one image (conv rows 2 i0 - 1 .. 2 i1 - 1 in order: one lead-in row per strip is computed twice);
// S is chosen for the best product of wave fill and (2S) / (2S + 1).
extern "C" int i2v_conv_stem_fwd_pool_supported(const i2v_conv_desc* d, int P2, int Q2) {
    return i2v_conv_stem_fwd_rows_supported(d) && d->P % 2 == 0 && d->Q % 2 == 0 && P2 == d->P / 2 && Q2 == d->Q / 2;
}

extern "C" int i2v_conv_stem_fwd_pool_f32(const i2v_conv_desc* d, int P2, int Q2, const float* x, const float* wk_hi,
                                          const float* wk_lo, const float* bias, float* pooled, uint8_t* argmax, int flags,
                                          i2v_stream_t stream) {
    I2V_REQUIRE(d && x && wk_hi && wk_lo && pooled && argmax, "null pointer");
    I2V_REQUIRE(i2v_conv_stem_fwd_pool_supported(d, P2, Q2), "shape not supported by the first-layer forward with fused pooling");
    I2V_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(wk_hi) | reinterpret_cast<uintptr_t>(wk_lo) |
                  reinterpret_cast<uintptr_t>(pooled) | reinterpret_cast<uintptr_t>(argmax) | reinterpret_cast<uintptr_t>(bias)) & 15) == 0,
                "all tensors must be 16-byte aligned");
    if (d->N == 0) return I2V_OK;
    if (int r = resolve_driver()) return r;
    I2V_REQUIRE((int64_t)d->N * d->P * d->Q < (int64_t)0x7fffffff, "too many pixels for one launch");
    const int pitch = (d->W + 7 + 3) / 4 * 4;
    CUtensorMap tmX, tmBhi, tmBlo;
    if (int r = get_map_planes(&tmX, x, d->N * 3, d->H, d->W, pitch, 7)) return r;
    if (int r = get_map_2d(&tmBhi, wk_hi, 64, SF_KB * TC_BK, 64)) return r;
    if (int r = get_map_2d(&tmBlo, wk_lo, 64, SF_KB * TC_BK, 64)) return r;
    StemFwdArgs a{};
    a.bias = bias; a.N = d->N; a.H = d->H; a.W = d->W; a.P = d->P; a.Q = d->Q; a.relu = (flags & I2V_EPI_RELU) ? 1 : 0;
    a.pitch = pitch; a.tiles = d->N * d->P;
    a.cstride = (7 * pitch * 4 + 127) / 128 * 32;
    a.pooled = pooled; a.argmax = argmax; a.P2 = P2; a.Q2 = Q2; a.mark_dead = (flags & 4) ? 1 : 0;
    const int sms = sm_count();
    double best = -1.0;
    for (int S = 1; S <= P2; ++S) {
        const int U = (P2 + S - 1) / S;
        const int64_t units = (int64_t)d->N * U;
        const int64_t waves = (units + sms - 1) / sms;
        // rows computed per image: 2 P2 + (U - 1) lead-in rows; fill of the last wave
        const double eff = ((double)units / (double)(waves * sms)) * (2.0 * P2 / (2.0 * P2 + U - 1));
        if (eff > best + 1e-9) { best = eff; a.S = S; a.U = U; }
    }
    a.units = d->N * a.U;
    const size_t in_stride = (size_t)3 * a.cstride * 4;
    const size_t smem = 1024 + SF_KB * SF_B_KB + (size_t)SF_ASLOTS * TC_A_BYTES + 2 * EPI_SLOT_BYTES + SF_ISTAGES * in_stride + 512;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(stem_fwd_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "i2v_conv_stem_fwd_pool_f32 (shared memory)");
        smem_set = smem;
    }
    const int grid = a.units < sms ? a.units : sms;
    stem_fwd_rows_kernel<true><<<grid, SF_THREADS, smem, as_stream(stream)>>>(tmX, tmBhi, tmBlo, tmBhi, a);
    I2V_LAUNCH_CHECK("i2v_conv_stem_fwd_pool_f32");
    return I2V_OK;
}

// First-layer forward on the tensor cores: a 12-byte pixel cannot be a TMA row, so the patch matrix
// col[(n,p,q)][Kp] (k = (c,r,s), Kp = 3*R*S rounded up to 32) is materialised by an im2col pass and the convolution
// becomes a plain GEMM with K = Kp (bias + ReLU in the TMA epilogue).  Frames are processed in groups that bound the
// scratch (see stem_group_bytes).
// wk_* = [Cout, Kp] K-major; col_scratch holds i2v_conv_stem_fwd_tc_group(d) * P*Q*Kp floats.
extern "C" int i2v_conv_stem_fwd_tc_group(const i2v_conv_desc* d) {
    if (!d) return 0;
    const int Kp = (3 * d->R * d->S + 31) / 32 * 32;
    const int64_t per_frame = (int64_t)d->P * d->Q * Kp * 4;
    int64_t g = stem_group_bytes() / (per_frame > 0 ? per_frame : 1);
    if (g < 1) g = 1;
    if (g > d->N) g = d->N;
    return (int)g;
}

extern "C" int i2v_conv_stem_fwd_tc_f32(const i2v_conv_desc* d, const float* x, const float* wk_hi, const float* wk_lo,
                                        const float* bias, float* col_scratch, float* y, int flags, i2v_stream_t stream) {
    I2V_REQUIRE(d && x && wk_hi && col_scratch && y, "null pointer");
    I2V_REQUIRE(d->Cin == 3 && d->R == d->S && d->Cout % 64 == 0 && d->stride >= 1 && d->pad < d->R, "not a first-layer shape");
    if (d->N == 0) return I2V_OK;
    const int Kp = (3 * d->R * d->S + 31) / 32 * 32;
    const int G = i2v_conv_stem_fwd_tc_group(d);
    for (int n0 = 0; n0 < d->N; n0 += G) {
        const int n = d->N - n0 < G ? d->N - n0 : G;
        if (int r = stem_im2col_launch(x, col_scratch, n0, n, d->H, d->W, d->P, d->Q, d->R, d->stride, d->pad, Kp, as_stream(stream)))
            return r;
        TcProblem pr{};
        pr.src = col_scratch; pr.N = n; pr.H = d->P; pr.W = d->Q; pr.C = Kp; pr.P = d->P; pr.Q = d->Q; pr.stride = 1;
        pr.taps_h = pr.taps_w = 1;
        pr.w_hi = wk_hi; pr.w_lo = wk_lo; pr.Cout = d->Cout; pr.bias = bias; pr.relu = (flags & I2V_EPI_RELU) ? 1 : 0;
        pr.dst = y + (int64_t)n0 * d->P * d->Q * d->Cout;
        if (int r = tc_run(pr, as_stream(stream))) return r;
    }
    return I2V_OK;
}

// EXPERIMENTAL ($I2V_STEM_DIRECT=1 in the engine): first-layer forward WITHOUT the patch matrix.  The image is packed once
// into a zero-padded NHWC4 copy xp [N, Hp, Wp, 4] (Hp = H + 2 pad, Wp >= W + 2 pad; 16 bytes per pixel), and filter row r
// of output pixel (p, q) is then the 32 contiguous floats xp[n, stride*p + r, stride*q .. stride*q + 7, 0..3] — S <= 8
// taps x 3 channels, the rest meets zero weights.  A TMA tile is a 16 x 8 box of output pixels (4-D tiled map whose q
// stride of 32 bytes overlaps the 128-byte window; one map per row parity, stride 2 only), K = R k-steps of 32, the
// GEMM and its TMA epilogue (4-D store of the same box) are the persistent dual-issuer kernel.
// Against the im2col path this drops ~2 GB of scratch traffic each way per 256 frames at 224^2 and adds 40 % MMA work.
// wr_* = [Cout, R*32] K-major, k = r*32 + s*4 + c (zero where s >= S or c == 3), TF32 hi / lo split.
extern "C" int i2v_conv_stem_fwd_direct_supported(const i2v_conv_desc* d) {
    return d && d->Cin == 3 && d->R == d->S && d->S >= 5 && d->S <= 8 && d->stride == 2 && d->Cout == 64 && d->pad < d->R;
}

static void stem_direct_geometry(const i2v_conv_desc* d, int* Hp, int* Wp) {
    *Hp = d->H + 2 * d->pad;
    const int need = (d->Q - 1) * d->stride + 8;
    const int wp = d->W + 2 * d->pad;
    *Wp = wp > need ? wp : need;
}

extern "C" int64_t i2v_conv_stem_fwd_direct_scratch_floats(const i2v_conv_desc* d) {
    if (!d) return 0;
    int Hp, Wp;
    stem_direct_geometry(d, &Hp, &Wp);
    return (int64_t)d->N * Hp * Wp * 4;
}

extern "C" int i2v_conv_stem_fwd_direct_f32(const i2v_conv_desc* d, const float* x, const float* wr_hi, const float* wr_lo,
                                            const float* bias, float* xp_scratch, float* y, int flags, i2v_stream_t stream) {
    I2V_REQUIRE(d && x && wr_hi && wr_lo && xp_scratch && y, "null pointer (the direct first-layer forward is 3xTF32 only)");
    I2V_REQUIRE(i2v_conv_stem_fwd_direct_supported(d), "shape not supported by the direct first-layer forward");
    if (d->N == 0) return I2V_OK;
    if (int r = resolve_driver()) return r;
    int Hp, Wp;
    stem_direct_geometry(d, &Hp, &Wp);
    if (int r = stem_pack_nhwc4_launch(x, xp_scratch, d->N, d->H, d->W, Hp, Wp, d->pad, as_stream(stream))) return r;
    const int st = d->stride;
    CUtensorMap tmA, tmAodd, tmBhi, tmBlo, tmOut;
    const uint64_t row_bytes = (uint64_t)Wp * 16, img_bytes = (uint64_t)Hp * row_bytes;
    {
        const uint64_t dims[4] = {TC_BK, (uint64_t)d->Q, (uint64_t)((Hp + 1) / 2), (uint64_t)d->N};
        const uint64_t strides[3] = {(uint64_t)st * 16, (uint64_t)st * row_bytes, img_bytes};
        if (int r = get_map_4d(&tmA, xp_scratch, dims, strides)) return r;
        const uint64_t dims_odd[4] = {TC_BK, (uint64_t)d->Q, (uint64_t)(Hp / 2), (uint64_t)d->N};
        if (int r = get_map_4d(&tmAodd, xp_scratch + (size_t)Wp * 4, dims_odd, strides)) return r;
    }
    {
        const uint64_t dims[4] = {(uint64_t)d->Cout, (uint64_t)d->Q, (uint64_t)d->P, (uint64_t)d->N};
        const uint64_t strides[3] = {(uint64_t)d->Cout * 4, (uint64_t)d->Q * d->Cout * 4, (uint64_t)d->P * d->Q * d->Cout * 4};
        if (int r = get_map_4d(&tmOut, y, dims, strides)) return r;
    }
    const int Ktot = d->R * TC_BK;
    if (int r = get_map_2d(&tmBhi, wr_hi, d->Cout, Ktot, 64)) return r;
    if (int r = get_map_2d(&tmBlo, wr_lo, d->Cout, Ktot, 64)) return r;
    TcArgs a{};
    a.bias = bias; a.dst = y;
    a.store_cols = d->Cout;
    a.stem4d = st; a.stem_tq = (d->Q + 15) / 16; a.stem_tp = (d->P + 7) / 8;
    a.M = (int64_t)d->N * a.stem_tq * a.stem_tp * TC_BM;        // every tile row is "valid": the TMA store clips the boxes
    a.Cout = d->Cout; a.P = d->P; a.Q = d->Q; a.stride = 1;
    a.taps_h = d->R; a.taps_w = 1; a.cblocks = 1; a.relu = (flags & I2V_EPI_RELU) ? 1 : 0;
    a.trace = g_trace; a.trace_tiles = g_trace_tiles;
    return tc_launch_persist<64, true, false, true, true>(tmA, tmBhi, tmBlo, tmOut, tmAodd, a, as_stream(stream));
}

// Strided data gradient, one stride-parity class per call.  Image rows h = stride*i + ph (columns likewise)
// receive only the filter taps r = r0 + stride*a with r0 = (ph + pad) mod stride, from dy row
// i + (ph + pad - r0)/stride - a: a dense, stride-1 problem over dy with A_h x A_w taps whose output rows are
// scattered with pitch `stride` into dx.  w_* = [Cin, (tap_h, tap_w, co)] with tap_h = A_h-1-a (host: see
// engine_native._class_weights).  Classes without taps (e.g. the odd rows of a 1x1/s2 conv) receive no gradient
// from this conv: the call is a no-op for them and the caller owns their contents (addend already in place).
extern "C" int i2v_conv_tc_dgrad_class_bits_f32(const i2v_conv_desc* d, int ph, int pw, const float* dy, const float* w_hi,
                                                const float* w_lo, const float* addend, const float* mask_src,
                                                const uint32_t* mask_bits, float* dx, i2v_stream_t stream);

extern "C" int i2v_conv_tc_dgrad_class_f32(const i2v_conv_desc* d, int ph, int pw, const float* dy, const float* w_hi,
                                           const float* w_lo, const float* addend, const float* mask_src, float* dx,
                                           i2v_stream_t stream) {
    return i2v_conv_tc_dgrad_class_bits_f32(d, ph, pw, dy, w_hi, w_lo, addend, mask_src, nullptr, dx, stream);
}

// ... with the ReLU-backward mask as BITS ([Cin/32][N*H*W] words, the forward epilogue's bits_out of the tensor dx belongs to)
// instead of the f32 activation: with mask_src == NULL the class runs the TMA epilogue (im2col-mode scatter store, addend by
// im2col-mode load) — no per-thread global access
extern "C" int i2v_conv_tc_dgrad_class_bits_f32(const i2v_conv_desc* d, int ph, int pw, const float* dy, const float* w_hi,
                                                const float* w_lo, const float* addend, const float* mask_src,
                                                const uint32_t* mask_bits, float* dx, i2v_stream_t stream) {
    I2V_REQUIRE(d && dy && dx, "null pointer");
    I2V_REQUIRE(!(mask_bits && mask_src), "pass either an f32 mask source or a bit mask, not both");
    I2V_REQUIRE(i2v_conv_tc_supported(d, 1), "shape not supported by the tensor-core path");
    const int st = d->stride;
    I2V_REQUIRE(ph >= 0 && ph < st && pw >= 0 && pw < st, "class index out of range");
    if (d->N == 0) return I2V_OK;
    const int r0 = (ph + d->pad) % st, s0 = (pw + d->pad) % st;
    const int Ah = r0 < d->R ? (d->R - 1 - r0) / st + 1 : 0;
    const int Aw = s0 < d->S ? (d->S - 1 - s0) / st + 1 : 0;
    const int Hc = ph < d->H ? (d->H - 1 - ph) / st + 1 : 0;
    const int Wc = pw < d->W ? (d->W - 1 - pw) / st + 1 : 0;
    if (Ah == 0 || Aw == 0 || Hc == 0 || Wc == 0) return I2V_OK;
    I2V_REQUIRE(w_hi, "null weight pointer");
    const int ch = (ph + d->pad - r0) / st, cw = (pw + d->pad - s0) / st;
    TcProblem pr{};
    pr.src = dy; pr.N = d->N; pr.H = d->P; pr.W = d->Q; pr.C = d->Cout;
    pr.P = Hc; pr.Q = Wc; pr.stride = 1;
    pr.lower_h = ch - (Ah - 1); pr.lower_w = cw - (Aw - 1);
    pr.upper_h = pr.lower_h + Hc - d->P; pr.upper_w = pr.lower_w + Wc - d->Q;
    pr.taps_h = Ah; pr.taps_w = Aw;
    pr.w_hi = w_hi; pr.w_lo = w_lo; pr.Cout = d->Cin;
    pr.bias = nullptr; pr.residual = addend; pr.mask_src = mask_src; pr.mask_bits = mask_bits; pr.dst = dx; pr.relu = 0;
    pr.out_s = st; pr.out_h0 = ph; pr.out_w0 = pw; pr.out_H = d->H; pr.out_W = d->W;
    return tc_run(pr, as_stream(stream));
}
