/* i2v_b200_debug.h — measurement hooks of libi2v_b200.so used by tools/ (tools/mma_probe.py, tools/tc_trace.py,
 * tools/stem_direct_trace.py).  NOT part of the product ABI of include/i2v_b200.h: nothing on the attack path calls
 * them, a binding of the reference does not need them, and they may change without notice.                        */
#ifndef I2V_B200_DEBUG_H_
#define I2V_B200_DEBUG_H_
#include "i2v_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Debug / measurement: issue-rate probe of tcgen05.mma.kind::tf32 (M = 128, K = 8): `count` MMAs of width N into
 * `accs` round-robin accumulators with A from shared memory (a_tmem = 0) or tensor memory, on `ctas` CTAs;
 * out[0] = cycles to issue, out[1] = cycles until completion (CTA 0).  tools/mma_probe.py prints the table.   */
int i2v_mma_probe(int N, int accs, int a_tmem, int count, int ctas, int issuers /* 1..3 concurrently issuing warps */,
                  long long* out, i2v_stream_t stream);

/* Debug: D[128 x 64] = A[shift .. shift+127] B^T with a SWIZZLE_128B A descriptor whose start is shifted by `shift_rows`
 * 128-byte rows (a = [256][32], b = [64][32], out = [128][64] device f32); base_offset_mode 1 also sets the descriptor's
 * base-offset field to (start >> 7) & 7.  tools/mma_shift_probe.py reports which rows the tensor core actually read.  */
int i2v_mma_shift_probe(int shift_rows, int base_offset_mode, const float* a, const float* b, float* out, i2v_stream_t stream);

/* Debug: subsequent tensor-core launches make CTA 0 stamp clock64() at 8 pipeline points of each of its first
 * `tiles` tiles into device_buf[tiles][8] (see TcArgs::trace in csrc/conv_tc.cu); NULL switches it off.     */
int i2v_conv_tc_set_trace(unsigned long long* device_buf, int tiles);

/* Tuning: tiles of at least `min_ksteps` k-steps (taps x Cin/32) of the 3xTF32 / TMA-epilogue configuration run on the
 * CTA-pair kernel (tcgen05.mma.cta_group::2, half of the weight rows per CTA; csrc/conv_tc.cu: conv_tc_pair_kernel);
 * 0 = never; -1 (the default) = only the convolutions that stream a residual / addend through the epilogue, from 4
 * k-steps — where the pair was measured to win in the attack step.  Initial value: $I2V_TC_PAIR.                  */
int i2v_conv_tc_set_pair_minkit(int min_ksteps);

/* Tuning: which 3x3 / stride-1 / pad-1 convolutions of the 3xTF32 / TMA-epilogue configuration run on the patch-once kernel
 * (csrc/conv_tc.cu: conv3x3_halo_kernel): 1 = wherever it fits, 2 = likewise but always with 64-channel tiles, 0 = never,
 * -1 (the default) = the 64-channel tiles only, where it was measured to win.  Initial value: $I2V_TC_HALO.            */
int i2v_conv_tc_set_halo_mode(int mode);

#ifdef __cplusplus
}
#endif
#endif /* I2V_B200_DEBUG_H_ */
