#!/usr/bin/env python
"""tcgen05.mma.kind::tf32 issue-rate table (run on the GPU box): cycles per M=128 x N x K=8 instruction for chains into
1 / 2 / 3 / 4 independent accumulators, A from shared memory or tensor memory, on 1 and 148 CTAs.  Floor = N/2 cycles."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from i2v_b200 import capi
import torch
capi.device_check(torch.device("cuda", 0))
count = 512
print("N  accs A     ctas issuers  issue/instr  done/instr(per issuer)  floor")
for ctas in (1, 148):
    for a_tmem in (0, 1):
        for N in (64, 128, 256):
            for accs, issuers in ((1, 1), (2, 1), (1, 2), (1, 3), (2, 2)):
                if accs * N > 384 or issuers * accs * N > 448:
                    continue
                capi.mma_probe(N, accs, a_tmem, count, ctas, issuers)
                iss, done = capi.mma_probe(N, accs, a_tmem, count, ctas, issuers)
                print("%-3d %d   %-5s %4d  %d   %10.1f  %10.1f  %6.1f" % (N, accs, "tmem" if a_tmem else "smem", ctas, issuers,
                                                                         iss / count, done / count, N / 2))
