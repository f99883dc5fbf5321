#!/usr/bin/env python
"""Which rows does tcgen05.mma read when a SWIZZLE_128B K-major A descriptor starts at a row that is not a multiple of 8?
(i2v_mma_shift_probe; run on the GPU box.)  Integer-valued operands: the expected product is exact."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from i2v_b200 import capi
dev = torch.device("cuda", 0)
capi.device_check(dev)
g = torch.Generator().manual_seed(0)
a = torch.randint(-8, 9, (256, 32), generator=g).float().to(dev)
b = torch.randint(-8, 9, (64, 32), generator=g).float().to(dev)
lib = capi.load()
for mode in (0, 1):
    for shift in (0, 8, 1, 2, 3, 5, 7, 9, 58, 59, 60, 116, 117, 118):
        out = torch.full((128, 64), float("nan"), device=dev)
        rc = lib.i2v_mma_shift_probe(shift, mode, a.data_ptr(), b.data_ptr(), out.data_ptr(), None)
        torch.cuda.synchronize()
        want = a[shift:shift + 128] @ b.t()
        ok = torch.equal(out, want)
        note = ""
        if not ok:
            # which source row did each output row come from?
            allp = a @ b.t()
            src = [(allp == out[r]).all(1).nonzero().flatten().tolist() for r in range(128)]
            note = " rows 0..11 read from %s" % ([s[0] if s else None for s in src[:12]],)
        print("mode %d shift %3d rc %d exact %s%s" % (mode, shift, rc, ok, note), flush=True)
