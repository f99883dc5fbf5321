#!/usr/bin/env python
"""Device time of the scratch-free first-layer data gradient with parts switched off ($I2V_STEM_DBG: 1 no MMA, 2 no col2im
math, 4 no A split, 8 no TMEM loads; results are wrong in these modes) next to the GEMM + col2im pair it replaces."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from i2v_b200 import capi
from i2v_b200.engine_native import _split_tf32
from tc_probe import timeit

n, H = int(os.environ.get("FRAMES", "256")), 224
P = 112
g = torch.Generator().manual_seed(1)
w = torch.randn(64, 3, 7, 7, generator=g) / 147 ** 0.5
w_stem = w.permute(1, 2, 3, 0).reshape(147, 64).contiguous().cuda()
hi, lo, _ = _split_tf32(capi.stem_direct_dgrad_weights(w_stem))
dy = torch.randn(n, P, P, 64, device="cuda")
dx = torch.empty(n, 3, H, H, device="cuda")
d = capi.ConvDesc(n, H, H, 3, 64, 7, 7, 2, 3, P, P)
out = {"dbg": int(os.environ.get("I2V_STEM_DBG", "0")), "frames": n}
out["direct_us"] = round(1e3 * timeit(lambda: capi.conv_stem_dgrad_direct(d, dy, hi, lo, dx), iters=3, reps=3), 1)
nz = capi.stem_dgrad_tc_rows(147)
wz = torch.cat([w_stem, w_stem.new_zeros(nz - 147, 64)], 0).contiguous()
h2, l2, _ = _split_tf32(wz)
z = torch.empty(capi.stem_dgrad_tc_scratch_floats(d), device="cuda")
out["gemm_col2im_us"] = round(1e3 * timeit(lambda: capi.conv_stem_dgrad_tc(d, dy, h2, l2, z, dx), iters=3, reps=3), 1)
print(json.dumps(out))
