#!/usr/bin/env python
"""Device time of the dual-source 1x1 launch of layer1.0 (64 + 64 -> 256 at 56x56: i2v_conv_tc_dual_f32) next to the same GEMM
from ONE source (128 -> 256) and its pipeline trace (run on the GPU box)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from i2v_b200 import capi
from i2v_b200.engine_native import _split_tf32
from tc_probe import timeit

n, H = int(os.environ.get("FRAMES", "256")), 56
dev = "cuda"
capi.device_check(torch.device(dev, 0))
g = torch.Generator().manual_seed(1)
w = torch.randn(256, 128, generator=g) / 128 ** 0.5
hi, lo, _ = _split_tf32(w.to(dev))
x = torch.randn(n, H, H, 64, device=dev)
t = torch.randn(n, H, H, 64, device=dev)
xt = torch.randn(n, H, H, 128, device=dev)
y = torch.empty(n, H, H, 256, device=dev)
bias = torch.zeros(256, device=dev)
bits = torch.empty(8, n * H * H, device=dev, dtype=torch.int32)
d2 = capi.ConvDesc(n, H, H, 64, 256, 1, 1, 1, 0, H, H)
d1 = capi.ConvDesc(n, H, H, 128, 256, 1, 1, 1, 0, H, H)
out = {"env": {k: v for k, v in os.environ.items() if k.startswith("I2V_")}}
out["dual_bits_us"] = round(1e3 * timeit(lambda: capi.conv_tc_dual(d2, x, t, hi, lo, bias, y, relu=True, mask_bits=bits), iters=3, reps=4), 1)
out["dual_us"] = round(1e3 * timeit(lambda: capi.conv_tc_dual(d2, x, t, hi, lo, bias, y, relu=True), iters=3, reps=4), 1)
out["single_bits_us"] = round(1e3 * timeit(lambda: capi.conv_tc(d1, 0, xt, hi, lo, bias, None, None, y, relu=True, mask_bits=bits), iters=3, reps=4), 1)
out["single_us"] = round(1e3 * timeit(lambda: capi.conv_tc(d1, 0, xt, hi, lo, bias, None, None, y, relu=True), iters=3, reps=4), 1)
print(json.dumps(out))
if os.environ.get("TRACE"):
    buf = torch.zeros(10, 8, dtype=torch.int64, device=dev)
    capi.conv_tc_set_trace(buf, 10)
    capi.conv_tc_dual(d2, x, t, hi, lo, bias, y, relu=True, mask_bits=bits)
    torch.cuda.synchronize()
    capi.conv_tc_set_trace(None)
    tt = buf.cpu().double(); t0 = tt[tt > 0].min(); us = (tt - t0) / 1965.0
    print("   tile   P0     P1     M2     M3     M4     E5     E6     S7")
    for i in range(10):
        print("   %3d " % i + " ".join("%6.2f" % v if tt[i, j] > 0 else "   -  " for j, v in enumerate(us[i])))
