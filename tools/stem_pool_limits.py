"""Timing experiments on i2v_conv_stem_dgrad_pool_f32 with parts of the kernel switched off ($I2V_STEM_DBG, re-read per call by the pooled launcher; results are wrong, only the time is read): 1 no MMA, 2 no col2im math, 4 no assemble at all, 8 no TMEM loads in the
epilogue, 16 no gradient gather, 32 no argmax compare, 64 no proxy fence, 128 no tensor-memory
store of a_lo, 256 no shared-memory store of a_hi, 512 no argmax loads.  Flags on the command line; one JSON line each.  Prints one JSON line."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from i2v_b200 import capi  # noqa: E402
from i2v_b200.engine_native import _split_tf32  # noqa: E402

n, H, W = 256, 224, 224
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
P = Q = 112
P2 = Q2 = 56
d = capi.ConvDesc(n, H, W, 3, 64, 7, 7, 2, 3, P, Q)
w = torch.randn(147, 64, generator=g).to(dev) / 12.0
hi, lo, _ = _split_tf32(capi.stem_direct_dgrad_weights(w))
act = torch.relu(torch.randn(n, P, Q, 64, device=dev))
pooled = torch.empty(n, P2, Q2, 64, device=dev)
am = torch.empty(n, P2, Q2, 64, device=dev, dtype=torch.uint8)
capi.maxpool_fwd(act, pooled, am, 3, 2, 1, mark_dead=True)
gp = torch.randn(n, P2, Q2, 64, device=dev)
gact = torch.empty(n, P, Q, 64, device=dev)
dx = torch.empty(n, 3, H, W, device=dev)


def timed(fn, it=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / it


flags = [int(v) for v in (sys.argv[1:] or [os.environ.get("I2V_STEM_DBG", "0")])]
for f in flags:
    os.environ["I2V_STEM_DBG"] = str(f)          # the pooled launcher re-reads it on every call
    print(json.dumps({"dbg": f, "pool_fused_us": round(timed(lambda: capi.conv_stem_dgrad_pool(d, gp, am, hi, lo, dx)), 1)}), flush=True)
os.environ["I2V_STEM_DBG"] = "0"
print(json.dumps({"maxpool_bwd_us": timed(lambda: capi.maxpool_bwd(gp, am, None, gact, 3, 2, 1)),
                  "stem_dgrad_direct_us": timed(lambda: capi.conv_stem_dgrad_direct(d, gact, hi, lo, dx))}))
