#!/usr/bin/env python
"""Pipeline timeline (clock64 stamps of CTA 0, see tools/tc_trace.py) of the two first-layer forward variants at 224^2:
the im2col + GEMM path (its GEMM launch) and the experimental direct path."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from i2v_b200 import capi
from i2v_b200.engine_native import _split_tf32

MHZ = float(os.environ.get("SM_MHZ", "1965"))
COLS = "   tile   P0     P1     M2     M3     M4     E5     E6     S7"


def show(title, buf, ms):
    t = buf.cpu().double()
    t0 = t[t > 0].min()
    us = (t - t0) / MHZ
    print("== %s: %.1f us" % (title, ms * 1000))
    print(COLS)
    for i in range(us.shape[0]):
        print("  %4d " % i + " ".join("%6.2f" % v if t[i, j] > 0 else "   -  " for j, v in enumerate(us[i].tolist())))


def main():
    dev = "cuda"
    capi.device_check(torch.device(dev, 0))
    n, H, k, s, p, Cout = int(os.environ.get("FRAMES", "64")), 224, 7, 2, 3, 64
    P = (H + 2 * p - k) // s + 1
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, 3, H, H, generator=g).to(dev)
    w = torch.randn(Cout, 3, k, k, generator=g) / (3 * k * k) ** 0.5
    d = capi.ConvDesc(n, H, H, 3, Cout, k, k, s, p, P, P)
    y = torch.empty(n, P, P, Cout, device=dev)
    K = 3 * k * k
    kp = (K + 31) // 32 * 32
    wk = torch.cat([w.reshape(Cout, K), torch.zeros(Cout, kp - K)], 1).contiguous().to(dev)
    hi, lo, _ = _split_tf32(wk)
    col = torch.empty(capi.stem_fwd_tc_scratch_floats(d), device=dev)
    wr = torch.zeros(Cout, k, 8, 4)
    wr[:, :, :k, :3] = w.permute(0, 2, 3, 1)
    dhi, dlo, _ = _split_tf32(wr.reshape(Cout, k * 32).contiguous().to(dev))
    xp = torch.empty(capi.stem_fwd_direct_scratch_floats(d), device=dev)
    tiles = 10
    for title, fn in (("im2col + GEMM", lambda: capi.conv_stem_fwd_tc(d, x, hi, lo, None, col, y, relu=True)),
                      ("direct", lambda: capi.conv_stem_fwd_direct(d, x, dhi, dlo, None, xp, y, relu=True))):
        for _ in range(3):
            fn()
        buf = torch.zeros(tiles, 8, dtype=torch.int64, device=dev)
        capi.conv_tc_set_trace(buf, tiles)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        capi.conv_tc_set_trace(None)
        show("%s, %d frames (whole entry point)" % (title, n), buf, e0.elapsed_time(e1))


if __name__ == "__main__":
    main()
