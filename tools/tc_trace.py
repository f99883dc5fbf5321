#!/usr/bin/env python
"""Pipeline timeline of the persistent tensor-core convolution kernel (run on the GPU box): CTA 0 stamps clock64() at
8 points of each tile (i2v_conv_tc_set_trace); prints per tile, in microseconds relative to the first stamp:
  P0 producer starts tile | P1 last load issued | M2 MMA owns accumulator | M3 operands landed (+split) |
  M4 last MMA issued | E5 epilogue sees accumulator | E6 epilogue done | S7 split warps done"""
import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from i2v_b200 import capi
from i2v_b200.engine_native import _split_tf32
from tc_probe import LAYERS


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--layers", default="")
    ap.add_argument("--tiles", type=int, default=12)
    ap.add_argument("--mhz", type=float, default=1965.0)
    args = ap.parse_args()
    dev = "cuda"
    capi.device_check(torch.device(dev, 0))
    want = [s for s in args.layers.split(",") if s]
    for name, H, Cin, Cout, k, s, p, cnt in LAYERS:
        if want and not any(w in name for w in want):
            continue
        P = (H + 2 * p - k) // s + 1
        g = torch.Generator().manual_seed(1)
        w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
        hi, lo, rna = _split_tf32(w.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous().to(dev))
        n = args.frames
        xd = torch.randn(n, H, H, Cin, device=dev)
        y = torch.empty(n, P, P, Cout, device=dev)
        res = torch.randn_like(y) if name.endswith("+res") else None
        d = capi.ConvDesc(n, H, H, Cin, Cout, k, k, s, p, P, P)
        for mode, (a, b) in (("x3", (hi, lo)), ("x1", (rna, None))):
            for _ in range(3):
                capi.conv_tc(d, 0, xd, a, b, None, res, None, y, relu=True)
            buf = torch.zeros(args.tiles, 8, dtype=torch.int64, device=dev)
            capi.conv_tc_set_trace(buf, args.tiles)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            capi.conv_tc(d, 0, xd, a, b, None, res, None, y, relu=True)
            e1.record()
            torch.cuda.synchronize()
            capi.conv_tc_set_trace(None)
            t = buf.cpu().double()
            t0 = t[t > 0].min()
            us = (t - t0) / args.mhz
            print("== %s %s: kernel %.1f us (%d frames)" % (name, mode, 1e3 * e0.elapsed_time(e1), n))
            print("   tile   P0     P1     M2     M3     M4     E5     E6     S7")
            for i in range(args.tiles):
                if (t[i] > 0).any():
                    print("   %3d " % i + " ".join("%6.2f" % v if t[i, j] > 0 else "   -  " for j, v in enumerate(us[i])))


if __name__ == "__main__":
    main()
