#!/usr/bin/env python
"""What bounds each convolution shape: device time of the tensor-core kernel with parts switched off through
$I2V_TC_PAIR_DBG (bit 0 / 1: the two MMA issuers, bit 2: the A split, bit 3: the weight-tile loads; results are wrong in
these modes, only the time is read).  One process per mode (the switch is read once):

    for d in 0 3 4 8 15; do I2V_TC_PAIR=0 I2V_TC_PAIR_DBG=$d python tools/tc_limits.py --frames 256; done
"""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from i2v_b200 import capi
from i2v_b200.engine_native import _split_tf32
from tc_probe import LAYERS, timeit


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--only", default="", help="substring filter on the layer names")
    args = ap.parse_args()
    dev = "cuda"
    out = {"dbg": int(os.environ.get("I2V_TC_PAIR_DBG", "0")), "halo": os.environ.get("I2V_TC_HALO", ""), "halo_dbg": os.environ.get("I2V_TC_HALO_DBG", ""), "pair": os.environ.get("I2V_TC_PAIR", "-1"), "frames": args.frames, "us": {}}
    for name, H, Cin, Cout, k, s, p, cnt in LAYERS:
        if args.only not in name:
            continue
        P = (H + 2 * p - k) // s + 1
        g = torch.Generator().manual_seed(1)
        w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
        hi, lo, rna = _split_tf32(w.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous().to(dev))
        n = args.frames
        xd = torch.randn(n, H, H, Cin, device=dev)
        y = torch.empty(n, P, P, Cout, device=dev)
        res = torch.randn_like(y) if name.endswith("+res") else None
        bits = torch.empty(Cout // 32, n * P * P, device=dev, dtype=torch.int32)
        d = capi.ConvDesc(n, H, H, Cin, Cout, k, k, s, p, P, P)
        ms = timeit(lambda: capi.conv_tc(d, 0, xd, hi, lo, None, res, None, y, relu=True, mask_bits=bits), iters=3, reps=4)
        out["us"][name] = round(ms * 1e3, 1)
        del xd, y, res, bits
    print(json.dumps(out))


if __name__ == "__main__":
    main()
