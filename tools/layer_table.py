#!/usr/bin/env python
"""Per-layer table: this repo's convolution kernels against cuDNN on the SAME layer shapes (SURVEY.md 7.2 step 5,
BASELINE.md section 4): every convolution of ResNet-50 up to layer2 at a 256-frame chunk of 224 x 224 frames, forward and
data gradient.

  native : timed IN SITU — one attack step's forward + backward through NativeEngine with CUDA events around every launch
           (capi.PROFILE_EVENTS), FP32-parity mode (3xTF32), epilogues (bias, ReLU, residual, masks) included;
  cuDNN  : torch.ops.aten.convolution / convolution_backward(output_mask = input only) on random tensors of the layer's
           shape, best layout of NCHW / channels_last, `cudnn.allow_tf32` False (the parity-equivalent arm) and True;
           conv ONLY — cuDNN's number excludes the BN / ReLU / residual kernels eager torch adds on top, so the comparison
           favours cuDNN.

    python tools/layer_table.py [--frames 256] > gpurun_out/layer_table.json
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from i2v_b200 import backbones, capi, engines   # noqa: E402


def time_ms(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def cudnn_times(n, cin, cout, h, w, k, s, p):
    out = {}
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        for cl in (False, True):
            fmt = torch.channels_last if cl else torch.contiguous_format
            x = torch.randn(n, cin, h, w, device="cuda").contiguous(memory_format=fmt)
            wt = torch.randn(cout, cin, k, k, device="cuda").contiguous(memory_format=fmt)
            y = torch.ops.aten.convolution(x, wt, None, [s, s], [p, p], [1, 1], False, [0, 0], 1)
            dy = torch.randn_like(y)
            tag = ("tf32" if tf32 else "fp32") + ("_cl" if cl else "_nchw")
            out["fwd_" + tag] = time_ms(lambda: torch.ops.aten.convolution(x, wt, None, [s, s], [p, p], [1, 1], False, [0, 0], 1))
            out["dgrad_" + tag] = time_ms(lambda: torch.ops.aten.convolution_backward(
                dy, x, wt, None, [s, s], [p, p], [1, 1], False, [0, 0], 1, [True, False, False]))
            del x, wt, y, dy
    torch.backends.cudnn.allow_tf32 = False
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--model", default="resnet50")
    ap.add_argument("--depth", type=int, default=2)
    args = ap.parse_args()
    backbones.set_weight_policy("random", 0)
    torch.backends.cudnn.benchmark = True
    eng = engines.make_engine(backbones.get_model(args.model), args.model, args.depth, "native")
    img = torch.randn(args.frames, 3, 224, 224, device="cuda")
    native = {}
    for it in range(3):
        capi.PROFILE_EVENTS = [] if it == 2 else None
        feats = eng.features(img, need_grad=True)
        eng.input_grad([torch.randn_like(f) * (f > 0) for f in feats])
        torch.cuda.synchronize()
    events, capi.PROFILE_EVENTS = capi.PROFILE_EVENTS, None
    stem = {}
    for name, e0, e1, nb, fl, detail in events:
        ms = e0.elapsed_time(e1)
        if detail is not None:
            d = native.setdefault(detail.split(" +")[0], {"ms": 0.0, "launches": 0, "variants": set()})
            d["ms"] += ms; d["launches"] += 1; d["variants"].add(detail)
        elif "stem" in name or "class" in name:
            d = stem.setdefault(name, {"ms": 0.0, "launches": 0})
            d["ms"] += ms; d["launches"] += 1
    plan = eng._last_fwd
    rows = []
    seen = {}
    for op in eng.ops:
        if op.kind != "conv":
            continue
        d = plan["descs"][op.name]
        key = (d.H, d.W, d.Cin, d.Cout, d.R, d.stride, d.pad)
        seen.setdefault(key, []).append(op.name)
    for key, names in seen.items():
        H, W, Cin, Cout, R, s, p = key
        cud = cudnn_times(args.frames, Cin, Cout, H, W, R, s, p)
        row = {"shape": "%dx%d %d->%d k%ds%d" % (H, W, Cin, Cout, R, s), "layers": names, "cudnn_ms": cud}
        for direction in ("fwd", "dgrad"):
            tag = "%s %dx%d %d->%d k%ds%d" % (direction, H, W, Cin, Cout, R, s)
            nat = native.get(tag)
            best32 = min(v for k, v in cud.items() if k.startswith(direction + "_fp32"))
            best_tf = min(v for k, v in cud.items() if k.startswith(direction + "_tf32"))
            if nat is not None:
                ms = nat["ms"] / nat["launches"]
                row[direction] = {"native_ms": ms, "cudnn_fp32_ms": best32, "cudnn_tf32_ms": best_tf,
                                  "native_over_cudnn_fp32": best32 / ms, "native_over_cudnn_tf32": best_tf / ms,
                                  "native_variants": sorted(nat["variants"])}
            else:
                row[direction] = {"native_ms": None, "cudnn_fp32_ms": best32, "cudnn_tf32_ms": best_tf,
                                  "note": "native path for this direction: see `other_native_kernels` (first layer: im2col/col2im "
                                          "+ GEMM; strided data gradient: stride-parity class launches)"}
        rows.append(row)
    out = {"frames": args.frames, "model": args.model, "depth": args.depth, "mode": "native = FP32 parity (3xTF32), in situ",
           "rows": rows, "other_native_kernels": {k: {"ms_per_step": v["ms"], "launches": v["launches"]} for k, v in stem.items()},
           "gpu": torch.cuda.get_device_name(0)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
