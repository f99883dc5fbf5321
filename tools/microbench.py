#!/usr/bin/env python
"""Isolated timing of the memory-bound kernels (K1, K3a, K3b, K3c) at config-2 size.

Each kernel is timed alone with CUDA events on the launching stream, 3 warm-ups, L2 flushed between
iterations by writing a 256 MB buffer (L2 is 126 MB); GB/s = ALGORITHMIC bytes / time, compared with the
measured HBM copy bandwidth in MEASURED_PEAKS.json (burst figure: these kernels are timed alone).
Prints one JSON object per kernel; `--out` also writes them to a file.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from i2v_b200 import capi  # noqa: E402


def time_kernel(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    times = []
    for _ in range(iters):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1)
        times.append(t)
        total += t
    times.sort()
    return times[len(times) // 2], times[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=512)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--out", default=None)
    ap.add_argument("--clusters", default="default,8,16,4")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    capi.device_check(dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    flush = torch.empty(64 * 1024 * 1024, device=dev)      # 256 MB
    N = args.frames
    results = []

    def report(name, nbytes, fn, extra=None):
        med, best = time_kernel(fn, args.iters, flush)
        r = {"kernel": name, "frames": N, "algorithmic_bytes": nbytes, "median_ms": med, "best_ms": best,
             "gbs_median": nbytes / med / 1e6, "gbs_best": nbytes / best / 1e6, "peak_gbs": peak,
             "frac_median": nbytes / med / 1e6 / peak}
        if extra:
            r.update(extra)
        results.append(r)
        print(json.dumps(r), flush=True)

    # K3a / K3b / compose at [N,3,224,224]
    shape = (N, 3, 224, 224)
    inner = 224 * 224
    n = N * 3 * inner
    g = torch.randn(shape, device=dev) * 1e-6
    x = torch.rand(shape, device=dev)
    m = torch.zeros(shape, device=dev)
    v = torch.zeros(shape, device=dev)
    mod = torch.full(shape, 0.01 / 255, device=dev)
    out = torch.empty(shape, device=dev)
    report("i2v_adam_compose_f32", 36 * n, lambda: capi.adam_compose(g, m, v, mod, x, out, 16 / 255, inner, 3, 0.005))
    report("i2v_compose_norm_f32", 12 * n, lambda: capi.compose_norm(x, mod, out, 16 / 255, inner))
    report("i2v_sign_step_project_f32", 16 * n, lambda: capi.sign_step_project(out, g, x, 16 / 2550, 16 / 255, inner))
    del m, v, mod
    B, T = N // 32, 32
    g5 = g.view(B, T, 3, 224, 224).permute(0, 2, 1, 3, 4).contiguous()
    x5 = x.view(B, T, 3, 224, 224).permute(0, 2, 1, 3, 4).contiguous()
    adv5 = out.view(B, T, 3, 224, 224).permute(0, 2, 1, 3, 4).contiguous()
    mom = torch.zeros_like(g5)
    norm = torch.empty(B, T, device=dev)

    def mi():
        capi.frame_absmean(g5, norm)
        capi.mi_sign_step_project(adv5, g5, mom, norm, x5, 1.0, 16 / 2550, 16 / 255)
    report("i2v_frame_absmean+mi_sign_step_project_f32", 28 * n, mi)
    del g, x, out, g5, x5, adv5, mom
    torch.cuda.empty_cache()

    # K1 at ResNet layer2 feature size, per cluster choice
    D = 512 * 28 * 28
    a = torch.relu(torch.randn(N, D, device=dev))
    b = torch.relu(torch.randn(N, D, device=dev))
    grad = torch.empty_like(a)
    cos = torch.empty(N, device=dev)
    for c in args.clusters.split(","):
        if c == "default":
            os.environ.pop("I2V_COS_CLUSTER", None)
        else:
            os.environ["I2V_COS_CLUSTER"] = c
        report("i2v_cosine_loss_grad_f32", 12 * N * D, lambda: capi.cosine_loss_grad(a, b, grad, cos), {"cluster": c})
    os.environ.pop("I2V_COS_CLUSTER", None)
    report("i2v_cosine_loss_grad_f32(loss only)", 8 * N * D, lambda: capi.cosine_loss_grad(a, b, None, cos))
    # a small-N case: one clip (config 1), 32 frames
    a32, b32, g32 = a[:32], b[:32], grad[:32]
    report("i2v_cosine_loss_grad_f32", 12 * 32 * D, lambda: capi.cosine_loss_grad(a32, b32, g32, cos[:32]), {"frames": 32})
    del a, b, grad, a32, b32, g32
    torch.cuda.empty_cache()

    # K3d (ILAF update), K8 (TemporalTranslation, 7 variants), K9 (ILAF layer loss), K6 (DR), max pooling of ResNet's stem
    clips = max(1, N // 32)
    shp = (clips, 3, 32, 224, 224)
    n5 = clips * 3 * 32 * 224 * 224
    inner5 = 32 * 224 * 224
    g = torch.randn(shp, device=dev)
    x = torch.rand(shp, device=dev)
    mod = torch.zeros(shp, device=dev)
    out = torch.empty(shp, device=dev)
    report("i2v_sign_descent_compose_f32", 20 * n5, lambda: capi.sign_descent_compose(g, mod, x, out, 16 / 255, 0.005, inner5))
    D7 = 7
    tclips = max(1, min(clips, 4))
    tshape = (tclips, 3, 32, 224, 224)
    tn = tclips * 3 * 32 * 224 * 224
    stack = torch.empty((D7,) + tshape, device=dev)
    adv = torch.randn(tshape, device=dev)
    moves = [-3, -2, -1, 0, 1, 2, 3]
    report("i2v_temporal_shift_stack_f32", 4 * tn * (1 + D7), lambda: capi.temporal_shift_stack(adv, stack, moves), {"clips": tclips})
    kern = [1.0 / D7] * D7
    tout = torch.empty(tshape, device=dev)
    report("i2v_temporal_combine_f32", 4 * tn * (1 + D7), lambda: capi.temporal_combine(stack, kern, moves, 0.5, tout), {"clips": tclips})
    del stack, adv, tout, g, x, mod, out
    torch.cuda.empty_cache()
    nf = N * 512 * 28 * 28 // 4
    f = torch.randn(nf, device=dev)
    o = torch.randn(nf, device=dev)
    d0 = torch.randn(nf, device=dev)
    d0 /= d0.norm()
    ws = capi.ila_workspace(dev)
    stats = torch.zeros(4, device=dev)
    gr = torch.empty(nf, device=dev)
    report("i2v_ila_loss_f32", 12 * nf, lambda: capi.ila_loss(f, o, d0, 1.0, ws, stats))
    report("i2v_ila_grad_f32", 16 * nf, lambda: capi.ila_grad(f, o, d0, gr, stats))
    acc = torch.zeros(2, device=dev, dtype=torch.float64)
    sw = capi.std_workspace(dev)
    report("i2v_std_accumulate_f32", 4 * nf, lambda: capi.std_accumulate(f, sw, acc))
    del f, o, d0, gr
    torch.cuda.empty_cache()
    pf = min(N, 256)
    xs = torch.relu(torch.randn(pf, 112, 112, 64, device=dev))
    ys = torch.empty(pf, 56, 56, 64, device=dev)
    am = torch.empty(pf, 56, 56, 64, device=dev, dtype=torch.uint8)
    report("i2v_maxpool_fwd_f32 (3,2,1) mark_dead", 4 * xs.numel() + 5 * ys.numel(), lambda: capi.maxpool_fwd(xs, ys, am, 3, 2, 1, mark_dead=True),
           {"frames": pf})
    dys = torch.randn_like(ys)
    dxs = torch.empty_like(xs)
    report("i2v_maxpool_bwd_f32 (3,2,1) no mask", 5 * ys.numel() + 4 * xs.numel(), lambda: capi.maxpool_bwd(dys, am, None, dxs, 3, 2, 1),
           {"frames": pf})
    if args.out:
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
