#!/usr/bin/env python
"""The reference's I2V loop as image_attacks.py:294-364 runs it ON A GPU, restated with torch ops only — comparison point
(i) of BASELINE.md section 4 and the "reference on its own platform" yardstick of the parity reports.  No kernel of this
repo is on the path: torchvision ResNet-50 (seeded random init, FULL forward incl. layer3/4, avgpool, fc), forward hook on
layer2[-1], F.cosine_similarity, `cost.backward()` with every weight requiring grad, torch.optim.Adam on the modifier
(CUDA foreach path), one host sync per step (the reference prints the cost).

As a tool: runs 60 steps on the config-1 fixture's clip and scores the result against tests/golden/
i2v_resnet50_d2_224_60step.npz (the same class run on the CPU) exactly as the GPU parity test scores the native engine:

    python tools/reference_on_gpu.py > gpurun_out/reference_gpu_vs_cpu_fixture.json
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]


def run(videos, steps, step_size, eps=16 / 255, tf32=False, timed_from=2, keep_first_grad=False):
    """videos [b,3,f,h,w] on a CUDA device.  Returns dict(adv, cost [steps], ms_per_step, first_grad)."""
    from i2v_b200 import backbones
    dev = videos.device
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    try:
        b, c, f, h, w = videos.shape
        mean = torch.tensor(MEAN, device=dev).view(1, 3, 1, 1)
        std = torch.tensor(STD, device=dev).view(1, 3, 1, 1)
        model = backbones.seeded_random_init("resnet50", 0).to(dev)
        model.train()                                                        # image_attacks.py:253-256
        for m in model.modules():
            if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm1d)):
                m.eval()
        acts = []
        model.layer2[-1].register_forward_hook(lambda mod, i, o: acts.append(o))
        image_inps = videos.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)   # 300-301
        modifier = torch.nn.Parameter(torch.full((b * f, c, h, w), 0.01 / 255, device=dev))   # 304-305
        opt = torch.optim.Adam([modifier], lr=step_size)                     # 306
        unnorm = (image_inps * std + mean).detach()                          # 308
        model(image_inps)                                                    # 318
        init = acts.pop().detach()
        costs, first_grad, t0 = [], None, None
        for i in range(steps):
            if i == timed_from:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            del acts[:]
            true_image = torch.clamp(unnorm + torch.clamp(modifier, min=-eps, max=eps), min=0, max=1)   # 331
            true_image = (true_image - mean) / std                           # 332
            model(true_image)                                                # 334
            cost = torch.sum(F.cosine_similarity(acts[0].view(b * f, -1), init.view(b * f, -1)))      # 341-347
            opt.zero_grad()
            cost.backward()                                                  # 352
            if i == 0 and keep_first_grad:
                first_grad = modifier.grad.detach().clone()
            opt.step()                                                       # 353
            costs.append(float(cost.detach().cpu()))                         # 349: the reference prints the cost
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / max(1, steps - timed_from) if t0 is not None else None
        with torch.no_grad():
            true_image = torch.clamp(unnorm + torch.clamp(modifier, min=-eps, max=eps), min=0, max=1)   # 360
            adv = ((true_image - mean) / std).reshape(b, f, c, h, w).permute(0, 2, 1, 3, 4)             # 361-363
        return {"adv": adv, "cost": np.array(costs, np.float32), "ms_per_step": ms, "first_grad": first_grad}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


def score_against_fixture(adv_minus_videos, cost, g_mod_first, g):
    """The figures tests/test_gpu_attacks.py::test_config1_60_steps_vs_reference_fixture records."""
    d = np.abs(adv_minus_videos - g["delta16"].astype(np.float32))
    n = int(np.prod(g["g_first_shape"]))
    gm = g_mod_first.reshape(-1)
    out = {"cost_rel_err_max": float(np.abs(cost / g["cost"] - 1).max()),
           "cost_rel_err_by_step": [float(v) for v in np.abs(cost / g["cost"] - 1)],
           "final_frac_within_1_255": float((d <= (1 / 255) / 0.225).mean()), "final_max_abs": float(d.max())}
    big = np.unpackbits(g["g_first_big_bits"])[:n].astype(bool)
    pos, neg = np.unpackbits(g["g_first_pos_bits"])[:n].astype(bool), np.unpackbits(g["g_first_neg_bits"])[:n].astype(bool)
    out["step1_sign_vs_reference_big"] = float((((gm > 0) == pos) & ((gm < 0) == neg))[big].mean())
    if "g64_first_pos_bits" in g:
        big64 = np.unpackbits(g["g64_first_big_bits"])[:n].astype(bool)
        pos64 = np.unpackbits(g["g64_first_pos_bits"])[:n].astype(bool)
        out["step1_sign_vs_float64_big"] = float(((gm > 0) == pos64)[big64].mean())
        out["reference_cpu_step1_sign_vs_float64_big"] = float(g["ref_vs_f64_sign_agreement_big"])
    return out


def main():
    from i2v_b200 import synth
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "i2v_resnet50_d2_224_60step.npz")))
    videos, _ = synth.clip(0, b=1, f=int(g["frames"]), h=int(g["side"]), w=int(g["side"]))
    out = {"what": "the reference loop on CUDA (torch/cuDNN, restated in tools/reference_on_gpu.py) against the same class on "
                   "the CPU (fixture), 32 frames x 3x224x224, 60 steps"}
    for tag, tf32 in (("fp32", False), ("tf32_torch_default_for_cudnn", True)):
        r = run(videos.cuda(), int(g["steps"]), float(g["step_size"]), tf32=tf32, keep_first_grad=True)
        sc = score_against_fixture((r["adv"].cpu() - videos).numpy(), r["cost"], r["first_grad"].cpu().numpy(), g)
        sc["ms_per_step"] = r["ms_per_step"]
        out[tag] = sc
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
