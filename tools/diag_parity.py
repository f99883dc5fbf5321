"""Diagnostic: isolate which stage of one attack step deviates from float64 (run on the GPU box)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from i2v_b200 import attack_loop, backbones, capi, engines, synth
from oracle import loops as OL, oracle as O

backbones.set_weight_policy("random", 0)
backbones.ARCH_OVERRIDE.update({"resnet": "resnet50"})
name, depth = sys.argv[1] if len(sys.argv) > 1 else "resnet", 2
videos, _ = synth.clip(2, b=1, f=2, h=64, w=64)
for engine in (sys.argv[2:] or ["cudnn", "native"]):
    os.environ.pop("I2V_COS_CLUSTER", None)
    eng = engines.make_engine(backbones.get_model(name), name, depth, engine)
    run = attack_loop.ImageGuidedRun([eng], 16 / 255, 3, 0.005)
    run.setup(videos)
    init = [t.clone() for t in run.init_feats[0]]
    for step in range(3):
        ti = run.true_img.clone()
        feats = [f.clone() for f in eng.features(ti, need_grad=True)]
        a, b = feats[0], init[0]
        N = a.shape[0]
        c64, g64 = O.cosine_loss_grad_f64(a.cpu().numpy().reshape(N, -1), b.cpu().numpy().reshape(N, -1))
        out = {}
        for cl in ("default", "1", "2", "8"):
            if cl == "default": os.environ.pop("I2V_COS_CLUSTER", None)
            else: os.environ["I2V_COS_CLUSTER"] = cl
            ga = torch.empty_like(a); cos = torch.empty(N, device="cuda")
            capi.cosine_loss_grad(a, b, ga, cos, relu_mask=False)
            err = np.linalg.norm(ga.cpu().numpy().reshape(N, -1) - g64) / np.linalg.norm(g64)
            out[cl] = (float(err), float(np.abs(cos.cpu().numpy() - c64).max()))
        os.environ.pop("I2V_COS_CLUSTER", None)
        at = a.detach().clone().requires_grad_(True)
        torch.nn.functional.cosine_similarity(at.reshape(N, -1), b.reshape(N, -1)).sum().backward()
        e32 = np.linalg.norm(at.grad.cpu().numpy().reshape(N, -1) - g64) / np.linalg.norm(g64)
        hm = OL.HookedModel(backbones.seeded_random_init(backbones.arch_of(name), 0).double(), backbones.family_of(name), depth)
        with torch.no_grad():
            f64 = hm.run(ti.cpu().double())[0]
        fa = a.permute(0, 3, 1, 2) if engine.startswith("native") else a
        ferr = float((fa.cpu().double() - f64).abs().max() / f64.abs().max())
        print(engine, "step", step, "K1 relL2 per cluster", out, "torch-f32 K1", float(e32), "feature err", ferr, "cos", c64, flush=True)
        run.step()
