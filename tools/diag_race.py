"""Diagnostic: repeatability of the native engine (tensor-core mode) — same inputs, R repetitions; reports for every
activation / gradient buffer how many repetitions differ bitwise from repetition 0 and the error vs the CUDA-core path."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from i2v_b200 import backbones
from i2v_b200.engine_native import NativeEngine
backbones.set_weight_policy("random", 0)
name, depth, side, frames, reps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
g = torch.Generator().manual_seed(7)
img = torch.randn(frames, 3, side, side, generator=g).cuda()


def run(mode):
    eng = NativeEngine(backbones.get_model(name), name, depth, tf32x3=True, use_tensor_cores=(mode == "tc"))
    out = []
    for r in range(reps if mode == "tc" else 1):
        feats = eng.features(img, need_grad=True)
        plan = eng._last
        ups = [torch.randn(f.shape, generator=torch.Generator().manual_seed(3)).cuda() * (f > 0) for f in feats]
        gimg = eng.input_grad([u.clone() for u in ups]).clone()
        torch.cuda.synchronize()
        out.append((dict((k, v.clone()) for k, v in plan["acts"].items()),
                    dict((k, v.clone()) for k, v in plan["grads"].items()), gimg))
    return out, [op.name for op in eng.ops]


simt, names = run("simt")
tc, _ = run("tc")
a_s, g_s, gi_s = simt[0]
bad = 0
for kind, idx in (("act", 0), ("grad", 1)):
    keys = list(tc[0][idx]) if kind == "act" else list(reversed(list(tc[0][idx])))
    for k in keys:
        ref = simt[0][idx][k]
        errs = [(t[idx][k] - ref).abs().max().item() / (ref.abs().max().item() + 1e-30) for t in tc]
        ndiff = sum(1 for t in tc[1:] if not torch.equal(t[idx][k], tc[0][idx][k]))
        flag = "" if max(errs) < 1e-4 and ndiff == 0 else "   <<<<"
        bad += bool(flag)
        print("%-4s %-16s %-22s vs simt max %.2e min %.2e  reps differing bitwise from rep0: %d/%d%s"
              % (kind, k, tuple(ref.shape), max(errs), min(errs), ndiff, len(tc) - 1, flag))
errs = [(t[2] - gi_s).abs().max().item() / gi_s.abs().max().item() for t in tc]
print("gimg vs simt max %.2e min %.2e; bitwise differing reps %d" % (max(errs), min(errs), sum(1 for t in tc[1:] if not torch.equal(t[2], tc[0][2]))))
print("SUSPECT BUFFERS:", bad)
