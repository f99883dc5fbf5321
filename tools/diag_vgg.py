"""Diagnostic: per-buffer comparison of the native engine's backward between the CUDA-core and tensor-core paths."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from i2v_b200 import backbones
from i2v_b200.engine_native import NativeEngine
backbones.set_weight_policy("random", 0)
name, depth, side = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
g = torch.Generator().manual_seed(7)
img = torch.randn(3, 3, side, side, generator=g).cuda()
res = {}
for mode in ("simt", "tc"):
    eng = NativeEngine(backbones.get_model(name), name, depth, tf32x3=True, use_tensor_cores=(mode == "tc"))
    feats = eng.features(img, need_grad=True)
    plan = eng._last
    ups = [torch.randn(f.shape, generator=torch.Generator().manual_seed(3)).cuda() * (f > 0) for f in feats]
    gimg = eng.input_grad([u.clone() for u in ups]).clone()
    res[mode] = (dict((k, v.clone()) for k, v in plan["acts"].items()), dict((k, v.clone()) for k, v in plan["grads"].items()), gimg, [op.name for op in eng.ops])
a_s, g_s, gi_s, names = res["simt"]
a_t, g_t, gi_t, _ = res["tc"]
for k in a_s:
    d = (a_s[k] - a_t[k]).abs().max().item() / (a_s[k].abs().max().item() + 1e-30)
    print("act ", k, tuple(a_s[k].shape), "rel diff %.3e" % d)
for k in reversed(list(g_s)):
    d = (g_s[k] - g_t[k]).abs().max().item() / (g_s[k].abs().max().item() + 1e-30)
    print("grad", k, tuple(g_s[k].shape), "rel diff %.3e" % d)
print("gimg rel diff %.3e" % ((gi_s - gi_t).abs().max().item() / gi_s.abs().max().item()))
