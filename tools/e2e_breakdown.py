#!/usr/bin/env python
"""Where the end-to-end call spends its time beyond the K steps (run on the GPU box): events around the stages of one
attack(videos, labels, names) call, second call of the same shape (warm allocator, captured graph reused)."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from i2v_b200 import attack_loop, backbones, capi, synth
import image_attacks

K = int(os.environ.get("STEPS", "10"))
clips = int(os.environ.get("CLIPS", "16"))
dev = torch.device("cuda", 0)
capi.device_check(dev)
backbones.set_weight_policy("random", 0)
videos = torch.cat([synth.clip(i, b=1, f=32, h=224, w=224)[0] for i in range(clips)], 0).pin_memory()
atk = image_attacks.ImageGuidedFMDirection_Adam(["resnet50"], depth=2, step_size=0.005, epsilon=16 / 255, steps=K)
names = ["c%d" % i for i in range(clips)]
out_host = torch.empty(videos.shape).pin_memory()
adv = atk(videos, None, names); out_host.copy_(adv); torch.cuda.synchronize()

marks = []
def mark(name):
    e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((name, e, time.perf_counter()))
run = atk._run_cache["run"]
mark("start")
dv = videos.to(dev, non_blocking=True); mark("h2d")
run.setup(dv); mark("setup (denorm, clean features, compose)")
for _ in range(K):
    run.step()
mark("%d steps" % K)
res = run.finish(); mark("finish (clone, cost log D2H)")
if os.environ.get("D2H_SPLIT"):
    adv_c = res.adv.contiguous(); mark("permuting copy on the device")
    out_host.copy_(adv_c, non_blocking=True); mark("d2h (contiguous source)")
else:
    out_host.copy_(res.adv, non_blocking=True); mark("d2h")
torch.cuda.synchronize()
mark("sync")
out = {}
for (n0, e0, t0), (n1, e1, t1) in zip(marks, marks[1:]):
    out[n1] = {"gpu_ms": round(e0.elapsed_time(e1), 2), "host_ms": round(1e3 * (t1 - t0), 2)}
out["total_gpu_ms"] = round(marks[0][1].elapsed_time(marks[-1][1]), 2)
print(json.dumps(out, indent=1))
