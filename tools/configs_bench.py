#!/usr/bin/env python
"""Throughput of BASELINE.json's configs[2..4] through the PUBLIC drop-in classes (configs[0] / [1] are bench.py's
`--impl reference` and default lines).  One JSON object per config on stdout, all of them in --out.

  config 3  AENS_I2V_MF over resnet50 / vgg16 / densenet121 / squeezenet1_1, depths [2,3] each (BASELINE.json's ensemble;
            the DenseNet hook is this repo's extension), one 32-frame 224^2 clip per call.  Single process = all four
            backbones on this GPU; under torchrun with world % 4 == 0, one backbone per GPU (placement='ensemble').
  config 4  Kinetics-val-shaped sweep: independent batch-size-1 calls of the config-1 attack (I2V ResNet-50 layer2,
            60 steps), clips dealt round-robin over the ranks (no collective); `--clips` clips in total.
  config 5  UCF-101 shape (16 clips of 16 x 3 x 112 x 112): I2V (config-1 attack) vs BIM / MIFGSM (10 steps) on a
            random-init torchvision r3d_18 white-box stand-in (gluoncv video models are not installable offline).

Timing: CUDA events around whole attack calls after one warm-up call, inputs resident on the device; frame-steps =
clips x frames x steps.  Synthetic clips (i2v_b200.synth), seeded random-init weights.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from i2v_b200 import backbones, dist as D, synth   # noqa: E402
import base_attacks                                 # noqa: E402
import image_attacks                                # noqa: E402
import TPAMI_attack                                 # noqa: E402


def timed(fn, reps):
    fn()                                                # warm-up (allocator, tensor maps, cuDNN heuristics)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 1000.0 / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=6, help="config 4: clips in the sweep (the full sweep is 400)")
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--only", default="3,4,5")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        rank, local_rank, world = D.init_from_env()
    else:
        rank, local_rank = 0, 0
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    backbones.set_weight_policy("random", 0)
    backbones.ARCH_OVERRIDE.update({"resnet": "resnet50"})
    only = set(args.only.split(","))
    results = []

    def emit(rec):
        if world > 1:
            t = torch.tensor([rec["seconds"]], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            rec["seconds"] = float(t)
        rec["frame_steps_per_sec"] = rec["frame_steps"] / rec["seconds"]
        rec["n_gpus"] = world
        if rank == 0:
            print(json.dumps(rec), flush=True)
            results.append(rec)

    if "3" in only:
        names = ["resnet", "vgg", "densenet121", "squeezenet"]          # BASELINE.json configs[2] (native DenseNet engine)
        ens = world > 1 and world % 4 == 0
        atk = TPAMI_attack.AENS_I2V_MF(names, {n: [2, 3] for n in names}, 0.005, momentum=0.5, steps=args.steps,
                                       placement="ensemble" if ens else None)
        v, lab = synth.clip(0, b=1, f=32, h=224, w=224)
        v = v.to(dev)
        sec = timed(lambda: atk(v, lab, ["c"]), 2)
        emit({"config": 3, "workload": "AENS_I2V_MF resnet50/vgg16/densenet121/squeezenet1_1 depths [2,3], 1 clip x 32 x 3x224x224, "
              "%d steps, %s" % (args.steps, "one backbone per GPU (NCCL all-reduce of dcost/dimage)" if ens else
                                "all backbones on one GPU"),
              "frame_steps": 32 * args.steps * (world // 4 if ens else 1), "seconds": sec})
        del atk
        torch.cuda.empty_cache()

    if "4" in only:
        atk = image_attacks.ImageGuidedFMDirection_Adam(["resnet"], depth=2, step_size=0.005, steps=args.steps)
        mine = D.clip_shard(args.clips, rank, world)
        clips = [synth.clip(1000 + i, b=1, f=32, h=224, w=224)[0].to(dev) for i in mine]
        lab = torch.zeros(1, dtype=torch.long)

        def sweep():
            for i, c in zip(mine, clips):
                atk(c, lab, ["clip%d" % i])
        sec = timed(sweep, 1)
        emit({"config": 4, "workload": "I2V ResNet-50 layer2, %d independent batch-size-1 calls of 32 x 3x224x224, %d steps, "
              "clips round-robin over ranks, no collective" % (args.clips, args.steps),
              "frame_steps": args.clips * 32 * args.steps, "seconds": sec})
        del atk, clips
        torch.cuda.empty_cache()

    if "5" in only:
        import torchvision
        b, f, side = 16, 16, 112
        v, lab = synth.clip(2000 + rank, b=b, f=f, h=side, w=side, num_classes=101)
        v = v.to(dev)
        i2v = image_attacks.ImageGuidedFMDirection_Adam(["resnet"], depth=2, step_size=0.005, steps=args.steps)
        sec = timed(lambda: i2v(v, lab, ["c%d" % k for k in range(b)]), 2)
        emit({"config": 5, "attack": "I2V", "workload": "I2V ResNet-50 layer2, 16 clips x 16 x 3x112x112 per GPU, %d steps" % args.steps,
              "frame_steps": b * f * args.steps * world, "seconds": sec})
        del i2v
        torch.manual_seed(0)
        wb = torchvision.models.video.r3d_18(weights=None, num_classes=101).to(dev).eval()
        for name, cls in (("BIM", base_attacks.BIM), ("MIFGSM", base_attacks.MIFGSM)):
            atk = cls(wb, steps=10)
            sec = timed(lambda: atk(v.clone(), lab), 2)
            emit({"config": 5, "attack": name, "workload": "%s on random-init r3d_18 (white-box stand-in; cuDNN, TF32 off), 16 clips x 16 x "
                  "3x112x112 per GPU, 10 steps" % name, "frame_steps": b * f * 10 * world, "seconds": sec})

    if rank == 0 and args.out:
        with open(args.out, "w") as fh:
            json.dump(results, fh, indent=1)


if __name__ == "__main__":
    main()
