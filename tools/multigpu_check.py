#!/usr/bin/env python
"""Multi-GPU checks on a real box (run under torchrun, >= 2 ranks, NCCL):

  1. clip sharding: every rank attacks its own clips with no collective; rank 0 re-runs one of rank 1's clips
     and must get the bit-identical adversarial clip (clips are independent units);
  2. one backbone per GPU (placement='ensemble'): ENS-I2V and AENS-I2V with the members dealt over the ranks
     and the per-step all-reduce of dcost/dtrue_image + cosine rows, against the single-process ensemble —
     costs / layer weights must agree to 1e-5, the eps-ball must hold, every rank must hold the same result.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from i2v_b200 import backbones, dist as D, synth   # noqa: E402
import image_attacks                                # noqa: E402
import TPAMI_attack                                 # noqa: E402


def main():
    rank, local_rank, world = D.init_from_env()
    assert world >= 2, "run under torchrun with >= 2 ranks"
    dev = torch.device("cuda", local_rank)
    backbones.set_weight_policy("random", 0)
    out = {"world": world}
    labels = torch.zeros(1, dtype=torch.long)

    # ---- 1. clip sharding --------------------------------------------------------------------------
    atk = image_attacks.ImageGuidedFMDirection_Adam(["resnet50"], depth=2, step_size=0.005, steps=3)
    mine = D.clip_shard(4, rank, world)
    advs = {}
    for i in mine:
        v, _ = synth.clip(i, b=1, f=2, h=64, w=64)
        advs[i] = atk(v, labels, ["clip%d" % i]).contiguous()
    probe = D.clip_shard(4, 1, world)[0]                    # a clip rank 1 owns
    buf = advs[probe].clone() if rank == 1 else torch.empty(1, 3, 2, 64, 64, device=dev)
    dist.broadcast(buf, src=1)
    if rank == 0:
        v, _ = synth.clip(probe, b=1, f=2, h=64, w=64)
        again = atk(v, labels, ["again"]).contiguous()
        out["shard_bitwise_equal"] = bool(torch.equal(again, buf))

    # ---- 2. one backbone per GPU -------------------------------------------------------------------
    names = ["resnet", "vgg", "squeezenet", "alexnet"]
    v, _ = synth.clip(7, b=1, f=2, h=64, w=64)
    ens_d = image_attacks.ImageGuidedFML2_Adam_MultiModels(names, {"resnet": 2, "vgg": 3, "squeezenet": 2, "alexnet": 3},
                                                           steps=3, placement="ensemble")
    adv_d = ens_d(v, labels, ["c"]).contiguous()
    cost_d = [float(ens_d.loss_info["c"][i]["cost"]) for i in range(3)]
    aens_d = TPAMI_attack.AENS_I2V_MF(names, {n: [2, 3] for n in names}, 0.005, momentum=0.5, steps=3, placement="ensemble")
    aadv_d, _, acost_d = aens_d(v, labels, ["c"])
    aadv_d = aadv_d.contiguous()
    # every rank must hold the same replicated state
    ref = adv_d.clone()
    dist.broadcast(ref, src=0)
    same = torch.tensor([float(torch.equal(ref, adv_d))], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    out["ensemble_ranks_identical"] = bool(same.item() == 1.0)
    out["members_rank%d" % rank] = ens_d._plan.members
    if rank == 0:
        ens_1 = image_attacks.ImageGuidedFML2_Adam_MultiModels(names, {"resnet": 2, "vgg": 3, "squeezenet": 2, "alexnet": 3}, steps=3)
        adv_1 = ens_1(v, labels, ["c"]).contiguous()
        cost_1 = [float(ens_1.loss_info["c"][i]["cost"]) for i in range(3)]
        aens_1 = TPAMI_attack.AENS_I2V_MF(names, {n: [2, 3] for n in names}, 0.005, momentum=0.5, steps=3)
        aadv_1, _, acost_1 = aens_1(v, labels, ["c"])
        out["ens_cost_rel_err"] = float(np.abs(np.array(cost_d) / np.array(cost_1) - 1).max())
        out["aens_cost_rel_err"] = float(np.abs(acost_d / acost_1 - 1).max())
        out["aens_weights_rel_err"] = float(np.abs(np.stack(aens_d.weights) / np.stack(aens_1.weights) - 1).max())
        d = (adv_d - adv_1).abs()
        out["ens_adv_frac_equal"] = float((d == 0).float().mean())
        out["ens_adv_max_abs"] = float(d.max())
        x01 = v.to(dev) * torch.tensor(synth.STD, device=dev).view(1, 3, 1, 1, 1) + torch.tensor(synth.MEAN, device=dev).view(1, 3, 1, 1, 1)
        a01 = adv_d * torch.tensor(synth.STD, device=dev).view(1, 3, 1, 1, 1) + torch.tensor(synth.MEAN, device=dev).view(1, 3, 1, 1, 1)
        out["eps_ok"] = bool((a01 - x01).abs().max() <= 16 / 255 + 1e-6)
        ok = (out["shard_bitwise_equal"] and out["ensemble_ranks_identical"] and out["ens_cost_rel_err"] <= 1e-5
              and out["aens_cost_rel_err"] <= 1e-5 and out["aens_weights_rel_err"] <= 1e-5 and out["eps_ok"])
        out["ok"] = bool(ok)
        print(json.dumps(out), flush=True)
    D.barrier()
    dist.destroy_process_group()
    if rank == 0 and not out["ok"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
