#!/usr/bin/env python
"""Drop-in for the reference's `image_main.py` (the driver `run_image_guided.py` launches, one process per GPU):
same flags (image_main.py:15-48), same attack construction (65-80), same artefacts — `{label}-adv.npy`
float32 [3,f,H,W] in normalised space (90-92) and `loss_info_{batch_index}.json` (94-95).

What differs, and why:
  * The clip source.  The reference iterates a gluoncv Kinetics-400 loader (`datasets.get_dataset`), which
    needs decord, the videos and a config that are not reachable offline.  If `datasets` imports, it is
    used exactly as in the reference; otherwise (or with `--synthetic`) clips come from `i2v_b200.synth.clip`
    (seeded, Kinetics-shaped: [b,3,32,224,224], labels i mod 400) — the 400-clip sweep of BASELINE.json
    configs[3].
  * Sharding.  `--batch_nums/--batch_index` keep their meaning (contiguous slices of the 400 loader steps,
    image_main.py:61-63).  Under `torchrun` (WORLD_SIZE > 1) they default to the rank layout, one process
    per GPU, round-robin over clips so that a sweep stays balanced — clips are independent units, no
    collective is needed.
  * `--attack_method AENS_I2V_MF` (TPAMI_attack.py:141-320) gets a branch; the reference never wired it (D6).
  * Saving overlaps the next clip: the adversarial clip is copied to pinned host memory on a side stream
    and a writer thread does the `np.save`.
"""
import argparse
import json
import os
import queue
import threading

import numpy as np
import torch

import image_attacks
import TPAMI_attack
from i2v_b200 import dist as D
from i2v_b200 import synth


def arg_parse(argv=None):
    parser = argparse.ArgumentParser(description="image-guided (I2V) attacks on B200")
    # parallel run (image_main.py:18-19)
    parser.add_argument("--batch_nums", type=int, default=None)
    parser.add_argument("--batch_index", type=int, default=None)
    parser.add_argument("--gpu", type=str, default=None, help="gpu device (ignored under torchrun: LOCAL_RANK wins)")
    parser.add_argument("--batch_size", type=int, default=1, metavar="N", help="clips per attack call")
    parser.add_argument("--attack_method", type=str, default="ImageGuidedFMDirection_Adam",
                        help="ImageGuidedStd_Adam | ImageGuidedFMDirection_Adam | ImageGuidedFML2_Adam_MultiModels | AENS_I2V_MF")
    parser.add_argument("--step", type=int, default=60, metavar="N")
    parser.add_argument("--file_prefix", type=str, default="")
    parser.add_argument("--depth", type=int, default=1, help="1,2,3,4")
    parser.add_argument("--lamb", type=float, default=0.1)
    parser.add_argument("--mode", type=str, default="direction")
    parser.add_argument("--step_size", type=float, default=0.004)
    parser.add_argument("--dropout", type=float, default=0.1)
    parser.add_argument("--direction_image_model", type=str, default="resnet", help="resnet, squeezenet, vgg, alexnet")
    # extensions
    parser.add_argument("--opt_path", type=str, default=os.environ.get("I2V_OPT_PATH", "./i2v_out"),
                        help="output root (the reference's utils.OPT_PATH)")
    parser.add_argument("--synthetic", action="store_true", help="seeded synthetic Kinetics-shaped clips")
    parser.add_argument("--num_clips", type=int, default=400, help="loader length with --synthetic")
    parser.add_argument("--frames", type=int, default=32)
    parser.add_argument("--side", type=int, default=224)
    parser.add_argument("--momentum", type=float, default=0.0, help="AENS_I2V_MF")
    parser.add_argument("--engine", type=str, default=None, help="native | native_tf32 | cudnn | cudnn_tf32")
    parser.add_argument("--weights", type=str, default=os.environ.get("I2V_WEIGHTS", "pretrained"),
                        help="pretrained (default, as the reference: fails when the ImageNet checkpoints are not cached) | "
                             "random (seeded random init: tests / benchmarks only) | auto (backbones.set_weight_policy)")
    args = parser.parse_args(argv)
    args.adv_path = os.path.join(args.opt_path, "{}-{}-{}-{}".format("Image", args.attack_method, args.step, args.file_prefix))
    return args


def build_attack(args):
    """image_main.py:65-80 (+ the AENS branch)."""
    name = args.attack_method
    if name in ("ImageGuidedStd_Adam", "ImageGuidedFMDirection_Adam"):
        return getattr(image_attacks, name)([args.direction_image_model], depth=args.depth, step_size=args.step_size,
                                            steps=args.step, engine=args.engine)
    model_name_lists = ["resnet", "vgg", "squeezenet", "alexnet"]
    if name == "ImageGuidedFML2_Adam_MultiModels":
        depths = {"resnet": 2, "vgg": 3, "squeezenet": 2, "alexnet": 3}
        # the reference passes neither --step nor --step_size here (image_main.py:80): ctor defaults 60 / 0.005
        return image_attacks.ImageGuidedFML2_Adam_MultiModels(model_name_lists, depths=depths, engine=args.engine)
    if name == "AENS_I2V_MF":
        depths = {m: [2, 3] for m in model_name_lists}
        return TPAMI_attack.AENS_I2V_MF(model_name_lists, depths, args.step_size, momentum=args.momentum, steps=args.step,
                                        engine=args.engine)
    raise ValueError("unknown --attack_method %r" % name)


class SyntheticLoader:
    """Yields (val_batch [b,3,f,H,W] normalised f32, val_label [b] int64, video_names) like the gluoncv loader."""

    def __init__(self, num_clips, batch_size, frames, side):
        self.n = (num_clips + batch_size - 1) // batch_size
        self.b, self.f, self.side, self.num_clips = batch_size, frames, side, num_clips

    def __len__(self):
        return self.n

    def step(self, i):
        b = min(self.b, self.num_clips - i * self.b)
        vids, labs, names = [], [], []
        for k in range(b):
            idx = i * self.b + k
            v, _ = synth.clip(idx, b=1, f=self.f, h=self.side, w=self.side)
            vids.append(v)
            labs.append(idx % 400)
            names.append("synthetic_%05d" % idx)
        return torch.cat(vids, 0), torch.tensor(labs, dtype=torch.long), names


def get_loader(args):
    """image_main.py:52-58.  Without --synthetic the reference's own data pipeline is REQUIRED: its `datasets` module
    and the gluoncv config helpers of its `utils.py` (CONFIG_PATHS, get_cfg_custom) must be importable from the
    reference environment ($I2V_REFERENCE_ROOT on sys.path).  Nothing is substituted silently: a missing pipeline is a
    SystemExit, a broken dataset raises whatever it raises."""
    if not args.synthetic:
        try:
            from datasets import get_dataset                      # the reference's loader (gluoncv + decord)
            import utils as ref_utils
            CONFIG_PATHS, get_cfg_custom = ref_utils.CONFIG_PATHS, ref_utils.get_cfg_custom
        except (ImportError, AttributeError) as exc:
            raise SystemExit("image_main.py: the reference data pipeline is not importable (%s: %s); run inside the "
                             "reference environment or pass --synthetic for seeded synthetic clips"
                             % (type(exc).__name__, exc))
        cfg = get_cfg_custom(CONFIG_PATHS["i3d_resnet101"], args.batch_size)
        loader = get_dataset(cfg)
        items = list(loader)
        return len(items), lambda i: items[i]
    sl = SyntheticLoader(args.num_clips, args.batch_size, args.frames, args.side)
    return len(sl), sl.step


class AsyncSaver:
    """D2H on a side stream into pinned buffers + a writer thread, so `np.save` overlaps the next clip's attack."""

    def __init__(self, adv_path):
        self.adv_path = adv_path
        self.q = queue.Queue(maxsize=4)
        self.stream = torch.cuda.Stream()
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def submit(self, adv_batches, labels):
        adv = adv_batches.detach()
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            host = torch.empty(adv.shape, dtype=adv.dtype, pin_memory=True)
            host.copy_(adv, non_blocking=True)
            adv.record_stream(self.stream)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.q.put((host, [int(l) for l in labels], ev))

    def _run(self):
        while True:
            item = self.q.get()
            if item is None:
                return
            host, labels, ev = item
            ev.synchronize()
            for ind, label in enumerate(labels):
                np.save(os.path.join(self.adv_path, "{}-adv".format(label)), host[ind].numpy())   # image_main.py:90-92

    def close(self):
        self.q.put(None)
        self.thread.join()


def main(argv=None):
    args = arg_parse(argv)
    rank, local_rank, world = D.env_world()
    if world > 1:
        torch.cuda.set_device(local_rank)
    elif args.gpu is not None:
        torch.cuda.set_device(int(args.gpu.split(",")[0]))
    from i2v_b200 import backbones
    backbones.set_weight_policy(args.weights, 0)
    os.makedirs(args.adv_path, exist_ok=True)
    print(args)

    n_steps, get_step = get_loader(args)
    if args.batch_nums is None and world > 1:
        mine = D.clip_shard(n_steps, rank, world)                 # one process per GPU, round-robin
        index = rank + 1
    else:
        nums, index = args.batch_nums or 1, args.batch_index or 1
        nums_contained = int(n_steps / nums)                      # image_main.py:61-63
        mine = list(range((index - 1) * nums_contained, index * nums_contained))

    attack_method = build_attack(args)
    saver = AsyncSaver(args.adv_path)
    for step in mine:
        print("Running {}, {}/{}".format(args.attack_method, step + 1, n_steps))
        val_batch, val_label, video_names = get_step(step)[:3]
        out = attack_method(val_batch, val_label, video_names)
        adv_batches = out[0] if isinstance(out, tuple) else out   # AENS returns (adv, used_time, cost_saved)
        saver.submit(adv_batches, val_label)
    saver.close()
    with open(os.path.join(args.adv_path, "loss_info_{}.json".format(index)), "w") as opt:   # image_main.py:94-95
        json.dump(attack_method.loss_info, opt)
    # provenance next to the artefacts: which weights the surrogate backbones really had and where the clips came from
    with open(os.path.join(args.adv_path, "run_info_{}.json".format(index)), "w") as opt:
        json.dump({"args": {k: v for k, v in vars(args).items()}, "weight_source": dict(backbones.WEIGHT_SOURCE),
                   "data_source": "synthetic (i2v_b200.synth.clip)" if args.synthetic else "reference loader",
                   "clips": len(mine)}, opt, indent=1)


if __name__ == "__main__":
    main()
