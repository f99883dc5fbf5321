#!/usr/bin/env python
"""Drop-in for the reference's `image_main_ucf101.py` — `image_main.py` with the four differences the reference has
(`diff image_main.py image_main_ucf101.py`): `--step` defaults to 10 (26), the clips come from the UCF-101 loader
`dataset_ucf101.attack_genearte_dataeset(batch_size)` (54; here: that loader if it imports, else seeded synthetic clips with
labels mod 101), the ENS-I2V attack is constructed with `steps=args.step` (75), and `video_names = str(val_label)` (83 —
so `loss_info` is keyed by the CHARACTERS of the label tensor's repr; kept, it only affects the json side file).
Everything else — flags, sharding, `{label}-adv.npy`, `loss_info_{index}.json`, the async saver — is image_main.py's.
"""
import json
import os

import torch

import image_attacks
import image_main as im
from i2v_b200 import dist as D
from i2v_b200 import synth


def arg_parse(argv=None):
    args = im.arg_parse(argv)
    if not _flag_given(argv, "--step"):
        args.step = 10                                             # image_main_ucf101.py:26
        args.adv_path = os.path.join(args.opt_path, "{}-{}-{}-{}".format("Image", args.attack_method, args.step, args.file_prefix))
    return args


def _flag_given(argv, flag):
    import sys
    items = sys.argv[1:] if argv is None else argv
    return any(a == flag or a.startswith(flag + "=") for a in items)


def build_attack(args):
    if args.attack_method == "ImageGuidedFML2_Adam_MultiModels":   # image_main_ucf101.py:75 passes the step count
        model_name_lists = ["resnet", "vgg", "squeezenet", "alexnet"]
        depths = {"resnet": 2, "vgg": 3, "squeezenet": 2, "alexnet": 3}
        return image_attacks.ImageGuidedFML2_Adam_MultiModels(model_name_lists, depths=depths, steps=args.step, engine=args.engine)
    return im.build_attack(args)


def get_loader(args):
    if not args.synthetic:
        try:
            from dataset_ucf101 import attack_genearte_dataeset    # the reference's UCF-101 loader (decord + the videos)
        except ImportError as exc:
            raise SystemExit("image_main_ucf101.py: the reference UCF-101 pipeline is not importable (%s); run inside the "
                             "reference environment or pass --synthetic for seeded synthetic clips" % exc)
        items = list(attack_genearte_dataeset(args.batch_size))
        return len(items), lambda i: items[i]
    sl = im.SyntheticLoader(args.num_clips, args.batch_size, args.frames, args.side)

    def step(i):
        vids, labs, names = sl.step(i)
        return vids, labs % 101, names                             # UCF-101 has 101 classes
    return len(sl), step


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
        mine = D.clip_shard(n_steps, rank, world)
        index = rank + 1
    else:
        nums, index = args.batch_nums or 1, args.batch_index or 1
        nums_contained = int(n_steps / nums)
        mine = list(range((index - 1) * nums_contained, index * nums_contained))
    attack_method = build_attack(args)
    saver = im.AsyncSaver(args.adv_path)
    for step in mine:
        print("Running {}, {}/{}".format(args.attack_method, step + 1, n_steps))
        data = get_step(step)
        val_batch, val_label = data[0], data[1]
        video_names = str(val_label)                               # image_main_ucf101.py:83
        out = attack_method(val_batch, val_label, video_names)
        saver.submit(out[0] if isinstance(out, tuple) else out, val_label)
    saver.close()
    with open(os.path.join(args.adv_path, "loss_info_{}.json".format(index)), "w") as opt:
        json.dump(attack_method.loss_info, opt)
    with open(os.path.join(args.adv_path, "run_info_{}.json".format(index)), "w") as opt:
        json.dump({"args": dict(vars(args)), "weight_source": dict(backbones.WEIGHT_SOURCE),
                   "data_source": "synthetic (i2v_b200.synth.clip)" if args.synthetic else "reference loader",
                   "clips": len(mine)}, opt, indent=1)
    return attack_method


if __name__ == "__main__":
    main()
