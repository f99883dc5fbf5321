#!/usr/bin/env python
"""Drop-in for the reference's `attack_ucf101.py` — `attack.py` with the reference's differences (`diff attack.py
attack_ucf101.py`): output directory `UCF101_Video-...` / `UCF101_Image-...` (57-59), the UCF-101 loader and fine-tuned
checkpoints (74-79; here: used when gluoncv / the checkpoints import, else the seeded stand-in model and synthetic clips
with labels mod 101), and fixed TemporalTranslation parameters `kernlen=15, momentum=False, weight=1.0, move_type='adj',
kernel_mode='gaussian'` (87).  Flags, dispatch, artefacts and the async saver are attack.py's.
"""
import os

import torch

import attack as at
import base_attacks
import video_attacks
from i2v_b200 import dist as D
from image_main import AsyncSaver


def arg_parse(argv=None):
    args = at.arg_parse(argv)
    if not at_flag_given(argv, "--num_classes"):
        args.num_classes = 101
    kind = "UCF101_Video" if args.attack_type == "video" else "UCF101_Image"
    args.adv_path = os.path.join(args.opt_path, "{}-{}-{}-{}-{}".format(kind, args.model, args.attack_method, args.step,
                                                                         args.file_prefix))
    return args


def at_flag_given(argv, flag):
    import sys
    items = sys.argv[1:] if argv is None else argv
    return any(a == flag or a.startswith(flag + "=") for a in items)


def get_model_and_loader(args):
    if not args.synthetic:
        try:
            from dataset_ucf101 import attack_genearte_dataeset   # the reference's pipeline, if its environment exists
            from gluoncv.torch.model_zoo import get_model
            from reference_ucf101 import MODEL_TO_CKPTS
            from utils import CONFIG_PATHS, get_cfg_custom
            cfg = get_cfg_custom(CONFIG_PATHS[args.model], args.batch_size)
            items = list(attack_genearte_dataeset(args.batch_size))                       # 74
            model = get_model(cfg)
            model.load_state_dict(torch.load(MODEL_TO_CKPTS[args.model])["state_dict"])    # 75-77
            return model.cuda().eval(), len(items), lambda i: items[i][:2]
        except Exception as exc:                                  # noqa: BLE001 — gluoncv / checkpoints / data are absent offline
            print("reference UCF-101 pipeline unavailable (%s: %s) -> stand-in model, synthetic clips" % (type(exc).__name__, exc))
    args.synthetic = True
    return at.get_model_and_loader(args)


def build_attack(args, model):
    if args.attack_type == "video" and args.attack_method == "TemporalTranslation":
        spe_params = {"kernlen": 15, "momentum": False, "weight": 1.0, "move_type": "adj", "kernel_mode": "gaussian"}   # 87
        print("Used Params")
        print(spe_params)
        return video_attacks.TemporalTranslation(model, params=spe_params, steps=args.step)
    return at.build_attack(args, model)


def main(argv=None):
    args = arg_parse(argv)
    rank, local_rank, world = D.env_world()
    if world > 1:
        torch.cuda.set_device(local_rank)
    elif args.gpu is not None:
        torch.cuda.set_device(int(args.gpu.split(",")[0]))
    os.makedirs(args.adv_path, exist_ok=True)
    print(args)
    model, n_steps, get_step = get_model_and_loader(args)
    attack_method = build_attack(args, model)
    mine = D.clip_shard(n_steps, rank, world) if world > 1 else list(range(n_steps))
    saver, ori = AsyncSaver(args.adv_path), at._OriSaver(args.adv_path)
    for step in mine:
        print("Running {}, {}/{}".format(args.attack_method, step + 1, n_steps))
        val_batch, val_label = get_step(step)
        adv_batches = attack_method(val_batch.cuda(), val_label.cuda())
        saver.submit(adv_batches, val_label)
        ori.save(val_batch, val_label)
    saver.close()


if __name__ == "__main__":
    main()
