#!/usr/bin/env python
"""Drop-in for the reference's `attack.py` (the white-box video-attack driver, attack.py:1-96): same flags, the same
class lookup — `getattr(base_attacks, name)(model, steps=...)` for `--attack_type image`, `getattr(video_attacks,
name)(model, params=..., steps=...)` for `--attack_type video` (76-84) — and the same artefacts: `{label}-adv.npy` and
`{label}-ori.npy`, float32 `[3,T,H,W]` in normalised space (92-96).

What differs, and why:
  * The white-box model and the clip source.  The reference builds a gluoncv video model from a yaml config and iterates
    its Kinetics-400 loader (65-72); neither the model zoo nor the data is reachable offline.  If `gluoncv` and the
    reference's `datasets` import, they are used exactly as in the reference; otherwise (or with `--synthetic`) the model
    is a seeded stand-in (`--model tiny` = i2v_b200.synth.TinyVideoNet, `--model r3d_18` = torchvision's, random init)
    and the clips come from `i2v_b200.synth.clip` (Kinetics-shaped, labels i mod num_classes).
  * Sharding: under `torchrun` every rank takes the loader steps `rank, rank + world, ...` (clips are independent units,
    no collective); the reference runs one process.
  * Saving overlaps the next clip (pinned buffers + writer thread, as in image_main.py).
The attack classes run their update blocks in this repo's sm_100a kernels (K3b/K3c/K7/K8); the model's forward/backward is
torch autograd on the given module with TF32 off (base_attacks.fp32_parity).
"""
import argparse
import os

import numpy as np
import torch

import base_attacks
import video_attacks
from i2v_b200 import dist as D
from i2v_b200 import synth
from image_main import AsyncSaver


def arg_parse(argv=None):
    parser = argparse.ArgumentParser(description="white-box video attacks (base_attacks / video_attacks) on B200")
    parser.add_argument("--gpu", type=str, default=None, help="gpu device (ignored under torchrun: LOCAL_RANK wins)")
    parser.add_argument("--batch_size", type=int, default=4, metavar="N")
    parser.add_argument("--model", type=str, default="i3d_resnet101",
                        help="i3d_resnet101 | i3d_slow_resnet101 | slowfast_resnet101 | tpn_resnet101 (gluoncv), or the stand-ins tiny | r3d_18")
    parser.add_argument("--attack_method", type=str, default="BIM",
                        help="FGSM | BIM | MIFGSM | DIFGSM | TIFGSM | SGM | SIM | TIFGSM3D (image) or TemporalTranslation (video)")
    parser.add_argument("--attack_type", type=str, default="image", help="image | video")
    parser.add_argument("--step", type=int, default=10, metavar="N")
    parser.add_argument("--kernlen", type=int, default=15, metavar="N")
    parser.add_argument("--file_prefix", type=str, default="")
    parser.add_argument("--kernel_mode", type=str, default="gaussian")
    parser.add_argument("--iterative_momentum", action="store_true", default=False)
    parser.add_argument("--augmentation_weight", type=float, default=1.0)
    parser.add_argument("--move_type", type=str, default="adj", help="adj | large | random")
    # extensions
    parser.add_argument("--opt_path", type=str, default=os.environ.get("I2V_OPT_PATH", "./i2v_out"),
                        help="output root (the reference's utils.OPT_PATH)")
    parser.add_argument("--synthetic", action="store_true", help="seeded stand-in model and synthetic clips")
    parser.add_argument("--num_clips", type=int, default=400)
    parser.add_argument("--frames", type=int, default=32)
    parser.add_argument("--side", type=int, default=224)
    parser.add_argument("--num_classes", type=int, default=400)
    args = parser.parse_args(argv)
    args.adv_path = os.path.join(args.opt_path, "{}-{}-{}-{}".format(args.model, args.attack_method, args.step, args.file_prefix))
    return args


def standin_model(name, num_classes):
    """Seeded random-init white-box stand-ins for the gluoncv zoo."""
    if name == "r3d_18":
        import torchvision
        state = torch.random.get_rng_state()
        torch.manual_seed(0)
        try:
            return torchvision.models.video.r3d_18(weights=None, num_classes=num_classes)
        finally:
            torch.random.set_rng_state(state)
    return synth.TinyVideoNet(num_classes=num_classes)


def get_model_and_loader(args):
    """(model on the current device, number of loader steps, step -> (val_batch, val_label))."""
    if not args.synthetic:
        try:
            from datasets import get_dataset                      # the reference's pipeline, if its environment exists
            from gluoncv.torch.model_zoo import get_model
            from utils import CONFIG_PATHS, get_cfg_custom
            cfg = get_cfg_custom(CONFIG_PATHS[args.model], args.batch_size)   # attack.py:65-66
            items = list(get_dataset(cfg))
            return get_model(cfg).cuda(), len(items), lambda i: items[i][:2]
        except Exception as exc:                                  # noqa: BLE001 — gluoncv / decord / data are absent offline
            print("reference model zoo / data pipeline unavailable (%s: %s) -> stand-in model, synthetic clips"
                  % (type(exc).__name__, exc))
    model = standin_model(args.model if args.model in ("tiny", "r3d_18") else "tiny", args.num_classes).cuda().eval()
    n_steps = (args.num_clips + args.batch_size - 1) // args.batch_size

    def step(i):
        b = min(args.batch_size, args.num_clips - i * args.batch_size)
        vids, labs = [], []
        for k in range(b):
            idx = i * args.batch_size + k
            vids.append(synth.clip(idx, b=1, f=args.frames, h=args.side, w=args.side)[0])
            labs.append(idx % args.num_classes)
        return torch.cat(vids, 0), torch.tensor(labs, dtype=torch.long)
    return model, n_steps, step


def build_attack(args, model):
    """attack.py:76-84"""
    if args.attack_type == "image":
        return getattr(base_attacks, args.attack_method)(model, steps=args.step)
    if args.attack_type == "video":
        if args.attack_method != "TemporalTranslation":
            raise ValueError("--attack_type video knows TemporalTranslation only (attack.py:79-84)")
        spe_params = {"kernlen": args.kernlen, "momentum": args.iterative_momentum, "weight": args.augmentation_weight,
                      "move_type": args.move_type, "kernel_mode": args.kernel_mode}
        print("Used Params")
        print(spe_params)
        return getattr(video_attacks, args.attack_method)(model, params=spe_params, steps=args.step)
    raise ValueError("--attack_type must be image or video, got %r" % (args.attack_type,))


class _OriSaver:
    """`{label}-ori.npy` next to `{label}-adv.npy` (attack.py:96); the originals are already host tensors."""

    def __init__(self, adv_path):
        self.adv_path = adv_path

    def save(self, val_batch, labels):
        host = val_batch.detach().cpu()
        for ind, label in enumerate(labels):
            np.save(os.path.join(self.adv_path, "{}-ori".format(int(label))), host[ind].numpy())


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
    saver, ori = AsyncSaver(args.adv_path), _OriSaver(args.adv_path)
    for step in mine:
        print("Running {}, {}/{}".format(args.attack_method, step + 1, n_steps))
        val_batch, val_label = get_step(step)
        adv_batches = attack_method(val_batch.cuda(), val_label.cuda())          # attack.py:88-90
        saver.submit(adv_batches, val_label)
        ori.save(val_batch, val_label)
    saver.close()


if __name__ == "__main__":
    main()
