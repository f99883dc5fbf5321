#!/usr/bin/env python
"""Drop-in for the reference's `image_fine_tune_attack.py` (1-81): ILAF fine-tuning of adversarial clips that an
image-guided attack saved earlier.  Same flags, the same `AdvDataset` over `{label}-adv.npy` / `{label}-ori.npy` pairs
(16-37), the same dispatch `getattr(image_attacks, args.attack_method)(model, args.white_model)` (67) and the same output
`{label}-adv.npy` under `--opt_path` (77-81) — i.e. the tensor ILAF returns, with the reference's frame / channel
reinterpretation (image_attacks.py:627-629) kept.

The white-box video model comes from gluoncv when it imports (55-62); otherwise, or with `--synthetic`, a seeded
stand-in whose hooked attribute matches the model-type string: `--white_model tpn_tiny` -> i2v_b200.synth.TinyTPNLike
(`layer2` is hooked, image_attacks.py:518-519).  `--steps` / `--step_size` expose ILAF's constructor defaults (60, 0.005).
"""
import argparse
import os

import numpy as np
import torch
from torch.utils.data import Dataset

import image_attacks
from i2v_b200 import synth


class AdvDataset(Dataset):
    """image_fine_tune_attack.py:16-37: item = (adv [1,3,T,H,W], ori [1,3,T,H,W], label [1] int64)."""

    def __init__(self, used_adv_path, used_ori_path):
        self.used_adv_path = used_adv_path
        self.files = sorted(i for i in os.listdir(self.used_adv_path) if "adv" in i)
        self.used_ori_path = used_ori_path

    def __len__(self):
        return len(self.files)

    def __getitem__(self, idx):
        file = self.files[idx]
        vid_id = file.split("-")[0]
        ori_file = os.path.join(self.used_ori_path, "{}-ori.npy".format(vid_id))
        vid = torch.from_numpy(np.load(os.path.join(self.used_adv_path, file)))[None]
        ori_vid = torch.from_numpy(np.load(ori_file))[None]
        label = torch.from_numpy(np.array([int(vid_id)]).astype(np.int32)).long()
        return vid, ori_vid, label


def arg_parse(argv=None):
    parser = argparse.ArgumentParser(description="ILAF fine-tuning of saved adversarial clips on B200")
    parser.add_argument("--gpu", type=str, default=None, help="gpu device")
    parser.add_argument("--batch_size", type=int, default=4, metavar="N")
    parser.add_argument("--attack_method", type=str, default="ILAF")
    parser.add_argument("--opt_path", type=str, default="")
    parser.add_argument("--used_adv", type=str, default="")
    parser.add_argument("--used_ori", type=str, default="")
    parser.add_argument("--white_model", type=str, default="i3d_resnet101",
                        help="i3d_resnet101 | slowfast_resnet101 | tpn_resnet101 (gluoncv), or the stand-in tpn_tiny")
    parser.add_argument("--dataset", type=str, default="Kinetics-400", help="Kinetics-400 | UCF-101")
    # extensions
    parser.add_argument("--synthetic", action="store_true", help="seeded stand-in model instead of the gluoncv zoo")
    parser.add_argument("--num_classes", type=int, default=400)
    parser.add_argument("--steps", type=int, default=60)
    parser.add_argument("--step_size", type=float, default=0.005)
    return parser.parse_args(argv)


def get_white_model(args):
    if not args.synthetic:
        try:
            from gluoncv.torch.model_zoo import get_model         # the reference's zoo, if its environment exists
            from utils import CONFIG_PATHS, get_cfg_custom
            cfg = get_cfg_custom(CONFIG_PATHS[args.white_model], args.batch_size)   # 58-59
            model = get_model(cfg)
            if args.dataset == "UCF-101":
                from reference_ucf101 import MODEL_TO_CKPTS
                model.load_state_dict(torch.load(MODEL_TO_CKPTS[args.white_model])["state_dict"])   # 61-63
            return model.cuda()
        except Exception as exc:                                  # noqa: BLE001 — gluoncv / checkpoints are absent offline
            print("reference model zoo unavailable (%s: %s) -> stand-in model" % (type(exc).__name__, exc))
    if "tpn" not in args.white_model:
        raise ValueError("the offline stand-in exists for a 'tpn' model type only (--white_model tpn_tiny)")
    return synth.TinyTPNLike(num_classes=args.num_classes).cuda().eval()


def main(argv=None):
    args = arg_parse(argv)
    if args.gpu is not None:
        torch.cuda.set_device(int(args.gpu.split(",")[0]))
    print(args)
    model = get_white_model(args)
    dataset = AdvDataset(used_adv_path=args.used_adv, used_ori_path=args.used_ori)
    os.makedirs(args.opt_path, exist_ok=True)
    attack_method = getattr(image_attacks, args.attack_method)(model, args.white_model, step_size=args.step_size,
                                                               steps=args.steps)          # 67
    for step in range(len(dataset)):
        print("Running {}, {}/{}".format(args.attack_method, step + 1, len(dataset)))
        val_batch, ori_batch, val_label = dataset[step]
        video_names = ["..."]                                                                # 74
        adv_batches = attack_method(val_batch, ori_batch, val_label, video_names)
        for ind, label in enumerate(val_label):
            np.save(os.path.join(args.opt_path, "{}-adv".format(label.item())), adv_batches[ind].detach().cpu().numpy())
    return attack_method


if __name__ == "__main__":
    main()
