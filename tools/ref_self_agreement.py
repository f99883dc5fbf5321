"""Noise floor of the 60-step parity figure: the UNMODIFIED reference against ITSELF.

Runs reference image_attacks.ImageGuidedFMDirection_Adam (via oracle/load_reference.py, CPU) on the 60-step config-1
fixture's input with a different — equally correct — float32 convolution backend than the fixture was generated with and
scores the two runs exactly as tests/test_gpu_attacks.py::test_config1_60_steps_vs_reference_fixture scores the CUDA path
against the fixture.  The fixture run uses oneDNN; `nomkldnn` switches torch to its own im2col + GEMM convolutions
(`torch.backends.mkldnn.flags(enabled=False)`): same mathematics, different summation order — the same kind of
perturbation as cuDNN vs oneDNN or this repo's kernels vs either.  (A different thread count alone changes nothing:
oneDNN's partitioning is deterministic in the result, measured 8 vs 4 threads -> bit-identical.)  Build container only:

    python tools/ref_self_agreement.py 8 nomkldnn > profiles/r02_reference_self_agreement.json
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import load_reference as LR   # noqa: E402
from oracle import make_golden as MG      # noqa: E402


def main():
    threads = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    torch.set_num_threads(threads)
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "i2v_resnet50_d2_224_60step.npz")))
    nomkl = len(sys.argv) > 2 and sys.argv[2] == "nomkldnn"
    with torch.backends.mkldnn.flags(enabled=not nomkl):
        rec = MG.run_config1_60step(LR.load(), int(g["frames"]), int(g["side"]), int(g["steps"]), float(g["step_size"]))
    d = np.abs(rec["delta16"].astype(np.float32) - g["delta16"].astype(np.float32))
    big = np.unpackbits(g["g_first_big_bits"]).astype(bool)
    pos_a, neg_a = np.unpackbits(g["g_first_pos_bits"]).astype(bool), np.unpackbits(g["g_first_neg_bits"]).astype(bool)
    pos_b, neg_b = np.unpackbits(rec["g_first_pos_bits"]).astype(bool), np.unpackbits(rec["g_first_neg_bits"]).astype(bool)
    same = (pos_a == pos_b) & (neg_a == neg_b)
    out = {"what": "reference vs reference (%s convolutions, %d threads, against the oneDNN / 4-thread fixture), 32 frames x "
                   "3x224x224, 60 steps" % ("torch-native im2col+GEMM" if nomkl else "oneDNN", threads),
           "cost_rel_err_max": float(np.abs(rec["cost"] / g["cost"] - 1).max()),
           "final_frac_within_1_255": float((d <= (1 / 255) / 0.225).mean()),
           "final_frac_equal_f16": float((d == 0).mean()),
           "final_max_abs": float(d.max()),
           "step1_sign_agreement_big": float(same[big].mean()), "step1_sign_agreement_all": float(same.mean()),
           "step1_gmax": [float(g["g_first_max"]), float(rec["g_first_max"])]}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
