"""TEST INFRASTRUCTURE ONLY — generate tests/golden/*.npz by running the UNMODIFIED reference classes
(/root/reference, loaded through oracle/load_reference.py) on seeded synthetic inputs, CPU, float32.

Run in the build container (the reference is not present on the GPU box):
    python -m oracle.make_golden
The fixtures pin the oracle (tests/test_oracle_golden.py) and, on the GPU box, the CUDA path
(tests/test_gpu_attacks.py).  Weights are torchvision random init under torch.manual_seed(0); a
checksum of the first conv weight is stored so a torch version that initialises differently is detected.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import load_reference as LR   # noqa: E402
import i2v_b200.synth as synth            # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
THREADS = 4


def weight_checksum(model):
    p = next(model.parameters()).detach().double()
    return np.array([p.sum().item(), p.abs().sum().item(), float(p.flatten()[0])])


def run_image_guided(ref, kind, names, depths, shape, steps, step_size, **kw):
    b, f, h, w = shape
    videos, labels = synth.clip(0, b=b, f=f, h=h, w=w)
    with LR.quiet():
        if kind == "i2v":
            atk = ref.image_attacks.ImageGuidedFMDirection_Adam(names, depth=depths, step_size=step_size, steps=steps)
            models = [atk.model]
        elif kind == "dr":
            atk = ref.image_attacks.ImageGuidedStd_Adam(names, depth=depths, step_size=step_size, steps=steps)
            models = [atk.model]
        elif kind == "ens":
            atk = ref.image_attacks.ImageGuidedFML2_Adam_MultiModels(names, depths, steps=steps)
            models = atk.models
        else:
            atk = ref.TPAMI_attack.AENS_I2V_MF(names, depths, step_size, steps=steps, **kw)
            models = atk.models
    vids = ["clip0"] * b
    with LR.AdamSpy(keep=("grad", "param", "param_before", "exp_avg", "exp_avg_sq")) as spy, LR.quiet():
        out = atk(videos.clone(), labels, vids)
    rec = {"videos": videos.numpy(), "steps": steps, "step_size": step_size, "epsilon": 16 / 255,
           "weight_checksums": np.stack([weight_checksum(m) for m in models])}
    if kind == "aens":
        adv, _, cost_saved = out
        rec["cost_saved"] = cost_saved
        rec["weights"] = np.stack(atk.weights)
        rec["coeffs_after"] = atk.coeffs.numpy()
    else:
        adv = out
    rec["adv"] = adv.detach().contiguous().numpy()
    rec["cost"] = np.array([float(atk.loss_info["clip0"][i]["cost"]) for i in range(steps)], dtype=np.float32)
    # teacher-forcing taps: gradient w.r.t. the modifier and the Adam state of the first and last step
    # (ensemble fixtures keep only the first gradient: the Adam arithmetic is pinned by the i2v ones)
    for tag, r in (("first", spy.records[0]), ("last", spy.records[-1])):
        rec["g_mod_" + tag] = r["grad"].numpy()
        if kind not in ("i2v", "dr"):
            break
        rec["mod_" + tag] = r["param"].numpy()
        rec["m_" + tag] = r["exp_avg"].numpy()
        rec["v_" + tag] = r["exp_avg_sq"].numpy()
        if tag == "last":   # state before the last step: the last step can be teacher-forced on its own
            rec["mod_before_last"] = r["param_before"].numpy()
            rec["m_before_last"] = r["exp_avg_before"].numpy()
            rec["v_before_last"] = r["exp_avg_sq_before"].numpy()
    return rec


def run_base(ref):
    torch.manual_seed(0)
    model = synth.TinyVideoNet()
    videos, _ = synth.clip(3, b=1, f=32, h=12, w=12)
    labels = torch.tensor([3])
    rec = {"videos": videos.numpy(), "labels": labels.numpy(), "weight_checksums": weight_checksum(model)[None]}
    rec["fgsm"] = ref.base_attacks.FGSM(model)(videos.clone(), labels).numpy()
    rec["bim3"] = ref.base_attacks.BIM(model, steps=3)(videos.clone(), labels).numpy()
    rec["mifgsm3"] = ref.base_attacks.MIFGSM(model, steps=3)(videos.clone(), labels).numpy()
    tgt = ref.base_attacks.BIM(model, steps=2)
    tgt.set_attack_mode("targeted", lambda images, labels: (labels + 1) % 10)
    rec["bim2_targeted"] = tgt(videos.clone(), labels).numpy()
    g = torch.randn(2, 3, 32, 6, 6, generator=torch.Generator().manual_seed(5))
    rec["norm_grads_in"] = g.numpy()
    rec["norm_grads_frame"] = ref.utils.norm_grads(g.clone(), True).numpy()
    rec["norm_grads_clip"] = ref.utils.norm_grads(g.clone(), False).numpy()
    return rec


def run_base_variants(ref):
    """DIFGSM / TIFGSM / SGM / SIM / TIFGSM3D of the unmodified reference (base_attacks.py:342-683) on seeded inputs."""
    import random
    rec = {}
    torch.manual_seed(0)
    model = synth.TinyVideoNet()
    videos, _ = synth.clip(3, b=1, f=32, h=12, w=12)
    labels = torch.tensor([3])
    rec["videos"], rec["labels"] = videos.numpy(), labels.numpy()
    rec["weight_checksums"] = weight_checksum(model)[None]
    for mom in (False, True):
        tag = "_mom" if mom else ""
        rec["tifgsm3" + tag] = ref.base_attacks.TIFGSM(model, steps=3, momentum=mom)(videos.clone(), labels).numpy()
        rec["sim2" + tag] = ref.base_attacks.SIM(model, steps=2, momentum=mom)(videos.clone(), labels).numpy()
        rec["tifgsm3d2" + tag] = ref.base_attacks.TIFGSM3D(model, steps=2, momentum=mom)(videos.clone(), labels).numpy()
    relu_model = synth.TinyReluVideoNet()
    rec["relu_weight_checksums"] = weight_checksum(relu_model)[None]
    for mom in (False, True):
        tag = "_mom" if mom else ""
        with LR.quiet():
            atk = ref.base_attacks.SGM(synth.TinyReluVideoNet(), steps=3, momentum=mom)
            rec["sgm3" + tag] = atk(videos.clone(), labels).numpy()
    rec["bim3_relu"] = ref.base_attacks.BIM(synth.TinyReluVideoNet(), steps=3)(videos.clone(), labels).numpy()
    # DI needs 224 x 224 frames (constants of base_attacks.py:356-376).  To keep the fixture small the input is
    # synth.clip(5, b=1, f=2, h=224, w=224) (regenerated by the tests) and only the perturbation adv - videos is
    # stored, as float16 (|delta| <= eps/std = 0.28: 1e-4 resolution, enough to compare sign patterns)
    di_videos, _ = synth.clip(5, b=1, f=2, h=224, w=224)
    for mom in (False, True):
        random.seed(11)
        torch.manual_seed(11)
        adv = ref.base_attacks.DIFGSM(model, steps=4, momentum=mom)(di_videos.clone(), labels)
        rec["difgsm4_delta16" + ("_mom" if mom else "")] = (adv - di_videos).numpy().astype(np.float16)
    return rec


def run_video_variants(ref):
    """TemporalTranslation (video_attacks.py), TAP (base_attacks.py:685-814) and ILAF (image_attacks.py:498-629) on the
    seeded `synth.TinyTPNLike` stand-in (model_type 'tpn' -> layer1 / layer2 hooks), one 32-frame 12x12 clip."""
    rec = {}
    videos, _ = synth.clip(3, b=1, f=32, h=12, w=12)
    labels = torch.tensor([3])
    rec["videos"], rec["labels"] = videos.numpy(), labels.numpy()
    rec["weight_checksums"] = weight_checksum(synth.TinyTPNLike())[None]

    def tt(kernlen, weight, mode, steps, mom):
        with LR.quiet():
            atk = ref.video_attacks.TemporalTranslation(
                synth.TinyTPNLike(), {"kernlen": kernlen, "momentum": mom, "weight": weight, "move_type": "adj",
                                      "kernel_mode": mode}, steps=steps)
            return atk(videos.clone(), labels).detach().numpy()
    rec["tt3_k5"] = tt(5, 0.5, "gaussian", 3, False)
    rec["tt3_k5_mom"] = tt(5, 0.5, "gaussian", 3, True)
    rec["tt2_k9_linear"] = tt(9, 0.3, "linear", 2, False)

    def tt_move(move_type):
        import random
        random.seed(21)
        with LR.quiet():
            atk = ref.video_attacks.TemporalTranslation(
                synth.TinyTPNLike(), {"kernlen": 5, "momentum": True, "weight": 0.7, "move_type": move_type,
                                      "kernel_mode": "random"}, steps=2)
            return atk(videos.clone(), labels).detach().numpy()
    rec["tt2_k5_large"] = tt_move("large")
    rec["tt2_k5_randommove"] = tt_move("random")
    for conv3d in (True, False):
        with LR.quiet():
            atk = ref.base_attacks.TAP(synth.TinyTPNLike(), {"kernlen": 3, "temporal_kernlen": 3, "eta": 1e3, "conv3d": conv3d,
                                                            "model_type": "tpn"}, steps=3)
            adv = atk(videos.clone(), labels).detach().numpy()
        tag = "3d" if conv3d else "2d"
        rec["tap3_" + tag] = adv
        last = list(atk.loss_info.values())[-1]      # the reference keys loss_info by a tensor (790 shadows `i`): one entry survives
        rec["tap3_%s_last_losses" % tag] = np.array([float(last["ce loss"]), float(last["reg_cost"]),
                                                     float(np.asarray(last["distance"]).reshape(-1)[0])], dtype=np.float64)
    with LR.quiet():
        model = synth.TinyTPNLike()
        atk = ref.image_attacks.ILAF(model, "tpn", step_size=0.005, steps=4)
        out = atk(torch.from_numpy(rec["tt3_k5"]).clone(), videos.clone(), labels, ["v0"])
    rec["ilaf4"] = out.detach().contiguous().numpy()
    rec["ilaf4_costs"] = np.array([float(atk.loss_info["v0"][i]["cost"]) for i in range(4)], dtype=np.float64)
    return rec


def run_config1_60step(ref, frames=32, side=224, steps=60, step_size=0.005):
    """BASELINE.json configs[0] at its own size and step budget (/root/reference/run_image_guided.py:63-70: 60 steps of
    0.005 on one 32-frame 224x224 clip; image_attacks.py:294-364).  The input is synth.clip(0, b=1, f=32, h=224, w=224),
    regenerated by the tests; stored are the 60 costs, the final perturbation adv - videos as float16 (|delta| <=
    eps/std = 0.28: 2.4e-4 resolution against the 1/255/std = 0.017 agreement threshold), and of the FIRST step's
    dcost/dmodifier only what the sign-agreement score needs: the sign bits, the |g| > 1e-3 max mask, max |g|."""
    videos, labels = synth.clip(0, b=1, f=frames, h=side, w=side)
    with LR.quiet():
        atk = ref.image_attacks.ImageGuidedFMDirection_Adam(["resnet"], depth=2, step_size=step_size, steps=steps)
    first = {}
    orig = torch.optim.Adam

    class FirstGrad(orig):
        def step(self, closure=None):
            if not first:
                first["g"] = self.param_groups[0]["params"][0].grad.detach().clone()
            return super().step(closure)
    torch.optim.Adam = FirstGrad
    try:
        with LR.quiet():
            adv = atk(videos.clone(), labels, ["clip0"])
    finally:
        torch.optim.Adam = orig
    g = first["g"].numpy()
    gmax = np.abs(g).max()
    return {"frames": frames, "side": side, "steps": steps, "step_size": step_size, "epsilon": 16 / 255,
            "weight_checksums": weight_checksum(atk.model)[None],
            "cost": np.array([float(atk.loss_info["clip0"][i]["cost"]) for i in range(steps)], dtype=np.float32),
            "delta16": (adv.detach().contiguous() - videos).numpy().astype(np.float16),
            "g_first_shape": np.array(g.shape), "g_first_max": np.float64(gmax),
            "g_first_pos_bits": np.packbits(g.reshape(-1) > 0), "g_first_neg_bits": np.packbits(g.reshape(-1) < 0),
            "g_first_big_bits": np.packbits(np.abs(g.reshape(-1)) > 1e-3 * gmax)}


def add_float64_arbiter(path, frames=32, side=224):
    """Adds the float64 arbiter of the FIRST step to the 60-step fixture: dcost/dtrue_image of image_attacks.py:334-352 for
    the step-1 true_image (x + 0.01/255, clamped, normalised) evaluated by torch autograd in float64 on the CPU
    (oracle.loops.teacher_forced_grad) — sign bits, the |g| > 1e-3 max mask and max |g|, of dcost/dmodifier = g / std.
    At this size the reference's own float32 gradient is ~10 % (relative L2) away from it (the cosine's f32 sums), so
    the arbiter, not the reference, is what a more accurate implementation can be held to."""
    from oracle import loops as OL
    from oracle import oracle as O
    from i2v_b200 import backbones
    rec = dict(np.load(path))
    videos, _ = synth.clip(0, b=1, f=frames, h=side, w=side)
    fr = OL._frames(videos)
    x = O.denorm(fr.numpy(), side * side)
    ti = O.compose_norm(x, np.full_like(x, np.float32(OL.INIT_MODIFIER)), 16 / 255, side * side)
    hooked = [OL.HookedModel(backbones.seeded_random_init("resnet50", 0), "resnet", 2)]
    _, g64, _ = OL.teacher_forced_grad(hooked, fr, ti, torch.float64)
    gm = (g64 / O.STD[None, :, None, None].astype(np.float64)).reshape(-1)
    gmax = np.abs(gm).max()
    rec["g64_first_max"] = np.float64(gmax)
    rec["g64_first_pos_bits"] = np.packbits(gm > 0)
    rec["g64_first_big_bits"] = np.packbits(np.abs(gm) > 1e-3 * gmax)
    n = gm.size
    pos = np.unpackbits(rec["g_first_pos_bits"])[:n].astype(bool)
    big = np.abs(gm) > 1e-3 * gmax
    rec["ref_vs_f64_sign_agreement_big"] = np.float64(((gm > 0) == pos)[big].mean())
    np.savez_compressed(path, **rec)
    print("float64 arbiter added: reference-vs-float64 step-1 sign agreement %.5f" % rec["ref_vs_f64_sign_agreement_big"])


def main():
    torch.set_num_threads(THREADS)
    os.makedirs(OUT, exist_ok=True)
    ref = LR.load()
    ens_names = ["resnet", "vgg", "squeezenet", "alexnet"]
    jobs = {
        "i2v_resnet50_d2_32": lambda: run_image_guided(ref, "i2v", ["resnet"], 2, (1, 2, 32, 32), 3, 0.005),
        "i2v_vgg_d3_32": lambda: run_image_guided(ref, "i2v", ["vgg"], 3, (1, 2, 32, 32), 2, 0.005),
        "dr_resnet50_d2_32": lambda: run_image_guided(ref, "dr", ["resnet"], 2, (1, 2, 32, 32), 3, 0.005),
        "dr_vgg_d2_32": lambda: run_image_guided(ref, "dr", ["vgg"], 2, (1, 3, 32, 32), 2, 0.005),
        "ens_4models_64": lambda: run_image_guided(
            ref, "ens", ens_names, {"resnet": 2, "vgg": 3, "squeezenet": 2, "alexnet": 3}, (1, 2, 64, 64), 3, 0.005),
        "aens_4models_64": lambda: run_image_guided(
            ref, "aens", ens_names, {n: [2, 3] for n in ens_names}, (1, 2, 64, 64), 3, 0.005, momentum=0.5),
        "aens_ce_2models_64": lambda: run_image_guided(
            ref, "aens", ["resnet", "squeezenet"], {"resnet": [1, 2], "squeezenet": [2, 3]}, (1, 2, 64, 64), 2, 0.005,
            coef_CE=True),
        "base_tiny3d": lambda: run_base(ref),
        "base_variants": lambda: run_base_variants(ref),
        "video_variants": lambda: run_video_variants(ref),
        "i2v_resnet50_d2_224_60step": lambda: run_config1_60step(ref),
    }
    only = sys.argv[1:]
    for name, job in jobs.items():
        if only and name not in only:
            continue
        rec = job()
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **rec)
        print("%-24s %8.1f KB" % (name, os.path.getsize(path) / 1024))
        if name == "i2v_resnet50_d2_224_60step":
            add_float64_arbiter(path)


if __name__ == "__main__":
    main()
