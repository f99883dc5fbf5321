"""Host-side logic that needs no GPU: depth -> layer mapping, name handling, sharding, step scalars."""
import os
import numpy as np
import pytest
import torch

from i2v_b200 import backbones, capi, dist as D
from oracle import oracle as O


def test_depth_to_layer_mapping_matches_reference_tables():
    m = backbones.seeded_random_init("resnet50")
    assert backbones.find_target_layers(m, "resnet", 2) == [m.layer2[-1]]
    assert backbones.find_target_layers(m, "resnet", [2, 3]) == [m.layer2[-1], m.layer3[-1]]
    v = backbones.seeded_random_init("vgg16")
    assert backbones.find_target_layers(v, "vgg", 3) == [v.features[20]]
    a = backbones.seeded_random_init("alexnet")
    assert backbones.find_target_layers(a, "alexnet", [2, 3]) == [a.features[4], a.features[7]]
    s = backbones.seeded_random_init("squeezenet1_1")
    assert backbones.find_target_layers(s, "squeezenet", 2) == [s.features[6].expand3x3_activation]
    # list of depths hooks the whole Fire module (TPAMI_attack.py:195-198)
    assert backbones.find_target_layers(s, "squeezenet", [2, 3]) == [s.features[6], s.features[9]]


def test_unknown_model_and_depth_raise_value_error():
    with pytest.raises(ValueError):
        backbones.arch_of("inception")
    m = backbones.seeded_random_init("alexnet")
    with pytest.raises(ValueError):
        backbones.find_target_layers(m, "alexnet", 5)


def test_reference_names_keep_reference_archs():
    saved = dict(backbones.ARCH_OVERRIDE)
    backbones.ARCH_OVERRIDE.clear()
    try:
        assert backbones.arch_of("resnet") == "resnet101"      # image_attacks.py:95
        assert backbones.arch_of("densenet") == "densenet161"  # image_attacks.py:97
        assert backbones.arch_of("vgg") == "vgg16"
        assert backbones.arch_of("resnet50") == "resnet50"
    finally:
        backbones.ARCH_OVERRIDE.update(saved)


def test_seeded_init_is_reproducible_and_leaves_rng_alone():
    torch.manual_seed(123)
    before = torch.rand(1)
    torch.manual_seed(123)
    a = backbones.seeded_random_init("squeezenet1_1", 0)
    after = torch.rand(1)
    b = backbones.seeded_random_init("squeezenet1_1", 0)
    assert torch.equal(before, after)
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert torch.equal(pa, pb)


def test_adam_step_table_matches_oracle_scalars():
    table = capi.adam_step_table(60, 0.005).numpy()
    for k in (0, 1, 9, 59):
        bc2, nss = O.adam_step_scalars(0.005, 0.9, 0.999, k + 1)
        assert table[k, 0] == np.float32(bc2) and table[k, 1] == np.float32(nss)
    # and the python-double recipe of torch/optim/adam.py
    t = 7
    assert table[t - 1, 0] == np.float32((1 - 0.999 ** t) ** 0.5)
    assert table[t - 1, 1] == np.float32(-(0.005 / (1 - 0.9 ** t)))


def test_contiguous_shard_partitions_exactly():
    for n in (0, 1, 7, 400, 401):
        for world in (1, 2, 3, 8):
            spans = [D.contiguous_shard(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    # the reference's own split: 400 loader steps, --batch_nums 8 (image_main.py:61-63)
    assert D.contiguous_shard(400, 3, 8) == (150, 200)


def test_round_robin_shard_and_ensemble_placement():
    got = sorted(i for r in range(4) for i in D.clip_shard(10, r, 4))
    assert got == list(range(10))
    names = ["resnet", "vgg", "densenet", "squeezenet"]
    assert D.ensemble_placement(names, 2, 4) == ([2], 0)
    assert D.ensemble_placement(names, 6, 8) == ([2], 1)
    assert D.ensemble_placement(names, 1, 2) == ([1, 3], 0)
    assert D.ensemble_placement(names, 4, 5) == ([], None)
    # ... which is why the plan refuses such a world: the idle rank would "attack" with no backbone at all
    with pytest.raises(ValueError):
        D.EnsemblePlan(names, [1, 1, 1, 1], rank=4, world=5)
    with pytest.raises(ValueError):
        D.EnsemblePlan(names, [2, 2, 2, 2], rank=0, world=6)
    plan = D.EnsemblePlan(names, [2, 2, 2, 2], rank=6, world=8)
    assert plan.members == [2] and plan.replica == 1 and plan.layer_offsets == [4] and plan.n_layers_total == 8 and plan.active
    plan = D.EnsemblePlan(names, [1, 1, 1, 1], rank=1, world=2)
    assert plan.members == [1, 3] and plan.layer_offsets == [1, 3]


def test_loss_info_format():
    from i2v_b200 import attack_loop
    info = {}
    attack_loop.record_loss_info(info, ["a", "b"], np.array([1.9999996, 1.5], dtype=np.float32))
    assert info["a"][0] == {"cost": "1.9999996"} and info["b"][1] == {"cost": "1.5"}


def test_forced_relu_masks_isolate_the_decisions():
    """oracle.loops.forced_relu_masks: a float64 backward evaluated with the ReLU decisions of a float32 forward
    reproduces the float32 gradient to rounding level, and counts the decisions that differ."""
    import torch
    from i2v_b200 import backbones
    from oracle import loops as OL
    name, depth = "resnet", 1
    m32 = backbones.freeze_for_attack(backbones.seeded_random_init(backbones.arch_of(name), 0))
    g = torch.Generator().manual_seed(3)
    img = torch.randn(2, 3, 32, 32, generator=g)
    masks, acts32 = [], []
    hs = [mod.register_forward_hook(lambda mo, i, o: masks.append((o.detach() > 0).clone())) for mod in m32.modules()
          if isinstance(mod, torch.nn.ReLU)]
    tgt = backbones.find_target_layers(m32, name, depth)[0]
    hs.append(tgt.register_forward_hook(lambda mo, i, o: acts32.append(o)))
    x32 = img.clone().requires_grad_(True)
    m32(x32)
    for h in hs:
        h.remove()
    n_relu = 1 + 3 * 3        # stem + layer1's three bottlenecks: the decisions upstream of the hook
    up = torch.randn(acts32[0].shape, generator=g)
    (g32,) = torch.autograd.grad(acts32[0], x32, up)
    m64 = backbones.freeze_for_attack(backbones.seeded_random_init(backbones.arch_of(name), 0)).double()
    acts64 = []
    h = backbones.find_target_layers(m64, name, depth)[0].register_forward_hook(lambda mo, i, o: acts64.append(o))
    x64 = img.double().requires_grad_(True)
    with OL.forced_relu_masks(m64, masks[:n_relu]) as fm:
        m64(x64)
    h.remove()
    assert fm.i == n_relu and fm.total == sum(mk.numel() for mk in masks[:n_relu])
    assert fm.flips <= 2
    (g64,) = torch.autograd.grad(acts64[0], x64, up.double())
    assert ((g32.double() - g64).abs().max() / g64.abs().max()) < 1e-5
    # the context manager restores the modules
    assert all("forward" not in vars(mod) for mod in m64.modules())


def test_forced_maxpool_winners_reproduce_autograd():
    """oracle.loops.forced_relu_masks with the model's OWN ReLU decisions and max-pool winners must reproduce the plain
    forward and gradient exactly, and count zero flips; a swapped winner must be counted and change the gradient."""
    import torch.nn as nn
    import torch.nn.functional as F
    from oracle import loops as OL
    torch.manual_seed(0)
    model = nn.Sequential(nn.Conv2d(3, 4, 3, padding=1), nn.ReLU(), nn.MaxPool2d(3, 2, 1), nn.Conv2d(4, 4, 3, padding=1),
                          nn.ReLU(), nn.MaxPool2d(2, 2, 0, ceil_mode=True)).double()
    x = torch.randn(2, 3, 9, 9, dtype=torch.float64, requires_grad=True)
    y = model(x)
    (g,) = torch.autograd.grad(y.square().sum(), x)
    masks, pools, h = [], [], x.detach()
    for m in model:
        if isinstance(m, nn.MaxPool2d):
            h, idx = F.max_pool2d(h, m.kernel_size, m.stride, m.padding, 1, m.ceil_mode, return_indices=True)
            pools.append(idx)
        else:
            h = m(h)
            if isinstance(m, nn.ReLU):
                masks.append(h > 0)
    with OL.forced_relu_masks(model, masks, pools) as fm:
        y2 = model(x)
        (g2,) = torch.autograd.grad(y2.square().sum(), x)
    assert fm.flips == 0 and fm.pool_flips == 0 and fm.pool_total == sum(p.numel() for p in pools)
    assert torch.equal(y2, y) and torch.allclose(g2, g, rtol=0, atol=1e-15)
    bad = [p.clone() for p in pools]
    bad[0][0, 0, 1, 1] = bad[0][0, 0, 1, 1] + 1                      # neighbouring column of the same window row
    with OL.forced_relu_masks(model, masks, bad) as fm:
        (g3,) = torch.autograd.grad(model(x).square().sum(), x)
    assert fm.pool_flips >= 1 and not torch.equal(g3, g)     # (the changed value may move a winner downstream too)
    assert model[2].forward.__func__ is nn.MaxPool2d.forward          # patches removed


def test_image_main_host_logic(tmp_path):
    """The drop-in driver: reference flags and defaults (image_main.py:15-48), the adv_path naming (45), the
    contiguous --batch_nums/--batch_index slices (61-63) and the synthetic loader's tuple layout."""
    import image_main
    args = image_main.arg_parse(["--attack_method", "ImageGuidedFMDirection_Adam", "--step", "7", "--depth", "2",
                                 "--file_prefix", "run1", "--opt_path", str(tmp_path)])
    assert args.step_size == 0.004 and args.batch_size == 1 and args.direction_image_model == "resnet"
    assert args.adv_path == os.path.join(str(tmp_path), "Image-ImageGuidedFMDirection_Adam-7-run1")
    sl = image_main.SyntheticLoader(5, 2, 4, 16)
    assert len(sl) == 3
    batch, labels, names = sl.step(2)                      # ragged last batch
    assert tuple(batch.shape) == (1, 3, 4, 16, 16) and labels.tolist() == [4] and names == ["synthetic_00004"]
    batch, labels, names = sl.step(0)
    assert tuple(batch.shape) == (2, 3, 4, 16, 16) and labels.dtype == torch.int64
    v, _ = __import__("i2v_b200.synth", fromlist=["clip"]).clip(1, b=1, f=4, h=16, w=16)
    assert torch.equal(batch[1:], v)
    with pytest.raises(ValueError):
        image_main.build_attack(image_main.arg_parse(["--attack_method", "nope"]))
    # no silent substitutions: the default weight policy is the reference's `pretrained=True`, and without --synthetic
    # the reference's data pipeline must really be importable
    assert args.weights == os.environ.get("I2V_WEIGHTS", "pretrained") and not args.synthetic
    with pytest.raises(SystemExit) as exc:
        image_main.get_loader(args)
    assert "--synthetic" in str(exc.value)


def test_weight_policy_defaults_to_pretrained_and_records_the_source():
    """image_attacks.py:86-98 hard-requires pretrained=True; so does the default policy — random init is an opt-in."""
    saved = dict(backbones._WEIGHT_POLICY)
    try:
        backbones.set_weight_policy()
        assert backbones._WEIGHT_POLICY["mode"] == "pretrained"
        if backbones._pretrained_cached("squeezenet1_1") is None:
            with pytest.raises(RuntimeError) as exc:
                backbones.get_model("squeezenet", device=torch.device("cpu"))
            assert "I2V_WEIGHTS=random" in str(exc.value)
        backbones.set_weight_policy("random", 3)
        backbones.get_model("squeezenet", device=torch.device("cpu"))
        assert backbones.WEIGHT_SOURCE["squeezenet1_1"] == "random:seed=3"
    finally:
        backbones._WEIGHT_POLICY.update(saved)


def test_temporal_translation_host_side():
    """video_attacks.TemporalTranslation: variant kernels (video_attacks.py:51-78) and frame moves (93-146) — host logic only."""
    import random
    import video_attacks
    from i2v_b200 import synth
    from oracle import loops as OL
    model = synth.TinyTPNLike()
    for mode in ("gaussian", "linear", "random"):
        for kernlen in (3, 5, 9):
            atk = video_attacks.TemporalTranslation(model, {"kernlen": kernlen, "momentum": False, "weight": 0.5,
                                                            "move_type": "adj", "kernel_mode": mode})
            assert np.array_equal(atk._kernel_host, OL.tt_kernel(kernlen, mode))
            assert atk.cycle_move_list == list(range(-(kernlen // 2), kernlen // 2 + 1))
            assert tuple(atk.kernel.shape) == (1, kernlen) and abs(float(atk.kernel.sum()) - 1) < 1e-6
            assert atk.step_size == atk.epsilon / atk.steps and atk.frames == 32
    atk.move_type = "adj"
    assert [atk._effective_move(m, 32) for m in (-3, 0, 2, 35)] == [-3, 0, 2, 3]
    atk.move_type = "large"            # 107-120: |m| -> (|m| + frames/2 - 1) mod frames, 0 stays
    assert [atk._effective_move(m, 32) for m in (-3, 0, 2)] == [-18, 0, 17]
    atk.move_type = "random"           # 122-135: one randint(0, 100) per non-zero move
    random.seed(4)
    got = [atk._effective_move(m, 32) for m in (-1, 0, 1)]
    random.seed(4)
    want = [-(random.randint(0, 100) % 32), 0, random.randint(0, 100) % 32]
    assert got == want
    with pytest.raises(ValueError):
        video_attacks.TemporalTranslation(model, {"kernlen": 3, "momentum": False, "weight": 0.5, "move_type": "adj",
                                                  "kernel_mode": "cubic"})


def test_tap_and_ilaf_constructors():
    import base_attacks
    import image_attacks
    from i2v_b200 import synth
    model = synth.TinyTPNLike()
    tap = base_attacks.TAP(model, {"kernlen": 3, "temporal_kernlen": 5, "eta": 1e3, "conv3d": True, "model_type": "tpn"})
    assert tuple(tap.stack_2d_kernel.shape) == (3, 1, 3, 3) and tuple(tap.stack_3d_kernel.shape) == (3, 1, 5, 3, 3)
    assert abs(float(tap.stack_3d_kernel[0].sum()) - 1) < 1e-6 and tap.step_size == tap.epsilon / tap.steps
    assert tap._find_target_layer() == [model.layer1, model.layer2]
    with pytest.raises(ValueError):
        base_attacks.TAP(model, {"kernlen": 4, "temporal_kernlen": 3, "conv3d": True, "model_type": "tpn"})
    with pytest.raises(ValueError):
        base_attacks.TAP(synth.TinyTPNLike(), {"kernlen": 3, "temporal_kernlen": 3, "conv3d": True, "model_type": "resnet"})
    il = image_attacks.ILAF(model, "tpn")
    assert il._find_target_layer() is model.layer2 and il.steps == 60 and il.step_size == 0.005
    assert image_attacks.ILAF(model, "other", target_layers=[model.layer1])._find_target_layer() == [model.layer1]
    with pytest.raises(ValueError):
        image_attacks.ILAF(model, "resnet")


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's own CPU loop on the host cores) prints exactly one JSON line with the
    contract's keys; under torchrun only rank 0 prints."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    r = json.loads(lines[0])
    assert r["impl"] == "reference" and r["metric"] == "attack_frame_steps_per_sec" and r["unit"] == "frame-steps/s"
    assert r["higher_is_better"] is True and r["value"] > 0 and r["gpu_launches"] == 0 and r["vs_baseline"] is None
    assert r["cpu_baseline"]["kind"] in ("reference", "port") and r["cpu_baseline"]["cores"] >= 1
    assert r["cpu_baseline"]["weight_grads"] is True and "kind=" + r["cpu_baseline"]["kind"] in r["cpu_baseline"]["sample"]
    assert r["e2e"] == {"value": r["value"], "unit": r["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in r["config"]
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2"],
                         capture_output=True, text=True, timeout=600, cwd=root, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_attack_driver_host_side(tmp_path):
    """attack.py: flags, output directory name (attack.py:54-57) and the class dispatch (76-84) — no GPU needed."""
    import attack
    import base_attacks
    import video_attacks
    from i2v_b200 import synth
    args = attack.arg_parse(["--model", "tiny", "--attack_method", "TIFGSM", "--step", "3", "--file_prefix", "x",
                             "--opt_path", str(tmp_path)])
    assert args.adv_path == os.path.join(str(tmp_path), "tiny-TIFGSM-3-x") and args.attack_type == "image"
    model = synth.TinyVideoNet()
    atk = attack.build_attack(args, model)
    assert isinstance(atk, base_attacks.TIFGSM) and atk.steps == 3
    args = attack.arg_parse(["--attack_type", "video", "--attack_method", "TemporalTranslation", "--kernlen", "7",
                             "--augmentation_weight", "0.3", "--move_type", "large", "--kernel_mode", "linear"])
    atk = attack.build_attack(args, model)
    assert isinstance(atk, video_attacks.TemporalTranslation) and atk.kernlen == 7 and atk.weight == 0.3
    assert atk.move_type == "large" and atk.momentum is False and len(atk.cycle_move_list) == 7
    assert attack.standin_model("r3d_18", 7).fc.out_features == 7


def test_fine_tune_driver_host_side(tmp_path):
    """image_fine_tune_attack.py: AdvDataset over saved pairs (16-37) and the flags — no GPU needed."""
    import image_fine_tune_attack as ft
    rng = np.random.default_rng(0)
    for label in (7, 12):
        np.save(os.path.join(str(tmp_path), "%d-adv" % label), rng.standard_normal((3, 4, 6, 6)).astype(np.float32))
        np.save(os.path.join(str(tmp_path), "%d-ori" % label), rng.standard_normal((3, 4, 6, 6)).astype(np.float32))
    ds = ft.AdvDataset(str(tmp_path), str(tmp_path))
    assert len(ds) == 2
    vid, ori, label = ds[0]
    assert tuple(vid.shape) == tuple(ori.shape) == (1, 3, 4, 6, 6) and label.dtype == torch.int64
    assert int(label) in (7, 12) and np.array_equal(vid[0].numpy(), np.load(os.path.join(str(tmp_path), "%d-adv.npy" % int(label))))
    args = ft.arg_parse(["--white_model", "tpn_tiny", "--used_adv", "a", "--used_ori", "b", "--opt_path", "c", "--synthetic"])
    assert args.attack_method == "ILAF" and args.steps == 60 and args.step_size == 0.005 and args.synthetic


def test_image_main_ucf101_host_side(tmp_path):
    """image_main_ucf101.py = image_main.py with the reference's four differences (step default 10, UCF loader /
    labels mod 101, ENS gets steps, video_names = str(label tensor)) — flags and loader only, no GPU."""
    import image_main_ucf101 as u
    args = u.arg_parse(["--synthetic", "--num_clips", "205", "--batch_size", "2", "--frames", "2", "--side", "8", "--opt_path",
                        str(tmp_path), "--file_prefix", "p"])
    assert args.step == 10 and args.adv_path == os.path.join(str(tmp_path), "Image-ImageGuidedFMDirection_Adam-10-p")
    assert u.arg_parse(["--step", "7"]).step == 7 and u.arg_parse(["--step=9"]).step == 9
    n, get_step = u.get_loader(args)
    assert n == 103
    vids, labs, names = get_step(101)                       # clips 202, 203 -> labels 0, 1 (mod 101)
    assert tuple(vids.shape) == (2, 3, 2, 8, 8) and labs.tolist() == [0, 1]
    vids, labs, names = get_step(102)
    assert vids.shape[0] == 1 and labs.tolist() == [204 % 400 % 101]


def test_attack_ucf101_host_side(tmp_path):
    """attack_ucf101.py: output directory names (57-59), 101 classes, fixed TemporalTranslation parameters (87)."""
    import attack_ucf101 as au
    import video_attacks
    from i2v_b200 import synth
    args = au.arg_parse(["--model", "tiny", "--attack_method", "BIM", "--step", "4", "--opt_path", str(tmp_path)])
    assert args.adv_path == os.path.join(str(tmp_path), "UCF101_Image-tiny-BIM-4-") and args.num_classes == 101
    args = au.arg_parse(["--model", "tiny", "--attack_type", "video", "--attack_method", "TemporalTranslation", "--kernlen", "5",
                         "--opt_path", str(tmp_path)])
    assert args.adv_path.endswith("UCF101_Video-tiny-TemporalTranslation-10-")
    atk = au.build_attack(args, synth.TinyVideoNet())
    assert isinstance(atk, video_attacks.TemporalTranslation) and atk.kernlen == 15 and atk.weight == 1.0 and atk.momentum is False
