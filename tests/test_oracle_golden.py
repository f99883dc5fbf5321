"""Pin the CPU oracle against outputs of the UNMODIFIED reference classes (tests/golden/*.npz, made by
oracle/make_golden.py from /root/reference).  The reference holds no tests or golden vectors of its
own, so these fixtures are the pin.

Tolerances (measured, see DESIGN.md "parity"): the per-pixel update arithmetic is IEEE-exact in the
oracle, while torch's CPU Adam takes sqrt through MKL VML which is off by 1 ulp on ~0.7 % of elements;
the softmax of K2 uses a float64 exp where torch uses a float32 one.  Everything else is bit-exact.
"""
import numpy as np
import pytest
import torch

from conftest import ulp_diff
from i2v_b200 import backbones, synth
from oracle import loops as OL
from oracle import oracle as O

EPS = 16 / 255


def _hooked(names, depths):
    out = []
    for n in names:
        model = backbones.seeded_random_init(backbones.arch_of(n), 0)
        d = depths[n] if isinstance(depths, dict) else depths
        out.append(OL.HookedModel(model, backbones.family_of(n), d))
    return out


def _check_weights(gold, hooked):
    for row, hm in zip(gold["weight_checksums"], hooked):
        p = next(hm.model.parameters()).detach().double()
        got = np.array([p.sum().item(), p.abs().sum().item(), float(p.flatten()[0])])
        if not np.allclose(got, row, rtol=0, atol=0):
            pytest.skip("torchvision random init differs from the one the fixture was made with")


# ------------------------------------------------------------------ per-kernel (teacher-forced) pins
def test_adam_compose_matches_reference_step(golden):
    """K3a oracle vs the reference's optimizer.step(): feed the reference's own modifier gradient of
    step 1 and compare (m, v, modifier) after the step."""
    g = golden("i2v_resnet50_d2_32")
    videos = torch.from_numpy(g["videos"])
    frames = OL._frames(videos).numpy()
    inner = frames.shape[-1] * frames.shape[-2]
    x = O.denorm(frames, inner)
    mod0 = np.full_like(x, np.float32(0.01 / 255))
    # the reference's p.grad is dcost/dmodifier = (dcost/dtrue_image)/std * masks; undo the /std so the
    # oracle sees what the K3a kernel sees (exact: multiply back is not exact, so rebuild from g*std)
    std = O.STD[None, :, None, None]
    g_true = g["g_mod_first"] * std
    # multiplication by std then division by std is not an identity in f32: check on the elements where it is
    back = (g_true / std).astype(np.float32)
    exact = back == g["g_mod_first"]
    assert exact.mean() > 0.7
    m, v, mod, _ = O.adam_compose(g_true, np.zeros_like(x), np.zeros_like(x), mod0, x, EPS, inner, 1, float(g["step_size"]))
    assert np.array_equal(m[exact], g["m_first"][exact])
    assert np.array_equal(v[exact], g["v_first"][exact])
    d = ulp_diff(mod[exact], g["mod_first"][exact])
    assert d.max() <= 2 and (d > 0).mean() < 0.02, (d.max(), (d > 0).mean())   # MKL sqrt: <=2 ulp on <2 % of elements


def test_adam_compose_matches_reference_last_step(golden):
    """Same, for the last recorded step (non-trivial m, v, bias corrections and clamp masks)."""
    for name in ("i2v_resnet50_d2_32", "i2v_vgg_d3_32"):
        g = golden(name)
        frames = OL._frames(torch.from_numpy(g["videos"])).numpy()
        inner = frames.shape[-1] * frames.shape[-2]
        x = O.denorm(frames, inner)
        std = O.STD[None, :, None, None]
        g_true = g["g_mod_last"] * std
        exact = (g_true / std).astype(np.float32) == g["g_mod_last"]
        m, v, mod, _ = O.adam_compose(g_true, g["m_before_last"], g["v_before_last"], g["mod_before_last"], x, EPS, inner,
                                      int(g["steps"]), float(g["step_size"]))
        assert np.array_equal(m[exact], g["m_last"][exact])
        assert np.array_equal(v[exact], g["v_last"][exact])
        # the update q = (-ss*m)/den carries MKL's 1-ulp sqrt; where mod + q cancels, ulps of the sum
        # are meaningless, so bound the absolute difference by 2 ulp of the largest addend (|q| <~ 0.015)
        d = np.abs(mod[exact] - g["mod_last"][exact])
        assert d.max() <= 2e-9 and (d > 0).mean() < 0.02, (d.max(), (d > 0).mean())


def test_sign_step_and_norm_grads_match_reference(golden):
    g = golden("base_tiny3d")
    ng = g["norm_grads_in"]
    norm = O.frame_absmean(ng)
    got = ng / norm[:, None, :, None, None]
    assert np.allclose(got, g["norm_grads_frame"], rtol=3e-7, atol=0)
    normc = O.frame_absmean(ng, clip_level=True)
    assert np.allclose(ng / normc[:, None, None, None, None], g["norm_grads_clip"], rtol=3e-7, atol=0)


# ------------------------------------------------------------------ full-loop pins
def _close_adv(adv, ref_adv, min_equal=0.85):
    """End state after a few free-running steps.  The only arithmetic difference between the oracle and
    the reference is MKL's 1-ulp sqrt in torch's CPU Adam (pinned above); through the chaotic dynamics
     it leaves ~5-10 % of the final pixels off by a few 1e-5 after 3 steps."""
    d = np.abs(adv - ref_adv)
    assert (d == 0).mean() > min_equal, (d == 0).mean()
    assert (d <= 1e-5).mean() > 0.99, (d <= 1e-5).mean()
    assert d.max() < 2e-3, d.max()          # 0.05 % of the normalised pixel range, << eps/std = 0.27


def _loop_case(golden, name, names, depths, **kw):
    g = golden(name)
    hooked = _hooked(names, depths)
    _check_weights(g, hooked)
    adv, cost, weights, coeffs = OL.image_guided_loop(hooked, g["videos"], float(g["epsilon"]), int(g["steps"]),
                                                      float(g["step_size"]), cos_mode="torch", **kw)
    return g, adv, cost, weights, coeffs


def test_i2v_loop_matches_reference(golden):
    g, adv, cost, _, _ = _loop_case(golden, "i2v_resnet50_d2_32", ["resnet"], 2)
    assert np.allclose(cost, g["cost"], rtol=1e-6)
    _close_adv(adv, g["adv"])


def test_port_with_weight_grads_is_the_same_loop(golden):
    """bench.py's CPU arm runs the port with weight_grads=True (`cost.backward()`, every backbone parameter requiring grad,
    image_attacks.py:351-353): same costs and adversarial clip bit for bit as the pruned-autograd variant the parity tests
    use, and the weight gradients really are computed (the cost the reference pays)."""
    g = golden("i2v_resnet50_d2_32")
    hooked = _hooked(["resnet"], 2)
    adv0, cost0, _, _ = OL.image_guided_loop(hooked, g["videos"], float(g["epsilon"]), 2, float(g["step_size"]))
    assert hooked[0].model.conv1.weight.grad is None
    hooked = _hooked(["resnet"], 2)
    adv1, cost1, _, _ = OL.image_guided_loop(hooked, g["videos"], float(g["epsilon"]), 2, float(g["step_size"]), weight_grads=True)
    assert np.array_equal(adv0, adv1) and np.array_equal(cost0, cost1)
    wg = hooked[0].model.layer2[0].conv1.weight.grad
    assert wg is not None and float(wg.abs().sum()) > 0


def test_config1_60step_fixture_first_steps(golden):
    """The benchmark-size fixture (32 frames x 3x224x224, 60 steps, unmodified reference class): the oracle port reproduces
    its first two costs and the sign pattern of the first dcost/dmodifier (the port runs the same torch ops in the same
    order, so the gradient is bit-identical; the later steps are pinned on the GPU side)."""
    from i2v_b200 import synth
    g = golden("i2v_resnet50_d2_224_60step")
    assert g["cost"].shape == (60,) and g["delta16"].shape == (1, 3, 32, 224, 224)
    assert np.abs(g["delta16"].astype(np.float32)).max() <= (16 / 255) / 0.224 * 1.001
    videos, _ = synth.clip(0, b=1, f=int(g["frames"]), h=int(g["side"]), w=int(g["side"]))
    hooked = _hooked(["resnet"], 2)
    _check_weights(g, hooked)
    taps = {}
    _, cost, _, _ = OL.image_guided_loop(hooked, videos.numpy(), float(g["epsilon"]), 2, float(g["step_size"]),
                                         tap=lambda i, d: taps.setdefault(i, d["g"].copy()))
    assert np.allclose(cost, g["cost"][:2], rtol=1e-6)
    gm = (taps[0] / O.STD[None, :, None, None]).reshape(-1)
    n = gm.size
    # the fixture holds dcost/dMODIFIER (zero where the [0,1] clamp is active, image_attacks.py:331); the port taps
    # dcost/dtrue_image and leaves the clamp masks to the Adam block: compare where the clamp is inactive
    x = O.denorm(OL._frames(videos).numpy(), int(g["side"]) ** 2).reshape(-1)
    live = (x + np.float32(OL.INIT_MODIFIER) <= 1) & (x + np.float32(OL.INIT_MODIFIER) >= 0)
    pos, neg = np.unpackbits(g["g_first_pos_bits"])[:n].astype(bool), np.unpackbits(g["g_first_neg_bits"])[:n].astype(bool)
    assert live.mean() > 0.999 and not (pos | neg)[~live].any()
    assert np.array_equal((gm > 0)[live], pos[live]) and np.array_equal((gm < 0)[live], neg[live])
    assert abs(np.abs(gm[live]).max() / float(g["g_first_max"]) - 1) < 1e-6


def test_i2v_vgg_loop_matches_reference(golden):
    g, adv, cost, _, _ = _loop_case(golden, "i2v_vgg_d3_32", ["vgg"], 3)
    assert np.allclose(cost, g["cost"], rtol=1e-6)
    _close_adv(adv, g["adv"])


def test_ens_loop_matches_reference(golden):
    names = ["resnet", "vgg", "squeezenet", "alexnet"]
    g, adv, cost, _, _ = _loop_case(golden, "ens_4models_64", names, {"resnet": 2, "vgg": 3, "squeezenet": 2, "alexnet": 3})
    assert np.allclose(cost, g["cost"], rtol=1e-6)
    _close_adv(adv, g["adv"])


def test_aens_loop_matches_reference(golden):
    names = ["resnet", "vgg", "squeezenet", "alexnet"]
    g, adv, cost, weights, coeffs = _loop_case(golden, "aens_4models_64", names, {n: [2, 3] for n in names},
                                               adaptive=True, coeffs=np.ones(8, np.float32), momentum=0.5)
    assert np.allclose(weights, g["weights"], rtol=2e-6)
    assert np.allclose(coeffs, g["coeffs_after"], rtol=2e-6)
    assert np.allclose(cost, g["cost_saved"], rtol=2e-6)
    _close_adv(adv, g["adv"], min_equal=0.6)   # + K2's float64 exp vs torch's float32 softmax


def test_aens_coef_ce_loop_matches_reference(golden):
    names = ["resnet", "squeezenet"]
    g, adv, cost, weights, _ = _loop_case(golden, "aens_ce_2models_64", names, {"resnet": [1, 2], "squeezenet": [2, 3]},
                                          adaptive=True, coeffs=np.ones(4, np.float32), momentum=0.0, coef_CE=True)
    assert np.allclose(weights, g["weights"], rtol=2e-6)
    assert np.allclose(cost, g["cost_saved"], rtol=2e-6)
    _close_adv(adv, g["adv"], min_equal=0.6)   # + K2's float64 exp vs torch's float32 softmax


@pytest.mark.parametrize("fixture,name,depth", [("dr_resnet50_d2_32", "resnet", 2), ("dr_vgg_d2_32", "vgg", 2)])
def test_dispersion_loop_matches_reference(golden, fixture, name, depth):
    """Dispersion Reduction (image_attacks.py:129-234): the loop restatement against the unmodified class, and the
    float64 analytic std gradient against torch's autograd of Tensor.std()."""
    g = golden(fixture)
    hooked = _hooked([name], depth)
    _check_weights(g, hooked)
    adv, cost = OL.dispersion_loop(hooked, g["videos"], float(g["epsilon"]), int(g["steps"]), float(g["step_size"]))
    assert np.allclose(cost, g["cost"], rtol=1e-6)
    _close_adv(adv, g["adv"])
    a = torch.randn(3, 8, 5, 5, dtype=torch.float64, generator=torch.Generator().manual_seed(1)).relu().requires_grad_(True)
    sd = a.std()
    (ga,) = torch.autograd.grad(sd, a)
    sd_o, _, g_o = O.std_loss_grad_f64(a.detach().numpy())
    assert abs(sd_o - float(sd)) <= 1e-14 and np.allclose(g_o, ga.numpy(), rtol=1e-12, atol=1e-18)


def test_base_attacks_match_reference(golden):
    g = golden("base_tiny3d")
    model = synth.TinyVideoNet()
    p = next(model.parameters()).detach().double()
    if not np.allclose([p.sum().item(), p.abs().sum().item(), float(p.flatten()[0])], g["weight_checksums"][0], rtol=0, atol=0):
        pytest.skip("random init differs from the fixture's")
    labels = torch.from_numpy(g["labels"])
    assert np.array_equal(OL.fgsm(model, g["videos"], labels), g["fgsm"])
    assert np.array_equal(OL.bim(model, g["videos"], labels, steps=3), g["bim3"])
    # 'targeted' only flips the sign of the loss: BIM.forward never calls _transform_label (base_attacks.py:272-295)
    assert np.array_equal(OL.bim(model, g["videos"], labels, steps=2, targeted=-1), g["bim2_targeted"])
    mi = OL.mifgsm(model, g["videos"], labels, steps=3)
    assert (mi == g["mifgsm3"]).mean() > 0.9999   # norm is a float64 mean here, a float32 one in torch


def test_base_variants_match_reference(golden):
    """DIFGSM / TIFGSM / SGM / SIM / TIFGSM3D restatements (oracle/loops.py) against the unmodified classes.  The
    update block is bit-exact given the gradient; what differs is float32 summation order inside torch ops on two
    code paths of the same library, so the bar is the fraction of identical pixels (sign steps amplify a gradient
    that rounds to the other side of zero into a full step)."""
    import random
    g = golden("base_variants")
    model = synth.TinyVideoNet()
    p = next(model.parameters()).detach().double()
    if not np.allclose([p.sum().item(), p.abs().sum().item(), float(p.flatten()[0])], g["weight_checksums"][0], rtol=0, atol=0):
        pytest.skip("random init differs from the fixture's")
    labels = torch.from_numpy(g["labels"])
    v = g["videos"]
    for mom in (False, True):
        tag = "_mom" if mom else ""
        assert (OL.tifgsm(model, v, labels, steps=3, momentum=mom) == g["tifgsm3" + tag]).mean() > 0.999
        assert (OL.sim(model, v, labels, steps=2, momentum=mom) == g["sim2" + tag]).mean() > 0.999
        assert (OL.tifgsm3d(model, v, labels, steps=2, momentum=mom) == g["tifgsm3d2" + tag]).mean() > 0.999
        assert (OL.sgm(synth.TinyReluVideoNet(), v, labels, steps=3, momentum=mom) == g["sgm3" + tag]).mean() > 0.999
        di_videos, _ = synth.clip(5, b=1, f=2, h=224, w=224)
        random.seed(11)
        torch.manual_seed(11)
        adv = OL.difgsm(model, di_videos.numpy(), labels, steps=4, momentum=mom)
        delta = adv - di_videos.numpy()
        assert (np.abs(delta - g["difgsm4_delta16" + tag].astype(np.float32)) < 2e-3).mean() > 0.999
    # the SGM hooks do change the result (gamma = 0.5 through three ReLU modules), i.e. the fixture exercises them
    assert (g["sgm3"] != g["bim3_relu"]).mean() > 0.01


def test_eps_bound_never_violated(golden):
    """clamp(modifier, ±eps) is exact; on adv - x the f32 add may exceed eps by 1 ulp(1.0)."""
    g = golden("i2v_resnet50_d2_32")
    frames = OL._frames(torch.from_numpy(g["videos"])).numpy()
    inner = frames.shape[-1] * frames.shape[-2]
    x = O.denorm(frames, inner)
    adv01 = O.denorm(OL._frames(torch.from_numpy(np.ascontiguousarray(g["adv"]))).numpy(), inner)
    assert np.abs(adv01 - x).max() <= np.float32(EPS) + 2 * np.finfo(np.float32).eps
    assert adv01.min() >= -1e-6 and adv01.max() <= 1 + 1e-6


# ---- TemporalTranslation / TAP / ILAF ----------------------------------------------
def test_temporal_pieces_match_their_torch_statements():
    """The C pieces of the TemporalTranslation restatement against the reference's own statements: `_cycle_move`
    (video_attacks.py:93-105) is a roll, `_grad_augmentation` (163-177) two [1,D] x [D,M] products and a blend."""
    rng = np.random.default_rng(5)
    adv = rng.standard_normal((2, 3, 8, 3, 5)).astype(np.float32)
    moves = [-2, -1, 0, 1, 2, 11, -9]
    st = O.temporal_shift_stack(adv, moves)
    for d, m in enumerate(moves):
        assert np.array_equal(st[d], OL._cycle(torch.from_numpy(adv), m, 8).numpy())
        assert np.array_equal(st[d], np.roll(adv, m, axis=2))
    D = 5
    g = rng.standard_normal((D, 1, 3, 8, 3, 5)).astype(np.float32)
    k = OL.tt_kernel(D, "gaussian")
    mv = [-2, -1, 0, 1, 2]
    out = O.temporal_combine(g, k, mv, 0.3)
    gt = torch.from_numpy(g)
    diff = torch.stack([OL._cycle(gt[i], -m, 8) for i, m in enumerate(mv)])
    kt = torch.from_numpy(k)[None]
    ref = (1 - 0.3) * torch.matmul(kt, gt.reshape(D, -1)) + 0.3 * torch.matmul(kt, diff.reshape(D, -1))
    assert np.allclose(out.reshape(-1), ref.numpy().reshape(-1), rtol=2e-6, atol=1e-7)


def test_sign_descent_and_ila_pieces_match_autograd():
    """K3d restated (image_attacks.py:615-617 through the compose block 582-585) and the ILAF layer loss (596-611)
    against torch autograd of the reference's own expressions."""
    rng = np.random.default_rng(9)
    eps, step = 16 / 255, 0.005
    x = rng.random((1, 3, 4, 6, 6)).astype(np.float32)
    mod = (rng.standard_normal(x.shape) * 0.08).astype(np.float32)       # some outside +-eps, some sums outside [0,1]
    g = rng.standard_normal(x.shape).astype(np.float32)
    g[0, 0, 0, 0, :3] = 0.0
    inner = 4 * 6 * 6
    mod2, img2 = O.sign_descent_compose(g, mod, x, eps, step, inner)
    m_t = torch.from_numpy(mod.copy()).requires_grad_(True)
    mean = torch.tensor(O.MEAN)[None, :, None, None, None]
    std = torch.tensor(O.STD)[None, :, None, None, None]
    ti = (torch.clamp(torch.from_numpy(x) + torch.clamp(m_t, min=-eps, max=eps), min=0, max=1) - mean) / std
    (gm,) = torch.autograd.grad(ti, m_t, grad_outputs=torch.from_numpy(g))
    want = m_t.detach() - step * gm.sign()
    assert np.array_equal(mod2, want.numpy())
    ti2 = (torch.clamp(torch.from_numpy(x) + torch.clamp(want, min=-eps, max=eps), min=0, max=1) - mean) / std
    assert np.array_equal(img2, ti2.numpy())

    f = rng.standard_normal(500).astype(np.float32)
    o = rng.standard_normal(500).astype(np.float32)
    d0 = rng.standard_normal(500)
    n0 = float(np.linalg.norm(d0)) * 1.7
    d0 = (d0 / np.linalg.norm(d0)).astype(np.float32)
    loss, grad = O.ila_loss_grad_f64(f, o, d0, n0)
    ft = torch.from_numpy(f).double().requires_grad_(True)
    sd = ft - torch.from_numpy(o).double()
    sn = torch.norm(sd, p=2)
    ref = -(0.5 * sn / n0 + torch.mm(torch.from_numpy(d0).double().view(1, -1), (sd / sn).view(1, -1).transpose(1, 0)))
    (gr,) = torch.autograd.grad(ref.sum(), ft)
    assert abs(loss - float(ref)) <= 1e-12 and np.allclose(grad, gr.numpy(), rtol=1e-10, atol=1e-14)


def test_video_variant_loops_match_reference(golden):
    """TemporalTranslation, TAP and ILAF restatements (oracle/loops.py) against the UNMODIFIED reference classes
    (tests/golden/video_variants.npz, oracle/make_golden.py:run_video_variants)."""
    g = golden("video_variants")
    videos, labels = g["videos"], torch.from_numpy(g["labels"])
    p = next(synth.TinyTPNLike().parameters()).detach().double()
    if not np.allclose([p.sum().item(), p.abs().sum().item(), float(p.flatten()[0])], g["weight_checksums"][0], rtol=0, atol=0):
        pytest.skip("random init differs from the fixture's")

    def frac(a, b):
        return float((np.abs(a - b) < 1e-6).mean())

    m = synth.TinyTPNLike()
    assert frac(OL.temporal_translation(m, videos, labels, 5, 0.5, steps=3), g["tt3_k5"]) == 1.0
    assert frac(OL.temporal_translation(m, videos, labels, 5, 0.5, momentum=True, steps=3), g["tt3_k5_mom"]) == 1.0
    assert frac(OL.temporal_translation(m, videos, labels, 9, 0.3, kernel_mode="linear", steps=2), g["tt2_k9_linear"]) == 1.0
    import random
    for move_type, key in (("large", "tt2_k5_large"), ("random", "tt2_k5_randommove")):      # uniform kernel, momentum, w = 0.7
        random.seed(21)
        got = OL.temporal_translation(m, videos, labels, 5, 0.7, momentum=True, kernel_mode="random", steps=2, move_type=move_type)
        assert frac(got, g[key]) == 1.0, move_type
    for conv3d, tag in ((True, "3d"), (False, "2d")):
        adv, info = OL.tap(m, [m.layer1, m.layer2], videos, labels, conv3d=conv3d, steps=3)
        assert frac(adv, g["tap3_" + tag]) == 1.0
        want = g["tap3_%s_last_losses" % tag]
        assert np.allclose([info[-1][0], info[-1][1], float(info[-1][2][0])], want, rtol=1e-5)
    out, clip, costs = OL.ilaf(m, [m.layer2], g["tt3_k5"], videos, steps=4)
    assert frac(out, g["ilaf4"]) == 1.0
    assert np.allclose(costs, g["ilaf4_costs"], rtol=1e-6)
    # 627-629 reinterpret [b,3,f,h,w] as [b,f,3,h,w]: the returned tensor is NOT the clip (a reference defect that is kept)
    assert not np.array_equal(out, clip) and np.array_equal(np.sort(out.reshape(-1)), np.sort(clip.reshape(-1)))
