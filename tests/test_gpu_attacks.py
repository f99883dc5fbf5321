"""GPU parity of the attack classes (drop-in API) against the golden fixtures produced by the
UNMODIFIED reference classes, and against the CPU oracle loops.

What can and cannot match: the reference does not reproduce itself across reduction
orders — at step 1 the gradient is a cancellation-level quantity.  The checks are therefore staged:
teacher-forced per-step state (tight), costs (1e-5 relative), gradient-sign agreement restricted to
gradients above the noise level, and the end state after a few free-running steps with the measured
oracle-vs-reference spread as the yardstick.  The eps-ball / [0,1] bounds are exact properties.
"""
import os

import numpy as np
import pytest
import torch

import TPAMI_attack
import base_attacks
import image_attacks
from i2v_b200 import attack_loop, backbones, capi, engines, synth
from oracle import loops as OL
from oracle import oracle as O

pytestmark = pytest.mark.gpu
EPS = 16 / 255
ENGINES = ["cudnn"]
try:   # the native engine joins the same tests once it exists
    from i2v_b200 import engine_native  # noqa: F401
    ENGINES.append("native")
except ImportError:
    pass


def _frames_np(clip):
    return OL._frames(torch.as_tensor(np.ascontiguousarray(clip))).numpy()


def _bounds_ok(videos, adv):
    fr = _frames_np(videos)
    inner = fr.shape[-1] * fr.shape[-2]
    x = O.denorm(fr, inner)
    a = O.denorm(_frames_np(adv), inner)
    assert np.abs(a - x).max() <= np.float32(EPS) + 2 * np.finfo(np.float32).eps
    assert a.min() >= -1e-6 and a.max() <= 1 + 1e-6


def _sign_agreement(g, g_ref, tau_frac=1e-3):
    big = np.abs(g_ref) > tau_frac * np.abs(g_ref).max()
    return (np.sign(g[big]) == np.sign(g_ref[big])).mean()


STATS = {}


def _record(key, **kw):
    """Parity numbers are also written to gpurun_out/parity_stats.json (copied into profiles/ and DESIGN.md)."""
    import json
    import os
    STATS[key] = {k: (float(v) if (v is not None and np.isscalar(v)) else v) for k, v in kw.items()}
    prev = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_stats.json")
    if len(STATS) == 1 and os.path.isfile(prev):            # keep figures recorded by an earlier pytest process of this run
        try:
            STATS.update({k: v for k, v in json.load(open(prev)).items() if k not in STATS})
        except ValueError:
            pass
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_stats.json"), "w") as f:
        json.dump(STATS, f, indent=1, sort_keys=True)


def _grad_scores(g, g64, tau=1e-3):
    """(sign agreement on |g64| > tau*max, relative L2 error) of a float32 gradient vs the float64 arbiter."""
    big = np.abs(g64) > tau * np.abs(g64).max()
    sign = (np.sign(g[big]) == np.sign(g64[big])).mean()
    rel = np.linalg.norm(g.astype(np.float64) - g64) / np.linalg.norm(g64)
    return sign, rel


def _hooked_cpu(names, depths):
    out = []
    for n in names:
        d = depths[n] if isinstance(depths, dict) else depths
        out.append(OL.HookedModel(backbones.seeded_random_init(backbones.arch_of(n), 0), backbones.family_of(n), d))
    return out


def _teacher_forced_steps(tag, names, depths, taps, videos, weights_per_step=None):
    """For every step of OUR free-running trajectory, recompute the gradient of that step's true_image on
    the CPU in float64 (arbiter) and float32 (the reference's arithmetic) and score both against the
    arbiter.  Ours must be at least as close as the reference's own float32 arithmetic is, up to a
    factor 4 on the L2 error and 0.5 % on sign agreement.

    ReLU decisions: 1[z > 0] of a pre-activation within rounding distance of 0 is decided differently by
    different correct float32 forward passes (ours, torch-f32, cuDNN), and one decision changes the gradient
    by O(1) over the element's receptive field.  When the engine exposes its decisions (native engine) the
    float64 arbiter is evaluated WITH those decisions (oracle.loops.forced_relu_masks; the winners of every max
    pooling are forced the same way — two window entries within rounding distance are the same discontinuity), so the tight L2 bound
    tests the gradient arithmetic; how many decisions differ from float64's own is bounded separately
    (<= 2e-5 of all activations) and the free-decision error is recorded next to it."""
    frames = OL._frames(torch.as_tensor(videos))
    for step, tp in sorted(taps.items()):
        ti = tp["true_image"].cpu().numpy()
        w = None if weights_per_step is None else weights_per_step[step]
        c64, g64, cos64 = OL.teacher_forced_grad(_hooked_cpu(names, depths), frames, ti, torch.float64, w)
        c32, g32, _ = OL.teacher_forced_grad(_hooked_cpu(names, depths), frames, ti, torch.float32, w)
        ours = tp["g"].cpu().numpy()
        s_free, r_free = _grad_scores(ours, g64)
        s_r, r_r = _grad_scores(g32.astype(np.float32), g64)
        masks = tp.get("relu_masks")
        flips = total = None
        if masks is not None and all(m is not None for m in masks):
            pools = tp.get("pool_indices")
            if pools is not None and any(p is None for p in pools):
                pools = None
            _, g64m, _, (flips, total) = OL.teacher_forced_grad(_hooked_cpu(names, depths), frames, ti, torch.float64, w,
                                                               relu_masks=masks, pool_indices=pools)
            s_o, r_o = _grad_scores(ours, g64m)
        else:
            s_o, r_o = s_free, r_free
        cos_err = np.abs(tp["cos"].cpu().numpy() - cos64).max() / np.abs(cos64).max()
        _record("%s/step%d" % (tag, step), sign_ours=s_o, sign_torch_f32=s_r, relL2_ours=r_o, relL2_torch_f32=r_r,
                relL2_ours_free_relu=r_free, sign_ours_free_relu=s_free, relu_flips=flips, relu_total=total,
                cos_rel_err=cos_err, gmax=float(np.abs(g64).max()))
        assert cos_err <= 1e-5, cos_err                       # north star: cosine within 1e-5 relative
        # cuDNN is free to pick Winograd / FFT algorithms (it does for VGG's 3x3 stacks), which are a few
        # times less accurate than the direct float32 convolution oneDNN runs on the CPU
        if flips is not None:
            assert r_o <= max(4.0 * r_r, 1e-4), (step, r_o, r_r)
        assert s_o >= s_r - 5e-3, (step, s_o, s_r)
        assert s_free >= s_r - 5e-3, (step, s_free, s_r)
        assert r_free <= 2e-2, (step, r_free)
        if flips is not None:
            assert flips <= max(2, 2e-5 * total), (step, flips, total)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("fixture,name,depth", [("i2v_resnet50_d2_32", "resnet", 2), ("i2v_vgg_d3_32", "vgg", 3)])
def test_i2v_matches_reference_fixture(golden, engine, fixture, name, depth):
    g = golden(fixture)
    steps = int(g["steps"])
    atk = image_attacks.ImageGuidedFMDirection_Adam([name], depth=depth, step_size=float(g["step_size"]),
                                                    steps=steps, engine=engine)
    videos = torch.from_numpy(g["videos"])
    adv = atk(videos, torch.zeros(1, dtype=torch.long), ["clip0"])
    assert tuple(adv.shape) == tuple(videos.shape) and adv.is_cuda
    assert not adv.is_contiguous()        # the reference returns the permuted view (image_attacks.py:362-363)
    cost = np.array([float(atk.loss_info["clip0"][i]["cost"]) for i in range(steps)], np.float32)
    assert np.allclose(cost, g["cost"], rtol=1e-5)
    adv = adv.cpu().numpy()
    _bounds_ok(g["videos"], adv)
    # free-running end state vs the reference's: reported, and bounded by what 3 Adam steps can move
    d = np.abs(adv - g["adv"])
    _record("%s/%s/final" % (fixture, engine), frac_equal=(d == 0).mean(), frac_1e4=(d <= 1e-4).mean(),
            frac_1_255=(d <= (1 / 255) / 0.225).mean(), max_abs=d.max())
    assert d.max() <= 2 * steps * float(g["step_size"]) / 0.224 * 1.01


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("fixture,name,depth", [("dr_resnet50_d2_32", "resnet", 2), ("dr_vgg_d2_32", "vgg", 2)])
def test_dispersion_matches_reference_fixture(golden, engine, fixture, name, depth):
    """ImageGuidedStd_Adam (Dispersion Reduction, image_attacks.py:129-234) against the unmodified class; also
    chunked over frames (two-pass statistics), which must give the same costs."""
    g = golden(fixture)
    steps = int(g["steps"])
    videos = torch.from_numpy(g["videos"])
    atk = image_attacks.ImageGuidedStd_Adam([name], depth=depth, step_size=float(g["step_size"]), steps=steps, engine=engine)
    adv = atk(videos, torch.zeros(1, dtype=torch.long), ["clip0"])
    assert tuple(adv.shape) == tuple(videos.shape) and adv.is_cuda and not adv.is_contiguous()
    cost = np.array([float(atk.loss_info["clip0"][i]["cost"]) for i in range(steps)], np.float32)
    # std() is not scale invariant (the cosine is): the tensor cores accumulate in FP32 with TRUNCATION, a systematic
    # relative bias of ~(K/16) ulp per convolution (csrc/conv_tc.cu) that compounds to ~1.3e-5 over the 24 layers below
    # layer2 — measured 1.26e-5 on this fixture; the cuDNN engine (round-to-nearest FMA chains) stays within 1e-5
    cost_tol = 3e-5 if engine.startswith("native") else 1e-5
    _record("%s/%s/cost" % (fixture, engine), rel_err=float(np.abs(cost / g["cost"] - 1).max()))
    assert np.allclose(cost, g["cost"], rtol=cost_tol), (cost, g["cost"])
    adv = adv.cpu().numpy()
    _bounds_ok(g["videos"], adv)
    d = np.abs(adv - g["adv"])
    _record("%s/%s/final" % (fixture, engine), frac_equal=(d == 0).mean(), frac_1e4=(d <= 1e-4).mean(),
            frac_1_255=(d <= (1 / 255) / 0.225).mean(), max_abs=d.max())
    assert d.max() <= 2 * steps * float(g["step_size"]) / 0.224 * 1.01
    # step-1 gradient against the reference's own tap (dcost/dmodifier) and the chunked two-pass variant
    taps, taps1 = {}, {}
    res = attack_loop.run_dispersion([atk._engine], videos, EPS, 1, float(g["step_size"]), tap=lambda i, t: taps.setdefault(i, t))
    res1 = attack_loop.run_dispersion([atk._engine], videos, EPS, 1, float(g["step_size"]), chunk=1,
                                      tap=lambda i, t: taps1.setdefault(i, t))
    std = O.STD[None, :, None, None]
    s, r = _grad_scores(taps[0]["g"].cpu().numpy() / std, g["g_mod_first"].astype(np.float64))
    s1, r1 = _grad_scores(taps1[0]["g"].cpu().numpy() / std, g["g_mod_first"].astype(np.float64))
    _record("%s/%s/step1_vs_reference" % (fixture, engine), sign=s, relL2=r, sign_chunked=s1, relL2_chunked=r1)
    assert np.allclose(res.cost, g["cost"][:1], rtol=cost_tol) and np.allclose(res1.cost, g["cost"][:1], rtol=cost_tol)
    # the std gradient is well conditioned (no cancellation as in I2V): what is left is float32-vs-float32 noise of the
    # backbone (torch's own f32 gradient is 2e-5 .. 4e-3 from float64 on these nets, profiles/r01_parity_stats.json)
    assert s >= 0.999 and r <= 5e-3, (s, r)
    assert s1 >= 0.999 and r1 <= 5e-3, (s1, r1)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name,depth", [("resnet", 2), ("vgg", 3), ("squeezenet", 2), ("alexnet", 3), ("densenet", 2)])
def test_i2v_gradients_vs_float64_arbiter(engine, name, depth):
    videos, _ = synth.clip(2, b=1, f=2, h=64, w=64)
    model = backbones.get_model(name)
    eng = engines.make_engine(model, name, depth, engine)
    taps = {}
    attack_loop.run_image_guided([eng], videos, EPS, 3, 0.005, tap=lambda i, d: taps.setdefault(i, d))
    _teacher_forced_steps("i2v_%s_d%d/%s" % (name, depth, engine), [name], depth, taps, videos)


@pytest.mark.parametrize("engine", ENGINES)
def test_i2v_step1_vs_reference_tap(golden, engine):
    """Step 1 against the reference's own tap (dcost/dmodifier recorded from its optimizer)."""
    g = golden("i2v_resnet50_d2_32")
    model = backbones.get_model("resnet")
    eng = engines.make_engine(model, "resnet", 2, engine)
    taps = {}
    res = attack_loop.run_image_guided([eng], torch.from_numpy(g["videos"]), EPS, 1, float(g["step_size"]),
                                       tap=lambda i, d: taps.setdefault(i, d))
    std = O.STD[None, :, None, None]
    g_mod = taps[0]["g"].cpu().numpy() / std
    ref = g["g_mod_first"]
    s, r = _grad_scores(g_mod, ref.astype(np.float64))
    _record("i2v_resnet50_d2_32/%s/step1_vs_reference" % engine, sign=s, relL2=r)
    assert np.allclose(res.cost, g["cost"][:1], rtol=1e-5)
    # float32-vs-float32 floor measured by the survey: 99.7 % (same code, threads differ).  Native engine: 99.77 %
    # measured; cuDNN's algorithm choice on these 4x4 feature maps is noisier (99.33 % measured, r2 run a)
    assert s >= (0.995 if engine.startswith("native") else 0.99), s


@pytest.mark.parametrize("engine", ENGINES)
def test_ens_matches_reference_fixture(golden, engine):
    g = golden("ens_4models_64")
    names = ["resnet", "vgg", "squeezenet", "alexnet"]
    depths = {"resnet": 2, "vgg": 3, "squeezenet": 2, "alexnet": 3}
    atk = image_attacks.ImageGuidedFML2_Adam_MultiModels(names, depths, steps=int(g["steps"]), engine=engine)
    assert atk.step_size == 0.005
    adv = atk(torch.from_numpy(g["videos"]), torch.zeros(1, dtype=torch.long), ["clip0"]).cpu().numpy()
    cost = np.array([float(atk.loss_info["clip0"][i]["cost"]) for i in range(int(g["steps"]))], np.float32)
    assert np.allclose(cost, g["cost"], rtol=1e-5)
    _bounds_ok(g["videos"], adv)
    d = np.abs(adv - g["adv"])
    _record("ens_4models_64/%s/final" % engine, frac_equal=(d == 0).mean(), frac_1e4=(d <= 1e-4).mean(),
            frac_1_255=(d <= (1 / 255) / 0.225).mean(), max_abs=d.max())
    # teacher-forced gradient of the whole ensemble, every step
    taps = {}
    engs = [engines.make_engine(backbones.get_model(n), n, depths[n], engine) for n in names]
    attack_loop.run_image_guided(engs, torch.from_numpy(g["videos"]), EPS, 2, 0.005, tap=lambda i, t: taps.setdefault(i, t))
    _teacher_forced_steps("ens/%s" % engine, names, depths, taps, g["videos"])


@pytest.mark.parametrize("engine", ENGINES)
def test_aens_matches_reference_fixture(golden, engine):
    g = golden("aens_4models_64")
    names = ["resnet", "vgg", "squeezenet", "alexnet"]
    atk = TPAMI_attack.AENS_I2V_MF(names, {n: [2, 3] for n in names}, 0.005, momentum=0.5, steps=int(g["steps"]),
                                   engine=engine)
    adv, used_time, cost_saved = atk(torch.from_numpy(g["videos"]), torch.zeros(1, dtype=torch.long), ["clip0"])
    assert isinstance(used_time, float) and used_time > 0
    assert cost_saved.shape == (int(g["steps"]),) and cost_saved.dtype == np.float64
    assert np.allclose(cost_saved, g["cost_saved"], rtol=1e-5)
    assert np.allclose(np.stack(atk.weights), g["weights"], rtol=1e-5)
    assert np.allclose(atk.coeffs.cpu().numpy(), g["coeffs_after"], rtol=1e-5)      # persists
    adv = adv.cpu().numpy()
    _bounds_ok(g["videos"], adv)
    d = np.abs(adv - g["adv"])
    _record("aens_4models_64/%s/final" % engine, frac_equal=(d == 0).mean(), frac_1e4=(d <= 1e-4).mean(),
            frac_1_255=(d <= (1 / 255) / 0.225).mean(), max_abs=d.max())
    # second call: coeffs carry over when momentum != 0 (TPAMI_attack.py:165, 265)
    first = atk.coeffs.clone()
    atk(torch.from_numpy(g["videos"]), torch.zeros(1, dtype=torch.long), ["clip0"])
    c0, _ = O.layer_reweight(first.cpu().numpy(), np.ones(8, np.float32), 0.5)
    assert np.allclose(atk.weights[0], c0, rtol=1e-6)


@pytest.mark.parametrize("engine", ENGINES)
def test_aens_coef_ce_and_validation(golden, engine):
    g = golden("aens_ce_2models_64")
    atk = TPAMI_attack.AENS_I2V_MF(["resnet", "squeezenet"], {"resnet": [1, 2], "squeezenet": [2, 3]}, 0.005,
                                   coef_CE=True, steps=int(g["steps"]), engine=engine)
    adv, _, cost_saved = atk(torch.from_numpy(g["videos"]), torch.zeros(1, dtype=torch.long), ["clip0"])
    assert np.allclose(cost_saved, g["cost_saved"], rtol=1e-5)
    assert np.allclose(np.stack(atk.weights), g["weights"], rtol=1e-5)
    _bounds_ok(g["videos"], adv.cpu().numpy())
    with pytest.raises(ValueError):   # reference silently needs exactly two depths per model (D10)
        TPAMI_attack.AENS_I2V_MF(["resnet"], {"resnet": [1, 2, 3]}, 0.005, engine=engine)


@pytest.mark.parametrize("engine", ENGINES)
def test_chunking_does_not_change_the_result(engine):
    """Frames are independent units: any chunking of N gives the same costs; on the native engine (deterministic
    per-frame arithmetic: fixed k-order per output element, per-frame K1 clusters, per-row pooling) the adversarial clip
    and the cost log are BIT-identical whatever the chunk size."""
    videos, _ = synth.clip(1, b=1, f=6, h=64, w=64)
    out = []
    for chunk in (6, 4, 1):
        model = backbones.get_model("resnet")
        eng = engines.make_engine(model, "resnet", 2, engine)
        res = attack_loop.run_image_guided([eng], videos, EPS, 3, 0.005, chunk=chunk)
        out.append((res.adv.cpu().numpy(), res.cost))
    for adv, cost in out[1:]:
        if engine.startswith("native"):
            assert np.array_equal(cost, out[0][1]) and np.array_equal(adv, out[0][0])
        else:
            # cuDNN may pick different algorithms for different batch sizes, so only the cost is compared
            # (the trajectories are chaotic in the conv rounding noise)
            assert np.allclose(cost, out[0][1], rtol=1e-5)


@pytest.mark.parametrize("adaptive", [False, True])
def test_cuda_graph_replay_is_bit_identical_to_the_eager_loop(adaptive, monkeypatch):
    """Steps 1.. of a native run are replays of a CUDA graph captured at step 1 (attack_loop.ImageGuidedRun.step): the
    adversarial clip, the cost log and (AENS) the coefficient log must be bit-identical to the eager loop, for several
    chunks per step too, and the launch accounting must count the replayed launches."""
    videos, _ = synth.clip(3, b=1, f=6, h=64, w=64)
    names = ["resnet", "squeezenet"] if adaptive else ["resnet"]
    engs = [engines.make_engine(backbones.get_model(n), n, [1, 2] if adaptive else 2, "native") for n in names]
    out = []
    for graph in ("0", "1"):
        monkeypatch.setenv("I2V_GRAPH", graph)
        coeffs = torch.ones(4, device="cuda") if adaptive else None
        capi.LAUNCHES.clear()
        run = attack_loop.ImageGuidedRun(engs, EPS, 7, 0.005, adaptive=adaptive, coeffs=coeffs, momentum=0.5, chunk=4)
        run.setup(videos)
        for _ in range(7):
            run.step()
        assert (run._graph is not None) == (graph == "1")
        res = run.finish()
        out.append((res.adv.clone(), res.cost, res.weights, dict(capi.LAUNCHES)))
    assert torch.equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    if adaptive:
        assert np.array_equal(out[0][2], out[1][2])
    assert out[0][3] == out[1][3] and sum(out[1][3].values()) > 100


def test_repeated_calls_reuse_the_run_and_its_graph():
    """A sweep calls the attack object clip after clip (image_main.py:82-89): the second call of the same shape reuses the
    first call's device state and captured graph and must give exactly what a fresh attack object gives; a call with another
    shape in between starts over; the result of an earlier call is not overwritten by a later one."""
    v1, _ = synth.clip(11, b=1, f=4, h=64, w=64)
    v2, _ = synth.clip(12, b=1, f=4, h=64, w=64)
    v3, _ = synth.clip(13, b=1, f=2, h=32, w=32)
    lab = torch.zeros(1, dtype=torch.long)

    def fresh(v, name):
        a = image_attacks.ImageGuidedFMDirection_Adam(["resnet"], depth=2, step_size=0.005, steps=5, engine="native")
        out = a(v, lab, [name])
        return out.clone(), a.loss_info[name]
    want = [fresh(v, "c%d" % i) for i, v in enumerate((v1, v2, v3))]
    atk = image_attacks.ImageGuidedFMDirection_Adam(["resnet"], depth=2, step_size=0.005, steps=5, engine="native")
    got1 = atk(v1, lab, ["c0"])
    run = atk._run_cache["run"]
    assert run._graph is not None
    got2 = atk(v2, lab, ["c1"])
    assert atk._run_cache["run"] is run and run._graph is not None
    assert torch.equal(got1, want[0][0]) and torch.equal(got2, want[1][0])          # got1 survived the second call
    got3 = atk(v3, lab, ["c2"])
    assert torch.equal(got3, want[2][0])
    got1b = atk(v1, lab, ["c0b"])
    assert torch.equal(got1b, want[0][0])
    for i in range(3):
        assert atk.loss_info["c%d" % i] == want[i][1]
    # AENS: the persistent coefficient vector carries over between calls exactly as without reuse (TPAMI_attack.py:165, 265)
    names = ["resnet", "squeezenet"]
    a1 = TPAMI_attack.AENS_I2V_MF(names, {n: [1, 2] for n in names}, 0.005, momentum=0.5, steps=4, engine="native")
    r1 = [a1(v, lab, ["x"])[0].clone() for v in (v1, v2)]
    w1 = np.stack(a1.weights)
    import os
    os.environ["I2V_GRAPH"] = "0"
    try:
        a2 = TPAMI_attack.AENS_I2V_MF(names, {n: [1, 2] for n in names}, 0.005, momentum=0.5, steps=4, engine="native")
        r2 = []
        for v in (v1, v2):
            a2.__dict__["_run_cache"] = {}               # no reuse, no graph: the plain loop
            r2.append(a2(v, lab, ["x"])[0].clone())
        w2 = np.stack(a2.weights)
    finally:
        del os.environ["I2V_GRAPH"]
    assert torch.equal(r1[0], r2[0]) and torch.equal(r1[1], r2[1]) and np.array_equal(w1, w2)


def test_native_chunking_bit_identical_at_benchmark_size():
    """bench.py's configuration (BASELINE.json configs[1]) relies on this: 224 x 224 frames on the native engine in
    256-frame chunks plus a remainder chunk with a different (n, h, w) buffer plan.  288 frames (9 clips x 32), two
    steps: chunk 256 (256 + 32), 100 (100 + 100 + 88), 7 (41 x 7 + 1) and the first clip alone must agree BIT for bit in
    the adversarial clip and the per-step cost (the cost of a sub-batch is compared through its own frames' cosines)."""
    clips = 9
    videos = torch.cat([synth.clip(40 + i, b=1, f=32, h=224, w=224)[0] for i in range(clips)], 0)
    eng = engines.make_engine(backbones.get_model("resnet"), "resnet", 2, "native")
    ref_adv = ref_cost = None
    for chunk in (256, 100, 7):
        run = attack_loop.ImageGuidedRun([eng], EPS, 2, 0.005, chunk=chunk)
        run.setup(videos)
        assert run.chunk == chunk and len(run.spans) == -(-clips * 32 // chunk)
        for _ in range(2):
            run.step()
        cos = run.cos.clone()
        res = run.finish()
        adv = res.adv.contiguous()
        if ref_adv is None:
            ref_adv, ref_cost, ref_cos = adv, res.cost, cos
            _bounds_ok(videos[:1].numpy(), adv[:1].cpu().numpy())
            assert float((adv != videos.cuda()).float().mean()) > 0.5
        else:
            assert np.array_equal(res.cost, ref_cost), (chunk, res.cost, ref_cost)
            assert torch.equal(cos, ref_cos), chunk
            assert torch.equal(adv, ref_adv), chunk
        del run, res
    one = attack_loop.run_image_guided([eng], videos[:1], EPS, 2, 0.005, chunk=32)
    assert torch.equal(one.adv.contiguous(), ref_adv[:1])


def test_i2v_gradients_vs_float64_arbiter_at_224():
    """The teacher-forced float64 arbiter at the benchmark's own layer shapes: two 224 x 224 frames through ResNet-50 up to
    layer2 (M = 2*112*112 stem rows, 56 x 56 and 28 x 28 bottlenecks), native engine, three steps, with the engine's ReLU
    and max-pool decisions forced on the arbiter — same bounds as the small-shape cases."""
    videos, _ = synth.clip(2, b=1, f=2, h=224, w=224)
    eng = engines.make_engine(backbones.get_model("resnet"), "resnet", 2, "native")
    taps = {}
    attack_loop.run_image_guided([eng], videos, EPS, 3, 0.005, tap=lambda i, d: taps.setdefault(i, d))
    _teacher_forced_steps("i2v_resnet_d2_224/native", ["resnet"], 2, taps, videos)


def test_config1_60_steps_vs_reference_fixture(golden):
    """BASELINE.json configs[0]/[1] at the reference's own size and step budget (run_image_guided.py:63-70: 60 steps of
    0.005, one 32-frame 224 x 224 clip; image_attacks.py:294-364), free running, native engine, against the UNMODIFIED
    reference class (tests/golden/i2v_resnet50_d2_224_60step.npz, oracle/make_golden.py) and against the float64 arbiter
    of the first step stored with it.

    What the fixture itself says about the bars (profiles/r02_parity_60step.json keeps all figures side by side):
      * step-1 gradient signs: the reference's own float32 gradient agrees with float64 on only 97.4 % of the significant
        elements (|g| > 1e-3 max) at this size — F.cosine_similarity's float32 sums put a ~10 % relative-L2 error on a
        quantity that is 1e-8 of its terms.  An implementation that is closer to float64 than that cannot agree with the
        reference on more than ~97.4 %; the north star's 99.99 % is asserted against the ARBITER (>= 0.995 here, 0.999
        measured), the agreement with the reference is recorded and bounded by the reference's own accuracy.
      * two runs of the reference that differ only in the CPU convolution backend (oneDNN vs torch's im2col + GEMM, which
        share the cosine arithmetic) end with 84.8 % of the final perturbation within 1/255 and costs 2.7e-5 apart
        (profiles/r02_reference_self_agreement.json): the trajectory is chaotic in the rounding noise.
    Asserted: first cost within 1e-5 (same input), every cost within 5e-4 of the reference's free-running trajectory
    (teacher-forced cosines are within 1e-8: test_i2v_gradients_vs_float64_arbiter_at_224), final perturbation within 1/255
    on at least the reference-vs-reference fraction minus 3 points, eps-ball and [0,1] exactly, and the run without the
    per-step tap bit-identical."""
    g = golden("i2v_resnet50_d2_224_60step")
    steps, f, side = int(g["steps"]), int(g["frames"]), int(g["side"])
    videos, _ = synth.clip(0, b=1, f=f, h=side, w=side)
    taps = {}

    def tap(i, d):
        if i == 0:
            taps[0] = d["g"].cpu().numpy()
    eng = engines.make_engine(backbones.get_model("resnet"), "resnet", 2, "native")
    res = attack_loop.run_image_guided([eng], videos, EPS, steps, float(g["step_size"]), tap=tap)
    adv = res.adv.cpu().numpy()
    _bounds_ok(videos.numpy(), adv)
    from tools import reference_on_gpu as RG
    g_mod = taps[0] / O.STD[None, :, None, None]                          # dcost/dmodifier = dcost/dtrue_image / std
    sc = RG.score_against_fixture(adv - videos.numpy(), res.cost, g_mod, g)
    # the same run without the per-step host sync of the tap must give the same clip (device-side logs only)
    res2 = attack_loop.run_image_guided([eng], videos, EPS, steps, float(g["step_size"]))
    assert torch.equal(res2.adv, res.adv) and np.array_equal(res2.cost, res.cost)
    _record("config1_60step/native", final_cost=float(res.cost[-1]), ref_final_cost=float(g["cost"][-1]), **sc)
    assert sc["cost_rel_err_by_step"][0] <= 1e-5, sc["cost_rel_err_by_step"][:3]
    assert sc["cost_rel_err_max"] <= 5e-4, sc["cost_rel_err_max"]
    assert sc["step1_sign_vs_float64_big"] >= 0.995, sc
    assert sc["step1_sign_vs_reference_big"] >= sc["reference_cpu_step1_sign_vs_float64_big"] - 0.01, sc
    assert sc["final_frac_within_1_255"] >= 0.848 - 0.03, sc["final_frac_within_1_255"]


def test_base_attacks_match_reference_fixture(golden):
    g = golden("base_tiny3d")
    model = synth.TinyVideoNet().cuda()
    videos = torch.from_numpy(g["videos"]).cuda()
    labels = torch.from_numpy(g["labels"]).cuda()
    step = EPS / 3

    def frac_equal(a, ref, tol):
        return (np.abs(a.cpu().numpy() - ref) <= tol).mean()

    # The model gradient comes from cuDNN here and from oneDNN in the fixture: signs of near-zero
    # gradients may flip, moving a pixel by 2*step/std.  Everything else is bit-exact arithmetic.
    fg = base_attacks.FGSM(model)(videos.clone(), labels)
    assert frac_equal(fg, g["fgsm"], 1e-6) >= 0.999
    bim = base_attacks.BIM(model, steps=3)
    assert bim.step_size == EPS / 3 and bim.attack == "FGSM"
    assert frac_equal(bim(videos.clone(), labels), g["bim3"], 1e-6) >= 0.98
    mi = base_attacks.MIFGSM(model, steps=3)(videos.clone(), labels)
    assert frac_equal(mi, g["mifgsm3"], 1e-6) >= 0.98
    tgt = base_attacks.BIM(model, steps=2)
    tgt.set_attack_mode("targeted", lambda images, labels: (labels + 1) % 10)
    assert tgt._targeted == -1
    assert frac_equal(tgt(videos.clone(), labels), g["bim2_targeted"], 1e-6) >= 0.98
    # uint8 return type and mode restoration (base_attacks.py:226-234)
    model.train()
    b2 = base_attacks.BIM(model, steps=1)
    b2.set_return_type("int")
    out = b2(videos.clone(), labels)
    assert out.dtype == torch.uint8 and model.training
    with pytest.raises(ValueError):
        b2.set_return_type("double")
    with pytest.raises(ValueError):
        b2.set_attack_mode("nonsense")


def test_mifgsm_runs_at_16_frames():
    """The reference asserts T == 32 in norm_grads (utils.py:61) and cannot run MI on UCF-shaped
    16-frame clips; the assert-free path must agree with the oracle loop."""
    model = synth.TinyVideoNet().cuda()
    videos, _ = synth.clip(5, b=2, f=16, h=12, w=12)
    labels = torch.tensor([1, 2])
    adv = base_attacks.MIFGSM(model, steps=2)(videos.cuda(), labels.cuda()).cpu().numpy()
    want = OL.mifgsm(synth.TinyVideoNet(), videos.numpy(), labels, steps=2)
    assert (np.abs(adv - want) <= 1e-6).mean() >= 0.98
    import utils
    gr = torch.randn(2, 3, 16, 6, 6, device="cuda")
    ref = gr / gr.abs().mean(dim=(1, 3, 4), keepdim=True)
    assert torch.allclose(utils.norm_grads(gr), ref, rtol=1e-6)


def test_transform_video_helpers():
    a = image_attacks.Attack("x")
    v = torch.rand(4, 3, 8, 8, device="cuda")
    want = O.normalize(v.cpu().numpy(), 64)
    got = a._transform_video(v.clone(), "forward")
    assert np.array_equal(got.cpu().numpy(), want)
    back = a._transform_video(got.clone(), "back")
    assert np.array_equal(back.cpu().numpy(), O.denorm(want, 64))


def test_image_main_driver_end_to_end(tmp_path):
    """`image_main.py` as run_image_guided.py launches it: artefact names, shapes and dtypes of image_main.py:90-95."""
    import json
    import image_main
    image_main.main(["--attack_method", "ImageGuidedFMDirection_Adam", "--depth", "2", "--step", "2", "--step_size", "0.005",
                     "--synthetic", "--num_clips", "3", "--frames", "2", "--side", "32", "--opt_path", str(tmp_path),
                     "--batch_nums", "1", "--batch_index", "1", "--weights", "random"])
    out = os.path.join(str(tmp_path), "Image-ImageGuidedFMDirection_Adam-2-")
    for idx in range(3):
        adv = np.load(os.path.join(out, "%d-adv.npy" % idx))
        assert adv.shape == (3, 2, 32, 32) and adv.dtype == np.float32 and np.isfinite(adv).all()
        clean, _ = synth.clip(idx, b=1, f=2, h=32, w=32)
        _bounds_ok(clean.numpy(), adv[None])
        assert np.abs(adv - clean.numpy()[0]).max() > 0
    info = json.load(open(os.path.join(out, "loss_info_1.json")))
    assert sorted(info) == ["synthetic_%05d" % i for i in range(3)] and sorted(info["synthetic_00000"]) == ["0", "1"]


def test_base_variants_match_reference_fixture(golden):
    """DIFGSM / TIFGSM / SGM / SIM / TIFGSM3D (base_attacks.py:342-683) on the GPU against the unmodified classes: the
    K3b update block is bit-exact given the gradient, the gradient itself comes from cuDNN / the K7 stencil instead of
    oneDNN, so the bar is the fraction of identical pixels after the sign steps and the exact eps-ball."""
    import random
    g = golden("base_variants")
    labels = torch.from_numpy(g["labels"])
    v = torch.from_numpy(g["videos"])

    def frac_equal(adv, ref):
        adv = adv.cpu().numpy()
        _bounds_ok(g["videos"], adv)
        return float((np.abs(adv - ref) < 1e-6).mean())

    stats = {}
    for mom in (False, True):
        tag = "_mom" if mom else ""
        model = synth.TinyVideoNet().cuda()
        stats["tifgsm" + tag] = frac_equal(base_attacks.TIFGSM(model, steps=3, momentum=mom)(v.clone(), labels), g["tifgsm3" + tag])
        stats["sim" + tag] = frac_equal(base_attacks.SIM(model, steps=2, momentum=mom)(v.clone(), labels), g["sim2" + tag])
        stats["tifgsm3d" + tag] = frac_equal(base_attacks.TIFGSM3D(model, steps=2, momentum=mom)(v.clone(), labels), g["tifgsm3d2" + tag])
        stats["sgm" + tag] = frac_equal(base_attacks.SGM(synth.TinyReluVideoNet().cuda(), steps=3, momentum=mom)(v.clone(), labels),
                                        g["sgm3" + tag])
        di_videos, _ = synth.clip(5, b=1, f=2, h=224, w=224)
        random.seed(11)
        torch.manual_seed(11)
        adv = base_attacks.DIFGSM(model, steps=4, momentum=mom)(di_videos.clone(), labels).cpu().numpy()
        _bounds_ok(di_videos.numpy(), adv)
        stats["difgsm" + tag] = float((np.abs(adv - di_videos.numpy() - g["difgsm4_delta16" + tag].astype(np.float32)) < 2e-3).mean())
    _record("base_variants/frac_equal", **stats)
    # TI variants smooth the gradient (stable signs); the others take the raw sign of a cuDNN-vs-oneDNN gradient whose
    # near-zero entries flip, each flip moving a pixel by 2*step/std per step (same floor as BIM / MI above: 0.98)
    # measured with TF32 off for the white-box model: 1.0 everywhere except DI (0.988 / 0.994), whose nearest-neighbour
    # resize turns one flipped gradient sign into a different sampled pixel
    for k, val in stats.items():
        assert val > (0.97 if k.startswith("difgsm") else 0.995), (k, val, stats)


def test_depthwise_stencil_vs_torch():
    """K7 against F.conv2d / F.conv3d (groups = 3, zero padding) in float64, ragged sizes included."""
    gen = torch.Generator().manual_seed(2)
    for (B, T, H, W, kt, kh, kw) in [(2, 3, 12, 12, 1, 15, 15), (1, 16, 9, 13, 15, 15, 15), (1, 2, 5, 7, 1, 3, 5), (1, 4, 6, 6, 3, 1, 1)]:
        x = torch.randn(B, 3, T, H, W, generator=gen)
        k = torch.rand(kt, kh, kw, generator=gen)
        out = torch.full_like(x, float("nan")).cuda()
        capi.depthwise_stencil(x.cuda(), out, k.cuda())
        ref = torch.nn.functional.conv3d(x.double(), k.double().expand(3, 1, kt, kh, kw).contiguous(), groups=3,
                                         padding=(kt // 2, kh // 2, kw // 2))
        err = (out.cpu().double() - ref).abs().max() / ref.abs().max()
        assert err < 2e-6, (B, T, H, W, kt, kh, kw, float(err))


def test_video_variants_match_reference_fixture(golden):
    """TemporalTranslation (video_attacks.py), TAP (base_attacks.py:685-814) and ILAF (image_attacks.py:498-629) on the
    GPU against the unmodified classes run on the CPU (tests/golden/video_variants.npz).  The K8 / K3b / K3c / K3d / K7 /
    K9 arithmetic is pinned bit-exactly or against float64 in test_gpu_kernels.py; here the white-box model's gradient
    comes from cuDNN instead of oneDNN, so the bar is the fraction of identical pixels after the sign steps (each flip of
    a near-zero gradient moves a pixel by 2*step/std), the exact eps-ball and the per-step costs."""
    import image_attacks
    import video_attacks
    g = golden("video_variants")
    v = torch.from_numpy(g["videos"])
    labels = torch.from_numpy(g["labels"])

    def frac_equal(adv, ref, base=None):
        adv = adv.detach().cpu().numpy()
        _bounds_ok(g["videos"] if base is None else base, adv)
        return float((np.abs(adv - ref) < 1e-6).mean())

    stats = {}

    def tt(kernlen, weight, mode, steps, mom):
        atk = video_attacks.TemporalTranslation(
            synth.TinyTPNLike().cuda(), {"kernlen": kernlen, "momentum": mom, "weight": weight, "move_type": "adj",
                                         "kernel_mode": mode}, steps=steps)
        return atk(v.clone(), labels)
    adv_tt = tt(5, 0.5, "gaussian", 3, False)
    stats["tt3_k5"] = frac_equal(adv_tt, g["tt3_k5"])
    stats["tt3_k5_mom"] = frac_equal(tt(5, 0.5, "gaussian", 3, True), g["tt3_k5_mom"])
    stats["tt2_k9_linear"] = frac_equal(tt(9, 0.3, "linear", 2, False), g["tt2_k9_linear"])

    def tt_move(move_type):
        import random
        random.seed(21)                                  # oracle/make_golden.py: tt_move (video_attacks.py:107-135)
        atk = video_attacks.TemporalTranslation(
            synth.TinyTPNLike().cuda(), {"kernlen": 5, "momentum": True, "weight": 0.7, "move_type": move_type,
                                         "kernel_mode": "random"}, steps=2)
        return atk(v.clone(), labels)
    stats["tt2_k5_large"] = frac_equal(tt_move("large"), g["tt2_k5_large"])
    stats["tt2_k5_randommove"] = frac_equal(tt_move("random"), g["tt2_k5_randommove"])
    # kernlen 7 makes the reference's 5-way split produce an empty model batch (torch.cat([]) raises there): runs here
    adv7 = tt(7, 0.5, "gaussian", 1, True)
    _bounds_ok(g["videos"], adv7.cpu().numpy())
    for conv3d, tag in ((True, "3d"), (False, "2d")):
        atk = base_attacks.TAP(synth.TinyTPNLike().cuda(), {"kernlen": 3, "temporal_kernlen": 3, "eta": 1e3, "conv3d": conv3d,
                                                            "model_type": "tpn"}, steps=3)
        stats["tap3_" + tag] = frac_equal(atk(v.clone(), labels), g["tap3_" + tag])
        last = atk.loss_info[2]
        got = np.array([float(last["ce loss"]), float(last["reg_cost"]), float(np.asarray(last["distance"]).reshape(-1)[0])])
        assert np.allclose(got, g["tap3_%s_last_losses" % tag], rtol=2e-2), (got, g["tap3_%s_last_losses" % tag])
    il = image_attacks.ILAF(synth.TinyTPNLike().cuda(), "tpn", step_size=0.005, steps=4)
    out = il(torch.from_numpy(g["tt3_k5"]).clone(), v.clone(), labels, ["v0"])
    assert tuple(out.shape) == tuple(g["ilaf4"].shape)
    stats["ilaf4"] = float((np.abs(out.detach().cpu().numpy() - g["ilaf4"]) < 1e-6).mean())
    _bounds_ok(g["videos"], il.last_adv.cpu().numpy())
    costs = np.array([float(il.loss_info["v0"][i]["cost"]) for i in range(4)])
    assert np.allclose(costs, g["ilaf4_costs"], rtol=2e-4), (costs, g["ilaf4_costs"])
    _record("video_variants/frac_equal", **stats)
    for k, val in stats.items():                  # measured: 1.0 except tap3_2d 0.99986
        assert val > 0.995, (k, val, stats)


def test_temporal_translation_helpers():
    """The reference's private helpers kept on the class (video_attacks.py:80-177) against their torch statements."""
    import random
    import video_attacks
    atk = video_attacks.TemporalTranslation(synth.TinyTPNLike().cuda(), {"kernlen": 5, "momentum": False, "weight": 0.3,
                                                                         "move_type": "adj", "kernel_mode": "linear"})
    g = torch.Generator().manual_seed(3)
    v = torch.randn(1, 3, 32, 6, 6, generator=g).cuda()
    assert torch.equal(atk._cycle_move(v, -2), torch.roll(v, -2, dims=2))
    assert torch.equal(atk._cycle_move_large(v, 2), torch.roll(v, 17, dims=2))
    assert torch.equal(atk._cycle_move_large(v, -3), torch.roll(v, -18, dims=2))
    random.seed(5)
    got = atk._cycle_move_random(v, 1)
    random.seed(5)
    assert torch.equal(got, torch.roll(v, random.randint(0, 100) % 32, dims=2))
    ex = atk._exchange_move(v, [(0, 5), (7, 9)])
    assert torch.equal(ex[:, :, 0], v[:, :, 5]) and torch.equal(ex[:, :, 9], v[:, :, 7]) and torch.equal(ex[:, :, 3], v[:, :, 3])
    grads = torch.randn(5, 1, 3, 32, 6, 6, generator=g).cuda()
    k = atk.kernel.double()
    s_conv = torch.matmul(k, grads.double().reshape(5, -1)).reshape(1, 3, 32, 6, 6)
    assert torch.allclose(atk._conv1d_frame(grads).double(), s_conv, rtol=1e-6, atol=1e-7)
    diff = torch.stack([torch.roll(grads[i], -m, dims=2) for i, m in enumerate(atk.cycle_move_list)])      # 172-173
    d_conv = torch.matmul(k, diff.double().reshape(5, -1)).reshape(1, 3, 32, 6, 6)
    assert torch.allclose(atk._grad_augmentation(grads).double(), 0.7 * s_conv + 0.3 * d_conv, rtol=1e-6, atol=1e-7)


def test_attack_driver_end_to_end(tmp_path):
    """`attack.py` (reference attack.py:1-96) with the stand-in model: both dispatch branches, artefact names / shapes /
    dtypes of 92-96, the eps-ball."""
    import attack
    common = ["--synthetic", "--model", "tiny", "--num_clips", "3", "--batch_size", "2", "--frames", "8", "--side", "16",
              "--num_classes", "10", "--opt_path", str(tmp_path), "--step", "2"]
    attack.main(common + ["--attack_type", "image", "--attack_method", "MIFGSM"])
    attack.main(common + ["--attack_type", "video", "--attack_method", "TemporalTranslation", "--kernlen", "5",
                          "--augmentation_weight", "0.5", "--iterative_momentum"])
    for sub in ("tiny-MIFGSM-2-", "tiny-TemporalTranslation-2-"):
        for idx in range(3):
            adv = np.load(os.path.join(str(tmp_path), sub, "%d-adv.npy" % idx))
            ori = np.load(os.path.join(str(tmp_path), sub, "%d-ori.npy" % idx))
            assert adv.shape == ori.shape == (3, 8, 16, 16) and adv.dtype == np.float32 and np.isfinite(adv).all()
            clean, _ = synth.clip(idx, b=1, f=8, h=16, w=16)
            assert np.array_equal(ori, clean.numpy()[0])
            _bounds_ok(clean.numpy(), adv[None])
            assert np.abs(adv - ori).max() > 0
    with pytest.raises(ValueError):
        attack.main(common + ["--attack_type", "video", "--attack_method", "BIM"])


def test_fine_tune_driver_end_to_end(tmp_path):
    """attack.py -> image_fine_tune_attack.py: ILAF over the saved `{label}-adv.npy` / `{label}-ori.npy` pairs."""
    import attack
    import image_fine_tune_attack as ft
    src, dst = os.path.join(str(tmp_path), "src"), os.path.join(str(tmp_path), "dst")
    attack.main(["--synthetic", "--model", "tiny", "--num_clips", "2", "--batch_size", "1", "--frames", "8", "--side", "16",
                 "--num_classes", "10", "--opt_path", src, "--step", "2", "--attack_method", "BIM"])
    pairs = os.path.join(src, "tiny-BIM-2-")
    atk = ft.main(["--synthetic", "--white_model", "tpn_tiny", "--num_classes", "10", "--used_adv", pairs, "--used_ori", pairs,
                   "--opt_path", dst, "--steps", "3"])
    for idx in range(2):
        out = np.load(os.path.join(dst, "%d-adv.npy" % idx))
        assert out.shape == (3, 8, 16, 16) and out.dtype == np.float32 and np.isfinite(out).all()
    clean, _ = synth.clip(1, b=1, f=8, h=16, w=16)
    _bounds_ok(clean.numpy(), atk.last_adv.cpu().numpy())                 # the plain (unscrambled) clip of the last call
    assert sorted(atk.loss_info["..."]) == [0, 1, 2]
