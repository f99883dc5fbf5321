"""GPU parity of the memory-bound kernels (K1, K2, K3a/b/c) against the CPU oracle, through the C ABI.

Bars (BASELINE.json north_star):
  * K3 update kernels: BIT-EXACT vs oracle/i2v_oracle.c on identical inputs.
  * K1 cosine: cos within 1e-5 relative of the float64 arbiter (we get ~1e-7); gradient closer to the
    float64 arbiter than torch-f32 autograd is.
  * K2: within 1 ulp of the oracle (device exp() vs libm exp() in float64, rounded to f32).
"""
import os

import numpy as np
import pytest
import torch

from conftest import ulp_diff
from i2v_b200 import capi
from oracle import oracle as O

pytestmark = pytest.mark.gpu
EPS = 16 / 255
DEV = "cuda"


def bits_equal(a, b):
    a = np.ascontiguousarray(a, dtype=np.float32).view(np.int32)
    b = np.ascontiguousarray(b, dtype=np.float32).view(np.int32)
    return np.array_equal(a, b)


def gpu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


LAYOUTS = [
    # shape, inner, channels
    ((5, 3, 16, 20), 320, 3),          # [N,3,H,W] planar, inner % 4 == 0
    ((3, 3, 7, 5), 35, 3),             # planar, inner % 4 != 0 and n % 4 != 0 (scalar tail)
    ((2, 3, 4, 6, 6), 144, 3),         # [B,3,T,H,W]
    ((4, 9, 9, 3), 1, 3),              # NHWC, 3 channels interleaved
    ((4, 8, 8, 4), 1, 4),              # NHWC4 (padded channel stays 0)
    ((0, 3, 8, 8), 64, 3),             # empty
]


def _rand01(shape, seed):
    return np.random.default_rng(seed).random(shape, dtype=np.float32)


@pytest.mark.parametrize("shape,inner,channels", LAYOUTS)
def test_denorm_normalize_compose_bit_exact(shape, inner, channels):
    rng = np.random.default_rng(1)
    inp = (rng.standard_normal(shape) * 2).astype(np.float32)
    x = torch.empty(shape, device=DEV)
    capi.denorm(gpu(inp), x, inner, channels)
    want = O.denorm(inp, inner, channels)
    assert bits_equal(x.cpu().numpy(), want)
    out = torch.empty(shape, device=DEV)
    capi.normalize(gpu(want), out, inner, channels)
    assert bits_equal(out.cpu().numpy(), O.normalize(want, inner, channels))
    x01 = _rand01(shape, 2)
    mod = ((rng.random(shape) - 0.5) * 0.3).astype(np.float32)      # exceeds ±eps on ~60 % of elements
    capi.compose_norm(gpu(x01), gpu(mod), out, EPS, inner, channels)
    assert bits_equal(out.cpu().numpy(), O.compose_norm(x01, mod, EPS, inner, channels))


@pytest.fixture(params=["cuda", "cpu"])
def adam_arith(request):
    """Both of torch's Adam arithmetics (include/i2v_b200.h: i2v_set_adam_arithmetic): K3a must reproduce the oracle bit for
    bit in either; 'cuda' is the product default."""
    prev = capi.get_adam_arithmetic()
    capi.set_adam_arithmetic(request.param)
    yield request.param
    capi.set_adam_arithmetic(prev)


@pytest.mark.parametrize("shape,inner,channels", LAYOUTS)
def test_adam_compose_bit_exact_over_steps(shape, inner, channels, adam_arith):
    """Five Adam steps with gradients spanning 1e-9..1e-2 (step-1 gradients are ~1e-8),
    pixels at the [0,1] borders and modifiers crossing ±eps so that both clamp masks are exercised."""
    rng = np.random.default_rng(3)
    x01 = _rand01(shape, 4)
    x01[x01 < 0.05] = 0.0
    x01[x01 > 0.95] = 1.0
    if channels == 4:
        x01[..., 3] = 0
    mod = np.full(shape, np.float32(0.01 / 255))
    m = np.zeros(shape, np.float32)
    v = np.zeros(shape, np.float32)
    d = dict(g=None, m=gpu(m), v=gpu(v), mod=gpu(mod), x=gpu(x01), out=torch.empty(shape, device=DEV))
    table = capi.adam_step_table(5, 0.02).to(DEV)
    step_idx = torch.zeros(1, dtype=torch.int32, device=DEV)
    for step in range(1, 6):
        g = (rng.standard_normal(shape) * 10.0 ** rng.uniform(-9, -2, size=shape)).astype(np.float32)
        g[rng.random(shape) < 0.05] = 0.0
        m, v, mod, want = O.adam_compose(g, m, v, mod, x01, EPS, inner, step, 0.02, channels=channels, arith=adam_arith)
        if step % 2:   # alternate the two entry points: scalars by value / device table
            capi.adam_compose(gpu(g), d["m"], d["v"], d["mod"], d["x"], d["out"], EPS, inner, step, 0.02, channels=channels)
        else:
            capi.adam_compose_table(gpu(g), d["m"], d["v"], d["mod"], d["x"], d["out"], EPS, inner, table, step_idx,
                                    channels=channels)
        capi.step_advance(step_idx)
        if channels == 4:   # padded lanes: state untouched, output forced to 0 on both sides
            sel = np.ones(shape, bool); sel[..., 3] = False
        else:
            sel = np.ones(shape, bool)
        for name, ref in (("m", m), ("v", v), ("mod", mod), ("out", want)):
            got = d[name].cpu().numpy()
            assert bits_equal(got[sel], ref[sel]), (name, step)
    assert int(step_idx.item()) == 5
    if mod.size:
        assert (np.abs(mod) > EPS).any(), "test should drive some modifiers outside the eps ball"


def test_adam_bit_exact_vs_torch_cuda_adam():
    """K3a against the optimiser the reference really hits: `torch.optim.Adam([modifier], lr)` on CUDA
    (image_attacks.py:306, `.cuda()` hard-coded at 304; with every tensor on the GPU torch takes its foreach path).  The
    same gradients go into both for six steps (1e-9..1e-2 magnitudes, 5 % exact zeros, both clamp masks active): exp_avg,
    exp_avg_sq and the modifier must be BIT-identical after every step in the default ('cuda') arithmetic.  torch's CPU
    kernels group addcmul_ / addcdiv_ differently (tools/adam_cuda_probe.py) — that variant is pinned to the CPU fixtures
    in tests/test_oracle_golden.py and to the oracle above."""
    assert capi.get_adam_arithmetic() == "cuda"
    shape, inner = (4, 3, 56, 56), 56 * 56
    rng = np.random.default_rng(9)
    x01 = _rand01(shape, 10)
    x01[x01 < 0.03] = 0.0
    x01[x01 > 0.97] = 1.0
    lr = 0.02
    mod = torch.nn.Parameter(torch.full(shape, 0.01 / 255, device=DEV))
    opt = torch.optim.Adam([mod], lr=lr)                      # default arguments, as the reference constructs it
    d = dict(m=torch.zeros(shape, device=DEV), v=torch.zeros(shape, device=DEV), mod=torch.full(shape, 0.01 / 255, device=DEV),
             x=gpu(x01), out=torch.empty(shape, device=DEV))
    std = np.asarray(O.STD, dtype=np.float32)[None, :, None, None]
    for step in range(1, 7):
        g = (rng.standard_normal(shape) * 10.0 ** rng.uniform(-9, -2, size=shape)).astype(np.float32)
        g[rng.random(shape) < 0.05] = 0.0
        # K3a takes dcost/dtrue_image and applies the chain rule through compose (1/std, closed-interval clamp masks)
        # itself; torch gets that same dcost/dmodifier
        modh = d["mod"].cpu().numpy()
        ssum = x01 + np.clip(modh, -np.float32(EPS), np.float32(EPS))
        live = (np.abs(modh) <= np.float32(EPS)) & (ssum >= 0) & (ssum <= 1)
        gm = np.where(live, g / std, np.float32(0)).astype(np.float32)
        mod.grad = gpu(gm)
        opt.step()
        before = [d[k].cpu().numpy() for k in ("m", "v", "mod")]
        capi.adam_compose(gpu(g), d["m"], d["v"], d["mod"], d["x"], d["out"], EPS, inner, step, lr)
        st = opt.state[mod]
        assert bits_equal(d["m"].cpu().numpy(), st["exp_avg"].cpu().numpy()), step
        assert bits_equal(d["v"].cpu().numpy(), st["exp_avg_sq"].cpu().numpy()), step
        assert bits_equal(d["mod"].cpu().numpy(), mod.detach().cpu().numpy()), step
        # and the C oracle in its 'cuda' arithmetic is the same function
        m_o, v_o, mod_o, want = O.adam_compose(g, before[0], before[1], before[2], x01, EPS, inner, step, lr, arith="cuda")
        assert bits_equal(m_o, d["m"].cpu().numpy()) and bits_equal(v_o, d["v"].cpu().numpy())
        assert bits_equal(mod_o, d["mod"].cpu().numpy()) and bits_equal(want, d["out"].cpu().numpy())
    assert (np.abs(d["mod"].cpu().numpy()) > EPS).any(), "the test should drive some modifiers outside the eps ball"


@pytest.mark.parametrize("shape,inner", [((2, 3, 4, 6, 6), 144), ((1, 3, 3, 5, 7), 105), ((3, 3, 8, 8), 64)])
@pytest.mark.parametrize("project", [True, False])
def test_sign_step_project_bit_exact(shape, inner, project):
    rng = np.random.default_rng(5)
    x01 = _rand01(shape, 6)
    adv = O.normalize(np.clip(x01 + (rng.random(shape).astype(np.float32) - 0.5) * 0.2, 0, 1), inner)
    g = rng.standard_normal(shape).astype(np.float32)
    g[rng.random(shape) < 0.1] = 0.0          # sign(0) = 0
    a = gpu(adv)
    capi.sign_step_project(a, gpu(g), gpu(x01) if project else None, EPS / 10, EPS, inner, project=project)
    want = O.sign_step_project(adv, g, x01 if project else None, EPS / 10, EPS, inner, project=project)
    assert bits_equal(a.cpu().numpy(), want)


@pytest.mark.parametrize("shape", [(2, 3, 32, 12, 12), (1, 3, 16, 7, 5), (3, 3, 4, 28, 28)])
@pytest.mark.parametrize("clip_level", [False, True])
def test_mi_update_bit_exact(shape, clip_level):
    rng = np.random.default_rng(7)
    inner = shape[2] * shape[3] * shape[4]
    x01 = _rand01(shape, 8)
    adv = O.normalize(x01, inner)
    g = (rng.standard_normal(shape) * 1e-3).astype(np.float32)
    mom = rng.standard_normal(shape).astype(np.float32)
    norm = torch.empty((shape[0],) if clip_level else (shape[0], shape[2]), device=DEV)
    capi.frame_absmean(gpu(g), norm, clip_level=clip_level)
    want_norm = O.frame_absmean(g, clip_level)
    assert ulp_diff(norm.cpu().numpy(), want_norm).max() <= 1       # tree vs sequential float64 sum
    a, mo = gpu(adv), gpu(mom)
    capi.mi_sign_step_project(a, gpu(g), mo, gpu(want_norm), gpu(x01), 1.0, EPS / 10, EPS, clip_level=clip_level)
    want_adv, want_mom = O.mi_sign_step_project(adv, g, mom, want_norm, x01, 1.0, EPS / 10, EPS, clip_level)
    assert bits_equal(mo.cpu().numpy(), want_mom)
    assert bits_equal(a.cpu().numpy(), want_adv)


def test_update_kernels_full_size_bit_exact():
    """Config-2 frame size (224x224) on 48 frames: 7.2 M elements through K3a and K3b."""
    shape = (48, 3, 224, 224)
    inner = 224 * 224
    rng = np.random.default_rng(9)
    x01 = _rand01(shape, 10)
    g = (rng.standard_normal(shape) * 1e-6).astype(np.float32)
    mod = ((rng.random(shape) - 0.5) * 0.2).astype(np.float32)
    m = (rng.standard_normal(shape) * 1e-6).astype(np.float32)
    v = (rng.random(shape) * 1e-12).astype(np.float32)
    dm, dv, dmod, out = gpu(m), gpu(v), gpu(mod), torch.empty(shape, device=DEV)
    capi.adam_compose(gpu(g), dm, dv, dmod, gpu(x01), out, EPS, inner, 17, 0.005)
    m2, v2, mod2, want = O.adam_compose(g, m, v, mod, x01, EPS, inner, 17, 0.005, arith=capi.get_adam_arithmetic())
    assert bits_equal(dm.cpu().numpy(), m2) and bits_equal(dv.cpu().numpy(), v2)
    assert bits_equal(dmod.cpu().numpy(), mod2) and bits_equal(out.cpu().numpy(), want)
    shape5 = (2, 3, 24, 224, 224)
    adv = want.reshape(2, 24, 3, 224, 224).transpose(0, 2, 1, 3, 4).copy()
    x5 = x01.reshape(2, 24, 3, 224, 224).transpose(0, 2, 1, 3, 4).copy()
    g5 = g.reshape(shape5)
    a = gpu(adv)
    capi.sign_step_project(a, gpu(g5), gpu(x5), EPS / 10, EPS, 24 * inner)
    assert bits_equal(a.cpu().numpy(), O.sign_step_project(adv, g5, x5, EPS / 10, EPS, 24 * inner))


# ------------------------------------------------------------------------------------- K1
def _torch_cos(a, b, dtype):
    at = torch.tensor(a, dtype=dtype, requires_grad=True)
    bt = torch.tensor(b, dtype=dtype)
    c = torch.nn.functional.cosine_similarity(at.view(at.shape[0], -1), bt.view(bt.shape[0], -1))
    c.sum().backward()
    return c.detach().numpy(), at.grad.numpy()


def _k1(a, b, w=1.0, relu_mask=False, w_dev=None, want_grad=True):
    da, db = gpu(a), gpu(b)
    grad = torch.empty_like(da) if want_grad else None
    cos = torch.empty(a.shape[0], device=DEV)
    capi.cosine_loss_grad(da, db, grad, cos, w_dev=w_dev, w_host=w, relu_mask=relu_mask)
    return cos.cpu().numpy(), (grad.cpu().numpy() if want_grad else None)


@pytest.mark.parametrize("N,D", [(3, 512 * 28 * 28), (5, 128 * 27 * 27), (2, 4099), (7, 384 * 13 * 13), (1, 12)])
@pytest.mark.parametrize("near", [False, True])
def test_cosine_loss_grad_vs_float64(N, D, near):
    """near=True is the step-1 regime: a = b + 1e-4 noise, cos ~ 1 - 5e-9, gradient ~1e-8 of its terms."""
    rng = np.random.default_rng(11)
    b = np.maximum(rng.standard_normal((N, D)), 0).astype(np.float32)        # post-ReLU-like
    if near:
        a = (b * (1 + 1e-4 * rng.standard_normal((N, D)))).astype(np.float32)
    else:
        a = np.maximum(rng.standard_normal((N, D)), 0).astype(np.float32)
    cos64, grad64 = O.cosine_loss_grad_f64(a, b)
    cos, grad = _k1(a, b)
    assert np.abs(cos - cos64).max() <= 1e-5 * np.abs(cos64).max()           # north-star tolerance
    assert np.abs(cos - cos64).max() <= 1.2e-7                               # what we actually get
    scale = np.abs(grad64).max(axis=1, keepdims=True)
    err = (np.abs(grad - grad64) / scale).max()
    # far from cancellation the gradient is correctly rounded (~1e-7).  In the step-1 regime |g| is ~1e-4 of
    # its two terms, so the ~1e-9..1e-7 relative error of the float32-partial frame sums is amplified 1e4x;
    # torch-f32 autograd (checked below) is another 10-1000x further from float64 there.
    assert err <= (1e-3 if near else 1e-6), err
    # torch-f32 autograd (the reference's arithmetic) is no closer to float64 than we are
    _, g32 = _torch_cos(a, b, torch.float32)
    err32 = (np.abs(g32 - grad64) / scale).max()
    assert err <= err32 * 1.0001 + 1e-7, (err, err32)
    # and torch-f64 autograd agrees with the analytic float64 gradient of the oracle
    c64t, g64t = _torch_cos(a, b, torch.float64)
    assert np.allclose(c64t, cos64, rtol=1e-12) and (np.abs(g64t - grad64) / scale).max() < 1e-9


def test_cosine_weights_mask_and_loss_only():
    rng = np.random.default_rng(12)
    N, D = 4, 256 * 14 * 14
    a = rng.standard_normal((N, D)).astype(np.float32)          # signed, so the ReLU mask bites
    b = rng.standard_normal((N, D)).astype(np.float32)
    cos64, grad64 = O.cosine_loss_grad_f64(a, b, w=0.125, relu_mask=True)
    w_dev = torch.tensor([0.125], device=DEV)
    cos, grad = _k1(a, b, w=99.0, relu_mask=True, w_dev=w_dev)   # w_dev wins over w_host
    assert np.abs(cos - cos64).max() <= 1.2e-7
    assert (grad[a <= 0] == 0).all()
    assert (np.abs(grad - grad64) / np.abs(grad64).max()).max() <= 1e-6
    cos2, none = _k1(a, b, want_grad=False)
    assert none is None and bits_equal(cos2, cos)


def test_cosine_zero_features_and_empty():
    """All-zero frame: |a| is clamped to 1e-8, cos = 0, no NaN (torch 2.x cosine_similarity semantics)."""
    N, D = 3, 4096
    rng = np.random.default_rng(13)
    a = rng.standard_normal((N, D)).astype(np.float32)
    b = rng.standard_normal((N, D)).astype(np.float32)
    a[1] = 0
    b[2] = 0
    cos, grad = _k1(a, b)
    cos64, grad64 = O.cosine_loss_grad_f64(a, b)
    assert np.isfinite(cos).all() and np.isfinite(grad).all()
    assert cos[1] == 0 and cos[2] == 0
    scale = np.abs(grad64).max(axis=1, keepdims=True) + 1e-300
    assert (np.abs(grad - grad64) / scale).max() <= 1e-6
    assert (grad[2] == 0).all()                                 # b = 0: alpha*b - beta*a with cos = 0
    ct, gt = _torch_cos(a, b, torch.float64)
    assert np.allclose(grad64, gt, rtol=1e-9, atol=1e-30)       # the oracle follows torch's clamping
    e = torch.empty(0, 16, device=DEV)
    capi.cosine_loss_grad(e, e, torch.empty_like(e), torch.empty(0, device=DEV))


@pytest.mark.parametrize("cluster", ["1", "2", "4", "8", "16"])
def test_cosine_cluster_sizes_agree(cluster):
    """Every cluster size (DSMEM reduction width, shared-memory stash coverage) gives the same answer;
    full ResNet layer2 feature size, N not a multiple of anything."""
    rng = np.random.default_rng(14)
    N, D = 5, 512 * 28 * 28
    a = np.maximum(rng.standard_normal((N, D)), 0).astype(np.float32)
    b = np.maximum(rng.standard_normal((N, D)), 0).astype(np.float32)
    cos64, grad64 = O.cosine_loss_grad_f64(a, b)
    old = os.environ.get("I2V_COS_CLUSTER")
    os.environ["I2V_COS_CLUSTER"] = cluster
    try:
        cos, grad = _k1(a, b)
    finally:
        if old is None:
            os.environ.pop("I2V_COS_CLUSTER")
        else:
            os.environ["I2V_COS_CLUSTER"] = old
    assert np.abs(cos - cos64).max() <= 1.2e-7
    assert (np.abs(grad - grad64) / np.abs(grad64).max()).max() <= 1e-6


# ------------------------------------------------------------------------------------- K2
@pytest.mark.parametrize("L", [1, 2, 8, 16, 32])
@pytest.mark.parametrize("momentum", [0.0, 0.5])
def test_layer_reweight_and_sums(L, momentum):
    rng = np.random.default_rng(15)
    N = 37
    coeffs = np.ones(L, np.float32)
    prev = np.ones(L, np.float32)
    dc, dp = gpu(coeffs), gpu(prev)
    w_out = torch.empty(L, device=DEV)
    log = torch.zeros(3, L, device=DEV)
    cost_log = torch.zeros(3, device=DEV)
    step_idx = torch.zeros(1, dtype=torch.int32, device=DEV)
    for step in range(3):
        coeffs, w = O.layer_reweight(coeffs, prev, momentum)
        capi.layer_reweight(dc, dp, momentum, w_out, log, step_idx)
        assert ulp_diff(dc.cpu().numpy(), coeffs).max() <= 1
        assert ulp_diff(w_out.cpu().numpy(), w).max() <= 1
        assert bits_equal(log[step].cpu().numpy(), dc.cpu().numpy())
        coeffs = dc.cpu().numpy().copy()           # teacher-force so 1-ulp differences do not accumulate
        cosv = rng.uniform(0.2, 1.0, size=(L, N)).astype(np.float32)
        for mode, ce in ((0, False), (1, False), (1, True)):
            cost, prev_o = O.layer_sums(cosv, coeffs, mode, ce)
            dprev = torch.zeros(L, device=DEV)
            capi.layer_sums(gpu(cosv), dc, dprev, cost_log, step_idx, mode=mode, coef_CE=ce)
            assert ulp_diff(cost_log[step].cpu().numpy(), np.float32(cost)).max() <= 1
            if mode == 1:
                assert ulp_diff(dprev.cpu().numpy(), prev_o).max() <= 1
        prev = prev_o
        dp.copy_(gpu(prev))
        capi.step_advance(step_idx)


@pytest.mark.parametrize("shape,chunks", [((2, 512, 28, 28), 1), ((5, 64, 7, 9), 2), ((3, 1, 1, 7), 3), ((1, 128, 56, 56), 1)])
@pytest.mark.parametrize("relu_mask", [False, True])
def test_std_loss_grad_vs_float64(shape, chunks, relu_mask):
    """K6 (Dispersion Reduction): unbiased std of the whole map and its gradient against the float64 oracle, with the
    map fed in frame chunks (ragged sizes included); repeated runs are bit-identical (fixed reduction order)."""
    rng = np.random.default_rng(3)
    a = np.maximum(rng.standard_normal(shape) + 0.3, 0).astype(np.float32)
    sd64, mean64, g64 = O.std_loss_grad_f64(a, relu_mask=relu_mask)
    ad = torch.from_numpy(a).cuda()
    ws = capi.std_workspace(ad.device)
    outs = []
    for rep in range(2):
        acc = torch.zeros(2, dtype=torch.float64, device="cuda")
        stats = torch.zeros(4, device="cuda")
        cost = torch.full((3,), 7.0, device="cuda")
        step = torch.tensor([1], dtype=torch.int32, device="cuda")
        bounds = np.linspace(0, shape[0], chunks + 1).astype(int)
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            if hi > lo:
                capi.std_accumulate(ad[lo:hi], ws, acc)
        capi.std_finalize(acc, a.size, stats, cost, step, add_to_cost=True)
        grad = torch.full_like(ad, float("nan"))
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            if hi > lo:
                capi.std_grad(ad[lo:hi], grad[lo:hi], stats, relu_mask=relu_mask)
        outs.append((stats.cpu().numpy(), grad.cpu().numpy(), cost.cpu().numpy()))
    (st, gr, cost), (st2, gr2, _) = outs
    assert np.array_equal(st, st2) and np.array_equal(gr, gr2)
    assert abs(st[0] - mean64) <= 2e-7 * abs(mean64) and abs(st[1] - sd64) <= 2e-7 * sd64
    assert cost[0] == 7.0 and cost[2] == 7.0 and abs(cost[1] - (7.0 + sd64)) <= 1e-6
    assert np.abs(gr - g64).max() <= 5e-7 * np.abs(g64).max()


# ---- K8 temporal translation, K3d / K9 ILAF ----------------------------------------------
@pytest.mark.parametrize("shape", [(1, 3, 32, 12, 12), (2, 3, 8, 7, 5), (1, 3, 16, 28, 28), (0, 3, 4, 4, 4)])
def test_temporal_shift_and_combine_bit_exact(shape):
    rng = np.random.default_rng(21)
    adv = rng.standard_normal(shape).astype(np.float32)
    T = shape[2]
    moves = [-3, -2, -1, 0, 1, 2, 3, T + 5, -(T + 1)]
    out = torch.full((len(moves),) + shape, float("nan"), device=DEV)
    capi.temporal_shift_stack(gpu(adv), out, moves)
    if adv.size:
        assert bits_equal(out.cpu().numpy(), O.temporal_shift_stack(adv, moves))
    D = 7
    grads = rng.standard_normal((D,) + shape).astype(np.float32)
    k = rng.random(D).astype(np.float32)
    k /= k.sum()
    mv = [-3, -2, -1, 0, 1, 2, 3]
    for weight in (0.5, 0.3, 0.0, 1.0):
        got = torch.full(shape, float("nan"), device=DEV)
        capi.temporal_combine(gpu(grads), k, mv, weight, got)
        if adv.size:
            assert bits_equal(got.cpu().numpy(), O.temporal_combine(grads, k, mv, weight)), weight


@pytest.mark.parametrize("shape,inner", [((1, 3, 8, 12, 12), 8 * 144), ((2, 3, 3, 5, 7), 105), ((3, 3, 8, 8), 64), ((0, 3, 2, 4, 4), 32)])
def test_sign_descent_compose_bit_exact(shape, inner):
    rng = np.random.default_rng(31)
    x = rng.random(shape, dtype=np.float32)
    mod = (rng.standard_normal(shape) * 0.08).astype(np.float32)
    g = rng.standard_normal(shape).astype(np.float32)
    if g.size:
        g.reshape(-1)[:5] = 0.0
        x.reshape(-1)[5:9] = [0.0, 1.0, 0.0, 1.0]
    md, img = gpu(mod), torch.full(shape, float("nan"), device=DEV)
    for step in range(3):
        capi.sign_descent_compose(gpu(g), md, gpu(x), img, EPS, 0.005, inner)
        mod, want = O.sign_descent_compose(g, mod, x, EPS, 0.005, inner)
        assert bits_equal(md.cpu().numpy(), mod) and bits_equal(img.cpu().numpy(), want), step
        g = np.roll(g, 3).copy()


@pytest.mark.parametrize("n", [16 * 16 * 6 * 6, 4099, 512 * 8 * 14 * 14, 3])
def test_ila_loss_grad_vs_float64(n):
    rng = np.random.default_rng(41)
    f = rng.standard_normal(n).astype(np.float32)
    o = (f + 0.3 * rng.standard_normal(n)).astype(np.float32)
    d0 = rng.standard_normal(n)
    n0 = float(np.linalg.norm(d0)) * 0.8
    d0 = (d0 / np.linalg.norm(d0)).astype(np.float32)
    loss64, grad64 = O.ila_loss_grad_f64(f, o, d0, n0)
    ws = capi.ila_workspace(torch.device(DEV))
    stats = torch.zeros(4, device=DEV)
    cost = torch.zeros(3, device=DEV)
    idx = torch.tensor([1], device=DEV, dtype=torch.int32)
    capi.ila_loss(gpu(f), gpu(o), gpu(d0), n0, ws, stats, cost, idx)
    capi.ila_loss(gpu(f), gpu(o), gpu(d0), n0, ws, stats, cost, idx, add_to_cost=True)
    grad = capi.ila_grad(gpu(f), gpu(o), gpu(d0), torch.empty(n, device=DEV), stats).cpu().numpy()
    st = stats.cpu().numpy()
    assert abs(st[2] - loss64) <= 2e-6 * abs(loss64)
    assert abs(st[3] - np.linalg.norm(f.astype(np.float64) - o)) <= 1e-6 * st[3]
    c = cost.cpu().numpy()
    assert c[0] == 0 and c[2] == 0 and abs(c[1] - 2 * loss64) <= 4e-6 * abs(loss64)
    assert np.abs(grad - grad64).max() <= 3e-6 * np.abs(grad64).max()
