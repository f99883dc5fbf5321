"""GPU parity of the backbone kernels (K4/K5): convolution forward / data-gradient, max pooling, and the
native engine's truncated forward + input gradient, against PyTorch float64 on the CPU (the arbiter)
and float32 (the reference's arithmetic).  Tolerance: max |err| <= 2e-5 * max |ref| for a single layer in
FP32 mode (observed ~1e-6); the float32 torch result must not be meaningfully closer than ours."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from i2v_b200 import backbones, capi, engines
from i2v_b200.engine_native import NativeEngine, _pad_cols, _pad_vec

pytestmark = pytest.mark.gpu
DEV = "cuda"

#        N  H   W   Cin Cout k s p  x_nchw
SHAPES = [
    (2, 56, 56, 64, 64, 1, 1, 0, False),     # layer1.0.conv1
    (2, 56, 56, 64, 64, 3, 1, 1, False),     # layer1.x.conv2
    (2, 56, 56, 256, 128, 1, 1, 0, False),   # layer2.0.conv1
    (2, 56, 56, 128, 128, 3, 2, 1, False),   # layer2.0.conv2 (strided 3x3)
    (2, 56, 56, 256, 512, 1, 2, 0, False),   # layer2.0.downsample (strided 1x1)
    (3, 28, 28, 128, 512, 1, 1, 0, False),   # layer2.x.conv3
    (1, 64, 64, 3, 64, 7, 2, 3, True),       # ResNet stem, reads / writes the [N,3,H,W] image
    (1, 64, 64, 3, 64, 11, 4, 2, True),      # AlexNet conv1
    (2, 15, 15, 64, 192, 5, 1, 2, False),    # AlexNet conv2
    (2, 27, 27, 64, 16, 1, 1, 0, False),     # SqueezeNet squeeze (16 channels)
    (1, 13, 13, 48, 192, 3, 1, 1, False),    # SqueezeNet expand3x3, odd spatial size
    (1, 7, 9, 8, 12, 3, 1, 1, False),        # tiny ragged
]


def _ref_conv(x, w, scale, shift, stride, pad, residual, relu, dtype):
    y = F.conv2d(x.to(dtype), (w * scale.view(-1, 1, 1, 1)).to(dtype), None, stride, pad) + shift.to(dtype).view(1, -1, 1, 1)
    if residual is not None:
        y = y + residual.to(dtype)
    return torch.relu(y) if relu else y


@pytest.mark.parametrize("shape", SHAPES)
def test_conv_fwd_and_dgrad_simt(shape):
    N, H, W, Cin, Cout, k, s, p, x_nchw = shape
    g = torch.Generator().manual_seed(hash(shape) % 1000)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    scale = torch.rand(Cout, generator=g) + 0.5
    shift = torch.randn(Cout, generator=g) * 0.1
    P, Q = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    res = torch.randn(N, Cout, P, Q, generator=g)
    ws = w * scale.view(-1, 1, 1, 1)
    d = capi.ConvDesc(N, H, W, Cin, Cout, k, k, s, p, P, Q)
    bf = _pad_cols(ws.permute(2, 3, 1, 0).reshape(k * k * Cin, Cout)).to(DEV)
    bd = _pad_cols(ws.permute(2, 3, 0, 1).reshape(k * k * Cout, Cin)).to(DEV)
    bias = _pad_vec(shift).to(DEV)
    xd = (x if x_nchw else x.permute(0, 2, 3, 1)).contiguous().to(DEV)
    resd = res.permute(0, 2, 3, 1).contiguous().to(DEV)
    for use_res, relu in ((False, False), (True, True)):
        y = torch.empty(N, P, Q, Cout, device=DEV)
        capi.conv_fwd_simt(d, xd, bf, bias, resd if use_res else None, y, relu=relu, x_nchw=x_nchw)
        ref64 = _ref_conv(x, w, scale, shift, s, p, res if use_res else None, relu, torch.float64)
        ref32 = _ref_conv(x, w, scale, shift, s, p, res if use_res else None, relu, torch.float32)
        got = y.permute(0, 3, 1, 2).cpu().double()
        err = (got - ref64).abs().max() / ref64.abs().max()
        err32 = (ref32.double() - ref64).abs().max() / ref64.abs().max()
        assert err <= 2e-5 and err <= 20 * err32 + 1e-6, (err, err32)
    # data gradient: dx = conv_transpose(dy, w*scale) [+ addend], masked by a forward activation
    dy = torch.randn(N, Cout, P, Q, generator=g)
    addend = torch.randn(N, Cin, H, W, generator=g)
    act = torch.randn(N, Cin, H, W, generator=g)
    dyd = dy.permute(0, 2, 3, 1).contiguous().to(DEV)
    lay = (lambda t: t.contiguous().to(DEV)) if x_nchw else (lambda t: t.permute(0, 2, 3, 1).contiguous().to(DEV))
    for use_add, use_mask in ((False, False), (True, True)):
        dx = torch.empty_like(xd)
        capi.conv_dgrad_simt(d, dyd, bd, lay(addend) if use_add else None, lay(act) if use_mask else None, dx, x_nchw=x_nchw)
        ref64 = torch.nn.grad.conv2d_input((N, Cin, H, W), ws.double(), dy.double(), s, p)
        if use_add:
            ref64 = ref64 + addend.double()
        if use_mask:
            ref64 = ref64 * (act > 0).double()
        got = (dx if x_nchw else dx.permute(0, 3, 1, 2)).cpu().double()
        err = (got - ref64).abs().max() / ref64.abs().max()
        assert err <= 2e-5, err


@pytest.mark.parametrize("H,W,k,s,p,ceil", [(112, 112, 3, 2, 1, False), (55, 55, 3, 2, 0, False), (27, 31, 3, 2, 0, True),
                                            (56, 56, 2, 2, 0, False), (13, 13, 3, 2, 0, True)])
def test_maxpool_fwd_bwd(H, W, k, s, p, ceil):
    g = torch.Generator().manual_seed(3)
    N, C = 2, 16
    x = torch.relu(torch.randn(N, C, H, W, generator=g))          # post-ReLU input: zeros tie, first wins
    mp = torch.nn.MaxPool2d(k, s, p, ceil_mode=ceil, return_indices=True)
    xr = x.clone().requires_grad_(True)
    yr, idx = mp(xr)
    P, Q = yr.shape[2:]
    dy = torch.randn(N, C, P, Q, generator=g)
    yr.backward(dy)
    xd = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    y = torch.empty(N, P, Q, C, device=DEV)
    am = torch.empty(N, P, Q, C, device=DEV, dtype=torch.uint8)
    capi.maxpool_fwd(xd, y, am, k, s, p)
    assert torch.equal(y.permute(0, 3, 1, 2).cpu(), yr.detach())
    dx = torch.empty_like(xd)
    capi.maxpool_bwd(dy.permute(0, 2, 3, 1).contiguous().to(DEV), am, None, dx, k, s, p)
    ref = xr.grad
    got = dx.permute(0, 3, 1, 2).cpu()
    # ties between equal zeros may be broken differently; the gradient through them is killed by the ReLU mask
    nz = x > 0
    assert torch.allclose(got[nz], ref[nz], rtol=1e-6, atol=1e-7)
    capi.maxpool_bwd(dy.permute(0, 2, 3, 1).contiguous().to(DEV), am, xd, dx, k, s, p)
    got = dx.permute(0, 3, 1, 2).cpu()
    assert torch.allclose(got, ref * nz, rtol=1e-6, atol=1e-7)
    # the same mask taken from the pooled output (1[y > 0] = 1[x[argmax] > 0]) is bit-identical, also with dx +=
    dx2 = torch.full_like(xd, float("nan"))
    capi.maxpool_bwd(dy.permute(0, 2, 3, 1).contiguous().to(DEV), am, y, dx2, k, s, p, mask_pooled=True)
    assert torch.equal(dx2, dx)
    base = torch.randn_like(xd)
    acc1, acc2 = base.clone(), base.clone()
    capi.maxpool_bwd(dy.permute(0, 2, 3, 1).contiguous().to(DEV), am, xd, acc1, k, s, p, accumulate=True)
    capi.maxpool_bwd(dy.permute(0, 2, 3, 1).contiguous().to(DEV), am, y, acc2, k, s, p, accumulate=True, mask_pooled=True)
    assert torch.equal(acc1, acc2)
    # mark_dead: the ReLU-backward mask folded into the argmax plane by the forward pass (255 = no winner) — same y,
    # and the unmasked backward over that plane is bit-identical to the masked one
    y3 = torch.empty_like(y)
    am3 = torch.empty_like(am)
    capi.maxpool_fwd(xd, y3, am3, k, s, p, mark_dead=True)
    assert torch.equal(y3, y)
    assert torch.equal(am3 == 255, ~(y > 0)) and torch.equal(am3[y > 0], am[y > 0])
    dx3 = torch.full_like(xd, float("nan"))
    capi.maxpool_bwd(dy.permute(0, 2, 3, 1).contiguous().to(DEV), am3, None, dx3, k, s, p)
    assert torch.equal(dx3, dx)


def test_copy_channels():
    a = torch.randn(5, 7, 3, 16, device=DEV)
    b = torch.randn(5, 7, 3, 24, device=DEV)
    cat = torch.empty(5, 7, 3, 40, device=DEV)
    capi.copy_channels(a, cat, 0, 0, 16)
    capi.copy_channels(b, cat, 0, 16, 24)
    assert torch.equal(cat, torch.cat([a, b], -1))
    back = torch.ones(5, 7, 3, 24, device=DEV)
    capi.copy_channels(cat, back, 16, 0, 24, accumulate=True)
    assert torch.equal(back, b + 1)


def test_bn_relu_and_avgpool2():
    """The DenseNet pieces (csrc/dense.cu) against their torch statements: pre-activation BatchNorm(eval) + ReLU of a
    channel prefix of a concat buffer (zero-padded dense output), 2x2 average pooling into / out of a channel slice."""
    g = torch.Generator().manual_seed(3)
    cat = torch.randn(3, 5, 7, 96, generator=g).to(DEV)
    scale = (torch.rand(64, generator=g) + 0.5).to(DEV)
    shift = torch.randn(64, generator=g).to(DEV)
    t = torch.full((3, 5, 7, 128), float("nan"), device=DEV)
    capi.bn_relu(cat, 64, scale, shift, t)
    want = torch.relu(torch.addcmul(shift, cat[..., :64], scale))
    assert torch.allclose(t[..., :64], want, rtol=1e-6, atol=1e-7) and torch.equal(t[..., 64:], torch.zeros_like(t[..., 64:]))
    x = torch.randn(2, 7, 6, 32, generator=g).to(DEV)                    # odd height: floor mode drops the last row
    y = torch.zeros(2, 3, 3, 80, device=DEV)
    capi.avgpool2_fwd(x, y, dst_off=16)
    ref = torch.nn.functional.avg_pool2d(x.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
    assert torch.allclose(y[..., 16:48], ref, rtol=1e-6, atol=1e-7)
    assert torch.equal(y[..., :16], torch.zeros_like(y[..., :16])) and torch.equal(y[..., 48:], torch.zeros_like(y[..., 48:]))
    dy = torch.randn(2, 3, 3, 80, generator=g).to(DEV)
    dx = torch.full_like(x, float("nan"))
    capi.avgpool2_bwd(dy, dx, src_off=16)
    xr = x.clone().requires_grad_(True)
    torch.nn.functional.avg_pool2d(xr.permute(0, 3, 1, 2), 2, 2).backward(dy[..., 16:48].permute(0, 3, 1, 2))
    assert torch.allclose(dx, xr.grad, rtol=1e-6, atol=1e-7)


ENGINE_CASES = [("resnet", 2, 64), ("resnet", [1, 2], 64), ("resnet", 3, 64), ("vgg", 3, 32), ("vgg", [2, 3], 32),
                ("alexnet", 3, 64), ("alexnet", [2, 3], 64), ("squeezenet", 2, 64), ("squeezenet", [2, 3], 64),
                ("squeezenet", 4, 64), ("densenet", 1, 64), ("densenet", [2, 3], 64), ("densenet161", 2, 64)]


@pytest.mark.parametrize("name,depth,side", ENGINE_CASES)
@pytest.mark.parametrize("mode", ["simt", "tc"])
def test_native_engine_matches_autograd(name, depth, side, mode):
    """Truncated forward features and dcost/dimage of the native engine vs torch autograd in float64 on the
    CPU, for every family and both hook styles (scalar depth / list of depths)."""
    if mode == "tc" and not hasattr(capi, "conv_tc_supported"):
        pytest.skip("tensor-core path not built yet")
    g = torch.Generator().manual_seed(7)
    img = torch.randn(3, 3, side, side, generator=g)
    model = backbones.get_model(name)
    if backbones.family_of(name) == "densenet":
        from i2v_b200.engine_densenet import DenseNetEngine
        eng = DenseNetEngine(model, name, depth, tf32x3=True, use_tensor_cores=(mode == "tc"))
    else:
        eng = NativeEngine(model, name, depth, tf32x3=True, use_tensor_cores=(mode == "tc"))
    feats = eng.features(img.to(DEV), need_grad=True)
    # float64 reference on the CPU through the reference's own hook mechanism.  ReLU is not differentiable at 0:
    # a pre-activation within rounding distance of 0 may be decided differently by two correct float32 forward
    # passes, and one decision is an O(1) change of the gradient over that element's receptive field.  So the
    # float64 backward is evaluated with the ENGINE's ReLU decisions (oracle.loops.forced_relu_masks) and compared
    # tightly; the number of decisions that differ from float64's own is bounded separately.
    from oracle import loops as OL
    ref_model = backbones.seeded_random_init(backbones.arch_of(name), 0).double()
    backbones.freeze_for_attack(ref_model)
    acts = []
    hs = [t.register_forward_hook(lambda m, i, o: acts.append(o)) for t in backbones.find_target_layers(ref_model, name, depth)]
    xi = img.double().requires_grad_(True)
    with OL.forced_relu_masks(ref_model, eng.relu_masks()) as fm:
        ref_model(xi)
    for h in hs:
        h.remove()
    assert fm.i == len(fm.masks)
    assert fm.flips <= max(2, 2e-5 * fm.total), ("ReLU decisions differing from float64", fm.flips, fm.total)
    assert len(acts) == len(feats)
    ups = []
    for a, f in zip(acts, feats):
        got = f.permute(0, 3, 1, 2).cpu().double()
        assert got.shape == a.shape
        err = (got - a.detach()).abs().max() / a.detach().abs().max()
        # 3xTF32: exact operands, but the tensor core accumulates with truncation (K/8 truncations of the main
        # accumulator): a systematic shrink of ~K * 2^-27 per layer that adds up over the depth of the stack
        assert err <= (5e-5 if mode == "tc" else 5e-6), ("feature", err)
        up = torch.randn(a.shape, generator=g, dtype=torch.float64)
        if eng.relu_masked_grads:
            up = up * (got > 0)                                                     # pre-activation gradient
        ups.append(up)
    (gref,) = torch.autograd.grad(acts, xi, ups)
    grads = [u.permute(0, 2, 3, 1).contiguous().float().to(DEV) for u in ups]
    gimg = eng.input_grad(grads).cpu().double()
    diff = (gimg - gref).abs()
    rel_l2 = diff.pow(2).sum().sqrt() / gref.pow(2).sum().sqrt()
    rel_max = diff.max() / gref.abs().max()
    print("engine parity %s depth=%s mode=%s: relL2 %.2e relmax %.2e relu flips %d / %d" % (name, depth, mode, rel_l2, rel_max, fm.flips, fm.total))
    assert rel_l2 <= (5e-5 if mode == "tc" else 5e-6), ("input gradient L2", rel_l2)
    assert rel_max <= (2e-4 if mode == "tc" else 2e-5), ("input gradient max", rel_max)


# ------------------------------------------------------------------------------------------------
# tensor-core path (tcgen05 / TMEM / TMA), per layer
# ------------------------------------------------------------------------------------------------
TC_SHAPES = [
    #  N  H   W   Cin  Cout k s p
    (2, 56, 56, 64, 64, 1, 1, 0),      # 2-D tiled TMA, 2 k-blocks, BN=64
    (2, 56, 56, 64, 256, 1, 1, 0),     # layer1 conv3
    (2, 56, 56, 256, 64, 1, 1, 0),     # layer1 conv1 (8 k-blocks: pipeline wraps)
    (2, 56, 56, 64, 64, 3, 1, 1),      # im2col TMA, padding rows/cols
    (3, 28, 28, 128, 128, 3, 1, 1),    # tile spans image boundaries (784 px / image)
    (2, 56, 56, 128, 128, 3, 2, 1),    # strided im2col (forward only on TC)
    (2, 56, 56, 256, 512, 1, 2, 0),    # strided 1x1 (forward only on TC)
    (1, 28, 28, 128, 512, 1, 1, 0),    # M = 784: last tile ragged (784 = 6*128 + 16)
    (2, 14, 14, 256, 256, 3, 1, 1),    # layer3-like, 196 px / image
    (1, 13, 13, 64, 192, 3, 1, 1),     # odd spatial size, Cout = 3*64
    (1, 15, 15, 64, 192, 5, 1, 2),     # AlexNet conv2: 25 taps
    (3, 32, 32, 64, 64, 3, 1, 1),      # VGG-on-32x32 shapes (tests): tiles cover several rows / several images
    (3, 16, 16, 128, 128, 3, 1, 1),
    (3, 8, 8, 256, 256, 3, 1, 1),
    (3, 4, 4, 512, 512, 3, 1, 1),      # 48 pixels in total (< one tile), K = 4608: 144 pipeline iterations
    (11, 4, 4, 64, 64, 3, 1, 1),       # one tile spans 8 whole images
]


def _tc_operands(w, scale):
    from i2v_b200.engine_native import _split_tf32
    ws = w * scale.view(-1, 1, 1, 1)
    cout, cin, k, _ = w.shape
    tf = ws.permute(0, 2, 3, 1).reshape(cout, k * k * cin).contiguous().to(DEV)
    td = ws.flip(2, 3).permute(1, 2, 3, 0).reshape(cin, k * k * cout).contiguous().to(DEV)
    return ws, _split_tf32(tf), _split_tf32(td)


@pytest.fixture
def pair_kernel():
    """Every 3xTF32 / TMA-epilogue launch of the test runs on the CTA-pair kernel (tcgen05.mma.cta_group::2)."""
    capi.conv_tc_set_pair_minkit(1)
    yield
    capi.conv_tc_set_pair_minkit(int(os.environ.get("I2V_TC_PAIR", "-1")))


@pytest.mark.parametrize("shape", TC_SHAPES)
def test_conv_tc_pair_kernel_is_bit_identical(shape, pair_kernel):
    """The CTA-pair kernel computes every output element with the same operands in the same k-order as the single-CTA
    dual-issuer kernel, so forward (+bias, residual, ReLU, activity bits) and data gradient (+addend, bit mask) must be
    BIT-identical to it — odd tile counts (a phantom second m-tile), ragged last tiles and multi-image tiles included."""
    N, H, W, Cin, Cout, k, s, p = shape
    g = torch.Generator().manual_seed(13)
    x = torch.randn(N, H, W, Cin, generator=g).to(DEV)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    scale = torch.rand(Cout, generator=g) + 0.5
    bias = (torch.randn(Cout, generator=g) * 0.1).to(DEV)
    P, Q = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    res = torch.randn(N, P, Q, Cout, generator=g).to(DEV)
    d = capi.ConvDesc(N, H, W, Cin, Cout, k, k, s, p, P, Q)
    ws, (fh, fl, fr), (dh, dl, dr) = _tc_operands(w, scale)
    out = {}
    for pair in (0, 1):
        capi.conv_tc_set_pair_minkit(pair)
        y = torch.full((N, P, Q, Cout), float("nan"), device=DEV)
        bits = torch.zeros(Cout // 32, N * P * Q, device=DEV, dtype=torch.int32)
        capi.conv_tc(d, 0, x, fh, fl, bias, res, None, y, relu=True, mask_bits=bits)
        y2 = torch.full((N, P, Q, Cout), float("nan"), device=DEV)
        capi.conv_tc(d, 0, x, fh, fl, bias, None, None, y2, relu=False)
        dx = None
        if s == 1 and capi.conv_tc_supported(d, 1):
            dy = torch.randn(N, P, Q, Cout, generator=torch.Generator().manual_seed(5)).to(DEV)
            add = torch.randn(N, H, W, Cin, generator=torch.Generator().manual_seed(6)).to(DEV)
            mb = torch.randint(-2 ** 31, 2 ** 31 - 1, (Cin // 32, N * H * W), generator=torch.Generator().manual_seed(7),
                               dtype=torch.int64).to(torch.int32).to(DEV)
            dx = torch.full((N, H, W, Cin), float("nan"), device=DEV)
            capi.conv_tc(d, 1, dy, dh, dl, None, add, None, dx, mask_bits=mb)
        out[pair] = (y, bits, y2, dx)
    assert torch.isfinite(out[1][0]).all() and torch.isfinite(out[1][2]).all()
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1]) and torch.equal(out[0][2], out[1][2])
    if out[0][3] is not None:
        assert torch.isfinite(out[1][3]).all() and torch.equal(out[0][3], out[1][3])


HALO_SHAPES = [
    #  N  H   W   Cin  Cout
    (2, 56, 56, 64, 64),       # the layer it runs for: two output rows per tile (58-pixel padded rows), 28 tiles per image
    (2, 28, 28, 128, 128),     # four rows per tile, BN = 128: one accumulator stage, two weight stages
    (3, 14, 14, 256, 64),      # eight rows per tile, H % R != 0: a short last tile per image through the second store map
    (2, 7, 9, 64, 128),        # H < R: one short tile per image
    (1, 57, 31, 64, 192),      # odd sizes, three n-tiles of 64
    (2, 20, 126, 64, 64),      # the widest row that fits a tile (W + 2 = 128): one row per tile
    (5, 3, 3, 512, 64),        # 16 channel blocks per tile
]


@pytest.mark.parametrize("shape", HALO_SHAPES)
@pytest.mark.parametrize("mode", [1, 2])
def test_conv3x3_halo_kernel(shape, mode):
    """The patch-once 3x3 kernel (conv3x3_halo_kernel: padded-raster tiles, the nine taps as row-shifted descriptor views of
    one TMA box) against the im2col-mode kernel on the same operands — same 3xTF32 products, a different accumulation order
    (taps inside a channel block instead of channel blocks inside a tap), so equal to within the accumulation bound — and both
    against float64: forward (+bias, ReLU, activity bits out) and data gradient (+ReLU-backward bit mask).  mode 2 forces
    64-channel tiles (two accumulator stages) where mode 1 would pick 128."""
    N, H, W, Cin, Cout = shape
    g = torch.Generator().manual_seed(17)
    x = torch.randn(N, H, W, Cin, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5
    scale = torch.rand(Cout, generator=g) + 0.5
    bias = torch.randn(Cout, generator=g) * 0.1
    d = capi.ConvDesc(N, H, W, Cin, Cout, 3, 3, 1, 1, H, W)
    ws, (fh, fl, fr), (dh, dl, dr) = _tc_operands(w, scale)
    dy = torch.randn(N, H, W, Cout, generator=g)
    mb = torch.randint(-2 ** 31, 2 ** 31 - 1, (Cin // 32, N * H * W), generator=g, dtype=torch.int64).to(torch.int32)
    xd, dyd, mbd, bd = x.to(DEV), dy.to(DEV), mb.to(DEV), bias.to(DEV)
    out = {}
    try:
        for m in (0, mode):
            capi.conv_tc_set_halo_mode(m)
            y = torch.full((N, H, W, Cout), float("nan"), device=DEV)
            bits = torch.zeros(Cout // 32, N * H * W, device=DEV, dtype=torch.int32)
            capi.conv_tc(d, 0, xd, fh, fl, bd, None, None, y, relu=True, mask_bits=bits)
            dx = None
            if capi.conv_tc_supported(d, 1):
                dx = torch.full((N, H, W, Cin), float("nan"), device=DEV)
                capi.conv_tc(d, 1, dyd, dh, dl, None, None, None, dx, mask_bits=mbd)
            out[m] = (y, bits, dx)
    finally:
        capi.conv_tc_set_halo_mode(int(os.environ.get("I2V_TC_HALO", "-1")))
    K = 9 * Cin
    y64 = torch.relu(torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), ws.double(), bias.double(), padding=1))
    y64 = y64.permute(0, 2, 3, 1)
    tol = (1e-5 + K * 2.0 ** -24) * float(y64.abs().max())
    for m in out:
        y = out[m][0].cpu().double()
        assert torch.isfinite(y).all()
        assert float((y - y64).abs().max()) <= tol, (m, float((y - y64).abs().max()), tol)
    assert float((out[0][0] - out[mode][0]).abs().max()) <= tol
    # activity bits: the same function of the kernel's own output in both kernels
    for m in out:
        y, bits = out[m][0], out[m][1]
        want = (y.reshape(-1, Cout // 32, 32) > 0).to(torch.int64)
        words = (want << torch.arange(32, device=DEV).view(1, 1, 32)).sum(-1)                       # [M, Cout/32]
        words = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32).t().contiguous()
        assert torch.equal(words, bits), m
    if out[0][2] is not None:
        keep = ((mb.view(Cin // 32, -1, 1) >> torch.arange(32).view(1, 1, 32)) & 1).permute(1, 0, 2).reshape(N, H, W, Cin)
        wd = ws.double()
        dx64 = torch.nn.functional.conv_transpose2d(dy.permute(0, 3, 1, 2).double(), wd, padding=1).permute(0, 2, 3, 1) * keep
        told = (1e-5 + 9 * Cout * 2.0 ** -24) * float(dx64.abs().max())
        for m in out:
            dx = out[m][2].cpu().double()
            assert torch.isfinite(dx).all()
            assert float((dx - dx64).abs().max()) <= told, (m, float((dx - dx64).abs().max()), told)


@pytest.mark.parametrize("shape", TC_SHAPES)
@pytest.mark.parametrize("x3", [True, False])
def test_conv_tc_fwd_and_dgrad(shape, x3):
    """Error model: plain TF32 truncates both operands (~1e-3); 3xTF32 recovers the operands exactly and is left
    with the tensor core's truncating FP32 accumulation, a bias of about (K/8)/2 ulp — bounded here by
    1e-5 + K * 2^-24 relative to the largest output."""
    N, H, W, Cin, Cout, k, s, p = shape
    g = torch.Generator().manual_seed(11)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    scale = torch.rand(Cout, generator=g) + 0.5
    shift = torch.randn(Cout, generator=g) * 0.1
    P, Q = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    res = torch.randn(N, Cout, P, Q, generator=g)
    d = capi.ConvDesc(N, H, W, Cin, Cout, k, k, s, p, P, Q)
    assert capi.conv_tc_supported(d, 0)
    ws, (fh, fl, fr), (dh, dl, dr) = _tc_operands(w, scale)
    xd = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    resd = res.permute(0, 2, 3, 1).contiguous().to(DEV)
    bias = shift.to(DEV)
    tol = (1e-5 + k * k * Cin * 2.0 ** -24) if x3 else 4e-3
    for use_res, relu in ((False, False), (True, True)):
        y = torch.full((N, P, Q, Cout), float("nan"), device=DEV)
        capi.conv_tc(d, 0, xd, fh if x3 else fr, fl if x3 else None, bias, resd if use_res else None, None, y, relu=relu)
        ref64 = _ref_conv(x, w, scale, shift, s, p, res if use_res else None, relu, torch.float64)
        got = y.permute(0, 3, 1, 2).cpu().double()
        assert torch.isfinite(got).all()
        err = (got - ref64).abs().max() / ref64.abs().max()
        assert err <= tol, ("fwd", err, tol)
    if s != 1:      # strided data gradients go through the stride-parity classes (test_conv_tc_strided_dgrad_classes)
        return
    assert capi.conv_tc_supported(d, 1)
    dy = torch.randn(N, Cout, P, Q, generator=g)
    addend = torch.randn(N, Cin, H, W, generator=g)
    act = torch.randn(N, Cin, H, W, generator=g)
    dyd = dy.permute(0, 2, 3, 1).contiguous().to(DEV)
    lay = lambda t: t.permute(0, 2, 3, 1).contiguous().to(DEV)
    tol = (1e-5 + k * k * Cout * 2.0 ** -24) if x3 else 4e-3
    for use_add, use_mask in ((False, False), (True, True)):
        dx = torch.full((N, H, W, Cin), float("nan"), device=DEV)
        capi.conv_tc(d, 1, dyd, dh if x3 else dr, dl if x3 else None, None, lay(addend) if use_add else None,
                     lay(act) if use_mask else None, dx)
        ref64 = torch.nn.grad.conv2d_input((N, Cin, H, W), ws.double(), dy.double(), s, p)
        if use_add:
            ref64 = ref64 + addend.double()
        if use_mask:
            ref64 = ref64 * (act > 0).double()
        got = dx.permute(0, 3, 1, 2).cpu().double()
        assert torch.isfinite(got).all()
        err = (got - ref64).abs().max() / ref64.abs().max()
        assert err <= tol, ("dgrad", err, tol)


def _pack_bits(t_nhwc):
    """[N,H,W,C] bool -> int32 [C/32, M] words, bit j of word (w, m) <-> channel 32w + j (include/i2v_b200.h)."""
    C = t_nhwc.shape[-1]
    b = t_nhwc.reshape(-1, C // 32, 32).to(torch.int64)
    words = (b << torch.arange(32, device=b.device, dtype=torch.int64)).sum(-1)        # [M, C/32]
    words = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)
    return words.t().contiguous()


@pytest.mark.parametrize("shape", [(2, 56, 56, 64, 256, 1, 1, 0), (3, 28, 28, 128, 128, 3, 1, 1), (1, 13, 13, 64, 192, 3, 1, 1),
                                   (5, 28, 28, 512, 128, 1, 1, 0), (11, 4, 4, 64, 64, 3, 1, 1), (2, 56, 56, 64, 64, 1, 1, 0)])
@pytest.mark.parametrize("x3", [True, False])
def test_conv_tc_bit_masks(shape, x3):
    """TMA epilogue: the forward's activity bits equal 1[y > 0] exactly, and a data gradient masked by bits (with an
    in-place addend) makes the same mask decisions as the same launch with the f32 mask source (register epilogue)."""
    N, H, W, Cin, Cout, k, s, p = shape
    g = torch.Generator().manual_seed(5)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    scale = torch.rand(Cout, generator=g) + 0.5
    shift = torch.randn(Cout, generator=g) * 0.1
    P, Q = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    d = capi.ConvDesc(N, H, W, Cin, Cout, k, k, s, p, P, Q)
    ws, (fh, fl, fr), (dh, dl, dr) = _tc_operands(w, scale)
    lay = lambda t: t.permute(0, 2, 3, 1).contiguous().to(DEV)
    xd, bias = lay(x), shift.to(DEV)
    res = lay(torch.randn(N, Cout, P, Q, generator=g))
    M = N * P * Q
    for use_res in (False, True):
        y = torch.full((N, P, Q, Cout), float("nan"), device=DEV)
        bits = torch.zeros(Cout // 32, M, dtype=torch.int32, device=DEV)
        capi.conv_tc(d, 0, xd, fh if x3 else fr, fl if x3 else None, bias, res if use_res else None, None, y, relu=True,
                     mask_bits=bits)
        y_plain = torch.full_like(y, float("nan"))
        capi.conv_tc(d, 0, xd, fh if x3 else fr, fl if x3 else None, bias, res if use_res else None, None, y_plain, relu=True)
        assert torch.isfinite(y).all()
        assert torch.equal(y, y_plain)
        assert torch.equal(bits, _pack_bits(y > 0))
    # data gradient: mask by bits == mask by the f32 tensor, also with dx += (in place)
    dy = lay(torch.randn(N, Cout, P, Q, generator=g))
    act = lay(torch.randn(N, Cin, H, W, generator=g))
    add = lay(torch.randn(N, Cin, H, W, generator=g))
    abits = _pack_bits(act > 0)
    for inplace in (False, True):
        want = add.clone() if inplace else torch.full_like(add, float("nan"))
        capi.conv_tc(d, 1, dy, dh if x3 else dr, dl if x3 else None, None, want if inplace else add, act, want)
        got = add.clone() if inplace else torch.full_like(add, float("nan"))
        capi.conv_tc(d, 1, dy, dh if x3 else dr, dl if x3 else None, None, got if inplace else add, None, got, mask_bits=abits)
        assert torch.isfinite(got).all()
        # same mask decisions exactly; values to rounding (the dual-issuer kernel keeps the two cross terms in separate
        # accumulators, the register-epilogue kernel in one)
        assert torch.equal(got == 0, want == 0)
        assert (got - want).abs().max() <= 2e-6 * want.abs().max()


@pytest.mark.parametrize("shape", [(2, 56, 56, 64, 1, 64, 256), (3, 56, 56, 256, 2, 128, 512), (1, 15, 13, 64, 2, 32, 64),
                                   (5, 9, 9, 32, 1, 96, 128), (2, 28, 28, 512, 2, 256, 1024)])
@pytest.mark.parametrize("x3", [True, False])
def test_conv_tc_dual_source(shape, x3):
    """A bottleneck's downsample branch (1x1 / stride s over x) and its last 1x1 (over t) as ONE GEMM over the concatenated
    K (second A source through its own tensor map): against float64 with the error model of the single-source launches,
    against the two-launch path (ds, then conv3 + residual) to rounding, activity bits = 1[y > 0] exactly; ResNet-50's
    layer1.0 / layer2.0 / layer3.0 shapes, odd sizes under stride 2, ragged last tiles."""
    N, H, W, C1, s, C2, Cout = shape
    g = torch.Generator().manual_seed(21)
    P, Q = (H - 1) // s + 1, (W - 1) // s + 1
    x = torch.randn(N, C1, H, W, generator=g)
    t = torch.randn(N, C2, P, Q, generator=g)
    w1 = torch.randn(Cout, C1, 1, 1, generator=g) / C1 ** 0.5
    w2 = torch.randn(Cout, C2, 1, 1, generator=g) / C2 ** 0.5
    sc1, sc2 = torch.rand(Cout, generator=g) + 0.5, torch.rand(Cout, generator=g) + 0.5
    b1, b2 = torch.randn(Cout, generator=g) * 0.1, torch.randn(Cout, generator=g) * 0.1
    d1 = capi.ConvDesc(N, H, W, C1, Cout, 1, 1, s, 0, P, Q)
    d2 = capi.ConvDesc(N, P, Q, C2, Cout, 1, 1, 1, 0, P, Q)
    _, (h1, l1, r1), _ = _tc_operands(w1, sc1)
    _, (h2, l2, r2), _ = _tc_operands(w2, sc2)
    hi, lo, rn = (torch.cat([a, b], 1).contiguous() for a, b in ((h1, h2), (l1, l2), (r1, r2)))
    lay = lambda v: v.permute(0, 2, 3, 1).contiguous().to(DEV)
    xd, td = lay(x), lay(t)
    bias = (b1 + b2).to(DEV)
    M = N * P * Q
    y = torch.full((N, P, Q, Cout), float("nan"), device=DEV)
    bits = torch.zeros(Cout // 32, M, dtype=torch.int32, device=DEV)
    capi.conv_tc_dual(d1, xd, td, hi if x3 else rn, lo if x3 else None, bias, y, relu=True, mask_bits=bits)
    assert torch.isfinite(y).all()
    assert torch.equal(bits, _pack_bits(y > 0))
    y2 = torch.full_like(y, float("nan"))
    capi.conv_tc_dual(d1, xd, td, hi if x3 else rn, lo if x3 else None, bias, y2, relu=True)
    assert torch.equal(y, y2)
    ref64 = torch.relu(_ref_conv(x, w1, sc1, b1, s, 0, None, False, torch.float64) + _ref_conv(t, w2, sc2, b2, 1, 0, None, False, torch.float64))
    got = y.permute(0, 3, 1, 2).cpu().double()
    err = (got - ref64).abs().max() / ref64.abs().max()
    assert err <= ((1e-5 + (C1 + C2) * 2.0 ** -24) if x3 else 4e-3), err
    # the two-launch path it replaces
    sc = torch.full((N, P, Q, Cout), float("nan"), device=DEV)
    capi.conv_tc(d1, 0, xd, h1 if x3 else r1, l1 if x3 else None, b1.to(DEV), None, None, sc)
    y3 = torch.full_like(y, float("nan"))
    capi.conv_tc(d2, 0, td, h2 if x3 else r2, l2 if x3 else None, b2.to(DEV), sc, None, y3, relu=True)
    assert (y - y3).abs().max() <= (4e-6 if x3 else 4e-3) * y3.abs().max()


def test_engine_fused_downsample_matches_two_launches(monkeypatch):
    """NativeEngine with the downsample branches absorbed into the blocks' last convolutions (default) against the same
    engine with $I2V_FUSE_DS=0: features and input gradient agree to FP32 rounding; the absorbed buffers are not allocated."""
    from i2v_b200 import backbones
    from i2v_b200.engine_native import NativeEngine
    g = torch.Generator().manual_seed(2)
    img = torch.randn(3, 3, 64, 64, generator=g).to(DEV)
    outs = []
    for fuse in ("1", "0"):
        monkeypatch.setenv("I2V_FUSE_DS", fuse)
        model = backbones.get_model("resnet50")
        eng = NativeEngine(model, "resnet50", 2)
        feats = eng.features(img, need_grad=True)
        plan = eng._last
        assert len(plan["fused_ds"]) == (2 if fuse == "1" else 0)
        assert all(("l%d.0.sc" % li in plan["acts"]) == (fuse == "0") for li in (1, 2))
        f0 = feats[0].clone()
        gin = eng.input_grad([torch.ones_like(f0) / f0.numel()]).clone()
        outs.append((f0, gin))
    (fa, ga), (fb, gb) = outs
    assert (fa - fb).abs().max() <= 1e-5 * fb.abs().max()
    assert torch.equal(fa > 0, fb > 0) or ((fa > 0) != (fb > 0)).float().mean() < 1e-4
    assert (ga - gb).abs().max() <= 1e-3 * gb.abs().max()


@pytest.mark.parametrize("H,W,k,s,p", [(224, 224, 7, 2, 3), (64, 64, 7, 2, 3), (64, 64, 11, 4, 2), (32, 32, 3, 1, 1),
                                       (64, 64, 3, 2, 0), (37, 53, 7, 2, 3), (31, 45, 3, 2, 0), (50, 70, 11, 4, 2)])
def test_stem_fwd_and_dgrad(H, W, k, s, p):
    """Dedicated first-layer kernels (Cin = 3 -> Cout = 64) for every attacked family, incl. ragged sizes."""
    g = torch.Generator().manual_seed(5)
    N, Cin, Cout = 2, 3, 64
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    scale = torch.rand(Cout, generator=g) + 0.5
    shift = torch.randn(Cout, generator=g) * 0.1
    P, Q = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    d = capi.ConvDesc(N, H, W, Cin, Cout, k, k, s, p, P, Q)
    assert capi.conv_stem_supported(d)
    ws = w * scale.view(-1, 1, 1, 1)
    wf = ws.permute(1, 2, 3, 0).reshape(Cin * k * k, Cout).contiguous().to(DEV)
    wd = ws.permute(2, 3, 1, 0).reshape(k * k * Cin, Cout).contiguous().to(DEV)
    y = torch.full((N, P, Q, Cout), float("nan"), device=DEV)
    capi.conv_stem_fwd(d, x.to(DEV), wf, shift.to(DEV), y, relu=True)
    ref = _ref_conv(x, w, scale, shift, s, p, None, True, torch.float64)
    got = y.permute(0, 3, 1, 2).cpu().double()
    assert torch.isfinite(got).all()
    assert (got - ref).abs().max() / ref.abs().max() <= 2e-6
    dy = torch.randn(N, Cout, P, Q, generator=g)
    dx = torch.full((N, Cin, H, W), float("nan"), device=DEV)
    capi.conv_stem_dgrad(d, dy.permute(0, 2, 3, 1).contiguous().to(DEV), wd, dx)
    ref = torch.nn.grad.conv2d_input((N, Cin, H, W), ws.double(), dy.double(), s, p)
    assert torch.isfinite(dx).all()
    assert (dx.cpu().double() - ref).abs().max() / ref.abs().max() <= 2e-6


@pytest.mark.parametrize("shape", [(2, 56, 56, 128, 128, 3, 2, 1), (2, 56, 56, 256, 512, 1, 2, 0), (3, 28, 28, 256, 256, 3, 2, 1),
                                   (2, 27, 27, 64, 64, 3, 2, 1), (1, 55, 41, 64, 128, 3, 2, 0), (2, 28, 28, 512, 1024, 1, 2, 0)])
@pytest.mark.parametrize("x3", [True, False])
@pytest.mark.parametrize("mask", ["f32", "bits", "none"])
def test_conv_tc_strided_dgrad_classes(shape, x3, mask):
    """Strided data gradient as stride^2 dense tensor-core problems scattered into dx (even/odd sizes, padding 0/1).
    mask = f32: the activation as mask source (register epilogue, per-thread scatter); bits / none: the TMA epilogue — rows
    scatter through an im2col-mode TMA store over the class's strided view of dx, the in-place addend comes in by an
    im2col-mode load, the mask is one word per row from the [Cin/32][N*H*W] bit planes."""
    from i2v_b200.engine_native import _class_weights
    N, H, W, Cin, Cout, k, s, p = shape
    g = torch.Generator().manual_seed(13)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    P, Q = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    d = capi.ConvDesc(N, H, W, Cin, Cout, k, k, s, p, P, Q)
    assert capi.conv_tc_supported(d, 1)
    cls = _class_weights(w.to(DEV), s, p)
    dy = torch.randn(N, Cout, P, Q, generator=g)
    addend = torch.randn(N, Cin, H, W, generator=g)
    act = torch.randn(N, Cin, H, W, generator=g)
    lay = lambda t: t.permute(0, 2, 3, 1).contiguous().to(DEV)
    ref = torch.nn.grad.conv2d_input((N, Cin, H, W), w.double(), dy.double(), s, p)
    tol = (1e-5 + k * k * Cout * 2.0 ** -24) if x3 else 4e-3
    abits = _pack_bits(lay(act) > 0) if mask == "bits" else None
    for use_add in (False, True):
        dx = lay(addend) if use_add else torch.zeros(N, H, W, Cin, device=DEV)
        for (ph, pw), wt in cls.items():
            if wt is None:
                continue
            hi, lo, rna = wt
            capi.conv_tc_dgrad_class(d, ph, pw, lay(dy), hi if x3 else rna, lo if x3 else None, dx if use_add else None,
                                     lay(act) if mask == "f32" else None, dx, mask_bits=abits)
        want = (ref + addend.double()) if use_add else ref
        if mask == "none":
            act = torch.ones_like(act)
        # classes with taps are masked by the kernel; classes without taps keep the addend (already masked in real use)
        got = dx.permute(0, 3, 1, 2).cpu().double()
        has = torch.zeros(H, W, dtype=torch.bool)
        for (ph, pw), wt in cls.items():
            if wt is not None:
                has[ph::s, pw::s] = True
        want = torch.where(has, want * (act > 0).double(), want)
        err = (got - want).abs().max() / want.abs().max()
        assert err <= tol, (err, tol)


@pytest.mark.parametrize("H,W,k,s,p,n", [(224, 224, 7, 2, 3, 3), (64, 64, 7, 2, 3, 2), (37, 53, 7, 2, 3, 1), (62, 30, 5, 2, 2, 2),
                                         (112, 112, 7, 2, 3, 5)])
def test_stem_fwd_direct(H, W, k, s, p, n):
    """EXPERIMENTAL first-layer forward without the patch matrix: padded NHWC4 copy + 4-D tiled TMA boxes of 16 x 8 output
    pixels whose q stride overlaps the 8-pixel window (ragged P / Q exercise the TMA zero fill and store clipping)."""
    from i2v_b200.engine_native import _split_tf32
    g = torch.Generator().manual_seed(4)
    Cout = 64
    x = torch.randn(n, 3, H, W, generator=g)
    w = torch.randn(Cout, 3, k, k, generator=g) / (3 * k * k) ** 0.5
    shift = torch.randn(Cout, generator=g) * 0.1
    P, Q = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    d = capi.ConvDesc(n, H, W, 3, Cout, k, k, s, p, P, Q)
    assert capi.conv_stem_fwd_direct_supported(d)
    wr = torch.zeros(Cout, k, 8, 4)
    wr[:, :, :k, :3] = w.permute(0, 2, 3, 1)
    hi, lo, _ = _split_tf32(wr.reshape(Cout, k * 32).contiguous().to(DEV))
    xp = torch.full((capi.stem_fwd_direct_scratch_floats(d),), float("nan"), device=DEV)
    y = torch.full((n, P, Q, Cout), float("nan"), device=DEV)
    capi.conv_stem_fwd_direct(d, x.to(DEV), hi, lo, shift.to(DEV), xp, y, relu=True)
    ref64 = _ref_conv(x, w, torch.ones(Cout), shift, s, p, None, True, torch.float64)
    got = y.permute(0, 3, 1, 2).cpu().double()
    assert torch.isfinite(got).all()
    err = (got - ref64).abs().max() / ref64.abs().max()
    assert err <= (1e-5 + 3 * k * k * 2.0 ** -24), err


@pytest.mark.parametrize("H,W,k,s,p,n", [(224, 224, 7, 2, 3, 2), (64, 64, 7, 2, 3, 3), (64, 64, 11, 4, 2, 1), (32, 32, 3, 1, 1, 2),
                                         (63, 63, 3, 2, 0, 4)])
@pytest.mark.parametrize("x3", [True, False])
def test_stem_dgrad_tc(H, W, k, s, p, n, x3):
    """First-layer data gradient as a tcgen05 GEMM over the output channels + col2im, against float64 autograd
    (same error model as test_conv_tc_fwd_and_dgrad with K = Cout = 64, plus <= ceil(k/s)^2 f32 additions)."""
    from i2v_b200.engine_native import _split_tf32
    g = torch.Generator().manual_seed(3)
    Cout = 64
    w = torch.randn(Cout, 3, k, k, generator=g) / (3 * k * k) ** 0.5
    P, Q = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    if (n * P * Q) % 4:
        pytest.skip("N*P*Q % 4 != 0: the engine uses the CUDA-core stem kernel")
    dy = torch.randn(n, Cout, P, Q, generator=g)
    d = capi.ConvDesc(n, H, W, 3, Cout, k, k, s, p, P, Q)
    rows = 3 * k * k
    nz = capi.stem_dgrad_tc_rows(rows)
    wz = torch.cat([w.permute(1, 2, 3, 0).reshape(rows, Cout), torch.zeros(nz - rows, Cout)], 0).contiguous().to(DEV)
    hi, lo, rna = _split_tf32(wz)
    z = torch.full((capi.stem_dgrad_tc_scratch_floats(d),), float("nan"), device=DEV)
    dx = torch.full((n, 3, H, W), float("nan"), device=DEV)
    capi.conv_stem_dgrad_tc(d, dy.permute(0, 2, 3, 1).contiguous().to(DEV), hi if x3 else rna, lo if x3 else None, z, dx)
    ref64 = torch.nn.grad.conv2d_input((n, 3, H, W), w.double(), dy.double(), s, p)
    got = dx.cpu().double()
    assert torch.isfinite(got).all()
    err = (got - ref64).abs().max() / ref64.abs().max()
    assert err <= ((2e-5 + Cout * 2.0 ** -24) if x3 else 4e-3), err


@pytest.mark.parametrize("H,W,n", [(224, 224, 5), (64, 64, 3), (32, 32, 2), (16, 16, 3), (62, 60, 2), (224, 200, 1), (8, 8, 1)])
def test_stem_fwd_rows(H, W, n):
    """First-layer forward without the patch matrix (one output row per tile, patch tile assembled on chip from TMA-staged
    input rows): against the float64 convolution with bias + ReLU, same error model as test_stem_fwd_tc; even sizes, widths
    whose last tile lanes are junk, tiny images; twice in a row (deterministic, NaN-prefilled output)."""
    from i2v_b200.engine_native import _split_tf32
    g = torch.Generator().manual_seed(4)
    Cout, k, s, p = 64, 7, 2, 3
    x = torch.randn(n, 3, H, W, generator=g)
    w = torch.randn(Cout, 3, k, k, generator=g) / (3 * k * k) ** 0.5
    shift = torch.randn(Cout, generator=g) * 0.1
    P, Q = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    d = capi.ConvDesc(n, H, W, 3, Cout, k, k, s, p, P, Q)
    assert capi.conv_stem_fwd_rows_supported(d)
    wk = torch.cat([w.reshape(Cout, 147), torch.zeros(Cout, 13)], 1).contiguous().to(DEV)
    hi, lo, _ = _split_tf32(wk)
    outs = []
    for _ in range(2):
        y = torch.full((n, P, Q, Cout), float("nan"), device=DEV)
        capi.conv_stem_fwd_rows(d, x.to(DEV), hi, lo, shift.to(DEV), y, relu=True)
        outs.append(y)
    assert torch.equal(outs[0], outs[1])
    ref64 = _ref_conv(x, w, torch.ones(Cout), shift, s, p, None, True, torch.float64)
    got = outs[0].permute(0, 3, 1, 2).cpu().double()
    assert torch.isfinite(got).all()
    err = (got - ref64).abs().max() / ref64.abs().max()
    assert err <= (1e-5 + 3 * k * k * 2.0 ** -24), err


@pytest.mark.parametrize("H,W,n", [(224, 224, 5), (224, 224, 150), (64, 64, 3), (32, 32, 2), (16, 16, 3), (224, 200, 1), (8, 8, 1),
                                   (112, 112, 33), (40, 244, 2)])
@pytest.mark.parametrize("mark_dead", [True, False])
def test_stem_fwd_pool_fused(H, W, n, mark_dead):
    """First-layer forward with the 3x3 / stride-2 / pad-1 max pooling fused into its epilogue (strips of pooled rows per CTA,
    running window maxima in registers, first maximum in window order wins): pooled tensor AND argmax plane bit-identical to
    i2v_conv_stem_fwd_rows_f32 followed by i2v_maxpool_fwd_f32 — many strips per CTA (150 frames), fewer units than SMs,
    widths with junk tile lanes, one-pooled-row images; a negative bias makes dead windows common."""
    from i2v_b200.engine_native import _split_tf32
    g = torch.Generator().manual_seed(21)
    Cout, k, s, p = 64, 7, 2, 3
    x = torch.randn(n, 3, H, W, generator=g).to(DEV)
    w = torch.randn(Cout, 3, k, k, generator=g) / (3 * k * k) ** 0.5
    shift = (torch.randn(Cout, generator=g) * 0.3 - 0.4).to(DEV)
    P, Q = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    P2, Q2 = (P + 2 - 3) // 2 + 1, (Q + 2 - 3) // 2 + 1
    d = capi.ConvDesc(n, H, W, 3, Cout, k, k, s, p, P, Q)
    assert capi.conv_stem_fwd_pool_supported(d, P2, Q2)
    wk = torch.cat([w.reshape(Cout, 147), torch.zeros(Cout, 13)], 1).contiguous().to(DEV)
    hi, lo, _ = _split_tf32(wk)
    y = torch.empty(n, P, Q, Cout, device=DEV)
    capi.conv_stem_fwd_rows(d, x, hi, lo, shift, y, relu=True)
    want = torch.empty(n, P2, Q2, Cout, device=DEV)
    want_am = torch.empty(n, P2, Q2, Cout, device=DEV, dtype=torch.uint8)
    capi.maxpool_fwd(y, want, want_am, 3, 2, 1, mark_dead=mark_dead)
    for _ in range(2):
        got = torch.full((n, P2, Q2, Cout), float("nan"), device=DEV)
        got_am = torch.full((n, P2, Q2, Cout), 77, device=DEV, dtype=torch.uint8)
        capi.conv_stem_fwd_pool(d, x, hi, lo, shift, got, got_am, relu=True, mark_dead=mark_dead)
        assert torch.equal(got.view(torch.int32), want.view(torch.int32))
        assert torch.equal(got_am, want_am)
    if mark_dead:
        assert int((want_am == 255).sum()) > 0


@pytest.mark.parametrize("H,W,n", [(224, 224, 5), (64, 64, 3), (32, 32, 2), (16, 16, 3), (63, 61, 2), (224, 200, 1), (8, 8, 1)])
def test_stem_dgrad_direct(H, W, n):
    """First-layer data gradient without scratch (one dy row per tile, on-chip col2im in a register window, half-image
    strips): against float64 autograd with the same error model as test_stem_dgrad_tc, on even / odd sizes, images too small
    for two strips, widths where the last tile lanes are junk — and twice in a row (deterministic, NaN-prefilled output)."""
    from i2v_b200.engine_native import _split_tf32
    g = torch.Generator().manual_seed(3)
    Cout, k, s, p = 64, 7, 2, 3
    w = torch.randn(Cout, 3, k, k, generator=g) / (3 * k * k) ** 0.5
    P, Q = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    dy = torch.randn(n, Cout, P, Q, generator=g)
    d = capi.ConvDesc(n, H, W, 3, Cout, k, k, s, p, P, Q)
    assert capi.conv_stem_dgrad_direct_supported(d)
    w_stem = w.permute(1, 2, 3, 0).reshape(147, Cout).contiguous().to(DEV)
    hi, lo, _ = _split_tf32(capi.stem_direct_dgrad_weights(w_stem))
    dyd = dy.permute(0, 2, 3, 1).contiguous().to(DEV)
    outs = []
    for _ in range(2):
        dx = torch.full((n, 3, H, W), float("nan"), device=DEV)
        capi.conv_stem_dgrad_direct(d, dyd, hi, lo, dx)
        outs.append(dx)
    assert torch.equal(outs[0], outs[1])
    ref64 = torch.nn.grad.conv2d_input((n, 3, H, W), w.double(), dy.double(), s, p)
    got = outs[0].cpu().double()
    assert torch.isfinite(got).all()
    err = (got - ref64).abs().max() / ref64.abs().max()
    assert err <= (2e-5 + Cout * 2.0 ** -24), err


@pytest.mark.parametrize("H,W,n", [(224, 224, 5), (112, 112, 3), (64, 64, 3), (32, 32, 2), (16, 16, 3), (63, 61, 2), (60, 58, 2),
                                   (224, 200, 1), (8, 8, 1)])
def test_stem_dgrad_pool_fused(H, W, n):
    """Max-pool backward (3x3 / stride 2 / pad 1, dead windows marked by the forward pass) fused into the first-layer data
    gradient: BIT-identical to i2v_maxpool_bwd_f32 followed by i2v_conv_stem_dgrad_direct_f32 — the pooling's routing is an
    exact selection and the window order is the same — on even / odd stem and pooled sizes, with ties and dead windows
    (the activation is a ReLU output with ~half zeros), twice in a row."""
    from i2v_b200.engine_native import _split_tf32
    g = torch.Generator().manual_seed(11)
    Cout, k, s, p = 64, 7, 2, 3
    w = torch.randn(Cout, 3, k, k, generator=g) / (3 * k * k) ** 0.5
    P, Q = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    P2, Q2 = (P - 1) // 2 + 1, (Q - 1) // 2 + 1
    d = capi.ConvDesc(n, H, W, 3, Cout, k, k, s, p, P, Q)
    if not capi.conv_stem_dgrad_pool_supported(d, P2, Q2):
        pytest.skip("pooled rows do not fit in shared memory at this width")
    act = torch.relu(torch.randn(n, P, Q, Cout, generator=g)).to(DEV)          # the stem's ReLU output, NHWC
    act[:, : P // 2, : Q // 3] = torch.round(act[:, : P // 2, : Q // 3])      # ties inside windows
    pooled = torch.empty(n, P2, Q2, Cout, device=DEV)
    am = torch.empty(n, P2, Q2, Cout, device=DEV, dtype=torch.uint8)
    capi.maxpool_fwd(act, pooled, am, 3, 2, 1, mark_dead=True)
    assert int((am == 255).sum()) > 0 or min(P, Q) > 8
    gp = torch.randn(n, P2, Q2, Cout, generator=g).to(DEV)
    w_stem = w.permute(1, 2, 3, 0).reshape(147, Cout).contiguous().to(DEV)
    hi, lo, _ = _split_tf32(capi.stem_direct_dgrad_weights(w_stem))
    gact = torch.full((n, P, Q, Cout), float("nan"), device=DEV)
    capi.maxpool_bwd(gp, am, None, gact, 3, 2, 1)
    ref = torch.full((n, 3, H, W), float("nan"), device=DEV)
    capi.conv_stem_dgrad_direct(d, gact, hi, lo, ref)
    for _ in range(2):
        dx = torch.full((n, 3, H, W), float("nan"), device=DEV)
        capi.conv_stem_dgrad_pool(d, gp, am, hi, lo, dx)
        assert torch.isfinite(dx).all()
        assert torch.equal(dx, ref), float((dx - ref).abs().max())


@pytest.mark.parametrize("H,W,k,s,p,n", [(224, 224, 7, 2, 3, 7), (64, 64, 7, 2, 3, 3), (64, 64, 11, 4, 2, 1), (32, 32, 3, 1, 1, 2),
                                         (63, 63, 3, 2, 0, 4)])
@pytest.mark.parametrize("x3", [True, False])
def test_stem_fwd_tc(H, W, k, s, p, n, x3):
    """First-layer forward as im2col + tcgen05 GEMM (K = 3*k*k padded to 32), bias + ReLU fused, against float64;
    n = 7 at 224x224 spans two frame groups."""
    from i2v_b200.engine_native import _split_tf32
    g = torch.Generator().manual_seed(4)
    Cout = 64
    x = torch.randn(n, 3, H, W, generator=g)
    w = torch.randn(Cout, 3, k, k, generator=g) / (3 * k * k) ** 0.5
    shift = torch.randn(Cout, generator=g) * 0.1
    P, Q = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    d = capi.ConvDesc(n, H, W, 3, Cout, k, k, s, p, P, Q)
    K = 3 * k * k
    kp = (K + 31) // 32 * 32
    wk = torch.cat([w.reshape(Cout, K), torch.zeros(Cout, kp - K)], 1).contiguous().to(DEV)
    hi, lo, rna = _split_tf32(wk)
    col = torch.full((capi.stem_fwd_tc_scratch_floats(d),), float("nan"), device=DEV)
    y = torch.full((n, P, Q, Cout), float("nan"), device=DEV)
    capi.conv_stem_fwd_tc(d, x.to(DEV), hi if x3 else rna, lo if x3 else None, shift.to(DEV), col, y, relu=True)
    ref64 = _ref_conv(x, w, torch.ones(Cout), shift, s, p, None, True, torch.float64)
    got = y.permute(0, 3, 1, 2).cpu().double()
    assert torch.isfinite(got).all()
    err = (got - ref64).abs().max() / ref64.abs().max()
    assert err <= ((1e-5 + K * 2.0 ** -24) if x3 else 4e-3), err
