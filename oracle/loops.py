"""TEST INFRASTRUCTURE ONLY — full-loop CPU restatements of the reference attack classes.

Every function follows the cited reference lines statement by statement; the per-pixel arithmetic goes
through the C oracle (oracle/i2v_oracle.c via oracle/oracle.py), the image / video model runs on the
CPU under torch autograd exactly as the reference runs it (full forward, hooks on the target layers).
Pinned against the unmodified reference classes by tests/test_oracle_golden.py.

Nothing here imports the product package.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import oracle as O

INIT_MODIFIER = 0.01 / 255   # image_attacks.py:304


# --------------------------------------------------------------------------------------------------
# reference `_find_target_layer` restated (image_attacks.py:260-271; TPAMI_attack.py:176-200)
# --------------------------------------------------------------------------------------------------
def target_layers(model, family, depth):
    is_list = isinstance(depth, (list, tuple))
    ds = list(depth) if is_list else [depth]
    if family == "resnet":
        return [getattr(model, "layer{}".format(d))[-1] for d in ds]
    if family == "alexnet":
        table = {1: 1, 2: 4, 3: 7, 4: 11}
        return [model.features[table[d]] for d in ds]
    if family == "vgg":
        table = {1: 1, 2: 11, 3: 20, 4: 29}
        return [model.features[table[d]] for d in ds]
    if family == "squeezenet":
        table = {1: 3, 2: 6, 3: 9, 4: 12}
        if is_list:
            return [model.features[table[d]] for d in ds]
        return [model.features[table[ds[0]]].expand3x3_activation]
    if family == "densenet":
        # NOT in the reference (its _find_target_layer has no DenseNet branch and would crash): the
        # extension this repo defines — depth d hooks features.denseblock{d} — restated here so that the float64
        # arbiter can check the native DenseNet engine
        return [getattr(model.features, "denseblock{}".format(d)) for d in ds]
    raise ValueError(family)


class HookedModel:
    """model.train() + BN eval (image_attacks.py:253-256) and forward hooks that append the target
    layers' outputs in execution order (273-292)."""

    def __init__(self, model, family, depth):
        self.model = model
        model.train()
        for m in model.modules():
            if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm1d)):
                m.eval()
        self.acts = []
        self.handles = [t.register_forward_hook(lambda mod, inp, out: self.acts.append(out))
                        for t in target_layers(model, family, depth)]

    def run(self, img):
        self.acts = []
        self.model(img)
        acts, self.acts = self.acts, []
        return acts

    def close(self):
        for h in self.handles:
            h.remove()


class _ForcedReLU(torch.autograd.Function):
    """y = x * mask: the linear branch of ReLU selected by a GIVEN activity mask (forward and backward)."""

    @staticmethod
    def forward(ctx, x, mask):
        ctx.save_for_backward(mask)
        return x * mask

    @staticmethod
    def backward(ctx, gy):
        (mask,) = ctx.saved_tensors
        return gy * mask, None


class _ForcedMaxPool(torch.autograd.Function):
    """y[n,c,p,q] = x[n,c].flatten()[idx[n,c,p,q]]: max pooling with GIVEN window winners (forward and backward)."""

    @staticmethod
    def forward(ctx, x, idx):
        ctx.save_for_backward(idx)
        ctx.in_shape = x.shape
        n, c = x.shape[:2]
        return x.reshape(n, c, -1).gather(2, idx.reshape(n, c, -1)).reshape(idx.shape)

    @staticmethod
    def backward(ctx, gy):
        (idx,) = ctx.saved_tensors
        n, c, h, w = ctx.in_shape
        gx = torch.zeros(n, c, h * w, dtype=gy.dtype)
        gx.scatter_add_(2, idx.reshape(n, c, -1), gy.reshape(n, c, -1))
        return gx.reshape(n, c, h, w), None


class forced_relu_masks:
    """Context manager: the first len(masks) ReLU calls of `model` (execution order) use the given activity
    masks instead of their own sign decision; a None entry and later calls (layers behind the hooks) are untouched.

    ReLU is not differentiable at 0: two correct float32 implementations of the forward pass can decide
    1[z > 0] differently for a pre-activation within rounding distance of 0, and that single decision changes
    the gradient by O(1) for every pixel in the element's receptive field.  Forcing the decisions of the
    implementation under test separates "were the decisions the same" (counted in `flips`) from "is the
    gradient arithmetic right given the decisions" (compared tightly).  `flips` = number of forced decisions
    that differ from the model's own, `total` = number of decisions."""

    def __init__(self, model, masks, pools=None):
        """pools: optional list of int64 [N,C,P,Q] flat indices (h*W + w into the pooled plane), one per MaxPool2d call
        in execution order — max pooling has the same discontinuity (two window entries within rounding distance of
        each other), so its winners are forced too; calls beyond the list keep their own decision."""
        self.model, self.masks = model, list(masks)
        self.pools = list(pools) if pools is not None else []
        self.flips = self.total = 0
        self.pool_flips = self.pool_total = 0

    def __enter__(self):
        self.i = 0
        self.saved = []
        outer = self

        def make(mod):
            def fwd(x):
                if outer.i < len(outer.masks) and outer.masks[outer.i] is None:
                    outer.i += 1                                  # a call the implementation under test does not make
                    return F.relu(x)
                if outer.i < len(outer.masks):
                    mk = outer.masks[outer.i].to(x.dtype)
                    outer.i += 1
                    if tuple(mk.shape) != tuple(x.shape):
                        raise ValueError("ReLU call %d: mask %s vs activation %s" % (outer.i - 1, tuple(mk.shape), tuple(x.shape)))
                    outer.flips += int(((x.detach() > 0) != (mk > 0)).sum())
                    outer.total += mk.numel()
                    return _ForcedReLU.apply(x, mk)
                return F.relu(x)
            return fwd
        self.pi = 0

        def make_pool(mod):
            def fwd(x):
                if outer.pi >= len(outer.pools) or outer.pools[outer.pi] is None:
                    outer.pi += 1
                    return F.max_pool2d(x, mod.kernel_size, mod.stride, mod.padding, mod.dilation, mod.ceil_mode)
                idx = outer.pools[outer.pi]
                outer.pi += 1
                _, own = F.max_pool2d(x.detach(), mod.kernel_size, mod.stride, mod.padding, mod.dilation, mod.ceil_mode,
                                      return_indices=True)
                if tuple(own.shape) != tuple(idx.shape):
                    raise ValueError("MaxPool call %d: indices %s vs %s" % (outer.pi - 1, tuple(idx.shape), tuple(own.shape)))
                outer.pool_flips += int((own != idx).sum())
                outer.pool_total += idx.numel()
                return _ForcedMaxPool.apply(x, idx)
            return fwd
        for mod in self.model.modules():
            if isinstance(mod, torch.nn.ReLU):
                self.saved.append(mod)
                mod.forward = make(mod)
            elif isinstance(mod, torch.nn.MaxPool2d) and self.pools:
                self.saved.append(mod)
                mod.forward = make_pool(mod)
        return self

    def __exit__(self, *exc):
        for mod in self.saved:
            del mod.forward
        return False


def _frames(videos):
    b, c, f, h, w = videos.shape
    return videos.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w).contiguous()


def image_guided_loop(hooked, videos, epsilon, steps, step_size, adaptive=False, coeffs=None, momentum=0.0,
                      coef_CE=False, cos_mode="torch", tap=None, weight_grads=False):
    """image_attacks.py:294-364 / 426-496 (adaptive=False) and TPAMI_attack.py:223-320 (adaptive=True).

    weight_grads: True = `cost.backward()` exactly as image_attacks.py:351-352 runs it — the backbone's parameters
              require grad, so every step also computes (and accumulates into .grad, nothing ever zeroes them) all
              weight gradients; this is the COST the reference pays and what the bench's CPU arm times.  False =
              `autograd.grad(cost, true_image)`: same dcost/dtrue_image bit for bit, weight gradients pruned (tests).

    hooked  : list of HookedModel (one per image model, reference order)
    cos_mode: 'torch' — F.cosine_similarity + autograd in f32, the reference's own arithmetic
              'f64'   — the float64 analytic cosine gradient of the C oracle (the accuracy arbiter)
    tap     : optional callable(step, dict) receiving per-step state (g, m, v, mod, true_image, cos)
    Returns (adv [b,3,f,h,w] float32 ndarray, cost [steps] float32, weights [steps, L] or None, coeffs)
    """
    videos = torch.as_tensor(videos, dtype=torch.float32)
    b, c, f, h, w = videos.shape
    N, inner = b * f, h * w
    frames = _frames(videos)                                                   # 300-301
    x = O.denorm(frames.numpy(), inner)                                        # 308
    mod = np.full_like(x, np.float32(INIT_MODIFIER))                           # 304
    m = np.zeros_like(x)                                                       # Adam state (306)
    v = np.zeros_like(x)

    with torch.no_grad():
        init = [[a.detach().clone() for a in hm.run(frames)] for hm in hooked]  # 318-323
    L = sum(len(i) for i in init)

    cost_log = np.zeros(steps, dtype=np.float32)
    weights = [] if adaptive else None
    if adaptive:
        coeffs = np.asarray(coeffs, dtype=np.float32).copy()
        prev = np.ones(L, dtype=np.float32)                                    # TPAMI_attack.py:257
    true_np = O.compose_norm(x, mod, epsilon, inner)                           # 331-332

    for i in range(steps):
        if adaptive:
            coeffs, w_up = O.layer_reweight(coeffs, prev, momentum)            # TPAMI_attack.py:265
            weights.append(coeffs.copy())                                      # 266
        true_image = torch.from_numpy(true_np.copy()).requires_grad_(True)
        cos_rows, feats, ups = [], [], []
        layer = 0
        for hm, init_feats in zip(hooked, init):                               # 334 / 469-470
            acts = hm.run(true_image)
            for a, a0 in zip(acts, init_feats):
                if cos_mode == "torch":
                    cs = F.cosine_similarity(a.view(N, -1), a0.view(N, -1))    # 341-343
                    cos_rows.append(cs)
                else:
                    cs, gr = O.cosine_loss_grad_f64(a.detach().numpy().reshape(N, -1), a0.numpy().reshape(N, -1),
                                                    w=float(w_up[layer]) if adaptive else 1.0)
                    cos_rows.append(torch.from_numpy(cs.astype(np.float32)))
                    feats.append(a)
                    ups.append(torch.from_numpy(gr.astype(np.float32)).view_as(a))
                layer += 1
        if cos_mode == "torch":
            stacked = torch.stack(cos_rows)                                    # [L, N]
            if adaptive:
                used = torch.from_numpy(coeffs).unsqueeze(1)                   # 289
                each = torch.sum(used * stacked, dim=1)                        # 290
                cost = torch.mean(each)                                        # 291
                prev = (each if coef_CE else torch.sum(stacked.detach(), dim=1)).detach().numpy().copy()  # 293-297
            else:
                cost = torch.sum(stacked)                                      # 347
            if weight_grads:
                cost.backward()                                                # 352, weights included
                g = true_image.grad
            else:
                (g,) = torch.autograd.grad(cost, true_image)                   # 352
            cost_val = np.float32(cost.detach().numpy())
            cos_np = stacked.detach().numpy()
        else:
            (g,) = torch.autograd.grad(feats, true_image, ups)
            cos_np = torch.stack(cos_rows).numpy()
            cost_val, prev_new = O.layer_sums(cos_np, coeffs if adaptive else None, mode=1 if adaptive else 0,
                                              coef_CE=coef_CE)
            if adaptive:
                prev = prev_new
        cost_log[i] = cost_val
        g_np = g.numpy()
        m, v, mod, true_np = O.adam_compose(g_np, m, v, mod, x, epsilon, inner, i + 1, step_size)   # 351-353, 331-332
        if tap is not None:
            tap(i, dict(g=g_np, m=m, v=v, mod=mod, true_image=true_np, cos=cos_np))

    adv = true_np.reshape(b, f, c, h, w).transpose(0, 2, 1, 3, 4)              # 360-363
    return adv, cost_log, (np.stack(weights) if adaptive and steps else None), coeffs


def dispersion_loop(hooked, videos, epsilon, steps, step_size, loss_mode="torch", tap=None):
    """image_attacks.py:190-234 (`ImageGuidedStd_Adam.forward`, Dispersion Reduction).

    hooked   : [HookedModel] (one image model)
    loss_mode: 'torch' — Tensor.std() + autograd in f32, the reference's own arithmetic
               'f64'   — the float64 analytic std gradient of oracle.std_loss_grad_f64
    Returns (adv [b,3,f,h,w] float32 ndarray, cost [steps] float32)."""
    videos = torch.as_tensor(videos, dtype=torch.float32)
    b, c, f, h, w = videos.shape
    inner = h * w
    frames = _frames(videos)                                                   # 195-196
    x = O.denorm(frames.numpy(), inner)                                        # 201
    mod = np.full_like(x, np.float32(INIT_MODIFIER))                           # 197
    m = np.zeros_like(x)                                                       # Adam state (199)
    v = np.zeros_like(x)
    cost_log = np.zeros(steps, dtype=np.float32)
    true_np = O.compose_norm(x, mod, epsilon, inner)                           # 211-212
    for i in range(steps):
        true_image = torch.from_numpy(true_np.copy()).requires_grad_(True)
        acts = [a for hm in hooked for a in hm.run(true_image)]                # 214
        if loss_mode == "torch":
            cost = torch.sum(torch.stack([a.std() for a in acts]))             # 216-220
            (g,) = torch.autograd.grad(cost, true_image)                       # 222
            cost_val = np.float32(cost.detach().numpy())
        else:
            ups, total = [], 0.0
            for a in acts:
                sd, _, gr = O.std_loss_grad_f64(a.detach().numpy())
                total += sd
                ups.append(torch.from_numpy(gr.astype(np.float32)).view_as(a))
            (g,) = torch.autograd.grad(acts, true_image, ups)
            cost_val = np.float32(total)
        cost_log[i] = cost_val
        g_np = g.numpy()
        m, v, mod, true_np = O.adam_compose(g_np, m, v, mod, x, epsilon, inner, i + 1, step_size)   # 221-223, 211-212
        if tap is not None:
            tap(i, dict(g=g_np, m=m, v=v, mod=mod, true_image=true_np))
    adv = true_np.reshape(b, f, c, h, w).transpose(0, 2, 1, 3, 4)              # 230-234
    return adv, cost_log


def teacher_forced_grad(hooked, frames, true_image, dtype=torch.float32, weights=None, relu_masks=None, pool_indices=None):
    """dcost/dtrue_image of ONE step for a GIVEN true_image (image_attacks.py:334-352), in `dtype`.

    float64 is the accuracy arbiter of SURVEY.md 8(c): both float32 implementations (torch on the CPU —
    the reference's arithmetic — and the CUDA path) are scored by their distance to it, because at
    step 1 the gradient is a cancellation-level quantity and the reference does not reproduce itself
    across reduction orders (D8).  `hooked` models are converted to `dtype` in place.
    relu_masks: optional, one list of [N,C,h,w] activity masks per hooked model (ReLU execution order) that the
    adversarial forward/backward is forced to use (see forced_relu_masks); then a 4th value (flips, total) is returned.
    pool_indices: optional, one list of int64 [N,C,P,Q] max-pool winners per hooked model (MaxPool2d execution order),
    forced together with the ReLU decisions (differing winners are counted in `flips`).
    Returns (cost, grad [N,3,H,W] as float64 ndarray, cos [L,N])."""
    frames = torch.as_tensor(frames).to(dtype)
    flips = total = 0
    ti = torch.as_tensor(np.ascontiguousarray(true_image)).to(dtype).requires_grad_(True)
    N = frames.shape[0]
    rows = []
    for hm in hooked:
        hm.model.to(dtype)
        with torch.no_grad():
            init = [a.detach().clone() for a in hm.run(frames)]
        if relu_masks is None:
            acts = hm.run(ti)
        else:
            pools = pool_indices[hooked.index(hm)] if pool_indices is not None else None
            with forced_relu_masks(hm.model, relu_masks[hooked.index(hm)], pools) as fm:
                acts = hm.run(ti)
            flips, total = flips + fm.flips + fm.pool_flips, total + fm.total + fm.pool_total
        for a, a0 in zip(acts, init):
            rows.append(F.cosine_similarity(a.view(N, -1), a0.view(N, -1)))
    stacked = torch.stack(rows)
    if weights is None:
        cost = torch.sum(stacked)
    else:
        cost = torch.mean(torch.sum(torch.as_tensor(weights).to(dtype).unsqueeze(1) * stacked, dim=1))
    (g,) = torch.autograd.grad(cost, ti)
    if relu_masks is not None:
        return float(cost.detach()), g.double().numpy(), stacked.detach().double().numpy(), (flips, total)
    return float(cost.detach()), g.double().numpy(), stacked.detach().double().numpy()


# --------------------------------------------------------------------------------------------------
# base_attacks.py:242-340
# --------------------------------------------------------------------------------------------------
def _ce_grad(model, adv, labels, targeted):
    adv_t = torch.from_numpy(adv.copy()).requires_grad_(True)
    out = model(adv_t)
    cost = targeted * torch.nn.CrossEntropyLoss()(out, labels)                 # 285
    (g,) = torch.autograd.grad(cost, adv_t)                                    # 286-287
    return g.numpy()


def fgsm(model, videos, labels, epsilon=16 / 255, targeted=1):
    """base_attacks.py:242-259"""
    model.eval()
    videos = np.ascontiguousarray(videos, dtype=np.float32)
    inner = videos.shape[2] * videos.shape[3] * videos.shape[4]
    g = _ce_grad(model, videos, labels, targeted)
    return O.sign_step_project(videos, g, None, epsilon, epsilon, inner, project=False)


def bim(model, videos, labels, epsilon=16 / 255, steps=10, targeted=1):
    """base_attacks.py:272-295"""
    model.eval()
    videos = np.ascontiguousarray(videos, dtype=np.float32)
    inner = videos.shape[2] * videos.shape[3] * videos.shape[4]
    step_size = epsilon / steps                                                # 270
    x = O.denorm(videos, inner)                                                # 279
    adv = videos.copy()                                                        # 280
    for _ in range(steps):
        g = _ce_grad(model, adv, labels, targeted)                             # 283-287
        adv = O.sign_step_project(adv, g, x, step_size, epsilon, inner)        # 289-293
    return adv


def mifgsm(model, videos, labels, epsilon=16 / 255, steps=10, decay=1.0, targeted=1):
    """base_attacks.py:309-340 with utils.py:58-67 (frame-level norm)"""
    model.eval()
    videos = np.ascontiguousarray(videos, dtype=np.float32)
    inner = videos.shape[2] * videos.shape[3] * videos.shape[4]
    step_size = epsilon / steps                                                # 306
    momentum = np.zeros_like(videos)                                           # 316
    x = O.denorm(videos, inner)                                                # 317
    adv = videos.copy()                                                        # 318
    for _ in range(steps):
        g = _ce_grad(model, adv, labels, targeted)                             # 321-326
        norm = O.frame_absmean(g)                                              # 328 -> utils.py:63
        adv, momentum = O.mi_sign_step_project(adv, g, momentum, norm, x, decay, step_size, epsilon)   # 328-338
    return adv


# --------------------------------------------------------------------------------------------------
# base_attacks.py:342-683 — transfer-enhancing variants (same update block, different gradient)
# --------------------------------------------------------------------------------------------------
def _sign_loop(model, videos, labels, epsilon, steps, grad_fn, accumulate=None, targeted=1):
    """The loop every variant repeats (e.g. base_attacks.py:378-411): grad -> [momentum] -> update block."""
    model.eval()
    videos = np.ascontiguousarray(videos, dtype=np.float32)
    inner = videos.shape[2] * videos.shape[3] * videos.shape[4]
    step_size = epsilon / steps
    momentum = np.zeros_like(videos)
    x = O.denorm(videos, inner)
    adv = videos.copy()
    for _ in range(steps):
        g = grad_fn(adv)
        if accumulate is not None:
            g, momentum = accumulate(g, momentum)
        adv = O.sign_step_project(adv, g, x, step_size, epsilon, inner)
    return adv


def _l1_momentum(decay):
    def acc(g, momentum):                                                       # base_attacks.py:394-398
        gt = torch.from_numpy(g)
        gt = gt / torch.norm(gt, p=1)
        gt = gt + torch.from_numpy(momentum) * decay
        return gt.numpy(), gt.numpy()
    return acc


def _plain_momentum(decay):
    def acc(g, momentum):                                                       # base_attacks.py:463-465
        gt = torch.from_numpy(g) + torch.from_numpy(momentum) * decay
        return gt.numpy(), gt.numpy()
    return acc


def input_diversity(videos):
    """base_attacks.py:356-376, statement by statement (RNG consumption order included)."""
    import random
    if random.random() < 0.5:
        return videos
    rnd = torch.randint(224, 250, size=(1, 1)).item()
    rescaled = videos.view((-1,) + videos.shape[2:])
    rescaled = F.interpolate(rescaled, size=[rnd, rnd], mode="nearest")
    h_rem = 250 - rnd
    w_rem = 250 - rnd
    pad_top = torch.randint(0, h_rem, size=(1, 1)).item()
    pad_bottom = h_rem - pad_top
    pad_left = torch.randint(0, w_rem, size=(1, 1)).item()
    pad_right = w_rem - pad_left
    padded = F.pad(rescaled, [pad_left, pad_right, pad_top, pad_bottom])
    padded = F.interpolate(padded, size=[224, 224], mode="nearest")
    return padded.view(videos.shape)


def difgsm(model, videos, labels, epsilon=16 / 255, steps=10, decay=1.0, momentum=False, targeted=1):
    """base_attacks.py:378-411"""
    def grad_fn(adv):
        adv_t = torch.from_numpy(adv.copy()).requires_grad_(True)
        cost = targeted * torch.nn.CrossEntropyLoss()(model(input_diversity(adv_t)), labels)
        return torch.autograd.grad(cost, adv_t)[0].numpy()
    return _sign_loop(model, videos, labels, epsilon, steps, grad_fn, _l1_momentum(decay) if momentum else None)


def gaussian_kernel(kernlen=15, nsig=3, dims=2):
    """base_attacks.py:427-432 (2-D) and 624-633 (3-D), float32 as the reference casts it."""
    x = np.linspace(-nsig, nsig, kernlen)
    kern1d = np.exp(-0.5 * x * x) / np.sqrt(2.0 * np.pi)                        # scipy.stats.norm.pdf
    raw = np.outer(kern1d, kern1d)
    if dims == 2:
        return (raw / raw.sum()).astype(np.float32)
    used = np.stack([kern1d[i] * raw for i in range(kernlen)])
    return (used / used.sum()).astype(np.float32)


def tifgsm(model, videos, labels, epsilon=16 / 255, steps=10, decay=1.0, momentum=False, targeted=1):
    """base_attacks.py:451-479 with _conv2d_frame 434-449 (depth-wise conv per frame, mean over dims (1,2,3))."""
    k = torch.from_numpy(gaussian_kernel(15, 3, 2))
    stack = k.expand(3, 1, 15, 15).contiguous()

    def grad_fn(adv):
        g = torch.from_numpy(_ce_grad(model, adv, labels, targeted))
        out = torch.zeros_like(g)
        for i in range(g.shape[2]):
            out[:, :, i] = F.conv2d(g[:, :, i], stack, groups=3, stride=1, padding=7)
        out = out / torch.mean(torch.abs(out), [1, 2, 3], True)
        return out.numpy()
    return _sign_loop(model, videos, labels, epsilon, steps, grad_fn, _plain_momentum(decay) if momentum else None)


def sim(model, videos, labels, epsilon=16 / 255, steps=10, decay=1.0, scale_step=5, momentum=False, targeted=1):
    """base_attacks.py:563-611"""
    def grad_fn(adv):
        mean_grad = None
        for i in range(scale_step):
            g = _ce_grad(model, (1 / 2 ** i * torch.from_numpy(adv)).numpy(), labels, targeted)
            mean_grad = g if mean_grad is None else mean_grad + g
        return mean_grad / scale_step
    return _sign_loop(model, videos, labels, epsilon, steps, grad_fn, _l1_momentum(decay) if momentum else None)


def sgm(model, videos, labels, epsilon=16 / 255, steps=10, decay=1.0, gamma=0.5, momentum=False, targeted=1):
    """base_attacks.py:495-551: backward hooks scale the gradient through the named ReLU modules by gamma**0.5."""
    scale = float(np.power(gamma, 0.5))
    handles = [m.register_full_backward_hook(lambda mod, gin, gout: (scale * gin[0],))
               for name, m in model.named_modules() if "relu" in name and "0.relu" not in name and isinstance(m, torch.nn.ReLU)]
    try:
        return _sign_loop(model, videos, labels, epsilon, steps, lambda adv: _ce_grad(model, adv, labels, targeted),
                          _l1_momentum(decay) if momentum else None)
    finally:
        for h in handles:
            h.remove()


def tifgsm3d(model, videos, labels, epsilon=16 / 255, steps=10, decay=1.0, momentum=False, targeted=1):
    """base_attacks.py:651-683 with _conv3d_frame 635-649 (conv3d + utils.norm_grads frame level)."""
    k = torch.from_numpy(gaussian_kernel(15, 3, 3))
    stack = k.expand(3, 1, 15, 15, 15).contiguous()

    def grad_fn(adv):
        g = torch.from_numpy(_ce_grad(model, adv, labels, targeted))
        out = F.conv3d(g, stack, groups=3, stride=1, padding=7)
        norm = O.frame_absmean(out.numpy())
        return out.numpy() / norm[:, None, :, None, None]
    return _sign_loop(model, videos, labels, epsilon, steps, grad_fn, _plain_momentum(decay) if momentum else None)


# --------------------------------------------------------------------------------------------------
# video_attacks.py:14-229 — TemporalTranslation
# --------------------------------------------------------------------------------------------------
def tt_kernel(kernlen, mode):
    """video_attacks.py:51-78 (`np.math.exp` restated with math.exp)."""
    import math
    if mode == "gaussian":
        k = (kernlen - 1) / 2
        sigma = k / 3
        k = int(k)
        kern1d = np.array([1 / (sigma * np.sqrt(2 * np.pi)) * math.exp(-(x ** 2) / (2 * (sigma ** 2))) for x in range(-k, k + 1)])
    elif mode == "linear":
        k = int((kernlen - 1) / 2)
        half = [1 - i / (k + 1) for i in range(k + 1)]
        kern1d = np.array(half[::-1][:-1] + half)
    else:
        kern1d = np.ones(kernlen)
    return (kern1d / kern1d.sum()).astype(np.float32)


def _cycle(videos, move, frames):
    """video_attacks.py:93-105: new[:, :, (i + direction*|move| % frames) % frames] = videos[:, :, i]."""
    direction = -1 if move < 0 else 1
    amount = abs(move) % frames
    new = torch.zeros_like(videos)
    for i in range(frames):
        new[:, :, (i + direction * amount) % frames] = videos[:, :, i]
    return new


def _cycle_amount(move, frames, move_type):
    """video_attacks.py:93-135: the frame shift each move type applies for a nominal `move` (random: one randint draw)."""
    import random
    direction = -1 if move < 0 else 1
    amount = abs(move)
    if move_type == "adj":
        amount = amount % frames
    elif move_type == "large":
        amount = amount % frames if amount == 0 else (amount + (int(frames / 2) - 1)) % frames
    else:
        amount = amount % frames if move == 0 else random.randint(0, 100) % frames
    return direction * amount


def temporal_translation(model, videos, labels, kernlen, weight, momentum=False, kernel_mode="gaussian", epsilon=16 / 255,
                         steps=10, delay=1.0, targeted=1, tpnet=False, move_type="adj"):
    """video_attacks.py:179-229; batch of one clip, as the reference needs.  The INPUT shifts follow `move_type`
    (192-199); the gradients are shifted back by the NOMINAL moves whatever the move type (172-173 call `_cycle_move`)."""
    import math
    model.eval()
    videos = np.ascontiguousarray(videos, dtype=np.float32)
    B, C, T, H, W = videos.shape
    inner = T * H * W
    frames = T                                                                  # 36 hard-codes 32
    step_size = epsilon / steps
    max_move = int((kernlen - 1) / 2)
    moves = [i for i in range(-max_move, max_move + 1)]                         # 47-50
    kernel = torch.from_numpy(tt_kernel(kernlen, kernel_mode))[None]            # 45
    mom = np.zeros_like(videos)                                                 # 182
    x = O.denorm(videos, inner)                                                 # 185
    adv = videos.copy()                                                         # 186
    for _ in range(steps):
        adv_t = torch.from_numpy(adv)
        batch_inps = torch.cat([_cycle(adv_t, _cycle_amount(m, frames, move_type), frames) for m in moves], dim=0)   # 191-200
        length = len(moves)
        batch_times = length if tpnet else 5                                    # 202-206
        batch_size = math.ceil(length / batch_times)
        grads = []
        for i in range(batch_times):                                            # 208-210
            sl = batch_inps[i * batch_size:min((i + 1) * batch_size, length)]
            if sl.shape[0] == 0:
                continue                                                        # (the reference raises in torch.cat([]) here)
            used_labels = torch.cat([labels] * sl.shape[0], dim=0)              # 152
            inp = sl.clone().requires_grad_(True)
            cost = targeted * torch.nn.CrossEntropyLoss()(model(inp), used_labels)
            grads.append(torch.autograd.grad(cost, inp)[0])
        grads = torch.unsqueeze(torch.cat(grads, dim=0), dim=1)                 # 212-213: [D, N, C, T, H, W]
        same = grads.clone()                                                    # 170
        diff = torch.zeros_like(grads)
        for ind, m in enumerate(moves):                                         # 172-173
            diff[ind] = _cycle(grads[ind], -m, frames)
        D = grads.shape[0]
        s_conv = torch.matmul(kernel, same.reshape(D, -1)).reshape(grads.shape[1:])        # 80-91
        d_conv = torch.matmul(kernel, diff.reshape(D, -1)).reshape(grads.shape[1:])
        g = ((1 - weight) * s_conv + weight * d_conv).numpy()                   # 176
        if momentum:                                                            # 217-220
            norm = O.frame_absmean(g)
            adv, mom = O.mi_sign_step_project(adv, g, mom, norm, x, delay, step_size, epsilon)
        else:
            adv = O.sign_step_project(adv, g, x, step_size, epsilon, inner)     # 224-228
    return adv


# --------------------------------------------------------------------------------------------------
# base_attacks.py:685-814 — TAP
# --------------------------------------------------------------------------------------------------
def tap(model, layers, videos, labels, kernlen=3, temporal_kernlen=3, conv3d=True, epsilon=16 / 255, steps=10, targeted=1):
    """base_attacks.py:757-814; `layers` = the hooked modules (738-744).  Returns (adv, [per-step (ce, reg, distance)])."""
    model.eval()
    store = []
    handles = [m.register_forward_hook(lambda mod, inp, out: store.append(out)) for m in layers]
    try:
        videos_np = np.ascontiguousarray(videos, dtype=np.float32)
        videos_t = torch.from_numpy(videos_np)
        batch_size = videos_np.shape[0]
        inner = videos_np.shape[2] * videos_np.shape[3] * videos_np.shape[4]
        step_size = epsilon / steps
        k2 = torch.full((3, 1, kernlen, kernlen), 1.0 / (kernlen * kernlen))                               # 701-703
        k3 = torch.full((3, 1, temporal_kernlen, kernlen, kernlen), 1.0 / (temporal_kernlen * kernlen * kernlen))   # 705-707
        std = torch.tensor(O.STD)[:, None, None, None]
        del store[:]
        model(videos_t)                                                          # 768-769
        ori = [f.detach() for f in store]
        x = O.denorm(videos_np, inner)                                           # 772
        adv = videos_np.copy()                                                   # 773
        info = []
        for _ in range(steps):
            del store[:]
            adv_t = torch.from_numpy(adv.copy()).requires_grad_(True)
            outputs = model(adv_t)                                               # 779
            cost1 = targeted * torch.nn.CrossEntropyLoss()(outputs, labels)      # 782
            dist = []
            for i, j in zip(list(store), ori):                                   # 787-789
                dist.append(torch.norm((torch.sign(i) * torch.sqrt(torch.abs(i))).reshape(batch_size, -1) -
                                       (torch.sign(j) * torch.sqrt(torch.abs(j))).reshape(batch_size, -1), p=2, dim=1))
            cost2 = torch.sum(torch.stack(dist), 0)                              # 790
            perts = (adv_t - videos_t) / std                                     # 792 (_transform_perts)
            if conv3d:                                                           # 793-796
                out = F.conv3d(perts, k3, groups=3, stride=1,
                               padding=[int((temporal_kernlen - 1) / 2), int((kernlen - 1) / 2), int((kernlen - 1) / 2)])
            else:
                out = torch.zeros_like(perts)
                for t in range(perts.shape[2]):
                    out[:, :, t] = F.conv2d(perts[:, :, t], k2, groups=3, stride=1, padding=[int((kernlen - 1) / 2)] * 2)
            reg = torch.sum(torch.abs(out))
            cost = cost1 + 1e3 * reg + 0.05 * cost2                              # 799
            (g,) = torch.autograd.grad(cost.sum(), adv_t)
            info.append((float(cost1), float(reg), cost2.detach().numpy().copy()))
            adv = O.sign_step_project(adv, g.numpy(), x, step_size, epsilon, inner)   # 806-810
        return adv, info
    finally:
        for h in handles:
            h.remove()


# --------------------------------------------------------------------------------------------------
# image_attacks.py:498-629 — ILAF
# --------------------------------------------------------------------------------------------------
def ilaf(model, layers, videos, ori_videos, epsilon=16 / 255, steps=60, step_size=0.005):
    """image_attacks.py:534-629; `layers` = the hooked modules (514-520).  Returns (returned tensor, unscrambled clip,
    [per-step cost])."""
    store = []
    handles = [m.register_forward_hook(lambda mod, inp, out: store.append(out)) for m in layers]
    try:
        videos_np = np.ascontiguousarray(videos, dtype=np.float32)
        ori_np = np.ascontiguousarray(ori_videos, dtype=np.float32)
        b, c, f, h, w = videos_np.shape
        inner = f * h * w
        with torch.no_grad():
            del store[:]
            model(torch.from_numpy(ori_np))                                      # 542-550
            ori_maps = [a.detach() for a in store]
            del store[:]
            model(torch.from_numpy(videos_np))                                   # 553-561
            adv_maps = [a.detach() for a in store]
        init_dirs, init_norms = [], []
        for o, a in zip(ori_maps, adv_maps):                                     # 563-569
            d = a - o
            n = torch.norm(d, p=2)
            init_norms.append(n)
            init_dirs.append(d / torch.norm(d, p=2, keepdim=True))
        ori_unnorm = O.denorm(ori_np, inner)                                     # 573
        modifier = O.denorm(videos_np, inner) - ori_unnorm                       # 572, 575
        costs = []
        true_image = O.compose_norm(ori_unnorm, modifier, epsilon, inner)        # 582-585
        for _ in range(steps):
            del store[:]
            inp = torch.from_numpy(true_image.copy()).requires_grad_(True)
            model(inp)                                                           # 588
            losses = []
            for o, s, d0, n0 in zip(ori_maps, list(store), init_dirs, init_norms):   # 596-611
                sd = s - o
                sn = torch.norm(sd, p=2)
                sdir = sd / torch.norm(sd, p=2, keepdim=True)
                magnitude_gain = sn / n0
                angle_loss = torch.mm(d0.view(1, -1), sdir.view(1, -1).transpose(1, 0))
                losses.append(-(0.5 * magnitude_gain + angle_loss))
            cost = torch.sum(torch.stack(losses))                                # 612
            (g,) = torch.autograd.grad(cost, inp)                                # 614 (through the compose block below)
            costs.append(float(cost))
            modifier, true_image = O.sign_descent_compose(g.numpy(), modifier, ori_unnorm, epsilon, step_size, inner)   # 615-617
        out = torch.from_numpy(true_image).reshape(b, f, c, h, w).permute([0, 2, 1, 3, 4])   # 627-629
        return out.numpy(), true_image, costs
    finally:
        for h in handles:
            h.remove()
