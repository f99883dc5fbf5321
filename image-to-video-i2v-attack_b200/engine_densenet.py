"""Native DenseNet feature extractor (BASELINE.json configs[2]: the DenseNet-121 member of the ensemble): the truncated
forward up to the deepest hooked dense block and its input data-gradient on this repo's sm_100a kernels.

The reference only constructs the net (image_attacks.py:95-98 — `models.densenet161(pretrained=True)`; its
`_find_target_layer`, 260-271, has no DenseNet branch and would crash); SURVEY.md D3 defines the hook this repo adds:
depth d -> `model.features.denseblock{d}`, whose output is the channel concatenation of the block input and every
layer's 32 (48 for densenet161) new feature maps — NOT a ReLU output, so K1 must not apply a ReLU-backward mask
(`relu_masked_grads = False`).

torchvision's graph (densenet.py) and how it maps onto the kernels:

    conv0 7x7/s2 + norm0 + relu0     first-layer kernels of NativeEngine (BN folded, ReLU in the epilogue)
    pool0 3x3/s2 max                 i2v_maxpool_fwd_f32 (ReLU-backward mask folded into the argmax plane)
    dense layer i of block b         t = relu(norm1(cat[:, :Cin]))        i2v_bn_relu_f32   (pre-activation: cannot be folded)
                                     u = relu(norm2(conv1 1x1 (t)))       tensor-core conv, norm2 folded, ReLU epilogue
                                     cat[:, Cin:Cin+g] = conv2 3x3 (u)    tensor-core conv (output channels zero-padded to 64)
                                                                          + i2v_copy_channels_f32
    transition b                     t = relu(norm(cat)); v = conv 1x1 (t); next cat[:, :C/2] = avgpool2(v)   i2v_avgpool2_fwd_f32

Backward (data gradient only; weight gradients are never needed).  G_cat[b] accumulates dcost/dcat:

    transition:  g_v = avgpool2 backward of G_cat[b+1][:, :C/2];  G_cat[b] (+)= dgrad_1x1(g_v) * 1[t > 0] * scale
    layer i (reverse order):  g_u = dgrad_3x3(G_cat[b][:, Cin:Cin+g]) * 1[u > 0]
                              G_cat[b][:, :Cin] += dgrad_1x1(g_u) * 1[t > 0] * scale1
    The per-channel BatchNorm scale of a pre-activation is folded into the data-gradient weights of its consumer and the
    1[t > 0] mask is the convolution epilogue's, so pre-activations have no backward kernel.

Channel counts that the tensor-core kernels do not take (Cin not a multiple of 64 in the 1x1 data gradient, 32 output
channels in the 3x3) are zero-padded on the host: t is written [M, ceil64(Cin)] with zero tail channels, the 3x3 weights
get 64 output channels of which the first g are real.
"""
import os

import torch
import torch.nn as nn

from . import backbones, capi
from .engine_native import NativeEngine, _Conv, _Pool, _split_tf32, _pad_cols


def _ceil64(c):
    return (c + 63) // 64 * 64


def _bn_affine(bn):
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    shift = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
    return scale.contiguous(), shift.contiguous()


def _padded_conv(conv, cin_pad=None, cout_pad=None):
    """A Conv2d with the same filter, input / output channels zero-padded (weights only; no bias in DenseNet convs)."""
    cout, cin, R, S = conv.weight.shape
    ci, co = cin_pad or cin, cout_pad or cout
    if ci == cin and co == cout:
        return conv
    c2 = nn.Conv2d(ci, co, (R, S), stride=conv.stride, padding=conv.padding, bias=False).to(conv.weight.device)
    with torch.no_grad():
        c2.weight.zero_()
        c2.weight[:cout, :cin] = conv.weight
    return c2


def _fold_dgrad_scale(op, scale):
    """dcost/d(pre-BN input)[c] = scale[c] * 1[t_c > 0] * (W^T g)[c]: scale the rows (= data-gradient output channels) of
    the data-gradient operands of `op` by the pre-activation's per-channel BatchNorm scale (zero for padded channels)."""
    s = torch.zeros(op.cin, device=scale.device)
    s[:scale.numel()] = scale
    w = op.tc_dgrad[0] * s.view(-1, 1)
    op.tc_dgrad = _split_tf32(w.contiguous())
    op.b_dgrad = _pad_cols(op.b_dgrad[:, :op.cin] * s.view(1, -1))


class _DenseLayer:
    def __init__(self, name, layer, cin, growth):
        self.name, self.cin, self.cp, self.growth = name, cin, _ceil64(cin), growth
        self.scale1, self.shift1 = _bn_affine(layer.norm1)
        self.conv1 = _Conv(name + ".conv1", name + ".t", name + ".u", _padded_conv(layer.conv1, cin_pad=self.cp), layer.norm2, True)
        _fold_dgrad_scale(self.conv1, self.scale1)
        self.gp = _ceil64(growth)
        self.conv2 = _Conv(name + ".conv2", name + ".u", name + ".new", _padded_conv(layer.conv2, cout_pad=self.gp), None, False)


class _Transition:
    def __init__(self, name, tr, cin):
        self.name, self.cin, self.cp, self.cout = name, cin, _ceil64(cin), tr.conv.out_channels
        self.scale, self.shift = _bn_affine(tr.norm)
        self.conv = _Conv(name + ".conv", name + ".t", name + ".v", _padded_conv(tr.conv, cin_pad=self.cp), None, False)
        _fold_dgrad_scale(self.conv, self.scale)


class DenseNetEngine(NativeEngine):
    relu_masked_grads = False    # the hooked tensor is a concatenation, not a ReLU output
    preferred_chunk = 32

    def __init__(self, model, model_name, depth, tf32x3=True, use_tensor_cores=None):
        self.model = backbones.freeze_for_attack(model)
        self.model_name, self.depth, self.tf32x3 = model_name, depth, tf32x3
        self.targets = backbones.find_target_layers(model, model_name, depth)
        f = model.features
        depths = sorted(set(depth) if isinstance(depth, (list, tuple)) else {depth})
        self.hook_blocks = depths                       # forward execution order (= ascending depth)
        self.last_block = depths[-1]
        self.stem = _Conv("conv0", "img", "stem", f.conv0, f.norm0, True, x_nchw=True)
        self.pool = _Pool("pool0", "stem", "pool", f.pool0)
        self.blocks, self.transitions = [], []
        c = f.conv0.out_channels
        for b in range(1, self.last_block + 1):
            block = getattr(f, "denseblock%d" % b)
            layers = []
            c0 = c
            for i, (_, layer) in enumerate(block.items()):
                g = layer.conv2.out_channels
                layers.append(_DenseLayer("b%d.l%d" % (b, i), layer, c, g))
                c += g
            self.blocks.append(dict(c0=c0, ctot=c, layers=layers))
            if b < self.last_block:
                tr = getattr(f, "transition%d" % b)
                self.transitions.append(_Transition("tr%d" % b, tr, c))
                c = tr.conv.out_channels
        if use_tensor_cores is None:
            use_tensor_cores = os.environ.get("I2V_NATIVE_TC", "1") != "0"
        self.use_tc = bool(use_tensor_cores)
        self.use_stem = os.environ.get("I2V_NATIVE_STEM", "1") != "0"
        self.use_bits = False
        self.use_stem_tc = os.environ.get("I2V_NATIVE_STEM_TC", "1") != "0"
        self._zbuf = None
        self.stem_dgrad_direct = os.environ.get("I2V_STEM_DGRAD_DIRECT", "1") != "0"
        self.stem_dgrad_pool = os.environ.get("I2V_STEM_DGRAD_POOL", "1") != "0"
        self.stem_fwd_rows = os.environ.get("I2V_STEM_FWD_ROWS", "1") != "0"
        self.stem_direct = False
        self._xpbuf = None
        self._cache = {}
        self.buffer_generation = 0
        self.hook_bufs = ["cat%d" % b for b in self.hook_blocks]

    @property
    def num_layers(self):
        return len(self.hook_blocks)

    # ---- geometry / memory -------------------------------------------------------------------------------------
    def _dims(self, h, w):
        sh, sw = self.stem.out_hw(h, w)
        ph, pw = self.pool.out_hw(sh, sw)
        dims = {"stem": (sh, sw)}
        bh, bw = ph, pw
        for b in range(1, self.last_block + 1):
            dims[b] = (bh, bw)
            bh, bw = bh // 2, bw // 2
        return dims

    def bytes_per_frame(self, h, w):
        dims = self._dims(h, w)
        sh, sw = dims["stem"]
        total = 3 * h * w * 8 + sh * sw * 64 * 8 + dims[1][0] * dims[1][1] * 64 * 9
        for b, blk in enumerate(self.blocks, start=1):
            px = dims[b][0] * dims[b][1]
            ch = 2 * blk["ctot"] + sum(l.cp + 128 for l in blk["layers"]) + 3 * _ceil64(blk["ctot"]) + 4 * 128
            total += px * ch * 4
        return total

    def frames_per_chunk(self, h, w, n_frames, device, share=1.0):
        free, _ = torch.cuda.mem_get_info(device)
        fit = int(free * 0.6 * share // max(1, self.bytes_per_frame(h, w)))
        return max(1, min(n_frames, 64, fit))

    def _plan(self, n, h, w, device):
        key = (n, h, w)
        plan = self._cache.get(key)
        if plan is not None:
            return plan
        dims = self._dims(h, w)
        sh, sw = dims["stem"]
        f32 = dict(device=device, dtype=torch.float32)
        plan = dict(dims=dims, gimg=torch.empty(n, 3, h, w, **f32))
        plan["d_stem"] = capi.ConvDesc(n, h, w, 3, self.stem.cout, self.stem.R, self.stem.R, self.stem.stride, self.stem.pad, sh, sw)
        plan["stem"] = torch.empty(n, sh, sw, self.stem.cout, **f32)
        plan["g_stem"] = torch.empty_like(plan["stem"])
        ph, pw = dims[1]
        plan["pool"] = torch.empty(n, ph, pw, self.stem.cout, **f32)
        plan["g_pool"] = torch.empty_like(plan["pool"])
        plan["argmax"] = torch.empty(n, ph, pw, self.stem.cout, device=device, dtype=torch.uint8)
        max_cp = max_gp = 0
        for b, blk in enumerate(self.blocks, start=1):
            bh, bw = dims[b]
            plan["cat%d" % b] = torch.empty(n, bh, bw, blk["ctot"], **f32)
            if b not in self.hook_blocks:
                plan["g_cat%d" % b] = torch.empty(n, bh, bw, blk["ctot"], **f32)
            for l in blk["layers"]:
                plan[l.name + ".t"] = torch.empty(n, bh, bw, l.cp, **f32)
                plan[l.name + ".u"] = torch.empty(n, bh, bw, 128 if l.conv1.cout == 128 else l.conv1.cout, **f32)
                plan["d_" + l.name + ".conv1"] = capi.ConvDesc(n, bh, bw, l.cp, l.conv1.cout, 1, 1, 1, 0, bh, bw)
                plan["d_" + l.name + ".conv2"] = capi.ConvDesc(n, bh, bw, l.conv2.cin, l.gp, 3, 3, 1, 1, bh, bw)
                max_cp, max_gp = max(max_cp, l.cp), max(max_gp, l.gp)
            # per-block scratch shared by its layers: the padded 3x3 output / its gradient, dcost/du, dcost/dt
            plan["new%d" % b] = torch.empty(n, bh, bw, max(l.gp for l in blk["layers"]), **f32)
            plan["g_u%d" % b] = torch.empty(n, bh, bw, blk["layers"][0].conv1.cout, **f32)
            plan["g_t%d" % b] = torch.empty(n, bh, bw, max(l.cp for l in blk["layers"]), **f32)
        for b, tr in enumerate(self.transitions, start=1):
            bh, bw = dims[b]
            plan[tr.name + ".t"] = torch.empty(n, bh, bw, tr.cp, **f32)
            plan[tr.name + ".v"] = torch.empty(n, bh, bw, tr.cout, **f32)
            plan["g_" + tr.name + ".v"] = torch.empty(n, bh, bw, tr.cout, **f32)
            plan["g_" + tr.name + ".t"] = torch.empty(n, bh, bw, tr.cp, **f32)
            plan["d_" + tr.name] = capi.ConvDesc(n, bh, bw, tr.cp, tr.cout, 1, 1, 1, 0, bh, bw)
        if len(self._cache) > 4:
            self._cache.clear()
            self.buffer_generation += 1      # captured CUDA graphs that point into the old buffers are stale
        self._cache[key] = plan
        return plan

    # ---- forward -------------------------------------------------------------------------------------------------
    def features(self, img, need_grad, clone=True):
        n, c, h, w = img.shape
        if c != 3 or not img.is_contiguous():
            raise ValueError("expected a contiguous [n,3,H,W] image batch")
        P = self._plan(n, h, w, img.device)
        self._conv_fwd(self.stem, P["d_stem"], img, P["stem"], None)
        capi.maxpool_fwd(P["stem"], P["pool"], P["argmax"], self.pool.k, self.pool.stride, self.pool.pad, mark_dead=True)
        capi.copy_channels(P["pool"], P["cat1"], 0, 0, self.stem.cout)
        for b, blk in enumerate(self.blocks, start=1):
            cat = P["cat%d" % b]
            for l in blk["layers"]:
                t, u, new = P[l.name + ".t"], P[l.name + ".u"], P["new%d" % b]
                capi.bn_relu(cat, l.cin, l.scale1, l.shift1, t)                                       # norm1 + relu1
                self._conv_fwd(l.conv1, P["d_" + l.name + ".conv1"], t, u, None)                      # conv1 + norm2 + relu2
                newv = new if new.shape[-1] == l.gp else new.view(-1)[:n * cat.shape[1] * cat.shape[2] * l.gp].view(n, cat.shape[1], cat.shape[2], l.gp)
                self._conv_fwd(l.conv2, P["d_" + l.name + ".conv2"], u, newv, None)                   # conv2
                capi.copy_channels(newv, cat, 0, l.cin, l.growth)                                     # torch.cat
            if b < self.last_block:
                tr = self.transitions[b - 1]
                capi.bn_relu(cat, tr.cin, tr.scale, tr.shift, P[tr.name + ".t"])
                self._conv_fwd(tr.conv, P["d_" + tr.name], P[tr.name + ".t"], P[tr.name + ".v"], None)
                capi.avgpool2_fwd(P[tr.name + ".v"], P["cat%d" % (b + 1)], 0)
        self._last = P if need_grad else None
        self._last_fwd = P
        feats = [P[name] for name in self.hook_bufs]
        return [f.clone() for f in feats] if (not need_grad and clone) else feats

    # ---- backward ------------------------------------------------------------------------------------------------
    def input_grad(self, grads, out=None):
        P = self._last
        if P is None:
            raise RuntimeError("input_grad() needs a preceding features(..., need_grad=True)")
        gimg = P["gimg"] if out is None else out
        if gimg.shape != P["gimg"].shape or not gimg.is_contiguous():
            raise ValueError("out must be a contiguous %s tensor" % (tuple(P["gimg"].shape),))
        G = {}
        for b, g in zip(self.hook_blocks, grads):
            G[b] = g.view_as(P["cat%d" % b])
        for b in range(self.last_block, 0, -1):
            blk = self.blocks[b - 1]
            cat = P["cat%d" % b]
            n, bh, bw, _ = cat.shape
            if b < self.last_block:                                  # contribution through the transition above
                tr = self.transitions[b - 1]
                g_v, g_t = P["g_" + tr.name + ".v"], P["g_" + tr.name + ".t"]
                capi.avgpool2_bwd(G[b + 1], g_v, 0)
                self._conv_dgrad(tr.conv, P["d_" + tr.name], g_v, None, P[tr.name + ".t"], g_t)
                if b in G:
                    capi.copy_channels(g_t, G[b], 0, 0, tr.cin, accumulate=True)
                elif tr.cp == tr.cin:
                    G[b] = g_t
                else:
                    G[b] = P["g_cat%d" % b]
                    capi.copy_channels(g_t, G[b], 0, 0, tr.cin)
            gcat = G[b]
            for l in reversed(blk["layers"]):
                new = P["new%d" % b]
                g_new = new if new.shape[-1] == l.gp else new.view(-1)[:n * bh * bw * l.gp].view(n, bh, bw, l.gp)
                if l.gp != l.growth:
                    g_new.zero_()
                capi.copy_channels(gcat, g_new, l.cin, 0, l.growth)
                g_u = P["g_u%d" % b]
                self._conv_dgrad(l.conv2, P["d_" + l.name + ".conv2"], g_new, None, P[l.name + ".u"], g_u)
                gt_full = P["g_t%d" % b]
                g_t = gt_full if gt_full.shape[-1] == l.cp else gt_full.view(-1)[:n * bh * bw * l.cp].view(n, bh, bw, l.cp)
                self._conv_dgrad(l.conv1, P["d_" + l.name + ".conv1"], g_u, None, P[l.name + ".t"], g_t)
                capi.copy_channels(g_t, gcat, 0, 0, l.cin, accumulate=True)
        capi.copy_channels(G[1], P["g_pool"], 0, 0, self.stem.cout)
        pool = self.pool
        if (self.stem_dgrad_pool and self.stem_dgrad_direct and self.use_tc and self.use_stem_tc and self.tf32x3
                and (pool.k, pool.stride, pool.pad) == (3, 2, 1) and not pool.ceil and self.stem.cout == 64
                and getattr(self.stem, "tc_stem_dgrad_direct", None) is not None
                and capi.conv_stem_dgrad_pool_supported(P["d_stem"], P["g_pool"].shape[1], P["g_pool"].shape[2])):
            # pooling backward inside the first layer's data-gradient kernel: g_stem (4x the pooled bytes) is never written
            self._conv_dgrad(self.stem, P["d_stem"], None, None, None, gimg, pooled=(P["g_pool"], P["argmax"]))
        else:
            capi.maxpool_bwd(P["g_pool"], P["argmax"], None, P["g_stem"], pool.k, pool.stride, pool.pad)
            self._conv_dgrad(self.stem, P["d_stem"], P["g_stem"], None, None, gimg)
        self._last = None
        return gimg

    # ---- decisions, for the parity tests ---------------------------------------------------------------------------
    def relu_masks(self):
        """1[activation > 0] of every ReLU of the last forward in the torch module's ReLU call order (relu0; per dense
        layer relu1, relu2; per transition relu), [n,C,h,w] bool on the CPU."""
        P = self._last_fwd

        def m(t, c):
            return (t[..., :c] > 0).permute(0, 3, 1, 2).contiguous().cpu()
        out = [m(P["stem"], self.stem.cout)]
        for b, blk in enumerate(self.blocks, start=1):
            for l in blk["layers"]:
                out.append(m(P[l.name + ".t"], l.cin))
                out.append(m(P[l.name + ".u"], l.conv1.cout))
            if b < self.last_block:
                tr = self.transitions[b - 1]
                out.append(m(P[tr.name + ".t"], tr.cin))
        return out

    def pool_indices(self):
        plan = self._last_fwd
        am = plan["argmax"].to(torch.int64)
        n, Pp, Q, C = am.shape
        ih, iw = plan["dims"]["stem"]
        k, st, pad = self.pool.k, self.pool.stride, self.pool.pad
        p = torch.arange(Pp, device=am.device).view(1, Pp, 1, 1) * st - pad
        q = torch.arange(Q, device=am.device).view(1, 1, Q, 1) * st - pad
        r0, s0 = (-p).clamp(min=0), (-q).clamp(min=0)
        am = torch.where(am == 255, r0 * k + s0, am)
        flat = (p + am // k) * iw + (q + am % k)
        return [flat.permute(0, 3, 1, 2).contiguous().cpu()]
