"""Native feature extractor: the truncated backbone forward and its input data-gradient on this repo's
own sm_100a kernels (K4/K5), NHWC float32 activations.

What the reference runs per step through torchvision + cuDNN + autograd (image_attacks.py:334, 352):
the FULL network forward, then a backward that also produces every weight gradient.  Here:

  * the graph is cut at the deepest hooked layer (nothing after it feeds the loss);
  * eval-mode BatchNorm (image_attacks.py:253-256) is folded into the convolution weights and bias;
  * ReLU, residual adds and the ReLU-backward masks live in the convolution epilogues;
  * backward is the data gradient only; gradients are kept as "pre-activation" gradients, i.e. already
    multiplied by 1[activation > 0], so no separate ReLU-backward pass exists.

The graph is a small list of ops over named buffers, built from the torchvision module (weights are
taken from it, so random-init / pretrained policies of backbones.py apply unchanged):

    Conv(x -> y, folded weights, relu, residual)     MaxPool(x -> y)     Concat(xs -> y)   (Fire)

Supported families: resnet (Bottleneck nets), vgg, alexnet, squeezenet — every hook point of
reference image_attacks.py:260-271 and TPAMI_attack.py:176-200.  DenseNet (BN-ReLU-Conv ordering, dense
concatenation, average pooling) has its own engine on the same kernels: engine_densenet.DenseNetEngine.

Kernel selection per conv: the tcgen05 tensor-core implicit GEMM (conv_tc.cu) when it supports the
shape, otherwise the CUDA-core gather-GEMM (conv_simt.cu).  `tf32x3=True` (default, "FP32 parity
mode") runs the tensor cores with 3xTF32 split accumulation; False is plain TF32.
"""
import os

import torch
import torch.nn as nn
import torchvision

from . import backbones, capi


class _Conv:
    kind = "conv"

    def __init__(self, name, x, y, conv, bn, relu, residual=None, x_nchw=False):
        self.name, self.x, self.y, self.relu, self.residual, self.x_nchw = name, x, y, relu, residual, x_nchw
        self.relu_skipped_before = 0     # ReLU calls of the torch module that this graph does not execute (see relu_masks)
        w = conv.weight.detach().float()
        cout, cin, R, S = w.shape
        if conv.groups != 1 or conv.dilation != (1, 1) or R != S or conv.stride[0] != conv.stride[1] \
                or conv.padding[0] != conv.padding[1]:
            raise NotImplementedError("unsupported convolution %s" % (conv,))
        self.cin, self.cout, self.R, self.stride, self.pad = cin, cout, R, conv.stride[0], conv.padding[0]
        if bn is not None:
            scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
            shift = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
        else:
            scale = torch.ones(cout, device=w.device)
            shift = conv.bias.detach().float() if conv.bias is not None else torch.zeros(cout, device=w.device)
        ws = w * scale.view(-1, 1, 1, 1)
        # fprop B[(r,s,ci), co]; dgrad B[(r,s,co), ci]; columns padded to a multiple of 4
        bf = ws.permute(2, 3, 1, 0).reshape(R * S * cin, cout)
        bd = ws.permute(2, 3, 0, 1).reshape(R * S * cout, cin)
        self.b_fwd = _pad_cols(bf)
        self.b_dgrad = _pad_cols(bd)
        self.bias = _pad_vec(shift)
        # tensor-core operands: [Cout, (r,s,ci)] for fprop, flipped/transposed [Cin, (r,s,co)] for dgrad,
        # each split exactly into hi = trunc_tf32 and lo = w - hi (3xTF32), plus a round-to-nearest TF32 copy
        # for the plain-TF32 mode
        tf = ws.permute(0, 2, 3, 1).reshape(cout, R * S * cin).contiguous()
        td = ws.flip(2, 3).permute(1, 2, 3, 0).reshape(cin, R * S * cout).contiguous()
        self.tc_fwd = _split_tf32(tf)
        self.tc_dgrad = _split_tf32(td)
        # strided data gradient: one weight matrix per stride-parity class (see i2v_conv_tc_dgrad_class_f32)
        self.tc_dgrad_cls = _class_weights(ws, self.stride, self.pad) if self.stride > 1 and not x_nchw else None
        # first layer (Cin = 3): [(c,r,s), co] for the stem forward; its dgrad uses b_fwd = [(r,s), c, co]
        self.w_stem = ws.permute(1, 2, 3, 0).reshape(cin * R * S, cout).contiguous() if x_nchw else None
        # first-layer data gradient on the tensor cores: rows (c,r,s) zero-padded to a multiple of 64, TF32 hi/lo split
        self.tc_stem_dgrad = None
        if x_nchw and cin == 3 and cout % 32 == 0:
            nz = capi.stem_dgrad_tc_rows(cin * R * S)
            wz = torch.cat([self.w_stem, self.w_stem.new_zeros(nz - cin * R * S, cout)], 0).contiguous()
            self.tc_stem_dgrad = _split_tf32(wz)
        # first-layer data gradient without the Z^T scratch (7x7 / s2 / p3 stems): [160, 64] taps in two halves
        self.tc_stem_dgrad_direct = None
        if x_nchw and cin == 3 and cout == 64 and R == 7 and self.stride == 2 and self.pad == 3:
            self.tc_stem_dgrad_direct = _split_tf32(capi.stem_direct_dgrad_weights(self.w_stem))
        # EXPERIMENTAL direct first-layer forward ($I2V_STEM_DIRECT=1): [Cout, R*32] K-major, k = r*32 + s*4 + c
        self.tc_stem_direct = None
        if x_nchw and cin == 3 and cout == 64 and R == S and 5 <= S <= 8 and self.stride == 2:
            wr = ws.new_zeros(cout, R, 8, 4)
            wr[:, :, :S, :3] = ws.permute(0, 2, 3, 1)
            self.tc_stem_direct = _split_tf32(wr.reshape(cout, R * 32).contiguous())
        # first-layer forward as im2col + GEMM: [Cout, Kp] K-major, k = (c,r,s) zero-padded to a multiple of 32
        self.tc_stem_fwd = None
        if x_nchw and cin == 3 and cout % 64 == 0:
            kp = (cin * R * S + 31) // 32 * 32
            wk = torch.cat([self.w_stem.t(), self.w_stem.new_zeros(cout, kp - cin * R * S)], 1).contiguous()
            self.tc_stem_fwd = _split_tf32(wk)

    def out_hw(self, h, w):
        return ((h + 2 * self.pad - self.R) // self.stride + 1, (w + 2 * self.pad - self.R) // self.stride + 1)


class _Pool:
    kind = "pool"

    def __init__(self, name, x, y, mp):
        self.name, self.x, self.y = name, x, y
        k = mp.kernel_size if isinstance(mp.kernel_size, int) else mp.kernel_size[0]
        s = mp.stride if isinstance(mp.stride, int) else mp.stride[0]
        p = mp.padding if isinstance(mp.padding, int) else mp.padding[0]
        self.k, self.stride, self.pad, self.ceil = k, s, p, bool(mp.ceil_mode)

    def out_hw(self, h, w):
        def one(n):
            if self.ceil:
                o = -(-(n + 2 * self.pad - self.k) // self.stride) + 1
                if (o - 1) * self.stride >= n + self.pad:   # last window must start inside the input (torch rule)
                    o -= 1
                return o
            return (n + 2 * self.pad - self.k) // self.stride + 1
        return one(h), one(w)


class _Concat:
    kind = "concat"

    def __init__(self, name, xs, y):
        self.name, self.xs, self.y = name, xs, y


def _split_tf32(w):
    """(raw, lo, rna) operands of the tensor-core kernels.  raw = w itself: the tensor core reads an f32 bit pattern
    as TF32, i.e. it ignores the low 13 mantissa bits, so the raw weights ARE B_hi; lo = w - trunc_tf32(w) (exact in
    f32) — a non-null lo is what selects the 3xTF32 mode; rna = w rounded to nearest TF32 for the plain-TF32 mode."""
    w = w.contiguous()
    bits = w.view(torch.int32)
    hi = (bits & -8192).view(torch.float32)
    lo = (w - hi).contiguous()
    rna = ((bits + 4096) & -8192).view(torch.float32)
    return w, lo, rna.contiguous()


def _class_weights(ws, stride, pad):
    """{(ph, pw): (hi, lo, rna) of [ci, (tap_h, tap_w, co)] or None}: rows h = stride*i + ph only see the taps
    r = r0 + stride*a (r0 = (ph + pad) mod stride); tap_h = A_h-1-a, so the tap order is r descending."""
    cout, cin, R, S = ws.shape
    out = {}
    for ph in range(stride):
        rs = list(range((ph + pad) % stride, R, stride))[::-1]
        for pw in range(stride):
            ss = list(range((pw + pad) % stride, S, stride))[::-1]
            if not rs or not ss:
                out[(ph, pw)] = None
                continue
            sel = ws[:, :, rs, :][:, :, :, ss]                      # [co, ci, A_h, A_w]
            b = sel.permute(1, 2, 3, 0).reshape(cin, len(rs) * len(ss) * cout).contiguous()
            out[(ph, pw)] = _split_tf32(b)
    return out


def _pad_cols(m):
    k, n = m.shape
    n4 = (n + 3) // 4 * 4
    if n4 != n:
        m = torch.cat([m, m.new_zeros(k, n4 - n)], 1)
    return m.contiguous()


def _pad_vec(v):
    n = v.numel()
    n4 = (n + 3) // 4 * 4
    if n4 != n:
        v = torch.cat([v, v.new_zeros(n4 - n)])
    return v.contiguous()


# --------------------------------------------------------------------------------------------------
# graph builders
# --------------------------------------------------------------------------------------------------
def _build_resnet(model, targets):
    ops, hooks, chans, relu_typed = [], {}, {"img": 3}, set()
    ops.append(_Conv("conv1", "img", "stem", model.conv1, model.bn1, True, x_nchw=True))
    chans["stem"] = model.conv1.out_channels
    relu_typed.add("stem")
    ops.append(_Pool("maxpool", "stem", "pool", model.maxpool))
    chans["pool"] = chans["stem"]
    cur = "pool"
    remaining = set(id(t) for t in targets)
    for li in range(1, 5):
        layer = getattr(model, "layer%d" % li)
        for bi, blk in enumerate(layer):
            if not isinstance(blk, torchvision.models.resnet.Bottleneck):
                raise NotImplementedError("native engine implements Bottleneck ResNets (resnet50/101/152)")
            pre = "l%d.%d." % (li, bi)
            sc = cur
            if blk.downsample is not None:
                ops.append(_Conv(pre + "ds", cur, pre + "sc", blk.downsample[0], blk.downsample[1], False))
                chans[pre + "sc"] = blk.downsample[0].out_channels
                sc = pre + "sc"
            ops.append(_Conv(pre + "conv1", cur, pre + "c1", blk.conv1, blk.bn1, True))
            ops.append(_Conv(pre + "conv2", pre + "c1", pre + "c2", blk.conv2, blk.bn2, True))
            ops.append(_Conv(pre + "conv3", pre + "c2", pre + "out", blk.conv3, blk.bn3, True, residual=sc))
            chans[pre + "c1"], chans[pre + "c2"], chans[pre + "out"] = blk.conv1.out_channels, blk.conv2.out_channels, blk.conv3.out_channels
            relu_typed.update([pre + "c1", pre + "c2", pre + "out"])
            cur = pre + "out"
            if id(blk) in remaining:
                hooks[id(blk)] = cur
                remaining.discard(id(blk))
            if not remaining:
                return ops, hooks, chans, relu_typed
    raise ValueError("hooked layer not found in the ResNet graph")


def _build_sequential(features, targets):
    """VGG / AlexNet / SqueezeNet `features` (nn.Sequential of Conv2d, ReLU, MaxPool2d, Fire)."""
    Fire = torchvision.models.squeezenet.Fire
    ops, hooks, chans, relu_typed = [], {}, {"img": 3}, set()
    remaining = set(id(t) for t in targets)
    cur, i, first = "img", 0, True
    mods = list(features)

    def conv_relu(name, x, conv, relu_mod):
        nonlocal first
        y = name
        ops.append(_Conv(name, x, y, conv, None, relu_mod is not None, x_nchw=first))
        first = False
        chans[y] = conv.out_channels
        if relu_mod is not None:
            relu_typed.add(y)
        return y

    while i < len(mods) and remaining:
        m = mods[i]
        if isinstance(m, nn.Conv2d):
            relu_mod = mods[i + 1] if i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU) else None
            if id(m) in remaining and relu_mod is not None:
                raise NotImplementedError("hook on a pre-ReLU convolution output is not a reference hook point")
            cur = conv_relu("f%d" % i, cur, m, relu_mod)
            if relu_mod is not None:
                if id(relu_mod) in remaining:
                    hooks[id(relu_mod)] = cur
                    remaining.discard(id(relu_mod))
                i += 1
            elif id(m) in remaining:
                hooks[id(m)] = cur
                remaining.discard(id(m))
        elif isinstance(m, nn.MaxPool2d):
            y = "f%d" % i
            ops.append(_Pool(y, cur, y, m))
            chans[y] = chans[cur]
            cur = y
        elif isinstance(m, Fire):
            pre = "f%d." % i
            sq = conv_relu(pre + "squeeze", cur, m.squeeze, m.squeeze_activation)
            e3 = pre + "e3"
            if id(m.expand3x3_activation) in remaining and len(remaining) == 1:
                # scalar-depth hook (image_attacks.py:271): only the 3x3 branch of the last Fire is needed
                ops.append(_Conv(e3, sq, e3, m.expand3x3, None, True))
                ops[-1].relu_skipped_before = 1          # torch still runs expand1x1_activation before expand3x3_activation
                chans[e3] = m.expand3x3.out_channels
                relu_typed.add(e3)
                hooks[id(m.expand3x3_activation)] = e3
                remaining.discard(id(m.expand3x3_activation))
                break
            e1 = conv_relu(pre + "e1", sq, m.expand1x1, m.expand1x1_activation)
            ops.append(_Conv(e3, sq, e3, m.expand3x3, None, True))
            chans[e3] = m.expand3x3.out_channels
            relu_typed.add(e3)
            y = pre + "cat"
            ops.append(_Concat(y, [e1, e3], y))
            chans[y] = chans[e1] + chans[e3]
            relu_typed.add(y)
            cur = y
            for key in (id(m), id(m.expand3x3_activation)):
                if key in remaining:
                    hooks[key] = y if key == id(m) else e3
                    remaining.discard(key)
        elif isinstance(m, (nn.ReLU, nn.Dropout)):
            pass
        else:
            raise NotImplementedError("native engine: unsupported module %s" % type(m).__name__)
        i += 1
    if remaining:
        raise ValueError("hooked layer not found in the feature stack")
    return ops, hooks, chans, relu_typed


class NativeEngine:
    relu_masked_grads = True    # K1 applies 1[feature > 0]: gradients are kept pre-activation
    preferred_chunk = 128       # fallback when the image size is not known (see frames_per_chunk)
    graph_safe = True           # a step is a fixed sequence of launches on buffers that live as long as the engine
    writes_input_grad_in_place = True   # input_grad(grads, out=...)

    def __init__(self, model, model_name, depth, tf32x3=True, use_tensor_cores=None):
        self.model = backbones.freeze_for_attack(model)
        self.model_name, self.depth, self.tf32x3 = model_name, depth, tf32x3
        self.targets = backbones.find_target_layers(model, model_name, depth)
        fam = backbones.family_of(model_name)
        if fam == "resnet":
            self.ops, hooks, self.chans, self.relu_typed = _build_resnet(model, self.targets)
        elif fam in ("vgg", "alexnet", "squeezenet"):
            self.ops, hooks, self.chans, self.relu_typed = _build_sequential(model.features, self.targets)
        else:
            raise NotImplementedError("NativeEngine builds Conv / MaxPool / Concat graphs; the %s family (pre-activation "
                                      "BN, average pooling) is engine_densenet.DenseNetEngine — engines.make_engine picks "
                                      "it" % fam)
        # hook buffers in reference order = forward execution order of the target modules
        self.hook_bufs = [hooks[id(t)] for t in self.targets]
        order = {op.y: i for i, op in enumerate(self.ops)}
        self.hook_bufs.sort(key=lambda b: order[b])
        for b in self.hook_bufs:
            if b not in self.relu_typed:
                raise NotImplementedError("hooked buffer %s is not a ReLU output" % b)
        if use_tensor_cores is None:
            use_tensor_cores = os.environ.get("I2V_NATIVE_TC", "1") != "0"
        self.use_tc = bool(use_tensor_cores)
        self.use_stem = os.environ.get("I2V_NATIVE_STEM", "1") != "0"   # dedicated first-layer kernels
        self.use_bits = os.environ.get("I2V_NATIVE_BITS", "1") != "0"   # ReLU-backward masks as bits (TMA epilogue)
        self.use_stem_tc = os.environ.get("I2V_NATIVE_STEM_TC", "1") != "0"   # first-layer dgrad as tcgen05 GEMM + col2im
        self._zbuf = None
        # first-layer data gradient without the Z^T scratch (i2v_conv_stem_dgrad_direct_f32); $I2V_STEM_DGRAD_DIRECT=0: GEMM + col2im
        self.stem_dgrad_direct = os.environ.get("I2V_STEM_DGRAD_DIRECT", "1") != "0"
        # max-pool backward fused into that kernel (i2v_conv_stem_dgrad_pool_f32): the stem activation's gradient never reaches
        # HBM; $I2V_STEM_DGRAD_POOL=0: i2v_maxpool_bwd_f32 + i2v_conv_stem_dgrad_direct_f32
        self.stem_dgrad_pool = os.environ.get("I2V_STEM_DGRAD_POOL", "1") != "0"
        # max pooling fused into the first-layer forward's epilogue (i2v_conv_stem_fwd_pool_f32): the stem activation is never
        # written.  Bit-identical, but OFF by default on numbers: the pooling runs in the four epilogue warps, in series with the
        # TMEM drain of the next row, and the fused kernel takes 970 us per 256 frames against 441 + 264 us for the two launches
        # (profiles/r02_stem_fwd_pool_fused.txt); $I2V_STEM_FWD_POOL=1 enables it
        self.stem_fwd_pool = os.environ.get("I2V_STEM_FWD_POOL", "0") == "1"
        # first-layer forward without the im2col patch matrix, one output row per tile (i2v_conv_stem_fwd_rows_f32); =0: im2col + GEMM
        self.stem_fwd_rows = os.environ.get("I2V_STEM_FWD_ROWS", "1") != "0"
        # EXPERIMENTAL: first-layer forward without the im2col patch matrix (i2v_conv_stem_fwd_direct_f32)
        self.stem_direct = os.environ.get("I2V_STEM_DIRECT", "0") == "1"
        self._xpbuf = None
        self._cache = {}
        self.buffer_generation = 0
        # A bottleneck's downsample branch and its last 1x1 convolution are summed anyway: run them as ONE GEMM over the
        # concatenated K (i2v_conv_tc_dual_f32) — the downsample output is never written nor read back as a residual.
        # $I2V_FUSE_DS=0: two launches.
        self.fuse_ds = os.environ.get("I2V_FUSE_DS", "1") != "0"
        self.dual = {}                       # name of the last convolution -> (downsample op, hi, lo, rna, bias)
        for op in self.ops:
            if op.kind != "conv" or op.residual is None or op.R != 1 or op.stride != 1 or op.pad != 0 or not op.relu:
                continue
            r = op.residual
            prod = [o for o in self.ops if o.kind == "conv" and o.y == r]
            users = [o for o in self.ops if (o.kind == "conv" and (o.x == r or o.residual == r)) or (o.kind == "pool" and o.x == r)
                     or (o.kind not in ("conv", "pool") and r in o.xs)]
            if len(prod) != 1 or users != [op] or r in self.hook_bufs or r in self.relu_typed:
                continue
            ds = prod[0]
            if ds.relu or ds.residual is not None or ds.R != 1 or ds.pad != 0 or ds.x_nchw or ds.cout != op.cout \
                    or ds.cin % 32 or op.cin % 32:
                continue
            self.dual[op.name] = (ds, torch.cat([ds.tc_fwd[0], op.tc_fwd[0]], 1).contiguous(),
                                  torch.cat([ds.tc_fwd[1], op.tc_fwd[1]], 1).contiguous(),
                                  torch.cat([ds.tc_fwd[2], op.tc_fwd[2]], 1).contiguous(), (ds.bias + op.bias).contiguous())

    @property
    def num_layers(self):
        return len(self.hook_bufs)

    def bytes_per_frame(self, h, w):
        """Activation + gradient + mask bytes one frame needs in a chunk's buffer plan (see _plan)."""
        dims, total = {"img": (h, w)}, 3 * h * w * 4
        for op in self.ops:
            if op.kind == "conv":
                dims[op.y] = op.out_hw(*dims[op.x])
            elif op.kind == "pool":
                dims[op.y] = op.out_hw(*dims[op.x])
            else:
                dims[op.y] = dims[op.xs[0]]
            oh, ow = dims[op.y]
            elems = oh * ow * self.chans[op.y]
            total += elems * 4 * (1 if op.y in self.hook_bufs else 2) + (elems if op.kind == "pool" else 0) + elems // 8
        return total

    def frames_per_chunk(self, h, w, n_frames, device, share=1.0):
        """Frames per forward/backward sub-batch.  Large chunks win (measured on B200, ResNet-50 layer2 at 224^2:
        8.8k / 11.3k / 12.2k / 12.9k frame-steps/s at 32 / 64 / 128 / 256 frames): every launch has ~15 us of fixed
        cost (prologue, pipeline ramp, last-wave imbalance) and the persistent kernels' tile counts quantise against
        the 148 SMs; the limit is memory — at most `share` x 60 % of what is free now, and 256 frames."""
        free, _ = torch.cuda.mem_get_info(device)
        fit = int(free * 0.6 * share // max(1, self.bytes_per_frame(h, w)))
        return max(1, min(n_frames, 256, fit))

    # ---- per-(n,h,w) buffer plan -------------------------------------------------------------------
    def _plan(self, n, h, w, device):
        key = (n, h, w)
        plan = self._cache.get(key)
        if plan is not None:
            return plan
        dims = {"img": (h, w)}
        acts, grads, argmax, descs, bits = {}, {}, {}, {}, {}
        fused_ds = {}                        # downsample op name -> name of the convolution that absorbs it in this plan
        for op in self.ops:
            if op.kind == "conv":
                ih, iw = dims[op.x]
                oh, ow = op.out_hw(ih, iw)
                dims[op.y] = (oh, ow)
                d = capi.ConvDesc(n, ih, iw, op.cin, op.cout, op.R, op.R, op.stride, op.pad, oh, ow)
                descs[op.name] = d
                if op.name in self.dual and self.fuse_ds and self.use_tc:
                    ds = self.dual[op.name][0]
                    if capi.conv_tc_supported(d, 0) and capi.conv_tc_supported(descs[ds.name], 0) and dims[ds.y] == (oh, ow):
                        fused_ds[ds.name] = op.name
                # ReLU outputs produced by the tensor-core kernel also leave their activity as BITS ([C/32, M] words):
                # the data gradients that need 1[activation > 0] then read 1/32 of the bytes (and run the TMA epilogue)
                if op.relu and self.use_bits and self.use_tc and not op.x_nchw and capi.conv_tc_supported(d, 0):
                    bits[op.y] = torch.empty(op.cout // 32, n * oh * ow, device=device, dtype=torch.int32)
            elif op.kind == "pool":
                ih, iw = dims[op.x]
                dims[op.y] = op.out_hw(ih, iw)
            else:
                dims[op.y] = dims[op.xs[0]]
            oh, ow = dims[op.y]
            acts[op.y] = torch.empty(n, oh, ow, self.chans[op.y], device=device, dtype=torch.float32)
            if op.kind == "pool":
                argmax[op.y] = torch.empty(n, oh, ow, self.chans[op.y], device=device, dtype=torch.uint8)
        for op in self.ops:                  # an absorbed downsample output is never materialised (nor is its gradient: the
            if op.kind == "conv" and op.name in fused_ds:        # block output's gradient IS the shortcut's)
                del acts[op.y]
        for name, t in acts.items():
            if name not in self.hook_bufs:
                grads[name] = torch.empty_like(t)
        plan = dict(dims=dims, acts=acts, grads=grads, argmax=argmax, descs=descs, bits=bits, fused_ds=fused_ds,
                    gimg=torch.empty(n, 3, h, w, device=device, dtype=torch.float32))
        if len(self._cache) > 4:
            self._cache.clear()
            self.buffer_generation += 1      # captured CUDA graphs that point into the old buffers are stale
        self._cache[key] = plan
        return plan

    # ---- forward -------------------------------------------------------------------------------------
    def _conv_fwd(self, op, d, x, y, residual, bits_out=None):
        if (op.x_nchw and residual is None and self.use_tc and self.use_stem_tc and self.tf32x3 and self.stem_fwd_rows
                and op.tc_stem_fwd is not None and op.tc_stem_fwd[0].shape == (64, 160) and capi.conv_stem_fwd_rows_supported(d)):
            hi, lo, _ = op.tc_stem_fwd
            capi.conv_stem_fwd_rows(d, x, hi, lo, op.bias, y, relu=op.relu)
        elif (op.x_nchw and residual is None and self.use_tc and self.use_stem_tc and self.tf32x3 and self.stem_direct
                and op.tc_stem_direct is not None and capi.conv_stem_fwd_direct_supported(d)):
            hi, lo, _ = op.tc_stem_direct
            nfl = capi.stem_fwd_direct_scratch_floats(d)
            if self._xpbuf is None or self._xpbuf.numel() < nfl:
                self._xpbuf = torch.empty(nfl, device=x.device, dtype=torch.float32)
                self.buffer_generation += 1
            capi.conv_stem_fwd_direct(d, x, hi, lo, op.bias, self._xpbuf, y, relu=op.relu)
        elif op.x_nchw and residual is None and self.use_tc and self.use_stem_tc and op.tc_stem_fwd is not None:
            hi, lo, rna = op.tc_stem_fwd
            nfl = capi.stem_fwd_tc_scratch_floats(d)
            if self._zbuf is None or self._zbuf.numel() < nfl:
                self._zbuf = torch.empty(nfl, device=x.device, dtype=torch.float32)
                self.buffer_generation += 1
            capi.conv_stem_fwd_tc(d, x, hi if self.tf32x3 else rna, lo if self.tf32x3 else None, op.bias, self._zbuf, y,
                                  relu=op.relu)
        elif op.x_nchw and residual is None and self.use_stem and capi.conv_stem_supported(d):
            capi.conv_stem_fwd(d, x, op.w_stem, op.bias, y, relu=op.relu)
        elif self.use_tc and not op.x_nchw and capi.conv_tc_supported(d, 0):
            hi, lo, rna = op.tc_fwd
            capi.conv_tc(d, 0, x, hi if self.tf32x3 else rna, lo if self.tf32x3 else None, op.bias, residual, None, y,
                         relu=op.relu, mask_bits=bits_out)
        else:
            capi.conv_fwd_simt(d, x, op.b_fwd, op.bias, residual, y, relu=op.relu, x_nchw=op.x_nchw)

    def _stem_pool_fusable(self, pool, plan):
        """True when `pool` (a max pooling whose gradient is about to be taken) and the convolution that produced its input
        can run as ONE kernel (i2v_conv_stem_dgrad_pool_f32): 3x3 / stride 2 / pad 1 over the ReLU output of a 7x7 / stride-2
        first layer that nothing else consumes."""
        if not (self.stem_dgrad_pool and self.stem_dgrad_direct and self.use_tc and self.use_stem_tc and self.tf32x3):
            return False
        if (pool.k, pool.stride, pool.pad) != (3, 2, 1) or pool.ceil or pool.x not in self.relu_typed or pool.x in self.hook_bufs:
            return False
        prod = [o for o in self.ops if o.kind == "conv" and o.y == pool.x]
        users = [o for o in self.ops if (o.kind == "conv" and (o.x == pool.x or o.residual == pool.x))
                 or (o.kind == "pool" and o.x == pool.x) or (o.kind not in ("conv", "pool") and pool.x in o.xs)]
        if len(prod) != 1 or users != [pool]:
            return False
        conv = prod[0]
        if not (conv.x_nchw and conv.residual is None and getattr(conv, "tc_stem_dgrad_direct", None) is not None):
            return False
        n, P2, Q2, c = plan["acts"][pool.y].shape
        return c == 64 and capi.conv_stem_dgrad_pool_supported(plan["descs"][conv.name], P2, Q2)

    def _stem_pool_fwd(self, plan):
        """{stem conv name: pool op} when the first layer and the max pooling behind it run as ONE forward kernel
        (i2v_conv_stem_fwd_pool_f32): the conditions of `_stem_pool_fusable` for the backward fusion — which must be on, the stem
        activation does not exist for a separate pooling backward's producer — plus the row-tile forward kernel's."""
        if "stem_pool_fwd" in plan:
            return plan["stem_pool_fwd"]
        out = {}
        if self.stem_fwd_pool and self.stem_fwd_rows:
            for pool in self.ops:
                if pool.kind != "pool" or not self._stem_pool_fusable(pool, plan):
                    continue
                conv = [o for o in self.ops if o.kind == "conv" and o.y == pool.x][0]
                n, P2, Q2, c = plan["acts"][pool.y].shape
                if (conv.relu and conv.tc_stem_fwd is not None and conv.tc_stem_fwd[0].shape == (64, 160)
                        and capi.conv_stem_fwd_pool_supported(plan["descs"][conv.name], P2, Q2)):
                    out[conv.name] = pool
        plan["stem_pool_fwd"] = out
        return out

    def _conv_dgrad(self, op, d, dy, addend, mask_src, dx, mask_bits=None, pooled=None):
        if pooled is not None:
            hi, lo, _ = op.tc_stem_dgrad_direct
            capi.conv_stem_dgrad_pool(d, pooled[0], pooled[1], hi, lo, dx)
        elif (op.x_nchw and addend is None and mask_src is None and self.use_tc and self.use_stem_tc and self.tf32x3
                and getattr(op, "tc_stem_dgrad_direct", None) is not None and self.stem_dgrad_direct
                and capi.conv_stem_dgrad_direct_supported(d)):
            hi, lo, _ = op.tc_stem_dgrad_direct
            capi.conv_stem_dgrad_direct(d, dy, hi, lo, dx)
        elif (op.x_nchw and addend is None and mask_src is None and self.use_tc and self.use_stem_tc
                and op.tc_stem_dgrad is not None and (d.N * d.P * d.Q) % 4 == 0):
            hi, lo, rna = op.tc_stem_dgrad
            nfl = capi.stem_dgrad_tc_scratch_floats(d)
            if self._zbuf is None or self._zbuf.numel() < nfl:
                self._zbuf = torch.empty(nfl, device=dy.device, dtype=torch.float32)
                self.buffer_generation += 1
            capi.conv_stem_dgrad_tc(d, dy, hi if self.tf32x3 else rna, lo if self.tf32x3 else None, self._zbuf, dx)
        elif op.x_nchw and addend is None and mask_src is None and self.use_stem and capi.conv_stem_supported(d):
            capi.conv_stem_dgrad(d, dy, op.b_fwd, dx)
        elif self.use_tc and not op.x_nchw and op.stride == 1 and capi.conv_tc_supported(d, 1):
            hi, lo, rna = op.tc_dgrad
            if mask_src is not None and mask_bits is not None:
                mask_src = None                                   # same mask, 1/32 of the bytes
            else:
                mask_bits = None
            capi.conv_tc(d, 1, dy, hi if self.tf32x3 else rna, lo if self.tf32x3 else None, None, addend, mask_src, dx,
                         mask_bits=mask_bits)
        elif (self.use_tc and not op.x_nchw and op.stride > 1 and capi.conv_tc_supported(d, 1)
              and (addend is None or addend is dx or all(v is not None for v in op.tc_dgrad_cls.values()))):
            # strided: one dense launch per stride-parity class; classes without taps keep what is there
            if addend is None and any(v is None for v in op.tc_dgrad_cls.values()):
                dx.zero_()
            if mask_src is not None and mask_bits is not None:
                mask_src = None                                   # same mask, 1/32 of the bytes — and the TMA epilogue
            else:
                mask_bits = None
            for (ph, pw), wt in op.tc_dgrad_cls.items():
                if wt is None:
                    continue
                hi, lo, rna = wt
                capi.conv_tc_dgrad_class(d, ph, pw, dy, hi if self.tf32x3 else rna, lo if self.tf32x3 else None,
                                         addend, mask_src, dx, mask_bits=mask_bits)
        else:
            capi.conv_dgrad_simt(d, dy, op.b_dgrad, addend, mask_src, dx, x_nchw=op.x_nchw)

    def features(self, img, need_grad, clone=True):
        """Hooked feature maps of `img` (NHWC views of this engine's reusable buffers when need_grad, or when the caller
        passes clone=False and copies them out itself before the next call)."""
        n, c, h, w = img.shape
        if c != 3 or not img.is_contiguous():
            raise ValueError("expected a contiguous [n,3,H,W] image batch")
        plan = self._plan(n, h, w, img.device)
        acts = plan["acts"]
        fused_ds = plan["fused_ds"]
        stem_pool = self._stem_pool_fwd(plan)
        pooled_by_stem = set()
        self._stem_act_stale = None
        for op in self.ops:
            if op.kind == "conv":
                if op.name in fused_ds:
                    continue                                       # runs inside the block's last convolution
                x = img if op.x == "img" else acts[op.x]
                if op.name in stem_pool:
                    pool = stem_pool[op.name]
                    hi, lo, _ = op.tc_stem_fwd
                    capi.conv_stem_fwd_pool(plan["descs"][op.name], x, hi, lo, op.bias, acts[pool.y], plan["argmax"][pool.y],
                                            relu=True, mark_dead=True)
                    pooled_by_stem.add(id(pool))
                    self._stem_act_stale = (op, x)                 # acts[op.y] was not written (relu_masks recomputes it on demand)
                    continue
                bits_out = plan["bits"].get(op.y) if need_grad else None
                if op.name in self.dual and self.dual[op.name][0].name in fused_ds:
                    ds, hi, lo, rna, bias = self.dual[op.name]
                    capi.conv_tc_dual(plan["descs"][ds.name], acts[ds.x], acts[op.x], hi if self.tf32x3 else rna,
                                      lo if self.tf32x3 else None, bias, acts[op.y], relu=True, mask_bits=bits_out)
                    continue
                self._conv_fwd(op, plan["descs"][op.name], x, acts[op.y], acts[op.residual] if op.residual else None,
                               bits_out)
            elif op.kind == "pool":
                if id(op) in pooled_by_stem:
                    continue                                       # ran inside the first layer's epilogue
                # a ReLU output is pooled: windows with nothing > 0 are marked "no winner", which IS the ReLU-backward mask
                capi.maxpool_fwd(acts[op.x], acts[op.y], plan["argmax"][op.y], op.k, op.stride, op.pad,
                                 mark_dead=op.x in self.relu_typed)
            else:
                off = 0
                for xn in op.xs:
                    capi.copy_channels(acts[xn], acts[op.y], 0, off, self.chans[xn])
                    off += self.chans[xn]
        self._last = plan if need_grad else None
        self._last_fwd = plan
        feats = [acts[b] for b in self.hook_bufs]
        # clean features are kept by the caller across the whole attack: hand out copies, the plan's buffers are reused
        return [f.clone() for f in feats] if (not need_grad and clone) else feats

    def relu_masks(self):
        """Activity masks 1[activation > 0] of every ReLU of the last forward, in the torch module's ReLU call
        order, as [n,C,h,w] bool tensors — the decisions the backward pass uses; None for a ReLU the truncated
        graph does not execute.  For parity tests only (forces a sync)."""
        acts = self._last_fwd["acts"]
        if getattr(self, "_stem_act_stale", None) is not None:      # the fused forward never wrote the stem activation
            op, x = self._stem_act_stale
            hi, lo, _ = op.tc_stem_fwd
            capi.conv_stem_fwd_rows(self._last_fwd["descs"][op.name], x, hi, lo, op.bias, acts[op.y], relu=op.relu)
            self._stem_act_stale = None
        out = []
        for op in self.ops:
            if op.kind == "conv" and op.relu:
                out.extend([None] * op.relu_skipped_before)
                out.append((acts[op.y] > 0).permute(0, 3, 1, 2).contiguous().cpu())
        return out

    def pool_indices(self):
        """Winners of every max pooling of the last forward, in the torch module's MaxPool2d call order, as int64
        [n,C,P,Q] flat indices h*W + w into the pooled plane (torch's `return_indices` convention) — the decisions
        the backward pass routes gradients by.  For parity tests only (forces a sync)."""
        plan = self._last_fwd
        out = []
        for op in self.ops:
            if op.kind != "pool":
                continue
            am = plan["argmax"][op.y].to(torch.int64)                       # [n,P,Q,C]: r*k + s of the first maximum
            n, P, Q, C = am.shape
            ih, iw = plan["dims"][op.x]
            p = torch.arange(P, device=am.device).view(1, P, 1, 1) * op.stride - op.pad
            q = torch.arange(Q, device=am.device).view(1, 1, Q, 1) * op.stride - op.pad
            # dead windows (mark_dead: nothing > 0, argmax = 255) pass no gradient whichever element is named the winner:
            # report the first valid element of the window, which is what torch picks among equal zeros
            r0 = (-p).clamp(min=0)
            s0 = (-q).clamp(min=0)
            am = torch.where(am == 255, r0 * op.k + s0, am)
            flat = (p + am // op.k) * iw + (q + am % op.k)
            out.append(flat.permute(0, 3, 1, 2).contiguous().cpu())
        return out

    # ---- backward ------------------------------------------------------------------------------------
    def input_grad(self, grads, out=None):
        """dcost/dimg [n,3,H,W] given dcost/dfeat of every hooked layer; written into `out` when given (the attack loop
        passes its slice of the batch-wide gradient, so no copy follows), else into this engine's own buffer."""
        plan = self._last
        if plan is None:
            raise RuntimeError("input_grad() needs a preceding features(..., need_grad=True)")
        gimg = plan["gimg"] if out is None else out
        if gimg.shape != plan["gimg"].shape or not gimg.is_contiguous():
            raise ValueError("out must be a contiguous %s tensor" % (tuple(plan["gimg"].shape),))
        acts = plan["acts"]
        G = dict(plan["grads"])
        ready = set()
        for b, g in zip(self.hook_bufs, grads):
            G[b] = g.view_as(acts[b])
            ready.add(b)
        pending = {}          # buffer -> gradient tensor of an identity (residual) contribution not yet merged
        pooled = {}           # buffer -> (gradient of its pooled map, argmax plane): pooling backward deferred to the producer's dgrad
        last = self.hook_bufs[-1]
        started = False
        for op in reversed(self.ops):
            if not started:
                if op.y != last:
                    continue
                started = True
            if op.y not in ready:
                if op.y in pending:      # only an identity contribution reached this buffer
                    raise RuntimeError("internal: unmerged residual gradient for %s" % op.y)
                continue                 # dead branch (not upstream of any hook)
            gy = G[op.y]
            if op.kind == "conv":
                if op.residual is not None:
                    r = op.residual
                    if r in self.relu_typed:
                        pending[r] = gy                   # merged (and masked by 1[r>0]) by the dgrad that writes G[r]
                    else:
                        G[r] = gy                         # pre-activation shortcut (downsample output): same gradient
                        ready.add(r)
                if op.x == "img":
                    dx, mask = gimg, None
                else:
                    dx, mask = G[op.x], (acts[op.x] if op.x in self.relu_typed else None)
                addend = None
                if op.x in ready:
                    addend = dx
                    if op.x in pending:
                        raise RuntimeError("internal: two addends for %s" % op.x)
                elif op.x in pending:
                    addend = pending.pop(op.x)
                self._conv_dgrad(op, plan["descs"][op.name], gy, addend, mask, dx, plan["bits"].get(op.x),
                                 pooled=pooled.pop(op.y, None))
                ready.add(op.x)
            elif op.kind == "pool":
                # ReLU-backward mask of the pooled tensor's producer: 1[x[argmax] > 0] = 1[y > 0] was folded into the argmax
                # plane by the forward pass (mark_dead), so no mask tensor is read here (the stem activation, 4x the pooled
                # bytes, is never touched in the backward pass).
                if op.x in pending:
                    raise NotImplementedError("pooling input that is also a residual source")
                if op.x not in ready and self._stem_pool_fusable(op, plan):
                    pooled[op.x] = (gy, plan["argmax"][op.y])    # the first layer's data-gradient kernel does the pooling too
                    ready.add(op.x)
                    continue
                capi.maxpool_bwd(gy, plan["argmax"][op.y], None, G[op.x], op.k, op.stride, op.pad,
                                 accumulate=op.x in ready)       # hooked input: K1 wrote its gradient first
                ready.add(op.x)
            else:
                off = 0
                for xn in op.xs:
                    if xn in ready:
                        capi.copy_channels(gy, G[xn], off, 0, self.chans[xn], accumulate=True)
                    else:
                        capi.copy_channels(gy, G[xn], off, 0, self.chans[xn])
                        ready.add(xn)
                    off += self.chans[xn]
        self._last = None
        return gimg
