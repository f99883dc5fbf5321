"""The per-step attack loop shared by I2V, ENS-I2V and adaptive ENS-I2V.

One function, `run_image_guided`, restates the common skeleton of reference
image_attacks.py:294-364 (I2V), 426-496 (ENS-I2V) and TPAMI_attack.py:223-320 (AENS-I2V) around this
repo's kernels:

    setup    x = denorm(frames)                       K3  i2v_denorm_f32          (308)
             modifier = 0.01/255, Adam m = v = 0                                   (304-306)
             clean features of every hooked layer     engine.features             (318-323)
             true_image = compose_norm(x, modifier)   K3  i2v_compose_norm_f32    (331-332)
    step     [AENS] coeffs <- softmax(softmax(prev)+momentum*coeffs)   K2         (TPAMI 265)
             features of true_image                   engine.features             (334)
             per layer: cos[n], dcos/dfeat * w_l      K1  i2v_cosine_loss_grad    (341-347 + autograd)
             dcost/dtrue_image                        engine.input_grad           (352)
             cost / layer sums / prev                 K2  i2v_layer_sums_f32      (347; TPAMI 289-297)
             Adam + next true_image                   K3a i2v_adam_compose_table  (351-353, 331-332)
    finish   the last true_image IS the returned adversarial clip                  (360-364)

Nothing in the loop synchronises with the host: the cost of every step and (AENS) the coefficient
vectors go to device logs that are copied once after the loop (the reference syncs 2-5 times per
step: print(cost), .cpu(), .item()).
"""
import os

import numpy as np
import torch

from . import capi

INIT_MODIFIER = 0.01 / 255   # image_attacks.py:304
BETA1, BETA2, ADAM_EPS = 0.9, 0.999, 1e-8   # torch.optim.Adam defaults (image_attacks.py:306)


class LoopResult:
    __slots__ = ("adv", "cost", "weights", "elapsed_ms")

    def __init__(self, adv, cost, weights, elapsed_ms):
        self.adv = adv
        self.cost = cost
        self.weights = weights
        self.elapsed_ms = elapsed_ms


def frames_of(videos):
    """[b,c,f,h,w] -> contiguous [b*f,c,h,w] (image_attacks.py:300-301)."""
    b, c, f, h, w = videos.shape
    return videos.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w).contiguous()


def clip_of(frames, b, f):
    """[b*f,c,h,w] -> the reference's returned view [b,c,f,h,w] (image_attacks.py:362-363)."""
    n, c, h, w = frames.shape
    return frames.reshape(b, f, c, h, w).permute(0, 2, 1, 3, 4)


def _chunk_frames(engines, N, chunk, h=None, w=None, device=None):
    """Frames per forward/backward sub-batch: explicit argument > $I2V_CHUNK > what the engines ask for given the
    image size and the free device memory (engines that do not size themselves report `preferred_chunk`)."""
    if chunk is None:
        chunk = int(os.environ.get("I2V_CHUNK", "0"))
    if not chunk and not engines:
        chunk = N
    if not chunk:
        asks = []
        for e in engines:
            if h is not None and hasattr(e, "frames_per_chunk"):
                asks.append(e.frames_per_chunk(h, w, N, device, share=1.0 / len(engines)))
            else:
                asks.append(getattr(e, "preferred_chunk", 128))
        chunk = min(asks)
    return max(1, min(int(chunk), N))


class ImageGuidedRun:
    """One attack call, split into setup() / step() / finish() so that a benchmark can time exactly K
    steps; `run_image_guided` is the plain composition the attack classes use.

    adaptive=False: cost = sum of all cosines (I2V / ENS-I2V).
    adaptive=True : AENS-I2V; `coeffs` is the persistent [L] device tensor (updated in place).
    chunk      : frames per forward/backward sub-batch (bounds activation memory and keeps one
                 chunk's activations L2-resident); frames are independent, so the result does not
                 depend on it.  Default: $I2V_CHUNK or the engines' preferred size.
    reduce_hook: optional object with .cos_rows(cos) / .grad(g) for multi-GPU runs (dist.py).
    tap        : optional callable(step, dict) for tests — forces a host sync per step when given.
    """

    def __init__(self, engines, epsilon, steps, step_size, adaptive=False, coeffs=None, momentum=0.0,
                 coef_CE=False, chunk=None, reduce_hook=None, tap=None, layer_offsets=None, n_layers_total=None):
        """layer_offsets / n_layers_total: one-backbone-per-GPU ensembles (dist.ensemble_plan) — `engines` are then only
        this rank's members, layer_offsets[i] is the row of engine i's first hooked layer in the ensemble-wide
        [L, N] cosine table and reduce_hook sums the rows and the input gradient over the ensemble group."""
        self.engines = list(engines)
        self.epsilon = float(epsilon)
        self.steps = int(steps)
        self.step_size = float(step_size)
        self.adaptive = adaptive
        self.coeffs = coeffs
        self.momentum = float(momentum)
        self.coef_CE = bool(coef_CE)
        self.chunk_request = chunk
        self.reduce_hook = reduce_hook
        self.tap = tap
        self.n_layers = sum(e.num_layers for e in self.engines) if n_layers_total is None else int(n_layers_total)
        if layer_offsets is None:
            layer_offsets, off = [], 0
            for e in self.engines:
                layer_offsets.append(off)
                off += e.num_layers
        self.layer_offsets = list(layer_offsets)
        self.owned_rows = sorted(o + k for o, e in zip(self.layer_offsets, self.engines) for k in range(e.num_layers))
        # CUDA-graph replay of the step (see step()): needs engines whose step is a fixed launch sequence on fixed
        # buffers (the native ones), no per-step host work (tap) and no collective inside the step; $I2V_GRAPH=0 disables
        self.use_graph = (os.environ.get("I2V_GRAPH", "1") != "0" and tap is None and reduce_hook is None
                          and bool(self.engines) and all(getattr(e, "graph_safe", False) for e in self.engines))
        self._graph = None
        self._graph_launches = {}
        self._graph_gen = None
        self._alloc_key = None
        self.reusable = False        # set by run_image_guided(cache=...): finish() then hands out a copy of the result
        if adaptive and (coeffs is None or coeffs.numel() != self.n_layers):
            raise ValueError("adaptive mode needs a coeffs tensor with one entry per hooked layer (%d)" % self.n_layers)
        self.step_no = 0

    def setup(self, videos):
        """Start an attack call on `videos`.  A run object can be set up again for another batch of the same shape: its
        device state and — the point — its captured CUDA graph are then reused, so a sweep of batch-size-1 calls
        (image_main.py:82-89) pays for capture and instantiation once, not per clip."""
        if videos.dim() != 5 or videos.shape[1] != 3:
            raise ValueError("videos must be [b,3,f,h,w], got %s" % (tuple(videos.shape),))
        device = torch.device("cuda", torch.cuda.current_device())
        capi.device_check(device)
        b, c, f, h, w = videos.shape
        self.b, self.f = b, f
        frames = frames_of(videos.to(device=device, dtype=torch.float32, non_blocking=True))
        N = self.N = b * f
        inner = self.inner = h * w
        steps = self.steps
        key = (b, f, h, w, device.index)
        reuse = self._alloc_key == key
        if not reuse:
            self._graph = None
            chunk = self.chunk = _chunk_frames(self.engines, N, self.chunk_request, h, w, device)
            self.spans = [(s, min(s + chunk, N)) for s in range(0, N, chunk)]
            self.x = torch.empty_like(frames)
            self.mod = torch.empty_like(frames)
            self.m = torch.empty_like(frames)
            self.v = torch.empty_like(frames)
            self.true_img = torch.empty_like(frames)
            self.g_total = torch.empty_like(frames)
            self.init_feats = None
        chunk = self.chunk
        capi.denorm(frames, self.x, inner)                                      # image_attacks.py:308
        capi.fill(self.mod, INIT_MODIFIER)                                      # image_attacks.py:304
        self.m.zero_()
        self.v.zero_()

        # clean features (image_attacks.py:318-323; TPAMI_attack.py:241-253), kept for all N frames
        init_feats = []
        for ei, e in enumerate(self.engines):
            per_layer = self.init_feats[ei] if reuse else None
            for (s0, s1) in self.spans:
                # native engines hand out views of their reusable buffers (clone=False): one copy into the per-call store
                fe = e.features(frames[s0:s1], need_grad=False, clone=False) if hasattr(e, "buffer_generation") \
                    else e.features(frames[s0:s1], need_grad=False)
                if per_layer is None:
                    per_layer = [torch.empty((N,) + tuple(t.shape[1:]), device=device, dtype=torch.float32) for t in fe]
                for dst, t in zip(per_layer, fe):
                    dst[s0:s1].copy_(t)
            init_feats.append(per_layer)
        self.init_feats = init_feats
        if not reuse:
            self.grads = [[torch.empty((chunk,) + tuple(t.shape[1:]), device=device, dtype=torch.float32) for t in fe]
                          for fe in self.init_feats]
            self.cos = torch.zeros(self.n_layers, N, device=device, dtype=torch.float32)
            foreign = [r for r in range(self.n_layers) if r not in set(self.owned_rows)]
            self.foreign_rows = torch.tensor(foreign, device=device, dtype=torch.long) if foreign else None
            self.step_idx = torch.zeros(1, device=device, dtype=torch.int32)
            self.cost_log = torch.zeros(max(steps, 1), device=device, dtype=torch.float32)
            self.table = capi.adam_step_table(steps, self.step_size, BETA1, BETA2).to(device)
            if self.adaptive:
                self.prev = torch.ones(self.n_layers, device=device, dtype=torch.float32)   # TPAMI_attack.py:257
                self.w_out = torch.empty(self.n_layers, device=device, dtype=torch.float32)
                self.weights_log = torch.zeros(max(steps, 1), self.n_layers, device=device, dtype=torch.float32)
            else:
                self.prev = self.w_out = self.weights_log = None
        else:
            self.cos.zero_()
            self.step_idx.zero_()
            self.cost_log.zero_()
            if self.adaptive:
                self.prev.fill_(1.0)
                self.weights_log.zero_()
        self._alloc_key = key
        capi.compose_norm(self.x, self.mod, self.true_img, self.epsilon, inner)   # image_attacks.py:331-332
        self.step_no = 0
        return self

    def step(self):
        """One attack step.  Step 0 runs eagerly (the engines size their buffers on first use); when nothing needs the
        host inside the step (no tap, no collective, no per-kernel profiling) step 1 is captured into a CUDA graph —
        every launch of the step, all chunks — and steps 1.. are replays of it: the device step counter and the Adam
        scalar table (K3a) make the step body iteration-invariant."""
        if self.step_no >= self.steps:
            raise RuntimeError("all %d steps of this run are done" % self.steps)
        if self._graph is not None and self._graph_gen != self._buffer_gen():
            self._graph = None                   # an engine re-allocated its buffers since the capture
        if self._graph is not None and capi.PROFILE_EVENTS is None:
            self._graph.replay()
            for k, v in self._graph_launches.items():
                capi.LAUNCHES[k] = capi.LAUNCHES.get(k, 0) + v
        elif (self.use_graph and self._graph is None and self.step_no >= 1 and self.steps - self.step_no >= 2
              and capi.PROFILE_EVENTS is None):
            before = dict(capi.LAUNCHES)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                self._step_body()
            self._graph_launches = {k: v - before.get(k, 0) for k, v in capi.LAUNCHES.items() if v != before.get(k, 0)}
            self._graph = graph
            self._graph_gen = self._buffer_gen()
            graph.replay()                       # capture only records: this replay IS the step
        else:
            self._step_body()
        self.step_no += 1

    def _buffer_gen(self):
        return tuple(getattr(e, "buffer_generation", 0) for e in self.engines)

    def _step_body(self):
        adaptive = self.adaptive
        if adaptive:
            capi.layer_reweight(self.coeffs, self.prev, self.momentum, self.w_out, self.weights_log, self.step_idx)
        if not self.engines:                                   # an idle rank of an ensemble group contributes zeros
            self.g_total.zero_()
        for (s0, s1) in self.spans:
            for ei, (e, f0, gr) in enumerate(zip(self.engines, self.init_feats, self.grads)):
                layer = self.layer_offsets[ei]
                feats = e.features(self.true_img[s0:s1], need_grad=True)
                gviews = []
                for a, a0, ga in zip(feats, f0, gr):
                    gv = ga[:s1 - s0]
                    capi.cosine_loss_grad(a, a0[s0:s1], gv, self.cos[layer, s0:s1],
                                          w_dev=self.w_out[layer:layer + 1] if adaptive else None, w_host=1.0,
                                          relu_mask=e.relu_masked_grads)
                    gviews.append(gv)
                    layer += 1
                if ei == 0 and getattr(e, "writes_input_grad_in_place", False):
                    e.input_grad(gviews, out=self.g_total[s0:s1])      # the first-layer data gradient lands in place
                    continue
                g = e.input_grad(gviews)
                if ei == 0:
                    self.g_total[s0:s1].copy_(g)
                else:
                    self.g_total[s0:s1].add_(g)
        if self.reduce_hook is not None:
            if self.foreign_rows is not None:                  # rows of peers still hold last step's sums
                self.cos.index_fill_(0, self.foreign_rows, 0.0)
            self.reduce_hook.grad(self.g_total)
            self.reduce_hook.cos_rows(self.cos)
        capi.layer_sums(self.cos, self.coeffs if adaptive else None, self.prev, self.cost_log, self.step_idx,
                        mode=1 if adaptive else 0, coef_CE=self.coef_CE)
        if self.tap is not None:
            # relu_masks: the activity decisions of this step's forward (valid when all frames fit one chunk)
            masks = [e.relu_masks() if hasattr(e, "relu_masks") and len(self.spans) == 1 else None for e in self.engines]
            pools = [e.pool_indices() if hasattr(e, "pool_indices") and len(self.spans) == 1 else None for e in self.engines]
            self.tap(self.step_no, dict(g=self.g_total.clone(), cos=self.cos.clone(), mod_before=self.mod.clone(),
                                        m_before=self.m.clone(), v_before=self.v.clone(),
                                        true_image=self.true_img.clone(), relu_masks=masks, pool_indices=pools))
        capi.adam_compose_table(self.g_total, self.m, self.v, self.mod, self.x, self.true_img, self.epsilon, self.inner,
                                self.table, self.step_idx, BETA1, BETA2, ADAM_EPS)
        capi.step_advance(self.step_idx)

    def finish(self):
        # `true_img` holds (clamp(x + clamp(mod, ±eps), 0, 1) - mean)/std for the final modifier, which is
        # exactly image_attacks.py:360-361.
        # (a reusable run keeps its buffers for the next call: the caller gets its own copy of the result)
        adv = clip_of(self.true_img.clone() if self.reusable else self.true_img, self.b, self.f)
        n = self.step_no
        cost_host = self.cost_log[:n].cpu().numpy() if n > 0 else np.zeros(0, dtype=np.float32)
        weights_host = self.weights_log[:n].cpu().numpy() if self.adaptive and n > 0 else None
        return LoopResult(adv, cost_host, weights_host, None)


class DispersionRun(ImageGuidedRun):
    """Dispersion Reduction (reference image_attacks.py:129-234, `ImageGuidedStd_Adam`): same skeleton as the
    I2V loop (modifier 196-199, compose 211-212, Adam 221-223, final compose 230-234) with
    cost = sum over hooked layers of `activations.std()` (214-220) and no clean-feature pass.

    std() runs over the WHOLE [N,C,h,w] map, so all frames of a call are coupled: when the frames are
    processed in several chunks the step makes two passes — forward + K6 accumulate over every chunk, then
    forward again + K6 gradient + data gradient per chunk.  With one chunk (N <= chunk, the reference's
    batch-size-1 runs) the single forward serves both."""

    def __init__(self, engines, epsilon, steps, step_size, chunk=None, tap=None):
        super().__init__(engines, epsilon, steps, step_size, chunk=chunk, tap=tap)

    def setup(self, videos):
        if videos.dim() != 5 or videos.shape[1] != 3:
            raise ValueError("videos must be [b,3,f,h,w], got %s" % (tuple(videos.shape),))
        device = torch.device("cuda", torch.cuda.current_device())
        capi.device_check(device)
        b, c, f, h, w = videos.shape
        self.b, self.f = b, f
        frames = frames_of(videos.to(device=device, dtype=torch.float32, non_blocking=True))
        N = self.N = b * f
        inner = self.inner = h * w
        chunk = self.chunk = _chunk_frames(self.engines, N, self.chunk_request, h, w, device)
        self.spans = [(s, min(s + chunk, N)) for s in range(0, N, chunk)]
        steps = self.steps
        self.x = torch.empty_like(frames)
        capi.denorm(frames, self.x, inner)                                      # image_attacks.py:201
        self.mod = torch.empty_like(frames)
        capi.fill(self.mod, INIT_MODIFIER)                                      # image_attacks.py:197
        self.m = torch.zeros_like(frames)
        self.v = torch.zeros_like(frames)
        self.true_img = torch.empty_like(frames)
        self.g_total = torch.empty_like(frames)
        self.acc = torch.zeros(self.n_layers, 2, device=device, dtype=torch.float64)
        self.stats = torch.zeros(self.n_layers, 4, device=device, dtype=torch.float32)
        self.workspace = capi.std_workspace(device)
        self.grads = None
        self.step_idx = torch.zeros(1, device=device, dtype=torch.int32)
        self.cost_log = torch.zeros(max(steps, 1), device=device, dtype=torch.float32)
        self.table = capi.adam_step_table(steps, self.step_size, BETA1, BETA2).to(device)
        capi.compose_norm(self.x, self.mod, self.true_img, self.epsilon, inner)   # image_attacks.py:211-212
        self.step_no = 0
        return self

    def _grad_bufs(self, feats_per_engine):
        if self.grads is None:
            self.grads = [[torch.empty((self.chunk,) + tuple(t.shape[1:]), device=t.device, dtype=torch.float32) for t in fe]
                          for fe in feats_per_engine]
        return self.grads

    def _backward_chunk(self, s0, s1, feats_per_engine):
        layer = 0
        grads = self._grad_bufs(feats_per_engine)
        for ei, (e, feats, gr) in enumerate(zip(self.engines, feats_per_engine, grads)):
            gviews = []
            for a, ga in zip(feats, gr):
                gv = ga[:s1 - s0]
                capi.std_grad(a, gv, self.stats[layer], relu_mask=e.relu_masked_grads)
                gviews.append(gv)
                layer += 1
            g = e.input_grad(gviews)
            if ei == 0:
                self.g_total[s0:s1].copy_(g)
            else:
                self.g_total[s0:s1].add_(g)

    def step(self):
        if self.step_no >= self.steps:
            raise RuntimeError("all %d steps of this run are done" % self.steps)
        self.acc.zero_()
        numel = [0] * self.n_layers
        single = len(self.spans) == 1
        kept = None
        for (s0, s1) in self.spans:                                             # pass 1: statistics of every hooked map
            layer = 0
            feats_per_engine = []
            for e in self.engines:
                feats = e.features(self.true_img[s0:s1], need_grad=True)        # image_attacks.py:214
                feats_per_engine.append(feats)
                for a in feats:
                    capi.std_accumulate(a, self.workspace, self.acc[layer])
                    numel[layer] += a.numel()
                    layer += 1
            kept = feats_per_engine
        for layer in range(self.n_layers):                                      # image_attacks.py:216-220
            capi.std_finalize(self.acc[layer], numel[layer], self.stats[layer], self.cost_log, self.step_idx,
                              add_to_cost=layer > 0)
        if single:
            self._backward_chunk(self.spans[0][0], self.spans[0][1], kept)      # image_attacks.py:222 (cost.backward())
        else:
            if len(self.engines) > 1:
                raise NotImplementedError("multi-chunk dispersion reduction with several backbones")
            for (s0, s1) in self.spans:                                         # pass 2: forward again, gradient, backward
                feats_per_engine = [e.features(self.true_img[s0:s1], need_grad=True) for e in self.engines]
                self._backward_chunk(s0, s1, feats_per_engine)
        if self.tap is not None:
            self.tap(self.step_no, dict(g=self.g_total.clone(), stats=self.stats.clone(), mod_before=self.mod.clone(),
                                        m_before=self.m.clone(), v_before=self.v.clone(), true_image=self.true_img.clone()))
        capi.adam_compose_table(self.g_total, self.m, self.v, self.mod, self.x, self.true_img, self.epsilon, self.inner,
                                self.table, self.step_idx, BETA1, BETA2, ADAM_EPS)   # image_attacks.py:221-223, 211-212
        capi.step_advance(self.step_idx)
        self.step_no += 1

    def finish(self):
        adv = clip_of(self.true_img, self.b, self.f)                            # image_attacks.py:230-234
        n = self.step_no
        cost_host = self.cost_log[:n].cpu().numpy() if n > 0 else np.zeros(0, dtype=np.float32)
        return LoopResult(adv, cost_host, None, None)


def run_dispersion(engines, videos, epsilon, steps, step_size, chunk=None, tap=None):
    run = DispersionRun(engines, epsilon, steps, step_size, chunk=chunk, tap=tap)
    run.setup(videos)
    for _ in range(run.steps):
        run.step()
    return run.finish()


def run_image_guided(engines, videos, epsilon, steps, step_size, adaptive=False, coeffs=None, momentum=0.0,
                     coef_CE=False, chunk=None, reduce_hook=None, tap=None, layer_offsets=None, n_layers_total=None,
                     cache=None):
    """cache: a dict owned by the caller (the attack object).  The run — device state, captured CUDA graph — is kept in it
    and set up again when the next call has the same configuration and clip shape."""
    run = None
    key = (float(epsilon), int(steps), float(step_size), bool(adaptive), float(momentum), bool(coef_CE), chunk,
           tuple(id(e) for e in engines), None if coeffs is None else coeffs.data_ptr())
    if cache is not None and tap is None and reduce_hook is None:
        if cache.get("key") == key:
            run = cache["run"]
        else:
            cache.clear()
    if run is None:
        run = ImageGuidedRun(engines, epsilon, steps, step_size, adaptive, coeffs, momentum, coef_CE, chunk, reduce_hook, tap,
                             layer_offsets, n_layers_total)
        if cache is not None and tap is None and reduce_hook is None:
            run.reusable = True
            cache["key"], cache["run"] = key, run
    run.setup(videos)
    for _ in range(run.steps):
        run.step()
    return run.finish()


def record_loss_info(loss_info, video_names, cost):
    """image_attacks.py:355-358: the batch-total cost of every step under every video name."""
    for i, cval in enumerate(cost):
        text = str(np.float32(cval))
        for name in video_names:
            loss_info.setdefault(name, {})[i] = {"cost": text}
