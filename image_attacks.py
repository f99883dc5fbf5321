"""Drop-in replacement for the reference's `image_attacks` module (I2V, CVPR'22; ENS-I2V).

Same class names, constructor arguments, call signatures, return types and public attributes as
reference image_attacks.py (`Attack` 12-82, `get_model(s)` 84-115, `ImageGuidedFMDirection_Adam`
236-364, `ImageGuidedFML2_Adam_MultiModels` 366-496), so `getattr(image_attacks, name)(...)` in the
reference's drivers (image_main.py:68-89) keeps working.  The per-step arithmetic runs in this repo's
sm_100a kernels through `i2v_b200` (see i2v_b200/attack_loop.py for the line-by-line mapping); there
is no CPU path.

Extensions (keyword-only, default to reference behaviour): `engine=` selects the convolution engine
('native' | 'cudnn', see i2v_b200/engines.py).
"""
import torch

from i2v_b200 import attack_loop, backbones, capi, engines

__all__ = ["Attack", "get_model", "get_models", "ImageGuidedStd_Adam", "ImageGuidedFMDirection_Adam",
           "ImageGuidedFML2_Adam_MultiModels", "ILAF"]


class Attack(object):
    """Base class of the image-guided attacks (reference image_attacks.py:12-82)."""

    def __init__(self, name, model=None):
        self.attack = name
        self.model = model
        self.model_name = str(model).split("(")[0]
        # ImageNet statistics, also used for Kinetics-400 (reference image_attacks.py:33-34)
        self.mean = [0.485, 0.456, 0.406]
        self.std = [0.229, 0.224, 0.225]

    def forward(self, *input):
        raise NotImplementedError

    def _transform_perts(self, perts):
        """perts / std, in place, [.., 3, H, W] (reference image_attacks.py:43-48)."""
        std = torch.as_tensor(self.std, dtype=perts.dtype, device=perts.device)
        perts.div_(std[:, None, None])
        return perts

    def _transform_video(self, video, mode="forward"):
        """In-place (x-mean)/std ('forward') or x*std+mean ('back') on [N,3,H,W]
        (reference image_attacks.py:50-63), through the K3 kernels."""
        if video.dim() < 3 or video.shape[-3] != 3:
            raise ValueError("expected [..,3,H,W], got %s" % (tuple(video.shape),))
        inner = video.shape[-1] * video.shape[-2]
        if mode == "forward":
            capi.normalize(video, video, inner)
        elif mode == "back":
            capi.denorm(video, video, inner)
        return video

    def _transform_video_ILAF(self, video, mode="forward"):
        """The 5-D form [b,3,f,h,w] of `_transform_video` (reference image_attacks.py:65-78), in place."""
        if video.dim() != 5 or video.shape[1] != 3:
            raise ValueError("expected [b,3,f,h,w], got %s" % (tuple(video.shape),))
        inner = video.shape[2] * video.shape[3] * video.shape[4]
        if mode == "forward":
            capi.normalize(video, video, inner)
        elif mode == "back":
            capi.denorm(video, video, inner)
        return video

    def __call__(self, *input, **kwargs):
        return self.forward(*input, **kwargs)


def get_model(model_name):
    """reference image_attacks.py:84-108 (names: alexnet, vgg, resnet, densenet, squeezenet)."""
    return backbones.get_model(model_name)


def get_models(model_name_lists):
    """reference image_attacks.py:110-115"""
    return backbones.get_models(model_name_lists)


class ImageGuidedStd_Adam(Attack):
    """Dispersion Reduction (DR) attack — reference image_attacks.py:129-234 (the baseline `image_main.py`
    dispatches next to I2V): minimise the standard deviation of the hooked feature map.

    parameters:
        model_name_lists: [one image model name]
        depth: {1,2,3,4}
    """

    def __init__(self, model_name_lists, depth, step_size, epsilon=16 / 255, steps=10, *, engine=None):
        super(ImageGuidedStd_Adam, self).__init__("ImageGuidedStd_Adam")
        self.epsilon = epsilon
        self.steps = steps
        self.step_size = step_size
        self.loss_info = {}
        self.depth = depth
        self.model = get_models(model_name_lists)[0]
        self.model_name = model_name_lists[0]
        self._engine = engines.make_engine(self.model, self.model_name, depth, engine)

    def forward(self, videos, labels, video_names):
        res = attack_loop.run_dispersion([self._engine], videos, self.epsilon, self.steps, self.step_size)
        attack_loop.record_loss_info(self.loss_info, video_names, res.cost)
        return res.adv


class ImageGuidedFMDirection_Adam(Attack):
    """The Image-to-Video (I2V) attack — reference image_attacks.py:236-364.

    parameters:
        model_name_lists: [one image model name]
        depth: {1,2,3,4}
    """

    def __init__(self, model_name_lists, depth, step_size, epsilon=16 / 255, steps=10, *, engine=None):
        super(ImageGuidedFMDirection_Adam, self).__init__("ImageGuidedFMDirection_Adam")
        self.epsilon = epsilon
        self.steps = steps
        self.step_size = step_size
        self.loss_info = {}
        self.depth = depth
        self.model = get_models(model_name_lists)[0]
        self.model_name = model_name_lists[0]
        self._engine = engines.make_engine(self.model, self.model_name, depth, engine)

    def forward(self, videos, labels, video_names):
        # labels are moved to the GPU and never read by the reference (image_attacks.py:298)
        res = attack_loop.run_image_guided([self._engine], videos, self.epsilon, self.steps, self.step_size,
                                           cache=self.__dict__.setdefault("_run_cache", {}))
        attack_loop.record_loss_info(self.loss_info, video_names, res.cost)
        return res.adv


class ImageGuidedFML2_Adam_MultiModels(Attack):
    """The ensemble I2V (ENS-I2V) attack — reference image_attacks.py:366-496.

    parameters:
        model_name_lists: image model names
        depths: {model name: depth in {1,2,3,4}}
    `step_size` is fixed at 0.005 as in the reference (image_attacks.py:376).
    """

    def __init__(self, model_name_lists, depths, epsilon=16 / 255, steps=60, *, engine=None, placement=None):
        super(ImageGuidedFML2_Adam_MultiModels, self).__init__("ImageGuidedFML2_Adam_MultiModels")
        self.epsilon = epsilon
        self.steps = steps
        self.step_size = 0.005
        self.loss_info = {}
        self.depths = depths
        self.model_names = model_name_lists
        # placement='ensemble' (extension): one backbone per GPU under torch.distributed (i2v_b200/dist.py: EnsemblePlan)
        self._plan = None
        if placement == "ensemble":
            from i2v_b200 import dist as D
            self._plan = D.EnsemblePlan(model_name_lists, [len(depths[n]) if isinstance(depths[n], (list, tuple)) else 1
                                                           for n in model_name_lists])
            mine = [model_name_lists[i] for i in self._plan.members]
        elif placement is None:
            mine = list(model_name_lists)
        else:
            raise ValueError("placement must be None or 'ensemble', got %r" % (placement,))
        self.models = get_models(mine)
        self._engines = [engines.make_engine(m, n, depths[n], engine) for m, n in zip(self.models, mine)]

    def forward(self, videos, labels, video_names):
        extra = {}
        if self._plan is not None:
            extra = dict(reduce_hook=self._plan.hook(), layer_offsets=self._plan.layer_offsets,
                         n_layers_total=self._plan.n_layers_total)
        res = attack_loop.run_image_guided(self._engines, videos, self.epsilon, self.steps, self.step_size,
                                           cache=self.__dict__.setdefault("_run_cache", {}), **extra)
        attack_loop.record_loss_info(self.loss_info, video_names, res.cost)
        return res.adv


class ILAF(Attack):
    """Intermediate Level Attack fine-tuning (reference image_attacks.py:498-629; "Enhancing adversarial example
    transferability with an intermediate level attack"): an existing adversarial clip is pushed further along its own
    feature-space displacement on a white-box *video* model.

        attack = ILAF(model, model_type, step_size=0.005, epsilon=16/255, steps=60)
        adv2 = attack(adv_videos, ori_videos, labels, video_names)

    `model_type` picks the hooked layer(s) as 514-520 does ('i3d' -> res_layers[1], 'slowfast' -> slow_res2 + fast_res2,
    'tpn' -> layer2); `target_layers=[modules]` is an extension for other models.  Per step: the model forward/backward is
    torch autograd on the opaque module; K9 (`i2v_ila_loss_f32` / `i2v_ila_grad_f32`) gives each layer's loss
    -(0.5 |d| / |d0| + <d0^, d^>) and its feature gradient; K3d (`i2v_sign_descent_compose_f32`) applies
    `modifier -= step_size * sign(grad)` and recomposes the next clip (615-617, 582-585).  No host sync inside the loop:
    the costs go to a device log read once at the end.

    The returned tensor reproduces 627-629 literally: the [b,3,f,h,w] result is *reinterpreted* as [b,f,3,h,w] and
    permuted back, which scrambles channels and frames unless f == 3 — a reference defect that a drop-in keeps
    (`image_fine_tune_attack.py` saves exactly this); `self.last_adv` holds the unscrambled clip."""

    def __init__(self, model, model_type, step_size=0.005, epsilon=16 / 255, steps=60, *, target_layers=None):
        super(ILAF, self).__init__("ILAF")
        self.epsilon = epsilon
        self.steps = steps
        self.step_size = step_size
        self.loss_info = {}
        self.model_type = model_type
        self.model = model
        self.target_layers = target_layers
        self.last_adv = None
        self._activation_hook()

    def _find_target_layer(self):
        if self.target_layers is not None:
            return list(self.target_layers)
        if "i3d" in self.model_type:
            return self.model.res_layers._modules["1"]
        if "slowfast" in self.model_type:
            return [self.model._modules["slow_res2"], self.model._modules["fast_res2"]]
        if "tpn" in self.model_type:
            return self.model.layer2
        raise ValueError("ILAF: model_type must contain i3d, slowfast or tpn (or pass target_layers=[...])")

    def _activation_hook(self):
        self.activations = {"value": []}

        def forward_hook(module, input, output):
            self.activations["value"] += [output]
            return None

        target_layer = self._find_target_layer()
        for layer in (target_layer if isinstance(target_layer, list) else [target_layer]):
            layer.register_forward_hook(forward_hook)

    def _features(self, x):
        self.activations = {"value": []}
        self.model(x)
        return list(self.activations["value"])

    def forward(self, videos, ori_videos, labels, video_names):
        from base_attacks import fp32_parity
        b, c, f, h, w = videos.shape
        device = torch.device("cuda", torch.cuda.current_device())
        capi.device_check(device)
        videos = videos.to(device).contiguous()
        ori_videos = ori_videos.to(device).contiguous()
        inner = f * h * w
        with fp32_parity(), torch.no_grad():
            ori_feature_maps = [a.detach().contiguous() for a in self._features(ori_videos)]        # 542-550
            adv_feature_maps = [a.detach() for a in self._features(videos)]                          # 553-561
        init_directions, init_norms = [], []
        for ori_di, adv_di in zip(ori_feature_maps, adv_feature_maps):                               # 563-569
            init_direction = adv_di - ori_di
            norm = torch.norm(init_direction, p=2)
            init_norms.append(float(norm))
            init_directions.append((init_direction / norm).contiguous())
        del adv_feature_maps
        ori_unnorm = torch.empty_like(ori_videos)
        capi.denorm(ori_videos, ori_unnorm, inner)                                                   # 573
        modifier = torch.empty_like(videos)
        capi.denorm(videos, modifier, inner)                                                         # 572
        modifier.sub_(ori_unnorm)                                                                    # 575-576
        true_image = torch.empty_like(videos)
        capi.compose_norm(ori_unnorm, modifier, true_image, float(self.epsilon), inner)              # 582-585
        workspace = capi.ila_workspace(device)
        stats = torch.zeros(len(ori_feature_maps), 4, device=device, dtype=torch.float32)
        cost_log = torch.zeros(max(self.steps, 1), device=device, dtype=torch.float32)
        step_idx = torch.zeros(1, device=device, dtype=torch.int32)
        for i in range(self.steps):
            inp = true_image.detach().requires_grad_(True)
            with fp32_parity():
                feats = self._features(inp)                                                          # 588-594
                grads = []
                for l, (fm, ori, d0, n0) in enumerate(zip(feats, ori_feature_maps, init_directions, init_norms)):
                    fmc = fm.detach().contiguous()
                    capi.ila_loss(fmc, ori, d0, n0, workspace, stats[l], cost_log, step_idx, add_to_cost=l > 0)   # 596-612
                    grads.append(capi.ila_grad(fmc, ori, d0, torch.empty_like(fmc), stats[l]).view_as(fm))
                g = torch.autograd.grad(feats, inp, grad_outputs=grads, retain_graph=False, create_graph=False)[0]  # 614
            capi.sign_descent_compose(g.contiguous(), modifier, ori_unnorm, true_image, float(self.epsilon),
                                      float(self.step_size), inner)                                  # 615-617, 582-585
            capi.step_advance(step_idx)
        costs = cost_log[: self.steps].cpu().numpy()
        for vid_name in video_names:                                                                 # 620-623
            self.loss_info.setdefault(vid_name, {})
            for i in range(self.steps):
                self.loss_info[vid_name][i] = {"cost": str(costs[i])}
        self.last_adv = true_image
        return true_image.reshape(b, f, c, h, w).permute([0, 2, 1, 3, 4])                           # 627-629
