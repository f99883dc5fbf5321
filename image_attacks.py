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
           "ImageGuidedFML2_Adam_MultiModels"]


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
        res = attack_loop.run_image_guided([self._engine], videos, self.epsilon, self.steps, self.step_size)
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
            self._plan = D.EnsemblePlan(model_name_lists, [1] * len(model_name_lists))
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
        res = attack_loop.run_image_guided(self._engines, videos, self.epsilon, self.steps, self.step_size, **extra)
        attack_loop.record_loss_info(self.loss_info, video_names, res.cost)
        return res.adv
