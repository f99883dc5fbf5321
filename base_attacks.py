"""Drop-in replacement for the hot-path part of the reference's `base_attacks` module.

`Attack` (base_attacks.py:12-234), `FGSM` (236-259), `BIM` (261-295), `MIFGSM` (297-340): gradient-sign
attacks on a white-box *video* model.  The model's forward/backward is the caller's `nn.Module`
(opaque, PyTorch autograd — exactly as in the reference); the sign step / eps-projection / [0,1] clamp /
re-normalise block that every class repeats (289-293, 334-338, ...) and MI's per-frame mean-|g|
normalisation + momentum (328-332, utils.py:58-67) run in this repo's K3b / K3c kernels.

The transfer-enhancing variants of SURVEY.md 8(f) rank 3 share that update block and differ only in how the
gradient is obtained / conditioned: `DIFGSM` (342-411, random resize + pad of the input), `TIFGSM` (413-479,
per-frame 15x15 Gaussian smoothing of the gradient: K7 stencil kernel), `SGM` (481-551, ReLU backward hooks),
`SIM` (553-611, gradient averaged over 5 input scales), `TIFGSM3D` (613-683, 15x15x15 smoothing + frame-level
mean-|g| normalisation), `TAP` (685-814, feature-magnification loss on hooked layers of a video model — `model_type`
i3d / slowfast / tpn pick the layers, the `target_layers` attribute any other modules — plus the box-filter regulariser on the K7
stencil kernel).
"""
import contextlib
import random

import numpy as np
import torch
import torch.nn as nn

from i2v_b200 import capi

# FP32-parity mode (BASELINE north star): gradients of the white-box model are taken with TF32 off; set to False to
# keep whatever torch.backends.* says
FP32_PARITY = True

__all__ = ["Attack", "FGSM", "BIM", "MIFGSM", "DIFGSM", "TIFGSM", "SGM", "SIM", "TIFGSM3D", "TAP"]


@contextlib.contextmanager
def fp32_parity():
    """The white-box model is an opaque nn.Module that runs on cuDNN / cuBLAS, whose defaults would silently compute
    convolutions in TF32 (1e-3 relative) and flip the sign of small gradient entries: TF32 off inside, restored after."""
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    if FP32_PARITY:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


class Attack(object):
    """Base class of the gradient-sign attacks (reference base_attacks.py:12-234).

    It takes the device from the model and switches the model to eval() for the duration of a call.
    """

    def __init__(self, name, model):
        self.attack = name
        self.model = model
        self.model_name = str(model).split("(")[0]
        self.training = model.training
        self.device = next(model.parameters()).device
        self._targeted = 1
        self._attack_mode = "default"
        self._return_type = "float"
        self._target_map_function = lambda images, labels: labels
        self.mean = [0.485, 0.456, 0.406]
        self.std = [0.229, 0.224, 0.225]

    def forward(self, *input):
        raise NotImplementedError

    # ---- attack mode / return type (reference base_attacks.py:49-93) --------------------------------
    def set_attack_mode(self, mode, target_map_function=None):
        if self._attack_mode == "only_default":
            raise ValueError("Changing attack mode is not supported in this attack method.")
        if mode == "targeted" and target_map_function is None:
            raise ValueError("Please give a target_map_function, e.g., lambda images, labels:(labels+1)%10.")
        if mode == "default":
            self._attack_mode = "default"
            self._targeted = 1
            self._transform_label = self._get_label
        elif mode == "targeted":
            self._attack_mode = "targeted"
            self._targeted = -1
            self._target_map_function = target_map_function
            self._transform_label = self._get_target_label
        elif mode == "least_likely":
            self._attack_mode = "least_likely"
            self._targeted = -1
            self._transform_label = self._get_least_likely_label
        else:
            raise ValueError(mode + " is not a valid mode. [Options : default, targeted, least_likely]")

    def set_return_type(self, type):
        if type == "float":
            self._return_type = "float"
        elif type == "int":
            self._return_type = "int"
        else:
            raise ValueError(type + " is not a valid type. [Options : float, int]")

    # ---- bulk generation helper (reference base_attacks.py:95-136) ---------------------------------
    def save(self, save_path, data_loader, verbose=True):
        self.model.eval()
        image_list, label_list = [], []
        correct = total = 0
        total_batch = len(data_loader)
        for step, (images, labels) in enumerate(data_loader):
            adv_images = self.__call__(images, labels)
            image_list.append(adv_images.cpu())
            label_list.append(labels.cpu())
            if self._return_type == "int":
                adv_images = adv_images.float() / 255
            if verbose:
                outputs = self.model(adv_images)
                _, predicted = torch.max(outputs.data, 1)
                total += labels.size(0)
                correct += (predicted == labels.to(self.device)).sum()
                acc = 100 * float(correct) / total
                print("- Save Progress : %2.2f %% / Accuracy : %2.2f %%" % ((step + 1) / total_batch * 100, acc),
                      end="\r")
        torch.save((torch.cat(image_list, 0), torch.cat(label_list, 0)), save_path)
        print("\n- Save Complete!")
        self._switch_model()

    # ---- normalisation helpers ([B,3,T,H,W], reference base_attacks.py:138-158) ---------------------
    def _transform_perts(self, perts):
        std = torch.as_tensor(self.std, dtype=perts.dtype, device=self.device)
        perts.div_(std[:, None, None, None])
        return perts

    def _transform_video(self, video, mode="forward"):
        if video.dim() < 4 or video.shape[-4] != 3:
            raise ValueError("expected [..,3,T,H,W], got %s" % (tuple(video.shape),))
        inner = video.shape[-1] * video.shape[-2] * video.shape[-3]
        if mode == "forward":
            capi.normalize(video, video, inner)
        elif mode == "back":
            capi.denorm(video, video, inner)
        return video

    # ---- label transforms (reference base_attacks.py:160-188) --------------------------------------
    def _transform_label(self, images, labels):
        return labels

    def _get_label(self, images, labels):
        return labels

    def _get_target_label(self, images, labels):
        return self._target_map_function(images, labels)

    def _get_least_likely_label(self, images, labels):
        outputs = self.model(images)
        _, labels = torch.min(outputs.data, 1)
        return labels.detach_()

    def _to_uint(self, images):
        return (images * 255).type(torch.uint8)

    def _switch_model(self):
        if self.training:
            self.model.train()
        else:
            self.model.eval()

    def __str__(self):
        info = {k: v for k, v in self.__dict__.items() if not k.startswith("_") and k not in ("model", "attack")}
        mode = self._attack_mode
        info["attack_mode"] = "default" if mode == "only_default" else mode
        info["return_type"] = self._return_type
        return self.attack + "(" + ", ".join("{}={}".format(k, v) for k, v in info.items()) + ")"

    def __call__(self, *input, **kwargs):
        self.model.eval()
        images = self.forward(*input, **kwargs)
        self._switch_model()
        if self._return_type == "int":
            images = self._to_uint(images)
        return images

    # ---- shared pieces of the gradient-sign loops ---------------------------------------------------
    def _ce_grad(self, adv_videos, labels, loss):
        """cost = _targeted * CE(model(adv), labels); d cost / d adv (reference base_attacks.py:283-287)."""
        adv_videos.requires_grad = True
        with fp32_parity():
            outputs = self.model(adv_videos)
            cost = self._targeted * loss(outputs, labels).to(self.device)
            grad = torch.autograd.grad(cost, adv_videos, retain_graph=False, create_graph=False)[0]
        return grad.contiguous()

    @staticmethod
    def _inner(videos):
        if videos.dim() != 5 or videos.shape[1] != 3:
            raise ValueError("videos must be [B,3,T,H,W], got %s" % (tuple(videos.shape),))
        return videos.shape[2] * videos.shape[3] * videos.shape[4]


class FGSM(Attack):
    """Fast Gradient Sign Method (reference base_attacks.py:236-259): one step of size epsilon, [0,1]
    clamp, no epsilon-projection."""

    def __init__(self, model, steps=None, epsilon=16 / 255):
        super(FGSM, self).__init__("FGSM", model)
        self.epsilon = epsilon

    def forward(self, videos, labels):
        videos = videos.to(self.device)
        labels = labels.to(self.device)
        loss = nn.CrossEntropyLoss()
        grad = self._ce_grad(videos, labels, loss)      # sets requires_grad on the caller's tensor, as 247 does
        adv_videos = videos.clone().detach().contiguous()
        capi.sign_step_project(adv_videos, grad, None, float(self.epsilon), float(self.epsilon), self._inner(videos),
                               project=False)
        return adv_videos


class BIM(Attack):
    """Basic Iterative Method (reference base_attacks.py:261-295); step_size = epsilon / steps."""

    def __init__(self, model, epsilon=16 / 255, steps=10):
        super(BIM, self).__init__("FGSM", model)   # the reference registers BIM under the name "FGSM" (267)
        self.epsilon = epsilon
        self.steps = steps
        self.step_size = self.epsilon / self.steps

    def forward(self, videos, labels):
        videos = videos.to(self.device)
        labels = labels.to(self.device)
        loss = nn.CrossEntropyLoss()
        inner = self._inner(videos)
        unnorm_videos = torch.empty_like(videos, memory_format=torch.contiguous_format)
        capi.denorm(videos.detach().contiguous(), unnorm_videos, inner)            # 279
        adv_videos = videos.clone().detach().contiguous()                           # 280
        for _ in range(self.steps):
            grad = self._ce_grad(adv_videos, labels, loss)                          # 283-287
            adv_videos = adv_videos.detach()
            capi.sign_step_project(adv_videos, grad, unnorm_videos, float(self.step_size), float(self.epsilon),
                                   inner, project=True)                             # 289-293
        return adv_videos


class MIFGSM(Attack):
    """Momentum Iterative FGSM (reference base_attacks.py:297-340) with the frame-level mean-|g|
    normalisation of utils.norm_grads (utils.py:58-67; the T == 32 assert is dropped)."""

    def __init__(self, model, epsilon=16 / 255, steps=10, decay=1.0):
        super(MIFGSM, self).__init__("MIFGSM", model)
        self.epsilon = epsilon
        self.steps = steps
        self.step_size = self.epsilon / self.steps
        self.decay = decay

    def forward(self, videos, labels):
        videos = videos.to(self.device)
        labels = labels.to(self.device)
        loss = nn.CrossEntropyLoss()
        inner = self._inner(videos)
        B, _, T = videos.shape[:3]
        momentum = torch.zeros_like(videos, memory_format=torch.contiguous_format)  # 316
        norm = torch.empty(B, T, device=self.device, dtype=torch.float32)
        unnorm_videos = torch.empty_like(videos, memory_format=torch.contiguous_format)
        capi.denorm(videos.detach().contiguous(), unnorm_videos, inner)             # 317
        adv_videos = videos.clone().detach().contiguous()                            # 318
        for _ in range(self.steps):
            grad = self._ce_grad(adv_videos, labels, loss)                           # 321-326
            adv_videos = adv_videos.detach()
            capi.frame_absmean(grad, norm, clip_level=False)                         # 328 -> utils.py:63
            capi.mi_sign_step_project(adv_videos, grad, momentum, norm, unnorm_videos, float(self.decay),
                                      float(self.step_size), float(self.epsilon))    # 328-338
        return adv_videos


# --------------------------------------------------------------------------------------------------------------
# transfer-enhancing variants: same K3b update block, different gradient
# --------------------------------------------------------------------------------------------------------------
class _SignLoopAttack(Attack):
    """Shared skeleton of DIFGSM / TIFGSM / SGM / SIM / TIFGSM3D (reference base_attacks.py:378-411 and its copies):

        for steps:  grad = self._gradient(adv, labels, loss)             # subclass
                    [momentum]  see _accumulate                          # subclass-specific normalisation
                    K3b: denorm -> + step*sign(grad) -> eps-projection -> [0,1] clamp -> renorm   (405-409)
    """

    def _setup(self, epsilon, steps, decay, momentum):
        self.epsilon = epsilon
        self.steps = steps
        self.step_size = self.epsilon / self.steps
        self.decay = decay
        self.momentum = momentum

    def _gradient(self, adv_videos, labels, loss):
        return self._ce_grad(adv_videos, labels, loss)

    def _accumulate(self, grad, momentum):
        """reference 394-398 (DI / SGM / SIM): grad /= ||grad||_1; grad += momentum*decay; momentum = grad."""
        grad = grad / torch.norm(grad, p=1)
        grad = grad + momentum * self.decay
        return grad, grad

    def forward(self, videos, labels):
        videos = videos.to(self.device)
        labels = labels.to(self.device)
        loss = nn.CrossEntropyLoss()
        inner = self._inner(videos)
        momentum = torch.zeros_like(videos, memory_format=torch.contiguous_format)
        unnorm_videos = torch.empty_like(videos, memory_format=torch.contiguous_format)
        capi.denorm(videos.detach().contiguous(), unnorm_videos, inner)
        adv_videos = videos.clone().detach().contiguous()
        for _ in range(self.steps):
            grad = self._gradient(adv_videos, labels, loss)
            if self.momentum:
                grad, momentum = self._accumulate(grad, momentum)
            adv_videos = adv_videos.detach()
            capi.sign_step_project(adv_videos, grad.contiguous(), unnorm_videos, float(self.step_size), float(self.epsilon),
                                   inner, project=True)
        return adv_videos


class DIFGSM(_SignLoopAttack):
    """Diverse Inputs Method (reference base_attacks.py:342-411).  The random resize (nearest) to rnd in [224, 250),
    zero padding to 250 x 250 at a random offset and resize back to 224 x 224 are the reference's constants (356-376):
    like the reference this class needs 224 x 224 frames.  RNG consumption order is the reference's
    (random.random, then three torch.randint calls), so seeded runs follow the same transformations."""

    def __init__(self, model, epsilon=16 / 255, steps=10, decay=1.0, momentum=False):
        super(DIFGSM, self).__init__("DIFGSM", model)
        self._setup(epsilon, steps, decay, momentum)

    def _input_diversity(self, videos):
        if random.random() < 0.5:
            return videos
        rnd = torch.randint(224, 250, size=(1, 1)).item()
        rescaled = videos.view((-1,) + videos.shape[2:])
        rescaled = torch.nn.functional.interpolate(rescaled, size=[rnd, rnd], mode="nearest")
        h_rem = 250 - rnd
        w_rem = 250 - rnd
        pad_top = torch.randint(0, h_rem, size=(1, 1)).item()
        pad_bottom = h_rem - pad_top
        pad_left = torch.randint(0, w_rem, size=(1, 1)).item()
        pad_right = w_rem - pad_left
        padded = nn.functional.pad(rescaled, [pad_left, pad_right, pad_top, pad_bottom])
        padded = torch.nn.functional.interpolate(padded, size=[224, 224], mode="nearest")
        return padded.view(videos.shape)

    def _gradient(self, adv_videos, labels, loss):
        adv_videos.requires_grad = True
        outputs = self.model(self._input_diversity(adv_videos))
        cost = self._targeted * loss(outputs, labels).to(self.device)
        return torch.autograd.grad(cost, adv_videos, retain_graph=False, create_graph=False)[0].contiguous()


def _gaussian_1d(kernlen, nsig):
    """scipy.stats.norm.pdf(np.linspace(-nsig, nsig, kernlen)) without the scipy import (reference 427-429)."""
    x = np.linspace(-nsig, nsig, kernlen)
    return np.exp(-0.5 * x * x) / np.sqrt(2.0 * np.pi)


class TIFGSM(_SignLoopAttack):
    """Translation-Invariant attack (reference base_attacks.py:413-479): every frame of the gradient is smoothed with
    a depth-wise 15 x 15 Gaussian (K7 `i2v_depthwise_stencil_f32` instead of 3*T grouped conv2d calls) and divided by
    mean(|g|) over dims (1,2,3) — i.e. per (clip, image COLUMN), the reference's axes (447)."""

    def __init__(self, model, epsilon=16 / 255, steps=10, decay=1.0, momentum=False):
        super(TIFGSM, self).__init__("MIFGSM", model)          # the reference registers it under "MIFGSM" (416)
        self._setup(epsilon, steps, decay, momentum)
        kern1d = _gaussian_1d(15, 3)
        kernel_raw = np.outer(kern1d, kern1d)
        kernel = (kernel_raw / kernel_raw.sum()).astype(np.float32)                  # 427-432
        self.stack_kernel = torch.from_numpy(np.expand_dims(np.stack([kernel, kernel, kernel]), 1)).to(self.device)
        self._kernel = torch.from_numpy(kernel).to(self.device).contiguous()

    def _conv2d_frame(self, grads):
        out_grads = torch.empty_like(grads, memory_format=torch.contiguous_format)
        capi.depthwise_stencil(grads.contiguous(), out_grads, self._kernel)           # 441-446
        return out_grads / torch.mean(torch.abs(out_grads), [1, 2, 3], True)          # 447

    def _gradient(self, adv_videos, labels, loss):
        return self._conv2d_frame(self._ce_grad(adv_videos, labels, loss))

    def _accumulate(self, grad, momentum):
        grad = grad + momentum * self.decay                                           # 463-465 (no L1 normalisation)
        return grad, grad


class SGM(_SignLoopAttack):
    """Skip Gradient Method (reference base_attacks.py:481-551): the gradient through every module whose name contains
    'relu' (but not '0.relu') is scaled by gamma**0.5 by a backward hook registered once in the constructor."""

    def __init__(self, model, epsilon=16 / 255, steps=10, decay=1.0, gamma=0.5, momentum=False):
        super(SGM, self).__init__("SGM", model)
        self._setup(epsilon, steps, decay, momentum)
        self.gamma = gamma
        self._register_hook_for_model(self.model)

    def _register_hook_for_model(self, model):
        scale = float(np.power(self.gamma, 0.5))

        def _backward_hook(module, grad_in, grad_out):
            if isinstance(module, nn.ReLU):
                return (scale * grad_in[0],)

        self._sgm_handles = []
        for name, module in model.named_modules():
            if "relu" in name and "0.relu" not in name:
                # the reference's (deprecated) non-full hook; it is exact for a single-op module such as nn.ReLU
                self._sgm_handles.append(module.register_backward_hook(_backward_hook))


class SIM(_SignLoopAttack):
    """Scale-Invariant attack (reference base_attacks.py:553-611): the gradient is the mean over `sclae_step` copies
    of the input scaled by 1 / 2**i (the reference's spelling of the keyword is kept)."""

    def __init__(self, model, epsilon=16 / 255, steps=10, decay=1.0, sclae_step=5, momentum=False):
        super(SIM, self).__init__("SIM", model)
        self._setup(epsilon, steps, decay, momentum)
        self.sclae_step = sclae_step

    def _gradient(self, adv_videos, labels, loss):
        mean_grad = None
        for i in range(self.sclae_step):
            tmp_videos = 1 / 2 ** i * adv_videos                                       # 572 (gradient w.r.t. the SCALED input)
            grad = self._ce_grad(tmp_videos.detach(), labels, loss)
            mean_grad = grad if mean_grad is None else mean_grad + grad
        return mean_grad / self.sclae_step


class TIFGSM3D(_SignLoopAttack):
    """Translation-Invariant attack with a 15 x 15 x 15 Gaussian over (T, H, W) (reference base_attacks.py:613-683):
    K7 stencil, then utils.norm_grads (frame-level mean-|g|, K3c reduction)."""

    def __init__(self, model, epsilon=16 / 255, steps=10, decay=1.0, momentum=False):
        super(TIFGSM3D, self).__init__("TIFGSM3D", model)
        self._setup(epsilon, steps, decay, momentum)
        kern1d = _gaussian_1d(15, 3)
        kernel_raw = np.outer(kern1d, kern1d)
        used_kernel = np.zeros((15, 15, 15))
        for i in range(15):
            used_kernel[i] = kern1d[i] * kernel_raw                                   # 629-631
        kernel = (used_kernel / used_kernel.sum()).astype(np.float32)
        self.stack_kernel = torch.from_numpy(np.expand_dims(np.stack([kernel, kernel, kernel]), 1)).to(self.device)
        self._kernel = torch.from_numpy(kernel).to(self.device).contiguous()

    def _conv3d_frame(self, grads):
        from utils import norm_grads
        out_grads = torch.empty_like(grads, memory_format=torch.contiguous_format)
        capi.depthwise_stencil(grads.contiguous(), out_grads, self._kernel)           # 640
        return norm_grads(out_grads, True)                                            # 649

    def _gradient(self, adv_videos, labels, loss):
        return self._conv3d_frame(self._ce_grad(adv_videos, labels, loss))

    def _accumulate(self, grad, momentum):
        grad = grad + momentum * self.decay                                           # 667-669
        return grad, grad


class TAP(Attack):
    """Transferable Adversarial Perturbations (reference base_attacks.py:685-814).

    params = {'kernlen': 3, 'temporal_kernlen': 3, 'eta': 1e3, 'conv3d': True, 'model_type': 'i3d'|'slowfast'|'tpn'}
    (`model_type` selects the hooked layers exactly as 738-744 does; `target_layers`, a list of modules, is an extension for
    other models).  cost = CE + 1e3 * sum|box(pert / std)| + 0.05 * sum_l |sgn(f_l) sqrt|f_l| - sgn(f0_l) sqrt|f0_l||_2
    (the reference hard-codes 1e3 and 0.05, 799, and ignores `eta`).

    The CE and feature-distance terms go through the opaque white-box model with torch autograd (so the reference's
    behaviour at exactly-zero features — a NaN derivative of sign*sqrt|.| — is inherited, not re-defined; K3b then treats a
    NaN gradient entry as sign 0 and leaves that pixel where it is, where torch.sign would turn it into NaN); the box-filter
    regulariser and its gradient are two passes of the K7 stencil (uniform kernels are symmetric, so the transposed
    convolution is the same stencil over sign(out)); the update is K3b.  `loss_info` is keyed by the step index (the
    reference's key `i` is shadowed by the feature loop, 790, and ends up being a tensor)."""

    def __init__(self, model, params, epsilon=16 / 255, steps=10):
        super(TAP, self).__init__("TAP", model)
        self.epsilon = epsilon
        self.steps = steps
        self.step_size = self.epsilon / self.steps
        self.target_layers = None
        for name, value in params.items():
            setattr(self, name, value)
        if self.kernlen % 2 != 1 or self.temporal_kernlen % 2 != 1:
            raise ValueError("TAP needs odd kernlen / temporal_kernlen ('same' padding, 725-734)")
        kernel = self._initial_kernel_uniform(self.kernlen).astype(np.float32)                        # 701
        stack_kernel = np.stack([kernel, kernel, kernel])
        self.stack_2d_kernel = torch.from_numpy(np.expand_dims(stack_kernel, 1)).to(self.device)      # 703
        kernel_3d = self._initial_kernel_uniform_3d(self.kernlen, self.temporal_kernlen)             # 705
        stack_kernel_3d = np.stack([kernel_3d, kernel_3d, kernel_3d])
        self.stack_3d_kernel = torch.from_numpy(np.expand_dims(stack_kernel_3d, 1)).to(self.device)   # 707
        self._k2d = torch.from_numpy(kernel).to(self.device).contiguous()
        self._k3d = torch.from_numpy(kernel_3d.astype(np.float32)).to(self.device).contiguous()
        self.loss_info = {}
        self._activation_hook()

    def _initial_kernel_uniform(self, kernlen):
        kern1d = np.ones(kernlen)
        kernel_raw = np.outer(kern1d, kern1d)
        return kernel_raw / kernel_raw.sum()

    def _initial_kernel_uniform_3d(self, kernlen, temporal_kernel):
        kern3d = np.ones((temporal_kernel, kernlen, kernlen))
        return kern3d / kern3d.sum()

    def _find_target_layer(self):
        if self.target_layers is not None:
            return list(self.target_layers)
        model_type = getattr(self, "model_type", "")
        if "i3d" in model_type:
            return [self.model.res_layers._modules["0"], self.model.res_layers._modules["1"]]
        if "slowfast" in model_type:
            return [self.model._modules["slow_res2"], self.model._modules["slow_res3"], self.model._modules["fast_res2"],
                    self.model._modules["fast_res3"]]
        if "tpn" in model_type:
            return [self.model.layer1, self.model.layer2]
        raise ValueError("TAP: params['model_type'] must contain i3d, slowfast or tpn (or pass params['target_layers'])")

    def _activation_hook(self):
        self.activations = {"value": []}

        def forward_hook(module, input, output):
            self.activations["value"] += [output]
            return None

        for layer in self._find_target_layer():
            layer.register_forward_hook(forward_hook)

    def _reg_cost_and_grad(self, adv_videos, videos):
        """reg = sum |box * ((adv - videos) / std)| (792-797, 720-735) and d reg / d adv."""
        std = torch.as_tensor(self.std, dtype=adv_videos.dtype, device=self.device)[None, :, None, None, None]
        perts = ((adv_videos - videos) / std).contiguous()                                            # 792, _transform_perts
        kernel = self._k3d if self.conv3d else self._k2d
        out = capi.depthwise_stencil(perts, torch.empty_like(perts), kernel)
        reg_cost = out.abs().sum()
        back = capi.depthwise_stencil(torch.sign(out), torch.empty_like(perts), kernel)
        return reg_cost, back / std

    def forward(self, videos, labels):
        batch_size = videos.shape[0]
        self.loss_info = {}
        videos = videos.to(self.device).contiguous()
        labels = labels.to(self.device)
        inner = self._inner(videos)
        with fp32_parity(), torch.no_grad():
            self.activations = {"value": []}
            self.model(videos)                                                                        # 768-769
            ori_feature_map = [f.detach() for f in self.activations["value"]]
        loss = nn.CrossEntropyLoss()
        unnorm_videos = torch.empty_like(videos)
        capi.denorm(videos.detach(), unnorm_videos, inner)                                            # 772
        adv_videos = videos.clone().detach()                                                          # 773
        for step in range(self.steps):
            self.activations = {"value": []}
            adv_videos.requires_grad = True
            with fp32_parity():
                outputs = self.model(adv_videos)                                                      # 779
                cost1 = self._targeted * loss(outputs, labels).to(self.device)                        # 782
                feat_distance = []
                for i, j in zip(self.activations["value"], ori_feature_map):                          # 787-789
                    this_distance = torch.norm((torch.sign(i) * torch.sqrt(torch.abs(i))).reshape(batch_size, -1) -
                                               (torch.sign(j) * torch.sqrt(torch.abs(j))).reshape(batch_size, -1), p=2, dim=1)
                    feat_distance.append(this_distance)
                cost2 = torch.sum(torch.stack(feat_distance), 0)                                      # 790
                grad = torch.autograd.grad((cost1 + 0.05 * cost2).sum(), adv_videos, retain_graph=False,
                                           create_graph=False)[0]
            adv_videos = adv_videos.detach()
            reg_cost, reg_grad = self._reg_cost_and_grad(adv_videos, videos)                          # 792-797
            grad = (grad + 1e3 * reg_grad).contiguous()                                               # 799
            capi.sign_step_project(adv_videos, grad, unnorm_videos, float(self.step_size), float(self.epsilon), inner,
                                   project=True)                                                      # 806-810
            self.loss_info[step] = {"ce loss": cost1.detach().cpu().numpy(), "reg_cost": reg_cost.detach().cpu().numpy(),
                                    "distance": cost2.detach().cpu().numpy()}
        return adv_videos
