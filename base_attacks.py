"""Drop-in replacement for the hot-path part of the reference's `base_attacks` module.

`Attack` (base_attacks.py:12-234), `FGSM` (236-259), `BIM` (261-295), `MIFGSM` (297-340): gradient-sign
attacks on a white-box *video* model.  The model's forward/backward is the caller's `nn.Module`
(opaque, PyTorch autograd — exactly as in the reference); the sign step / eps-projection / [0,1] clamp /
re-normalise block that every class repeats (289-293, 334-338, ...) and MI's per-frame mean-|g|
normalisation + momentum (328-332, utils.py:58-67) run in this repo's K3b / K3c kernels.

The remaining variants (DIFGSM, TIFGSM, SGM, SIM, TIFGSM3D, TAP) share the same update block and are
listed as "next" in SURVEY.md 8(f).
"""
import torch
import torch.nn as nn

from i2v_b200 import capi

__all__ = ["Attack", "FGSM", "BIM", "MIFGSM"]


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
