"""Library-path feature extractor: torchvision modules through PyTorch/cuDNN, truncated at the deepest
hooked layer and differentiated with respect to the input only.

This is the BASELINE engine (comparison point (ii) of BASELINE.md section 4): it already removes the
reference's wasted work — layers after the hook, avgpool/fc/dropout, every weight gradient
 — but the convolutions are cuDNN's.  The native engine (engine_native.py) replaces
them with this repo's own sm_100a kernels behind the same interface:

    feats = engine.features(img, need_grad)   # list of [N, ...] contiguous feature maps, hook order
    g     = engine.input_grad(grads)          # dcost/dimg given dcost/dfeat for every hooked layer
"""
import torch

from . import backbones


class _StopForward(Exception):
    pass


class CudnnEngine:
    relu_masked_grads = False   # autograd applies the hooked ReLU's backward itself
    preferred_chunk = 128       # frames per forward/backward: bounds autograd's saved activations

    def __init__(self, model, model_name, depth, allow_tf32=False, channels_last=False):
        self.model = backbones.freeze_for_attack(model)
        self.channels_last = bool(channels_last)
        if self.channels_last:
            self.model.to(memory_format=torch.channels_last)
        self.model_name = model_name
        self.depth = depth
        self.allow_tf32 = allow_tf32
        self.targets = backbones.find_target_layers(model, model_name, depth)
        self._acts = []
        self._armed = False
        self._img = None
        self._feats = None
        n_targets = len(self.targets)

        def hook(module, inputs, output):
            if not self._armed:
                return None
            self._acts.append(output)
            if len(self._acts) == n_targets:
                raise _StopForward()   # nothing after the last hooked layer feeds the loss
            return None

        # Forward hooks fire in execution order == the order the reference's activation list is filled
        # (image_attacks.py:281-283).  Targets are registered once; a list with a repeated module would
        # fire once per registration in the reference too.
        self._handles = [t.register_forward_hook(hook) for t in self.targets]

    @property
    def num_layers(self):
        return len(self.targets)

    def close(self):
        for h in self._handles:
            h.remove()
        self._handles = []

    def _run(self, img):
        self._acts = []
        self._armed = True
        prev_tf32 = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = self.allow_tf32
        try:
            self.model(img.contiguous(memory_format=torch.channels_last) if self.channels_last else img)
        except _StopForward:
            pass
        finally:
            self._armed = False
            torch.backends.cudnn.allow_tf32 = prev_tf32
        acts, self._acts = self._acts, []
        if len(acts) != len(self.targets):
            raise RuntimeError("hooked %d layers but captured %d activations" % (len(self.targets), len(acts)))
        return acts

    def features(self, img, need_grad):
        if not need_grad:
            with torch.no_grad():
                return [a.contiguous() for a in self._run(img)]
        self._img = img.detach().requires_grad_(True)
        with torch.enable_grad():
            self._feats = self._run(self._img)
        return [a.detach() if a.is_contiguous() else a.detach().contiguous() for a in self._feats]

    def input_grad(self, grads):
        prev_tf32 = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = self.allow_tf32
        try:
            (g,) = torch.autograd.grad(self._feats, self._img, [gr.view_as(f) for gr, f in zip(grads, self._feats)])
        finally:
            torch.backends.cudnn.allow_tf32 = prev_tf32
        self._feats = None
        self._img = None
        return g.contiguous()
