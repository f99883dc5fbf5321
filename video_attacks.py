"""Drop-in replacement for the reference's `video_attacks` module: `TemporalTranslation`
(reference video_attacks.py:14-229, "Boosting the transferability of video adversarial examples via temporal
translation").  Same constructor (`model, params, epsilon, steps, delay`), the same `params` keys (`kernlen`,
`momentum`, `weight`, `move_type` in adj|large|random, `kernel_mode` in gaussian|linear|random) and the same call
`attack(videos, labels) -> adv_videos`.

Per step (reference lines in brackets):
    K8  i2v_temporal_shift_stack_f32   the kernlen cyclically shifted copies of the clip           [190-200, 93-146]
        white-box model forward/backward on slices of the stack (opaque nn.Module, torch autograd)  [201-210, 150-157]
    K8  i2v_temporal_combine_f32       (1-w) * sum_d k_d G_d  +  w * sum_d k_d shift_back(G_d)      [163-177, 80-91]
    K3c i2v_frame_absmean_f32 + i2v_mi_sign_step_project_f32   (momentum: norm_grads, += momentum*delay)   [217-228]
 or K3b i2v_sign_step_project_f32                              (no momentum)                               [224-228]

Differences from the reference, all where the reference cannot run: `frames` follows the clip length instead of the
hard-coded 32 (36; `norm_grads` asserts T == 32); empty model batches (kernlen in {3, 7, 11, ...} with
the 5-way split of 203-208 make `torch.cat([])` raise in `_get_grad`) are skipped; a batch of B > 1 clips repeats the
labels per variant instead of failing the CE shape check; `np.math.exp` (72) is `math.exp`; the per-step
`print('now_time', ...)` is dropped.
"""
import math
import random

import numpy as np
import torch
import torch.nn as nn

from base_attacks import Attack
from i2v_b200 import capi

__all__ = ["TemporalTranslation"]


class TemporalTranslation(Attack):
    def __init__(self, model, params, epsilon=16 / 255, steps=10, delay=1.0):
        super(TemporalTranslation, self).__init__("TemporalTranslation", model)
        self.epsilon = epsilon
        self.steps = steps
        self.step_size = self.epsilon / self.steps
        self.delay = delay
        for name, value in params.items():
            setattr(self, name, value)
        self.frames = 32                                           # 36; forward() follows the clip's own length
        self.cycle_move_list = self._move_info_generation()
        if self.kernel_mode == "gaussian":
            kernel = self._initial_kernel_gaussian(self.kernlen).astype(np.float32)
        elif self.kernel_mode == "linear":
            kernel = self._initial_kernel_linear(self.kernlen).astype(np.float32)
        elif self.kernel_mode == "random":
            kernel = self._initial_kernel_uniform(self.kernlen).astype(np.float32)
        else:
            raise ValueError("kernel_mode must be gaussian, linear or random, got %r" % (self.kernel_mode,))
        if self.move_type not in ("adj", "large", "random"):
            raise ValueError("move_type must be adj, large or random, got %r" % (self.move_type,))
        self._kernel_host = kernel
        self.kernel = torch.from_numpy(np.expand_dims(kernel, 0)).to(self.device)   # [1, kernlen] (45)

    # ---- kernels over the variants (47-78) ----------------------------------------------------------------
    def _move_info_generation(self):
        max_move = int((self.kernlen - 1) / 2)
        return [i for i in range(-max_move, max_move + 1)]

    def _initial_kernel_linear(self, kernlen):
        k = int((kernlen - 1) / 2)
        kern1d = [1 - i / (k + 1) for i in range(k + 1)]
        kern1d = np.array(kern1d[::-1][:-1] + kern1d)
        return kern1d / kern1d.sum()

    def _initial_kernel_uniform(self, kernlen):
        kern1d = np.ones(kernlen)
        return kern1d / kern1d.sum()

    def _initial_kernel_gaussian(self, kernlen):
        assert kernlen % 2 == 1
        k = (kernlen - 1) / 2
        sigma = k / 3
        k = int(k)
        kern1d = np.array([1 / (sigma * np.sqrt(2 * np.pi)) * math.exp(-(x ** 2) / (2 * (sigma ** 2))) for x in range(-k, k + 1)])
        return kern1d / kern1d.sum()

    # ---- frame moves (93-146): new[:, :, (i + move) mod frames] = adv[:, :, i] --------------------------------
    def _effective_move(self, cycle_move, frames):
        direction = -1 if cycle_move < 0 else 1
        if self.move_type == "adj":
            amount = abs(cycle_move) % frames
        elif self.move_type == "large":
            amount = abs(cycle_move)
            amount = amount % frames if amount == 0 else (amount + (int(frames / 2) - 1)) % frames
        else:                                                      # 'random': one randint per non-zero move (128-131)
            amount = 0 if cycle_move == 0 else random.randint(0, 100) % frames
        return direction * amount

    def _cycle_move(self, adv_videos, cycle_move):
        """The reference's helper, kept for callers that use it directly (93-105)."""
        out = torch.empty((1,) + tuple(adv_videos.shape), device=adv_videos.device, dtype=adv_videos.dtype)
        direction = -1 if cycle_move < 0 else 1
        capi.temporal_shift_stack(adv_videos.contiguous(), out, [direction * (abs(cycle_move) % adv_videos.shape[2])])
        return out[0]

    def _cycle_move_large(self, adv_videos, cycle_move):
        """107-120: |move| -> (|move| + frames/2 - 1) mod frames (0 stays 0)."""
        frames = adv_videos.shape[2]
        direction = -1 if cycle_move < 0 else 1
        amount = abs(cycle_move)
        amount = amount % frames if amount == 0 else (amount + (int(frames / 2) - 1)) % frames
        out = torch.empty((1,) + tuple(adv_videos.shape), device=adv_videos.device, dtype=adv_videos.dtype)
        capi.temporal_shift_stack(adv_videos.contiguous(), out, [direction * amount])
        return out[0]

    def _cycle_move_random(self, adv_videos, cycle_move):
        """122-135: a random amount (one `random.randint(0, 100)`) for every non-zero move."""
        frames = adv_videos.shape[2]
        direction = -1 if cycle_move < 0 else 1
        amount = 0 if cycle_move == 0 else random.randint(0, 100) % frames
        out = torch.empty((1,) + tuple(adv_videos.shape), device=adv_videos.device, dtype=adv_videos.dtype)
        capi.temporal_shift_stack(adv_videos.contiguous(), out, [direction * amount])
        return out[0]

    def _exchange_move(self, adv_videos, exchange_lists):
        """137-143: swap pairs of frames (index plumbing; not used by forward())."""
        new_videos = adv_videos.clone()
        for one_frame, ano_frame in exchange_lists:
            new_videos[:, :, one_frame] = adv_videos[:, :, ano_frame]
            new_videos[:, :, ano_frame] = adv_videos[:, :, one_frame]
        return new_videos

    def _conv1d_frame(self, grads):
        """80-91: grads [D,N,C,T,H,W] -> sum_d kernel[d] * grads[d] (K8 with weight 0)."""
        D, N, C, T, H, W = grads.shape
        out = torch.empty((N, C, T, H, W), device=grads.device, dtype=grads.dtype)
        capi.temporal_combine(grads.contiguous().view(D, N, C, T, H, W), self._kernel_host, [0] * D, 0.0, out)
        return out

    def _grad_augmentation(self, grads):
        """163-177: (1-weight) * conv1d(grads) + weight * conv1d(grads shifted back by their nominal moves), one K8 pass."""
        D, N, C, T, H, W = grads.shape
        out = torch.empty((N, C, T, H, W), device=grads.device, dtype=grads.dtype)
        capi.temporal_combine(grads.contiguous(), self._kernel_host, self.cycle_move_list, self.weight, out)
        return out

    def _get_grad(self, adv_videos, labels, loss):
        """150-157: CE gradient of one slice of the variant stack ([n_var * B, 3, T, H, W])."""
        return self._ce_grad(adv_videos, labels, loss)

    def forward(self, videos, labels):
        videos = videos.to(self.device)
        labels = labels.to(self.device)
        if videos.dim() != 5 or videos.shape[1] != 3:
            raise ValueError("videos must be [B,3,T,H,W], got %s" % (tuple(videos.shape),))
        B, C, T, H, W = videos.shape
        inner = T * H * W
        loss = nn.CrossEntropyLoss()
        momentum = torch.zeros_like(videos, memory_format=torch.contiguous_format)          # 182
        norm = torch.empty(B, T, device=self.device, dtype=torch.float32)
        unnorm_videos = torch.empty_like(videos, memory_format=torch.contiguous_format)
        capi.denorm(videos.detach().contiguous(), unnorm_videos, inner)                     # 185
        adv_videos = videos.clone().detach().contiguous()                                   # 186
        length = len(self.cycle_move_list)
        stack = torch.empty((length, B, C, T, H, W), device=self.device, dtype=torch.float32)
        grads = torch.empty_like(stack)
        grad = torch.empty_like(adv_videos)
        for _ in range(self.steps):
            moves = [self._effective_move(m, T) for m in self.cycle_move_list]              # 192-199
            capi.temporal_shift_stack(adv_videos, stack, moves)                             # 200
            batch_times = 5                                                                 # 202
            if self.model_name == "TPNet":
                batch_times = length                                                        # 204-206
            batch_size = math.ceil(length / batch_times)
            for i in range(batch_times):                                                    # 208-210
                lo, hi = i * batch_size, min((i + 1) * batch_size, length)
                if hi <= lo:
                    continue
                inp = stack[lo:hi].reshape((hi - lo) * B, C, T, H, W).detach()
                used_labels = labels.repeat(hi - lo)                                        # 152 (B = 1 in the reference)
                grads[lo:hi] = self._get_grad(inp, used_labels, loss).reshape(hi - lo, B, C, T, H, W)
            capi.temporal_combine(grads, self._kernel_host, self.cycle_move_list, self.weight, grad)   # 213-214
            if self.momentum:                                                               # 217-220
                capi.frame_absmean(grad, norm, clip_level=False)
                capi.mi_sign_step_project(adv_videos, grad, momentum, norm, unnorm_videos, float(self.delay),
                                          float(self.step_size), float(self.epsilon))
            else:
                capi.sign_step_project(adv_videos, grad, unnorm_videos, float(self.step_size), float(self.epsilon), inner,
                                       project=True)                                        # 224-228
        return adv_videos
