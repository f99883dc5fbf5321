"""Drop-in for the hot-path part of the reference's `utils` module: `norm_grads` (utils.py:58-67) and
`AverageMeter` (utils.py:40-56).  The gluoncv config helpers (utils.py:1-38) configure the video
models of the evaluation scripts and are out of scope."""
import torch

from i2v_b200 import capi


class AverageMeter(object):
    """Computes and stores the average and current value."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = 0
        self.avg = 0
        self.sum = 0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def norm_grads(grads, frame_level=True):
    """grads / mean(|grads|) per (clip, frame) over (C,H,W), or per clip over (C,T,H,W).

    The reference asserts T == 32 (utils.py:61), which makes MI-FGSM unusable on 16-frame clips
    (SURVEY.md D4); the assert is dropped here.  The norm comes from the K3c reduction kernel."""
    if grads.dim() != 5:
        raise ValueError("norm_grads expects [B,C,T,H,W], got %s" % (tuple(grads.shape),))
    g = grads.contiguous()
    B, C, T, H, W = g.shape
    norm = torch.empty((B,) if not frame_level else (B, T), device=g.device, dtype=torch.float32)
    capi.frame_absmean(g, norm, clip_level=not frame_level)
    shape = (B, 1, T, 1, 1) if frame_level else (B, 1, 1, 1, 1)
    return g / norm.view(shape)
