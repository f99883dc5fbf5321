"""Seeded synthetic inputs (no dataset or checkpoint is reachable offline) — SURVEY.md 8(d).

clip i : g = torch.Generator().manual_seed(1000 + i); x01 = torch.rand(b,3,f,H,W, generator=g);
         videos = (x01 - mean_c) / std_c  (ImageNet-normalised, like datasets.py:86-93 delivers them)
labels : i mod 400 (Kinetics-400) or mod 101 (UCF-101)
`smooth=True` gives the low-frequency variant (8x8x4 noise, trilinear up-sampling): iid noise is the
worst case for trajectory chaos, natural video is much smoother.
"""
import torch

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


def clip(index=0, b=1, f=32, h=224, w=224, smooth=False, num_classes=400):
    g = torch.Generator().manual_seed(1000 + index)
    if smooth:
        low = torch.rand(b, 3, 4, 8, 8, generator=g)
        x01 = torch.nn.functional.interpolate(low, size=(f, h, w), mode="trilinear", align_corners=False)
        x01 = x01.clamp_(0, 1)
    else:
        x01 = torch.rand(b, 3, f, h, w, generator=g)
    mean = torch.tensor(MEAN).view(1, 3, 1, 1, 1)
    std = torch.tensor(STD).view(1, 3, 1, 1, 1)
    videos = (x01 - mean) / std
    labels = torch.tensor([(index * b + k) % num_classes for k in range(b)], dtype=torch.long)
    return videos, labels


class TinyVideoNet(torch.nn.Module):
    """A small seeded white-box *video* classifier [B,3,T,H,W] -> logits, standing in for the gluoncv
    I3D/SlowFast/TPN models of the reference's attack.py (not installable offline)."""

    def __init__(self, num_classes=10, width=8, seed=0):
        super().__init__()
        state = torch.random.get_rng_state()
        torch.manual_seed(seed)
        try:
            self.conv1 = torch.nn.Conv3d(3, width, 3, padding=1)
            self.conv2 = torch.nn.Conv3d(width, width * 2, 3, stride=(1, 2, 2), padding=1)
            self.fc = torch.nn.Linear(width * 2, num_classes)
        finally:
            torch.random.set_rng_state(state)

    def forward(self, x):
        x = torch.relu(self.conv1(x))
        x = torch.relu(self.conv2(x))
        return self.fc(x.mean(dim=(2, 3, 4)))


class TinyReluVideoNet(torch.nn.Module):
    """TinyVideoNet with nn.ReLU MODULES named `layer.N.relu` — what SGM's hook registration looks for
    (reference base_attacks.py:512-514: names containing 'relu' but not '0.relu')."""

    class Block(torch.nn.Module):
        def __init__(self, cin, cout, stride):
            super().__init__()
            self.conv = torch.nn.Conv3d(cin, cout, 3, stride=stride, padding=1)
            self.relu = torch.nn.ReLU()
            self.skip = cin == cout and stride == 1      # residual block: scaling the ReLU branch's gradient then matters

        def forward(self, x):
            y = self.relu(self.conv(x))
            return x + y if self.skip else y

    def __init__(self, num_classes=10, width=8, seed=0):
        super().__init__()
        state = torch.random.get_rng_state()
        torch.manual_seed(seed)
        try:
            self.layer = torch.nn.Sequential(self.Block(3, width, 1), self.Block(width, width * 2, (1, 2, 2)),
                                             self.Block(width * 2, width * 2, 1))
            self.fc = torch.nn.Linear(width * 2, num_classes)
        finally:
            torch.random.set_rng_state(state)

    def forward(self, x):
        return self.fc(self.layer(x).mean(dim=(2, 3, 4)))


class TinyTPNLike(torch.nn.Module):
    """Seeded stand-in with the attribute names the reference's `model_type == 'tpn'` branches hook (`layer1`, `layer2`:
    TAP base_attacks.py:738-744, ILAF image_attacks.py:514-520).  The hooked modules are bare Conv3d layers (the ReLUs are
    functional), so hooked features have no exact zeros — TAP's sign(f)*sqrt(|f|) has a NaN derivative at 0."""

    def __init__(self, num_classes=10, width=8, seed=0):
        super().__init__()
        state = torch.random.get_rng_state()
        torch.manual_seed(seed)
        try:
            self.layer1 = torch.nn.Conv3d(3, width, 3, padding=1)
            self.layer2 = torch.nn.Conv3d(width, width * 2, 3, stride=(1, 2, 2), padding=1)
            self.fc = torch.nn.Linear(width * 2, num_classes)
        finally:
            torch.random.set_rng_state(state)

    def forward(self, x):
        x = torch.relu(self.layer1(x))
        x = torch.relu(self.layer2(x))
        return self.fc(x.mean(dim=(2, 3, 4)))
