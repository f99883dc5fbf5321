"""Image backbones of the attack and the reference's depth -> hooked-layer mapping.

Mirrors reference image_attacks.py:84-115 (`get_model`, `get_models`) and the three copies of
`_find_target_layer` (image_attacks.py:260-271, 391-403; list-aware TPAMI_attack.py:176-200).

Differences from the reference, all deliberate:
  * D2  model names: the reference's names keep the reference's architectures ('resnet' -> resnet101,
        'densenet' -> densenet161, 'vgg' -> vgg16, 'squeezenet' -> squeezenet1_1, 'alexnet'); the
        explicit names 'resnet50', 'resnet101', 'densenet121', 'densenet161' are added, and
        `ARCH_OVERRIDE` / $I2V_ARCH_MAP can re-point a reference name (BASELINE.json benchmarks
        'resnet' as ResNet-50; truncated at depth <= 2 the two are the same network).
  * D3  DenseNet has no branch in the reference's `_find_target_layer` (it would crash); here depth d
        hooks `features.denseblock{d}` — an extension, not reference behaviour.
  * weights: the reference hard-requires `pretrained=True` (image_attacks.py:86-98) and so does the default policy
        here, 'pretrained': torchvision's IMAGENET1K_V1 weights from the local hub cache, RuntimeError when they are
        missing (there is no network to fetch them).  Seeded random init is an explicit opt-in — policy 'random'
        (`set_weight_policy('random')`, `--weights random`, $I2V_WEIGHTS=random) — used by tests and benchmarks;
        'auto' (cached weights if present, else random init with a warning) must be asked for as well.
        `WEIGHT_SOURCE[arch]` records what every constructed backbone actually got; the drivers write it out.
  * unknown names raise ValueError listing the supported ones (the reference raises
        UnboundLocalError).
"""
import os
import warnings

import torch
import torch.nn as nn
import torchvision

REFERENCE_ARCH = {
    "alexnet": "alexnet",
    "vgg": "vgg16",
    "resnet": "resnet101",
    "densenet": "densenet161",
    "squeezenet": "squeezenet1_1",
}
EXTRA_ARCH = {
    "resnet50": "resnet50",
    "resnet101": "resnet101",
    "densenet121": "densenet121",
    "densenet161": "densenet161",
    "vgg16": "vgg16",
    "squeezenet1_1": "squeezenet1_1",
}
FAMILY = {
    "alexnet": "alexnet", "vgg": "vgg", "vgg16": "vgg", "resnet": "resnet", "resnet50": "resnet",
    "resnet101": "resnet", "densenet": "densenet", "densenet121": "densenet", "densenet161": "densenet",
    "squeezenet": "squeezenet", "squeezenet1_1": "squeezenet",
}
SUPPORTED = tuple(sorted(FAMILY))

# name -> torchvision constructor name; tests set {'resnet': 'resnet50'} to follow BASELINE.json
ARCH_OVERRIDE = {}
_WEIGHT_POLICY = {"mode": os.environ.get("I2V_WEIGHTS", "pretrained"), "seed": int(os.environ.get("I2V_WEIGHT_SEED", "0"))}
WEIGHT_SOURCE = {}     # arch -> "pretrained:<file>" | "random:seed=<n>", what get_model() last built for it

for _kv in filter(None, os.environ.get("I2V_ARCH_MAP", "").split(",")):
    _k, _v = _kv.split("=")
    ARCH_OVERRIDE[_k.strip()] = _v.strip()


def set_weight_policy(mode="pretrained", seed=0):
    if mode not in ("auto", "random", "pretrained"):
        raise ValueError("weight policy must be auto|random|pretrained, got %r" % (mode,))
    _WEIGHT_POLICY["mode"] = mode
    _WEIGHT_POLICY["seed"] = seed


def family_of(model_name):
    try:
        return FAMILY[model_name]
    except KeyError:
        raise ValueError("unknown image model %r; supported: %s" % (model_name, ", ".join(SUPPORTED)))


def arch_of(model_name):
    family_of(model_name)
    if model_name in ARCH_OVERRIDE:
        return ARCH_OVERRIDE[model_name]
    return REFERENCE_ARCH.get(model_name) or EXTRA_ARCH[model_name]


def seeded_random_init(arch, seed=0):
    """`torch.manual_seed(seed)` immediately before the torchvision constructor with weights=None —
    the synthetic-weights recipe of SURVEY.md 8(d); the global RNG state is restored afterwards."""
    state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    try:
        return getattr(torchvision.models, arch)(weights=None)
    finally:
        torch.random.set_rng_state(state)


def _pretrained_cached(arch):
    try:
        enum = torchvision.models.get_model_weights(arch)
        w = enum.DEFAULT if not hasattr(enum, "IMAGENET1K_V1") else enum.IMAGENET1K_V1
        fname = os.path.basename(w.url)
        path = os.path.join(torch.hub.get_dir(), "checkpoints", fname)
        return w if os.path.isfile(path) else None
    except Exception:
        return None


def get_model(model_name, device=None):
    """image_attacks.py:84-108.  Returns the torchvision module in eval mode on `device`."""
    arch = arch_of(model_name)
    mode = _WEIGHT_POLICY["mode"]
    model = None
    if mode in ("auto", "pretrained"):
        w = _pretrained_cached(arch)
        if w is not None:
            model = getattr(torchvision.models, arch)(weights=w)
            WEIGHT_SOURCE[arch] = "pretrained:" + os.path.basename(w.url)
        elif mode == "pretrained":
            raise RuntimeError("pretrained ImageNet weights for %s are not in the local torch hub cache (%s) and cannot be "
                               "downloaded here; a random-weight surrogate has no transfer value, so it is opt-in: "
                               "--weights random / I2V_WEIGHTS=random / backbones.set_weight_policy('random')"
                               % (arch, os.path.join(torch.hub.get_dir(), "checkpoints")))
        else:
            warnings.warn("i2v_b200: no cached ImageNet weights for %s; using seeded random init (seed %d)"
                          % (arch, _WEIGHT_POLICY["seed"]))
    if model is None:
        model = seeded_random_init(arch, _WEIGHT_POLICY["seed"])
        WEIGHT_SOURCE[arch] = "random:seed=%d" % _WEIGHT_POLICY["seed"]
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    model.to(device)
    model.eval()
    return model


def get_models(model_name_lists, device=None):
    """image_attacks.py:110-115"""
    return [get_model(n, device) for n in model_name_lists]


_ALEXNET_DEPTH = {1: 1, 2: 4, 3: 7, 4: 11}       # image_attacks.py:264
_VGG_DEPTH = {1: 1, 2: 11, 3: 20, 4: 29}          # image_attacks.py:267
_SQUEEZE_DEPTH = {1: 3, 2: 6, 3: 9, 4: 12}        # image_attacks.py:270


def find_target_layers(model, model_name, depth):
    """depth -> list of hooked modules, in the order the reference registers them.

    Scalar depth: image_attacks.py:260-271 (SqueezeNet hooks `.expand3x3_activation`).
    List of depths: TPAMI_attack.py:176-200 (SqueezeNet hooks the whole `Fire`).
    """
    fam = family_of(model_name)
    is_list = isinstance(depth, (list, tuple))
    depths = list(depth) if is_list else [depth]
    for d in depths:
        if d not in (1, 2, 3, 4):
            raise ValueError("depth must be in {1,2,3,4}, got %r" % (d,))
    if fam == "resnet":
        return [getattr(model, "layer%d" % d)[-1] for d in depths]
    if fam == "alexnet":
        return [model.features[_ALEXNET_DEPTH[d]] for d in depths]
    if fam == "vgg":
        return [model.features[_VGG_DEPTH[d]] for d in depths]
    if fam == "squeezenet":
        if is_list:
            return [model.features[_SQUEEZE_DEPTH[d]] for d in depths]
        return [model.features[_SQUEEZE_DEPTH[depths[0]]].expand3x3_activation]
    if fam == "densenet":   # extension
        return [getattr(model.features, "denseblock%d" % d) for d in depths]
    raise ValueError("unsupported model family %r" % fam)


def freeze_for_attack(model):
    """image_attacks.py:253-256: model.train() with every BatchNorm back in eval().  Nothing up to the
    hooked layers behaves differently in train mode (Dropout lives in the classifiers, after every
    hook), so the observable module state is kept identical to the reference's.  Weights never need
    gradients because nobody reads them."""
    model.train()
    for p in model.parameters():
        p.requires_grad_(False)
    for m in model.modules():
        if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
            m.eval()
    return model
