"""TEST INFRASTRUCTURE ONLY — loads the *unmodified* reference classes from /root/reference.

Used by `oracle/make_golden.py` (fixture generation, in the build container only) and by
`bench.py --impl reference` / the `cpu_baseline` leg when /root/reference exists.  Nothing in the
product path (`i2v_b200`, `image_attacks`, `TPAMI_attack`, `base_attacks`) may import this module.

The reference modules have the same top-level names as this repo's drop-in modules
(`image_attacks`, `TPAMI_attack`, `base_attacks`, `utils`), so they are loaded under the aliases
`ref_image_attacks`, `ref_TPAMI_attack`, `ref_base_attacks`, `ref_utils` with `sys.modules`
temporarily pointed at stubs for their un-installable imports:

  * `timm.models.create_model`           (TPAMI_attack.py:13, only used by dead get_vits 88-98)
  * `gluoncv.torch.engine.config`        (utils.py:2)
  * `image_cam`, `image_cam_utils`       (image_attacks.py:7,9 — dead GradCAM code that pulls cv2)
  * `torchvision.models.<ctor>(pretrained=True)` → seeded random init (no network)
  * `Tensor.cuda` / `Module.cuda` → identity when no GPU is visible (image_attacks.py:45,103,297)
  * `numpy.math` → the `math` module (video_attacks.py:72 calls `np.math.exp`, removed in numpy 2)

No reference source is copied or edited.
"""
import contextlib
import importlib.util
import io
import os
import sys
import types

import torch
import torchvision

REFERENCE_ROOT = os.environ.get("I2V_REFERENCE_ROOT", "/root/reference")

# BASELINE.json names ResNet-50 / DenseNet-121 where the reference constructs resnet101 / densenet161
# (SURVEY.md D2).  `arch_map` selects which; truncated at depth <= 2 the two ResNets are the same net.
DEFAULT_ARCH_MAP = {"resnet101": "resnet50", "densenet161": "densenet121"}
WEIGHT_SEED = 0

_loaded = {}


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "image_attacks.py"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def seeded_ctor(arch, seed=WEIGHT_SEED):
    """torchvision constructor with `weights=None` after `torch.manual_seed(seed)`; shared with the
    new implementation so both sides see identical random-init weights."""
    def ctor(pretrained=False, **kw):
        gen_state = torch.random.get_rng_state()
        torch.manual_seed(seed)
        try:
            return _ORIG[arch](weights=None)
        finally:
            torch.random.set_rng_state(gen_state)
    return ctor


_ORIG = {}


def install_shims(arch_map=None, cpu_cuda_identity=None):
    """Idempotently install the torchvision / .cuda() shims.  Returns nothing."""
    arch_map = DEFAULT_ARCH_MAP if arch_map is None else arch_map
    tvm = torchvision.models
    for name in ("resnet101", "resnet50", "vgg16", "squeezenet1_1", "alexnet", "densenet161", "densenet121"):
        if name not in _ORIG:
            _ORIG[name] = getattr(tvm, name)
    for name in ("resnet101", "vgg16", "squeezenet1_1", "alexnet", "densenet161"):
        setattr(tvm, name, seeded_ctor(arch_map.get(name, name)))
    if cpu_cuda_identity is None:
        cpu_cuda_identity = not torch.cuda.is_available()
    if cpu_cuda_identity and not getattr(torch.Tensor, "_i2v_cuda_identity", False):
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        torch.Tensor._i2v_cuda_identity = True


def _load_as(alias, filename, extra_modules):
    saved = {k: sys.modules.get(k) for k in extra_modules}
    sys.modules.update(extra_modules)
    old_flag = sys.dont_write_bytecode
    sys.dont_write_bytecode = True  # /root/reference is read-only
    try:
        spec = importlib.util.spec_from_file_location(alias, os.path.join(REFERENCE_ROOT, filename))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[alias] = mod
        with contextlib.redirect_stdout(io.StringIO()):
            spec.loader.exec_module(mod)
        return mod
    finally:
        sys.dont_write_bytecode = old_flag
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def load(arch_map=None):
    """Return a namespace with the reference modules: .image_attacks, .TPAMI_attack, .base_attacks, .video_attacks, .utils"""
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    if "ns" in _loaded:
        return _loaded["ns"]
    install_shims(arch_map)
    gluon = {
        "gluoncv": _stub("gluoncv"),
        "gluoncv.torch": _stub("gluoncv.torch"),
        "gluoncv.torch.engine": _stub("gluoncv.torch.engine"),
        "gluoncv.torch.engine.config": _stub("gluoncv.torch.engine.config", get_cfg_defaults=lambda: None),
    }
    ref_utils = _load_as("ref_utils", "utils.py", gluon)
    cam = {
        "image_cam": _stub("image_cam", GradCAM=object),
        "image_cam_utils": _stub(
            "image_cam_utils",
            find_alexnet_layer=None, find_vgg_layer=None, find_resnet_layer=None,
            find_densenet_layer=None, find_squeezenet_layer=None),
        "timm": _stub("timm"),
        "timm.models": _stub("timm.models", create_model=None),
        "utils": ref_utils,
    }
    ns = types.SimpleNamespace(
        utils=ref_utils,
        image_attacks=_load_as("ref_image_attacks", "image_attacks.py", cam),
        TPAMI_attack=_load_as("ref_TPAMI_attack", "TPAMI_attack.py", cam),
        base_attacks=_load_as("ref_base_attacks", "base_attacks.py", cam),
    )
    import math
    import numpy
    if not hasattr(numpy, "math"):
        numpy.math = math                                   # video_attacks.py:72 (np.math.exp)
    ns.video_attacks = _load_as("ref_video_attacks", "video_attacks.py", dict(cam, base_attacks=ns.base_attacks))
    _loaded["ns"] = ns
    return ns


@contextlib.contextmanager
def quiet():
    """The reference prints the hooked layer and the cost every step (image_attacks.py:285, 349)."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield


class AdamSpy:
    """Wraps torch.optim.Adam so a test can read (modifier, grad, exp_avg, exp_avg_sq) every step of a
    reference attack without editing it — the teacher-forcing tap described in SURVEY.md Appendix D.6."""

    def __init__(self, keep=("grad", "param", "exp_avg", "exp_avg_sq")):
        self.records = []
        self.keep = keep
        self._orig = None

    def __enter__(self):
        spy = self
        orig = torch.optim.Adam
        self._orig = orig

        class SpiedAdam(orig):
            def step(self, closure=None):
                p = self.param_groups[0]["params"][0]
                rec = {}
                if "grad" in spy.keep:
                    rec["grad"] = p.grad.detach().clone()
                if "param_before" in spy.keep:
                    rec["param_before"] = p.detach().clone()
                    st0 = self.state.get(p, {})
                    rec["exp_avg_before"] = st0["exp_avg"].detach().clone() if "exp_avg" in st0 else torch.zeros_like(p)
                    rec["exp_avg_sq_before"] = st0["exp_avg_sq"].detach().clone() if "exp_avg_sq" in st0 else torch.zeros_like(p)
                out = super().step(closure)
                st = self.state[p]
                if "param" in spy.keep:
                    rec["param"] = p.detach().clone()
                if "exp_avg" in spy.keep:
                    rec["exp_avg"] = st["exp_avg"].detach().clone()
                if "exp_avg_sq" in spy.keep:
                    rec["exp_avg_sq"] = st["exp_avg_sq"].detach().clone()
                spy.records.append(rec)
                return out

        torch.optim.Adam = SpiedAdam
        return self

    def __exit__(self, *exc):
        torch.optim.Adam = self._orig
        return False
