"""Engine selection for the image-guided attacks.

'native' — this repo's own sm_100a convolution kernels (engine_native.py), the product path.
'cudnn'  — torchvision modules through PyTorch/cuDNN, truncated + input-gradient only
           (engine_cudnn.py): the library baseline the native kernels are benchmarked against
           ('cudnn_tf32': cudnn.allow_tf32 = True; suffix '_cl': channels_last weights and activations).
Both feed the same K1/K2/K3 kernels.  Choose with the `engine=` keyword of the attack classes or
$I2V_ENGINE; there is no silent switching — asking the native engine for a graph it does not
implement raises.
"""
import os

DEFAULT_ENGINE = "native"


def resolve(engine=None):
    name = engine or os.environ.get("I2V_ENGINE") or DEFAULT_ENGINE
    if name not in ("cudnn", "cudnn_tf32", "cudnn_cl", "cudnn_tf32_cl", "native", "native_tf32"):
        raise ValueError("engine must be one of cudnn, cudnn_tf32, cudnn_cl, cudnn_tf32_cl, native, native_tf32 (got %r)"
                         % (name,))
    return name


def make_engine(model, model_name, depth, engine=None):
    name = resolve(engine)
    if name.startswith("cudnn"):
        from .engine_cudnn import CudnnEngine
        return CudnnEngine(model, model_name, depth, allow_tf32="tf32" in name, channels_last=name.endswith("_cl"))
    from . import backbones
    if backbones.family_of(model_name) == "densenet":
        from .engine_densenet import DenseNetEngine
        return DenseNetEngine(model, model_name, depth, tf32x3=not name.endswith("tf32"))
    from .engine_native import NativeEngine
    return NativeEngine(model, model_name, depth, tf32x3=not name.endswith("tf32"))
