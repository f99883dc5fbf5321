"""TEST INFRASTRUCTURE ONLY — Python face of the CPU oracle.

Two layers:
  * ctypes wrappers over oracle/_build/libi2v_oracle.so (plain C, oracle/i2v_oracle.c) for the
    memory-bound arithmetic: denorm, compose, Adam+compose (K3a), sign-step (K3b), MI (K3c),
    cosine loss/grad in float64 (K1 arbiter), layer re-weighting (K2).
  * full-loop restatements of the reference attack classes (image_attacks.py:294-364, 426-496,
    TPAMI_attack.py:223-320, base_attacks.py:242-340) that run a torchvision backbone on the CPU with
    torch autograd and use the C pieces for everything else.  These are pinned against the unmodified
    reference classes by tests/test_oracle_golden.py (fixtures from oracle/make_golden.py).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libi2v_oracle.so")
_lib = None

MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32)
STD = np.array([0.229, 0.224, 0.225], dtype=np.float32)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


i64 = ctypes.c_int64
ci = ctypes.c_int
cf = ctypes.c_float
cd = ctypes.c_double


def denorm(inp, inner, channels=3):
    inp = _f32(inp)
    out = np.empty_like(inp)
    lib().oracle_denorm_f32(_ptr(inp), _ptr(out), i64(inp.size), i64(inner), ci(channels))
    return out


def normalize(x, inner, channels=3):
    x = _f32(x)
    out = np.empty_like(x)
    lib().oracle_normalize_f32(_ptr(x), _ptr(out), i64(x.size), i64(inner), ci(channels))
    return out


def compose_norm(x, mod, eps, inner, channels=3):
    x, mod = _f32(x), _f32(mod)
    out = np.empty_like(x)
    lib().oracle_compose_norm_f32(_ptr(x), _ptr(mod), _ptr(out), i64(x.size), i64(inner), ci(channels), cf(eps))
    return out


def adam_compose(g, m, v, mod, x, eps, inner, step, lr, beta1=0.9, beta2=0.999, adam_eps=1e-8, channels=3, arith="cpu"):
    """In-place on copies; returns (m, v, mod, next_img).  arith: 'cpu' = torch's CPU Adam kernels (what the CPU fixtures
    were generated with), 'cuda' = torch's CUDA foreach Adam kernels (what the reference hits on its own platform)."""
    g, x = _f32(g), _f32(x)
    m, v, mod = _f32(m).copy(), _f32(v).copy(), _f32(mod).copy()
    out = np.empty_like(x)
    lib().oracle_adam_compose_f32(_ptr(g), _ptr(m), _ptr(v), _ptr(mod), _ptr(x), _ptr(out), i64(x.size), i64(inner),
                                  ci(channels), cf(eps), cd(lr), cd(beta1), cd(beta2), cd(adam_eps), ci(step),
                                  ci(1 if arith == "cuda" else 0))
    return m, v, mod, out


def sign_descent_compose(g, mod, x, eps, step_size, inner, channels=3):
    """ILAF's update block (image_attacks.py:615-617 + recompose); returns (mod, next_img)."""
    g, x = _f32(g), _f32(x)
    mod = _f32(mod).copy()
    out = np.empty_like(x)
    lib().oracle_sign_descent_compose_f32(_ptr(g), _ptr(mod), _ptr(x), _ptr(out), i64(x.size), i64(inner), ci(channels),
                                          cf(eps), cf(step_size))
    return mod, out


def temporal_shift_stack(adv, moves):
    """adv [B,C,T,H,W] -> [D,B,C,T,H,W] with frame t of variant d moved to (t + moves[d]) mod T (video_attacks.py:93-105)."""
    adv = _f32(adv)
    B, C, T, H, W = adv.shape
    mv = np.ascontiguousarray(moves, dtype=np.int32)
    out = np.empty((len(mv),) + adv.shape, dtype=np.float32)
    lib().oracle_temporal_shift_stack_f32(_ptr(adv), _ptr(out), i64(B * C), ci(T), i64(H * W),
                                          mv.ctypes.data_as(ctypes.c_void_p), ci(len(mv)))
    return out


def temporal_combine(grads, kernel, moves, weight):
    """grads [D,B,C,T,H,W] -> [B,C,T,H,W] (video_attacks.py:163-177)."""
    grads = _f32(grads)
    D, B, C, T, H, W = grads.shape
    mv = np.ascontiguousarray(moves, dtype=np.int32)
    k = _f32(kernel)
    out = np.empty(grads.shape[1:], dtype=np.float32)
    lib().oracle_temporal_combine_f32(_ptr(grads), _ptr(k), mv.ctypes.data_as(ctypes.c_void_p), ci(D), cd(weight), _ptr(out),
                                      i64(B * C), ci(T), i64(H * W))
    return out


def ila_loss_grad_f64(f, f_ori, d0, init_norm, want_grad=True):
    """ILAF per-layer loss (float64) and d loss / d f (image_attacks.py:596-611)."""
    f, f_ori, d0 = _f32(f), _f32(f_ori), _f32(d0)
    loss = ctypes.c_double()
    grad = np.empty(f.shape, dtype=np.float64) if want_grad else None
    lib().oracle_ila_loss_grad_f64(_ptr(f), _ptr(f_ori), _ptr(d0), cd(init_norm), i64(f.size), ctypes.byref(loss),
                                   grad.ctypes.data_as(ctypes.c_void_p) if want_grad else None)
    return loss.value, grad


def adam_step_scalars(lr, beta1, beta2, step):
    a, b = cf(), cf()
    lib().oracle_adam_step_scalars(cd(lr), cd(beta1), cd(beta2), ci(step), ctypes.byref(a), ctypes.byref(b))
    return a.value, b.value


def sign_step_project(adv, g, x, step_size, eps, inner, project=True, channels=3):
    adv = _f32(adv).copy()
    g = _f32(g)
    xx = _f32(x) if x is not None else None
    lib().oracle_sign_step_project_f32(_ptr(adv), _ptr(g), _ptr(xx) if xx is not None else None, i64(adv.size),
                                       i64(inner), ci(channels), cf(step_size), cf(eps), ci(1 if project else 0))
    return adv


def frame_absmean(g, clip_level=False):
    g = _f32(g)
    B, C, T, H, W = g.shape
    norm = np.empty((B,) if clip_level else (B, T), dtype=np.float32)
    lib().oracle_frame_absmean_f32(_ptr(g), _ptr(norm), ci(B), ci(C), ci(T), i64(H * W), ci(int(clip_level)))
    return norm


def mi_sign_step_project(adv, g, momentum, norm, x, decay, step_size, eps, clip_level=False):
    adv, momentum = _f32(adv).copy(), _f32(momentum).copy()
    g, norm, x = _f32(g), _f32(norm), _f32(x)
    B, C, T, H, W = g.shape
    lib().oracle_mi_sign_step_project_f32(_ptr(adv), _ptr(g), _ptr(momentum), _ptr(norm), _ptr(x), ci(B), ci(C),
                                          ci(T), i64(H * W), ci(int(clip_level)), cf(decay), cf(step_size), cf(eps))
    return adv, momentum


def cosine_loss_grad_f64(a, b, w=1.0, relu_mask=False, want_grad=True):
    a, b = _f32(a), _f32(b)
    N = a.shape[0]
    D = a.size // N
    cos = np.empty(N, dtype=np.float64)
    grad = np.empty(a.shape, dtype=np.float64) if want_grad else None
    lib().oracle_cosine_loss_grad_f64(_ptr(a), _ptr(b), _ptr(grad) if want_grad else None, _ptr(cos), i64(N), i64(D),
                                      cd(w), ci(int(relu_mask)))
    return cos, grad


def layer_reweight(coeffs, prev, momentum):
    coeffs = _f32(coeffs).copy()
    prev = _f32(prev)
    w = np.empty_like(coeffs)
    lib().oracle_layer_reweight_f32(_ptr(coeffs), _ptr(prev), ci(coeffs.size), cf(momentum), _ptr(w))
    return coeffs, w


def layer_sums(cosv, coeffs=None, mode=0, coef_CE=False):
    cosv = _f32(cosv)
    L, N = cosv.shape
    prev = np.zeros(L, dtype=np.float32)
    cost = cf()
    cz = _f32(coeffs) if coeffs is not None else None
    lib().oracle_layer_sums_f32(_ptr(cosv), _ptr(cz) if cz is not None else None, _ptr(prev), ctypes.byref(cost),
                                ci(L), i64(N), ci(mode), ci(int(coef_CE)))
    return cost.value, prev


def std_loss_grad_f64(a, relu_mask=False):
    """Dispersion-Reduction loss of reference image_attacks.py:216-220 restated in float64 numpy:
    cost = torch.Tensor.std() of the whole feature map (unbiased, n-1), and its autograd gradient
    d std / d a_i = (a_i - mean) / ((n - 1) std).  relu_mask zeroes the gradient where a <= 0 (the
    pre-activation convention of the native engine: the hooked map is a ReLU output).
    Returns (std, mean, grad float64 of a's shape)."""
    a64 = np.asarray(a, dtype=np.float64)
    n = a64.size
    mean = a64.sum() / n
    d = a64 - mean
    sd = np.sqrt((d * d).sum() / (n - 1))
    g = d / ((n - 1) * sd)
    if relu_mask:
        g = np.where(a64 > 0, g, 0.0)
    return sd, mean, g
