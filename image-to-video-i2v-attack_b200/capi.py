"""ctypes binding of the C ABI declared in include/i2v_b200.h.

Host code stays Python/PyTorch (as the reference is); torch is used for device memory and streams
only — every function here hands raw device pointers and the current CUDA stream to
libi2v_b200.so.  There is no fallback: a missing library or a non-CUDA tensor raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("I2V_B200_LIB", os.path.join(_HERE, "libi2v_b200.so"))

_c_i64 = ctypes.c_int64
_c_int = ctypes.c_int
_c_f = ctypes.c_float
_c_d = ctypes.c_double
_c_p = ctypes.c_void_p

# name -> argtypes; must list every function include/i2v_b200.h declares (tests/test_capi_symbols.py)
SIGNATURES = {
    "i2v_version": ([], _c_int),
    "i2v_last_error": ([], ctypes.c_char_p),
    "i2v_device_check": ([_c_int], _c_int),
    "i2v_denorm_f32": ([_c_p, _c_p, _c_i64, _c_i64, _c_int, _c_p], _c_int),
    "i2v_normalize_f32": ([_c_p, _c_p, _c_i64, _c_i64, _c_int, _c_p], _c_int),
    "i2v_compose_norm_f32": ([_c_p, _c_p, _c_p, _c_i64, _c_i64, _c_int, _c_f, _c_p], _c_int),
    "i2v_fill_f32": ([_c_p, _c_f, _c_i64, _c_p], _c_int),
    "i2v_adam_compose_f32": ([_c_p] * 6 + [_c_i64, _c_i64, _c_int, _c_f, _c_d, _c_d, _c_d, _c_d, _c_int, _c_p], _c_int),
    "i2v_set_adam_arithmetic": ([_c_int], _c_int),
    "i2v_get_adam_arithmetic": ([], _c_int),
    "i2v_adam_step_table": ([_c_p, _c_int, _c_d, _c_d, _c_d], _c_int),
    "i2v_adam_compose_table_f32": ([_c_p] * 6 + [_c_i64, _c_i64, _c_int, _c_f, _c_f, _c_f, _c_f, _c_f, _c_p, _c_p, _c_p], _c_int),
    "i2v_step_advance": ([_c_p, _c_p], _c_int),
    "i2v_sign_step_project_f32": ([_c_p, _c_p, _c_p, _c_i64, _c_i64, _c_int, _c_f, _c_f, _c_int, _c_p], _c_int),
    "i2v_frame_absmean_f32": ([_c_p, _c_p, _c_int, _c_int, _c_int, _c_i64, _c_int, _c_p], _c_int),
    "i2v_mi_sign_step_project_f32": ([_c_p] * 5 + [_c_int, _c_int, _c_int, _c_i64, _c_int, _c_f, _c_f, _c_f, _c_p], _c_int),
    "i2v_cosine_loss_grad_f32": ([_c_p, _c_p, _c_p, _c_p, _c_i64, _c_i64, _c_p, _c_f, _c_int, _c_p], _c_int),
    "i2v_layer_reweight_f32": ([_c_p, _c_p, _c_int, _c_f, _c_p, _c_p, _c_p, _c_p], _c_int),
    "i2v_layer_sums_f32": ([_c_p, _c_p, _c_p, _c_p, _c_p, _c_int, _c_i64, _c_int, _c_int, _c_p], _c_int),
    "i2v_depthwise_stencil_f32": ([_c_p, _c_p, _c_i64, _c_int, _c_int, _c_int, _c_p, _c_int, _c_int, _c_int, _c_p], _c_int),
    "i2v_temporal_shift_stack_f32": ([_c_p, _c_p, _c_i64, _c_int, _c_i64, _c_p, _c_int, _c_p], _c_int),
    "i2v_temporal_combine_f32": ([_c_p, _c_p, _c_p, _c_int, _c_d, _c_p, _c_i64, _c_int, _c_i64, _c_p], _c_int),
    "i2v_ila_workspace_doubles": ([], _c_int),
    "i2v_ila_loss_f32": ([_c_p, _c_p, _c_p, _c_i64, _c_f, _c_p, _c_p, _c_p, _c_p, _c_int, _c_p], _c_int),
    "i2v_ila_grad_f32": ([_c_p, _c_p, _c_p, _c_p, _c_i64, _c_p, _c_p], _c_int),
    "i2v_sign_descent_compose_f32": ([_c_p, _c_p, _c_p, _c_p, _c_i64, _c_i64, _c_int, _c_f, _c_f, _c_p], _c_int),
    "i2v_std_workspace_doubles": ([], _c_int),
    "i2v_std_accumulate_f32": ([_c_p, _c_i64, _c_p, _c_p, _c_p], _c_int),
    "i2v_std_finalize_f32": ([_c_p, _c_i64, _c_p, _c_p, _c_p, _c_int, _c_p], _c_int),
    "i2v_std_grad_f32": ([_c_p, _c_p, _c_i64, _c_p, _c_int, _c_p], _c_int),
    "i2v_conv_fwd_simt_f32": ([_c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_int, _c_p], _c_int),
    "i2v_conv_dgrad_simt_f32": ([_c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_int, _c_p], _c_int),
    "i2v_conv_stem_supported": ([_c_p], _c_int),
    "i2v_conv_stem_fwd_f32": ([_c_p, _c_p, _c_p, _c_p, _c_p, _c_int, _c_p], _c_int),
    "i2v_conv_stem_dgrad_f32": ([_c_p, _c_p, _c_p, _c_p, _c_p], _c_int),
    "i2v_conv_stem_dgrad_tc_f32": ([_c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p], _c_int),
    "i2v_conv_stem_dgrad_tc_group": ([_c_p], _c_int),
    "i2v_conv_stem_dgrad_tc_rows": ([_c_int], _c_int),
    "i2v_conv_stem_fwd_tc_group": ([_c_p], _c_int),
    "i2v_conv_stem_fwd_direct_supported": ([_c_p], _c_int),
    "i2v_conv_stem_fwd_direct_scratch_floats": ([_c_p], _c_i64),
    "i2v_conv_stem_fwd_direct_f32": ([_c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_int, _c_p], _c_int),
    "i2v_conv_stem_dgrad_direct_supported": ([_c_p], _c_int),
    "i2v_conv_stem_dgrad_direct_f32": ([_c_p, _c_p, _c_p, _c_p, _c_p, _c_p], _c_int),
    "i2v_conv_stem_dgrad_pool_supported": ([_c_p, _c_int, _c_int], _c_int),
    "i2v_conv_stem_dgrad_pool_f32": ([_c_p, _c_int, _c_int, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p], _c_int),
    "i2v_conv_stem_fwd_rows_supported": ([_c_p], _c_int),
    "i2v_conv_stem_fwd_rows_f32": ([_c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_int, _c_p], _c_int),
    "i2v_conv_stem_fwd_pool_supported": ([_c_p, _c_int, _c_int], _c_int),
    "i2v_conv_stem_fwd_pool_f32": ([_c_p, _c_int, _c_int, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_int, _c_p], _c_int),
    "i2v_conv_stem_fwd_tc_f32": ([_c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_int, _c_p], _c_int),
    "i2v_conv_tc_supported": ([_c_p, _c_int], _c_int),
    "i2v_conv_tc_set_trace": ([_c_p, _c_int], _c_int),
    "i2v_conv_tc_set_pair_minkit": ([_c_int], _c_int),
    "i2v_conv_tc_set_halo_mode": ([_c_int], _c_int),
    "i2v_mma_shift_probe": ([_c_int, _c_int, _c_p, _c_p, _c_p, _c_p], _c_int),
    "i2v_mma_probe": ([_c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_p, _c_p], _c_int),
    "i2v_conv_tc_f32": ([_c_p, _c_int, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_int, _c_p], _c_int),
    "i2v_conv_tc_bits_f32": ([_c_p, _c_int, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_int, _c_p], _c_int),
    "i2v_conv_tc_dgrad_class_f32": ([_c_p, _c_int, _c_int, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p], _c_int),
    "i2v_conv_tc_dgrad_class_bits_f32": ([_c_p, _c_int, _c_int, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p], _c_int),
    "i2v_conv_tc_dual_f32": ([_c_p, _c_p, _c_int, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_int, _c_p], _c_int),
    "i2v_maxpool_fwd_f32": ([_c_p, _c_p, _c_p] + [_c_int] * 9 + [_c_p], _c_int),
    "i2v_maxpool_fwd_flags_f32": ([_c_p, _c_p, _c_p] + [_c_int] * 10 + [_c_p], _c_int),
    "i2v_maxpool_bwd_f32": ([_c_p, _c_p, _c_p, _c_p] + [_c_int] * 10 + [_c_p], _c_int),
    "i2v_copy_channels_f32": ([_c_p, _c_p, _c_i64] + [_c_int] * 6 + [_c_p], _c_int),
    "i2v_bn_relu_f32": ([_c_p, _c_i64, _c_int, _c_int, _c_int, _c_p, _c_p, _c_p, _c_p], _c_int),
    "i2v_avgpool2_fwd_f32": ([_c_p, _c_p] + [_c_int] * 6 + [_c_p], _c_int),
    "i2v_avgpool2_bwd_f32": ([_c_p, _c_p] + [_c_int] * 6 + [_c_p], _c_int),
}

EPI_RELU = 1
LAYOUT_X_NCHW = 16


class ConvDesc(ctypes.Structure):
    """i2v_conv_desc"""
    _fields_ = [(n, ctypes.c_int32) for n in ("N", "H", "W", "Cin", "Cout", "R", "S", "stride", "pad", "P", "Q")]

_lib = None


class I2VError(RuntimeError):
    pass


def load():
    """dlopen libi2v_b200.so (no CUDA call is made) and attach the signatures."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise I2VError(
            "libi2v_b200.so not found at %s — build it with `python -m i2v_b200.build` "
            "(there is no CPU or PyTorch fallback for these kernels)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (argtypes, restype) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


# Launch accounting for bench.py: every successful kernel-launching call bumps LAUNCHES[name]; with
# PROFILE_EVENTS set to a list, (name, start_event, end_event, bytes) tuples are appended so the
# benchmark can time each kernel on the launching stream with CUDA events (bytes / flops = ALGORITHMIC figures of
# tensors read + written once, 2 x MACs of the convolution).
LAUNCHES = {}
PROFILE_EVENTS = None
_NO_KERNEL = ("i2v_set_adam_arithmetic", "i2v_conv_stem_dgrad_pool_supported", "i2v_conv_stem_fwd_rows_supported", "i2v_conv_stem_fwd_pool_supported", "i2v_conv_stem_dgrad_direct_supported", "i2v_conv_tc_set_pair_minkit", "i2v_conv_tc_set_halo_mode", "i2v_conv_stem_fwd_direct_supported", "i2v_conv_stem_fwd_direct_scratch_floats", "i2v_conv_stem_dgrad_tc_rows", "i2v_device_check", "i2v_std_workspace_doubles", "i2v_ila_workspace_doubles", "i2v_adam_step_table", "i2v_conv_tc_supported", "i2v_conv_stem_supported",
              "i2v_conv_tc_set_trace")


class _Timed:
    __slots__ = ("name", "nbytes", "flops", "ev", "detail")

    def __init__(self, name, nbytes=0, flops=0, detail=None):
        self.name = name
        self.nbytes = nbytes
        self.flops = flops
        self.ev = None
        self.detail = detail

    def __enter__(self):
        if PROFILE_EVENTS is not None:
            self.ev = torch.cuda.Event(enable_timing=True)
            self.ev.record()
        return self

    def __exit__(self, *exc):
        if self.ev is not None and exc[0] is None:
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            PROFILE_EVENTS.append((self.name, self.ev, end, self.nbytes, self.flops, self.detail))
        return False


def _check(rc, what):
    if rc == 0 and what not in _NO_KERNEL:
        LAUNCHES[what] = LAUNCHES.get(what, 0) + 1
    if rc != 0:
        msg = load().i2v_last_error()
        raise I2VError("%s failed (%d): %s" % (what, rc, msg.decode(errors="replace") if msg else "?"))


def _dev(t, dtype=torch.float32, name="tensor"):
    if t is None:
        return None
    if not t.is_cuda:
        raise I2VError("%s must live on a CUDA device (no CPU path exists)" % name)
    if t.dtype != dtype:
        raise I2VError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise I2VError("%s must be contiguous" % name)
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


_checked_devices = set()


def device_check(device=None):
    idx = torch.cuda.current_device() if device is None else torch.device(device).index
    if idx not in _checked_devices:
        _check(load().i2v_device_check(idx), "i2v_device_check")
        _checked_devices.add(idx)


def version():
    return load().i2v_version()


# ------------------------------------------------------------------------------- K3 family
def denorm(inp, out, inner, channels=3):
    _check(load().i2v_denorm_f32(_dev(inp), _dev(out), inp.numel(), inner, channels, _stream()), "i2v_denorm_f32")
    return out


def normalize(x, out, inner, channels=3):
    _check(load().i2v_normalize_f32(_dev(x), _dev(out), x.numel(), inner, channels, _stream()), "i2v_normalize_f32")
    return out


def compose_norm(x, mod, out, eps, inner, channels=3):
    _check(load().i2v_compose_norm_f32(_dev(x), _dev(mod), _dev(out), x.numel(), inner, channels, eps, _stream()),
           "i2v_compose_norm_f32")
    return out


def fill(t, value):
    _check(load().i2v_fill_f32(_dev(t), value, t.numel(), _stream()), "i2v_fill_f32")
    return t


def adam_compose(g, m, v, mod, x, next_img, eps, inner, step, lr, beta1=0.9, beta2=0.999, adam_eps=1e-8, channels=3):
    _check(load().i2v_adam_compose_f32(_dev(g), _dev(m), _dev(v), _dev(mod), _dev(x), _dev(next_img), g.numel(), inner,
                                       channels, eps, lr, beta1, beta2, adam_eps, step, _stream()),
           "i2v_adam_compose_f32")


def set_adam_arithmetic(which):
    """'cuda' (default): K3a reproduces torch's CUDA foreach Adam bit for bit — the optimiser the reference hits;
    'cpu': torch's CPU Adam, the arithmetic of the committed CPU fixtures and of oracle.adam_compose(arith='cpu')."""
    if which not in ("cuda", "cpu"):
        raise ValueError("adam arithmetic must be 'cuda' or 'cpu', got %r" % (which,))
    _check(load().i2v_set_adam_arithmetic(1 if which == "cuda" else 0), "i2v_set_adam_arithmetic")


def get_adam_arithmetic():
    return "cuda" if load().i2v_get_adam_arithmetic() else "cpu"


def adam_step_table(steps, lr, beta1=0.9, beta2=0.999):
    """Host float32 tensor [steps, 2] of (sqrt(1-beta2^t), -lr/(1-beta1^t)) for t = 1..steps."""
    table = torch.empty(max(steps, 1), 2, dtype=torch.float32)
    _check(load().i2v_adam_step_table(table.data_ptr(), steps, lr, beta1, beta2), "i2v_adam_step_table")
    return table


def adam_compose_table(g, m, v, mod, x, next_img, eps, inner, step_table, step_idx, beta1=0.9, beta2=0.999,
                       adam_eps=1e-8, channels=3):
    import numpy as np
    w1 = float(np.float32(1.0 - beta1))
    b2 = float(np.float32(beta2))
    a2 = float(np.float32(1.0 - beta2))
    ae = float(np.float32(adam_eps))
    with _Timed("i2v_adam_compose_table_f32", 36 * g.numel()):
        _check(load().i2v_adam_compose_table_f32(_dev(g), _dev(m), _dev(v), _dev(mod), _dev(x), _dev(next_img),
                                                 g.numel(), inner, channels, eps, w1, b2, a2, ae, _dev(step_table),
                                                 _dev(step_idx, torch.int32), _stream()),
               "i2v_adam_compose_table_f32")


def step_advance(step_idx):
    _check(load().i2v_step_advance(_dev(step_idx, torch.int32), _stream()), "i2v_step_advance")


def sign_step_project(adv, g, x, step_size, eps, inner, project=True, channels=3):
    _check(load().i2v_sign_step_project_f32(_dev(adv), _dev(g), _dev(x), adv.numel(), inner, channels, step_size, eps,
                                            1 if project else 0, _stream()), "i2v_sign_step_project_f32")
    return adv


def frame_absmean(g, norm, clip_level=False):
    B, C, T, H, W = g.shape
    _check(load().i2v_frame_absmean_f32(_dev(g), _dev(norm), B, C, T, H * W, int(clip_level), _stream()),
           "i2v_frame_absmean_f32")
    return norm


def mi_sign_step_project(adv, g, momentum, norm, x, decay, step_size, eps, clip_level=False):
    B, C, T, H, W = g.shape
    _check(load().i2v_mi_sign_step_project_f32(_dev(adv), _dev(g), _dev(momentum), _dev(norm), _dev(x), B, C, T, H * W,
                                               int(clip_level), decay, step_size, eps, _stream()),
           "i2v_mi_sign_step_project_f32")
    return adv


# ------------------------------------------------------------------------------- K1 / K2
def cosine_loss_grad(a, b, grad_a, cos_out, w_dev=None, w_host=1.0, relu_mask=False):
    """a, b: [N, ...] feature maps (flattened per frame); grad_a may be None (loss only)."""
    N = a.shape[0]
    D = a.numel() // max(N, 1)
    if b.numel() != a.numel():
        raise I2VError("feature maps differ in size")
    with _Timed("i2v_cosine_loss_grad_f32", (12 if grad_a is not None else 8) * a.numel()):
        _check(load().i2v_cosine_loss_grad_f32(_dev(a), _dev(b), _dev(grad_a), _dev(cos_out), N, D, _dev(w_dev), w_host,
                                               int(relu_mask), _stream()), "i2v_cosine_loss_grad_f32")


def layer_reweight(coeffs, prev, momentum, w_out=None, weights_log=None, step_idx=None):
    _check(load().i2v_layer_reweight_f32(_dev(coeffs), _dev(prev), coeffs.numel(), momentum, _dev(w_out),
                                         _dev(weights_log), _dev(step_idx, torch.int32), _stream()),
           "i2v_layer_reweight_f32")


def layer_sums(cos, coeffs=None, prev=None, cost_log=None, step_idx=None, mode=0, coef_CE=False):
    L, N = cos.shape
    _check(load().i2v_layer_sums_f32(_dev(cos), _dev(coeffs), _dev(prev), _dev(cost_log), _dev(step_idx, torch.int32),
                                     L, N, mode, int(coef_CE), _stream()), "i2v_layer_sums_f32")


# ------------------------------------------------------------------------------- K4 / K5 accounting
def _conv_cost(d):
    """(algorithmic bytes, flops) of one convolution launch: input + output once, 2 x MACs."""
    nin = d.N * d.H * d.W * d.Cin
    nout = d.N * d.P * d.Q * d.Cout
    flops = 2.0 * nout * d.Cin * d.R * d.S
    return 4 * (nin + nout), flops


# ------------------------------------------------------------------------------- K7 (TI gradient smoothing)
def depthwise_stencil(src, dst, kernel):
    """src/dst [B,C,T,H,W] contiguous; kernel [kt,kh,kw] (or [kh,kw]: per-frame 2-D) float32 device tensor."""
    B, C, T, H, W = src.shape
    if kernel.dim() == 2:
        kernel = kernel.unsqueeze(0)
    kt, kh, kw = kernel.shape
    _check(load().i2v_depthwise_stencil_f32(_dev(src), _dev(dst), B * C, T, H, W, _dev(kernel), kt, kh, kw, _stream()),
           "i2v_depthwise_stencil_f32")
    return dst


# ------------------------------------------------------------------------------- K8 (temporal translation)
def _host_i32(values):
    import numpy as np
    return np.ascontiguousarray(values, dtype=np.int32)


def temporal_shift_stack(adv, out, moves):
    """adv [B,C,T,H,W] -> out [D,B,C,T,H,W]: frame t of variant d goes to frame (t + moves[d]) mod T
    (reference video_attacks.py:93-105, one variant per entry of cycle_move_list)."""
    B, C, T, H, W = adv.shape
    mv = _host_i32(moves)
    if tuple(out.shape) != (len(mv),) + tuple(adv.shape):
        raise I2VError("out must be [D,B,C,T,H,W]")
    with _Timed("i2v_temporal_shift_stack_f32", 4 * adv.numel() * (1 + len(mv))):
        _check(load().i2v_temporal_shift_stack_f32(_dev(adv), _dev(out), B * C, T, H * W, mv.ctypes.data, len(mv), _stream()),
               "i2v_temporal_shift_stack_f32")
    return out


def temporal_combine(grads, kernel, moves, weight, out):
    """grads [D,B,C,T,H,W] -> out [B,C,T,H,W] = (1-weight) * sum_d k_d G_d[t] + weight * sum_d k_d G_d[(t+moves[d]) mod T]
    (reference video_attacks.py:163-177); `kernel` and `moves` are HOST sequences of length D."""
    import numpy as np
    D, B, C, T, H, W = grads.shape
    mv = _host_i32(moves)
    k = np.ascontiguousarray(kernel, dtype=np.float32).reshape(-1)
    if len(mv) != D or k.size != D:
        raise I2VError("kernel / moves must have one entry per gradient variant")
    with _Timed("i2v_temporal_combine_f32", 4 * out.numel() * (1 + D)):
        _check(load().i2v_temporal_combine_f32(_dev(grads), k.ctypes.data, mv.ctypes.data, D, float(weight), _dev(out), B * C, T,
                                               H * W, _stream()), "i2v_temporal_combine_f32")
    return out


# ------------------------------------------------------------------------------- K9 (ILAF loss) / K3d (ILAF update)
def ila_workspace(device):
    return torch.empty(load().i2v_ila_workspace_doubles(), device=device, dtype=torch.float64)


def ila_loss(f, f_ori, d0, init_norm, workspace, stats, cost_log=None, step_idx=None, add_to_cost=False):
    """stats[0:4] = cA, cB, loss, |f - f_ori| of one hooked layer (reference image_attacks.py:596-611)."""
    with _Timed("i2v_ila_loss_f32", 12 * f.numel()):
        _check(load().i2v_ila_loss_f32(_dev(f), _dev(f_ori), _dev(d0), f.numel(), float(init_norm),
                                       _dev(workspace, torch.float64), _dev(stats), _dev(cost_log),
                                       _dev(step_idx, torch.int32), int(add_to_cost), _stream()), "i2v_ila_loss_f32")


def ila_grad(f, f_ori, d0, grad, stats):
    with _Timed("i2v_ila_grad_f32", 16 * f.numel()):
        _check(load().i2v_ila_grad_f32(_dev(f), _dev(f_ori), _dev(d0), _dev(grad), f.numel(), _dev(stats), _stream()),
               "i2v_ila_grad_f32")
    return grad


def sign_descent_compose(g, mod, x, next_img, eps, step_size, inner, channels=3):
    """modifier -= step_size * sign(dcost/dmodifier) through the compose block, then the next true_image
    (reference image_attacks.py:615-617, 582-585)."""
    with _Timed("i2v_sign_descent_compose_f32", 20 * g.numel()):
        _check(load().i2v_sign_descent_compose_f32(_dev(g), _dev(mod), _dev(x), _dev(next_img), g.numel(), inner, channels, eps,
                                                   step_size, _stream()), "i2v_sign_descent_compose_f32")


# ------------------------------------------------------------------------------- K6 (dispersion reduction)
def std_workspace(device):
    return torch.empty(load().i2v_std_workspace_doubles(), device=device, dtype=torch.float64)


def std_accumulate(a, workspace, acc):
    """acc[0] += sum(a), acc[1] += sum(a*a) (float64 [2] device tensor)."""
    with _Timed("i2v_std_accumulate_f32", 4 * a.numel()):
        _check(load().i2v_std_accumulate_f32(_dev(a), a.numel(), _dev(workspace, torch.float64), _dev(acc, torch.float64),
                                             _stream()), "i2v_std_accumulate_f32")


def std_finalize(acc, n_total, stats, cost_log=None, step_idx=None, add_to_cost=False):
    """stats[0:3] = mean, unbiased std, 1/((n-1) std); cost_log[step] (+)= std."""
    _check(load().i2v_std_finalize_f32(_dev(acc, torch.float64), int(n_total), _dev(stats), _dev(cost_log),
                                       _dev(step_idx, torch.int32), int(add_to_cost), _stream()), "i2v_std_finalize_f32")


def std_grad(a, grad, stats, relu_mask=False):
    with _Timed("i2v_std_grad_f32", 8 * a.numel()):
        _check(load().i2v_std_grad_f32(_dev(a), _dev(grad), a.numel(), _dev(stats), int(bool(relu_mask)), _stream()),
               "i2v_std_grad_f32")


# ------------------------------------------------------------------------------- K4 / K5 (CUDA-core path)
def conv_fwd_simt(desc, x, bmat, bias, residual, y, relu=False, x_nchw=False):
    flags = (EPI_RELU if relu else 0) | (LAYOUT_X_NCHW if x_nchw else 0)
    _check(load().i2v_conv_fwd_simt_f32(ctypes.addressof(desc), _dev(x), _dev(bmat), _dev(bias), _dev(residual), _dev(y),
                                        flags, _stream()), "i2v_conv_fwd_simt_f32")


def conv_dgrad_simt(desc, dy, bmat, addend, mask_src, dx, x_nchw=False):
    flags = LAYOUT_X_NCHW if x_nchw else 0
    _check(load().i2v_conv_dgrad_simt_f32(ctypes.addressof(desc), _dev(dy), _dev(bmat), _dev(addend), _dev(mask_src),
                                          _dev(dx), flags, _stream()), "i2v_conv_dgrad_simt_f32")


def conv_stem_supported(desc):
    return bool(load().i2v_conv_stem_supported(ctypes.addressof(desc)))


def conv_stem_fwd(desc, x, w, bias, y, relu=True):
    nb, fl = _conv_cost(desc)
    with _Timed("i2v_conv_stem_fwd_f32", nb, fl):
        _check(load().i2v_conv_stem_fwd_f32(ctypes.addressof(desc), _dev(x), _dev(w), _dev(bias), _dev(y),
                                            EPI_RELU if relu else 0, _stream()), "i2v_conv_stem_fwd_f32")


def conv_stem_dgrad(desc, dy, w, dx):
    nb, fl = _conv_cost(desc)
    with _Timed("i2v_conv_stem_dgrad_f32", nb, fl):
        _check(load().i2v_conv_stem_dgrad_f32(ctypes.addressof(desc), _dev(dy), _dev(w), _dev(dx), _stream()),
               "i2v_conv_stem_dgrad_f32")


def conv_stem_dgrad_tc(desc, dy, wz_hi, wz_lo, z_scratch, dx):
    """First-layer data gradient as a tcgen05 GEMM over the output channels + col2im (see include/i2v_b200.h)."""
    nb, fl = _conv_cost(desc)
    with _Timed("i2v_conv_stem_dgrad_f32", nb, fl):
        _check(load().i2v_conv_stem_dgrad_tc_f32(ctypes.addressof(desc), _dev(dy), _dev(wz_hi), _dev(wz_lo), _dev(z_scratch),
                                                 _dev(dx), _stream()), "i2v_conv_stem_dgrad_tc_f32")


def conv_stem_dgrad_direct_supported(desc):
    return bool(load().i2v_conv_stem_dgrad_direct_supported(ctypes.addressof(desc)))


def conv_stem_dgrad_direct(desc, dy, wd_hi, wd_lo, dx):
    """First-layer data gradient without scratch (see include/i2v_b200.h); wd_* from stem_direct_dgrad_weights()."""
    nb, fl = _conv_cost(desc)
    with _Timed("i2v_conv_stem_dgrad_f32", nb, fl):
        _check(load().i2v_conv_stem_dgrad_direct_f32(ctypes.addressof(desc), _dev(dy), _dev(wd_hi), _dev(wd_lo), _dev(dx),
                                                     _stream()), "i2v_conv_stem_dgrad_direct_f32")


def conv_stem_dgrad_pool_supported(desc, P2, Q2):
    return bool(load().i2v_conv_stem_dgrad_pool_supported(ctypes.addressof(desc), int(P2), int(Q2)))


def conv_stem_dgrad_pool(desc, dy_pooled, argmax, wd_hi, wd_lo, dx):
    """Max-pool backward (3x3 / stride 2 / pad 1) + first-layer data gradient in one kernel (see include/i2v_b200.h):
    dy_pooled / argmax = [N, P2, Q2, 64] f32 / u8."""
    n, P2, Q2, c = dy_pooled.shape
    assert c == 64 and argmax.shape == dy_pooled.shape and argmax.dtype == torch.uint8
    nb, fl = _conv_cost(desc)
    # bytes: the pooled gradient + argmax in, the image gradient out (the stem activation's gradient is never materialised)
    nb = dy_pooled.numel() * 5 + desc.N * 3 * desc.H * desc.W * 4
    with _Timed("i2v_conv_stem_dgrad_pool_f32", nb, fl):
        _check(load().i2v_conv_stem_dgrad_pool_f32(ctypes.addressof(desc), int(P2), int(Q2), _dev(dy_pooled), _dev(argmax, torch.uint8),
                                                   _dev(wd_hi), _dev(wd_lo), _dev(dx), _stream()), "i2v_conv_stem_dgrad_pool_f32")


def stem_direct_dgrad_weights(w_stem):
    """[(c,r,s) = 147, 64] -> [160, 64]: the taps followed by 13 zero rows (UMMA N = 160)."""
    assert w_stem.shape == (147, 64)
    return torch.cat([w_stem, w_stem.new_zeros(13, 64)], 0).contiguous()


def stem_dgrad_tc_rows(cols):
    """Rows of the zero-padded [(c,r,s), Cout] weight matrix i2v_conv_stem_dgrad_tc_f32 expects."""
    return int(load().i2v_conv_stem_dgrad_tc_rows(int(cols)))


def stem_dgrad_tc_scratch_floats(desc):
    g = load().i2v_conv_stem_dgrad_tc_group(ctypes.addressof(desc))
    return (3 * desc.R * desc.S + 31) // 32 * 32 * g * desc.P * desc.Q


def conv_stem_fwd_tc(desc, x, wk_hi, wk_lo, bias, col_scratch, y, relu=False):
    """First-layer forward as im2col + tcgen05 GEMM (see include/i2v_b200.h)."""
    nb, fl = _conv_cost(desc)
    with _Timed("i2v_conv_stem_fwd_f32", nb, fl):
        _check(load().i2v_conv_stem_fwd_tc_f32(ctypes.addressof(desc), _dev(x), _dev(wk_hi), _dev(wk_lo), _dev(bias),
                                               _dev(col_scratch), _dev(y), EPI_RELU if relu else 0, _stream()),
               "i2v_conv_stem_fwd_tc_f32")


def conv_stem_fwd_rows_supported(desc):
    return bool(load().i2v_conv_stem_fwd_rows_supported(ctypes.addressof(desc)))


def conv_stem_fwd_rows(desc, x, wk_hi, wk_lo, bias, y, relu=False):
    """First-layer forward without the patch matrix: one output row per tile (see include/i2v_b200.h)."""
    nb, fl = _conv_cost(desc)
    with _Timed("i2v_conv_stem_fwd_f32", nb, fl):
        _check(load().i2v_conv_stem_fwd_rows_f32(ctypes.addressof(desc), _dev(x), _dev(wk_hi), _dev(wk_lo), _dev(bias), _dev(y),
                                                 EPI_RELU if relu else 0, _stream()), "i2v_conv_stem_fwd_rows_f32")


def conv_stem_fwd_pool_supported(desc, P2, Q2):
    return bool(load().i2v_conv_stem_fwd_pool_supported(ctypes.addressof(desc), int(P2), int(Q2)))


def conv_stem_fwd_pool(desc, x, wk_hi, wk_lo, bias, pooled, argmax, relu=True, mark_dead=True):
    """First-layer forward with the 3x3 / stride-2 / pad-1 max pooling fused into its epilogue (see include/i2v_b200.h)."""
    _, fl = _conv_cost(desc)
    nb = 4 * x.numel() + 5 * pooled.numel()          # the image once, the pooled tensor and its argmax plane once
    P2, Q2 = pooled.shape[1], pooled.shape[2]
    with _Timed("i2v_conv_stem_fwd_pool_f32", nb, fl):
        _check(load().i2v_conv_stem_fwd_pool_f32(ctypes.addressof(desc), int(P2), int(Q2), _dev(x), _dev(wk_hi), _dev(wk_lo),
                                                 _dev(bias), _dev(pooled), _dev(argmax, torch.uint8),
                                                 (EPI_RELU if relu else 0) | (4 if mark_dead else 0), _stream()),
               "i2v_conv_stem_fwd_pool_f32")


def conv_stem_fwd_direct_supported(desc):
    return bool(load().i2v_conv_stem_fwd_direct_supported(ctypes.addressof(desc)))


def stem_fwd_direct_scratch_floats(desc):
    return int(load().i2v_conv_stem_fwd_direct_scratch_floats(ctypes.addressof(desc)))


def conv_stem_fwd_direct(desc, x, wr_hi, wr_lo, bias, xp_scratch, y, relu=False):
    """EXPERIMENTAL first-layer forward without the patch matrix (see include/i2v_b200.h)."""
    nb, fl = _conv_cost(desc)
    with _Timed("i2v_conv_stem_fwd_f32", nb, fl):
        _check(load().i2v_conv_stem_fwd_direct_f32(ctypes.addressof(desc), _dev(x), _dev(wr_hi), _dev(wr_lo), _dev(bias),
                                                   _dev(xp_scratch), _dev(y), EPI_RELU if relu else 0, _stream()),
               "i2v_conv_stem_fwd_direct_f32")


def stem_fwd_tc_scratch_floats(desc):
    g = load().i2v_conv_stem_fwd_tc_group(ctypes.addressof(desc))
    return (3 * desc.R * desc.S + 31) // 32 * 32 * g * desc.P * desc.Q


def mma_probe(N, accs, a_tmem, count, ctas=1, issuers=1):
    """(issue cycles, completion cycles) of `count` tcgen05 TF32 MMAs (debug / measurement, see include/i2v_b200.h)."""
    out = torch.zeros(2, dtype=torch.int64, device="cuda")
    _check(load().i2v_mma_probe(N, accs, int(a_tmem), count, ctas, issuers, _dev(out, torch.int64), _stream()), "i2v_mma_probe")
    torch.cuda.synchronize()
    return tuple(int(v) for v in out.tolist())


def conv_tc_set_trace(buf, tiles=0):
    """Debug: buf = int64 CUDA tensor [tiles, 8] (or None) receiving CTA 0's pipeline time stamps."""
    _check(load().i2v_conv_tc_set_trace(None if buf is None else _dev(buf, torch.int64), tiles), "i2v_conv_tc_set_trace")


def conv_tc_set_pair_minkit(min_ksteps):
    """Tiles of >= min_ksteps k-steps run on the CTA-pair (cta_group::2) kernel; 0 = never."""
    _check(load().i2v_conv_tc_set_pair_minkit(int(min_ksteps)), "i2v_conv_tc_set_pair_minkit")


def conv_tc_set_halo_mode(mode):
    """Which 3x3 / s1 / p1 convolutions run on the patch-once kernel: 1 wherever it fits, 0 never, -1 the default rule."""
    _check(load().i2v_conv_tc_set_halo_mode(int(mode)), "i2v_conv_tc_set_halo_mode")


def conv_tc_supported(desc, dgrad):
    return bool(load().i2v_conv_tc_supported(ctypes.addressof(desc), int(dgrad)))


def conv_tc(desc, dgrad, src, w_hi, w_lo, bias, residual, mask_src, dst, relu=False, mask_bits=None):
    """Tensor-core implicit GEMM (tcgen05/TMEM/TMA).  w_lo=None -> plain TF32, else 3xTF32 FP32-parity mode.
    mask_bits: int32 [C_dst/32, M] — forward: receives the activity bits of dst; dgrad: the ReLU-backward mask."""
    nb, fl = _conv_cost(desc)
    nb += 4 * dst.numel() * ((residual is not None) + (mask_src is not None))
    if mask_bits is not None:
        nb += 4 * mask_bits.numel()
    detail = None
    if PROFILE_EVENTS is not None:
        detail = "%s %dx%d %d->%d k%ds%d%s%s%s" % ("dgrad" if dgrad else "fwd", desc.H, desc.W, desc.Cin, desc.Cout, desc.R,
                                                  desc.stride, " +res" if residual is not None else "",
                                                  " +mask" if mask_src is not None else "",
                                                  " +bits" if mask_bits is not None else "")
    with _Timed("i2v_conv_tc_f32", nb, fl, detail):
        _check(load().i2v_conv_tc_bits_f32(ctypes.addressof(desc), int(dgrad), _dev(src), _dev(w_hi), _dev(w_lo), _dev(bias),
                                           _dev(residual), _dev(mask_src), _dev(mask_bits, torch.int32), _dev(dst),
                                           EPI_RELU if relu else 0, _stream()),
               "i2v_conv_tc_f32")


def conv_tc_dual(desc, x, t, w_hi, w_lo, bias, dst, relu=True, mask_bits=None):
    """dst = act(conv1x1_stride_s(x) + conv1x1(t) + bias) as ONE GEMM over K = Cin followed by C2 (see include/i2v_b200.h):
    desc describes the first convolution (over x), t is [N, P, Q, C2], w_* = [Cout, Cin + C2] K-major."""
    C2 = t.shape[-1]
    assert t.shape[:3] == dst.shape[:3] and w_hi.shape == (desc.Cout, desc.Cin + C2)
    m = desc.N * desc.P * desc.Q
    # algorithmic bytes: the pixels of x the strided window touches, t, and the output (+ its activity bits)
    nb = 4 * m * (desc.Cin + C2 + desc.Cout)
    if mask_bits is not None:
        nb += 4 * mask_bits.numel()
    fl = 2.0 * m * desc.Cout * (desc.Cin + C2)
    detail = None
    if PROFILE_EVENTS is not None:
        detail = "fwd %dx%d %d+%d->%d k1s%d dual%s" % (desc.H, desc.W, desc.Cin, C2, desc.Cout, desc.stride,
                                                     " +bits" if mask_bits is not None else "")
    with _Timed("i2v_conv_tc_f32", nb, fl, detail):
        _check(load().i2v_conv_tc_dual_f32(ctypes.addressof(desc), _dev(x), int(C2), _dev(t), _dev(w_hi), _dev(w_lo), _dev(bias),
                                           _dev(mask_bits, torch.int32), _dev(dst), EPI_RELU if relu else 0, _stream()),
               "i2v_conv_tc_dual_f32")


def conv_tc_dgrad_class(desc, ph, pw, dy, w_hi, w_lo, addend, mask_src, dx, mask_bits=None):
    """One stride-parity class of a strided data gradient on the tensor cores (see include/i2v_b200.h); mask_bits (int32
    [Cin/32, N*H*W]) instead of mask_src selects the TMA epilogue."""
    # one stride-parity class: dy once, 1/stride^2 of dx (+ mask, + addend) and of the taps
    st2 = desc.stride * desc.stride
    nb, fl = _conv_cost(desc)
    nout = desc.N * desc.P * desc.Q * desc.Cout
    nb = 4 * nout + 4 * (dx.numel() // st2) * (1 + (addend is not None) + (mask_src is not None))
    if mask_bits is not None:
        nb += 4 * (mask_bits.numel() // st2)
    with _Timed("i2v_conv_tc_dgrad_class_f32", nb, fl / st2):
        _check(load().i2v_conv_tc_dgrad_class_bits_f32(ctypes.addressof(desc), ph, pw, _dev(dy), _dev(w_hi), _dev(w_lo),
                                                       _dev(addend), _dev(mask_src), _dev(mask_bits, torch.int32), _dev(dx),
                                                       _stream()),
               "i2v_conv_tc_dgrad_class_f32")


def maxpool_fwd(x, y, argmax, k, stride, pad, mark_dead=False):
    """mark_dead: x is a ReLU output; windows whose maximum is not > 0 get argmax = 255 (no winner) so that maxpool_bwd
    needs no ReLU-backward mask."""
    N, H, W, C = x.shape
    _, P, Q, _ = y.shape
    with _Timed("i2v_maxpool_fwd_f32", 4 * x.numel() + 5 * y.numel()):
        _check(load().i2v_maxpool_fwd_flags_f32(_dev(x), _dev(y), _dev(argmax, torch.uint8), N, H, W, C, P, Q, k, stride, pad,
                                                4 if mark_dead else 0, _stream()), "i2v_maxpool_fwd_f32")


def maxpool_bwd(dy, argmax, mask_src, dx, k, stride, pad, accumulate=False, mask_pooled=False):
    """mask_pooled: mask_src is the pooled OUTPUT y (1[y > 0] = 1[x[argmax] > 0]) instead of the input activation."""
    N, H, W, C = dx.shape
    _, P, Q, _ = dy.shape
    nb = 5 * dy.numel() + 4 * dx.numel() * (1 + bool(accumulate))
    if mask_src is not None:
        nb += 4 * mask_src.numel()
    with _Timed("i2v_maxpool_bwd_f32", nb):
        _check(load().i2v_maxpool_bwd_f32(_dev(dy), _dev(argmax, torch.uint8), _dev(mask_src), _dev(dx), N, H, W, C, P, Q,
                                          k, stride, pad, int(bool(accumulate)) | (2 if mask_pooled else 0), _stream()),
               "i2v_maxpool_bwd_f32")


def copy_channels(src, dst, src_off, dst_off, ccopy, accumulate=False):
    M = src.numel() // src.shape[-1]
    _check(load().i2v_copy_channels_f32(_dev(src), _dev(dst), M, src.shape[-1], src_off, dst.shape[-1], dst_off, ccopy,
                                        int(accumulate), _stream()), "i2v_copy_channels_f32")


# ------------------------------------------------------------------------------- DenseNet pieces
def bn_relu(src, C, scale, shift, dst):
    """dst [M, Cp] = relu(scale * src[:, :C] + shift) (zero for channels C..Cp-1); src is [M, src_ld] with src_ld >= C."""
    M = src.numel() // src.shape[-1]
    with _Timed("i2v_bn_relu_f32", 4 * M * (C + dst.shape[-1])):
        _check(load().i2v_bn_relu_f32(_dev(src), M, C, dst.shape[-1], src.shape[-1], _dev(scale), _dev(shift), _dev(dst),
                                      _stream()), "i2v_bn_relu_f32")


def avgpool2_fwd(x, y, dst_off=0):
    """x [N,H,W,C] -> y[..., dst_off : dst_off+C] of y [N,H//2,W//2,dst_ld] (2x2 / stride 2, floor mode)."""
    N, H, W, C = x.shape
    with _Timed("i2v_avgpool2_fwd_f32", 4 * x.numel() + x.numel()):
        _check(load().i2v_avgpool2_fwd_f32(_dev(x), _dev(y), N, H, W, C, y.shape[-1], dst_off, _stream()), "i2v_avgpool2_fwd_f32")


def avgpool2_bwd(dy, dx, src_off=0):
    """dx [N,H,W,C] = the gradient of avgpool2_fwd given dy[..., src_off : src_off+C] of dy [N,H//2,W//2,src_ld]."""
    N, H, W, C = dx.shape
    with _Timed("i2v_avgpool2_bwd_f32", 4 * dx.numel() + dx.numel()):
        _check(load().i2v_avgpool2_bwd_f32(_dev(dy), _dev(dx), N, H, W, C, dy.shape[-1], src_off, _stream()), "i2v_avgpool2_bwd_f32")
