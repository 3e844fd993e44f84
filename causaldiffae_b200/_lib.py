"""ctypes binding of libcdae.so (include/cdae.h).  No fallback: if the library is missing or the device is not
sm_100 every op raises.  PyTorch is used only for device memory, streams and torch.distributed."""
import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libcdae.so")
_lock = threading.Lock()
_lib = None
_inited_devices = set()

MAX_SEG = 48


class Seg(C.Structure):
    _fields_ = [("src", C.c_int32), ("dh", C.c_int32), ("dw", C.c_int32), ("c0", C.c_int32), ("nchunk", C.c_int32),
                ("wk", C.c_int32)]


class IgemmDesc(C.Structure):
    _fields_ = [("src", C.c_void_p * 4), ("src_c", C.c_int32 * 4),
                ("nsrc", C.c_int32), ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("in_stride", C.c_int32),
                ("nseg", C.c_int32), ("seg", Seg * MAX_SEG),
                ("wgt", C.c_void_p), ("wrows", C.c_int32), ("wk", C.c_int32),
                ("out", C.c_void_p), ("out_mode", C.c_int32),
                ("OH", C.c_int32), ("OW", C.c_int32), ("ldo", C.c_int32), ("cout", C.c_int32),
                ("sps", C.c_int32), ("ooh", C.c_int32), ("oow", C.c_int32),
                ("bias", C.c_void_p), ("bias2", C.c_void_p),
                ("resid", C.c_void_p), ("ldr", C.c_int32),
                ("bn", C.c_int32),
                ("stats", C.c_void_p),
                ("gnb_ws", C.c_void_p), ("gnb_ab", C.c_void_p), ("gnb_x0", C.c_void_p), ("gnb_x1", C.c_void_p),
                ("gnb_c0", C.c_int32), ("gnb_ld0", C.c_int32), ("gnb_ld1", C.c_int32), ("gnb_silu", C.c_int32),
                ("bias_img", C.c_void_p), ("bias_img_ld", C.c_int32),
                ("gn_ab", C.c_void_p), ("gn_c", C.c_int32), ("gn_off", C.c_int32 * 4), ("up2x", C.c_int32)]


class WgradDesc(C.Structure):
    _fields_ = [("dy", C.c_void_p), ("ldy", C.c_int32), ("cout", C.c_int32),
                ("src", C.c_void_p), ("src_c", C.c_int32),
                ("c0", C.c_int32), ("cin", C.c_int32),
                ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("OH", C.c_int32), ("OW", C.c_int32),
                ("in_stride", C.c_int32), ("ksize", C.c_int32),
                ("dw", C.c_void_p), ("dw_ld", C.c_int32),
                ("ci_off", C.c_int32), ("cin_real", C.c_int32), ("splits", C.c_int32),
                ("dbias", C.c_void_p)]


class SgemmDesc(C.Structure):
    _fields_ = [("A", C.c_void_p), ("a_sm", C.c_int64), ("a_sk", C.c_int64), ("a_mode", C.c_int32),
                ("B", C.c_void_p), ("b_sk", C.c_int64), ("b_sn", C.c_int64), ("b_mode", C.c_int32),
                ("C", C.c_void_p), ("c_sm", C.c_int64), ("c_sn", C.c_int64), ("c_mode", C.c_int32),
                ("bias", C.c_void_p), ("act_out", C.c_int32),
                ("colstats", C.c_void_p),
                ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("splits", C.c_int32),
                ("g_sb", C.c_int64), ("g_sc", C.c_int64), ("g_sh", C.c_int64), ("g_sw", C.c_int64),
                ("g_cin", C.c_int32), ("g_h", C.c_int32), ("g_w", C.c_int32), ("g_oh", C.c_int32), ("g_ow", C.c_int32),
                ("g_ab", C.c_void_p)]


class PackEntry(C.Structure):
    _fields_ = [("src_off", C.c_int64), ("dst_fwd_off", C.c_int64), ("dst_tr_off", C.c_int64),
                ("cout", C.c_int32), ("cin", C.c_int32), ("taps", C.c_int32), ("cout_pad", C.c_int32),
                ("cin_pad", C.c_int32), ("fwd_ld", C.c_int32), ("tr_ld", C.c_int32), ("_pad", C.c_int32)]


P, I32, I64, F32, F64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double
_SIGS = {
    "cdae_version": ([], C.c_int),
    "cdae_init": ([], C.c_int),
    "cdae_q_sample": ([P, P, P, P, P, P, I64, I64, P], C.c_int),
    "cdae_mse_loss": ([P, P, P, P, F32, P, I64, I64, P], C.c_int),
    "cdae_ddim_step": ([P, P, P, F32, I32, P, P, I32, P, P, P, I64, I64, P], C.c_int),
    "cdae_adam_ema": ([P, P, I32, P, P, P, P, P, P, P, P, I64, P], C.c_int),
    "cdae_sumsq": ([P, I32, P, I64, P], C.c_int),
    "cdae_cast_bf16": ([P, P, I64, P], C.c_int),
    "cdae_ema_update": ([P, P, F32, I64, P], C.c_int),
    "cdae_zero": ([P, I64, P], C.c_int),
    "cdae_nchw_to_nhwc_pad": ([P, P, I32, I32, I32, I32, I32, P], C.c_int),
    "cdae_nhwc_to_nchw": ([P, P, I32, I32, I32, I32, I32, P], C.c_int),
    "cdae_pack_weights": ([P, P, P, I32, I64, P], C.c_int),
    "cdae_upsample2x": ([P, P, I32, I32, I32, I32, P], C.c_int),
    "cdae_sumpool2x": ([P, P, I32, I32, I32, I32, I32, P], C.c_int),
    "cdae_zero_insert2x": ([P, P, I32, I32, I32, I32, P], C.c_int),
    "cdae_colsum": ([P, P, I64, I32, I32, I32, I32, P], C.c_int),
    "cdae_dropout": ([P, I64, P, I64, P, P], C.c_int),
    "cdae_gn_fwd": ([P, I32, P, I32, I32, I32, P, P, P, I32, I32, I32, P, P, P, P], C.c_int),
    "cdae_gn_apply_fwd": ([P, I32, P, P, I32, P, I32, I32, P, P, P, I32, I32, I32, P, P, P, P, P], C.c_int),
    "cdae_gn_bwd": ([P, P, I32, P, I32, I32, I32, P, P, P, I32, I32, I32, P, P, P, P, P, I32, P, P, P, P], C.c_int),
    "cdae_gn_bwd_apply": ([P, P, I32, P, I32, I32, I32, P, P, P, I32, I32, P, P, P, P, P, P, I32, P, P, P, P], C.c_int),
    "cdae_igemm": ([C.POINTER(IgemmDesc), P], C.c_int),
    "cdae_wgrad": ([C.POINTER(WgradDesc), P], C.c_int),
    "cdae_attn_fwd": ([P, P, P, I32, I32, I32, I32, P], C.c_int),
    "cdae_attn_bwd": ([P, P, P, P, P, P, I32, I32, I32, I32, P], C.c_int),
    "cdae_dag_fwd": ([P, P, P, P, I32, I32, I32, I32, P], C.c_int),
    "cdae_gather_images": ([P, P, P, P, P, I32, I32, I32, I32, I32, I32, P], C.c_int),
    "cdae_dag_bwd": ([P, P, P, P, P, P, P, P, I32, I32, I32, I32, P], C.c_int),
    "cdae_sgemm": ([C.POINTER(SgemmDesc), P], C.c_int),
    "cdae_timestep_embedding": ([P, I32, I32, P, F32, P, P, I32, I32, P], C.c_int),
    "cdae_step_tick": ([P, P, I32, P], C.c_int),
    "cdae_randn": ([P, I64, P, I32, F32, P], C.c_int),
    "cdae_silu_bwd": ([P, P, I64, P], C.c_int),
    "cdae_silu_cast": ([P, P, I64, I32, P], C.c_int),
    "cdae_softplus_bwd": ([P, P, I64, P], C.c_int),
    "cdae_embed_rows": ([P, P, P, I32, I32, I32, P, P], C.c_int),
    "cdae_bn_finalize": ([P, F64, P, P, P, P, P, I32, P, P, I32, P], C.c_int),
    "cdae_enc_head": ([P, P, P, I32, I32, I32, I32, P], C.c_int),
    "cdae_bn_lrelu_bwd": ([P, P, P, P, P, P, P, P, I64, I32, P], C.c_int),
    "cdae_latent_fwd": ([P, P, P, P, P, P, P, P, P, I32, I32, I32, I32, F32, P], C.c_int),
    "cdae_latent_bwd": ([P, P, P, P, P, P, P, P, P, P, P, P, P, P, I32, I32, I32, I32, F32, P], C.c_int),
    "cdae_step_loss": ([P, P, P, P, P, P, I32, I32, P, P, P, P, P, P], C.c_int),
}
# entry points that later files add; absent symbols are only an error when called
_OPTIONAL = set()


class CdaeError(RuntimeError):
    pass


# kernels launched per C-ABI call (default 1; memsets are not kernels): lets callers COUNT the launches they issue
_KERNELS_PER_CALL = {"cdae_zero": 0, "cdae_version": 0, "cdae_init": 0, "cdae_last_error": 0, "cdae_adam_ema": 2, "cdae_randn": 2,
                     "cdae_bn_lrelu_bwd": 2, "cdae_dag_bwd": 2, "cdae_attn_bwd": 2}
_kernel_count = [0]


def kernel_count():
    """kernels launched through the C ABI by this process so far (graph replays do not pass through here: count a capture)"""
    return _kernel_count[0]


class _Counting:
    """thin proxy over the ctypes library: every entry point bumps the launch counter"""

    def __init__(self, lib):
        self._lib = lib
        self._cache = {}

    def __getattr__(self, name):
        fn = self._cache.get(name)
        if fn is None:
            raw = getattr(self._lib, name)
            n = _KERNELS_PER_CALL.get(name, 1)
            if n == 0:
                fn = raw
            else:
                def fn(*a, _raw=raw, _n=n):
                    _kernel_count[0] += _n
                    return _raw(*a)
            self._cache[name] = fn
        return fn


def load():
    """dlopen libcdae.so and type its entry points (no CUDA calls: safe on a CPU-only box)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(_LIB_PATH):
            raise CdaeError(f"{_LIB_PATH} is missing: run `python -m causaldiffae_b200.build` (there is no fallback path)")
        lib = C.CDLL(_LIB_PATH)
        for name, (args, res) in _SIGS.items():
            try:
                fn = getattr(lib, name)
            except AttributeError:
                if name in _OPTIONAL:
                    continue
                raise CdaeError(f"libcdae.so does not export {name}")
            fn.argtypes, fn.restype = args, res
        lib.cdae_last_error.restype = C.c_char_p
        _lib = _Counting(lib)
        return _lib


def lib():
    """Library handle checked against the current CUDA device (sm_100 only)."""
    l = load()
    if not torch.cuda.is_available():
        raise CdaeError("causaldiffae_b200 needs a CUDA device (sm_100a); there is no CPU path")
    dev = torch.cuda.current_device()
    if dev not in _inited_devices:
        rc = l.cdae_init()
        if rc != 0:
            raise CdaeError(f"cdae_init failed ({rc}): {l.cdae_last_error().decode()}")
        _inited_devices.add(dev)
    return l


def check(rc):
    if rc != 0:
        raise CdaeError(f"libcdae call failed ({rc}): {_lib.cdae_last_error().decode()}")


def stream():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()
