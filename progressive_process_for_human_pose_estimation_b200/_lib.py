"""ctypes binding of libhg_sm100a.so (C ABI declared in include/hg_sm100a.h).

This is the thin host layer: every call takes raw device pointers (``tensor.data_ptr()``) and the current
CUDA stream and only enqueues kernels.  There is no CPU or PyTorch fallback: if the library is missing or a
call fails, a RuntimeError is raised.
"""
import ctypes as C
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhg_sm100a.so")
CSRC = os.path.join(_HERE, "csrc")

HG_BF16, HG_F32, HG_F16 = 0, 1, 2
HG_CE_MAX_TERMS = 16


class HgConvDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("N", "H", "W", "Cin", "Cout", "R", "S", "stride", "pad", "dil", "dtype")]


class HgBnDesc(C.Structure):
    _fields_ = [("M", C.c_int64), ("C", C.c_int32), ("dtype", C.c_int32), ("eps", C.c_float),
                ("relu", C.c_int32), ("use_running", C.c_int32)]


class HgBnFold(C.Structure):
    _fields_ = [("stats", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("running_mean", C.c_void_p),
                ("running_var", C.c_void_p), ("eps", C.c_float), ("relu", C.c_int32), ("use_running", C.c_int32),
                ("pad_", C.c_int32)]


class HgBnStatsSlot(C.Structure):
    _fields_ = [("stats", C.c_void_p), ("pivot_src", C.c_void_p), ("C", C.c_int32), ("Cp", C.c_int32)]


class HgBnRunningSite(C.Structure):
    _fields_ = [("stats", C.c_void_p), ("count", C.c_float), ("pad_", C.c_int32)]


class HgBnRunningModule(C.Structure):
    _fields_ = [("running_mean", C.c_void_p), ("running_var", C.c_void_p), ("num_batches_tracked", C.c_void_p),
                ("C", C.c_int32), ("Cp", C.c_int32), ("first_site", C.c_int32), ("num_sites", C.c_int32),
                ("momentum", C.c_float), ("pad_", C.c_int32)]


class HgGaussDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "P", "J", "H", "W", "center_mode", "truncate", "accumulate")] + \
               [("pre_scale", C.c_double), ("sigma", C.c_double), ("amplitude", C.c_double)]


class HgMseDesc(C.Structure):
    _fields_ = [("numel", C.c_int64), ("num_stacks", C.c_int32), ("grad_scale", C.c_float)]


class HgCeTerm(C.Structure):
    _fields_ = [("logits", C.c_void_p), ("dlogits", C.c_void_p), ("target", C.c_void_p), ("logits_bstride", C.c_int64),
                ("dlogits_bstride", C.c_int64), ("channels", C.c_int32), ("norm", C.c_float),
                ("pixel_weight", C.c_void_p), ("nll_out", C.c_void_p)]


class HgCeDesc(C.Structure):
    _fields_ = [("num_terms", C.c_int32), ("B", C.c_int32), ("HW", C.c_int32), ("ignore_index", C.c_int32),
                ("grad_scale", C.c_float)]


class HgAdamChunk(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("n", C.c_int64)]


class HgAdamDesc(C.Structure):
    _fields_ = [("lr_d", C.c_double), ("beta1_d", C.c_double), ("beta2_d", C.c_double), ("lr", C.c_float),
                ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("weight_decay", C.c_float),
                ("step", C.c_int32), ("num_chunks", C.c_int32), ("pad_", C.c_int32)]


class HgResizeImage(C.Structure):
    _fields_ = [("src", C.c_void_p), ("w", C.c_int32), ("h", C.c_int32), ("kx_off", C.c_int32), ("ky_off", C.c_int32),
                ("bx_off", C.c_int32), ("by_off", C.c_int32), ("ksize_x", C.c_int32), ("ksize_y", C.c_int32),
                ("tmp_off", C.c_int64)]


class HgAnnotDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "P", "J", "mode")]


class HgLabelDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "P", "J", "L", "H", "W", "center_mode", "draw_points", "draw_lines",
                                         "line_value")]


_P = C.c_void_p
_I = C.c_int
_LL = C.c_longlong

# name -> argtypes (all return int status unless listed in _SPECIAL)
SIGNATURES = {
    "hg_device_ok": [],
    "hg_set_option": [C.c_char_p, _I],
    "hg_pack_conv_weight": [C.POINTER(HgConvDesc), _P, _P, _P, _P],
    "hg_conv_fprop": [C.POINTER(HgConvDesc), _P, _P, _P, _P, _P, _P, _P],
    "hg_conv_fprop_ex": [C.POINTER(HgConvDesc), _P, _P, _P, _P, _P, _P, _P, _P],
    "hg_conv_dgrad": [C.POINTER(HgConvDesc), _P, _P, _P, _P, _P],
    "hg_conv_wgrad": [C.POINTER(HgConvDesc), _P, _P, _P, _P, _P],
    "hg_conv_tc_eligible": [C.POINTER(HgConvDesc)],
    "hg_conv_fold_eligible": [C.POINTER(HgConvDesc)],
    "hg_conv_fprop_bn": [C.POINTER(HgConvDesc), C.POINTER(HgBnFold), _P, _P, _P, _P, _P, _P, _P, _P],
    "hg_conv_fprop_bnout": [C.POINTER(HgConvDesc), C.POINTER(HgBnFold), _P, _P, _P, _P, _P, _P],
    "hg_conv_wgrad_bn": [C.POINTER(HgConvDesc), C.POINTER(HgBnFold), _P, _P, _P, _P, _P],
    "hg_conv_dgrad_bn": [C.POINTER(HgConvDesc), C.POINTER(HgBnFold), _P, _P, _P, _P, _P, _P],
    "hg_unpack_conv_wgrad": [C.POINTER(HgConvDesc), _P, _P, _I, _P],
    "hg_bn_stats": [C.POINTER(HgBnDesc), _P, _P, _P],
    "hg_bn_apply": [C.POINTER(HgBnDesc), _P, _P, _P, _P, _P, _P, _P, _P],
    "hg_bn_bwd_reduce": [C.POINTER(HgBnDesc), _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "hg_bn_bwd_apply": [C.POINTER(HgBnDesc)] + [_P] * 14,
    "hg_bn_update_running": [_P, _P, _I, _P],
    "hg_bn_prepare_stats": [_P, _I, _P],
    "hg_maxpool2_fwd": [_I, _P, _I, _I, _I, _I, _P, _P, _P],
    "hg_maxpool2_bwd": [_I, _P, _P, _P, _I, _I, _I, _I, _P, _P],
    "hg_upsample2x_add_fwd": [_I, _I, _P, _P, _I, _I, _I, _I, _P, _P, _P],
    "hg_upsample2x_bwd": [_I, _I, _P, _P, _I, _I, _I, _I, _P, _P],
    "hg_spatial_mean": [_I, _P, _I, _I, _I, _I, C.c_float, _P, _P, _P],
    "hg_spatial_broadcast": [_I, _P, _I, _I, _I, _I, C.c_float, _P, _P, _P],
    "hg_channel_copy": [_I, _P, _I, _I, _P, _P, _I, _I, _I, _LL, _P],
    "hg_add": [_I, _P, _P, _P, _LL, _P],
    "hg_nchw_f32_to_nhwc": [_I, _P, _P, _I, _I, _I, _I, _P, _P],
    "hg_nhwc_to_nchw_f32": [_I, _P, _I, _I, _I, _I, _P, _P],
    "hg_stem_fwd": [_I, _P, _P, _P, _I, _I, _I, _I, _P, _P],
    "hg_stem_bwd": [_I, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P],
    "hg_pack_conv_weight_slice": [C.POINTER(HgConvDesc), _P, _I, _I, _P, _P, _P],
    "hg_unpack_conv_wgrad_slice": [C.POINTER(HgConvDesc), _P, _P, _I, _I, _I, _P],
    "hg_mix_rows": [_P, _P, _P, _I, _I, _I, _I, _P],
    "hg_mix_rows_rect": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "hg_gather_annotations": [C.POINTER(HgAnnotDesc), _P, _P, _P, _P, _P, _P, _P, _P],
    "hg_mse_multi": [C.POINTER(HgMseDesc), _P, _P, _P, _P, _P],
    "hg_scale_multi": [_LL, _I, _P, _P, _P],
    "hg_zero_async": [_P, _LL, _P],
    "hg_ce_multi": [C.POINTER(HgCeDesc), _P, _P, _P, _P, _P],
    "hg_image_u8_to_nchw_f32": [_P, _I, _I, _I, _I, _P, _P, _P, _P],
    "hg_resize_bicubic_u8": [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P],
    "hg_adam_multi": [C.POINTER(HgAdamDesc), _P, _P],
    "hg_mse_weighted": [_P, _P, _P, _I, _I, _I, _I, C.c_float, C.c_float, _P, _P, _P, _P],
    "hg_topk_mask": [_P, _I, _I, _I, _P, _P, _P],
    "hg_render_gauss": [C.POINTER(HgGaussDesc), _P, _P, _P, _P, _P],
    "hg_render_labels": [C.POINTER(HgLabelDesc), _P, _P, _P, _P, _P, _P],
    "hg_decode_argmax": [_P, _I, _I, _I, _I, _P, _P, _P],
    "hg_pckh_sweep": [_P, _I, _I, _I, _I, _I, _P, _P, _I, _I, _P, _I, _P, _P, _P, _P, _P, _P, _P],
    "hg_pckh_abs": [_P, _I, _I, _I, _I, _I, _P, _P, _I, _I, _P, _I, _P, _P, _P, _P, _P, _P, _P],
    "hg_softmax_stats": [_P, _I, _I, _I, _I, _P, _P],
    "hg_pckh_logits": [_P, _P, _I, _I, _I, _I, _I, _P, _P, _I, _I, _P, _I, _P, _P, _P, _P, _P, _P, _P],
    "hg_pckh_a": [_P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P],
}
_SPECIAL = {"hg_last_error_string": ([], C.c_char_p), "hg_launch_count": ([], C.c_ulonglong)}
EXPORTED = sorted(list(SIGNATURES) + list(_SPECIAL))

_lib = None


def build(verbose=False):
    """Compile csrc/*.cu for sm_100a into libhg_sm100a.so (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if verbose:
        print(out.stdout[-2000:], out.stderr[-2000:])
    if out.returncode != 0:
        raise RuntimeError("building libhg_sm100a.so failed:\n" + out.stdout[-4000:] + out.stderr[-4000:])
    return LIB_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (or make -C {CSRC}). "
            "There is no CPU / PyTorch fallback for the hourglass hot path.")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    for name, (argtypes, restype) in _SPECIAL.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def last_error():
    return load().hg_last_error_string().decode("utf-8", "replace")


def launch_count():
    return int(load().hg_launch_count())


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError(f"libhg_sm100a: {what} failed with status {rc}: {last_error()}")


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def zero_(t):
    """Clear a contiguous CUDA tensor with cudaMemsetAsync on the current stream (no fill kernel)."""
    call("hg_zero_async", ptr(t), t.numel() * t.element_size(), stream_ptr())
    return t


def ptr(t):
    """Device pointer of a tensor (or NULL for None)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def hg_dtype(torch_dtype):
    if torch_dtype == torch.bfloat16:
        return HG_BF16
    if torch_dtype == torch.float32:
        return HG_F32
    if torch_dtype == torch.float16:
        return HG_F16
    raise RuntimeError(f"unsupported dtype {torch_dtype}")


def call(name, *args):
    """Invoke an entry point and raise on a non-zero status."""
    rc = getattr(load(), name)(*args)
    check(rc, name)


def pad64(c):
    return (c + 63) // 64 * 64
