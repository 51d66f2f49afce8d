"""ctypes binding of libmog.so (the C ABI declared in include/mog.h).

The library is mandatory: there is no PyTorch/CPU fallback for any op on the hot path.  A
missing or unloadable ``libmog.so`` raises ``RuntimeError`` at first use (build it with
``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C multiple-objects-gan_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmog.so")

ACT_NONE, ACT_RELU, ACT_LRELU, ACT_GLU, ACT_TANH, ACT_SIGMOID = range(6)
PREC_FP32, PREC_BF16X3, PREC_BF16 = range(3)
PREC_NAMES = {"fp32": PREC_FP32, "bf16x3": PREC_BF16X3, "bf16": PREC_BF16}


class MogConvDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("N", "H", "W", "Cin", "Cout", "KH", "KW", "stride", "pad", "up2x", "act", "precision", "pad_w1")]


class MogPackEntry(C.Structure):
    _fields_ = [("w", C.c_void_p), ("hi", C.c_void_p), ("lo", C.c_void_p)] + \
               [(n, C.c_int32) for n in ("Cout", "Cin", "KHW", "transpose", "ntaps", "Nreal", "Npad", "Cs", "CsReal", "K", "Kpad",
                                         "nxb", "nyb", "block_start")] + [("taps", (C.c_int32 * 4) * 16)]


class MogPackGroup(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("first", "count", "block_start", "nxb")]


_p = C.c_void_p
_i = C.c_int
_f = C.c_float
_sz = C.c_size_t
_dp = C.POINTER(MogConvDesc)

# name -> (restype, argtypes); must list every symbol of include/mog.h (checked by tests)
SIGNATURES = {
    "mog_version": (_i, []),
    "mog_last_error": (C.c_char_p, []),
    "mog_launch_count": (C.c_ulonglong, []),
    "mog_nchw_to_nhwc": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "mog_nhwc_to_nchw": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "mog_packed_weight_bytes": (_sz, [_dp, _i]),
    "mog_packed_weight_layout": (_i, [_dp, _i]),
    "mog_pack_weight": (_i, [_dp, _i, _p, _p, _p]),
    "mog_pack_plan": (_i, [_dp, _i, _p, _p, C.POINTER(MogPackEntry), _i]),
    "mog_pack_multi": (_i, [_p, _p, _i, _i, _p]),
    "mog_conv_out_hw": (_i, [_dp, C.POINTER(_i), C.POINTER(_i)]),
    "mog_conv_workspace_bytes": (_sz, [_dp, _i]),
    "mog_planes_bytes": (_sz, [C.c_longlong, _i, _i]),
    "mog_split_planes": (_i, [_p, C.c_longlong, _i, _i, _p, _p]),
    "mog_split_planes_act": (_i, [_p, _p, _i, C.c_longlong, _i, _i, _p, _p]),
    "mog_patch_planes": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p]),
    "mog_col2im_act": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _i, _p, _p]),
    "mog_conv2d_fwd": (_i, [_dp, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "mog_conv2d_dgrad": (_i, [_dp, _p, _p, _p, _p, _p, _sz, _p]),
    "mog_conv2d_wgrad": (_i, [_dp, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "mog_bn_parts": (_i, [_i, _i, _i, _i, _i]),
    "mog_bn_stats": (_i, [_p, _i, _i, _i, _p, _i, _p]),
    "mog_bn_finalize": (_i, [_p, _i, _i, _i, _i, _p, _p, _f, _f, _p, _p, _p, _p, _p, _p, _p]),
    "mog_affine_act_fwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "mog_affine_act_fwd_planes": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "mog_bn_act_bwd_reduce": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _i, _p, _p, _p, _p, _p]),
    "mog_bn_act_bwd_apply": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p]),
    "mog_bn_act_bwd_apply_planes": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _i, _p]),
    "mog_act_bwd": (_i, [_p, _p, _p, _sz, _i, _p]),
    "mog_sumpool2x2": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "mog_stn_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "mog_stn_bwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "mog_word_attention_fwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "mog_word_attention_bwd_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "mog_word_attention_bwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _sz, _p]),
    "mog_damsm_words_fwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _f, _f, _p]),
    "mog_damsm_bwd_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "mog_damsm_words_bwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _f, _p, _sz, _p]),
    "mog_sigmoid_bce_fwd": (_i, [_p, _p, _f, _i, _p, _p, _i, _i, _p]),
    "mog_sigmoid_bce_bwd": (_i, [_p, _p, _f, _i, _p, _p, _i, _p]),
    "mog_pool2d_out_hw": (_i, [_i, _i, _i, _i, _i, C.POINTER(_i), C.POINTER(_i)]),
    "mog_pool2d_fwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "mog_pool2d_bwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "mog_resize_bilinear_fwd": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "mog_resize_bilinear_bwd": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "mog_adam_multi": (_i, [_i, _p, _p, _p, _p, _p, _p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_longlong, C.c_double, _f, _p]),
    "mog_adam_multi_dev": (_i, [_i, _p, _p, _p, _p, _p, _p, C.c_double, C.c_double, C.c_double, C.c_double, _p, C.c_double, _f, _p]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libmog.so not found at %s -- the CUDA library is mandatory (no CPU fallback). "
                "Build it: python -c 'import __graft_entry__ as g; g.build()'" % LIB_PATH)
        try:
            L = C.CDLL(LIB_PATH)
        except OSError as e:  # pragma: no cover
            raise RuntimeError("cannot load %s: %s" % (LIB_PATH, e))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def launch_count() -> int:
    """Kernels launched by libmog since load (counted inside the library)."""
    return int(lib().mog_launch_count())


_fn_cache = {}


def call(name, *args):
    """Call an int-returning entry point; raise RuntimeError(mog_last_error()) on failure."""
    fn = _fn_cache.get(name)
    if fn is None:
        fn = _fn_cache[name] = getattr(lib(), name)     # (bound once: a step makes ~2000 of these calls)
    rc = fn(*args)
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (name, rc, lib().mog_last_error().decode()))
    return rc
