"""torch.autograd.Function wrappers over the libmog C ABI.

All activation tensors handled here are contiguous fp32 CUDA tensors in **NHWC** order
(``[N, H, W, C]``; 2-D ``[N, C]`` for the fully-connected parts).  PyTorch supplies device
memory, streams and the autograd tape only -- every computation is a libmog kernel.
"""
from __future__ import annotations

import os

import torch

from . import _lib
from ._lib import (ACT_GLU, ACT_LRELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, MogConvDesc,
                   PREC_BF16, PREC_BF16X3, PREC_FP32, PREC_NAMES, call)

import contextlib
import ctypes as C

# Product default: bf16x3 (tcgen05, fp32-equivalent: meets the 1e-3 end-to-end bound).  MOG_PRECISION / cfg.MOG.PRECISION
# (applied by the trainers through ``set_precision``) / ``set_precision`` override it; 'fp32' is the CUDA-core parity mode.
_default_precision = PREC_NAMES[os.environ.get("MOG_PRECISION", "bf16x3")]


def set_precision(name: str):
    """Operand precision of the convolution kernels: 'fp32' | 'bf16x3' | 'bf16'."""
    global _default_precision
    if isinstance(name, int) and name in PREC_NAMES.values():      # a value returned by get_precision()
        _default_precision = name
        return
    if name not in PREC_NAMES:
        raise ValueError("unknown precision %r (expected one of %s)" % (name, sorted(PREC_NAMES)))
    _default_precision = PREC_NAMES[name]


def precision_from_cfg(cfg):
    """``cfg.MOG.PRECISION`` -> the conv operand precision (called by the trainers' constructors); the environment
    variable MOG_PRECISION, when set, wins (A/B runs without editing YAML files)."""
    name = os.environ.get("MOG_PRECISION") or str(getattr(getattr(cfg, "MOG", None), "PRECISION", "") or "bf16x3")
    set_precision(name)
    return name


def get_precision() -> int:
    return _default_precision


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def _chk(t, name):
    if t is None:
        return
    if not t.is_cuda:
        raise RuntimeError("%s: libmog ops need CUDA tensors (no CPU fallback); got %s" % (name, t.device))
    if t.dtype != torch.float32:
        raise RuntimeError("%s: expected float32, got %s" % (name, t.dtype))
    if not t.is_contiguous():
        raise RuntimeError("%s: expected a contiguous tensor" % name)


# ---------------------------------------------------------------------------------------------
# layout helpers (API boundary: the reference's tensors are NCHW)
# ---------------------------------------------------------------------------------------------
def to_nhwc(x: torch.Tensor) -> torch.Tensor:
    """Logical NCHW tensor -> contiguous [N,H,W,C].  Zero-copy when x already is channels_last."""
    v = x.permute(0, 2, 3, 1)
    if v.is_contiguous():
        return v
    x = x.contiguous()
    _chk(x, "to_nhwc")
    N, Cc, H, W = x.shape
    out = torch.empty((N, H, W, Cc), device=x.device, dtype=x.dtype)
    call("mog_nchw_to_nhwc", x.data_ptr(), out.data_ptr(), N, Cc, H, W, _stream())
    return out


def to_nchw_view(x: torch.Tensor) -> torch.Tensor:
    """Contiguous [N,H,W,C] -> logical NCHW view (channels_last strides, no copy)."""
    return x.permute(0, 3, 1, 2)


class _ToNHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return to_nhwc(x)

    @staticmethod
    def backward(ctx, g):
        return to_nchw_view(g.contiguous())


def nhwc(x):
    """Differentiable NCHW -> NHWC (a free autograd-tracked view when x is channels_last)."""
    v = x.permute(0, 2, 3, 1)
    if v.is_contiguous():
        return v
    return _ToNHWC.apply(x) if x.requires_grad else to_nhwc(x)


# ---------------------------------------------------------------------------------------------
# convolution / linear
# ---------------------------------------------------------------------------------------------
def _weight_version(weight):
    # torch's version counter catches torch-side in-place updates; `_mog_ver` is bumped by mog_b200.optim.Adam, whose fused
    # kernel writes the parameter through its raw pointer (invisible to the version counter)
    return (weight._version, getattr(weight, "_mog_ver", 0), weight.data_ptr())


def _weight4(weight):
    w = weight.detach()
    if w.dim() == 2:
        w = w.reshape(w.shape[0], w.shape[1], 1, 1)
    return w.contiguous()


def _packed(weight: torch.Tensor, which: str, d: MogConvDesc, dkey=None) -> torch.Tensor:
    """Pack an OIHW parameter into the GEMM B operand of the kernel selected by ``d`` (fp32 matrix or
    bf16 hi/lo planes per stride phase).  One persistent buffer per (weight, conv geometry), refreshed when the weight's
    version moves -- lazily here, or for all weights of a network at once by :func:`repack` after an optimiser step."""
    cache = getattr(weight, "_mog_pack", None)
    if cache is None:
        cache = {}
        try:
            weight._mog_pack = cache
        except Exception:
            pass
    ver = _weight_version(weight)
    # the dgrad packing depends on stride/pad (phase tap subsets) and, with odd sizes, on H/W parity
    wi = 0 if which == "fwd" else 1
    tag = _layout_cache.get((dkey, wi)) if dkey is not None else None
    if tag is None:
        tag = _lib.lib().mog_packed_weight_layout(C.byref(d), wi)
        if dkey is not None:
            _layout_cache[(dkey, wi)] = tag
    key = (which, tag, d.precision, d.stride, d.pad, d.pad_w1, d.up2x, (d.H << d.up2x) >= d.stride, (d.W << d.up2x) >= d.stride)
    ent = cache.get(key)
    if ent is None:
        w = _weight4(weight)
        _chk(w, "weight")
        n = _lib.lib().mog_packed_weight_bytes(C.byref(d), wi)
        ent = cache[key] = {"out": torch.empty((n + 3) // 4, device=w.device, dtype=torch.float32), "ver": None,
                            "d": MogConvDesc.from_buffer_copy(d), "wi": wi}
    if ent["ver"] != ver:
        w = _weight4(weight)
        _chk(w, "weight")
        call("mog_pack_weight", C.byref(d), wi, w.data_ptr(), ent["out"].data_ptr(), _stream())
        ent["ver"] = ver
    return ent["out"]


def invalidate_packed(module_or_params):
    """Mark the packed GEMM operands of the given parameters stale.  Needed after any write that bypasses both torch's
    version counter and ``mog_b200.optim.Adam``: ``p.data.copy_(...)`` / ``p.data.normal_()`` / ``dist.broadcast(p.data)``
    do not bump ``p._version`` (``load_params``, ``weights_init`` and ``parallel.broadcast_params`` call this)."""
    params = module_or_params.parameters() if hasattr(module_or_params, "parameters") else module_or_params
    for p in params:
        p._mog_ver = getattr(p, "_mog_ver", 0) + 1


_repack_tables = {}


def repack(params):
    """Refresh every packed operand of the given parameters with ONE multi-tensor launch (``mog_pack_multi``): called by
    ``mog_b200.optim.Adam.step`` right after the update, instead of ~350 per-problem pack launches spread over the next
    step.  Entries the multi-tensor form does not cover (filters with more than 16 taps, fp32 mode) stay on the lazy path."""
    params = [p for p in params if getattr(p, "_mog_pack", None)]
    if not params:
        return
    sig = tuple((id(p), p.data_ptr(), len(p._mog_pack)) for p in params)
    tkey = tuple(id(p) for p in params)
    tab = _repack_tables.get(tkey)
    if tab is None or tab["sig"] != sig:
        L = _lib.lib()
        buf = (_lib.MogPackEntry * 16)()
        entries, groups, ents = [], [], []
        for p in params:
            w = _weight4(p)
            for ent in p._mog_pack.values():
                if ent["d"].precision == PREC_FP32:
                    continue
                n = L.mog_pack_plan(C.byref(ent["d"]), ent["wi"], w.data_ptr(), ent["out"].data_ptr(), buf, 16)
                if n <= 0:
                    continue            # not covered: refreshed lazily by _packed
                for i in range(n):
                    e = _lib.MogPackEntry.from_buffer_copy(buf[i])
                    g = groups[-1] if groups else None
                    # problems of the same weight that read the same source tiles (sub-pixel phases, parity views, stride
                    # phases) share one group: the tile is loaded once and written into each of them
                    if i > 0 and g is not None and (e.Cs, e.Npad, e.transpose, e.nxb, e.nyb) == g["sig"]:
                        g["count"] += 1
                    else:
                        groups.append({"first": len(entries), "count": 1, "sig": (e.Cs, e.Npad, e.transpose, e.nxb, e.nyb),
                                       "blocks": e.nxb * e.nyb, "nxb": e.nxb})
                    entries.append(e)
                ents.append((p, ent))
        blocks = 0
        garr = (_lib.MogPackGroup * max(len(groups), 1))()
        for i, g in enumerate(groups):
            garr[i].first, garr[i].count, garr[i].block_start, garr[i].nxb = g["first"], g["count"], blocks, g["nxb"]
            blocks += g["blocks"]
        dev = gdev = None
        if entries:
            arr = (_lib.MogPackEntry * len(entries))(*entries)
            dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(params[0].device)
            gdev = torch.frombuffer(bytearray(bytes(garr)), dtype=torch.uint8).to(params[0].device)
        tab = _repack_tables[tkey] = {"sig": sig, "dev": dev, "gdev": gdev, "n": len(groups), "blocks": blocks, "ents": ents}
    if tab["n"]:
        call("mog_pack_multi", tab["dev"].data_ptr(), tab["gdev"].data_ptr(), tab["n"], tab["blocks"], _stream())
        for p, ent in tab["ents"]:
            ent["ver"] = _weight_version(p)


_desc_cache = {}
_ws_cache = {}
_layout_cache = {}
_planes_bytes_cache = {}


def _planes_bytes(rows, Cc, precision):
    k = (rows, Cc, precision)
    v = _planes_bytes_cache.get(k)
    if v is None:
        v = _planes_bytes_cache[k] = _lib.lib().mog_planes_bytes(rows, Cc, precision)
    return v


def _desc(x_shape, w_shape, stride, pad, up2x, act, precision):
    """Conv descriptor + output size for a call site; cached per geometry (the host side of a step issues ~2000 launches,
    so every avoidable ctypes round trip matters: the GPU must not wait for Python)."""
    key = (tuple(x_shape), tuple(w_shape), stride, pad if not isinstance(pad, list) else tuple(pad), bool(up2x), act, precision)
    hit = _desc_cache.get(key)
    if hit is not None:
        return hit
    hit = _desc_cache[key] = _desc_uncached(x_shape, w_shape, stride, pad, up2x, act, precision) + (key,)
    return hit


def _desc_uncached(x_shape, w_shape, stride, pad, up2x, act, precision):
    N, H, W, Ci = x_shape
    Co, Ci2, KH, KW = w_shape
    if Ci != Ci2:
        raise RuntimeError("conv: input has %d channels, weight expects %d" % (Ci, Ci2))
    if isinstance(pad, (tuple, list)):   # (pad_h, pad_w): the 1x7 / 7x1 / 1x3 / 3x1 filters of the image encoder
        d = MogConvDesc(N, H, W, Ci, Co, KH, KW, stride, int(pad[0]), int(up2x), act, precision, int(pad[1]) + 1)
    else:
        d = MogConvDesc(N, H, W, Ci, Co, KH, KW, stride, pad, int(up2x), act, precision, 0)
    ho, wo = C.c_int(), C.c_int()
    call("mog_conv_out_hw", C.byref(d), C.byref(ho), C.byref(wo))
    return d, ho.value, wo.value


def _workspace(d, which, device, key=None):
    n = _ws_cache.get((key, which)) if key is not None else None
    if n is None:
        n = _lib.lib().mog_conv_workspace_bytes(C.byref(d), which)
        if key is not None:
            _ws_cache[(key, which)] = n
    if n == 0:
        return None, 0
    ws = torch.empty((n + 3) // 4, device=device, dtype=torch.float32)
    return ws, n


def _alloc_planes(shape, precision, device):
    Cc = shape[-1]
    rows = 1
    for v in shape[:-1]:
        rows *= int(v)
    n = _planes_bytes(rows, Cc, precision)
    return torch.empty((n + 3) // 4, device=device, dtype=torch.float32)


_NO_PRODUCER_PLANES = os.environ.get("MOG_NO_PRODUCER_PLANES", "0") == "1"


def _attach_planes(y: torch.Tensor, planes: torch.Tensor, precision: int):
    """Remember the pre-split planes of ``y`` (written by the producing kernel) for a consuming convolution."""
    if _NO_PRODUCER_PLANES:      # debug knob (MOG_NO_PRODUCER_PLANES=1): every consumer runs its own split pass
        return
    y._mog_planes = (planes, precision, y._version, y.data_ptr())


def planes_of(x: torch.Tensor, precision: int):
    """Planes of x: those its producer emitted (bn_act / activation epilogue) if still valid, else a split pass."""
    cached = getattr(x, "_mog_planes", None)
    if cached is not None and cached[1] == precision and cached[2] == x._version and cached[3] == x.data_ptr():
        return cached[0]
    planes = split_planes(x, precision)
    # remember them on x: a tensor with several consumers (the input of an Inception block feeds 3-4 convolutions; x of a
    # conv is read again by its weight gradient) is split once, not once per consumer
    try:
        _attach_planes(x, planes, precision)
    except Exception:
        pass
    return planes


def split_planes(x: torch.Tensor, precision: int):
    """fp32 [..., C] -> bf16 planes buffer (hi [rows][C8], then lo for bf16x3) for the tcgen05 kernels."""
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    n = _planes_bytes(rows, Cc, precision)
    planes = torch.empty((n + 3) // 4, device=x.device, dtype=torch.float32)
    call("mog_split_planes", x.data_ptr(), rows, Cc, precision, planes.data_ptr(), _stream())
    return planes


# ---- weight gradients on a side stream -----------------------------------------------------------------------------------
# Inside ``with async_wgrad():`` (the trainers wrap their step in it) the weight gradient of a convolution is enqueued on a
# side stream of the stream its backward runs on, and is handed to ``weight.grad`` when the whole backward pass has finished
# (autograd's end-of-pass callback) instead of being returned through autograd: the data-gradient chain -- BatchNorm backward
# (HBM bound) -> dgrad (tensor bound) -> ... -- does not wait for the weight gradients, which fill the tensor pipe while the
# bandwidth-bound kernels of the next layer run.  Same kernels, same accumulation order for weights used more than once:
# bit-identical to the synchronous order.  Only ``loss.backward()`` style passes (gradients accumulated into leaves) may use
# it -- ``torch.autograd.grad(..., weight)`` would see no gradient -- hence opt-in.
_async = {"on": False, "pending": [], "sides": {}, "used": [], "cb": False}


@contextlib.contextmanager
def async_wgrad(enabled=True):
    prev = _async["on"]
    _async["on"] = bool(enabled)
    try:
        yield
    finally:
        _async["on"] = prev
        if _async["pending"]:      # a backward pass that did not complete: drop what it left
            _async["pending"].clear()
            _async["used"].clear()
            _async["cb"] = False


def _async_side(cur):
    s = _async["sides"].get(cur.cuda_stream)
    if s is None:
        s = _async["sides"][cur.cuda_stream] = torch.cuda.Stream(device=cur.device)
    return s


def _async_defer(weight, dw, keep, side):
    _async["pending"].append((weight, dw, keep))
    if side not in _async["used"]:
        _async["used"].append(side)
    if not _async["cb"]:
        _async["cb"] = True
        torch.autograd.Variable._execution_engine.queue_callback(_async_flush)


def _async_flush():
    """End of the backward pass (runs with the caller's current stream set): join the side streams, hand the gradients over."""
    cur = torch.cuda.current_stream()
    for s in _async["used"]:
        cur.wait_stream(s)
    with torch.no_grad():
        for weight, dw, _keep in _async["pending"]:
            if weight.grad is None:
                weight.grad = dw
            else:
                weight.grad.add_(dw)
    _async["pending"].clear()
    _async["used"].clear()
    _async["cb"] = False


def _split_dy(dy, y, act, precision, need_db, needed):
    """Output gradient of a conv whose epilogue applied ``act``: (fp32 dz or None, planes or None).  In the tcgen05 precisions
    the activation backward is fused into the plane split (dz never exists in fp32) unless the bias gradient needs it."""
    dyp = None
    if act != ACT_NONE and precision != PREC_FP32 and not need_db and needed:
        Cc = dy.shape[-1]
        rows = dy.numel() // Cc
        dyp = torch.empty((_planes_bytes(rows, Cc, precision) + 3) // 4, device=dy.device, dtype=torch.float32)
        call("mog_split_planes_act", dy.data_ptr(), y.data_ptr(), act, rows, Cc, precision, dyp.data_ptr(), _stream())
        dy = None
    elif act != ACT_NONE:
        dz = torch.empty_like(dy)
        call("mog_act_bwd", dy.data_ptr(), y.data_ptr(), dz.data_ptr(), dy.numel(), act, _stream())
        dy = dz
    if dyp is None and precision != PREC_FP32 and needed:
        dyp = planes_of(dy, precision)     # emitted by the BatchNorm backward that produced dy, else a split pass
    return dy, dyp


# ---- thin ends of the networks: 3-channel inputs / outputs on the tensor cores through the patch (im2col) matrix ----------
# A conv with C <= 4 input channels spends one K = 16 MMA step and one TMA box per filter tap on 3 real channels (D_NET256's
# first layer: forward 0.42 ms, weight gradient 0.87 ms at B = 32 for 0.01 TFLOP).  Its patch matrix [pixels][KH*KW*C -> KP]
# is small (KP = 48 for 4x4x3); on it the conv is a 1x1 conv with KP channels and the weight gradient a plain GEMM.  The same
# holds for the BACKWARD of a conv with <= 4 output channels (GET_IMAGE_G): its data gradient is a conv of the 3-channel dz,
# its weight gradient the GEMM x^T . patches(dz).
def _patch_planes(x, KH, KW, stride, pad, precision):
    N, H, W, Cc = x.shape
    Ho, Wo = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
    KP = (KH * KW * Cc + 7) // 8 * 8
    planes = torch.empty((_planes_bytes(N * Ho * Wo, KP, precision) + 3) // 4, device=x.device, dtype=torch.float32)
    call("mog_patch_planes", x.data_ptr(), N, H, W, Cc, KH, KW, stride, pad, precision, planes.data_ptr(), _stream())
    return planes, Ho, Wo, KP


def _pack_matrix(w2, d1):
    """Packed forward operand of a temporary 1x1 weight [Cout][KP] (not cached: rebuilt from the live parameter every call,
    so it is never stale -- inside a captured step the rebuild is part of the graph)."""
    n = _lib.lib().mog_packed_weight_bytes(C.byref(d1), 0)
    out = torch.empty((n + 3) // 4, device=w2.device, dtype=torch.float32)
    call("mog_pack_weight", C.byref(d1), 0, w2.data_ptr(), out.data_ptr(), _stream())
    return out


def _padded_matrix(m2, KP):
    """[R][K] -> contiguous [R][KP], zero pad columns."""
    if m2.shape[1] == KP:
        return m2.contiguous()
    out = torch.zeros(m2.shape[0], KP, device=m2.device, dtype=torch.float32)
    out[:, :m2.shape[1]] = m2
    return out


def _gemm_grid(N, H, W):
    """Pixel grid for a 1x1 problem on a patch matrix: neighbours do not matter, so all M pixels become one [M/8, 8] grid whose
    8 x 8 / 8 x 16 pixel tiles are CONTIGUOUS runs of the planes (one multi-KB DRAM burst per TMA box instead of 8 - 16 rows
    of a few hundred bytes: the 2M-pixel weight-gradient GEMM of GET_IMAGE_G ran at 1.9 TB/s on the image-shaped grid)."""
    M = N * H * W
    return (1, M // 8, 8) if M % 8 == 0 else (N, H, W)


def _padded_rows(m2, KP):
    """[K][C] -> contiguous [KP][C], zero pad rows."""
    if m2.shape[0] == KP:
        return m2.contiguous()
    out = torch.zeros(KP, m2.shape[1], device=m2.device, dtype=torch.float32)
    out[:m2.shape[0]] = m2
    return out


def _gemm_then_col2im(src, src_planes, w2, N, Hs, Ws, out_shape, Cc, KH, KW, stride, pad, bias, act, precision):
    """out = col2im_act(z), z = the 1x1 problem src [N*Hs*Ws][Cs] . w2^T [Cs][KP] (fp32, [N*Hs*Ws][KP]); see mog_col2im_act."""
    KP, Cs = w2.shape
    d1, _, _, dkey1 = _desc(_gemm_grid(N, Hs, Ws) + (Cs,), (KP, Cs, 1, 1), 1, 0, False, ACT_NONE, precision)
    dev = w2.device
    z = torch.empty((N * Hs * Ws, KP), device=dev, dtype=torch.float32)
    ws, nws = _workspace(d1, 0, dev, dkey1)
    call("mog_conv2d_fwd", C.byref(d1), _ptr(src), _ptr(src_planes), _pack_matrix(w2, d1).data_ptr(), None, z.data_ptr(), _ptr(ws), nws,
         _stream())
    out = torch.empty(out_shape, device=dev, dtype=torch.float32)
    _, H, W, _ = out_shape
    call("mog_col2im_act", z.data_ptr(), KP, N, H, W, Cc, KH, KW, stride, pad, _ptr(bias), act, out.data_ptr(), _stream())
    return out


def _thin_cin(x, weight, stride, pad, up2x, precision):
    if precision == PREC_FP32 or up2x or weight.dim() != 4 or x.dim() != 4 or isinstance(pad, (tuple, list)):
        return False
    Co, Ci, KH, KW = weight.shape
    return Ci <= 4 and KH * KW > 1 and KH * KW * Ci <= 64


def _thin_cout_bwd(w_shape, stride, pad, up2x, precision, has_bias):
    if precision == PREC_FP32 or up2x or stride != 1 or has_bias or isinstance(pad, (tuple, list)):
        return False
    Co, Ci, KH, KW = w_shape
    return Co <= 4 and Ci > 4 and KH == KW and KH == 2 * pad + 1 and KH > 1 and KH * KW * Co <= 64


class PatchConv2dFn(torch.autograd.Function):
    """Conv2dFn for <= 4 input channels: forward and weight gradient on the patch matrix (see above); the data gradient is the
    ordinary one.  replaces nn.Conv2d(3, ndf, 4, 2, 1) (model.py:598) and Inception's Conv2d_1a_3x3 (model.py:258)."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, pad, act, precision):
        _chk(x, "conv input")
        Co, Ci, KH, KW = weight.shape
        N, H, W, Cx = x.shape
        if Cx != Ci:
            raise RuntimeError("conv: input has %d channels, weight expects %d" % (Cx, Ci))
        xP, Ho, Wo, KP = _patch_planes(x, KH, KW, stride, pad, precision)
        d1, _, _, dkey1 = _desc(_gemm_grid(N, Ho, Wo) + (KP,), (Co, KP, 1, 1), 1, 0, False, act, precision)
        frozen = not weight.requires_grad
        ent = getattr(weight, "_mog_patch_pack", None) if frozen else None
        ver = _weight_version(weight)
        if ent is None or ent[0] != (ver, KP, precision):
            w2 = _padded_matrix(weight.detach().permute(0, 2, 3, 1).reshape(Co, KH * KW * Ci), KP)
            ent = ((ver, KP, precision), _pack_matrix(w2, d1))
            if frozen:
                weight._mog_patch_pack = ent
        y = torch.empty((N, Ho, Wo, Co), device=x.device, dtype=torch.float32)
        ws, nws = _workspace(d1, 0, x.device, dkey1)
        b = None if bias is None else bias.detach().contiguous()
        call("mog_conv2d_fwd", C.byref(d1), None, xP.data_ptr(), ent[1].data_ptr(), _ptr(b), y.data_ptr(), _ptr(ws), nws, _stream())
        ctx.cfg = (stride, pad, act, precision, tuple(x.shape), (Ho, Wo, KP))
        ctx.has_bias = bias is not None
        need_w = weight.requires_grad or (bias is not None and bias.requires_grad)
        ctx.save_for_backward(xP if need_w else None, weight, y if act != ACT_NONE else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        xP, weight, y = ctx.saved_tensors
        stride, pad, act, precision, xshape, (Ho, Wo, KP) = ctx.cfg
        dy = dy.contiguous()
        Co, Ci, KH, KW = weight.shape
        N = xshape[0]
        st, dev = _stream(), dy.device
        need_dx = ctx.needs_input_grad[0]
        need_dw = ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2])
        need_db = ctx.has_bias and ctx.needs_input_grad[2]
        dy, dyp = _split_dy(dy, y, act, precision, need_db, need_dx or need_dw)
        dx = dw = db = None
        if need_dx and KH * KW <= 16:
            # dx = col2im(dy . W2), W2[(kh, kw, ci)][co] = w[co][ci][kh][kw]: one GEMM over the output pixels + a gather, instead
            # of stride^2 phase problems whose N tile holds 3 real columns
            w2 = _padded_rows(weight.detach().permute(2, 3, 1, 0).reshape(KH * KW * Ci, Co), KP)
            dx = _gemm_then_col2im(dy, dyp, w2, N, Ho, Wo, xshape, Ci, KH, KW, stride, pad, None, ACT_NONE, precision)
        elif need_dx:
            d, _, _, dkey = _desc(xshape, weight.shape, stride, pad, False, ACT_NONE, precision)
            dx = torch.empty(xshape, device=dev, dtype=torch.float32)
            ws, nws = _workspace(d, 1, dev, dkey)
            call("mog_conv2d_dgrad", C.byref(d), _ptr(dy), _ptr(dyp), _packed(weight, "dgrad", d, dkey).data_ptr(),
                 dx.data_ptr(), _ptr(ws), nws, st)
        if need_dw:
            d1, _, _, dkey1 = _desc(_gemm_grid(N, Ho, Wo) + (KP,), (Co, KP, 1, 1), 1, 0, False, ACT_NONE, precision)
            dw1 = torch.empty((Co, KP), device=dev, dtype=torch.float32)
            if ctx.has_bias:
                db = torch.empty(Co, device=dev, dtype=torch.float32)
            ws, nws = _workspace(d1, 2, dev, dkey1)
            call("mog_conv2d_wgrad", C.byref(d1), None, xP.data_ptr(), _ptr(dy), _ptr(dyp), dw1.data_ptr(), _ptr(db), _ptr(ws), nws, st)
            dw = dw1[:, :KH * KW * Ci].reshape(Co, KH, KW, Ci).permute(0, 3, 1, 2).contiguous()
        return dx, dw, db, None, None, None, None


class Conv2dFn(torch.autograd.Function):
    """y = act(conv(up2x?(x), w) + b), NHWC.  replaces nn.Conv2d (+ nn.Upsample) of model.py.
    In the tcgen05 precisions the input is split once into bf16 planes (kept for the weight
    gradient instead of the fp32 tensor); the output gradient is split once in backward and
    shared by the data and weight gradients."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, pad, up2x, act, precision):
        _chk(x, "conv input")
        w4 = weight if weight.dim() == 4 else weight.reshape(weight.shape[0], weight.shape[1], 1, 1)
        d, Ho, Wo, dkey = _desc(x.shape, w4.shape, stride, pad, up2x, act, precision)
        b = None if bias is None else bias.detach().contiguous()
        xp = planes_of(x, precision) if precision != PREC_FP32 else None
        if _thin_cout_bwd(w4.shape, stride, pad, up2x, precision, False) and w4.shape[2] * w4.shape[3] <= 16:
            # <= 4 output channels ('same' conv: GET_IMAGE_G): z[q][(a, b, co)] = sum_ci x[q][ci] w[co][ci][KH-1-a][KW-1-b] is one
            # GEMM over the pixels with KP = KH*KW*Cout -> 8 columns; y = col2im(z) + bias, activation (mog_col2im_act)
            Co, Ci, KH, KW = w4.shape
            KP = (KH * KW * Co + 7) // 8 * 8
            w2 = _padded_rows(w4.detach().flip(2, 3).permute(2, 3, 0, 1).reshape(KH * KW * Co, Ci), KP)
            y = _gemm_then_col2im(x, xp, w2, d.N, d.H, d.W, (d.N, Ho, Wo, Co), Co, KH, KW, 1, pad, b, act, precision)
        else:
            y = torch.empty((d.N, Ho, Wo, d.Cout), device=x.device, dtype=torch.float32)
            ws, nws = _workspace(d, 0, x.device, dkey)
            call("mog_conv2d_fwd", C.byref(d), x.data_ptr(), _ptr(xp), _packed(weight, "fwd", d, dkey).data_ptr(), _ptr(b),
                 y.data_ptr(), _ptr(ws), nws, _stream())
        ctx.cfg = (stride, pad, up2x, act, precision, tuple(x.shape))
        ctx.has_bias = bias is not None
        # the input is only needed again for the weight / bias gradient: a frozen conv (image encoder) keeps nothing of it
        need_w = weight.requires_grad or (bias is not None and bias.requires_grad)
        ctx.save_for_backward(x if (xp is None and need_w) else None, xp if need_w else None, weight,
                              y if act != ACT_NONE else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, xp, weight, y = ctx.saved_tensors
        stride, pad, up2x, act, precision, xshape = ctx.cfg
        dy = dy.contiguous()
        w4 = weight if weight.dim() == 4 else weight.reshape(weight.shape[0], weight.shape[1], 1, 1)
        d, Ho, Wo, dkey = _desc(xshape, w4.shape, stride, pad, up2x, ACT_NONE, precision)
        st = _stream()
        dev = dy.device
        need_dx = ctx.needs_input_grad[0]
        need_dw = ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2])
        need_db = ctx.has_bias and ctx.needs_input_grad[2]
        if _thin_cout_bwd(w4.shape, stride, pad, up2x, precision, ctx.has_bias) and (need_dx or need_dw):
            return Conv2dFn._backward_thin_cout(dy, y, xp, weight, act, precision, xshape, need_dx, need_dw)
        dy, dyp = _split_dy(dy, y, act, precision, need_db, need_dx or need_dw)
        dx = dw = db = None
        if need_dx:
            dx = torch.empty(xshape, device=dev, dtype=torch.float32)
            ws, nws = _workspace(d, 1, dev, dkey)
            call("mog_conv2d_dgrad", C.byref(d), _ptr(dy), _ptr(dyp), _packed(weight, "dgrad", d, dkey).data_ptr(),
                 dx.data_ptr(), _ptr(ws), nws, st)
        if need_dw:
            side = None
            if _async["on"] and not ctx.has_bias and weight.is_leaf and dev.type == "cuda":
                cur = torch.cuda.current_stream()
                side = _async_side(cur)
                side.wait_stream(cur)          # the operands (planes of x and dy) are ready on cur
            with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
                dw = torch.empty(w4.shape, device=dev, dtype=torch.float32)
                if ctx.has_bias:
                    db = torch.empty(d.Cout, device=dev, dtype=torch.float32)
                ws, nws = _workspace(d, 2, dev, dkey)
                call("mog_conv2d_wgrad", C.byref(d), _ptr(x), _ptr(xp), _ptr(dy), _ptr(dyp), dw.data_ptr(), _ptr(db),
                     _ptr(ws), nws, _stream())
                dw = dw.reshape(weight.shape)
            if side is not None:
                # (the operands stay referenced until the end of the pass: their memory must not be reused on cur meanwhile)
                _async_defer(weight, dw, (x, xp, dy, dyp, ws), side)
                dw = None
        return dx, dw, db, None, None, None, None, None

    @staticmethod
    def _backward_thin_cout(dy, y, xp, weight, act, precision, xshape, need_dx, need_dw):
        """Backward of a 'same' conv with <= 4 output channels (GET_IMAGE_G, model.py:464-474) through the patch matrix of
        dz = dy * act'(y), dzP[p][(a, b, co)] = dz[p + (a - pad, b - pad)][co]:
          dx[p][ci]          = sum_k dzP[p][k] * w[co][ci][KH-1-a][KW-1-b]        (1x1 conv, KP -> Cin)
          dW[co][ci][kh][kw] = sum_p x[p][ci] * dzP[p][(KH-1-kh, KW-1-kw, co)]    (1x1 weight gradient: dzP input, x 'gradient')"""
        Co, Ci, KH, KW = weight.shape
        N, H, W, _ = xshape
        st, dev = _stream(), dy.device
        if act != ACT_NONE:
            dz = torch.empty_like(dy)
            call("mog_act_bwd", dy.data_ptr(), y.data_ptr(), dz.data_ptr(), dy.numel(), act, st)
        else:
            dz = dy
        dzP, _, _, KP = _patch_planes(dz, KH, KW, 1, (KH - 1) // 2, precision)
        K = KH * KW * Co
        d1, _, _, dkey1 = _desc(_gemm_grid(N, H, W) + (KP,), (Ci, KP, 1, 1), 1, 0, False, ACT_NONE, precision)
        dx = dw = None
        if need_dx:
            w2 = _padded_matrix(weight.detach().flip(2, 3).permute(1, 2, 3, 0).reshape(Ci, K), KP)
            dx = torch.empty(xshape, device=dev, dtype=torch.float32)
            ws, nws = _workspace(d1, 0, dev, dkey1)
            call("mog_conv2d_fwd", C.byref(d1), None, dzP.data_ptr(), _pack_matrix(w2, d1).data_ptr(), None, dx.data_ptr(), _ptr(ws), nws, st)
        if need_dw:
            g = torch.empty((Ci, KP), device=dev, dtype=torch.float32)
            ws, nws = _workspace(d1, 2, dev, dkey1)
            call("mog_conv2d_wgrad", C.byref(d1), None, dzP.data_ptr(), None, xp.data_ptr(), g.data_ptr(), None, _ptr(ws), nws, st)
            dw = g[:, :K].reshape(Ci, KH, KW, Co).flip(1, 2).permute(3, 0, 1, 2).contiguous()
        return dx, dw, None, None, None, None, None, None


def _tile_eff(H, W):
    """Fraction of the 8 (w) x 16 (h) pixel sub-tiles of the halo-tile kernels that holds real pixels."""
    return (H * W) / float(((H + 15) // 16) * ((W + 7) // 8) * 128)


def _retile(x, weight, stride, pad, up2x, precision):
    """Pixel-grid view that tiles better, for filters without taps along an axis (tcgen05 precisions).  The halo-tile kernels
    cover an image with 8 x 16 pixel sub-tiles: a 17 x 17 map of the DAMSM image encoder fills 38 % of its 6 sub-tiles.
    A 1 x 1 filter does not care which pixels are neighbours -- all N*H*W pixels become one [M/8, 8] grid (~100 %); a 1 x k
    filter (taps along W only) lets the rows of all images stack into one [N*H, W] image (17 wide: 71 %).  Pure views: the
    kernels, the packed weights and the bf16 planes ([pixels][C8]) are unchanged."""
    if precision == PREC_FP32 or up2x or stride != 1 or x.dim() != 4 or weight.dim() != 4:
        return None
    N, H, W, Cc = x.shape
    KH, KW = weight.shape[2], weight.shape[3]
    ph, pw = (pad if isinstance(pad, (tuple, list)) else (pad, pad))
    if KH != 1 or ph != 0 or N * H * W < 4096:
        return None
    M = N * H * W
    if KW == 1 and pw == 0 and M % 8 == 0:
        shape = (1, M // 8, 8, Cc)
    elif N > 1:
        shape = (1, N * H, W, Cc)
    else:
        return None
    if _tile_eff(shape[1], shape[2]) < 1.15 * _tile_eff(H, W):
        return None
    return shape


def conv2d(x, weight, bias=None, stride=1, pad=0, up2x=False, act=ACT_NONE, precision=None):
    if precision is None:
        precision = _default_precision
    if _thin_cin(x, weight, stride, pad, up2x, precision):
        return PatchConv2dFn.apply(x, weight, bias, stride, pad, act, precision)
    shape = _retile(x, weight, stride, pad, up2x, precision)
    if shape is not None:
        N, H, W, _ = x.shape
        xv = x.reshape(shape)
        cached = getattr(x, "_mog_planes", None)
        if cached is not None and cached[2] == x._version and cached[3] == x.data_ptr():
            xv._mog_planes = (cached[0], cached[1], xv._version, xv.data_ptr())     # same rows, same planes
        y = Conv2dFn.apply(xv, weight, bias, stride, pad, False, act, precision)
        return y.reshape(N, H, W, y.shape[-1])
    return Conv2dFn.apply(x, weight, bias, stride, pad, bool(up2x), act, precision)


def linear(x, weight, bias=None, act=ACT_NONE, precision=None):
    """x [N, F] -> [N, O] through the same implicit-GEMM kernels (H=W=KH=KW=1)."""
    N, Fi = x.shape
    y = conv2d(x.reshape(N, 1, 1, Fi), weight, bias, 1, 0, False, act, precision)
    return y.reshape(N, -1)


# ---------------------------------------------------------------------------------------------
# BatchNorm (train) + activation (+ residual)
# ---------------------------------------------------------------------------------------------
_bn_parts_cache = {}


def _bn_parts(S, M, Cc, act, which):
    k = (S, M, Cc, act, which)
    v = _bn_parts_cache.get(k)
    if v is None:
        v = _bn_parts_cache[k] = _lib.lib().mog_bn_parts(S, M, Cc, act, which)
    return v


class BnActFn(torch.autograd.Function):
    """y = act(BN_train(x)) (+ residual) over rows [S*M, C] with per-segment statistics."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, residual, S, act, momentum, eps, planes, precision):
        _chk(x, "bn input")
        Cc = x.shape[-1]
        rows = x.numel() // Cc
        if rows % S:
            raise RuntimeError("bn: %d rows not divisible by %d segments" % (rows, S))
        M = rows // S
        dev = x.device
        st = _stream()
        P = _bn_parts(S, M, Cc, act, 0)
        part = torch.empty((P, 2, S, Cc), device=dev, dtype=torch.float64)     # per-block partial sums (summed in a fixed order)
        call("mog_bn_stats", x.data_ptr(), S, M, Cc, part.data_ptr(), P, st)
        mis = torch.empty((4, S, Cc), device=dev, dtype=torch.float32)  # mean, invstd, scale, shift
        g = gamma.detach().contiguous()
        b = beta.detach().contiguous()
        call("mog_bn_finalize", part.data_ptr(), P, S, M, Cc, g.data_ptr(), b.data_ptr(),
             eps, momentum, _ptr(running_mean), _ptr(running_var), mis[0].data_ptr(), mis[1].data_ptr(),
             mis[2].data_ptr(), mis[3].data_ptr(), st)
        Co = Cc // 2 if act == ACT_GLU else Cc
        y = torch.empty(x.shape[:-1] + (Co,), device=dev, dtype=torch.float32)
        if residual is not None:
            _chk(residual, "bn residual")
        call("mog_affine_act_fwd_planes", x.data_ptr(), mis[2].data_ptr(), mis[3].data_ptr(), _ptr(residual),
             y.data_ptr(), _ptr(planes), precision, S, M, Cc, act, st)
        ctx.cfg = (S, M, Cc, act)
        ctx.precision = precision
        ctx.has_res = residual is not None
        ctx.save_for_backward(x, g, b, mis)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, g, b, mis = ctx.saved_tensors
        S, M, Cc, act = ctx.cfg
        dy = dy.contiguous()
        st = _stream()
        dev = x.device
        P = _bn_parts(S, M, Cc, act, 1)
        red = torch.empty((P + 1, 2, S, Cc), device=dev, dtype=torch.float64)   # P partials + the per-segment sums
        dgb = torch.empty((2, Cc), device=dev, dtype=torch.float32)
        call("mog_bn_act_bwd_reduce", x.data_ptr(), dy.data_ptr(), mis[0].data_ptr(), mis[1].data_ptr(),
             g.data_ptr(), b.data_ptr(), S, M, Cc, act, red.data_ptr(), P, red[P, 0].data_ptr(), red[P, 1].data_ptr(),
             dgb[0].data_ptr(), dgb[1].data_ptr(), st)
        dx = torch.empty_like(x)
        # dx usually is the output gradient of a convolution: emit it as bf16 planes too (saves that conv's split pass)
        prec = ctx.precision
        planes = None
        if prec != PREC_FP32 and x.dim() == 4 and Cc % 8 == 0:
            planes = torch.empty((_planes_bytes(S * M, Cc, prec) + 3) // 4, device=dev, dtype=torch.float32)
        call("mog_bn_act_bwd_apply_planes", x.data_ptr(), dy.data_ptr(), mis[0].data_ptr(), mis[1].data_ptr(),
             g.data_ptr(), b.data_ptr(), red[P, 0].data_ptr(), red[P, 1].data_ptr(), S, M, Cc, act, dx.data_ptr(), _ptr(planes), prec,
             st)
        if planes is not None:
            _attach_planes(dx, planes, prec)
        dres = dy if ctx.has_res else None
        return dx, dgb[0], dgb[1], None, None, dres, None, None, None, None, None, None


def bn_act(x, bn, act=ACT_NONE, residual=None, segments=1):
    """Apply a train-mode ``nn.BatchNorm*d``-compatible module ``bn`` (weight, bias,
    running_mean, running_var, num_batches_tracked, momentum, eps) followed by ``act``."""
    if not bn.training:
        return _bn_act_eval(x, bn, act, residual)
    if bn.num_batches_tracked is not None:
        bn.num_batches_tracked += segments
    # in the tensor-core precisions the apply kernel also emits y as pre-split bf16 planes: the usual consumer is a conv
    precision = _default_precision
    Co = x.shape[-1] // 2 if act == ACT_GLU else x.shape[-1]
    planes = None
    if precision != PREC_FP32 and x.dim() == 4 and Co % 4 == 0 and x.shape[-1] % 4 == 0 and x.is_cuda:
        planes = _alloc_planes(tuple(x.shape[:-1]) + (Co,), precision, x.device)
    y = BnActFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, residual, segments, act,
                      float(bn.momentum), float(bn.eps), planes, precision)
    if planes is not None:
        _attach_planes(y, planes, precision)
    return y


def _bn_act_eval(x, bn, act, residual):
    """Inference form (``netG.eval()`` in sampling / gen_example, trainer.py:399,504,599): y = act(x * scale + shift) with
    the running statistics folded into per-channel scale / shift; forward only."""
    if torch.is_grad_enabled() and (x.requires_grad or bn.weight.requires_grad):
        raise RuntimeError("libmog: eval-mode BatchNorm is forward-only (sampling); wrap the call in torch.no_grad()")
    _chk(x, "bn input")
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    with torch.no_grad():
        scale = (bn.weight * torch.rsqrt(bn.running_var + bn.eps)).contiguous()
        shift = (bn.bias - bn.running_mean * scale).contiguous()
        Co = Cc // 2 if act == ACT_GLU else Cc
        y = torch.empty(x.shape[:-1] + (Co,), device=x.device, dtype=torch.float32)
        if residual is not None:
            _chk(residual, "bn residual")
        call("mog_affine_act_fwd", x.data_ptr(), scale.data_ptr(), shift.data_ptr(), _ptr(residual), y.data_ptr(), 1, rows, Cc,
             act, _stream())
    return y


# ---------------------------------------------------------------------------------------------
# spatial transformer
# ---------------------------------------------------------------------------------------------
_STN_CHECK = os.environ.get("MOG_DEBUG", "0") == "1"


class StnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, theta, extra, mode, B, S, Ho, Wo, align):
        _chk(x, "stn input")
        theta = theta.contiguous()
        _chk(theta, "stn theta")
        n_in, Hi, Wi, Cc = x.shape
        if theta.shape != (B, S, 2, 3):
            raise RuntimeError("stn: theta must be [B,S,2,3], got %s" % (tuple(theta.shape),))
        if _STN_CHECK and bool((theta[..., 0, 1] != 0).any() | (theta[..., 1, 0] != 0).any()):
            # (host sync: only with MOG_DEBUG=1) the backward kernel builds separable row / column tables
            raise RuntimeError("stn: only axis-aligned theta (no rotation / shear terms) is supported, as produced by "
                               "compute_transformation_matrix[_inverse] (miscc/utils.py:16-49)")
        if n_in != (S * B if mode == 0 else B):
            raise RuntimeError("stn: input batch %d inconsistent with mode %d, B=%d, S=%d" % (n_in, mode, B, S))
        Cy = Cc
        if extra is not None:
            extra = extra.contiguous()
            _chk(extra, "stn extra")
            Cy = Cc + extra.shape[-1]
        n_out = B if mode == 0 else S * B
        y = torch.empty((n_out, Ho, Wo, Cy), device=x.device, dtype=torch.float32)
        call("mog_stn_fwd", x.data_ptr(), theta.data_ptr(), _ptr(extra), y.data_ptr(), mode, B, S, Hi, Wi, Cc,
             Ho, Wo, Cy, int(align), _stream())
        ctx.cfg = (mode, B, S, Hi, Wi, Cc, Ho, Wo, Cy, int(align))
        ctx.save_for_backward(theta)
        return y

    @staticmethod
    def backward(ctx, dy):
        (theta,) = ctx.saved_tensors
        mode, B, S, Hi, Wi, Cc, Ho, Wo, Cy, align = ctx.cfg
        dx = None
        if ctx.needs_input_grad[0]:
            dy = dy.contiguous()
            n_in = S * B if mode == 0 else B
            dx = torch.empty((n_in, Hi, Wi, Cc), device=dy.device, dtype=torch.float32)
            call("mog_stn_bwd", dy.data_ptr(), theta.data_ptr(), dx.data_ptr(), mode, B, S, Hi, Wi, Cc, Ho, Wo,
                 Cy, align, _stream())
        return dx, None, None, None, None, None, None, None, None


def stn_scatter_sum(x, theta, B, S, out_hw, align_corners=False):
    """x [S*B,Hi,Wi,C] (segment-major), theta [B,S,2,3] -> [B,Ho,Wo,C] = sum_s stn(x[s*B+b], theta[b,s])."""
    return StnFn.apply(x, theta, None, 0, B, S, out_hw[0], out_hw[1], align_corners)


def stn_crop(x, theta, S, out_hw, extra=None, align_corners=False):
    """x [B,Hi,Wi,C], theta [B,S,2,3] -> [S*B,Ho,Wo,C(+E)]; ``extra`` [B,S,E] is broadcast into
    the trailing channels (the label planes of D_NET64's object pathway)."""
    return StnFn.apply(x, theta, extra, 1, x.shape[0], S, out_hw[0], out_hw[1], align_corners)


# ---------------------------------------------------------------------------------------------
# word attention
# ---------------------------------------------------------------------------------------------
class WordAttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, src, mask, quirk, want_attn):
        _chk(h, "attention h")
        _chk(src, "attention src")
        B, Q, D = h.shape
        T = src.shape[1]
        out = torch.empty_like(h)
        attn = torch.empty((B, T, Q), device=h.device, dtype=torch.float32) if want_attn else None
        m = None
        if mask is not None:
            if mask.dim() != 2 or mask.shape[0] != B or mask.shape[1] < T:
                raise RuntimeError("word attention: mask must be [B=%d, >=T=%d], got %s" % (B, T, tuple(mask.shape)))
            m = mask[:, :T].to(torch.uint8).contiguous()   # a caption mask wider than the word embeddings is trimmed (trainer.py:288-289)
        call("mog_word_attention_fwd", h.data_ptr(), src.data_ptr(), _ptr(m), out.data_ptr(), _ptr(attn),
             B, Q, D, T, int(quirk), _stream())
        ctx.cfg = (B, Q, D, T, int(quirk))
        ctx.save_for_backward(h, src, m)
        if want_attn:
            ctx.mark_non_differentiable(attn)
            return out, attn
        return out, None

    @staticmethod
    def backward(ctx, dout, _dattn):
        h, src, m = ctx.saved_tensors
        B, Q, D, T, quirk = ctx.cfg
        dout = dout.contiguous()
        dh = torch.empty_like(h)
        nws = _lib.lib().mog_word_attention_bwd_workspace_bytes(B, Q, D, T)
        ws = torch.empty((nws + 3) // 4, device=h.device, dtype=torch.float32) if nws else None
        dsrc = torch.empty_like(src) if nws else torch.zeros_like(src)     # (scalar fallback kernel accumulates atomically)
        call("mog_word_attention_bwd", h.data_ptr(), src.data_ptr(), _ptr(m), dout.data_ptr(), dh.data_ptr(),
             dsrc.data_ptr(), B, Q, D, T, quirk, _ptr(ws), nws, _stream())
        return dh, dsrc, None, None, None


def word_attention(h, src, mask=None, mask_quirk=True, want_attn=True):
    """h [B,Q,D], src [B,T,D] -> (weighted context [B,Q,D], attn [B,T,Q] or None)."""
    return WordAttnFn.apply(h, src, mask, mask_quirk, want_attn)


# ---------------------------------------------------------------------------------------------
# sigmoid + BCE head
# ---------------------------------------------------------------------------------------------
class SigmoidBceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, target, weight, with_logits):
        z = z.contiguous()
        _chk(z, "bce logits")
        target = target.detach().to(torch.float32).contiguous()
        _chk(target, "bce target")
        if target.numel() != z.numel():
            raise RuntimeError("bce: %d logits vs %d targets" % (z.numel(), target.numel()))
        loss = torch.empty(1, device=z.device, dtype=torch.float32)
        call("mog_sigmoid_bce_fwd", z.data_ptr(), target.data_ptr(), float(weight), z.numel(), None,
             loss.data_ptr(), 0, int(with_logits), _stream())
        ctx.weight = float(weight)
        ctx.with_logits = int(with_logits)
        ctx.save_for_backward(z, target)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        z, target = ctx.saved_tensors
        g = g.contiguous().reshape(1)
        dz = torch.empty_like(z)
        call("mog_sigmoid_bce_bwd", z.data_ptr(), target.data_ptr(), ctx.weight, z.numel(), g.data_ptr(),
             dz.data_ptr(), ctx.with_logits, _stream())
        return dz, None, None, None


def sigmoid_bce(z, target, weight: float = 1.0, with_logits: bool = False):
    """weight * mean BCE(sigmoid(z), target): nn.Sigmoid + nn.BCELoss of the AttnGAN heads, or
    (with_logits) nn.BCEWithLogitsLoss of the StackGAN / CLEVR / Multi-MNIST programs."""
    return SigmoidBceFn.apply(z, target, weight, with_logits)


# ---------------------------------------------------------------------------------------------
# plain activation (no BatchNorm), e.g. the GLU of CA_NET (model.py:328)
# ---------------------------------------------------------------------------------------------
class ActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, act):
        _chk(x, "act input")
        Cc = x.shape[-1]
        rows = x.numel() // Cc
        Co = Cc // 2 if act == ACT_GLU else Cc
        y = torch.empty(x.shape[:-1] + (Co,), device=x.device, dtype=torch.float32)
        call("mog_affine_act_fwd", x.data_ptr(), None, None, None, y.data_ptr(), 1, rows, Cc, act, _stream())
        ctx.act = act
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = dy.contiguous()
        Cc = x.shape[-1]
        rows = x.numel() // Cc
        dx = torch.empty_like(x)
        call("mog_bn_act_bwd_apply", x.data_ptr(), dy.data_ptr(), None, None, None, None, None, None, 1, rows, Cc,
             ctx.act, dx.data_ptr(), _stream())
        return dx, None


def activation(x, act):
    return ActFn.apply(x, act)


# ---------------------------------------------------------------------------------------------
# DAMSM word-region matching
# ---------------------------------------------------------------------------------------------
class DamsmSimsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, words, lens, g1, g2):
        _chk(feat, "damsm region features")
        words = words.detach().contiguous()
        _chk(words, "damsm words")
        B, R, D = feat.shape
        NI, D2, Tw = words.shape
        if D != D2:
            raise RuntimeError("damsm: feature dim %d vs word dim %d" % (D, D2))
        lens = lens.to(device=feat.device, dtype=torch.int32).contiguous()
        sims = torch.empty((B, NI), device=feat.device, dtype=torch.float32)
        call("mog_damsm_words_fwd", feat.data_ptr(), words.data_ptr(), lens.data_ptr(), sims.data_ptr(), None, None,
             B, NI, R, D, Tw, 0, float(g1), float(g2), _stream())
        ctx.cfg = (B, NI, R, D, Tw, float(g1), float(g2))
        ctx.save_for_backward(feat, words, lens)
        return sims

    @staticmethod
    def backward(ctx, dsims):
        feat, words, lens = ctx.saved_tensors
        B, NI, R, D, Tw, g1, g2 = ctx.cfg
        dsims = dsims.contiguous()
        dfeat = torch.empty_like(feat)
        nws = _lib.lib().mog_damsm_bwd_workspace_bytes(B, NI, R, D)
        ws = torch.empty((nws + 3) // 4, device=feat.device, dtype=torch.float32)
        call("mog_damsm_words_bwd", feat.data_ptr(), words.data_ptr(), lens.data_ptr(), dsims.data_ptr(),
             dfeat.data_ptr(), B, NI, R, D, Tw, g1, g2, ws.data_ptr(), nws, _stream())
        return dfeat, None, None, None, None


def damsm_similarities(feat, words, lens, gamma1, gamma2):
    """feat [B,R,D] (NHWC regions), words [NI,D,Tw], lens [NI] -> sims [B,NI] (before gamma3)."""
    return DamsmSimsFn.apply(feat, words, lens, gamma1, gamma2)


def func_attention_paired(query, context_nhwc, gamma1):
    """GlobalAttention.func_attention for pairs (b, b): query [B,D,Tq], context [B,R,D] ->
    (weightedContext [B,D,Tq], attn [B,Tq,R]).  Forward only."""
    B, R, D = context_nhwc.shape
    Tq = query.shape[2]
    q = query.detach().contiguous()
    lens = torch.full((B,), Tq, device=q.device, dtype=torch.int32)
    wei = torch.zeros((B, D, Tq), device=q.device, dtype=torch.float32)
    attn = torch.zeros((B, Tq, R), device=q.device, dtype=torch.float32)
    call("mog_damsm_words_fwd", context_nhwc.detach().contiguous().data_ptr(), q.data_ptr(), lens.data_ptr(), None,
         wei.data_ptr(), attn.data_ptr(), B, B, R, D, Tq, 1, float(gamma1), 1.0, _stream())
    return wei, attn


# ---------------------------------------------------------------------------------------------
# pooling / bilinear resize (DAMSM image encoder)
# ---------------------------------------------------------------------------------------------
class Pool2dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k, stride, pad, mode):
        _chk(x, "pool input")
        N, H, W, Cc = x.shape
        ho, wo = C.c_int(), C.c_int()
        call("mog_pool2d_out_hw", H, W, k, stride, pad, C.byref(ho), C.byref(wo))
        y = torch.empty((N, ho.value, wo.value, Cc), device=x.device, dtype=torch.float32)
        # max pooling records each window's arg-max position (1 byte per output) instead of keeping x for backward
        idx = torch.empty(y.shape, device=x.device, dtype=torch.uint8) if (mode == 0 and ctx.needs_input_grad[0]) else None
        call("mog_pool2d_fwd", x.data_ptr(), y.data_ptr(), _ptr(idx), N, H, W, Cc, k, stride, pad, mode, _stream())
        ctx.cfg = (N, H, W, Cc, k, stride, pad, mode)
        ctx.save_for_backward(idx)
        return y

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        N, H, W, Cc, k, stride, pad, mode = ctx.cfg
        dy = dy.contiguous()
        dx = torch.empty((N, H, W, Cc), device=dy.device, dtype=torch.float32)
        call("mog_pool2d_bwd", None, _ptr(idx), dy.data_ptr(), dx.data_ptr(), N, H, W, Cc, k, stride, pad, mode, _stream())
        return dx, None, None, None, None


def max_pool2d(x, kernel_size, stride=None, padding=0):
    """F.max_pool2d on NHWC."""
    return Pool2dFn.apply(x, kernel_size, stride or kernel_size, padding, 0)


def avg_pool2d(x, kernel_size, stride=None, padding=0):
    """F.avg_pool2d (count_include_pad=True) on NHWC."""
    return Pool2dFn.apply(x, kernel_size, stride or kernel_size, padding, 1)


class ResizeBilinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, Ho, Wo, align):
        _chk(x, "resize input")
        N, Hi, Wi, Cc = x.shape
        y = torch.empty((N, Ho, Wo, Cc), device=x.device, dtype=torch.float32)
        call("mog_resize_bilinear_fwd", x.data_ptr(), y.data_ptr(), N, Hi, Wi, Cc, Ho, Wo, int(align), _stream())
        ctx.cfg = (N, Hi, Wi, Cc, Ho, Wo, int(align))
        return y

    @staticmethod
    def backward(ctx, dy):
        N, Hi, Wi, Cc, Ho, Wo, align = ctx.cfg
        dy = dy.contiguous()
        dx = torch.empty((N, Hi, Wi, Cc), device=dy.device, dtype=torch.float32)
        call("mog_resize_bilinear_bwd", dy.data_ptr(), dx.data_ptr(), N, Hi, Wi, Cc, Ho, Wo, align, _stream())
        return dx, None, None, None


def resize_bilinear(x, size, align_corners=False):
    """nn.Upsample(size=size, mode='bilinear') on NHWC."""
    return ResizeBilinearFn.apply(x, int(size[0]), int(size[1]), bool(align_corners))
