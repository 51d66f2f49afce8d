"""Fused Adam (+ EMA) on libmog -- drop-in for ``torch.optim.Adam`` as the reference uses it
(``optim.Adam(params, lr, betas=(0.5, 0.999))``: attngan/trainer.py:141-148, stackgan/trainer.py:136-137,
multi-mnist/trainer.py:103-104, clevr/trainer.py:100-101) and for the EMA loop of attngan/trainer.py:341-342.

Same constructor, ``param_groups`` and ``state_dict`` layout as ``torch.optim.Adam`` (per-parameter ``step``,
``exp_avg``, ``exp_avg_sq``), so checkpoints written by either load in the other.  ``step()`` is ONE
``mog_adam_multi`` call per parameter group (a handful of launches for a whole network instead of ~7 foreach
kernels per 30 tensors); the EMA copy of the generator is updated in the same pass when ``ema_params`` is given.
There is no CPU path: parameters must live on a CUDA device.

The step count the bias corrections need lives in DEVICE memory (one double per parameter group, incremented by the
kernel launch itself: ``mog_adam_multi_dev``), so a captured CUDA graph replays the optimiser step correctly; the
per-parameter ``state['step']`` tensors of the ``state_dict`` are kept in step on the host (``advance_host_steps`` after
a graph replay).
"""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import call


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if weight_decay != 0 or amsgrad:
            raise ValueError("mog_b200.optim.Adam implements the reference's configuration only (no weight decay / amsgrad)")
        if not 0.0 <= lr or not 0.0 <= eps or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("invalid Adam hyper-parameter")
        # the extra keys mirror torch.optim.Adam's defaults so that state_dicts interchange
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=0, amsgrad=False, maximize=False,
                                      foreach=None, capturable=False, differentiable=False, fused=None,
                                      decoupled_weight_decay=False))

    @torch.no_grad()
    def step(self, closure=None, ema_params=None, ema_decay=0.999, grad_scale=1.0):
        """``ema_params``: optional dict ``param -> EMA tensor`` (or a list parallel to all parameters of all groups):
        ``ema = ema_decay * ema + (1 - ema_decay) * param_new`` fused into the same pass."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if ema_params is not None and not isinstance(ema_params, dict):
            allp = [p for g in self.param_groups for p in g["params"]]
            ema_params = dict(zip(allp, ema_params))
        for gi, group in enumerate(self.param_groups):
            ps, gs, ms, vs, es, ns = [], [], [], [], [], []
            step = None
            for p in group["params"]:
                if p.grad is None:
                    # no Adam update, but the reference's EMA loop (trainer.py:341-342) still averages every parameter
                    e = ema_params.get(p) if ema_params is not None else None
                    if e is not None:
                        e.mul_(ema_decay).add_(p.detach(), alpha=1.0 - ema_decay)
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("mog_b200.optim.Adam: parameters must be contiguous fp32 CUDA tensors (no CPU fallback)")
                g = p.grad
                if not g.is_contiguous():
                    g = g.contiguous()
                state = self.state[p]
                if len(state) == 0:
                    state["step"] = torch.tensor(0.0, dtype=torch.float32)
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                state["step"] += 1
                t = int(state["step"].item()) if state["step"].device.type == "cpu" else int(state["step"])
                if step is None:
                    step = t
                elif step != t:   # parameters of one group at different step counts: flush what we have
                    self._launch(group, ps, gs, ms, vs, es, ns, step, ema_decay, grad_scale)
                    ps, gs, ms, vs, es, ns = [], [], [], [], [], []
                    step = t
                e = ema_params.get(p) if ema_params is not None else None
                ps.append(p.data_ptr()); gs.append(g.data_ptr()); ms.append(state["exp_avg"].data_ptr())
                vs.append(state["exp_avg_sq"].data_ptr()); es.append(e.data_ptr() if e is not None else None)
                ns.append(p.numel())
                state["_keep"] = g   # the gradient must outlive the asynchronous launch
                # the kernel writes p through its raw pointer: tell the packed-weight cache (ops._packed) that p changed
                p._mog_ver = getattr(p, "_mog_ver", 0) + 1
            if ps:
                self._launch(group, ps, gs, ms, vs, es, ns, step, ema_decay, grad_scale, dev_slot=gi)
        updated = []
        for group in self.param_groups:
            for p in group["params"]:
                if p in self.state and self.state[p].pop("_keep", None) is not None:
                    updated.append(p)
        # the convolutions' packed operands of every updated weight, refreshed in one multi-tensor launch
        from . import ops
        ops.repack(updated)
        return loss

    def advance_host_steps(self, n=1):
        """After ``n`` replays of a CUDA graph that captured ``step()``: the device counters advanced by themselves, the
        per-parameter ``state['step']`` of the ``state_dict`` (host tensors) are brought along."""
        for slot in getattr(self, "_dev_steps", {}).values():
            slot[1] += n
        for group in self.param_groups:
            for p in group["params"]:
                st = self.state.get(p)
                if st and "step" in st:
                    st["step"] += n

    def _launch(self, group, ps, gs, ms, vs, es, ns, step, ema_decay, grad_scale, dev_slot=None):
        n = len(ps)
        arr = lambda xs: (C.c_void_p * n)(*xs)
        b1, b2 = group["betas"]
        stream = torch.cuda.current_stream().cuda_stream
        ema = arr(es) if any(e is not None for e in es) else None
        if dev_slot is not None:
            # all parameters of the group are at the same step (always, in the reference's loops): device-resident counter
            # [tensor, host mirror], kept outside param_groups / state so that state_dict() stays torch.optim.Adam's.
            # It is (re)synchronised with the host count only when they disagree (first step, after load_state_dict).
            slots = self.__dict__.setdefault("_dev_steps", {})
            slot = slots.get(dev_slot)
            if slot is None:
                slot = slots[dev_slot] = [torch.zeros(1, dtype=torch.float64, device="cuda"), 0]
            if slot[1] != step - 1:
                if torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("mog_b200.optim.Adam: the device step counter is out of sync inside a CUDA graph "
                                       "capture -- run at least one eager step before capturing")
                slot[0].fill_(float(step - 1))
            call("mog_adam_multi_dev", n, arr(ps), arr(gs), arr(ms), arr(vs), ema, (C.c_longlong * n)(*ns), float(group["lr"]),
                 float(b1), float(b2), float(group["eps"]), slot[0].data_ptr(), float(ema_decay), float(grad_scale), stream)
            slot[1] = step
            return
        call("mog_adam_multi", n, arr(ps), arr(gs), arr(ms), arr(vs), ema, (C.c_longlong * n)(*ns), float(group["lr"]), float(b1),
             float(b2), float(group["eps"]), int(step), float(ema_decay), float(grad_scale), stream)
