"""Drop-in replacements for the optimiser-side loops of the pre-training caller.

  * ``LARS`` / ``add_weight_decay``  -- ``pixcontrast_18/contrast/lars.py`` (same names, constructor, wrapper
    semantics: ``LARS(torch.optim.SGD(add_weight_decay(model, wd), lr, momentum))``).  The reference calls
    ``p.norm()`` / ``p.grad.norm()`` per tensor and branches on the results in Python (:127-133), a host
    synchronisation for every weight tensor and step; here each parameter group is two multi-tensor
    launches (norms, update) per 48 tensors with every decision taken on the device
    (``stswin_lars_sgd_step``).  State stays in the wrapped optimiser (``momentum_buffer``), so
    ``state_dict`` / ``load_state_dict`` interoperate with the reference.
  * ``momentum_update``              -- the key-encoder EMA of ``PixPro._momentum_update_key_encoder``
    (``PixPro_swin_v5.py:258-289``): one multi-tensor stream instead of three eager ops per parameter.

CUDA fp32 parameters only -- no CPU path.
"""
from __future__ import annotations

import ctypes
from typing import Iterable, Sequence

import torch
from torch.optim.optimizer import Optimizer

from . import _lib, ops
from ._lib import StswinError

__all__ = ["LARS", "FusedAdam", "add_weight_decay", "momentum_update", "bf16_weight"]


def add_weight_decay(model, weight_decay=1e-5, skip_list=()):
    """The two parameter groups of ``lars.py:7-32``: vectors (biases, norm scales) and anything named in ``skip_list``
    get no weight decay and are exempt from LARS scaling (``ignore``); every other trainable tensor decays."""
    groups = {True: [], False: []}
    for name, param in model.named_parameters():
        if param.requires_grad:
            groups[param.dim() == 1 or name in skip_list].append(param)
    return [dict(params=groups[True], weight_decay=0, ignore=True),
            dict(params=groups[False], weight_decay=weight_decay, ignore=False)]


def _check_f32(tensors: Iterable[torch.Tensor], what: str) -> None:
    for t in tensors:
        if not t.is_cuda:
            raise StswinError(f"{what}: tensors must be CUDA tensors (stswincl_b200 has no CPU path)")
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise StswinError(f"{what}: tensors must be contiguous fp32, got {t.dtype} (contiguous={t.is_contiguous()})")


def _ptrs(tensors: Sequence[torch.Tensor]):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def _numels(tensors: Sequence[torch.Tensor]):
    return (ctypes.c_int64 * len(tensors))(*[t.numel() for t in tensors])


@torch.no_grad()
def momentum_update(params_q: Iterable[torch.Tensor], params_k: Iterable[torch.Tensor], momentum: float) -> None:
    """``param_k = param_k * momentum + param_q * (1. - momentum)`` for every pair, in place
    (``PixPro_swin_v5.py:266-267``), bit-exact with the eager expression."""
    q = [p.data for p in params_q]
    k = [p.data for p in params_k]
    if len(q) != len(k):
        raise ValueError(f"momentum_update: {len(q)} query tensors vs {len(k)} key tensors")
    if not q:
        return
    _check_f32(q, "momentum_update"); _check_f32(k, "momentum_update")
    for a, b in zip(q, k):
        if a.shape != b.shape or a.device != b.device:
            raise ValueError("momentum_update: query / key parameter mismatch")
    n_bytes = sum(t.numel() for t in k) * 4
    with ops._launch("ema_update", 3.0 * n_bytes, k[0]):                     # read k, q; write k
        st = _lib.load().stswin_ema_update(_ptrs(k), _ptrs(q), _numels(k), len(k), float(momentum), float(1. - momentum),
                                           ops._stream(k[0]))
    _lib.check(st, "stswin_ema_update")


class LARS(Optimizer):
    """'LARS (Layer-wise Adaptive Rate Scaling)' as a wrapper of ``torch.optim.SGD`` (``lars.py:34-152``)."""

    def __init__(self, optimizer, eps=1e-8, trust_coef=0.001):
        if eps < 0.0:
            raise ValueError('invalid epsilon value: , %f' % eps)
        if trust_coef < 0.0:
            raise ValueError("invalid trust coefficient: %f" % trust_coef)
        if not isinstance(optimizer, torch.optim.SGD):
            raise StswinError("LARS: the fused step wraps torch.optim.SGD (the optimiser main_pretrain_swinv5.py:43-47 builds)")
        self.optim = optimizer
        self.eps = eps
        self.trust_coef = trust_coef

    def __getstate__(self):
        return (self.optim, {'eps': self.eps, 'trust_coef': self.trust_coef})

    def __setstate__(self, state):
        self.optim, lars_dict = state
        self.eps = lars_dict['eps']
        self.trust_coef = lars_dict['trust_coef']

    def __repr__(self):
        return '%s(%r)' % (self.__class__.__name__, self.optim)

    @property
    def param_groups(self):
        return self.optim.param_groups

    @property
    def state(self):
        return self.optim.state

    def state_dict(self):
        return self.optim.state_dict()

    def load_state_dict(self, state_dict):
        self.optim.load_state_dict(state_dict)

    def zero_grad(self, set_to_none: bool = True):
        self.optim.zero_grad(set_to_none=set_to_none)

    def add_param_group(self, param_group):
        self.optim.add_param_group(param_group)

    @torch.no_grad()
    def step(self, closure=None):
        """Weight decay into the gradient, LARS scaling of the non-``ignore`` group, momentum SGD update
        (``lars.py:109-152``); ``p.grad`` holds the decayed / scaled gradient afterwards, as in the reference."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for group in self.optim.param_groups:
            params = [p for p in group['params'] if p.grad is not None]
            if not params:
                continue
            if group.get('maximize', False):
                raise StswinError("LARS: maximize=True is not supported")
            grads = [p.grad for p in params]
            _check_f32(params, "LARS.step"); _check_f32(grads, "LARS.step (grad)")
            momentum, dampening = float(group['momentum']), float(group['dampening'])
            ignore = group.get('ignore', None)
            lars = ignore is not None and not ignore          # lars.py:125
            bufs, first = None, None
            if momentum != 0:
                bufs, flags = [], []
                for p in params:
                    st = self.optim.state[p]
                    buf = st.get('momentum_buffer')
                    flags.append(1 if buf is None else 0)
                    if buf is None:
                        buf = st['momentum_buffer'] = torch.empty_like(p, memory_format=torch.contiguous_format)
                    bufs.append(buf)
                _check_f32(bufs, "LARS.step (momentum_buffer)")
                first = (ctypes.c_uint8 * len(flags))(*flags)
            dev = params[0].device
            norms = torch.empty(2 * len(params), dtype=torch.float64, device=dev) if lars else None
            n_bytes = sum(p.numel() for p in params) * 4
            passes = (2 if lars else 0) + 4 + (2 if momentum != 0 else 0)       # norms: p, g; update: p, g in / out (+ buf)
            with ops._launch("lars_sgd_step", float(passes * n_bytes), params[0]):
                st = lib.stswin_lars_sgd_step(_ptrs(params), _ptrs(grads), _ptrs(bufs) if bufs else None, _numels(params), first,
                                              len(params), float(group['lr']), momentum, dampening, int(bool(group['nesterov'])),
                                              float(group['weight_decay']), int(lars), float(self.trust_coef), float(self.eps),
                                              norms.data_ptr() if norms is not None else None, ops._stream(params[0]))
            _lib.check(st, "stswin_lars_sgd_step")
        return loss


# ---------------------------------------------------------------------------------------------------------------
# bf16 shadow of the weights
# ---------------------------------------------------------------------------------------------------------------
def bf16_weight(w: torch.Tensor) -> torch.Tensor:
    """The bf16 copy of a weight the GEMM kernels consume.  ``FusedAdam`` keeps a shadow copy per parameter that its
    update kernel rewrites every step, so the forward needs no cast kernel; the shadow is valid as long as nobody else
    wrote the parameter (tracked by the tensor's version counter -- the kernel does not bump it).  Without a valid
    shadow this is ``w.to(bfloat16)``."""
    sh = getattr(w, "_stswin_shadow", None)
    if sh is not None:
        if sh[1] == w._version and sh[0].device == w.device:
            return sh[0]
        with torch.no_grad():                       # someone else updated the parameter: refresh
            sh[0].copy_(w)
        w._stswin_shadow = (sh[0], w._version)
        return sh[0]
    return w.to(torch.bfloat16)


class FusedAdam(Optimizer):
    """``torch.optim.Adam`` semantics (the optimiser of ``seg18/train_swin.py``; no amsgrad / maximize) as multi-tensor
    launches of ``stswin_adam_step``: 32 tensors per launch, step count on the device (CUDA-graph capturable), and the
    bf16 shadow of every >= 2-D weight written in the same pass (``bf16_weight``).  State keys ``exp_avg`` /
    ``exp_avg_sq`` as in torch; the step count is one device scalar per group (``group['step_t']``).

    ``step(grads=...)`` takes the gradients from a list (fp32 or bf16, one per parameter with a gradient) instead of
    ``p.grad`` -- the bf16 bucket views of a data-parallel all-reduce (``dist.SegmentedStep``)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, shadow: bool = True):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("FusedAdam: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.shadow = shadow

    def _state(self, p):
        st = self.state[p]
        if not st:
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            if self.shadow and p.dim() >= 2:
                p._stswin_shadow = (p.detach().to(torch.bfloat16), p._version)
        return st

    @torch.no_grad()
    def step(self, closure=None, grads=None, grad_scale: float = 1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        gi = 0
        for group in self.param_groups:
            if grads is None:
                params = [p for p in group["params"] if p.grad is not None]
                gl = [p.grad for p in params]
            else:
                params = [p for p in group["params"] if p.requires_grad]
                gl = list(grads[gi:gi + len(params)])
                gi += len(params)
            if not params:
                continue
            _check_f32(params, "FusedAdam.step")
            gdt = gl[0].dtype
            if gdt not in (torch.float32, torch.bfloat16) or any(g.dtype != gdt or not g.is_contiguous() or not g.is_cuda for g in gl):
                raise StswinError("FusedAdam.step: gradients must be contiguous CUDA tensors, all fp32 or all bf16")
            states = [self._state(p) for p in params]
            if "step_t" not in group:
                group["step_t"] = torch.zeros((), dtype=torch.float32, device=params[0].device)
            shadows = []
            for p in params:
                sh = getattr(p, "_stswin_shadow", None)
                if sh is not None and sh[1] != p._version:       # stale shadow: the kernel rewrites it anyway
                    p._stswin_shadow = sh = (sh[0], p._version)
                shadows.append(sh[0].data_ptr() if sh is not None else None)
            n_bytes = sum(p.numel() for p in params) * (28 + gl[0].element_size() - 4 + 2)
            b1, b2 = group["betas"]
            with ops._launch("adam", float(n_bytes), params[0]):
                st = lib.stswin_adam_step(_ptrs(params), _ptrs(gl), _ptrs([s["exp_avg"] for s in states]),
                                          _ptrs([s["exp_avg_sq"] for s in states]),
                                          (ctypes.c_void_p * len(shadows))(*shadows), _numels(params), len(params),
                                          int(gdt == torch.bfloat16), float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                          float(group["weight_decay"]), float(grad_scale), group["step_t"].data_ptr(),
                                          ops._stream(params[0]))
            _lib.check(st, "stswin_adam_step")
            ops.count_extra_launches((len(params) + 31) // 32)
        return loss
