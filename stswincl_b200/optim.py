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

__all__ = ["LARS", "add_weight_decay", "momentum_update"]


def add_weight_decay(model, weight_decay=1e-5, skip_list=()):
    """Split parameters into a no-decay group (1-D tensors: biases, norms; flagged ``ignore`` for LARS) and a
    decay group (``lars.py:7-32``)."""
    decay, no_decay = [], []
    for name, param in model.named_parameters():
        if not param.requires_grad:
            continue
        if len(param.shape) == 1 or name in skip_list:
            no_decay.append(param)
        else:
            decay.append(param)
    return [{'params': no_decay, 'weight_decay': 0, 'ignore': True},
            {'params': decay, 'weight_decay': weight_decay, 'ignore': False}]


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
