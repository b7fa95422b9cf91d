"""Optimizers of the AR-DAE step as ONE vectorised pass over the flat parameter arena.

`Adam`    -- the reference's own utils.Adam (utils/optim.py:9-108: PyTorch-1.2 epsilon placement,
             denom = (sqrt(v)+eps)/sqrt(1-b2^t), step = lr/(1-b1^t)), constructed at ivae_ardae.py:550.
`RMSprop` -- torch.optim.RMSprop(params, lr, momentum) semantics (alpha=0.99, eps=1e-8,
             centered=False), constructed at ivae_ardae.py:626.
Same constructor signatures, `zero_grad()/step()/state_dict()/param_groups`.  Parameters whose
`.grad` is None are not updated (reference behaviour: `neglogprob.fc.bias`).
28 bytes/parameter/step: read p,g,s1,s2 - write p,s1,s2.
"""
import ctypes

import torch
from torch.optim.optimizer import Optimizer

from . import _lib


def _find_arena(params):
    arenas = {id(getattr(p, '_ardae_arena', None)): getattr(p, '_ardae_arena', None) for p in params}
    if None in arenas.values() or len(arenas) != 1:
        return None
    return list(arenas.values())[0]


class _FlatOptimizer(Optimizer):
    STATE_NAMES = ()

    def _setup(self):
        params = [p for g in self.param_groups for p in g['params']]
        if len(self.param_groups) != 1:
            raise RuntimeError('ardae optimizers take a single parameter group')
        owner = None
        for p in params:
            owner = getattr(p, '_ardae_owner', None)
            if owner is None:
                raise RuntimeError('parameters must belong to an ardae module (ToyIPVAE / MNISTIPVAE / MLPGradCARDAE)')
        mod = owner()
        ar = mod._ensure()
        if len(params) != len(ar.params) or any(a is not b for a, b in zip(params, ar.params)):
            raise RuntimeError('pass module.parameters() of exactly one ardae module')
        if getattr(self, '_ar', None) is not ar or getattr(self, '_flat_id', None) != ar.flat.data_ptr():
            self._ar, self._flat_id = ar, ar.flat.data_ptr()
            old = getattr(self, '_bufs', None)
            self._bufs = [torch.zeros_like(ar.flat) for _ in self.STATE_NAMES]
            for k, p in enumerate(ar.params):
                st = self.state[p]
                for name, buf in zip(self.STATE_NAMES, self._bufs):
                    v = ar.view(buf, k)
                    if name in st and torch.is_tensor(st[name]):
                        v.copy_(st[name])
                    st[name] = v
                st.setdefault('step', 0)
            del old
        return ar

    def zero_grad(self, set_to_none=False):
        ar = self._setup()
        ar.grad_flat.zero_()
        if set_to_none:
            for p in ar.params:
                p.grad = None

    def _prepare_grads(self, ar):
        """Make the grad arena hold exactly the gradients to apply; returns skipped indices."""
        skipped = []
        base = ar.grad_flat.data_ptr()
        for k, (p, o) in enumerate(zip(ar.params, ar.offsets)):
            g = ar.view(ar.grad_flat, k)
            if p.grad is None:
                g.zero_()
                skipped.append(k)
            elif p.grad.data_ptr() != base + 4 * o:
                g.copy_(p.grad)
        return skipped

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._ar = None  # re-alias the loaded state tensors into flat buffers on the next step
        self._state_sharded = False

    def state_dict(self):
        # fused data-parallel updates advance each slice of the state on its owner rank only (ardae/dp.py)
        if getattr(self, '_state_sharded', False):
            raise RuntimeError('optimizer state is sharded over the data-parallel ranks: call '
                               'TrainStep.gather_optimizer_state() on every rank before state_dict()')
        return super().state_dict()


class Adam(_FlatOptimizer):
    STATE_NAMES = ('exp_avg', 'exp_avg_sq')

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if weight_decay != 0 or amsgrad:
            raise NotImplementedError('weight_decay / amsgrad are not used by the AR-DAE configs')
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad))
        self.grad_scale = 1.0

    @torch.no_grad()
    def step(self, closure=None):
        ar = self._setup()
        group = self.param_groups[0]
        skipped = self._prepare_grads(ar)
        t = None
        for k, p in enumerate(ar.params):
            if k in skipped:
                continue
            self.state[p]['step'] += 1
            t = self.state[p]['step']
        if t is None:
            return None
        b1, b2 = group['betas']
        m, v = self._bufs
        _lib.check(_lib.lib().ardae_adam_step(_lib.ptr(ar.flat), _lib.ptr(ar.grad_flat), _lib.ptr(m), _lib.ptr(v),
                                              ctypes.c_size_t(ar.total), ctypes.c_float(group['lr']),
                                              ctypes.c_float(b1), ctypes.c_float(b2), ctypes.c_float(group['eps']),
                                              int(t), ctypes.c_float(self.grad_scale), _lib.stream_ptr()))
        return None


    @torch.no_grad()
    def step_flat(self, grad_flat, skip=()):
        """Fused-driver entry: apply `grad_flat` (arena layout) without touching `.grad`."""
        ar = self._setup()
        group = self.param_groups[0]
        t = None
        for k, p in enumerate(ar.params):
            if k in skip:
                continue
            self.state[p]['step'] += 1
            t = self.state[p]['step']
        b1, b2 = group['betas']
        m, v = self._bufs
        _lib.check(_lib.lib().ardae_adam_step(_lib.ptr(ar.flat), _lib.ptr(grad_flat), _lib.ptr(m), _lib.ptr(v),
                                              ctypes.c_size_t(ar.total), ctypes.c_float(group['lr']),
                                              ctypes.c_float(b1), ctypes.c_float(b2), ctypes.c_float(group['eps']),
                                              int(t), ctypes.c_float(self.grad_scale), _lib.stream_ptr()))


    @torch.no_grad()
    def step_flat_dp(self, grad_flat, comm, skip=()):
        """Data-parallel fused-driver entry: gradient exchange over peer memory + update in ONE kernel (ardae/dp.py)."""
        ar = self._setup()
        group = self.param_groups[0]
        t = None
        for k, p in enumerate(ar.params):
            if k in skip:
                continue
            self.state[p]['step'] += 1
            t = self.state[p]['step']
        b1, b2 = group['betas']
        m, v = self._bufs
        comm.fused_step(0, ar.flat, grad_flat, m, v, ar.total, group['lr'], b1, b2, group['eps'], 0.0, int(t),
                        self.grad_scale)
        self._state_sharded = True


class RMSprop(_FlatOptimizer):
    STATE_NAMES = ('square_avg', 'momentum_buffer')

    def __init__(self, params, lr=1e-2, alpha=0.99, eps=1e-8, weight_decay=0, momentum=0, centered=False):
        if weight_decay != 0 or centered:
            raise NotImplementedError('weight_decay / centered are not used by the AR-DAE configs')
        super().__init__(params, dict(lr=lr, alpha=alpha, eps=eps, weight_decay=weight_decay, momentum=momentum,
                                      centered=centered))
        self.grad_scale = 1.0

    @torch.no_grad()
    def step(self, closure=None):
        ar = self._setup()
        group = self.param_groups[0]
        skipped = self._prepare_grads(ar)
        for k, p in enumerate(ar.params):
            if k not in skipped:
                self.state[p]['step'] += 1
        sq, buf = self._bufs
        _lib.check(_lib.lib().ardae_rmsprop_step(_lib.ptr(ar.flat), _lib.ptr(ar.grad_flat), _lib.ptr(sq),
                                                 _lib.ptr(buf), ctypes.c_size_t(ar.total), ctypes.c_float(group['lr']),
                                                 ctypes.c_float(group['alpha']), ctypes.c_float(group['eps']),
                                                 ctypes.c_float(group['momentum']), ctypes.c_float(self.grad_scale),
                                                 _lib.stream_ptr()))
        return None

    @torch.no_grad()
    def step_flat(self, grad_flat, skip=()):
        ar = self._setup()
        group = self.param_groups[0]
        for k, p in enumerate(ar.params):
            if k not in skip:
                self.state[p]['step'] += 1
        sq, buf = self._bufs
        _lib.check(_lib.lib().ardae_rmsprop_step(_lib.ptr(ar.flat), _lib.ptr(grad_flat), _lib.ptr(sq),
                                                 _lib.ptr(buf), ctypes.c_size_t(ar.total), ctypes.c_float(group['lr']),
                                                 ctypes.c_float(group['alpha']), ctypes.c_float(group['eps']),
                                                 ctypes.c_float(group['momentum']), ctypes.c_float(self.grad_scale),
                                                 _lib.stream_ptr()))

    @torch.no_grad()
    def step_flat_dp(self, grad_flat, comm, skip=()):
        ar = self._setup()
        group = self.param_groups[0]
        for k, p in enumerate(ar.params):
            if k not in skip:
                self.state[p]['step'] += 1
        sq, buf = self._bufs
        comm.fused_step(1, ar.flat, grad_flat, sq, buf, ar.total, group['lr'], 0.0, group['alpha'], group['eps'],
                        group['momentum'], 1, self.grad_scale)
        self._state_sharded = True
