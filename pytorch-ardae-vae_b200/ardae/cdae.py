"""`MLPGradCARDAE` -- drop-in for the reference's net.MLPGradCARDAE
(= models/graddae/mlp.py:341-483 `ConditionalARDAE`, registered at models/__init__.py).

Same constructor, method names, return conventions and state_dict keys; the arithmetic is the
libardae CUDA plan (csrc/cdae.cuh): tcgen05 GEMM chain with fused softplus / sigmoid / residual-loss /
tangent / adjoint epilogues.  `forward` returns `(None, loss)` with `loss` attached to autograd:
`loss.backward()` accumulates the hand-derived double-backprop gradient into `.grad` of the 6L+1
parameters that receive one (`neglogprob.fc.bias` never does -- reference behaviour).
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib
from .arena import ParamArena
from .layers import MLP


class _TrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, loss_dev, *params):
        ctx.module = module
        ctx.gen = module._stage_gen
        return loss_dev.clone().reshape(())

    @staticmethod
    def backward(ctx, gloss):
        m = ctx.module
        if ctx.gen != m._stage_gen:
            # the gradients of this forward lived in the staging arena and a later forward() overwrote them
            raise RuntimeError('ConditionalARDAE: forward() was called again before this loss.backward(); the fused '
                               'plan keeps one set of staged gradients -- call backward() before the next forward()')
        m._arena.accumulate_staged(gloss, skip=m.no_grad_params)
        return (None, None) + (None,) * len(m._arena.params)


class ConditionalARDAE(nn.Module):
    KIND = 0           # ardae_cdae_config.kind: 0 = mlp-grad (energy network), 1 = mlp-res (score network)
    HEAD = 'neglogprob'  # attribute / state_dict prefix of the last MLP

    def __init__(self, input_dim=2, h_dim=128, context_dim=2, std=0.01, num_hidden_layers=1,
                 nonlinearity='tanh', noise_type='gaussian', enc_input=True, enc_ctx=True,
                 std_method='default'):
        super().__init__()
        if nonlinearity != 'softplus' or noise_type != 'gaussian' or not enc_input or not enc_ctx:
            raise NotImplementedError('the B200 path implements the configuration ivae_ardae.py:595-606 builds: '
                                      'softplus, gaussian noise, enc_input = enc_ctx = True')
        if num_hidden_layers < 2:
            raise NotImplementedError('num_hidden_layers >= 2 required')
        self.input_dim, self.h_dim, self.context_dim = input_dim, h_dim, context_dim
        self.std, self.num_hidden_layers = std, num_hidden_layers
        self.nonlinearity, self.noise_type = nonlinearity, noise_type
        self.enc_input, self.enc_ctx = enc_input, enc_ctx
        # models/graddae/mlp.py:374-378
        self.ctx_encode = MLP(context_dim, h_dim, h_dim, nonlinearity=nonlinearity,
                              num_hidden_layers=num_hidden_layers - 1, use_nonlinearity_output=True)
        self.inp_encode = MLP(input_dim, h_dim, h_dim, nonlinearity=nonlinearity,
                              num_hidden_layers=num_hidden_layers - 1, use_nonlinearity_output=True)
        # graddae: neglogprob -> scalar energy (:378); resdae: dae -> the score itself, input_dim wide (resdae/mlp.py:323)
        setattr(self, self.HEAD, MLP(2 * h_dim + 1, h_dim, 1 if self.KIND == 0 else input_dim, nonlinearity=nonlinearity,
                                     num_hidden_layers=num_hidden_layers, use_nonlinearity_output=False))
        self._arena = ParamArena(self)
        self._plans = {}
        self.inv_count_override = None  # data parallel: 1 / (global N * d)
        self.last_score = None
        self._stage_gen = 0  # bumped by every training forward (see _TrainFn.backward)

    @property
    def no_grad_params(self):
        """Indices (arena order) of parameters that never receive a gradient: the energy network's output bias
        (`neglogprob.fc.bias`, reference: .grad stays None); none for the residual CDAE."""
        return (len(self._arena.params) - 1,) if self.KIND == 0 else ()

    def reset_parameters(self):  # exists in the reference (graddae/mlp.py:380-382); never called there
        nn.init.normal_(getattr(self, self.HEAD).fc.weight)
        if self.KIND == 0:
            for m in self.inp_encode.linears():
                m.weight.data.mul_(0.001)

    # ------------------------------------------------------------------ plumbing
    def _ensure(self):
        if not self._arena.device_ok():
            self._arena.ensure()
            for _, (h, _ws) in self._plans.items():
                _lib.lib().ardae_cdae_destroy(h)
            self._plans = {}
        return self._arena

    def _plan(self, B, S, train):
        key = (B, S, train)
        if key not in self._plans:
            L = _lib.lib()
            ar = self._ensure()
            cfg = _lib.CdaeConfig(self.input_dim, self.context_dim, self.h_dim, self.num_hidden_layers, B, S,
                                  1 if train else 0, self.KIND)
            nbytes = ctypes.c_size_t(0)
            _lib.check(L.ardae_cdae_workspace_bytes(ctypes.byref(cfg), ctypes.byref(nbytes)))
            ws = torch.empty(nbytes.value + 256, dtype=torch.uint8, device=ar.flat.device)
            off = (-ws.data_ptr()) % 256
            params = _lib.ptr_array(ar.views(ar.flat))
            grads = _lib.ptr_array(ar.views(ar.stage_flat))
            h = ctypes.c_void_p(0)
            _lib.check(L.ardae_cdae_create(ctypes.byref(cfg), params, grads, len(ar.params),
                                           ctypes.c_void_p(ws.data_ptr() + off), nbytes.value, ctypes.byref(h)))
            self._plans[key] = (h, ws)
        return self._plans[key][0]

    def __del__(self):
        try:
            for _, (h, _ws) in self._plans.items():
                _lib.lib().ardae_cdae_destroy(h)
        except Exception:
            pass

    def add_noise(self, input, std=None):
        std = self.std if std is None else std
        eps = torch.randn_like(input)
        return input + std * eps, eps

    # ------------------------------------------------------------------ reference API
    def forward(self, input, context, std=None, scale=None, eps=None, seed=None):
        """graddae/mlp.py:400-444.  `eps` (optional, [B,S,d]) injects the Gaussian noise for parity
        runs; otherwise it is drawn in-kernel (Philox) -- the reference draws torch.randn_like."""
        assert input.dim() == 3   # bsz x ssz x x_dim
        assert context.dim() == 3  # bsz x 1 x ctx_dim
        B, S, d = input.shape
        assert d == self.input_dim and context.size(0) == B
        x = _lib.require_cuda(input.detach(), 'input').view(B * S, d)
        ctxt = _lib.require_cuda(context.detach(), 'context').view(B, -1)
        if std is None:
            sig = x.new_zeros(B * S)
        else:
            assert torch.is_tensor(std)
            sig = _lib.require_cuda(std.detach(), 'std').reshape(B * S)
        ar = self._ensure()
        h = self._plan(B, S, True)
        gen = eps is None
        if gen:
            epsb = torch.empty(B * S, d, dtype=torch.float32, device=x.device)
            if seed is None:
                seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        else:
            epsb = _lib.require_cuda(eps.detach(), 'eps').reshape(B * S, d).clone()
            seed = 0
        ar.stage_flat.zero_()
        self._stage_gen += 1
        loss_dev = torch.empty(1, dtype=torch.float32, device=x.device)
        score = torch.empty(B * S, d, dtype=torch.float32, device=x.device)
        inv = self.inv_count_override if self.inv_count_override is not None else 1.0 / float(B * S * d)
        _lib.check(_lib.lib().ardae_cdae_train(h, _lib.ptr(x), _lib.ptr(ctxt), _lib.ptr(sig), _lib.ptr(epsb),
                                               1 if gen else 0, ctypes.c_uint64(seed), ctypes.c_float(inv),
                                               _lib.ptr(loss_dev), _lib.ptr(score), _lib.stream_ptr()))
        self.last_score = score.view(B, S, d)
        self.last_eps = epsb.view(B, S, d)
        loss = _TrainFn.apply(self, loss_dev, *ar.params)
        return None, loss

    def glogprob(self, input, context, std=None, scale=None):
        """graddae/mlp.py:446-483."""
        assert input.dim() == 3
        assert context.dim() == 3
        B, S, d = input.shape
        x = _lib.require_cuda(input.detach(), 'input').view(B * S, d)
        ctxt = _lib.require_cuda(context.detach(), 'context').view(B, -1)
        if std is None:
            sig = x.new_zeros(B * S)
        else:
            assert torch.is_tensor(std)
            sig = _lib.require_cuda(std.detach(), 'std').reshape(B * S)
        self._ensure()
        h = self._plan(B, S, False)
        out = torch.empty(B * S, d, dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().ardae_cdae_score(h, _lib.ptr(x), _lib.ptr(ctxt), _lib.ptr(sig), _lib.ptr(out),
                                               _lib.stream_ptr()))
        return out.view(B, S, d)


MLPGradCARDAE = ConditionalARDAE


class ResidualConditionalARDAE(ConditionalARDAE):
    """Drop-in for net.MLPResCARDAE (= models/resdae/mlp.py:286-413 `ConditionalARDAE`, `--cdae mlp-res`,
    ivae_ardae.py:583-594): the network `dae([inp_encode(x~), ctx_encode(c), sigma])` outputs the score estimate
    directly; loss = mse(sigma * f, -eps); plain back-propagation (every parameter receives a gradient).  Same
    kernels as the energy variant: 3xTF32 forward chain, MUL_SIG backward chain with fused bias-gradient column
    sums, weight-gradient contractions."""
    KIND = 1
    HEAD = 'dae'


MLPResCARDAE = ResidualConditionalARDAE
