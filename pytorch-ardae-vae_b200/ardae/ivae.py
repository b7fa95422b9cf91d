"""`ToyIPVAE` / `MNISTIPVAE` -- drop-ins for the reference's implicit-posterior VAEs with the
noise-concat MLP encoder (models/ivae/toy.py, models/ivae/mnist.py; enc_type='concat', the only
encoder type ivae_ardae.py's configs 1-3 build).

Same constructor kwargs, attribute / state_dict names, method signatures and return tuples
(SURVEY.md 8b); the arithmetic is the libardae CUDA plan (csrc/model.cuh).  Autograd contract: the
`z` and `loss` returned by `forward` are attached to the graph (ivae_ardae.py calls `.backward` on
both, :804 and :834); `encode` / `forward_hidden` return plain tensors (every call site in the
reference step detaches them).
"""
import ctypes
import math
import weakref

import torch
import torch.nn as nn

from . import _lib
from .arena import ParamArena
from .layers import MLP, ContextConcatMLP


def normal_energy_func(x, mu=0., logvar=0.):
    """utils/energy.py:69-77 (kept for API parity: model.energy_func)."""
    x = x.view(x.size(0), -1)
    return torch.sum(0.5 * (logvar + (x - mu) ** 2 / math.exp(logvar) + math.log(2. * math.pi)), dim=1)


class Identity(nn.Module):
    def forward(self, input):
        return input


class NormalDistributionLinear(nn.Module):
    """models/reparam.py:62-71 (parameter container + sampler)."""

    def __init__(self, input_size, output_size, nonlinearity=None):
        super().__init__()
        if nonlinearity is not None:
            raise NotImplementedError('logvar clipping is not used by ToyIPVAE')
        self.input_size, self.output_size, self.nonlinearity = input_size, output_size, nonlinearity
        self.mean_fn = nn.Linear(input_size, output_size)
        self.logvar_fn = nn.Linear(input_size, output_size)

    def sample_gaussian(self, mu, logvar):
        return mu + torch.exp(0.5 * logvar) * torch.randn_like(logvar)


class BernoulliDistributionLinear(nn.Module):
    """models/reparam.py:163-174 (parameter container + relaxed-Bernoulli sampler :106-120)."""

    def __init__(self, input_size, output_size, hard=False):
        super().__init__()
        self.input_size, self.output_size, self.hard = input_size, output_size, hard
        self.logit_fn = nn.Linear(input_size, output_size)

    def sample_logistic_sigmoid(self, logits, temperature=1.0, hard=False):
        noise = torch.rand_like(logits)
        y = logits + torch.log(torch.div(noise, 1. - noise) + 1e-20)
        return torch.sigmoid(y / temperature)


class BernoulliDistributionConvTranspose2d(nn.Module):
    """models/reparam.py:191-202 (parameter container + the relaxed-Bernoulli sampler :106-120)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, output_padding=0, bias=True,
                 hard=False):
        super().__init__()
        self.hard = hard
        self.logit_fn = nn.ConvTranspose2d(in_channels, out_channels, kernel_size, stride, padding, output_padding,
                                           bias=bias)

    sample_logistic_sigmoid = BernoulliDistributionLinear.sample_logistic_sigmoid


def _weight_init(m):  # models/ivae/mnist.py:20-25
    if isinstance(m, (nn.Conv2d, nn.Linear)):
        nn.init.xavier_uniform_(m.weight)
        if m.bias is not None:
            nn.init.zeros_(m.bias)


class ConcatEncoder(nn.Module):
    """models/ivae/toy.py:154-194 (kind='toy') / models/ivae/mnist.py:123-165 (kind='mnist')."""

    def __init__(self, kind, input_dim, noise_dim, h_dim, z_dim, nonlinearity, num_hidden_layers, std=1.,
                 init='gaussian'):
        super().__init__()
        self.kind = kind
        self.input_dim, self.noise_dim, self.h_dim, self.z_dim = input_dim, noise_dim, h_dim, z_dim
        self.nonlinearity, self.num_hidden_layers, self.std, self.init = nonlinearity, num_hidden_layers, std, init
        self.enc_noise = False
        if kind == 'toy':
            self.inp_encode = MLP(input_dim, h_dim, h_dim, nonlinearity, num_hidden_layers - 1, True)
            self.nos_encode = Identity()
            self.fc = ContextConcatMLP(h_dim, noise_dim, h_dim, z_dim, nonlinearity, num_hidden_layers, False)
        else:
            self.inp_encode = MLP(input_dim, h_dim, h_dim, nonlinearity, num_hidden_layers, True)
            self.nos_encode = Identity()
            self.fc = MLP(h_dim + noise_dim, h_dim, z_dim, nonlinearity, 1, False)
        if init == 'gaussian':
            self.reset_parameters()

    def reset_parameters(self):
        nn.init.normal_(self.fc.fc.weight)

    def sample_noise(self, batch_size, std=None, device=None):
        """ivae/toy.py:61-65 draws on the CPU generator and copies; here the draw is on-device."""
        std = std if std is not None else self.std
        device = device if device is not None else next(self.parameters()).device
        if std == 0:
            return torch.zeros(batch_size, self.noise_dim, device=device)
        return std * torch.randn(batch_size, self.noise_dim, device=device)

    def forward(self, x, noise=None, std=None, nz=1):
        owner = self._owner()
        batch_size = x.size(0)
        if noise is None:
            zero = std is not None and std == 0
            noise = None if zero else self.sample_noise(batch_size * nz, std=std, device=x.device)
        else:
            assert noise.size(0) == batch_size * nz
            assert noise.size(1) == self.noise_dim
        return owner._encode(x, noise, nz)

    def _forward_inp(self, x):
        """ivae/mnist.py:76-86 / ivae/toy.py:67-75: features of the data rows, [B, h_dim]."""
        return self._owner()._forward_inp(x)

    def _forward_nos(self, batch_size=None, noise=None, std=None, device=None):
        """ivae/mnist.py:88-97: nos_encode is the identity for enc_type 'concat'."""
        assert batch_size is not None or noise is not None
        if noise is None:
            noise = self.sample_noise(batch_size, std=std, device=device)
        return noise

    def _forward_all(self, inp, nos):
        """ivae/mnist.py:161-165 / ivae/toy.py:192-194: z = fc([inp, nos]) on already expanded rows."""
        return self._owner()._forward_all(inp, nos)


class Decoder(nn.Module):
    def __init__(self, kind, input_dim, h_dim, z_dim, nonlinearity, num_hidden_layers, init='gaussian'):
        super().__init__()
        self.kind = kind
        self.input_dim, self.h_dim, self.z_dim = input_dim, h_dim, z_dim
        self.nonlinearity, self.num_hidden_layers, self.init = nonlinearity, num_hidden_layers, init
        if kind == 'toy':  # models/ivae/toy.py:711-720
            self.main = MLP(z_dim, h_dim, h_dim, nonlinearity, num_hidden_layers - 1, True)
            self.reparam = NormalDistributionLinear(h_dim, input_dim)
            if init == 'gaussian':
                nn.init.normal_(self.reparam.mean_fn.weight)
        else:  # models/ivae/mnist.py:180-181
            self.main = MLP(z_dim, h_dim, h_dim, nonlinearity, num_hidden_layers, True)
            self.reparam = BernoulliDistributionLinear(h_dim, input_dim)

    def sample(self, *heads):
        if self.kind == 'toy':
            return self.reparam.sample_gaussian(*heads)
        return self.reparam.sample_logistic_sigmoid(*heads)

    def forward(self, z):
        """ivae/toy.py:725-737 -> (x, mu, logvar); ivae/mnist.py:188-199 -> (x, logit)."""
        heads = self._owner()._decode_heads(z)
        return (self.sample(*heads),) + tuple(heads)


class _ForwardFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, key, z, sums, *params):
        ctx.model, ctx.key = model, key
        ctx.gen = model._plan_gen.get(key, 0)
        ctx.set_materialize_grads(False)
        return z, sums[0].clone()

    @staticmethod
    def backward(ctx, gz, gloss):
        m = ctx.model
        if ctx.key not in m._plans or m._plan_gen.get(ctx.key, 0) != ctx.gen:
            # the activations of this forward lived in the plan workspace and a later forward() overwrote them
            raise RuntimeError('ImplicitPosteriorVAE: forward() was called again with the same (batch, nz) before this '
                               'backward(); the plan keeps one set of activations -- call backward() first')
        h = m._plans[ctx.key][0]
        ar = m._arena
        L = _lib.lib()
        if gloss is not None:
            ar.stage_flat.zero_()
            _lib.check(L.ardae_model_backward(h, ctypes.c_float(1.0), None, ctypes.c_float(0.0), _lib.stream_ptr()))
            ar.accumulate_staged(gloss)
        if gz is not None:
            ar.stage_flat.zero_()
            g = _lib.require_cuda(gz, 'grad_z')
            _lib.check(L.ardae_model_backward(h, ctypes.c_float(0.0), _lib.ptr(g), ctypes.c_float(1.0),
                                              _lib.stream_ptr()))
            # only encoder tensors receive this pull-back; decoder slots of the stage arena stay zero
            ar.accumulate_staged(1.0)
        return (None, None, None, None) + (None,) * len(ar.params)


class ImplicitPosteriorVAE(nn.Module):
    KIND = None

    def __init__(self, energy_func=normal_energy_func, input_dim=2, noise_dim=2, h_dim=64, z_dim=2,
                 nonlinearity='tanh', num_hidden_layers=1, init='gaussian', enc_type='concat'):
        super().__init__()
        if enc_type != 'concat':
            raise NotImplementedError("the B200 path implements enc_type='concat' (what ivae_ardae.py builds)")
        if nonlinearity not in ('relu', 'softplus'):
            raise NotImplementedError("nonlinearity must be 'relu' or 'softplus'")
        if energy_func is not normal_energy_func:
            raise NotImplementedError('only the N(0,I) prior energy is fused')
        self.energy_func = energy_func
        self.input_dim, self.noise_dim, self.h_dim, self.z_dim = input_dim, noise_dim, h_dim, z_dim
        self.latent_dim = z_dim
        self.nonlinearity, self.num_hidden_layers, self.init, self.enc_type = nonlinearity, num_hidden_layers, init, enc_type
        kind = self.KIND
        if kind == 'toy':
            self.encode = ConcatEncoder(kind, input_dim, noise_dim, h_dim, z_dim, nonlinearity, num_hidden_layers, init=init)
            self.decode = Decoder(kind, input_dim, h_dim, z_dim, nonlinearity, num_hidden_layers, init=init)
            self._n_inp, self._n_fc, self._n_dec = num_hidden_layers, num_hidden_layers, num_hidden_layers
        else:
            self.encode = ConcatEncoder(kind, input_dim, noise_dim, h_dim, z_dim, nonlinearity, num_hidden_layers + 1, init=init)
            self.decode = Decoder(kind, input_dim, h_dim, z_dim, nonlinearity, num_hidden_layers)
            self.decode.apply(_weight_init)  # mnist.py:233-238
            if init == 'gaussian':
                self.encode.reset_parameters()
            self._n_inp, self._n_fc, self._n_dec = num_hidden_layers + 2, 1, num_hidden_layers + 1
        object.__setattr__(self.encode, '_owner', weakref.ref(self))
        object.__setattr__(self.decode, '_owner', weakref.ref(self))
        self._arena = ParamArena(self)
        self._plans = {}
        self.inv_rows_override = None  # data parallel: 1 / global row count

    # ------------------------------------------------------------------ plumbing
    def _ensure(self):
        if not self._arena.device_ok():
            self._arena.ensure()
            self._drop_plans()
        return self._arena

    @property
    def _plan_gen(self):
        """forward() generation per training plan: backward() of a stale forward raises instead of using the
        activations of a newer one."""
        g = self.__dict__.get('_plan_gen_')
        if g is None:
            g = self.__dict__['_plan_gen_'] = {}
        return g

    def _drop_plans(self):
        for _, (h, _ws) in self._plans.items():
            _lib.lib().ardae_model_destroy(h)
        self._plans = {}

    def __del__(self):
        try:
            self._drop_plans()
        except Exception:
            pass

    def _plan(self, B, nz, mode, slot=0):
        key = (B, nz, mode, slot)
        if key not in self._plans:
            L = _lib.lib()
            ar = self._ensure()
            cfg = _lib.ModelConfig({'toy': 0, 'mnist': 1, 'conv': 2, 'auxmnist': 3}[self.KIND], self.input_dim, self.noise_dim,
                                   self.h_dim, self.z_dim, self._n_inp, self._n_fc, self._n_dec,
                                   1 if self.nonlinearity == 'softplus' else 0, B, nz, mode,
                                   getattr(self, 'input_height', 0), getattr(self, 'input_channels', 0))
            nbytes = ctypes.c_size_t(0)
            _lib.check(L.ardae_model_workspace_bytes(ctypes.byref(cfg), ctypes.byref(nbytes)))
            ws = torch.empty(nbytes.value + 256, dtype=torch.uint8, device=ar.flat.device)
            off = (-ws.data_ptr()) % 256
            params = _lib.ptr_array(ar.views(ar.flat))
            grads = _lib.ptr_array(ar.views(ar.stage_flat))
            h = ctypes.c_void_p(0)
            _lib.check(L.ardae_model_create(ctypes.byref(cfg), params, grads, len(ar.params),
                                            ctypes.c_void_p(ws.data_ptr() + off), nbytes.value, ctypes.byref(h)))
            self._plans[key] = (h, ws)
        return key

    def _encode(self, x, noise, nz, slot=0):
        B = x.size(0)
        xf = _lib.require_cuda(x.detach(), 'input').view(B, self.input_dim)
        nf = None if noise is None else _lib.require_cuda(noise.detach(), 'noise')
        key = self._plan(B, nz, 0, slot)
        z = torch.empty(B * nz, self.z_dim, dtype=torch.float32, device=xf.device)
        _lib.check(_lib.lib().ardae_model_encode(self._plans[key][0], _lib.ptr(xf), _lib.ptr(nf), _lib.ptr(z),
                                                 _lib.stream_ptr()))
        return z.view(B, nz, self.z_dim)

    def _encode_with_mean(self, x, noise, nz):
        """(z [B,nz,d], zbar [B,1,d]) = (encode(x, noise, nz), encode(x, std=0)) from ONE pass over the input stack."""
        B = x.size(0)
        xf = _lib.require_cuda(x.detach(), 'input').view(B, self.input_dim)
        nf = _lib.require_cuda(noise.detach(), 'noise')
        key = self._plan(B, nz, 0)
        z = torch.empty(B * nz, self.z_dim, dtype=torch.float32, device=xf.device)
        zbar = torch.empty(B, self.z_dim, dtype=torch.float32, device=xf.device)
        _lib.check(_lib.lib().ardae_model_encode_with_mean(self._plans[key][0], _lib.ptr(xf), _lib.ptr(nf), _lib.ptr(z),
                                                           _lib.ptr(zbar), _lib.stream_ptr()))
        return z.view(B, nz, self.z_dim), zbar.view(B, 1, self.z_dim)

    @property
    def _noise_width(self):
        """columns of the `noise` operand of the plans: noise_dim; the hierarchical models pack (eps0 | eps)."""
        return self.noise_dim + (self.z_dim if self.KIND == 'auxmnist' else 0)

    def _encode_hidden(self, x, noise, nz, slot=0):
        """auxmnist: (z, zbar, hidden) = (encode(x, noise, nz), encode(x, std=0), encode.forward_hidden(x, std=0))."""
        B = x.size(0)
        xf = _lib.require_cuda(x.detach(), 'input').view(B, self.input_dim)
        nf = None if noise is None else _lib.require_cuda(noise.detach(), 'noise')
        key = self._plan(B, nz, 0, slot)
        z = torch.empty(B * nz, self.z_dim, dtype=torch.float32, device=xf.device)
        zbar = torch.empty(B, self.z_dim, dtype=torch.float32, device=xf.device)
        hid = torch.empty(B, 2 * self.h_dim, dtype=torch.float32, device=xf.device)
        _lib.check(_lib.lib().ardae_model_encode_hidden(self._plans[key][0], _lib.ptr(xf), _lib.ptr(nf), _lib.ptr(z),
                                                        _lib.ptr(zbar), _lib.ptr(hid), _lib.stream_ptr()))
        return z.view(B, nz, self.z_dim), zbar.view(B, 1, self.z_dim), hid

    def _feat_dim(self):
        return self.encode.s_h8 * self.encode.s_h8 * 32 if self.KIND == 'conv' else self.h_dim

    def _decode_heads(self, z):
        """Decoder heads for given latents (mode-3 plan): toy -> (mu, logvar), mnist / conv -> (logit,)."""
        R = z.size(0)
        zf = _lib.require_cuda(z.detach(), 'z').reshape(R, self.z_dim)
        self._ensure()
        key = self._plan(R, 1, 3)
        nH = 2 if self.KIND == 'toy' else 1
        heads = torch.empty(nH, R, self.input_dim, dtype=torch.float32, device=zf.device)
        _lib.check(_lib.lib().ardae_model_decode(self._plans[key][0], _lib.ptr(zf), _lib.ptr(heads), _lib.stream_ptr()))
        if self.KIND == 'conv':
            return (heads[0].view(R, self.input_channels, self.input_height, self.input_height),)
        return tuple(heads[k] for k in range(nH))

    def _forward_inp(self, x):
        B = x.size(0)
        xf = _lib.require_cuda(x.detach(), 'input').reshape(B, self.input_dim)
        self._ensure()
        key = self._plan(B, 1, 4)
        out = torch.empty(B, self._feat_dim(), dtype=torch.float32, device=xf.device)
        _lib.check(_lib.lib().ardae_model_forward_inp(self._plans[key][0], _lib.ptr(xf), _lib.ptr(out), _lib.stream_ptr()))
        return out

    def _forward_all(self, inp, nos):
        R = inp.size(0)
        f = _lib.require_cuda(inp.detach(), 'inp').reshape(R, self._feat_dim())
        nf = None if nos is None else _lib.require_cuda(nos.detach(), 'nos').reshape(R, self.noise_dim)
        self._ensure()
        key = self._plan(R, 1, 5)
        z = torch.empty(R, self.z_dim, dtype=torch.float32, device=f.device)
        _lib.check(_lib.lib().ardae_model_forward_all(self._plans[key][0], _lib.ptr(f), _lib.ptr(nf), _lib.ptr(z),
                                                      _lib.stream_ptr()))
        return z

    # ------------------------------------------------------------------ reference API
    def generate(self, batch_size=1):
        """ivae/toy.py:860-873 -> (x, mu, z); ivae/mnist.py:303-316 / ivae/conv.py:232-245 -> (x, sigmoid(logit), z)."""
        dev = next(self.parameters()).device
        z = torch.randn(batch_size, self.z_dim, device=dev)
        out = self.decode(z)
        if self.KIND == 'toy':
            return out[0], out[1], z
        return out[0], torch.sigmoid(out[1]), z

    def forward_hidden(self, input, std=None, nz=1):
        """toy.py:811-822 / mnist.py:254-265."""
        batch_size = input.size(0)
        input = input.view(batch_size, self.input_dim)
        eps = self.encode.sample_noise(batch_size * nz, std=std, device=input.device)
        return self.encode(input, noise=eps, std=std, nz=nz)

    def forward(self, input, beta=1.0, eta=0.0, lmbd=0.0, std=None, nz=1, noise=None):
        """toy.py:824-858 / mnist.py:267-301.  `noise` (optional [B*nz, n]) injects the encoder noise."""
        if lmbd > 0:
            raise NotImplementedError  # same as the reference (toy.py:845-846)
        batch_size = input.size(0)
        x = _lib.require_cuda(input.detach(), 'input').view(batch_size, self.input_dim)
        if noise is None:
            noise = self.encode.sample_noise(batch_size * nz, std=std, device=x.device)
        nf = _lib.require_cuda(noise.detach(), 'noise')
        ar = self._ensure()
        key = self._plan(batch_size, nz, 1)
        R = batch_size * nz
        z = torch.empty(R, self.z_dim, dtype=torch.float32, device=x.device)
        sums = torch.empty(3, dtype=torch.float32, device=x.device)
        nH = 2 if self.KIND == 'toy' else 1
        heads = torch.empty(nH, R, self.input_dim, dtype=torch.float32, device=x.device)
        inv_rows = self.inv_rows_override if self.inv_rows_override is not None else 1.0 / R
        self._plan_gen[key] = self._plan_gen.get(key, 0) + 1
        _lib.check(_lib.lib().ardae_model_forward(self._plans[key][0], _lib.ptr(x), _lib.ptr(nf),
                                                  ctypes.c_float(beta), ctypes.c_float(inv_rows), _lib.ptr(z),
                                                  _lib.ptr(sums), _lib.ptr(heads), _lib.stream_ptr()))
        z3, loss = _ForwardFn.apply(self, key, z.view(batch_size, nz, self.z_dim), sums, *ar.params)
        if self.KIND == 'toy':
            mu, logvar = heads[0], heads[1]
            xhat = self.decode.reparam.sample_gaussian(mu, logvar)
            mean = mu
        else:
            logit = heads[0]
            xhat = self.decode.reparam.sample_logistic_sigmoid(logit)
            mean = torch.sigmoid(logit)
        return xhat, mean, z3, loss, sums[1].detach(), sums[2].detach()

    def logprob(self, input, sample_size=128, z=None, std=None, noise=None, eta=None, return_per_image=False):
        """toy.py:875-939 / mnist.py:318-319,378-437 (`logprob_w_cov_gaussian_posterior`): importance-weighted
        log-likelihood with the moment-matched full-covariance Gaussian proposal, batched over the images.
        `noise` [b, S, n] / `eta` [b, S, z] inject the encoder noise and the MVN.rsample normal draw."""
        batch_size = input.size(0)
        assert sample_size >= 2 * self.z_dim  # mnist.py:382
        x = _lib.require_cuda(input.detach(), 'input').view(batch_size, self.input_dim)
        S = int(sample_size)
        if noise is None:
            noise = self.encode.sample_noise(batch_size * S, std=std, device=x.device)
        nf = _lib.require_cuda(noise.detach(), 'noise').reshape(batch_size * S, self._noise_width)
        ef = None if eta is None else _lib.require_cuda(eta.detach(), 'eta').reshape(batch_size * S, self.z_dim)
        self._ensure()
        key = self._plan(batch_size, S, 2)
        out = torch.empty(batch_size, dtype=torch.float32, device=x.device)
        status = torch.zeros(1, dtype=torch.int32, device=x.device)
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if ef is None else 0
        _lib.check(_lib.lib().ardae_model_iws(self._plans[key][0], _lib.ptr(x), _lib.ptr(nf), _lib.ptr(ef),
                                              ctypes.c_uint64(seed), _lib.ptr(out), None, _lib.ptr(status),
                                              _lib.stream_ptr()))
        self.last_iws_status = status
        if return_per_image:
            return out
        return out.mean()


class ToyIPVAE(ImplicitPosteriorVAE):
    """net.ToyIPVAE (models/__init__.py; models/ivae/toy.py:739-1024)."""
    KIND = 'toy'


class MNISTIPVAE(ImplicitPosteriorVAE):
    """net.MNISTIPVAE (models/ivae/mnist.py:201-518)."""
    KIND = 'mnist'

    def __init__(self, energy_func=normal_energy_func, input_dim=784, noise_dim=100, h_dim=300, z_dim=32,
                 nonlinearity='softplus', num_hidden_layers=1, init='gaussian', enc_type='concat'):
        super().__init__(energy_func, input_dim, noise_dim, h_dim, z_dim, nonlinearity, num_hidden_layers, init,
                         enc_type)


class _AuxStage(nn.Module):
    """Parameter container of one Gaussian stage of the hierarchical encoder: `AuxEncoder` (models/vae/auxmnist.py:
    31-53: main + reparam) or `SimpleEncoder` (:147-183: identities + fc + reparam)."""

    def __init__(self, name, in_dim, h_dim, out_dim, nonlinearity, num_hidden_layers, simple):
        super().__init__()
        if simple:
            self.inp_encode, self.nos_encode = Identity(), Identity()
        self.add_module(name, MLP(in_dim, h_dim, h_dim, nonlinearity, num_hidden_layers - 1, True))
        self.reparam = NormalDistributionLinear(h_dim, out_dim)


class _AuxEncoder(nn.Module):
    """models/ivae/auxmnist.py:47-131 `Encoder`: z0 ~ q(z0|x) (aux_encode), z ~ q(z|x, z0) (encode)."""

    def __init__(self, input_dim, noise_dim, h_dim, z_dim, nonlinearity, num_hidden_layers, enc_type):
        super().__init__()
        self.input_dim, self.noise_dim, self.h_dim, self.z_dim = input_dim, noise_dim, h_dim, z_dim
        self.nonlinearity, self.num_hidden_layers, self.enc_type = nonlinearity, num_hidden_layers, enc_type
        self.clip_z0_logvar = self.clip_z_logvar = None
        self.aux_encode = _AuxStage('main', input_dim, h_dim, noise_dim, nonlinearity, num_hidden_layers, False)
        self.encode = _AuxStage('fc', input_dim + noise_dim, h_dim, z_dim, nonlinearity, num_hidden_layers, True)

    def sample_noise(self, batch_size, std=None, device=None):
        """The two draws of Encoder._forward (:107-113) packed as [rows, noise_dim + z_dim] = (eps0 | eps), already
        scaled by `std` (sample_gaussian multiplies both standard deviations by it, :33-40)."""
        std = std if std is not None else 1
        device = device if device is not None else next(self.parameters()).device
        if std == 0:
            return torch.zeros(batch_size, self.noise_dim + self.z_dim, device=device)
        return std * torch.randn(batch_size, self.noise_dim + self.z_dim, device=device)

    def forward(self, x, std=None, nz=1, noise=None):
        """(:115-120) -> z [B, nz, z_dim].  `noise` (optional [B*nz, noise_dim + z_dim]) injects (eps0 | eps)."""
        batch_size = x.size(0)
        if noise is None:
            zero = std is not None and std == 0
            noise = None if zero else self.sample_noise(batch_size * nz, std=std, device=x.device)
        else:
            assert noise.size(0) == batch_size * nz and noise.size(1) == self.noise_dim + self.z_dim
        return self._owner()._encode(x, noise, nz)

    def forward_hidden(self, x, std=None, nz=1):
        """(:122-131) -> cat(h0, h) [B, 2 h_dim].  Only the deterministic call the training step makes
        (std = 0, ivae_ardae.py:739-741) is planned."""
        assert nz == 1
        if std is None or std != 0:
            raise NotImplementedError('forward_hidden is planned for std=0 (the hidden1a context)')
        return self._owner()._encode_hidden(x, None, 1)[2]


class MNISTAuxIPVAE(ImplicitPosteriorVAE):
    """net.MNISTAuxIPVAE (models/ivae/auxmnist.py:133-360): hierarchical implicit posterior
    q(z|x) = int q(z|x, z0) q(z0|x) dz0 with two Gaussian stages, Bernoulli MLP decoder (models/vae/mnist.py)."""
    KIND = 'auxmnist'

    def __init__(self, energy_func=normal_energy_func, input_dim=784, noise_dim=100, h_dim=300, z_dim=32,
                 nonlinearity='softplus', num_hidden_layers=2, enc_type='simple', clip_z0_logvar=None,
                 clip_z_logvar=None, do_xavier=True):
        nn.Module.__init__(self)
        if enc_type != 'simple':
            raise NotImplementedError  # same as the reference (:164)
        if nonlinearity != 'softplus':
            raise NotImplementedError("the auxmnist plan is softplus only (what ivae_ardae.py builds)")
        if energy_func is not normal_energy_func:
            raise NotImplementedError('only the N(0,I) prior energy is fused')
        if clip_z0_logvar not in (None, 'none') or clip_z_logvar not in (None, 'none'):
            raise NotImplementedError('logvar clipping is not planned (ivae_ardae.py default: none)')
        if num_hidden_layers < 1:
            raise ValueError('num_hidden_layers must be >= 1')
        self.energy_func = energy_func
        self.input_dim, self.noise_dim, self.h_dim, self.z_dim = input_dim, noise_dim, h_dim, z_dim
        self.latent_dim = z_dim
        self.nonlinearity, self.num_hidden_layers, self.enc_type = nonlinearity, num_hidden_layers, enc_type
        self.clip_z0_logvar = self.clip_z_logvar = None
        self.do_xavier = do_xavier
        self.encode = _AuxEncoder(input_dim, noise_dim, h_dim, z_dim, nonlinearity, num_hidden_layers, enc_type)
        self.decode = Decoder('mnist', input_dim, h_dim, z_dim, nonlinearity, num_hidden_layers - 1)
        if do_xavier:
            self.apply(_weight_init)
        self._n_inp = self._n_fc = self._n_dec = num_hidden_layers
        object.__setattr__(self.encode, '_owner', weakref.ref(self))
        object.__setattr__(self.decode, '_owner', weakref.ref(self))
        self._arena = ParamArena(self)
        self._plans = {}
        self.inv_rows_override = None

    def forward_hidden(self, input, std=None, nz=1):
        """(:236-249) -> z [B, nz, z_dim]."""
        return self.encode(input.view(input.size(0), self.input_dim), std=std, nz=nz)


class _ConvEncoder(nn.Module):
    """models/ivae/conv.py:44-136 (parameter container; attribute names = state_dict keys)."""

    def __init__(self, input_height, input_channels, noise_dim, z_dim, nonlinearity):
        super().__init__()
        self.input_height, self.input_channels = input_height, input_channels
        self.noise_dim, self.z_dim, self.nonlinearity, self.enc_noise = noise_dim, z_dim, nonlinearity, False
        co = lambda hin: int((hin + 2 * 2 - 1 * (5 - 1) - 1) / 2 + 1)  # utils/msc.py:43-45
        s_h8 = co(co(co(input_height)))
        self.conv1 = nn.Conv2d(input_channels, 16, 5, 2, 2, bias=True)
        self.conv2 = nn.Conv2d(16, 32, 5, 2, 2, bias=True)
        self.conv3 = nn.Conv2d(32, 32, 5, 2, 2, bias=True)
        self.fc4 = nn.Linear(s_h8 * s_h8 * 32 + noise_dim, 800, bias=True)
        self.fc5 = nn.Linear(800, z_dim, bias=True)
        self.nos_encode = Identity()
        self.s_h8 = s_h8

    def sample_noise(self, batch_size, std=None, device=None):
        std = std if std is not None else 1
        device = device if device is not None else next(self.parameters()).device
        if std == 0:
            return torch.zeros(batch_size, self.noise_dim, device=device)
        return std * torch.randn(batch_size, self.noise_dim, device=device)

    def forward(self, x, noise=None, std=None, nz=1):
        owner = self._owner()
        batch_size = x.size(0)
        if noise is None:
            zero = std is not None and std == 0
            noise = None if zero else self.sample_noise(batch_size * nz, std=std, device=x.device)
        else:
            assert noise.size(0) == batch_size * nz
            assert noise.size(1) == self.noise_dim
        return owner._encode(x, noise, nz)

    def _forward_inp(self, x):
        """ivae/conv.py:84-96: flattened conv3 features, [B, 32 * s_h8^2]."""
        return self._owner()._forward_inp(x)

    _forward_nos = ConcatEncoder._forward_nos

    def _forward_all(self, inp, nos):
        """ivae/conv.py:107-115."""
        return self._owner()._forward_all(inp, nos)


class _ConvDecoder(nn.Module):
    """models/vae/conv.py:79-136 (parameter container)."""

    def __init__(self, input_height, input_channels, z_dim, nonlinearity, s_h8):
        super().__init__()
        self.input_height, self.input_channels, self.z_dim, self.nonlinearity = input_height, input_channels, z_dim, nonlinearity
        self.s_h8 = s_h8
        self.fc = MLP(z_dim, 300, s_h8 * s_h8 * 32, nonlinearity, 1, True)
        self.deconv1 = nn.ConvTranspose2d(32, 32, 5, 2, 2, 0, bias=True)
        self.deconv2 = nn.ConvTranspose2d(32, 16, 5, 2, 2, 0, bias=True)
        self.reparam = BernoulliDistributionConvTranspose2d(16, input_channels, 5, 2, 2, 0, bias=True)

    def sample(self, logit):
        return self.reparam.sample_logistic_sigmoid(logit)

    def forward(self, z):
        """vae/conv.py:118-136 -> (x, logit), both [rows, C, H, W]."""
        (logit,) = self._owner()._decode_heads(z)
        return self.sample(logit), logit


class ConvIPVAE(ImplicitPosteriorVAE):
    """net.ConvIPVAE (models/ivae/conv.py:138-304): 3x(conv 5x5 s2) + noise-concat fc encoder, fc + 3x deconv
    Bernoulli decoder.  The conv / deconv layers run as direct fp32 CUDA kernels on the data rows; the
    per-sample layers (fc4, fc5) and the decoder fc run on the tcgen05 GEMMs."""
    KIND = 'conv'

    def __init__(self, energy_func=normal_energy_func, input_height=28, input_channels=1, z_dim=32, noise_dim=100,
                 nonlinearity='softplus', do_xavier=True):
        nn.Module.__init__(self)
        if nonlinearity not in ('relu', 'softplus'):
            raise NotImplementedError("nonlinearity must be 'relu' or 'softplus'")
        if energy_func is not normal_energy_func:
            raise NotImplementedError('only the N(0,I) prior energy is fused')
        self.energy_func = energy_func
        self.input_height, self.input_channels = input_height, input_channels
        self.input_dim = input_channels * input_height * input_height
        self.z_dim, self.latent_dim, self.noise_dim = z_dim, z_dim, noise_dim
        self.nonlinearity, self.do_xavier = nonlinearity, do_xavier
        self.h_dim = 800
        self.encode = _ConvEncoder(input_height, input_channels, noise_dim, z_dim, nonlinearity)
        self.decode = _ConvDecoder(input_height, input_channels, z_dim, nonlinearity, self.encode.s_h8)
        if do_xavier:  # models/vae/auxconv.py:18-23: Conv2d and Linear only (not ConvTranspose2d)
            self.apply(_weight_init)
        self._n_inp, self._n_fc, self._n_dec = 3, 1, 2
        object.__setattr__(self.encode, '_owner', weakref.ref(self))
        object.__setattr__(self.decode, '_owner', weakref.ref(self))
        self._arena = ParamArena(self)
        self._plans = {}
        self.inv_rows_override = None

    def forward(self, input, beta=1.0, eta=0.0, lmbd=0.0, std=None, nz=1, noise=None):
        out = super().forward(input.reshape(input.size(0), -1), beta, eta, lmbd, std, nz, noise)
        shp = (-1, self.input_channels, self.input_height, self.input_height)
        return (out[0].view(shp), out[1].view(shp)) + out[2:]
