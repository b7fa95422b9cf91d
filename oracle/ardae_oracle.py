"""CPU oracle for the AR-DAE hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a plain-numpy restatement (explicit formulas, explicit backward passes, no autograd,
no torch) of the reference's AR-DAE training step and IWS evaluator.  It exists to CHECK the CUDA
path; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product package never imports anything under ``oracle/``.

Parity status: PINNED.  The reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the reference's own modules executed in the build container:
``oracle/make_golden.py`` imports ``/root/reference`` (stubbing its plotting imports), runs the step
body of ``ivae_ardae.py:707-846`` and ``evaluate_iws`` (``:644-673``) with injected noise, and writes
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays those fixtures through this file.

Reference lines each function follows are cited in its docstring (paths relative to the reference
repo root).  Parameter dicts use the reference's ``state_dict`` key names; weights are ``[out, in]``.
All arithmetic runs in the dtype of the arrays passed in (tests use float64).
"""
import math

import numpy as np

LOG2PI = math.log(2.0 * math.pi)


# ----------------------------------------------------------------------------- activations
def softplus(x):
    """F.softplus (beta=1, threshold=20): utils/models.py:14-32 -> torch."""
    return np.where(x > 20.0, x, np.log1p(np.exp(np.minimum(x, 20.0))))


def sigmoid(x):
    return 0.5 * (1.0 + np.tanh(0.5 * x))


def act_fwd(name, a):
    if name == 'softplus':
        return softplus(a)
    if name == 'relu':
        return np.maximum(a, 0.0)
    raise NotImplementedError(name)


def act_grad(name, a):
    """d act / d a evaluated at pre-activation a."""
    if name == 'softplus':
        return np.where(a > 20.0, 1.0, sigmoid(a))
    if name == 'relu':
        return (a > 0.0).astype(a.dtype)
    raise NotImplementedError(name)


# ----------------------------------------------------------------------------- MLP helpers
def mlp_keys(prefix, num_hidden_layers):
    """models/layers.py:477-499: `layers.{i}` then `fc`."""
    return ['%s.layers.%d' % (prefix, i) for i in range(num_hidden_layers)] + ['%s.fc' % prefix]


def mlp_forward(P, keys, x, nonlin, out_nonlin):
    """models/layers.py:501-515.  Returns output and the list of (input, pre-activation) per layer."""
    tape = []
    h = x
    for i, k in enumerate(keys):
        a = h @ P[k + '.weight'].T + P[k + '.bias']
        tape.append((h, a))
        last = i == len(keys) - 1
        h = act_fwd(nonlin, a) if (not last or out_nonlin) else a
    return h, tape


def mlp_backward(P, keys, tape, dout, nonlin, out_nonlin, G):
    """Reverse of mlp_forward; accumulates parameter grads into G, returns d input."""
    d = dout
    for i in reversed(range(len(keys))):
        k = keys[i]
        h_in, a = tape[i]
        last = i == len(keys) - 1
        da = d * act_grad(nonlin, a) if (not last or out_nonlin) else d
        G[k + '.weight'] = G.get(k + '.weight', 0.0) + da.T @ h_in
        G[k + '.bias'] = G.get(k + '.bias', 0.0) + da.sum(0)
        d = da @ P[k + '.weight']
    return d


# ----------------------------------------------------------------------------- 5x5 stride-2 pad-2 convolutions
def conv5s2(x, W, b):
    """nn.Conv2d(Ci, Co, 5, 2, 2) (models/ivae/conv.py:70-72): x [B,Ci,H,H], W [Co,Ci,5,5] -> [B,Co,Ho,Ho]."""
    B, Ci, H, _ = x.shape
    Ho = (H + 4 - 5) // 2 + 1
    xp = np.pad(x, ((0, 0), (0, 0), (2, 2), (2, 2)))
    out = np.zeros((B, W.shape[0], Ho, Ho), dtype=x.dtype)
    for kh in range(5):
        for kw in range(5):
            out += np.einsum('bchw,oc->bohw', xp[:, :, kh:kh + 2 * Ho:2, kw:kw + 2 * Ho:2], W[:, :, kh, kw])
    return out + (b[None, :, None, None] if b is not None else 0.0)


def conv5s2_backward(x, W, dout):
    """Returns (dx, dW, db) of conv5s2."""
    B, Ci, H, _ = x.shape
    Ho = dout.shape[2]
    xp = np.pad(x, ((0, 0), (0, 0), (2, 2), (2, 2)))
    dxp = np.zeros_like(xp)
    dW = np.zeros_like(W)
    for kh in range(5):
        for kw in range(5):
            patch = xp[:, :, kh:kh + 2 * Ho:2, kw:kw + 2 * Ho:2]
            dW[:, :, kh, kw] = np.einsum('bohw,bchw->oc', dout, patch)
            dxp[:, :, kh:kh + 2 * Ho:2, kw:kw + 2 * Ho:2] += np.einsum('bohw,oc->bchw', dout, W[:, :, kh, kw])
    return dxp[:, :, 2:2 + H, 2:2 + H], dW, dout.sum((0, 2, 3))


def deconv5s2(x, W, b):
    """nn.ConvTranspose2d(Ci, Co, 5, 2, 2, 0) (models/vae/conv.py:112-115): x [B,Ci,H,H], W [Ci,Co,5,5]
    -> [B,Co,2H-1,2H-1]; it is the adjoint of conv5s2 with the same weight memory."""
    B, Ci, H, _ = x.shape
    Hout = 2 * H - 1
    canvas = np.zeros((B, W.shape[1], Hout + 4, Hout + 4), dtype=x.dtype)
    for kh in range(5):
        for kw in range(5):
            canvas[:, :, kh:kh + 2 * H:2, kw:kw + 2 * H:2] += np.einsum('bchw,co->bohw', x, W[:, :, kh, kw])
    return canvas[:, :, 2:2 + Hout, 2:2 + Hout] + b[None, :, None, None]


def deconv5s2_backward(x, W, dout):
    B, Ci, H, _ = x.shape
    dp = np.pad(dout, ((0, 0), (0, 0), (2, 2), (2, 2)))
    dx = np.zeros_like(x)
    dW = np.zeros_like(W)
    for kh in range(5):
        for kw in range(5):
            patch = dp[:, :, kh:kh + 2 * H:2, kw:kw + 2 * H:2]
            dx += np.einsum('bohw,co->bchw', patch, W[:, :, kh, kw])
            dW[:, :, kh, kw] = np.einsum('bchw,bohw->co', x, patch)
    return dx, dW, dout.sum((0, 2, 3))


# ----------------------------------------------------------------------------- implicit encoders
class ModelSpec(object):
    """Architecture of an ImplicitPosteriorVAE as built by ivae_ardae.py:295-314."""

    def __init__(self, kind, input_dim, noise_dim, h_dim, z_dim, num_hidden_layers, nonlin):
        assert kind in ('toy', 'mnist', 'conv', 'auxmnist')
        self.kind, self.input_dim, self.noise_dim = kind, input_dim, noise_dim
        self.h_dim, self.z_dim, self.n_layers, self.nonlin = h_dim, z_dim, num_hidden_layers, nonlin
        if kind == 'toy':
            # models/ivae/toy.py:176-179,711-712
            self.inp_keys = mlp_keys('encode.inp_encode', num_hidden_layers - 1)
            self.fc_keys = mlp_keys('encode.fc', num_hidden_layers)
            self.dec_keys = mlp_keys('decode.main', num_hidden_layers - 1)
        elif kind == 'conv':
            # models/ivae/conv.py:44-136, models/vae/conv.py:79-136; input_dim = C*H*H, h_dim = 800
            self.dec_keys = mlp_keys('decode.fc', 1)
        elif kind == 'auxmnist':
            # hierarchical encoder models/ivae/auxmnist.py:47-131: AuxEncoder (models/vae/auxmnist.py:31-68) +
            # SimpleEncoder (:147-191); decoder models/vae/mnist.py Decoder (MLP of num_hidden_layers linears + logits)
            self.aux_keys = mlp_keys('encode.aux_encode.main', num_hidden_layers - 1)
            self.fc_keys = mlp_keys('encode.encode.fc', num_hidden_layers - 1)
            self.dec_keys = mlp_keys('decode.main', num_hidden_layers - 1)
        else:
            # models/ivae/mnist.py:148-151,180-181,227 (encoder gets num_hidden_layers+1)
            self.inp_keys = mlp_keys('encode.inp_encode', num_hidden_layers + 1)
            self.fc_keys = mlp_keys('encode.fc', 1)
            self.dec_keys = mlp_keys('decode.main', num_hidden_layers)


def encoder_forward(spec, P, x, eps, nz):
    """z = f(x, eps).  toy: models/ivae/toy.py:67-109,192-194 + layers.py:707-724 (noise re-concatenated
    at every layer); mnist: models/ivae/mnist.py:76-121,161-165 (x <- 2x-1, one concat).
    x [B, D], eps [B*nz, n] (row b*nz+k) -> z [B, nz, d] and a tape for the backward pass."""
    B = x.shape[0]
    if spec.kind == 'conv':
        return _conv_encoder_forward(spec, P, x, eps, nz)
    if spec.kind == 'auxmnist':
        return _aux_encoder_forward(spec, P, x, eps, nz)
    xin = 2.0 * x - 1.0 if spec.kind == 'mnist' else x
    inp, tape_inp = mlp_forward(P, spec.inp_keys, xin, spec.nonlin, True)
    inp_rep = np.repeat(inp, nz, axis=0)  # row b*nz+k  (unsqueeze(1).expand(-1,nz,-1))
    tape_fc = []
    if spec.kind == 'toy':
        hid = inp_rep
        for i, k in enumerate(spec.fc_keys):
            cat = np.concatenate([hid, eps], axis=1)
            a = cat @ P[k + '.weight'].T + P[k + '.bias']
            tape_fc.append((cat, a))
            hid = act_fwd(spec.nonlin, a) if i < len(spec.fc_keys) - 1 else a
        z = hid
    else:
        cat = np.concatenate([inp_rep, eps], axis=1)
        z, tape_fc = mlp_forward(P, spec.fc_keys, cat, spec.nonlin, False)
    return z.reshape(B, nz, spec.z_dim), (tape_inp, tape_fc, xin)


def encoder_backward(spec, P, tape, dz, nz, G):
    """Backward of encoder_forward for upstream dz [B, nz, d]; accumulates grads into G."""
    if spec.kind == 'conv':
        return _conv_encoder_backward(spec, P, tape, dz, nz, G)
    if spec.kind == 'auxmnist':
        return _aux_encoder_backward(spec, P, tape, dz, nz, G)
    tape_inp, tape_fc, xin = tape
    B = xin.shape[0]
    H = spec.h_dim
    d = dz.reshape(B * nz, spec.z_dim)
    if spec.kind == 'toy':
        for i in reversed(range(len(spec.fc_keys))):
            k = spec.fc_keys[i]
            cat, a = tape_fc[i]
            da = d * act_grad(spec.nonlin, a) if i < len(spec.fc_keys) - 1 else d
            G[k + '.weight'] = G.get(k + '.weight', 0.0) + da.T @ cat
            G[k + '.bias'] = G.get(k + '.bias', 0.0) + da.sum(0)
            d = (da @ P[k + '.weight'])[:, :H]
        dinp_rep = d
    else:
        dcat = mlp_backward(P, spec.fc_keys, tape_fc, d, spec.nonlin, False, G)
        dinp_rep = dcat[:, :H]
    dinp = dinp_rep.reshape(B, nz, H).sum(1)
    mlp_backward(P, spec.inp_keys, tape_inp, dinp, spec.nonlin, True, G)


def _aux_encoder_forward(spec, P, x, eps, nz):
    """Hierarchical encoder (models/ivae/auxmnist.py:74-105).  eps [B*nz, n + d] packs the two draws of
    Encoder._forward: eps0 (columns [0, n), the reparametrisation noise of the auxiliary latent z0) and eps (columns
    [n, n + d), that of z); eps = 0 is `std = 0` (z0 = mu0, z = mu).
      h0 = MLP(2x - 1) ; mu0, lv0 = heads(h0) ; z0 = mu0 + exp(lv0 / 2) * eps0        (AuxEncoder, vae/auxmnist.py:55-68)
      h = MLP([2x - 1, z0]) ; mu, lv = heads(h) ; z = mu + exp(lv / 2) * eps          (SimpleEncoder, :185-190)"""
    B, n, d = x.shape[0], spec.noise_dim, spec.z_dim
    xin = 2.0 * x - 1.0
    h0, tape_aux = mlp_forward(P, spec.aux_keys, xin, spec.nonlin, True)
    mu0 = h0 @ P['encode.aux_encode.reparam.mean_fn.weight'].T + P['encode.aux_encode.reparam.mean_fn.bias']
    lv0 = h0 @ P['encode.aux_encode.reparam.logvar_fn.weight'].T + P['encode.aux_encode.reparam.logvar_fn.bias']
    eps0, eps1 = eps[:, :n], eps[:, n:n + d]
    mu0r, lv0r = np.repeat(mu0, nz, axis=0), np.repeat(lv0, nz, axis=0)
    z0 = mu0r + np.exp(0.5 * lv0r) * eps0
    cat = np.concatenate([np.repeat(xin, nz, axis=0), z0], axis=1)
    h, tape_fc = mlp_forward(P, spec.fc_keys, cat, spec.nonlin, True)
    mu = h @ P['encode.encode.reparam.mean_fn.weight'].T + P['encode.encode.reparam.mean_fn.bias']
    lv = h @ P['encode.encode.reparam.logvar_fn.weight'].T + P['encode.encode.reparam.logvar_fn.bias']
    z = mu + np.exp(0.5 * lv) * eps1
    return z.reshape(B, nz, d), (tape_aux, tape_fc, xin, h0, lv0r, eps0, h, lv, eps1)


def aux_encoder_hidden(spec, P, x):
    """Encoder.forward_hidden(x, std=0) (models/ivae/auxmnist.py:123-131): cat(h0, h) at eps = 0, [B, 2 h_dim] --
    the CDAE context of cdae_ctx_type 'hidden1a' (ivae_ardae.py:739-741)."""
    B = x.shape[0]
    _, tape = _aux_encoder_forward(spec, P, x, np.zeros((B, spec.noise_dim + spec.z_dim), dtype=x.dtype), 1)
    return np.concatenate([tape[3], tape[6]], axis=1)


def _aux_encoder_backward(spec, P, tape, dz, nz, G):
    tape_aux, tape_fc, xin, h0, lv0r, eps0, h, lv, eps1 = tape
    B, D, n = xin.shape[0], xin.shape[1], spec.noise_dim
    d = dz.reshape(B * nz, spec.z_dim)
    dmu, dlv = d, d * (0.5 * np.exp(0.5 * lv) * eps1)
    dh = 0.0
    for nm, g in (('encode.encode.reparam.mean_fn', dmu), ('encode.encode.reparam.logvar_fn', dlv)):
        G[nm + '.weight'] = G.get(nm + '.weight', 0.0) + g.T @ h
        G[nm + '.bias'] = G.get(nm + '.bias', 0.0) + g.sum(0)
        dh = dh + g @ P[nm + '.weight']
    dcat = mlp_backward(P, spec.fc_keys, tape_fc, dh, spec.nonlin, True, G)
    dz0 = dcat[:, D:D + n]
    dmu0 = dz0.reshape(B, nz, n).sum(1)
    dlv0 = (dz0 * (0.5 * np.exp(0.5 * lv0r) * eps0)).reshape(B, nz, n).sum(1)
    dh0 = 0.0
    for nm, g in (('encode.aux_encode.reparam.mean_fn', dmu0), ('encode.aux_encode.reparam.logvar_fn', dlv0)):
        G[nm + '.weight'] = G.get(nm + '.weight', 0.0) + g.T @ h0
        G[nm + '.bias'] = G.get(nm + '.bias', 0.0) + g.sum(0)
        dh0 = dh0 + g @ P[nm + '.weight']
    mlp_backward(P, spec.aux_keys, tape_aux, dh0, spec.nonlin, True, G)


def _img_h(spec):
    c = getattr(spec, 'img_c', 1)
    return int(round(math.sqrt(spec.input_dim // c))), c


def _conv_encoder_forward(spec, P, x, eps, nz):
    """models/ivae/conv.py:84-136: x <- 2x-1; conv1..3 + act; flatten (NCHW); fc4([inp, eps]) + act; fc5."""
    B = x.shape[0]
    Himg, C = _img_h(spec)
    x4 = (2.0 * x - 1.0).reshape(B, C, Himg, Himg)
    maps, pre = [x4], []
    hcur = x4
    for k in ('encode.conv1', 'encode.conv2', 'encode.conv3'):
        a = conv5s2(hcur, P[k + '.weight'], P[k + '.bias'])
        pre.append(a)
        hcur = act_fwd(spec.nonlin, a)
        maps.append(hcur)
    inp = hcur.reshape(B, -1)
    cat = np.concatenate([np.repeat(inp, nz, axis=0), eps], axis=1)
    a4 = cat @ P['encode.fc4.weight'].T + P['encode.fc4.bias']
    h4 = act_fwd(spec.nonlin, a4)
    z = h4 @ P['encode.fc5.weight'].T + P['encode.fc5.bias']
    return z.reshape(B, nz, spec.z_dim), (maps, pre, cat, a4, h4)


def _conv_encoder_backward(spec, P, tape, dz, nz, G):
    maps, pre, cat, a4, h4 = tape
    B = maps[0].shape[0]
    d = dz.reshape(B * nz, spec.z_dim)
    G['encode.fc5.weight'] = G.get('encode.fc5.weight', 0.0) + d.T @ h4
    G['encode.fc5.bias'] = G.get('encode.fc5.bias', 0.0) + d.sum(0)
    da4 = (d @ P['encode.fc5.weight']) * act_grad(spec.nonlin, a4)
    G['encode.fc4.weight'] = G.get('encode.fc4.weight', 0.0) + da4.T @ cat
    G['encode.fc4.bias'] = G.get('encode.fc4.bias', 0.0) + da4.sum(0)
    feat = maps[3].reshape(B, -1).shape[1]
    dinp = (da4 @ P['encode.fc4.weight'])[:, :feat].reshape(B, nz, feat).sum(1)
    dh = dinp.reshape(maps[3].shape)
    for i, k in reversed(list(enumerate(('encode.conv1', 'encode.conv2', 'encode.conv3')))):
        da = dh * act_grad(spec.nonlin, pre[i])
        dh, dW, db = conv5s2_backward(maps[i], P[k + '.weight'], da)
        G[k + '.weight'] = G.get(k + '.weight', 0.0) + dW
        G[k + '.bias'] = G.get(k + '.bias', 0.0) + db


def _conv_decoder_head(spec, P, h, tape):
    """models/vae/conv.py:126-131: view [32,s8,s8] -> act(deconv1) -> ZeroPad2d((0,1,0,1)) -> act(deconv2)
    -> logit deconv -> crop last row / column."""
    R = h.shape[0]
    Himg, C = _img_h(spec)
    s8 = int(round(math.sqrt(h.shape[1] // 32)))
    h1 = h.reshape(R, 32, s8, s8)
    a2 = deconv5s2(h1, P['decode.deconv1.weight'], P['decode.deconv1.bias'])
    h2 = np.pad(act_fwd(spec.nonlin, a2), ((0, 0), (0, 0), (0, 1), (0, 1)))
    a3 = deconv5s2(h2, P['decode.deconv2.weight'], P['decode.deconv2.bias'])
    h3 = act_fwd(spec.nonlin, a3)
    lg = deconv5s2(h3, P['decode.reparam.logit_fn.weight'], P['decode.reparam.logit_fn.bias'])[:, :, :-1, :-1]
    assert lg.shape[2] == Himg
    return (lg.reshape(R, -1),), (h, (tape, h1, a2, h2, a3, h3))


def _conv_decoder_head_backward(spec, P, tape_all, dlogit, G):
    tape, h1, a2, h2, a3, h3 = tape_all
    R = h1.shape[0]
    Himg, C = _img_h(spec)
    dl = np.pad(dlogit.reshape(R, C, Himg, Himg), ((0, 0), (0, 0), (0, 1), (0, 1)))
    dh3, dW, db = deconv5s2_backward(h3, P['decode.reparam.logit_fn.weight'], dl)
    G['decode.reparam.logit_fn.weight'] = G.get('decode.reparam.logit_fn.weight', 0.0) + dW
    G['decode.reparam.logit_fn.bias'] = G.get('decode.reparam.logit_fn.bias', 0.0) + db
    dh2, dW, db = deconv5s2_backward(h2, P['decode.deconv2.weight'], dh3 * act_grad(spec.nonlin, a3))
    G['decode.deconv2.weight'] = G.get('decode.deconv2.weight', 0.0) + dW
    G['decode.deconv2.bias'] = G.get('decode.deconv2.bias', 0.0) + db
    da2 = dh2[:, :, :-1, :-1] * act_grad(spec.nonlin, a2)
    dh1, dW, db = deconv5s2_backward(h1, P['decode.deconv1.weight'], da2)
    G['decode.deconv1.weight'] = G.get('decode.deconv1.weight', 0.0) + dW
    G['decode.deconv1.bias'] = G.get('decode.deconv1.bias', 0.0) + db
    return dh1.reshape(R, -1)


# ----------------------------------------------------------------------------- decoders + ELBO terms
def decoder_forward(spec, P, z):
    """toy: models/ivae/toy.py:725-737 + reparam.py:55-58 (mu, logvar heads, no clipping);
    mnist: models/ivae/mnist.py:188-199 + reparam.py:170-172 (logits)."""
    h, tape = mlp_forward(P, spec.dec_keys, z, spec.nonlin, True)
    if spec.kind == 'conv':
        return _conv_decoder_head(spec, P, h, tape)
    if spec.kind == 'toy':
        mu = h @ P['decode.reparam.mean_fn.weight'].T + P['decode.reparam.mean_fn.bias']
        lv = h @ P['decode.reparam.logvar_fn.weight'].T + P['decode.reparam.logvar_fn.bias']
        return (mu, lv), (h, tape)
    logit = h @ P['decode.reparam.logit_fn.weight'].T + P['decode.reparam.logit_fn.bias']
    return (logit,), (h, tape)


def recon_rows(spec, heads, x):
    """Per-row -log p(x|z).  toy: utils/vae.py:36-52 (do_sum=False); mnist: utils/vae.py:21-30."""
    if spec.kind == 'toy':
        mu, lv = heads
        return 0.5 * (lv + (x - mu) ** 2 / np.exp(lv) + LOG2PI).sum(1)
    (logit,) = heads
    return (softplus(logit) - x * logit).sum(1)


def recon_rows_grad(spec, heads, x):
    if spec.kind == 'toy':
        mu, lv = heads
        return (-(x - mu) / np.exp(lv), 0.5 * (1.0 - (x - mu) ** 2 / np.exp(lv)))
    (logit,) = heads
    return (sigmoid(logit) - x,)


def prior_rows(z):
    """utils/energy.py:69-77: 0.5 * sum_d (z^2 + log 2pi)."""
    return 0.5 * (z ** 2 + LOG2PI).sum(1)


def model_forward(spec, P, x, eps, beta, nz=1):
    """ImplicitPosteriorVAE.forward: toy.py:824-858 / mnist.py:267-301 (lmbd = 0).
    Returns dict(z, loss, recon, prior, mean) and a tape."""
    B = x.shape[0]
    z3, tape_enc = encoder_forward(spec, P, x, eps, nz)
    z = z3.reshape(B * nz, spec.z_dim)
    heads, tape_dec = decoder_forward(spec, P, z)
    xrep = np.repeat(x, nz, axis=0)
    rec = recon_rows(spec, heads, xrep)
    pri = prior_rows(z)
    out = dict(z=z3, loss=(rec + beta * pri).mean(), recon=rec.mean(), prior=pri.mean(),
               mean=heads[0] if spec.kind == 'toy' else sigmoid(heads[0]))
    return out, (tape_enc, tape_dec, heads, z, xrep)


def model_backward(spec, P, tape, beta, nz, dz_extra, G, loss_scale=1.0):
    """d(loss_scale*loss)/d params plus an injected upstream gradient dz_extra [B,nz,d] on z
    (ivae_ardae.py:804 and :834 folded into one pass)."""
    tape_enc, tape_dec, heads, z, xrep = tape
    n = z.shape[0]
    h, tape_main = tape_dec
    hg = recon_rows_grad(spec, heads, xrep)
    if spec.kind == 'conv':
        dh = _conv_decoder_head_backward(spec, P, tape_main, hg[0] * (loss_scale / n), G)
        tape_main = tape_main[0]
        names, hg = [], []
    elif spec.kind == 'toy':
        names = ['decode.reparam.mean_fn', 'decode.reparam.logvar_fn']
    else:
        names = ['decode.reparam.logit_fn']
    if spec.kind != 'conv':
        dh = 0.0
    for nm, g in zip(names, hg):
        g = g * (loss_scale / n)
        G[nm + '.weight'] = G.get(nm + '.weight', 0.0) + g.T @ h
        G[nm + '.bias'] = G.get(nm + '.bias', 0.0) + g.sum(0)
        dh = dh + g @ P[nm + '.weight']
    dz = mlp_backward(P, spec.dec_keys, tape_main, dh, spec.nonlin, True, G)
    dz = dz + (loss_scale * beta / n) * z
    dz3 = dz.reshape(-1, nz, spec.z_dim)
    if dz_extra is not None:
        dz3 = dz3 + dz_extra
    encoder_backward(spec, P, tape_enc, dz3, nz, G)
    return dz3


# ----------------------------------------------------------------------------- CDAE (mlp-grad)
class CdaeSpec(object):
    """MLPGradCARDAE as built by ivae_ardae.py:595-606 (enc_ctx = enc_input = True, softplus)."""

    def __init__(self, input_dim, context_dim, h_dim, num_hidden_layers, kind='grad'):
        self.d, self.c, self.H, self.L = input_dim, context_dim, h_dim, num_hidden_layers
        self.kind = kind  # 'grad': MLPGradCARDAE (graddae/mlp.py), 'res': MLPResCARDAE (resdae/mlp.py:286-413)
        # models/graddae/mlp.py:374-378 ; models/resdae/mlp.py:319-323
        self.ctx_keys = mlp_keys('ctx_encode', num_hidden_layers - 1)
        self.inp_keys = mlp_keys('inp_encode', num_hidden_layers - 1)
        self.nlp_keys = mlp_keys('neglogprob' if kind == 'grad' else 'dae', num_hidden_layers)


# ----------------------------------------------------------------------------- residual CDAE (mlp-res)
def _rescdae_forward(cs, P, xt, ctx, std, sample_size):
    """resdae/mlp.py:373-383 / :403-411: f = dae([inp_encode(xt), ctx_encode(ctx), std]) -- the network outputs the
    score directly.  xt [N,d], ctx [B,c], std [N,1]."""
    uL, tape_inp = mlp_forward(P, cs.inp_keys, xt, 'softplus', True)
    cL, tape_ctx = mlp_forward(P, cs.ctx_keys, ctx, 'softplus', True)
    hcat = np.concatenate([uL, np.repeat(cL, sample_size, axis=0), std], axis=1)
    f, tape_dae = mlp_forward(P, cs.nlp_keys, hcat, 'softplus', False)
    return f, dict(tape_inp=tape_inp, tape_ctx=tape_ctx, tape_dae=tape_dae)


def rescdae_loss_and_grads(cs, P, x, ctx, std, eps):
    """ConditionalARDAE.forward + loss.backward() of the residual CDAE: resdae/mlp.py:353-389,
    loss = F.mse_loss(std * f, -eps) (:386); plain back-propagation (every parameter gets a gradient)."""
    B, S, d = x.shape
    N, H = B * S, cs.H
    xf, sf, ef = x.reshape(N, d), std.reshape(N, 1), eps.reshape(N, d)
    xt = xf + sf * ef  # add_gaussian_noise
    f, T = _rescdae_forward(cs, P, xt, ctx.reshape(B, -1), sf, S)
    resid = sf * f + ef
    loss = (resid ** 2).mean()
    r = (2.0 / (N * d)) * sf * resid
    G = {}
    dh = mlp_backward(P, cs.nlp_keys, T['tape_dae'], r, 'softplus', False, G)  # [N, 2H+1]
    mlp_backward(P, cs.inp_keys, T['tape_inp'], dh[:, :H], 'softplus', True, G)
    dc = dh[:, H:2 * H].reshape(B, S, H).sum(1)
    mlp_backward(P, cs.ctx_keys, T['tape_ctx'], dc, 'softplus', True, G)
    return loss, f.reshape(B, S, d), G


def _cdae_primal_and_score(cs, P, xt, ctx, std, sample_size):
    """Sweeps (1)-(2) of SURVEY.md 8a-3: graddae/mlp.py:426-437.  xt [N,d], ctx [B,c], std [N,1]."""
    H = cs.H
    uL, tape_inp = mlp_forward(P, cs.inp_keys, xt, 'softplus', True)
    cL, tape_ctx = mlp_forward(P, cs.ctx_keys, ctx, 'softplus', True)
    hcat = np.concatenate([uL, np.repeat(cL, sample_size, axis=0), std], axis=1)
    nk = cs.nlp_keys
    tape_nlp = []
    v = hcat
    for k in nk[:-1]:
        p = v @ P[k + '.weight'].T + P[k + '.bias']
        tape_nlp.append((v, p))
        v = softplus(p)
    wo = P[nk[-1] + '.weight']  # [1, H]
    # score sweep: delta's
    n = xt.shape[0]
    dv = np.broadcast_to(-wo, (n, H)).copy()
    dps = [None] * len(tape_nlp)
    for i in reversed(range(len(tape_nlp))):
        _, p = tape_nlp[i]
        dps[i] = dv * sigmoid(p)
        dv = dps[i] @ P[nk[i] + '.weight']
    du = dv[:, :H]
    das = [None] * len(tape_inp)
    for i in reversed(range(len(tape_inp))):
        _, a = tape_inp[i]
        das[i] = du * sigmoid(a)
        du = das[i] @ P[cs.inp_keys[i] + '.weight']
    g = du  # [N, d]  = d(-sum E)/d xt
    return g, dict(tape_inp=tape_inp, tape_ctx=tape_ctx, tape_nlp=tape_nlp, v_last=v, dps=dps,
                   das=das, cL=cL)


def cdae_glogprob(cs, P, x, ctx, std):
    """ConditionalARDAE.glogprob: graddae/mlp.py:446-483.  x [B,S,d], ctx [B,1,c], std [B,S,1]."""
    B, S, d = x.shape
    if cs.kind == 'res':  # resdae/mlp.py:391-413
        f, _ = _rescdae_forward(cs, P, x.reshape(B * S, d), ctx.reshape(B, -1), std.reshape(B * S, 1), S)
        return f.reshape(B, S, d)
    g, _ = _cdae_primal_and_score(cs, P, x.reshape(B * S, d), ctx.reshape(B, -1),
                                  std.reshape(B * S, 1), S)
    return g.reshape(B, S, d)


def cdae_loss_and_grads(cs, P, x, ctx, std, eps):
    """ConditionalARDAE.forward + loss.backward(): graddae/mlp.py:400-444 and ivae_ardae.py:771.
    The parameter gradient of the double-backprop loss is computed by hand with the
    tangent/adjoint sweeps (3)-(4) of SURVEY.md 8a-3.  Returns loss, score g [B,S,d], grads dict
    (no entry for neglogprob.fc.bias: it never receives a gradient)."""
    if cs.kind == 'res':
        return rescdae_loss_and_grads(cs, P, x, ctx, std, eps)
    B, S, d = x.shape
    N, H = B * S, cs.H
    xf, sf, ef = x.reshape(N, d), std.reshape(N, 1), eps.reshape(N, d)
    xt = xf + sf * ef  # add_gaussian_noise, graddae/mlp.py:21-23
    g, T = _cdae_primal_and_score(cs, P, xt, ctx.reshape(B, -1), sf, S)
    resid = sf * g + ef
    loss = (resid ** 2).mean()  # F.mse_loss(std*glogprob, -eps), graddae/mlp.py:395-398,441
    r = (2.0 / (N * d)) * sf * resid
    G = {}
    ik, nk = cs.inp_keys, cs.nlp_keys
    # (3) tangent forward
    adots, udots = [], []
    ud = r
    for i, k in enumerate(ik):
        _, a = T['tape_inp'][i]
        ad = ud @ P[k + '.weight'].T
        adots.append(ad)
        udots.append(ud)  # tangent of the layer INPUT
        ud = ad * sigmoid(a)
    pdots, vdots = [], []
    vd = np.concatenate([ud, np.zeros((N, H + 1), dtype=ud.dtype)], axis=1)
    for i, k in enumerate(nk[:-1]):
        _, p = T['tape_nlp'][i]
        pd = vd @ P[k + '.weight'].T
        pdots.append(pd)
        vdots.append(vd)
        vd = pd * sigmoid(p)
    G[nk[-1] + '.weight'] = -vd.sum(0, keepdims=True)
    # (4) adjoint backward
    wo = P[nk[-1] + '.weight']
    dv = np.broadcast_to(-wo, (N, H))
    adj_v = np.zeros((N, H), dtype=xt.dtype)
    for i in reversed(range(len(nk) - 1)):
        k = nk[i]
        v_in, p = T['tape_nlp'][i]
        s = sigmoid(p)
        # delta v_i (score-sweep value feeding layer i's sigmoid) = dps[i]/s, use the identity
        # dv_i * pdot * s' = dps[i] * pdot * (1 - s)
        adj_p = adj_v * s + T['dps'][i] * pdots[i] * (1.0 - s)
        G[k + '.weight'] = adj_p.T @ v_in + T['dps'][i].T @ vdots[i]
        G[k + '.bias'] = adj_p.sum(0)
        adj_v = (adj_p @ P[k + '.weight'])
    adj_h = adj_v  # [N, 2H+1]
    adj_u = adj_h[:, :H]
    adj_c = adj_h[:, H:2 * H].reshape(B, S, H).sum(1)
    for i in reversed(range(len(ik))):
        k = ik[i]
        u_in, a = T['tape_inp'][i]
        s = sigmoid(a)
        adj_a = adj_u * s + T['das'][i] * adots[i] * (1.0 - s)
        G[k + '.weight'] = adj_a.T @ u_in + T['das'][i].T @ udots[i]
        G[k + '.bias'] = adj_a.sum(0)
        adj_u = adj_a @ P[k + '.weight']
    mlp_backward(P, cs.ctx_keys, T['tape_ctx'], adj_c, 'softplus', True, G)
    return loss, g.reshape(B, S, d), G


# ----------------------------------------------------------------------------- sigma schedule
def sigma_schedule(z, zbar, std_scale, delta):
    """ivae_ardae.py:753-755: lsm = S(z - zbar); std over the nz samples (unbiased) -> mean over
    dims -> * delta.  z [B,nz,d], zbar [B,1,d] -> lsm [B,nz,d], std [B,1,1]."""
    lsm = std_scale * (z - zbar)
    std_qz = lsm.std(axis=1, ddof=1, keepdims=True)
    return lsm, delta * std_qz.mean(axis=2, keepdims=True)


# ----------------------------------------------------------------------------- optimizers
def adam_step(P, G, state, lr, beta1, beta2=0.999, eps=1e-8):
    """utils/optim.py:49-108 (PyTorch-1.2 epsilon placement); params without a grad are skipped."""
    for k, g in G.items():
        st = state.setdefault(k, dict(step=0, exp_avg=np.zeros_like(P[k]), exp_avg_sq=np.zeros_like(P[k])))
        st['step'] += 1
        st['exp_avg'] = beta1 * st['exp_avg'] + (1 - beta1) * g
        st['exp_avg_sq'] = beta2 * st['exp_avg_sq'] + (1 - beta2) * g * g
        bc1 = 1 - beta1 ** st['step']
        bc2 = 1 - beta2 ** st['step']
        denom = (np.sqrt(st['exp_avg_sq']) + eps) / math.sqrt(bc2)
        P[k] = P[k] - (lr / bc1) * st['exp_avg'] / denom


def rmsprop_step(P, G, state, lr, momentum, alpha=0.99, eps=1e-8):
    """torch.optim.RMSprop (centered=False, weight_decay=0) as constructed at ivae_ardae.py:626."""
    for k, g in G.items():
        st = state.setdefault(k, dict(step=0, square_avg=np.zeros_like(P[k]), momentum_buffer=np.zeros_like(P[k])))
        st['step'] += 1
        st['square_avg'] = alpha * st['square_avg'] + (1 - alpha) * g * g
        avg = np.sqrt(st['square_avg']) + eps
        if momentum > 0:
            st['momentum_buffer'] = momentum * st['momentum_buffer'] + g / avg
            P[k] = P[k] - lr * st['momentum_buffer']
        else:
            P[k] = P[k] - lr * g / avg


# ----------------------------------------------------------------------------- the training step
def train_step(spec, cs, Pm, Pc, x_cdae, x_model, noise, hp, opt_state=None):
    """One iteration of train(): ivae_ardae.py:707-846 with cdae_ctx_type='lt0',
    num_cdae_updates=1, every random draw supplied in `noise`:
      enc_cdae [B*nz, n]  (:749), xi [B, nz*nstd, 1] (:761), eps_cdae [B, nz*nstd, d]
      (graddae/mlp.py:22), enc_model [B*nz_model, n] (:801).
    hp: std_scale, delta, nz_cdae, nstd, nz_model, beta, m_lr, m_beta1, d_lr, d_momentum.
    Returns a dict of every intermediate the parity tests compare; updates Pm/Pc in place when
    opt_state is given."""
    S_, delta = hp['std_scale'], hp['delta']
    nz, nstd, nzm, beta = hp['nz_cdae'], hp['nstd'], hp['nz_model'], hp['beta']
    B = x_cdae.shape[0]
    n = spec.noise_dim + (spec.z_dim if spec.kind == 'auxmnist' else 0)  # aux: eps0 | eps packed (see _aux_encoder_forward)
    out = {}
    ctx_type = hp.get('ctx_type', 'lt0')  # ivae_ardae.py:729-741: 'lt0' = mean code, 'data' = the input (2x-1 for MNIST)

    def context_of(x, zbar_):
        if ctx_type == 'data':
            xx = x.reshape(x.shape[0], 1, -1)
            return 2.0 * xx - 1.0 if spec.kind in ('mnist', 'conv', 'auxmnist') else xx
        if ctx_type == 'hidden1a':  # ivae_ardae.py:739-741: model.encode.forward_hidden(x, std=0)
            return aux_encoder_hidden(spec, Pm, x)[:, None, :]
        return zbar_
    # ---- CDAE update (:713-779)
    zbar, _ = encoder_forward(spec, Pm, x_cdae, np.zeros((B, n), dtype=x_cdae.dtype), 1)  # encode(std=0)
    z, _ = encoder_forward(spec, Pm, x_cdae, noise['enc_cdae'], nz)
    lsm, std = sigma_schedule(z, zbar, S_, delta)
    stdmat = std * noise['xi']
    lsm_e = np.repeat(lsm, nstd, axis=1)  # unsqueeze(2).expand(..nstd..).reshape  (:765-767)
    closs, g, Gc = cdae_loss_and_grads(cs, Pc, lsm_e, context_of(x_cdae, zbar), stdmat, noise['eps_cdae'])
    out.update(zbar=zbar, z_cdae=z, std=std, cdae_loss=closs, cdae_score=g, cdae_grads=Gc)
    if opt_state is not None:
        rmsprop_step(Pc, Gc, opt_state.setdefault('cdae', {}), hp['d_lr'], hp['d_momentum'])
    # ---- model update (:781-846)
    Bm = x_model.shape[0]
    mo, tape = model_forward(spec, Pm, x_model, noise['enc_model'], beta, nzm)
    zbar_m, _ = encoder_forward(spec, Pm, x_model, np.zeros((Bm, n), dtype=x_model.dtype), 1)
    lsm_m = S_ * (mo['z'] - zbar_m)
    gm = cdae_glogprob(cs, Pc, lsm_m, context_of(x_model, zbar_m), np.zeros((Bm, nzm, 1), dtype=x_model.dtype))
    dz_extra = S_ * beta * gm / float(Bm * nzm)
    Gm = {}
    model_backward(spec, Pm, tape, beta, nzm, dz_extra, Gm)
    out.update(model_loss=mo['loss'], recon=mo['recon'], prior=mo['prior'], z_model=mo['z'],
               entropy_grad=gm, model_grads=Gm)
    if opt_state is not None:
        adam_step(Pm, Gm, opt_state.setdefault('model', {}), hp['m_lr'], hp['m_beta1'])
    return out


# ----------------------------------------------------------------------------- IWS evaluator
def iws_logprob(spec, P, x, enc_noise, eta):
    """ImplicitPosteriorVAE.logprob_w_cov_gaussian_posterior: toy.py:878-939 / mnist.py:378-437
    (+ utils/stat.py:65-85,127-158, torch MultivariateNormal).  x [b,D]; enc_noise [b,S,n] encoder
    noise per image; eta [b,S,d] the standard-normal draw behind MVN.rsample.
    Returns (mean log p_hat(x), per-image values)."""
    b = x.shape[0]
    S = enc_noise.shape[1]
    d = spec.z_dim
    assert S >= 2 * d  # mnist.py:382
    vals = np.zeros(b, dtype=x.dtype)
    for i in range(b):
        z3, _ = encoder_forward(spec, P, x[i:i + 1], enc_noise[i], S)
        z = z3[0]  # [S, d]
        mu = z.mean(0)
        zc = z - mu
        cov = zc.T @ zc / (S - 1)  # get_covmat, utils/stat.py:127-158
        if spec.kind == 'auxmnist':
            cov = cov + 1e-5 * np.eye(d)  # models/ivae/auxmnist.py:350
        Lc = np.linalg.cholesky(cov)
        newz = mu + eta[i] @ Lc.T
        logq = -0.5 * (eta[i] ** 2).sum(1) - np.log(np.diag(Lc)).sum() - 0.5 * d * LOG2PI
        logprior = (-0.5 * (newz ** 2 + LOG2PI)).sum(1)
        heads, _ = decoder_forward(spec, P, newz)
        xi = np.repeat(x[i:i + 1], S, axis=0)
        loglik = -recon_rows(spec, heads, xi)
        w = loglik + logprior - logq
        m = w.max()
        vals[i] = math.log(np.exp(w - m).mean() + 1e-10) + m  # mnist.py:431-434
    return vals.mean(), vals
