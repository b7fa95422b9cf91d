"""Run the UNMODIFIED reference modules with injected noise  --  TEST INFRASTRUCTURE ONLY.

Used (a) by ``oracle/make_golden.py`` in the build container to generate ``tests/golden/*.npz`` and
(b) by ``bench.py --impl reference`` / the ``cpu_baseline`` leg when a copy of the reference is
reachable (``$ARDAE_REF``, ``baseline/_ref`` or ``/root/reference``).  Nothing in the product package
imports this file.  The step body restates ``ivae_ardae.py:707-846`` line by line as a function (the
script itself cannot run: missing matplotlib/seaborn/tensorboardX/torchcontrib and a py2-style
``iterator.next()``) but every model / CDAE / optimizer call goes to the reference's
own classes.
"""
import os
import sys
import types

_REF_MODULES = None


def find_reference():
    here = os.path.dirname(os.path.abspath(__file__))
    if os.environ.get('ARDAE_NO_REF'):
        return None
    for cand in (os.environ.get('ARDAE_REF'), os.path.join(here, '..', 'baseline', '_ref'), '/root/reference'):
        if cand and os.path.isfile(os.path.join(cand, 'models', 'graddae', 'mlp.py')):
            return os.path.abspath(cand)
    return None


def import_reference():
    """Returns (utils, models) of the reference, stubbing its plotting imports (utils/msc.py:12-17)."""
    global _REF_MODULES
    if _REF_MODULES is not None:
        return _REF_MODULES
    ref = find_reference()
    if ref is None:
        raise ImportError('reference tree not found (set $ARDAE_REF)')

    class _Any(types.ModuleType):
        def __getattr__(self, n):
            if n.startswith('__'):
                raise AttributeError(n)
            return lambda *a, **k: None

    for m in ('matplotlib', 'matplotlib.pyplot', 'seaborn'):
        if m not in sys.modules:
            sys.modules[m] = _Any(m)
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k in ('utils', 'models') or
             k.startswith('utils.') or k.startswith('models.')}
    sys.path.insert(0, ref)
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            import utils as ref_utils
            import models as ref_models
    finally:
        sys.path.remove(ref)
        # keep the reference's modules importable only through the returned handles
        for k in list(sys.modules):
            if k in ('utils', 'models') or k.startswith('utils.') or k.startswith('models.'):
                sys.modules['_ardae_ref_' + k] = sys.modules.pop(k)
        sys.modules.update(saved)
    _REF_MODULES = (ref_utils, ref_models)
    return _REF_MODULES


def build_reference(kind, model_kwargs, cdae_kwargs, dtype=None, seed=0, cdae_kind='grad'):
    """Construct the reference model + CDAE as ivae_ardae.py:295-314,583-606 does
    (cdae_kind 'grad' = --cdae mlp-grad, 'res' = --cdae mlp-res)."""
    import torch
    _, net = import_reference()
    torch.manual_seed(seed)
    if kind == 'conv':
        model = net.ConvIPVAE(**model_kwargs)
    elif kind == 'auxmnist':  # ivae_ardae.py:455-466
        model = net.MNISTAuxIPVAE(enc_type='simple', clip_z0_logvar='none', clip_z_logvar='none', **model_kwargs)
    else:
        model = (net.ToyIPVAE if kind == 'toy' else net.MNISTIPVAE)(enc_type='concat', **model_kwargs)
    cls = net.MLPGradCARDAE if cdae_kind == 'grad' else net.MLPResCARDAE
    cdae = cls(std=1., noise_type='gaussian', enc_ctx=True, enc_input=True, **cdae_kwargs)
    if dtype is not None:
        model, cdae = model.to(dtype), cdae.to(dtype)
    return model, cdae


def build_optimizers(model, cdae, hp):
    """ivae_ardae.py:550 (utils.Adam) and :626 (torch.optim.RMSprop)."""
    import torch
    utils, _ = import_reference()
    mopt = utils.Adam(model.parameters(), lr=hp['m_lr'], betas=(hp['m_beta1'], 0.999))
    copt = torch.optim.RMSprop(cdae.parameters(), lr=hp['d_lr'], momentum=hp['d_momentum'])
    return mopt, copt


class _NoiseQueue(object):
    """Replaces Encoder.sample_noise (ivae/toy.py:61-65) by a queue of pre-drawn tensors."""

    def __init__(self, enc):
        self.enc, self.q = enc, []

    def __call__(self, batch_size, std=None, device=None):
        import torch
        std = std if std is not None else getattr(self.enc, 'std', 1.)
        if std == 0:
            w = next(self.enc.parameters())
            return torch.zeros(batch_size, self.enc.noise_dim, dtype=w.dtype, device=w.device)
        eps = self.q.pop(0)
        assert eps.shape == (batch_size, self.enc.noise_dim), (eps.shape, batch_size)
        return std * eps


def _install_aux_noise(model, cmod, q):
    """Encoder._forward (ivae/auxmnist.py:107-115) draws eps0 [R, n] then eps [R, 1, d] through the module-level
    sample_noise(sz, device); the injected tensors pack them as [R, n + d].  Calls made with an empty queue are the
    std = 0 passes (the draws are multiplied by 0 there): they get zeros."""
    import torch
    n_, d_ = model.noise_dim, model.z_dim
    pend = []

    def aux_noise(sz, device):
        w = next(model.parameters())
        if len(sz) == 2:
            if q.q:
                t = q.q.pop(0)
                assert t.shape == (sz[0], n_ + d_), (t.shape, sz)
                pend.append(t[:, n_:])
                return t[:, :n_].to(w.dtype)
            pend.append(None)
            return torch.zeros(*sz, dtype=w.dtype)
        e = pend.pop(0)
        return torch.zeros(*sz, dtype=w.dtype) if e is None else e.reshape(sz).to(w.dtype)
    cmod.sample_noise = aux_noise


def _ref_context(model, x, ctx_type, mnist_like):
    """ivae_ardae.py:729-741 / :807-819: cdae_ctx_type 'lt0' (latent of the mean code), 'data' (the input itself) or
    'hidden1a' (hidden features of the hierarchical encoder)."""
    if ctx_type == 'hidden1a':
        return model.encode.forward_hidden(x, std=0).detach().unsqueeze(1)
    if ctx_type == 'data':
        context = x.unsqueeze(1)
        if mnist_like:  # "if 'mnist' in opt.dataset"
            context = 2 * context - 1
            context = context.view(x.size(0), 1, -1)
        return context
    return model.encode(x, std=0).detach()


def ref_train_step(model, cdae, mopt, copt, x_cdae, x_model, noise, hp, do_step=True, ctx_type='lt0', mnist_like=True):
    """ivae_ardae.py:707-846 with cdae_ctx_type in {'lt0', 'data'}, num_cdae_updates=1, injected noise.
    `noise` holds torch tensors: enc_cdae, xi, eps_cdae, enc_model (see oracle.train_step)."""
    import warnings
    import torch
    S_, delta = hp['std_scale'], hp['delta']
    nz, nstd, nzm, beta = hp['nz_cdae'], hp['nstd'], hp['nz_model'], hp['beta']
    aux = model.__class__.__module__.endswith('auxmnist')
    q = _NoiseQueue(model.encode)
    if not aux:
        model.encode.sample_noise = q
    # ConvIPVAE.forward / forward_hidden draw through a module-level sample_noise(sz, std, device)
    # (ivae/conv.py:24-27,190,207): route that to the same queue
    cmod = sys.modules.get('_ardae_ref_' + model.__class__.__module__) or sys.modules.get(model.__class__.__module__)
    c_orig = getattr(cmod, 'sample_noise', None) if cmod is not None else None
    if c_orig is not None and not aux:
        cmod.sample_noise = lambda sz, std=None, device=None: q(sz[0], std=std, device=device)
    if aux:
        _install_aux_noise(model, cmod, q)
    out = {}
    model.train(); cdae.train()
    # ---- update cdae (:713-779)
    copt.zero_grad()
    B = x_cdae.size(0)
    context = _ref_context(model, x_cdae, ctx_type, mnist_like)        # :729-741
    latent_mean = model.encode(x_cdae, std=0).detach()                  # :748
    q.q.append(noise['enc_cdae'])
    latent = model.forward_hidden(x_cdae, nz=nz).detach()               # :749
    latent_sub_mean = S_ * (latent - latent_mean)                       # :753
    std_qz = torch.std(latent_sub_mean, dim=1, keepdim=True)            # :754
    std = delta * torch.mean(std_qz, dim=2, keepdim=True)               # :755
    stdmat = std * noise['xi']                                          # :761
    sz = list(latent_sub_mean.size())
    lsm = latent_sub_mean.unsqueeze(2).expand(B, nz, nstd, sz[-1]).reshape(B, nz * nstd, sz[-1])  # :765
    eps_c = noise['eps_cdae'].reshape(B * nz * nstd, -1)
    cdae.add_noise = lambda inp, s=None: (inp + s * eps_c, eps_c)      # graddae/mlp.py:21-23
    _, cdae_loss = cdae(lsm, context, std=stdmat, scale=S_)             # :768
    cdae_loss.backward()                                                # :771
    # the score itself, for comparison (same perturbed input)
    g = cdae.glogprob((lsm + stdmat * noise['eps_cdae']).detach().clone(), context, std=stdmat, scale=S_).detach()
    out.update(zbar=latent_mean, z_cdae=latent, std=std, cdae_loss=cdae_loss.detach(), cdae_score=g,
               cdae_grads={k: (p.grad.detach().clone() if p.grad is not None else None)
                           for k, p in cdae.named_parameters()})
    if do_step:
        copt.step()                                                     # :779
    # ---- update model (:781-846)
    model.train(); cdae.eval()
    mopt.zero_grad()
    Bm = x_model.size(0)
    q.q.append(noise['enc_model'])
    _, _, latent, model_loss, recon_loss, prior_loss = model(x_model, beta=beta, eta=0., lmbd=0., nz=nzm)  # :801
    model_loss.backward(retain_graph=True)                              # :804
    context = _ref_context(model, x_model, ctx_type, mnist_like)       # :807-819
    latent_mean = model.encode(x_model, std=0).detach()                 # :826
    lsm_m = S_ * (latent - latent_mean).detach()                        # :827
    stdmat0 = torch.zeros(Bm, nzm, 1, dtype=x_model.dtype)              # :828
    latent = latent.view(Bm, nzm, -1)
    grad = cdae.glogprob(lsm_m, context, std=stdmat0, scale=S_).detach()  # :829
    (S_ * (latent - latent_mean)).backward(beta * grad.detach() / float(Bm * nzm))  # :834
    out.update(model_loss=model_loss.detach(), recon=recon_loss, prior=prior_loss, z_model=latent.detach(),
               entropy_grad=grad,
               model_grads={k: p.grad.detach().clone() for k, p in model.named_parameters()})
    if do_step:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            mopt.step()                                                 # :846
    if not aux:
        del model.encode.sample_noise
    del cdae.add_noise
    if c_orig is not None:
        cmod.sample_noise = c_orig
    return out


def ref_iws(model, x, enc_noise, eta):
    """evaluate_iws -> model.logprob (ivae_ardae.py:662; ivae/mnist.py:378-437) with injected noise:
    enc_noise [b,S,n] (one sample_noise draw per image), eta [b,S,d] (MVN.rsample's normal draw)."""
    import torch
    import torch.distributions.multivariate_normal as mvn
    b, S, _ = enc_noise.shape
    q = _NoiseQueue(model.encode)
    aux = model.__class__.__module__.endswith('auxmnist')
    cmod = sys.modules.get('_ardae_ref_' + model.__class__.__module__) or sys.modules.get(model.__class__.__module__)
    c_orig = getattr(cmod, 'sample_noise', None) if cmod is not None else None
    etas = [eta[i].reshape(1, S, -1) for i in range(b)]
    if aux:  # one Encoder._forward call over all images (ivae/auxmnist.py:346)
        q.q = [enc_noise.reshape(b * S, -1)]
        _install_aux_noise(model, cmod, q)
    else:
        q.q = [enc_noise[i] for i in range(b)]
        model.encode.sample_noise = q
    orig = mvn._standard_normal
    mvn._standard_normal = lambda shape, dtype, device: etas.pop(0).reshape(shape).to(dtype)
    try:
        model.eval()
        with torch.no_grad():
            val = model.logprob(x, sample_size=S)
    finally:
        mvn._standard_normal = orig
        if aux:
            cmod.sample_noise = c_orig
        else:
            del model.encode.sample_noise
        model.train()
    return val
