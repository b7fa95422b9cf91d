"""GPU parity of the whole AR-DAE training step (CDAE update + model update + both optimizers):
fused TrainStep and the drop-in module API vs the reference-generated fixtures (2 consecutive steps).

Tolerances (tf32 backward sweeps, 3xTF32 forward):
  losses (cdae, vae, recon, prior)      rel <= 2e-3
  z, zbar                               rel <= 1e-5      (fp32-accurate forward)
  sigma scale std_b                     rel <= 1e-4
  CDAE score / entropy gradient         rel-L2 <= 1e-2
  every parameter gradient tensor       rel-L2 <= 2e-2   (auxmnist_small, second step: 6e-2, see below)
  parameter UPDATE (after - before)     rel-L2 <= 5e-2 (RMSprop) / 0.15 (Adam): at t = 1 Adam's update is
                                        lr*g/(|g|+eps) ~ lr*sign(g), so the few elements whose |g| is below
                                        the 1e-3 gradient error flip sign; the optimizer arithmetic itself is
                                        pinned to 1e-6 in test_optimizers_match_oracle.
"""
import os

import numpy as np
import pytest
import torch

from golden_util import CASES, GOLDEN_DIR, build_cdae, build_model, is_lite, load_case, num_steps, pick, rel_err, sub

pytestmark = pytest.mark.gpu


def arena_np(mod, flat):
    ar = mod._arena
    return {n: ar.view(flat, k).detach().cpu().numpy().copy() for k, n in enumerate(ar.names)}


def build(meta, z):
    import ardae
    c, hp = meta['cdae'], meta['hp']
    model = build_model(meta)
    cdae = build_cdae(meta)
    f = lambda d: {k: torch.from_numpy(np.asarray(v)).float() for k, v in d.items()}
    model.load_state_dict(f(sub(z, 'm0/')))
    cdae.load_state_dict(f(sub(z, 'c0/')))
    model, cdae = model.cuda(), cdae.cuda()
    mopt = ardae.Adam(model.parameters(), lr=hp['m_lr'], betas=(hp['m_beta1'], 0.999))
    copt = ardae.RMSprop(cdae.parameters(), lr=hp['d_lr'], momentum=hp['d_momentum'])
    return model, cdae, mopt, copt


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).float().cuda()


def params_np(mod):
    return {k: v.detach().cpu().numpy().astype(np.float64) for k, v in mod.state_dict().items()}


def update_err(before, after, ref_before, ref_after):
    got = np.concatenate([(after[k] - before[k]).ravel() for k in sorted(before)])
    ref = np.concatenate([(ref_after[k] - ref_before[k]).ravel() for k in sorted(before)])
    return rel_err(got, ref)


@pytest.mark.parametrize('name', CASES)
def test_fused_step_matches_reference_fixture(name):
    import ardae
    z, meta = load_case(name)
    hp = meta['hp']
    model, cdae, mopt, copt = build(meta, z)
    step = ardae.TrainStep(model, cdae, mopt, copt, std_scale=hp['std_scale'], delta=hp['delta'],
                           nz_cdae=hp['nz_cdae'], nstd=hp['nstd'], nz_model=hp['nz_model'],
                           ctx_type=meta.get('ctx_type', 'lt0'))
    ref_m_prev, ref_c_prev = sub(z, 'm0/'), sub(z, 'c0/')
    lite = is_lite(name, meta) and not meta.get('sampled')  # fixture holds one step and no post-step weights
    def pk(d):  # `sampled` fixtures keep big model tensors at fixed positions (idempotent: already-picked arrays pass)
        return {k: (v if ('sample_idx/' + k) in z.files and np.asarray(v).size == z['sample_idx/' + k].size
                    else pick(z, k, v)) for k, v in d.items()}
    for s in range(num_steps(name, meta)):
        p = 's%d/' % s
        noise = {k: t(v) for k, v in sub(z, p + 'noise/').items()}
        m_before, c_before = params_np(model), params_np(cdae)
        out = step(t(z[p + 'x_cdae']), t(z[p + 'x_model']), beta=hp['beta'], noise=noise)
        torch.cuda.synchronize()
        # x3 case, step 1: the fixture itself is ill-conditioned there (see test_oracle_golden)
        loose = name.endswith('_x3') and s == 1
        ltol = 2e-2 if loose else 2e-3
        for k in ('cdae_loss', 'model_loss', 'recon', 'prior'):
            e = abs(out[k].item() - float(z[p + k])) / abs(float(z[p + k]))
            assert e <= ltol, (s, k, e)
        assert rel_err(out['std'].cpu().numpy(), z[p + 'std'].ravel()) <= (1e-2 if loose else 1e-4)
        # step >= 1 starts from parameters that carry the (tf32-level, Adam-normalised) update error of the step before
        assert rel_err(out['z_model'].cpu().numpy().ravel(), z[p + 'z_model'].ravel()) <= (1e-3 if loose else ((2e-5 if meta['kind'] == 'conv' else 1e-5) if s == 0 else (2e-4 if meta['kind'] == 'conv' else 1e-4)))
        eg = rel_err(out['entropy_grad'].cpu().numpy().ravel(), z[p + 'entropy_grad'].ravel())
        assert eg <= (5e-2 if loose else 1e-2), (s, 'entropy_grad', eg)
        m_after, c_after = params_np(model), params_np(cdae)
        if lite:
            ref_m_after, ref_c_after, ue_c, ue_m = ref_m_prev, ref_c_prev, 0.0, 0.0
        else:
            ref_m_after, ref_c_after = sub(z, p + 'm_after/'), sub(z, p + 'c_after/')
            ue_c = update_err(c_before, c_after, ref_c_prev, ref_c_after)
            ue_m = update_err(pk(m_before), pk(m_after), pk(ref_m_prev), ref_m_after)
        print('%s step %d: cdae_loss %.6g (ref %.6g) model_loss %.6g (ref %.6g) entropy_grad rel %.2e '
              'update rel: cdae %.2e model %.2e' % (name, s, out['cdae_loss'].item(), float(z[p + 'cdae_loss']),
                                                    out['model_loss'].item(), float(z[p + 'model_loss']), eg, ue_c, ue_m))
        assert ue_c <= (0.3 if loose else 5e-2) and ue_m <= (0.3 if loose else 0.15)
        # the gradients the optimizers consumed are still in the stage arenas
        # step >= 1 of the hierarchical case starts from parameters carrying step 0's update error (4-5 % of an
        # RMSprop / Adam update); with the fp64 oracle alone, that perturbation moves the step-1 CDAE gradients by
        # 1.5e-2 .. 3.6e-2 (scripts/step1_sensitivity.py).  Step 0 (same kernels, exact start) holds 2e-2.
        gtol = 0.2 if loose else (6e-2 if (meta['kind'] == 'auxmnist' and s >= 1) else 2e-2)
        for mod, pref in ((cdae, 'cdae_grads/'), (model, 'model_grads/')):
            ar, ref = mod._arena, sub(z, p + pref)
            for k, nme in enumerate(ar.names):
                if nme in ref:
                    got = ar.view(ar.stage_flat, k).cpu().numpy()
                    e = rel_err(pick(z, nme, got) if mod is model else got, ref[nme])
                    assert e <= gtol, (s, nme, e)
        if meta.get('cdae_kind', 'grad') == 'grad':
            assert np.array_equal(c_after['neglogprob.fc.bias'], c_before['neglogprob.fc.bias'])  # never updated
        else:
            assert not np.array_equal(c_after['dae.fc.bias'], c_before['dae.fc.bias'])  # residual CDAE: it is
        ref_m_prev, ref_c_prev = ref_m_after, ref_c_after


@pytest.mark.parametrize('name', ['toy_small', 'mnist_small', 'conv_small', 'mnist_small_res'])
def test_dropin_loop_matches_fused(name):
    """The reference's own step body (ivae_ardae.py:713-846) written against the drop-in module API
    (autograd .backward() calls, optimizer objects) must give what the fused driver gives."""
    import ardae
    z, meta = load_case(name)
    hp = meta['hp']
    S_, delta, nz, nstd, nzm, beta = hp['std_scale'], hp['delta'], hp['nz_cdae'], hp['nstd'], hp['nz_model'], hp['beta']
    noise = {k: t(v) for k, v in sub(z, 's0/noise/').items()}
    xc, xm = t(z['s0/x_cdae']), t(z['s0/x_model'])
    # ---- fused
    model, cdae, mopt, copt = build(meta, z)
    out = ardae.TrainStep(model, cdae, mopt, copt, std_scale=S_, delta=delta, nz_cdae=nz, nstd=nstd,
                          nz_model=nzm)(xc, xm, beta=beta, noise=noise)
    pm_f, pc_f = params_np(model), params_np(cdae)
    gm_f, gc_f = arena_np(model, model._arena.stage_flat), arena_np(cdae, cdae._arena.stage_flat)
    # ---- drop-in loop, written like the reference
    model, cdae, mopt, copt = build(meta, z)
    B = xc.size(0)
    cdae_optimizer, model_optimizer = copt, mopt
    cdae_optimizer.zero_grad()
    context = model.encode(xc, std=0).detach()
    latent_mean = model.encode(xc, std=0).detach()
    latent = model.encode(xc, noise=noise['enc_cdae'], nz=nz).detach()
    latent_sub_mean = S_ * (latent - latent_mean)
    std_qz = torch.std(latent_sub_mean, dim=1, keepdim=True)
    std = delta * torch.mean(std_qz, dim=2, keepdim=True)
    stdmat = std * noise['xi']
    lsm = latent_sub_mean.unsqueeze(2).expand(B, nz, nstd, latent.size(-1)).reshape(B, nz * nstd, -1)
    _, cdae_loss = cdae(lsm, context, std=stdmat, scale=S_, eps=noise['eps_cdae'])
    cdae_loss.backward()
    cdae_optimizer.step()
    model_optimizer.zero_grad()
    _, _, latent, model_loss, recon_loss, prior_loss = model(xm, beta=beta, eta=0., lmbd=0., nz=nzm,
                                                             noise=noise['enc_model'])
    model_loss.backward(retain_graph=True)
    context = model.encode(xm, std=0).detach()
    latent_mean = model.encode(xm, std=0).detach()
    latent_sub_mean = S_ * (latent - latent_mean).detach()
    stdmat0 = torch.zeros(xm.size(0), nzm, 1, device='cuda')
    grad = cdae.glogprob(latent_sub_mean, context, std=stdmat0, scale=S_).detach()
    (S_ * (latent - latent_mean)).backward(beta * grad.detach() / float(xm.size(0) * nzm))
    model_optimizer.step()
    torch.cuda.synchronize()
    assert abs(cdae_loss.item() - out['cdae_loss'].item()) <= 1e-5 * abs(cdae_loss.item())
    assert abs(model_loss.item() - out['model_loss'].item()) <= 1e-5 * abs(model_loss.item())
    pm_d, pc_d = params_np(model), params_np(cdae)
    # gradients: the drop-in loop accumulates two encoder backward passes into .grad, the fused driver folds them
    # into one pass, so tf32 rounding differs slightly; parameters: Adam/RMSprop's g/(|g|+eps) turns tiny
    # gradient differences into O(lr) update differences, so allow 5 % of the update on top of 1e-5 of the parameter
    gm_d, gc_d = arena_np(model, model._arena.grad_flat), arena_np(cdae, cdae._arena.grad_flat)
    for k in gm_f:
        assert rel_err(gm_d[k], gm_f[k]) <= 1e-2, ('model grad', k)
    for k in gc_f:
        if k != 'neglogprob.fc.bias':
            assert rel_err(gc_d[k], gc_f[k]) <= 1e-2, ('cdae grad', k)

    def close(d, f, p0):
        return np.linalg.norm(d - f) <= 1e-5 * np.linalg.norm(f) + 5e-2 * np.linalg.norm(f - p0)
    for k in pm_f:
        assert close(pm_d[k], pm_f[k], np.asarray(z['m0/' + k], dtype=np.float32)), (k, rel_err(pm_d[k], pm_f[k]))
    for k in pc_f:
        assert close(pc_d[k], pc_f[k], np.asarray(z['c0/' + k], dtype=np.float32)), (k, rel_err(pc_d[k], pc_f[k]))
    if meta.get('cdae_kind', 'grad') == 'grad':
        assert cdae.neglogprob.fc.bias.grad is None
    else:
        assert cdae.dae.fc.bias.grad is not None


@pytest.mark.parametrize('kind', ['adam', 'rmsprop'])
def test_optimizers_match_oracle(kind):
    """Flat optimizer kernels vs the oracle's restatement of utils/optim.py:49-108 and torch RMSprop
    (themselves pinned to the reference by test_oracle_golden), 3 steps, exact gradients."""
    import ctypes
    import ardae_oracle as orc
    from ardae import _lib
    rng = np.random.RandomState(0)
    n = 4096 + 64
    P = {'w': rng.randn(n)}
    st = {}
    p = t(P['w']); s1 = torch.zeros_like(p); s2 = torch.zeros_like(p)
    for step in range(1, 4):
        g = rng.randn(n) * 10.0 ** rng.uniform(-9, 1, size=n)
        if kind == 'adam':
            orc.adam_step(P, {'w': g}, st, lr=1e-3, beta1=0.5)
            _lib.check(_lib.lib().ardae_adam_step(_lib.ptr(p), _lib.ptr(t(g)), _lib.ptr(s1), _lib.ptr(s2), n, 1e-3, 0.5,
                                                  0.999, 1e-8, step, 1.0, _lib.stream_ptr()))
        else:
            orc.rmsprop_step(P, {'w': g}, st, lr=1e-3, momentum=0.5)
            _lib.check(_lib.lib().ardae_rmsprop_step(_lib.ptr(p), _lib.ptr(t(g)), _lib.ptr(s1), _lib.ptr(s2), n, 1e-3,
                                                     0.99, 1e-8, 0.5, 1.0, _lib.stream_ptr()))
        torch.cuda.synchronize()
        assert np.max(np.abs(p.cpu().numpy() - P['w'])) <= 2e-6


def test_graph_replay_matches_eager():
    """TrainStep(graph=True): two eager iterations, then the iteration is captured into one CUDA graph and replayed.
    The first replay runs the very launches (and Philox seeds) the third eager iteration would have issued, so the
    parameters after three iterations must agree; later replays draw fresh noise through the device counter and
    keep the host-side optimizer step counters in line."""
    import ardae
    z, meta = load_case('mnist_small')
    hp = meta['hp']
    xc, xm = t(z['s0/x_cdae']), t(z['s0/x_model'])
    res = []
    for graph in (False, True):
        model, cdae, mopt, copt = build(meta, z)
        step = ardae.TrainStep(model, cdae, mopt, copt, std_scale=hp['std_scale'], delta=hp['delta'],
                               nz_cdae=hp['nz_cdae'], nstd=hp['nstd'], nz_model=hp['nz_model'], graph=graph, seed=7)
        for _ in range(3):
            step(xc, xm, beta=hp['beta'])
        torch.cuda.synchronize()
        p3 = (params_np(model), params_np(cdae))
        losses = []
        for _ in range(3):
            out = step(xc, xm, beta=hp['beta'])
            losses.append((out['cdae_loss'].item(), out['model_loss'].item()))
        if graph:
            assert step._g is not None, 'graph was not captured'
        t_m = mopt.state[next(iter(model.parameters()))]['step']
        t_c = copt.state[next(iter(cdae.parameters()))]['step']
        res.append((p3, losses, t_m, t_c))
    (pe, le, tme, tce), (pg, lg, tmg, tcg) = res
    assert (tme, tce) == (tmg, tcg) == (6, 6)
    for a, b in zip(pe, pg):
        for k in a:
            assert rel_err(b[k], a[k]) <= 1e-5, k
    for (c1, m1), (c2, m2) in zip(le, lg):
        assert np.isfinite([c1, m1, c2, m2]).all()
    # fresh noise per replay: the CDAE loss is a noisy estimate, three identical values would mean frozen seeds
    assert len({round(c, 7) for c, _ in lg}) == 3


def test_resume_from_reference_checkpoint():
    """Checkpoint interchange: load the files the REFERENCE wrote after its first iteration (weights + utils.Adam /
    torch RMSprop state), run the second iteration on the B200 path with the fixture's inputs and noise, and land on
    the parameters the reference reached after ITS second iteration."""
    import os
    import ardae
    from golden_util import GOLDEN_DIR
    z, meta = load_case('mnist_small')
    hp = meta['hp']
    model, cdae = build_model(meta), build_cdae(meta)
    mopt = ardae.Adam(model.parameters(), lr=hp['m_lr'], betas=(hp['m_beta1'], 0.999))
    copt = ardae.RMSprop(cdae.parameters(), lr=hp['d_lr'], momentum=hp['d_momentum'])
    ck = os.path.join(GOLDEN_DIR, 'ref_ckpt_mnist_small')
    ardae.load_checkpoint(model, mopt, ck, filename='model-checkpoint.pth.tar')
    ardae.load_checkpoint(cdae, copt, ck, filename='cdae-checkpoint.pth.tar')
    model, cdae = model.cuda(), cdae.cuda()
    # optimizer state follows the parameters to the device
    for opt in (mopt, copt):
        for st in opt.state.values():
            for k, v in st.items():
                if torch.is_tensor(v):
                    st[k] = v.cuda()
    step = ardae.TrainStep(model, cdae, mopt, copt, std_scale=hp['std_scale'], delta=hp['delta'],
                           nz_cdae=hp['nz_cdae'], nstd=hp['nstd'], nz_model=hp['nz_model'])
    m_before, c_before = params_np(model), params_np(cdae)
    noise = {k: t(v) for k, v in sub(z, 's1/noise/').items()}
    out = step(t(z['s1/x_cdae']), t(z['s1/x_model']), beta=hp['beta'], noise=noise)
    torch.cuda.synchronize()
    assert abs(out['cdae_loss'].item() - float(z['s1/cdae_loss'])) <= 2e-3 * abs(float(z['s1/cdae_loss']))
    assert abs(out['model_loss'].item() - float(z['s1/model_loss'])) <= 2e-3 * abs(float(z['s1/model_loss']))
    ue_c = update_err(c_before, params_np(cdae), sub(z, 's0/c_after/'), sub(z, 's1/c_after/'))
    ue_m = update_err(m_before, params_np(model), sub(z, 's0/m_after/'), sub(z, 's1/m_after/'))
    print('resume: update rel err cdae %.2e model %.2e' % (ue_c, ue_m))
    assert ue_c <= 5e-2 and ue_m <= 5e-2
    assert mopt.state[next(iter(model.parameters()))]['step'] == 2


def test_two_cdae_updates_per_iteration():
    """num_cdae_updates = 2 (run_vae_dbmnist.sh:25): the fused driver takes a list of CDAE minibatches; it must equal
    cdae_update, cdae_update, model_update issued by hand (same injected noise for both)."""
    import ardae
    z, meta = load_case('mnist_small')
    hp = meta['hp']
    noise = {k: t(v) for k, v in sub(z, 's0/noise/').items()}
    x1, x2, xm = t(z['s0/x_cdae']), t(z['s1/x_cdae']), t(z['s0/x_model'])
    res = []
    for fused in (True, False):
        model, cdae, mopt, copt = build(meta, z)
        step = ardae.TrainStep(model, cdae, mopt, copt, std_scale=hp['std_scale'], delta=hp['delta'],
                               nz_cdae=hp['nz_cdae'], nstd=hp['nstd'], nz_model=hp['nz_model'], num_cdae_updates=2)
        if fused:
            step([x1, x2], xm, beta=hp['beta'], noise=noise)
        else:
            step.overlap = False
            step.cdae_update(x1, noise)
            step.cdae_update(x2, noise)
            step.model_update(xm, hp['beta'], noise)
        torch.cuda.synchronize()
        res.append((params_np(model), params_np(cdae), copt.state[next(iter(cdae.parameters()))]['step']))
    (pm_a, pc_a, sa), (pm_b, pc_b, sb) = res
    assert sa == sb == 2
    for a, b in ((pm_a, pm_b), (pc_a, pc_b)):
        for k in a:
            assert rel_err(a[k], b[k]) <= 1e-5, k


def test_aux_encoder_full_width_vs_oracle():
    """Hierarchical encoder at the reference's MNIST widths (784/300/100/32, 2 hidden layers, ivae_ardae.py:455-467):
    z, the mean code, the 'hidden1a' context cat(h0, h) and the decoder logits against the fp64 oracle."""
    import ardae
    import ardae_oracle as orc
    torch.manual_seed(5)
    model = ardae.MNISTAuxIPVAE(input_dim=784, noise_dim=100, h_dim=300, z_dim=32, num_hidden_layers=2).cuda()
    spec = orc.ModelSpec('auxmnist', 784, 100, 300, 32, 2, 'softplus')
    P = params_np(model)
    rng = np.random.RandomState(2)
    B, nz = 6, 8
    x = (rng.rand(B, 784) > 0.6).astype(np.float64)
    eps = rng.randn(B * nz, 132)
    z_ref, _ = orc.encoder_forward(spec, P, x, eps, nz)
    zbar_ref, _ = orc.encoder_forward(spec, P, x, np.zeros((B, 132)), 1)
    z, zbar, hid = model._encode_hidden(t(x), t(eps), nz)
    assert rel_err(z.cpu().numpy(), z_ref) <= 1e-5
    assert rel_err(zbar.cpu().numpy(), zbar_ref) <= 1e-5
    assert rel_err(hid.cpu().numpy(), orc.aux_encoder_hidden(spec, P, x)) <= 1e-5
    assert rel_err(model.encode(t(x), nz=nz, noise=t(eps)).cpu().numpy(), z_ref) <= 1e-5
    assert rel_err(model.encode.forward_hidden(t(x), std=0).cpu().numpy(), orc.aux_encoder_hidden(spec, P, x)) <= 1e-5
    assert rel_err(model.forward_hidden(t(x), std=0).cpu().numpy(), zbar_ref) <= 1e-5
    zz = rng.randn(10, 32)
    heads, _ = orc.decoder_forward(spec, P, zz)
    xs, logit = model.decode(t(zz))
    assert xs.shape == logit.shape == (10, 784)
    assert rel_err(logit.cpu().numpy(), heads[0] if isinstance(heads, (tuple, list)) else heads) <= 1e-5
    out, mean, zg = model.generate(4)
    assert out.shape == mean.shape == (4, 784) and zg.shape == (4, 32)


def test_stale_forward_raises_in_backward():
    """The plans keep ONE set of activations / staged gradients per shape: backward() of a forward that was
    superseded by a newer forward() of the same shape must raise instead of returning the newer input's gradients."""
    z, meta = load_case('mnist_small')
    model, cdae, _, _ = build(meta, z)
    x = t(z['s0/x_model'])
    loss_a = model(x, beta=1.0, nz=1)[3]
    loss_b = model(x * 0.5, beta=1.0, nz=1)[3]
    with pytest.raises(RuntimeError, match='forward\\(\\) was called again'):
        loss_a.backward()
    loss_b.backward()
    assert model.decode.reparam.logit_fn.weight.grad is not None
    B = x.size(0)
    xs, ctx = torch.randn(B, 4, meta['cdae']['input_dim'], device='cuda'), torch.randn(B, 1, meta['cdae']['context_dim'], device='cuda')
    std = torch.rand(B, 4, 1, device='cuda')
    _, l1 = cdae(xs, ctx, std=std)
    _, l2 = cdae(xs * 2, ctx, std=std)
    with pytest.raises(RuntimeError, match='forward\\(\\) was called again'):
        l1.backward()
    l2.backward()


@pytest.mark.parametrize('graph', [False, True])
def test_staged_inputs_match_direct_inputs(graph):
    """TrainStep.stage(pinned host pair) + step() must give what step(device pair) gives (same Philox seeds),
    eagerly and through the CUDA-graph replay, and `losses` packs the four scalars."""
    import ardae
    z, meta = load_case('mnist_small')
    hp = meta['hp']
    xs = [(torch.from_numpy(np.ascontiguousarray(z['s%d/x_cdae' % (i % 2)])).float().pin_memory(),
           torch.from_numpy(np.ascontiguousarray(z['s%d/x_model' % (i % 2)])).float().pin_memory()) for i in range(5)]
    outs = []
    for staged in (False, True):
        model, cdae, mopt, copt = build(meta, z)
        step = ardae.TrainStep(model, cdae, mopt, copt, std_scale=hp['std_scale'], delta=hp['delta'], nz_cdae=hp['nz_cdae'],
                               nstd=hp['nstd'], nz_model=hp['nz_model'], seed=7, graph=graph)
        rec = []
        if staged:
            step.stage(*xs[0])
        for i in range(5):
            if staged:
                o = step(beta=hp['beta'])
                if i + 1 < 5:
                    step.stage(*xs[i + 1])
            else:
                o = step(xs[i][0].cuda(), xs[i][1].cuda(), beta=hp['beta'])
            l = o['losses'].cpu().numpy().copy()
            assert l.shape == (4,)
            assert l[0] == o['cdae_loss'].item() and l[1] == o['model_loss'].item()
            rec.append(l)
        torch.cuda.synchronize()
        outs.append((np.stack(rec), params_np(model)))
        if staged:
            with pytest.raises(RuntimeError, match='nothing staged'):
                step(beta=hp['beta'])
    # same seeds, same kernels; the scalar reductions use float atomics, so two runs agree to an ulp, not bitwise
    assert np.allclose(outs[0][0], outs[1][0], rtol=1e-6, atol=0)
    for k in outs[0][1]:
        assert rel_err(outs[0][1][k], outs[1][1][k]) <= 1e-5, k


def test_training_trajectory_tracks_reference():
    """400 consecutive iterations on the 25-Gaussians problem, same initial weights, data and injected noise as the
    reference's own fp64 run (oracle/make_curve.py -> tests/golden/toy_curve.npz).  With std_scale = 1e4 the problem
    amplifies rounding along a trajectory: the reference's own fp32 run -- what a user of the reference executes --
    drifts from its fp64 run by up to 14 % (CDAE loss, first 25 iterations) and 12 % (25-iteration means).  That drift
    is the yardstick: the GPU trajectory (itself not bit-reproducible: float atomics) must stay within 3x of it (+2 %)
    of the fp64 trajectory, per quantity, both per iteration over the first 25 iterations and in 25-iteration means over
    the whole run; the final importance-weighted log-likelihood of 64 held-out points within 0.25 nat."""
    import ardae
    import curve_util as cu
    import json
    z = np.load(os.path.join(GOLDEN_DIR, 'toy_curve.npz'), allow_pickle=True)
    C = json.loads(str(z['meta']))
    hp, B, T, seed = C['hp'], C['B'], C['T'], C['seed']
    meta = dict(kind='toy', model=C['model'], cdae=C['cdae'], hp=hp)
    model, cdae, mopt, copt = build(meta, z)
    step = ardae.TrainStep(model, cdae, mopt, copt, std_scale=hp['std_scale'], delta=hp['delta'], nz_cdae=hp['nz_cdae'],
                           nstd=hp['nstd'], nz_model=hp['nz_model'])
    n, d = C['model']['noise_dim'], C['model']['z_dim']
    got = torch.zeros(T, 5, device='cuda')
    for it in range(T):
        nz = {k: t(v) for k, v in cu.noise(seed, it, B, n, d, hp).items()}
        o = step(t(cu.batch(seed, it, B, 0)), t(cu.batch(seed, it, B, 1)), beta=hp['beta'], noise=nz)
        got[it, :4] = o['losses']
        got[it, 4] = o['std'].mean()
    got = got.cpu().numpy().astype(np.float64)
    ref, ref32 = z['curve'], z['curve_ref_fp32']
    if os.environ.get('ARDAE_CURVE_DUMP'):
        np.save(os.environ['ARDAE_CURVE_DUMP'], got)
    assert np.isfinite(got).all()
    win = lambda a: a.reshape(T // 25, 25, 5).mean(axis=1)
    early = lambda a: np.abs(a[:25] / ref[:25] - 1.0).max(axis=0)
    smooth = lambda a: np.abs(win(a) / win(ref) - 1.0).max(axis=0)
    x, en, eta = cu.iws_inputs(seed, 64, 64, n, d)
    lp = model.logprob(t(x), sample_size=64, noise=t(en), eta=t(eta)).item()
    print('quantities: cdae_loss, model_loss, recon, prior, sigma scale')
    print('first 25 iterations, max rel deviation from the fp64 reference: GPU', early(got), ' reference fp32', early(ref32))
    print('25-iteration means,  max rel deviation from the fp64 reference: GPU', smooth(got), ' reference fp32', smooth(ref32))
    print('last 25 iterations: model_loss %.4f (fp64 ref %.4f, fp32 ref %.4f); iws %.4f (fp64 ref %.4f, fp32 ref %.4f)' % (
        got[-25:, 1].mean(), ref[-25:, 1].mean(), ref32[-25:, 1].mean(), lp, float(z['iws_logprob']),
        float(z['iws_logprob_ref_fp32'])))
    assert (early(got) <= 3.0 * early(ref32) + 0.02).all(), (early(got), early(ref32))
    assert (smooth(got) <= 3.0 * smooth(ref32) + 0.02).all(), (smooth(got), smooth(ref32))
    assert abs(lp - float(z['iws_logprob'])) <= 0.25
    # the run must actually have trained: the ELBO loss falls from 32.8 to ~5.5
    assert got[0, 1] > 30.0 and got[-25:, 1].mean() < 6.5
