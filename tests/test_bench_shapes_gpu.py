"""GPU parity at the BENCHMARKED widths (BASELINE.json configs 1 and 2), where the fused tcgen05 chain kernels, the
bf16 activation spill, the 300 -> 320 zero-padded model layers and the in-kernel Philox noise actually run -- the
reference-generated fixtures (tests/golden) are tiny and only reach the per-layer fallback kernels.

The checker is the numpy oracle (oracle/ardae_oracle.py, itself pinned to the reference by tests/test_oracle_golden.py)
in fp64 on the same weights, inputs and noise.  Row counts are reduced (B = 16, nz = 64: N = 1024 CDAE rows) so the
oracle finishes in seconds; widths, depths, activations and hyper-parameters are the configs' own
(run_vae_25gaussians.sh / run_vae_dbmnist.sh:37).

Tolerances (tf32 tensor-core operands, bf16 spill storage; stated in DESIGN.md):
  losses rel <= 2e-3 | z rel <= 1e-5 | sigma scale rel <= 1e-4 | score / entropy gradient rel-L2 <= 1e-2
  every gradient tensor rel-L2 <= 2e-2 | RMSprop update rel-L2 <= 0.1, Adam update <= 0.15 (sign flips at t = 1)
These hold with the oracle evaluated on the latents the GPU produced (kernel isolation); end to end from the inputs the
entropy gradient and the two CDAE layers next to the input are looser (1e-1 / 5e-2) because std_scale = 1e4 amplifies
the 1e-6..1e-5 accuracy of z -- see check_against_oracle.
"""
import numpy as np
import pytest
import torch

import ardae_oracle as orc
from golden_util import rel_err

pytestmark = pytest.mark.gpu

HP = dict(std_scale=10000., delta=0.1, nz_cdae=64, nstd=1, nz_model=1, beta=1.0, m_lr=1e-4, m_beta1=0.5, d_lr=1e-4,
          d_momentum=0.5)
CONFIGS = {
    # configs[1] / [2]: dbMNIST-shape MNISTIPVAE 784/300/100/32 + mlp-grad CDAE h=256 L=5
    'config2': dict(kind='mnist', D=784, n=100, h=300, z=32, layers=2, nonlin='softplus', cdae_h=256, cdae_L=5),
    # configs[0]: 25gaussians ToyIPVAE relu h=256 n=10 z=2 + mlp-grad CDAE h=256 L=3
    'config1': dict(kind='toy', D=2, n=10, h=256, z=2, layers=2, nonlin='relu', cdae_h=256, cdae_L=3),
}


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).float().cuda()


def build(cfg, seed=11):
    import ardae
    torch.manual_seed(seed)
    cls = ardae.ToyIPVAE if cfg['kind'] == 'toy' else ardae.MNISTIPVAE
    model = cls(input_dim=cfg['D'], noise_dim=cfg['n'], h_dim=cfg['h'], num_hidden_layers=cfg['layers'],
                nonlinearity=cfg['nonlin'], enc_type='concat', z_dim=cfg['z']).cuda()
    cdae = ardae.MLPGradCARDAE(input_dim=cfg['z'], context_dim=cfg['z'], std=1., h_dim=cfg['cdae_h'],
                               num_hidden_layers=cfg['cdae_L'], nonlinearity='softplus').cuda()
    mopt = ardae.Adam(model.parameters(), lr=HP['m_lr'], betas=(HP['m_beta1'], 0.999))
    copt = ardae.RMSprop(cdae.parameters(), lr=HP['d_lr'], momentum=HP['d_momentum'])
    return model, cdae, mopt, copt


def params64(mod):
    return {k: v.detach().cpu().numpy().astype(np.float64) for k, v in mod.state_dict().items()}


def specs(cfg):
    spec = orc.ModelSpec(cfg['kind'], cfg['D'], cfg['n'], cfg['h'], cfg['z'], cfg['layers'], cfg['nonlin'])
    cs = orc.CdaeSpec(cfg['z'], cfg['z'], cfg['cdae_h'], cfg['cdae_L'])
    return spec, cs


def inputs(cfg, B, rng):
    if cfg['kind'] == 'toy':  # 25-Gaussians grid (datasets/toy.py:195-250)
        centers = np.linspace(-4, 4, 5)
        mk = lambda: np.stack([rng.choice(centers, B), rng.choice(centers, B)], 1) + np.sqrt(0.1) * rng.randn(B, 2)
    else:
        mk = lambda: (rng.rand(B, cfg['D']) < 0.13).astype(np.float64)
    return mk(), mk()


def check_against_oracle(cfg, step, model, cdae, xc, xm, noise_np, out, Pm0, Pc0):
    """Replays the iteration in the oracle and compares everything the step produced, at two levels.

    A. END TO END from the inputs (fp64 oracle).  At the reference initialisation z ~ 20 and std_scale = 1e4, so the
       1e-6..1e-5 relative accuracy of z (3xTF32, ~22 significant bits) becomes an absolute error of 0.1..1 in the CDAE
       input S (z - zbar) ~ 1e4..1e5, which moves the softplus kinks of the randomly initialised score network: the
       quantities closest to that input (entropy gradient, first inp_encode layers) see it amplified.  The fp32
       arithmetic of the reference has the same sensitivity (printed: oracle in float32 vs float64).  Bounds: losses
       2e-3, z 1e-5, sigma scale 1e-4, gradient tensors 2e-2 (5e-2 for the two inp_encode layers next to the input),
       entropy gradient 1e-1.
    B. KERNEL ISOLATION: the oracle is fed the latents the GPU produced (z, zbar of both minibatches), so only the CDAE
       kernels are compared: score / entropy gradient 1e-2, every CDAE gradient tensor 2e-2 -- the stated tolerances."""
    spec, cs = specs(cfg)
    Pm, Pc = {k: v.copy() for k, v in Pm0.items()}, {k: v.copy() for k, v in Pc0.items()}
    ref = orc.train_step(spec, cs, Pm, Pc, xc, xm, noise_np, HP, opt_state={})
    f32 = lambda d: {k: np.asarray(v, dtype=np.float32) for k, v in d.items()}
    ref32 = orc.train_step(spec, cs, f32(Pm0), f32(Pc0), xc.astype(np.float32), xm.astype(np.float32), f32(noise_np), HP,
                           opt_state={})
    # ---------------- A: end to end
    for k in ('cdae_loss', 'model_loss', 'recon', 'prior'):
        e = abs(out[k].item() - float(ref[k])) / abs(float(ref[k]))
        assert e <= 2e-3, (k, e, out[k].item(), float(ref[k]))
    assert rel_err(out['std'].cpu().numpy().ravel(), ref['std'].ravel()) <= 1e-4
    assert rel_err(out['z_model'].cpu().numpy().ravel(), ref['z_model'].ravel()) <= 1e-5
    eg = rel_err(out['entropy_grad'].cpu().numpy().ravel(), ref['entropy_grad'].ravel())
    eg32 = rel_err(ref32['entropy_grad'].ravel(), ref['entropy_grad'].ravel())
    print('A entropy gradient: cuda vs fp64 oracle %.2e | fp32 oracle vs fp64 oracle %.2e' % (eg, eg32))
    assert eg <= 1e-1, ('entropy_grad', eg, eg32)
    got_grads = {}
    for mod, key in ((cdae, 'cdae_grads'), (model, 'model_grads')):
        ar = mod._arena
        errs, errs32 = {}, {}
        for k, nme in enumerate(ar.names):
            if nme in ref[key]:
                got_grads[(key, nme)] = ar.view(ar.stage_flat, k).cpu().numpy().astype(np.float64)
                errs[nme] = rel_err(got_grads[(key, nme)], ref[key][nme])
                errs32[nme] = rel_err(ref32[key][nme], ref[key][nme])
        print('A', key, 'cuda vs fp64 | fp32 oracle vs fp64:', {n: '%.1e|%.1e' % (errs[n], errs32[n]) for n in errs})
        for nme in errs:
            near_input = key == 'cdae_grads' and nme.startswith('inp_encode.layers.') and int(nme.split('.')[2]) < 2
            assert errs[nme] <= (5e-2 if near_input else 2e-2), (key, nme, errs[nme], errs32[nme])
    # post-step parameters (optimizers consumed the gradients above).  t = 1: RMSprop's lr*g/(sqrt(0.01 g^2)+eps) and
    # Adam's lr*g/(|g|+eps) are sign-like, elements with |g| below the gradient error flip
    def upd(after, before, ref_after):
        got = np.concatenate([(after[k] - before[k]).ravel() for k in sorted(before)])
        exp = np.concatenate([(ref_after[k] - before[k]).ravel() for k in sorted(before)])
        return rel_err(got, exp)
    Pc_after = params64(cdae)
    ue_c, ue_m = upd(Pc_after, Pc0, Pc), upd(params64(model), Pm0, Pm)
    print('A update rel: cdae %.2e model %.2e' % (ue_c, ue_m))
    assert ue_c <= 0.1 and ue_m <= 0.15
    assert np.array_equal(Pc_after['neglogprob.fc.bias'], Pc0['neglogprob.fc.bias'])  # never receives a gradient
    # ---------------- B: the CDAE kernels on the latents the GPU produced
    ln = {k: v.detach().cpu().numpy().astype(np.float64) for k, v in step.last_noise.items()}
    B = xc.shape[0]
    nz, d = HP['nz_cdae'], cfg['z']
    zc, zbc = ln['z_cdae'].reshape(B, nz, d), ln['zbar_cdae'].reshape(B, 1, d)
    lsm, std = orc.sigma_schedule(zc, zbc, HP['std_scale'], HP['delta'])
    closs, g, Gc = orc.cdae_loss_and_grads(cs, Pc0, lsm, zbc, std * noise_np['xi'], noise_np['eps_cdae'])
    assert abs(out['cdae_loss'].item() - float(closs)) <= 2e-3 * abs(float(closs))
    errsB = {n: rel_err(got_grads[('cdae_grads', n)], Gc[n]) for n in Gc if ('cdae_grads', n) in got_grads}
    print('B cdae_grads (oracle on the GPU latents):', {n: '%.1e' % e for n, e in errsB.items()})
    for n, e in errsB.items():
        assert e <= 2e-2, ('B cdae_grad', n, e)
    zm, zbm = out['z_model'].cpu().numpy().astype(np.float64).reshape(B, 1, d), ln['zbar_model'].reshape(B, 1, d)
    gm = orc.cdae_glogprob(cs, Pc_after, HP['std_scale'] * (zm - zbm), zbm, np.zeros((B, 1, 1)))
    egB = rel_err(out['entropy_grad'].cpu().numpy().ravel(), gm.ravel())
    print('B entropy gradient (oracle on the GPU latents, updated CDAE): %.2e' % egB)
    assert egB <= 1e-2, ('B entropy_grad', egB)
    return ref


@pytest.mark.parametrize('name', sorted(CONFIGS))
def test_full_step_injected_noise_vs_oracle(name):
    """(i)/(ii): whole TrainStep (CDAE update + model update + RMSprop + Adam) with injected noise."""
    import ardae
    cfg = CONFIGS[name]
    B, nz, d, n = 16, HP['nz_cdae'], cfg['z'], cfg['n']
    model, cdae, mopt, copt = build(cfg)
    Pm0, Pc0 = params64(model), params64(cdae)
    rng = np.random.RandomState(5)
    xc, xm = inputs(cfg, B, rng)
    noise = dict(enc_cdae=rng.randn(B * nz, n), xi=rng.randn(B, nz, 1), eps_cdae=rng.randn(B, nz, d),
                 enc_model=rng.randn(B, n))
    step = ardae.TrainStep(model, cdae, mopt, copt, std_scale=HP['std_scale'], delta=HP['delta'], nz_cdae=nz,
                           nstd=1, nz_model=1)
    step.keep_noise = True
    out = step(t(xc), t(xm), beta=HP['beta'], noise={k: t(v) for k, v in noise.items()})
    torch.cuda.synchronize()
    check_against_oracle(cfg, step, model, cdae, xc, xm, noise, out, Pm0, Pc0)


def drawn_noise(step, B, nz, d):
    """What the in-kernel Philox draws of the last iteration were, in the oracle's layout (xi recovered from
    sigma = std_b * xi)."""
    ln = {k: v.detach().cpu().numpy().astype(np.float64) for k, v in step.last_noise.items()}
    xi = ln['sigma'].reshape(B, nz) / ln['std'].reshape(B, 1)
    return dict(enc_cdae=ln['enc_cdae'], xi=xi.reshape(B, nz, 1), eps_cdae=ln['eps_cdae'].reshape(B, nz, d),
                enc_model=ln['enc_model'])


@pytest.mark.parametrize('graph', [False, True])
def test_philox_path_vs_oracle(graph):
    """(iv): the path bench.py times -- noise drawn on the device (Philox), side-stream overlap, and (graph=True) a
    CUDA-graph replay -- checked by feeding the drawn noise to the oracle."""
    import ardae
    cfg = CONFIGS['config2']
    B, nz, d = 16, HP['nz_cdae'], cfg['z']
    model, cdae, mopt, copt = build(cfg)
    rng = np.random.RandomState(6)
    xc, xm = inputs(cfg, B, rng)
    step = ardae.TrainStep(model, cdae, mopt, copt, std_scale=HP['std_scale'], delta=HP['delta'], nz_cdae=nz,
                           nstd=1, nz_model=1, graph=graph, seed=3)
    step.keep_noise = True
    # graph mode: two eager iterations, the third call captures and replays; compare the fourth (a pure replay)
    for _ in range(3 if graph else 0):
        step(t(xc), t(xm), beta=HP['beta'])
    torch.cuda.synchronize()
    if graph:
        assert step._g is not None, 'graph was not captured'
    Pm0, Pc0 = params64(model), params64(cdae)
    # the oracle restarts its optimizers: compare gradients / losses of this iteration, not the (stateful) update
    out = step(t(xc), t(xm), beta=HP['beta'])
    torch.cuda.synchronize()
    noise = drawn_noise(step, B, nz, d)
    for k, v in noise.items():
        assert np.isfinite(v).all() and abs(v.mean()) < 0.2 and 0.8 < v.std() < 1.2, k  # N(0,1) draws
    spec, cs = specs(cfg)
    ref = orc.train_step(spec, cs, dict(Pm0), dict(Pc0), xc, xm, noise, HP, opt_state=None)
    for k in ('cdae_loss', 'model_loss'):
        e = abs(out[k].item() - float(ref[k])) / abs(float(ref[k]))
        assert e <= 2e-3, (k, e)
    for mod, key in ((cdae, 'cdae_grads'), (model, 'model_grads')):
        ar = mod._arena
        for k, nme in enumerate(ar.names):
            if nme in ref[key]:
                e = rel_err(ar.view(ar.stage_flat, k).cpu().numpy(), ref[key][nme])
                # end to end from the inputs: the two CDAE layers next to S (z - zbar) see the z accuracy amplified
                # (check_against_oracle, level A).  In graph mode the compared iteration follows three updates whose
                # lr*sign(g) steps are not bit-reproducible (float atomics), so the parameters -- and with them this
                # amplified error -- differ from run to run: measured 3e-2 ... 1.1e-1 over repeated runs.  There the two
                # tensors get a sanity bound only (0.3, i.e. cosine > 0.95); kernel accuracy is pinned by level B (2e-2)
                # and by the eager variant of this test (5e-2).
                near_input = key == 'cdae_grads' and nme.startswith('inp_encode.layers.') and int(nme.split('.')[2]) < 2
                assert e <= ((0.3 if graph else 5e-2) if near_input else 2e-2), (key, nme, e)


def test_replays_draw_disjoint_noise():
    """Consecutive graph replays (and the eager iterations before them) must not share noise: with a replay-seed
    stride of 1 the model-update encoder noise of replay r+1 equalled the first rows of the CDAE-update encoder noise
    of replay r.  Checks every pair of noise buffers over 6 iterations for common values."""
    import ardae
    cfg = CONFIGS['config2']
    B, nz = 16, HP['nz_cdae']
    model, cdae, mopt, copt = build(cfg)
    rng = np.random.RandomState(7)
    xc, xm = inputs(cfg, B, rng)
    step = ardae.TrainStep(model, cdae, mopt, copt, std_scale=HP['std_scale'], delta=HP['delta'], nz_cdae=nz,
                           nstd=1, nz_model=1, graph=True, seed=9)
    step.keep_noise = True
    seen = []
    for it in range(6):
        step(t(xc), t(xm), beta=HP['beta'])
        torch.cuda.synchronize()
        for k in ('enc_cdae', 'eps_cdae', 'enc_model', 'sigma'):
            seen.append((it, k, step.last_noise[k].detach().cpu().numpy().ravel()[:1600].copy()))
    assert step._g is not None
    for i in range(len(seen)):
        for j in range(i + 1, len(seen)):
            a, b = seen[i][2], seen[j][2]
            common = np.intersect1d(a.view(np.uint32), b.view(np.uint32)).size
            assert common <= 2, (seen[i][:2], seen[j][:2], common)  # 1600 fp32 normals: chance collisions ~ 0


def test_beta_annealing_does_not_recapture():
    """beta lives in a device scalar: an annealed beta (utils/msc.py:53-55, --beta-annealing) replays the SAME graph,
    and the run equals the eager run with the same betas (same seeds -> same noise)."""
    import ardae
    cfg = CONFIGS['config1']
    B, nz = 16, HP['nz_cdae']
    rng = np.random.RandomState(8)
    xc, xm = inputs(cfg, B, rng)
    betas = [ardae.annealing_func(0.01, 1.0, 50., i) for i in range(6)]
    res = []
    for graph in (False, True):
        model, cdae, mopt, copt = build(cfg)
        step = ardae.TrainStep(model, cdae, mopt, copt, std_scale=HP['std_scale'], delta=HP['delta'], nz_cdae=nz,
                               nstd=1, nz_model=1, graph=graph, seed=4)
        graphs, losses = set(), []
        for b in betas:
            out = step(t(xc), t(xm), beta=b)
            losses.append(out['model_loss'].item())
            if graph and step._g is not None:
                graphs.add(id(step._g[0]))
        torch.cuda.synchronize()
        if graph:
            assert len(graphs) == 1, 'beta changes must not re-capture'
        res.append((params64(model), losses))
    (pe, le), (pg, lg) = res
    # eager and replayed runs issue the same launches with the same seeds; red.global.add summation order differs from
    # run to run and the sign-like RMSprop / Adam updates of the first iterations amplify it
    for a, b in zip(le, lg):
        assert abs(a - b) <= 5e-3 * abs(a), (le, lg)
    for k in pe:
        assert rel_err(pg[k], pe[k]) <= 5e-4, k  # ~2 % of the accumulated update (8 steps of lr 1e-4 on weights ~0.04)
    assert len({round(v, 4) for v in lg}) > 3  # the loss does respond to beta


@pytest.mark.parametrize('name', ['config2', 'config1', 'conv28'])
def test_submodule_calls(name):
    """decode(z) / generate / encode._forward_inp / encode._forward_all (SURVEY 8b 'must expose') vs the oracle."""
    import ardae
    from golden_util import build_model, load_case, sub
    if name == 'conv28':
        z, meta = load_case('conv28')
        model = build_model(meta)
        model.load_state_dict({k: torch.from_numpy(np.asarray(v)).float() for k, v in sub(z, 'm0/').items()})
        model = model.cuda()
        spec = orc.ModelSpec('conv', 784, 100, 800, 32, 0, 'softplus')
        spec.img_c = 1
        D, n, zd = 784, 100, 32
    else:
        cfg = CONFIGS[name]
        model = build(cfg)[0]
        spec = specs(cfg)[0]
        D, n, zd = cfg['D'], cfg['n'], cfg['z']
    P = params64(model)
    rng = np.random.RandomState(9)
    B, nz = 6, 5
    x = (rng.rand(B, D) < 0.2).astype(np.float64) if spec.kind != 'toy' else rng.randn(B, D) * 2
    eps = rng.randn(B * nz, n)
    zz = rng.randn(7, zd)
    # decode
    heads_ref, _ = orc.decoder_forward(spec, P, zz)
    out = model.decode(t(zz))
    heads = [h.reshape(7, -1).cpu().numpy() for h in out[1:]]
    refs = list(heads_ref) if isinstance(heads_ref, (tuple, list)) else [heads_ref]
    assert len(heads) == len(refs) == (2 if spec.kind == 'toy' else 1)
    for h, r in zip(heads, refs):
        assert rel_err(h, np.asarray(r).reshape(7, -1)) <= 1e-5
    assert out[0].shape == out[1].shape
    # generate: shapes + the decoder really ran on the drawn z
    xg, mean, zg = model.generate(batch_size=3)
    assert zg.shape == (3, zd) and xg.shape[0] == 3 and torch.isfinite(mean).all()
    hz, _ = orc.decoder_forward(spec, P, zg.cpu().numpy().astype(np.float64))
    hz0 = np.asarray(hz[0] if isinstance(hz, (tuple, list)) else hz).reshape(3, -1)
    exp = hz0 if spec.kind == 'toy' else 1.0 / (1.0 + np.exp(-hz0))
    assert rel_err(mean.reshape(3, -1).cpu().numpy(), exp) <= 1e-4
    # _forward_inp + expand + _forward_all == encode (the reference's Encoder.forward, ivae/mnist.py:99-121)
    z_ref, _ = orc.encoder_forward(spec, P, x, eps, nz)
    inp = model.encode._forward_inp(t(x))
    nos = model.encode._forward_nos(noise=t(eps))
    inp_e = inp.unsqueeze(1).expand(-1, nz, -1).contiguous().view(B * nz, -1)
    z_parts = model.encode._forward_all(inp_e, nos).view(B, nz, -1)
    z_full = model.encode(t(x), noise=t(eps), nz=nz)
    ztol = 2e-5 if spec.kind == 'conv' else 1e-5  # conv: 612- and 800-long fp32 contractions
    assert rel_err(z_parts.cpu().numpy(), z_ref) <= ztol
    assert rel_err(z_full.cpu().numpy(), z_ref) <= ztol


def test_conv_gemm_path_equals_direct_kernels(monkeypatch):
    """ConvIPVAE 28x28: the im2col + tcgen05-GEMM conv layers (default) against the direct fp32 CUDA-core kernels
    (ARDAE_CONV_GEMM=0) on the same weights, inputs and noise: logits / z to 1e-5 (both forwards are fp32-accurate),
    every gradient tensor to 5e-3 (tf32 backward GEMMs vs exact fp32 accumulation)."""
    import ardae
    torch.manual_seed(3)
    ref = ardae.ConvIPVAE(input_height=28, input_channels=1, z_dim=32, noise_dim=100, nonlinearity='softplus')
    state = {k: v.clone() for k, v in ref.state_dict().items()}
    x = (torch.rand(12, 784, device='cuda') < 0.2).float()
    noise = torch.randn(12 * 2, 100, device='cuda')
    res = {}
    for flag in ('0', '1'):
        monkeypatch.setenv('ARDAE_CONV_GEMM', flag)
        m = ardae.ConvIPVAE(input_height=28, input_channels=1, z_dim=32, noise_dim=100, nonlinearity='softplus')
        m.load_state_dict(state)
        m = m.cuda()
        xhat, mean, z, loss, recon, prior = m(x, beta=0.7, nz=2, noise=noise)
        loss.backward()
        torch.cuda.synchronize()
        res[flag] = dict(mean=mean.detach().cpu().numpy().astype(np.float64), z=z.detach().cpu().numpy().astype(np.float64),
                         loss=loss.item(), grads={k: p.grad.detach().cpu().numpy().astype(np.float64)
                                                  for k, p in m.named_parameters()})
    a, b = res['0'], res['1']
    assert abs(a['loss'] - b['loss']) <= 1e-6 * abs(a['loss'])
    assert rel_err(b['z'], a['z']) <= 1e-5 and rel_err(b['mean'], a['mean']) <= 1e-5
    for k in a['grads']:
        assert rel_err(b['grads'][k], a['grads'][k]) <= 5e-3, (k, rel_err(b['grads'][k], a['grads'][k]))
