"""GPU parity: CUDA MLPGradCARDAE (through the C ABI) vs the numpy oracle and the reference-generated
golden fixtures.  Tolerances (tf32 tensor-core operands, fp32 accumulate; stated per north_star):
  loss            rel err <= 2e-3
  score g         rel-L2 <= 1e-2 and cosine >= 0.9999
  every parameter gradient tensor: rel-L2 <= 2e-2 and cosine >= 0.999
"""
import numpy as np
import pytest
import torch

import ardae_oracle as orc
from golden_util import CASES, build_cdae, cdae_spec, cosine, load_case, rel_err, sub

pytestmark = pytest.mark.gpu

LOSS_TOL, SCORE_TOL, GRAD_TOL = 2e-3, 1e-2, 2e-2


def make_cdae(d, c, H, L, state=None, seed=0, kind='grad'):
    import ardae
    torch.manual_seed(seed)
    cls = ardae.MLPGradCARDAE if kind == 'grad' else ardae.MLPResCARDAE
    m = cls(input_dim=d, context_dim=c, std=1., h_dim=H, num_hidden_layers=L,
            nonlinearity='softplus', noise_type='gaussian', enc_ctx=True, enc_input=True)
    if state is not None:
        m.load_state_dict({k: torch.from_numpy(np.asarray(v)).float() for k, v in state.items()})
    return m.cuda()


def check_against(m, cs, P64, x, ctx, std, eps, label):
    loss_o, g_o, G_o = orc.cdae_loss_and_grads(cs, P64, x, ctx, std, eps)
    dev = 'cuda'
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().to(dev)
    m.zero_grad()
    _, loss = m(t(x), t(ctx), std=t(std), scale=1.0, eps=t(eps))
    loss.backward()
    torch.cuda.synchronize()
    lerr = abs(loss.item() - loss_o) / abs(loss_o)
    serr, scos = rel_err(m.last_score.cpu().numpy(), g_o), cosine(m.last_score.cpu().numpy(), g_o)
    worst = (0.0, None)
    worst_cos = (1.0, None)
    for k, p in m.named_parameters():
        if k == 'neglogprob.fc.bias':  # energy network: the output bias never receives a gradient
            assert p.grad is None
            continue
        assert p.grad is not None, k
        e = rel_err(p.grad.cpu().numpy(), G_o[k])
        cth = cosine(p.grad.cpu().numpy(), G_o[k])
        worst = max(worst, (e, k))
        worst_cos = min(worst_cos, (cth, k))
    print('%s: loss %.6g (oracle %.6g, rel %.2e) score rel %.2e cos %.6f worst grad rel %.2e (%s) cos %.6f (%s)' % (
        label, loss.item(), loss_o, lerr, serr, scos, worst[0], worst[1], worst_cos[0], worst_cos[1]))
    assert lerr <= LOSS_TOL
    assert serr <= SCORE_TOL and scos >= 0.9999
    assert worst[0] <= GRAD_TOL, worst
    assert worst_cos[0] >= 0.999, worst_cos
    # glogprob at the perturbed point must reproduce the score of the training pass
    B, S, d = x.shape
    g2 = m.glogprob(t(x + std * eps), t(ctx), std=t(std), scale=1.0)
    assert rel_err(g2.cpu().numpy(), g_o) <= SCORE_TOL


@pytest.mark.parametrize('name', CASES)
def test_cdae_matches_reference_fixture(name):
    """Same weights / inputs / noise as the reference run that produced the fixture (step 0)."""
    z, meta = load_case(name)
    c = meta['cdae']
    cs = cdae_spec(meta)
    P64 = sub(z, 'c0/')
    m = make_cdae(c['input_dim'], c['context_dim'], c['h_dim'], c['num_hidden_layers'], state=P64,
                  kind=meta.get('cdae_kind', 'grad'))
    hp = meta['hp']
    lsm = hp['std_scale'] * (z['s0/z_cdae'] - z['s0/zbar'])
    lsm = np.repeat(lsm, hp['nstd'], axis=1)
    std = z['s0/std'] * z['s0/noise/xi']
    eps = z['s0/noise/eps_cdae']
    # context rows: the mean code ('lt0') or the input mapped to 2x-1 ('data'), ivae_ardae.py:729-741
    ct = meta.get('ctx_type', 'lt0')
    if ct == 'hidden1a':  # cat(h0, h) of the hierarchical encoder at std = 0 (:739-741), from the oracle (not stored)
        from test_oracle_golden import specs
        Pm = {k: np.asarray(v, dtype=np.float64) for k, v in sub(z, 'm0/').items()}
        ctx = orc.aux_encoder_hidden(specs(meta)[0], Pm, np.asarray(z['s0/x_cdae'], dtype=np.float64))[:, None, :]
    else:
        ctx = z['s0/zbar'] if ct == 'lt0' else (2.0 * z['s0/x_cdae'] - 1.0)[:, None, :]
    # the fixture's own numbers (reference, fp64) agree with the oracle by test_oracle_golden
    check_against(m, cs, P64, lsm, ctx, std, eps, name)
    ref_loss = float(z['s0/cdae_loss'])
    m.zero_grad()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().cuda()
    _, loss = m(t(lsm), t(ctx), std=t(std), eps=t(eps))
    assert abs(loss.item() - ref_loss) / abs(ref_loss) <= LOSS_TOL
    assert rel_err(m.last_score.cpu().numpy(), z['s0/cdae_score']) <= SCORE_TOL


@pytest.mark.parametrize('cfg', [
    dict(d=8, c=8, H=64, L=3, B=8, S=32, wscale=1.0),
    dict(d=2, c=2, H=256, L=3, B=16, S=64, wscale=1.0),     # config-1 widths
    dict(d=32, c=32, H=256, L=5, B=32, S=256, wscale=1.0),  # config-2 widths, N = 8192
    dict(d=32, c=32, H=256, L=5, B=8, S=128, wscale=3.0),   # saturated activations
    dict(d=5, c=3, H=36, L=2, B=3, S=50, wscale=1.0),       # ragged: N=150, odd dims
    # residual CDAE (--cdae mlp-res): fused chains (H = 256 / 64) and the per-layer path (H = 36, and H = 512 as in
    # run_vae_dbmnist.sh:25)
    dict(d=32, c=32, H=256, L=5, B=32, S=256, wscale=1.0, kind='res'),
    dict(d=8, c=8, H=64, L=3, B=8, S=32, wscale=3.0, kind='res'),
    dict(d=5, c=3, H=36, L=2, B=3, S=50, wscale=1.0, kind='res'),
    dict(d=32, c=32, H=512, L=3, B=4, S=100, wscale=1.0, kind='res'),
])
def test_cdae_matches_oracle(cfg):
    rng = np.random.RandomState(7)
    d, c, H, L, B, S = cfg['d'], cfg['c'], cfg['H'], cfg['L'], cfg['B'], cfg['S']
    m = make_cdae(d, c, H, L, seed=11, kind=cfg.get('kind', 'grad'))
    if cfg['wscale'] != 1.0:
        with torch.no_grad():
            for p in m.parameters():
                if p.dim() == 2:
                    p.mul_(cfg['wscale'])
    P64 = {k: v.detach().cpu().numpy().astype(np.float64) for k, v in m.state_dict().items()}
    cs = orc.CdaeSpec(d, c, H, L, kind=cfg.get('kind', 'grad'))
    x = rng.randn(B, S, d) * 2.0
    ctx = rng.randn(B, 1, c)
    std = 0.3 * rng.randn(B, S, 1)
    eps = rng.randn(B, S, d)
    check_against(m, cs, P64, x, ctx, std, eps, str(cfg))


def test_cdae_generated_noise_and_accumulation():
    """In-kernel Philox noise: eps is N(0,1); backward accumulates (autograd semantics)."""
    m = make_cdae(4, 4, 32, 3, seed=3)
    x = torch.randn(16, 64, 4, device='cuda')
    ctx = torch.randn(16, 1, 4, device='cuda')
    std = 0.1 * torch.randn(16, 64, 1, device='cuda')
    _, loss = m(x, ctx, std=std, seed=123)
    e1 = m.last_eps.clone()
    loss.backward()
    g1 = m.inp_encode.fc.weight.grad.clone()
    _, loss2 = m(x, ctx, std=std, seed=123)
    assert torch.equal(e1, m.last_eps)
    loss2.backward()
    assert torch.allclose(m.inp_encode.fc.weight.grad, 2 * g1, rtol=1e-4, atol=1e-7)
    assert abs(e1.mean().item()) < 0.05 and abs(e1.std().item() - 1.0) < 0.05
    assert m.neglogprob.fc.bias.grad is None


def test_cdae_shape_errors():
    m = make_cdae(4, 4, 32, 3)
    with pytest.raises(AssertionError):
        m(torch.zeros(8, 4, device='cuda'), torch.zeros(8, 1, 4, device='cuda'))
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 3, 4), torch.zeros(2, 1, 4))  # CPU tensors: no fallback


# ---------------------------------------------------------------------------------------------------------
# Full size (BASELINE.json configs[1]: B=512 data rows x 256 samples = 131072 CDAE rows, d=c=32, H=256, L=5):
# the oracle cannot run the whole batch in seconds, so parity is checked through size-independent properties.
FULL = dict(d=32, c=32, H=256, L=5, B=512, S=256)


def _full_inputs(seed=3):
    g = torch.Generator(device='cuda').manual_seed(seed)
    B, S, d, c = FULL['B'], FULL['S'], FULL['d'], FULL['c']
    x = 30.0 * torch.randn(B, S, d, device='cuda', generator=g)
    ctx = torch.randn(B, 1, c, device='cuda', generator=g)
    std = 0.5 * torch.randn(B, S, 1, device='cuda', generator=g)
    eps = torch.randn(B, S, d, device='cuda', generator=g)
    return x, ctx, std, eps


def test_full_size_rows_match_oracle_on_a_subset():
    """Rows are independent in sweeps 1-2: the scores of 4 of the 512 data rows taken from the
    full-size launch (1024 row tiles, fused chains) must equal the oracle run on those rows alone."""
    m = make_cdae(FULL['d'], FULL['c'], FULL['H'], FULL['L'], seed=5)
    x, ctx, std, eps = _full_inputs()
    _, loss = m(x, ctx, std=std, scale=1.0, eps=eps)
    torch.cuda.synchronize()
    assert np.isfinite(loss.item())
    pick = [0, 137, 300, 511]
    P64 = {k: v.detach().cpu().numpy().astype(np.float64) for k, v in m.state_dict().items()}
    cs = orc.CdaeSpec(FULL['d'], FULL['c'], FULL['H'], FULL['L'])
    n = lambda a: a[pick].detach().cpu().numpy().astype(np.float64)
    loss_o, g_o, _ = orc.cdae_loss_and_grads(cs, P64, n(x), n(ctx), n(std), n(eps))
    g = m.last_score[pick].cpu().numpy()
    assert rel_err(g, g_o) <= SCORE_TOL and cosine(g, g_o) >= 0.9999
    # the subset's own loss term: mean((sigma*g + eps)^2) over the picked rows
    res = (n(std) * g + n(eps)) ** 2
    assert abs(res.mean() - loss_o) / abs(loss_o) <= LOSS_TOL


def test_full_size_duplicated_halves_give_the_half_batch_gradient():
    """loss = mean over N*d: a batch whose second half repeats the first has the loss and every parameter gradient
    of the half batch.  Compares the B=512 launch (1024 tiles) with the B=256 launch (different tiling, different
    split-K partition of the weight-gradient contractions)."""
    x, ctx, std, eps = _full_inputs(seed=9)
    h = FULL['B'] // 2
    dup = lambda a: torch.cat([a[:h], a[:h]], dim=0).contiguous()
    m = make_cdae(FULL['d'], FULL['c'], FULL['H'], FULL['L'], seed=6)
    m.zero_grad()
    _, l_full = m(dup(x), dup(ctx), std=dup(std), scale=1.0, eps=dup(eps))
    l_full.backward()
    g_full = {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}
    m.zero_grad()
    _, l_half = m(x[:h].contiguous(), ctx[:h].contiguous(), std=std[:h].contiguous(), scale=1.0, eps=eps[:h].contiguous())
    l_half.backward()
    torch.cuda.synchronize()
    assert abs(l_full.item() - l_half.item()) <= 1e-5 * abs(l_half.item())
    for k, p in m.named_parameters():
        if p.grad is None:
            continue
        e = rel_err(g_full[k].cpu().numpy(), p.grad.cpu().numpy())
        assert e <= 2e-4, (k, e)  # same tf32 products, only the fp32 summation order differs
