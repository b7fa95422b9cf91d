"""Pin oracle/ardae_oracle.py against fixtures produced by the reference itself
(oracle/make_golden.py -> tests/golden/*.npz).  CPU only, fp64, tolerance 1e-9 relative."""
import numpy as np
import pytest

import ardae_oracle as orc
from golden_util import CASES, cdae_spec, hp_of, is_lite, load_case, model_dims, num_steps, pick, rel_err, sub



def specs(meta):
    c = meta['cdae']
    spec = orc.ModelSpec(meta['kind'], *model_dims(meta))
    spec.img_c = meta['model'].get('input_channels', 1)
    cs = cdae_spec(meta)
    return spec, cs


@pytest.mark.parametrize('name', CASES)
def test_train_step_matches_reference(name):
    z, meta = load_case(name)
    spec, cs = specs(meta)
    f64 = lambda d: {k: np.asarray(v, dtype=np.float64) for k, v in d.items()}
    Pm, Pc = f64(sub(z, 'm0/')), f64(sub(z, 'c0/'))
    state = {}
    lite = is_lite(name, meta)  # weights / gradients stored as float32
    for step in range(num_steps(name, meta)):
        p = 's%d/' % step
        # x3 weights (saturated, ill-conditioned) + RMSprop's g/(|g|+eps) normalisation amplify the
        # 1e-9 fp64 differences of step 0 to ~5e-6 in step 1; step 0 pins the formulas.
        TOL, GTOL = (2e-5, 1e-4) if (name.endswith('_x3') and step == 1) else ((1e-7, 2e-6) if lite else (1e-7, 1e-6))
        out = orc.train_step(spec, cs, Pm, Pc, z[p + 'x_cdae'], z[p + 'x_model'], sub(z, p + 'noise/'),
                             hp_of(meta), opt_state=state)
        for k in ('zbar', 'z_cdae', 'std', 'cdae_loss', 'cdae_score', 'model_loss', 'recon', 'prior',
                  'z_model', 'entropy_grad'):
            assert rel_err(out[k], z[p + k]) < TOL, (step, k)
        ref_cg = sub(z, p + 'cdae_grads/')
        assert set(ref_cg) == set(out['cdae_grads'])  # incl.: no grad for neglogprob.fc.bias
        for k, v in ref_cg.items():
            assert rel_err(out['cdae_grads'][k], v) < GTOL, (step, 'cdae_grad', k)
        ref_mg = sub(z, p + 'model_grads/')
        assert set(ref_mg) == set(out['model_grads'])
        for k, v in ref_mg.items():
            assert rel_err(pick(z, k, out['model_grads'][k]), v) < GTOL, (step, 'model_grad', k)
        if lite and not meta.get('sampled'):
            continue
        # optimizer semantics: reference Adam (eps placement) and torch RMSprop w/ momentum
        for k, v in sub(z, p + 'm_after/').items():
            assert rel_err(pick(z, k, Pm[k]), v) < (1e-6 if lite else TOL), (step, 'adam', k)
        for k, v in sub(z, p + 'c_after/').items():
            assert rel_err(Pc[k], v) < TOL, (step, 'rmsprop', k)


@pytest.mark.parametrize('name', CASES)
def test_iws_matches_reference(name):
    z, meta = load_case(name)
    spec, _ = specs(meta)
    if is_lite(name, meta):  # IWS on the initial weights (the stepped ones are not stored in full)
        Pm = {k: np.asarray(v, dtype=np.float64) for k, v in sub(z, 'm0/').items()}
    else:
        Pm = sub(z, 's1/m_after/')
    val, per = orc.iws_logprob(spec, Pm, z['iws/x'], z['iws/enc_noise'], z['iws/eta'])
    assert abs(val - float(z['iws/logprob'])) < 1e-8 * max(1.0, abs(val))
    assert per.shape == (meta['iws']['b'],)


def test_known_facts():
    """SURVEY.md 8c 'already-verified facts' the oracle must encode."""
    z, meta = load_case('mnist_small')
    # (ii) neglogprob.fc.bias never receives a gradient
    assert 'neglogprob.fc.bias' not in sub(z, 's0/cdae_grads/')
    # (iv) sigma is signed: std * xi has both signs
    std = z['s0/std'] * z['s0/noise/xi']
    assert (std > 0).any() and (std < 0).any()
    # (v) Adam epsilon placement differs from modern torch Adam at small t
    P = {'w': np.array([1.0])}
    orc.adam_step(P, {'w': np.array([1e-9])}, {}, lr=0.1, beta1=0.5)
    modern = 1.0 - 0.1 * 1e-9 / (1e-9 + 1e-8)
    assert abs(P['w'][0] - modern) > 1e-3


def test_oracle_follows_reference_trajectory():
    """The first iterations of the 400-iteration reference run (oracle/make_curve.py, fp64): the oracle carries the
    optimizer state across steps (Adam bias correction, RMSprop momentum) exactly as the reference does.  The problem
    is chaotic at std_scale 1e4: a 1e-13 relative perturbation of the initial weights grows to 1e-7 by iteration 4 and
    to 1e-2 by iteration 8 (measured with the oracle against itself), so fp64 summation-order differences are pinned
    to 1e-6 over iterations 0-3 and only sanity-checked (10 %) up to iteration 7."""
    import json
    import os
    import curve_util as cu
    from golden_util import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, 'toy_curve.npz'), allow_pickle=True)
    C = json.loads(str(z['meta']))
    hp, B, seed = C['hp'], C['B'], C['seed']
    m, c = C['model'], C['cdae']
    spec = orc.ModelSpec('toy', m['input_dim'], m['noise_dim'], m['h_dim'], m['z_dim'], m['num_hidden_layers'], m['nonlinearity'])
    cs = orc.CdaeSpec(c['input_dim'], c['context_dim'], c['h_dim'], c['num_hidden_layers'])
    f64 = lambda d: {k: np.asarray(v, dtype=np.float64) for k, v in d.items()}
    Pm, Pc, state = f64(sub(z, 'm0/')), f64(sub(z, 'c0/')), {}
    hp_o = dict(hp)
    for t in range(8):
        nz = {k: v.astype(np.float64) for k, v in cu.noise(seed, t, B, m['noise_dim'], m['z_dim'], hp).items()}
        out = orc.train_step(spec, cs, Pm, Pc, cu.batch(seed, t, B, 0).astype(np.float64),
                             cu.batch(seed, t, B, 1).astype(np.float64), nz, hp_o, opt_state=state)
        got = np.array([out['cdae_loss'], out['model_loss'], out['recon'], out['prior'], np.mean(out['std'])], dtype=np.float64)
        assert np.abs(got / z['curve'][t] - 1.0).max() < (1e-6 if t < 4 else 0.1), (t, got, z['curve'][t])
