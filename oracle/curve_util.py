"""Deterministic inputs of the long-run trajectory fixture (tests/golden/toy_curve.npz): the data minibatches and
all injected noise of iteration `t` are functions of (seed, t) alone, so the reference run (oracle/make_curve.py,
build container) and the GPU run (tests/test_step_gpu.py, GPU box) see the same numbers without storing them.
Test infrastructure only (see oracle/ardae_oracle.py's header)."""
import numpy as np

CURVE = dict(
    kind='toy',
    # 25gaussians-shaped (run_vae_25gaussians.sh): relu ToyIPVAE + softplus mlp-grad CDAE L=3, at widths that put the
    # CDAE on the fused chain kernels (H = 128)
    model=dict(input_dim=2, noise_dim=10, h_dim=64, num_hidden_layers=2, nonlinearity='relu', z_dim=2),
    cdae=dict(input_dim=2, context_dim=2, h_dim=128, num_hidden_layers=3, nonlinearity='softplus'),
    B=64, T=400, seed=4242,
    hp=dict(std_scale=10000., delta=0.1, nz_cdae=32, nstd=1, nz_model=1, beta=1.0,
            m_lr=1e-3, m_beta1=0.5, d_lr=1e-4, d_momentum=0.5))


def batch(seed, t, B, which):
    """Minibatch `which` (0: CDAE update, 1: model update) of iteration t: 25 Gaussians on the 5x5 grid {-4..4}^2,
    std 0.05, divided by 2.828 as datasets/toy.py's exp4 does."""
    rng = np.random.RandomState((seed * 1000003 + 2 * t + which) % (2 ** 31 - 1))
    idx = rng.randint(0, 25, size=B)
    centers = np.stack([(idx // 5 - 2) * 2.0, (idx % 5 - 2) * 2.0], axis=1)
    return ((centers + 0.05 * rng.randn(B, 2)) / 2.828).astype(np.float32)


def noise(seed, t, B, n, d, hp):
    rng = np.random.RandomState((seed * 7919 + 104729 * (t + 1)) % (2 ** 31 - 1))
    nz, nstd, nzm = hp['nz_cdae'], hp['nstd'], hp['nz_model']
    return dict(enc_cdae=rng.randn(B * nz, n).astype(np.float32), xi=rng.randn(B, nz * nstd, 1).astype(np.float32),
                eps_cdae=rng.randn(B, nz * nstd, d).astype(np.float32), enc_model=rng.randn(B * nzm, n).astype(np.float32))


def iws_inputs(seed, b, S, n, d):
    rng = np.random.RandomState(seed + 99)
    x = batch(seed, 10 ** 6, b, 0)
    return x, rng.randn(b, S, n).astype(np.float32), rng.randn(b, S, d).astype(np.float32)
