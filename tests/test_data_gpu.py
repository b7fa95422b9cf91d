"""On-device data pipeline: distributional checks against the reference's definitions
(datasets/mnist.py:39-40 torch.bernoulli transform; datasets/toy.py:195-228 exp4)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_dynamic_binarize_is_bernoulli():
    import ardae
    g = torch.rand(512, 784, device='cuda')
    g[:, :10] = 0.0
    g[:, 10:20] = 1.0
    x = ardae.dynamic_binarize(g, seed=1)
    assert set(torch.unique(x).tolist()) <= {0.0, 1.0}
    assert float(x[:, :10].sum()) == 0.0 and float(x[:, 10:20].mean()) == 1.0
    # E[x] = g: average over many draws of the same image
    p = torch.linspace(0.05, 0.95, 784, device='cuda').repeat(4096, 1)
    m = ardae.dynamic_binarize(p, seed=2).mean(0)
    assert float((m - p[0]).abs().max()) < 4.5 * 0.5 / np.sqrt(4096)
    # different seeds -> different masks, same seed -> same mask
    a, b, c = ardae.dynamic_binarize(g, seed=3), ardae.dynamic_binarize(g, seed=3), ardae.dynamic_binarize(g, seed=4)
    assert torch.equal(a, b) and not torch.equal(a, c)


def test_toy_exp4_matches_the_reference_definition():
    import ardae
    x, label = ardae.toy_exp4(num_data=25 * 4000, seed=5)
    assert x.shape == (100000, 2) and label.shape == (100000,)
    lin = np.linspace(-4, 4, 5)
    xv, yv = np.meshgrid(lin, lin)
    mu = np.stack([xv.reshape(-1), yv.reshape(-1)], axis=1)  # datasets/toy.py:209-215
    xs, ls = x.cpu().numpy(), label.cpu().numpy()
    for i in range(25):
        pts = xs[ls == i]
        assert pts.shape[0] == 4000
        assert np.abs(pts.mean(0) - mu[i]).max() < 5 * np.sqrt(0.1 / 4000)
        assert np.abs(pts.var(0) - 0.1).max() < 0.01
    with pytest.raises(ValueError):
        ardae.toy_exp4(num_data=1001)
    s = ardae.MinibatchSampler(1000, 128, seed=0)
    seen = torch.cat([s.next() for _ in range(7)])
    assert seen.numel() == 896 and seen.unique().numel() == 896  # one epoch: no repeats, drop_last
