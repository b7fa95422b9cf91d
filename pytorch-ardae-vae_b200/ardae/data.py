"""On-device data pipeline pieces: what the reference does on the host per sample / per run.

* `dynamic_binarize`: the `torch.bernoulli` DataLoader transform of dbMNIST (`datasets/mnist.py:39-40,129`) as one
  Philox kernel over a resident batch of grey-level images -- no host work per step.
* `toy_exp4`: the 25-Gaussians data set of `run_vae_25gaussians.sh` (`datasets/toy.py:195-250`, `exp4`): a 5x5 grid of
  means on linspace(-4, 4, 5)^2, variance 0.1, `num_data / 25` points per mixture, generated on the device.
* `MinibatchSampler`: uniform minibatch indices over a resident data set (shuffled epochs, drop_last), the role of
  the reference's DataLoader for data that already lives in HBM.
"""
import ctypes
import math

import torch

from . import _lib


def dynamic_binarize(gray, seed, out=None):
    """x ~ Bernoulli(gray) elementwise; `gray` float32 CUDA tensor with values in [0, 1]."""
    g = _lib.require_cuda(gray, 'gray')
    if out is None:
        out = torch.empty_like(g)
    _lib.check(_lib.lib().ardae_bernoulli(_lib.ptr(g), _lib.ptr(out), ctypes.c_size_t(g.numel()),
                                          ctypes.c_uint64(int(seed) & ((1 << 64) - 1)), _lib.stream_ptr()))
    return out


def toy_exp4(num_data=1000, seed=0, device='cuda'):
    """(x [num_data, 2], label [num_data]) as datasets/toy.py:195-228 builds them."""
    n = 5
    N = n * n
    if num_data % N != 0:
        raise ValueError('num_data should be multiple of {} (num_data = {})'.format(N, num_data))
    lin = torch.linspace(-4.0, 4.0, n, device=device)
    yv, xv = torch.meshgrid(lin, lin, indexing='ij')  # np.meshgrid(x, y): xv varies along columns
    mu = torch.stack([xv.reshape(N), yv.reshape(N)], dim=1)
    per = num_data // N
    eps = torch.empty(num_data, 2, dtype=torch.float32, device=device)
    _lib.check(_lib.lib().ardae_randn(_lib.ptr(eps), ctypes.c_size_t(eps.numel()),
                                      ctypes.c_uint64(int(seed) & ((1 << 64) - 1)), 19, _lib.stream_ptr()))
    label = torch.arange(N, device=device).repeat_interleave(per)
    x = mu[label] + math.sqrt(0.1) * eps
    return x, label


class MinibatchSampler(object):
    """Shuffled epochs of minibatch index tensors over `n` resident samples (drop_last like the reference loaders)."""

    def __init__(self, n, batch_size, seed=0, device='cuda'):
        self.n, self.bs, self.device = int(n), int(batch_size), device
        self.gen = torch.Generator(device=device)
        self.gen.manual_seed(int(seed))
        self._perm, self._pos = None, 0

    def next(self):
        if self._perm is None or self._pos + self.bs > self.n:
            self._perm = torch.randperm(self.n, device=self.device, generator=self.gen)
            self._pos = 0
        idx = self._perm[self._pos:self._pos + self.bs]
        self._pos += self.bs
        return idx
