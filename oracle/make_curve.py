"""Generates tests/golden/toy_curve.npz: T = 400 iterations of the REFERENCE's own training step
(ivae_ardae.py:707-846 through oracle/ref_harness.ref_train_step, the unmodified reference modules in fp64 on the CPU)
on the 25-Gaussians toy problem, with data and noise from oracle/curve_util.py.  Stored: initial weights, the four
losses of every iteration, the sigma scale, the IWS log-likelihood of a fixed batch under the final weights.
The GPU test replays the same iterations through ardae.TrainStep and compares the trajectories.

    python oracle/make_curve.py        (build container only: needs /root/reference)
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import curve_util as cu  # noqa: E402
import ref_harness as rh  # noqa: E402


def run(dtype, init=None):
    C = cu.CURVE
    hp, B, T, seed = C['hp'], C['B'], C['T'], C['seed']
    model, cdae = rh.build_reference('toy', C['model'], C['cdae'], dtype=dtype, seed=11)
    mopt, copt = rh.build_optimizers(model, cdae, hp)
    if init is None:
        init = {}
        for k, v in model.state_dict().items():
            init['m0/' + k] = v.numpy().astype(np.float32)
        for k, v in cdae.state_dict().items():
            init['c0/' + k] = v.numpy().astype(np.float32)
    # the stored initial weights are float32: every run starts from exactly those values
    npdt = np.float64 if dtype == torch.float64 else np.float32
    model.load_state_dict({k: torch.from_numpy(init['m0/' + k].astype(npdt)) for k in model.state_dict()})
    cdae.load_state_dict({k: torch.from_numpy(init['c0/' + k].astype(npdt)) for k in cdae.state_dict()})
    n, d = C['model']['noise_dim'], C['model']['z_dim']
    curve = np.zeros((T, 5))
    tt = lambda a: torch.from_numpy(np.asarray(a, dtype=npdt))
    for t in range(T):
        nz = {k: tt(v) for k, v in cu.noise(seed, t, B, n, d, hp).items()}
        o = rh.ref_train_step(model, cdae, mopt, copt, tt(cu.batch(seed, t, B, 0)), tt(cu.batch(seed, t, B, 1)), nz, hp,
                              mnist_like=False)
        curve[t] = [float(o['cdae_loss']), float(o['model_loss']), float(o['recon']), float(o['prior']),
                    float(o['std'].mean())]
        if t % 100 == 0:
            print(dtype, t, curve[t])
    x, en, eta = cu.iws_inputs(seed, 64, 64, n, d)
    return init, curve, float(rh.ref_iws(model, tt(x), tt(en), tt(eta)))


def main():
    # the fp64 run is the trajectory; the reference's own fp32 run (what a user of the reference actually executes)
    # measures how far rounding alone moves a trajectory of this problem -- the yardstick of the GPU test
    out, curve, lp = run(torch.float64)
    _, curve32, lp32 = run(torch.float32, init=out)
    out['curve'], out['iws_logprob'] = curve, lp
    out['curve_ref_fp32'], out['iws_logprob_ref_fp32'] = curve32, lp32
    out['meta'] = json.dumps(cu.CURVE)
    np.savez_compressed(os.path.join(HERE, '..', 'tests', 'golden', 'toy_curve.npz'), **out)
    dev = np.abs(curve32 / curve - 1.0)
    print('final', curve[-1], 'iws', lp, lp32)
    print('reference fp32 vs fp64: max rel first 25', dev[:25].max(0), ' windows',
          np.abs(curve32.reshape(-1, 25, 5).mean(1) / curve.reshape(-1, 25, 5).mean(1) - 1).max(0))


if __name__ == '__main__':
    main()
