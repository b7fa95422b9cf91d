"""How much of a fixture's SECOND-step gradient error is inherited from the first step's update error?

The GPU step matches step 0 of every fixture to <= 2e-2 per gradient tensor, and its parameter UPDATE to a few per
cent (Adam / RMSprop normalise by |g|, so tf32-level gradient error becomes an O(lr) update error).  Step 1 then starts
from slightly different parameters.  This script measures that inheritance with the fp64 oracle alone: run step 0
exactly, perturb the update of every tensor by `rel` of its norm in a random direction, run step 1 and compare its
gradients with the fixture.  For `auxmnist_small` (std_scale 1e4 multiplies every change of the encoder into the CDAE
inputs) a 4-5 % update perturbation moves the step-1 CDAE gradients by 1.5e-2 .. 3.6e-2 -- the tolerance
tests/test_step_gpu.py grants that case at step 1.

    python scripts/step1_sensitivity.py auxmnist_small
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')]
import ardae_oracle as orc  # noqa: E402
from golden_util import hp_of, load_case, rel_err, sub  # noqa: E402
from test_oracle_golden import specs  # noqa: E402


def main(name, rel_model=0.05, rel_cdae=0.04, trials=3):
    z, meta = load_case(name)
    spec, cs = specs(meta)
    f64 = lambda d: {k: np.asarray(v, dtype=np.float64) for k, v in d.items()}
    for trial in range(trials):
        rng = np.random.RandomState(trial)
        Pm, Pc = f64(sub(z, 'm0/')), f64(sub(z, 'c0/'))
        state = {}
        orc.train_step(spec, cs, Pm, Pc, z['s0/x_cdae'], z['s0/x_model'], sub(z, 's0/noise/'), hp_of(meta), opt_state=state)
        for P, P0, rel in ((Pm, f64(sub(z, 'm0/')), rel_model), (Pc, f64(sub(z, 'c0/')), rel_cdae)):
            for k in P:
                u = P[k] - P0[k]
                nrm = np.linalg.norm(u)
                if nrm > 0:
                    e = rng.randn(*u.shape)
                    P[k] = P[k] + e * (rel * nrm / np.linalg.norm(e))
        out = orc.train_step(spec, cs, Pm, Pc, z['s1/x_cdae'], z['s1/x_model'], sub(z, 's1/noise/'), hp_of(meta),
                             opt_state=state)
        errs = {k: rel_err(out['cdae_grads'][k], v) for k, v in sub(z, 's1/cdae_grads/').items()}
        worst = max(errs, key=errs.get)
        print('trial %d: worst step-1 cdae gradient %s %.2e, median %.2e, entropy_grad %.2e'
              % (trial, worst, errs[worst], float(np.median(list(errs.values()))),
                 rel_err(out['entropy_grad'], z['s1/entropy_grad'])))


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'auxmnist_small')
