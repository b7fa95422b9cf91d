import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'pytorch-ardae-vae_b200'), os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')]
import numpy as np, torch
import ardae
from ardae import _lib
from golden_util import rel_err, cosine, load_case, sub
from test_step_gpu import build, t
name = sys.argv[1]
z, meta = load_case(name); hp = meta['hp']
model, cdae, mopt, copt = build(meta, z)
step = ardae.TrainStep(model, cdae, mopt, copt, std_scale=hp['std_scale'], delta=hp['delta'], nz_cdae=hp['nz_cdae'], nstd=hp['nstd'], nz_model=hp['nz_model'])
noise = {k: t(v) for k, v in sub(z, 's0/noise/').items()}
# model update only, without optimizer step: replicate model_update but keep grads
mopt.step_flat = lambda *a, **k: None
sums, g, zz = step.model_update(t(z['s0/x_model']), hp['beta'], noise)
torch.cuda.synchronize()
ar = model._arena
print('loss', sums.cpu().numpy(), float(z['s0/model_loss']), float(z['s0/recon']), float(z['s0/prior']))
ref = sub(z, 's0/model_grads/')
for k, (nme, p) in enumerate(zip(ar.names, ar.params)):
    got = ar.view(ar.stage_flat, k).cpu().numpy()
    print('%-36s rel %.3e cos %.6f |ref| %.3e |got| %.3e' % (nme, rel_err(got, ref[nme]), cosine(got, ref[nme]), np.linalg.norm(ref[nme]), np.linalg.norm(got)))
