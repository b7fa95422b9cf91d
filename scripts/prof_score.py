"""Per-launch timing of the B-row score plan (glogprob of the model minibatch with the updated CDAE): the 0.23 ms that
sit on the critical path of every iteration.  python scripts/prof_score.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'pytorch-ardae-vae_b200'))
import torch, ardae
from ardae import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
torch.manual_seed(0)
cdae = ardae.MLPGradCARDAE(input_dim=32, context_dim=32, std=1., h_dim=256, num_hidden_layers=5, nonlinearity='softplus').cuda()
x = torch.randn(B, 1, 32, device='cuda') * 1e4
ctx = torch.randn(B, 1, 32, device='cuda')
std = torch.zeros(B, 1, 1, device='cuda')
for _ in range(3):
    g = cdae.glogprob(x, ctx, std=std, scale=1e4)
h = cdae._plan(B, 1, False)
_lib.check(_lib.lib().ardae_cdae_set_profile(h, 1))
for _ in range(3):
    g = cdae.glogprob(x, ctx, std=std, scale=1e4)
torch.cuda.synchronize()
ops = _lib.read_cdae_profile(h)
tot = 0.0
for tag, ms in ops:
    print('%-20s %8.1f us' % (tag, ms * 1e3)); tot += ms
print('sum %.1f us over %d launches' % (tot * 1e3, len(ops)))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
_lib.check(_lib.lib().ardae_cdae_set_profile(h, 0))
e0.record()
for _ in range(20):
    g = cdae.glogprob(x, ctx, std=std, scale=1e4)
e1.record(); torch.cuda.synchronize()
print('glogprob eager: %.1f us per call' % (e0.elapsed_time(e1) / 20 * 1e3))
