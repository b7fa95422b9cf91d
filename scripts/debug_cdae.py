import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'pytorch-ardae-vae_b200'), os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')]
import numpy as np, torch
import ardae, ardae_oracle as orc
from golden_util import rel_err, cosine
d, c, H, L, B, S = [int(v) for v in (sys.argv[1:7] if len(sys.argv) > 6 else (4, 4, 32, 3, 4, 32))]
torch.manual_seed(0)
m = ardae.MLPGradCARDAE(input_dim=d, context_dim=c, std=1., h_dim=H, num_hidden_layers=L, nonlinearity='softplus').cuda()
P64 = {k: v.detach().cpu().numpy().astype(np.float64) for k, v in m.state_dict().items()}
cs = orc.CdaeSpec(d, c, H, L)
rng = np.random.RandomState(0)
x = rng.randn(B, S, d) * 2; ctx = rng.randn(B, 1, c); std = 0.3 * rng.randn(B, S, 1); eps = rng.randn(B, S, d)
loss_o, g_o, G_o = orc.cdae_loss_and_grads(cs, P64, x, ctx, std, eps)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().cuda()
_, loss = m(t(x), t(ctx), std=t(std), eps=t(eps)); loss.backward(); torch.cuda.synchronize()
print('loss', loss.item(), loss_o, 'score rel', rel_err(m.last_score.cpu().numpy(), g_o))
for k, p in m.named_parameters():
    if p.grad is None: print(k, 'None'); continue
    g = p.grad.cpu().numpy()
    print('%-32s rel %.3e cos %.6f |ref| %.3e |got| %.3e' % (k, rel_err(g, G_o[k]), cosine(g, G_o[k]), np.linalg.norm(G_o[k]), np.linalg.norm(g)))
