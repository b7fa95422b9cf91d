import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'pytorch-ardae-vae_b200'), os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')]
import numpy as np, torch
import ardae, ardae_oracle as orc
from golden_util import rel_err, cosine, load_case, sub
name = sys.argv[1]
z, meta = load_case(name); c = meta['cdae']; hp = meta['hp']
d, cc, H, L = c['input_dim'], c['context_dim'], c['h_dim'], c['num_hidden_layers']
cs = orc.CdaeSpec(d, cc, H, L)
P64 = sub(z, 'c0/')
m = ardae.MLPGradCARDAE(input_dim=d, context_dim=cc, std=1., h_dim=H, num_hidden_layers=L, nonlinearity='softplus')
m.load_state_dict({k: torch.from_numpy(np.asarray(v)).float() for k, v in P64.items()}); m = m.cuda()
x = np.repeat(hp['std_scale'] * (z['s0/z_cdae'] - z['s0/zbar']), hp['nstd'], axis=1)
ctx = z['s0/zbar']; std = z['s0/std'] * z['s0/noise/xi']; eps = z['s0/noise/eps_cdae']
print('x abs max', np.abs(x).max(), 'std abs max', np.abs(std).max())
loss_o, g_o, G_o = orc.cdae_loss_and_grads(cs, P64, x, ctx, std, eps)
f32 = lambda a: np.asarray(a, dtype=np.float32)
loss_f, g_f, G_f = orc.cdae_loss_and_grads(cs, {k: f32(v) for k, v in P64.items()}, f32(x), f32(ctx), f32(std), f32(eps))
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().cuda()
_, loss = m(t(x), t(ctx), std=t(std), eps=t(eps)); loss.backward(); torch.cuda.synchronize()
print('loss', loss.item(), loss_o, loss_f, 'score rel cuda %.3e fp32-oracle %.3e' % (rel_err(m.last_score.cpu().numpy(), g_o), rel_err(g_f, g_o)))
for k, p in m.named_parameters():
    if p.grad is None: print(k, 'None'); continue
    g = p.grad.cpu().numpy()
    print('%-32s cuda rel %.3e cos %.6f | fp32-oracle rel %.3e | |ref| %.3e' % (k, rel_err(g, G_o[k]), cosine(g, G_o[k]), rel_err(G_f[k], G_o[k]), np.linalg.norm(G_o[k])))
