"""Graph-replay check: (1) a graph-mode TrainStep follows the eager one (same seeds are impossible across modes, so compare
loss statistics and parameter drift), (2) timing of eager vs graph, device-resident and host-fed (e2e)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'pytorch-ardae-vae_b200')); sys.path.insert(0, ROOT)
import torch, ardae
from bench import CFG as c

def build(graph):
    torch.manual_seed(1234)
    model = ardae.MNISTIPVAE(input_dim=c['D'], noise_dim=c['n'], h_dim=c['h'], num_hidden_layers=c['model_layers'], nonlinearity=c['nonlin'], z_dim=c['z']).cuda()
    cdae = ardae.MLPGradCARDAE(input_dim=c['z'], context_dim=c['z'], std=1., h_dim=c['cdae_h'], num_hidden_layers=c['cdae_L'], nonlinearity='softplus').cuda()
    mopt = ardae.Adam(model.parameters(), lr=1e-4, betas=(0.5, 0.999)); copt = ardae.RMSprop(cdae.parameters(), lr=1e-4, momentum=0.5)
    return model, cdae, mopt, copt, ardae.TrainStep(model, cdae, mopt, copt, nz_cdae=c['nz'], graph=graph)

x = (torch.rand(512, 784, device='cuda') < 0.13).float()
xh = x.cpu().pin_memory()
res = {}
for graph in (False, True):
    model, cdae, mopt, copt, step = build(graph)
    losses = []
    for i in range(12):
        o = step(x, x, beta=1.0)
        losses.append((o['cdae_loss'].item(), o['model_loss'].item()))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        o = step(x, x, beta=1.0)
    e1.record(); torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1) / 20
    xb = torch.empty_like(x); hl = torch.empty(2, pin_memory=True)
    t0 = time.perf_counter()
    for i in range(20):
        xb.copy_(xh, non_blocking=True)
        o = step(xb, xb, beta=1.0)
        hl.copy_(torch.cat([o['cdae_loss'], o['model_loss']]), non_blocking=True)
        torch.cuda.current_stream().synchronize()
    e2e_ms = (time.perf_counter() - t0) / 20 * 1e3
    p = torch.cat([q.detach().flatten() for q in model.parameters()]).double()
    st = mopt.state[next(iter(model.parameters()))]['step']
    res[graph] = (losses, p.norm().item(), st)
    print('graph=%s: device-resident %.3f ms/step, e2e %.3f ms/step, adam step %d, |params| %.6f' % (graph, dev_ms, e2e_ms, st, p.norm().item()))
    print('   losses', ['%.4f/%.2f' % l for l in losses[:3]], '...', ['%.4f/%.2f' % l for l in losses[-3:]])
a, b = res[False], res[True]
assert a[2] == b[2], 'host step counters diverged'
for (ca, ma), (cb, mb) in zip(a[0], b[0]):
    assert abs(ca - cb) < 0.05 * abs(ca) + 1e-3 and abs(ma - mb) < 0.02 * abs(ma), ((ca, ma), (cb, mb))
assert abs(a[1] - b[1]) < 1e-3 * a[1]
print('graph replay OK')
