"""Config 1 (BASELINE.json configs[0]: 25gaussians, ToyIPVAE mlp-concat z=2 h=256 relu + mlp-grad CDAE h=256 L=3,
batch 512, nz-cdae 256) at full size on one GPU: sanity (finite losses, training signal) + throughput."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'pytorch-ardae-vae_b200'))
import torch, ardae
torch.manual_seed(0)
model = ardae.ToyIPVAE(input_dim=2, noise_dim=10, h_dim=256, num_hidden_layers=2, nonlinearity='relu', enc_type='concat', z_dim=2).cuda()
cdae = ardae.MLPGradCARDAE(input_dim=2, context_dim=2, std=1., h_dim=256, num_hidden_layers=3, nonlinearity='softplus').cuda()
mopt = ardae.Adam(model.parameters(), lr=1e-4, betas=(0.5, 0.999)); copt = ardae.RMSprop(cdae.parameters(), lr=1e-4, momentum=0.5)
step = ardae.TrainStep(model, cdae, mopt, copt, std_scale=10000., delta=0.1, nz_cdae=256, graph=True)
data, _ = ardae.toy_exp4(num_data=50000, seed=1)
sampler = ardae.MinibatchSampler(data.size(0), 512, seed=2)
first = None
for i in range(60):
    o = step(data[sampler.next()], data[sampler.next()], beta=1.0)
    if i == 3:
        first = (o['cdae_loss'].item(), o['model_loss'].item())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(20):
    o = step(data[sampler.next()], data[sampler.next()], beta=1.0)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
last = (o['cdae_loss'].item(), o['model_loss'].item())
assert all(map(lambda v: v == v and abs(v) < 1e9, first + last)), (first, last)
print('config 1: %.3f ms/step = %.0f samples/s; cdae_loss %.4f -> %.4f, model_loss %.3f -> %.3f (80 iterations)' % (
    ms, 512 / ms * 1e3, first[0], last[0], first[1], last[1]))
