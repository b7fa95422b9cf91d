"""Profiling driver: 3 full fused train steps at config 2 (for ncu launch lists)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'pytorch-ardae-vae_b200'))
import torch, ardae
sys.path.insert(0, ROOT)
from bench import CFG as c
torch.manual_seed(1234)
model = ardae.MNISTIPVAE(input_dim=c['D'], noise_dim=c['n'], h_dim=c['h'], num_hidden_layers=c['model_layers'], nonlinearity=c['nonlin'], z_dim=c['z']).cuda()
cdae = ardae.MLPGradCARDAE(input_dim=c['z'], context_dim=c['z'], std=1., h_dim=c['cdae_h'], num_hidden_layers=c['cdae_L'], nonlinearity='softplus').cuda()
mopt = ardae.Adam(model.parameters(), lr=1e-4, betas=(0.5, 0.999)); copt = ardae.RMSprop(cdae.parameters(), lr=1e-4, momentum=0.5)
step = ardae.TrainStep(model, cdae, mopt, copt, nz_cdae=c['nz'])
x = (torch.rand(512, 784, device='cuda') < 0.13).float()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for i in range(n):
    out = step(x, x, beta=1.0)
torch.cuda.synchronize()
print('launches per step', step.count_launches(512), out['cdae_loss'].item())
