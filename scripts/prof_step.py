"""Profiling driver: N full fused train steps of one bench configuration (for ncu launch lists).

    python scripts/prof_step.py [steps] [config2|config1|config4|aux]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'pytorch-ardae-vae_b200'))
import torch, ardae
sys.path.insert(0, ROOT)
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
name = sys.argv[2] if len(sys.argv) > 2 else 'config2'
c = {'config2': bench.CFG, 'config1': bench.CFG1, 'config4': bench.CFG4, 'aux': bench.CFG_AUX}[name]
torch.manual_seed(1234)
model, cdae, mopt, copt = bench.build_models(c, torch.device('cuda'))
step = ardae.TrainStep(model, cdae, mopt, copt, nz_cdae=c['nz'], ctx_type=c.get('ctx_type', 'lt0'))
B = c['B']
x = (torch.rand(B, c['D'], device='cuda') < 0.13).float() if c['kind'] != 'toy' else torch.randn(B, c['D'], device='cuda')
for i in range(n):
    out = step(x, x, beta=1.0)
torch.cuda.synchronize()
print('launches per step', step.count_launches(B), out['cdae_loss'].item())
