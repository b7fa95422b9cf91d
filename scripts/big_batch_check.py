"""Config 3 shard sizes (batch 4096 over 2/4/8 GPUs -> 2048/1024/512 data rows per GPU): one GPU, per-GPU batch B."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'pytorch-ardae-vae_b200')); sys.path.insert(0, ROOT)
import torch, ardae
from bench import CFG as c
for B in [int(a) for a in sys.argv[1:]] or [1024, 2048]:
    torch.manual_seed(1234)
    model = ardae.MNISTIPVAE(input_dim=c['D'], noise_dim=c['n'], h_dim=c['h'], num_hidden_layers=c['model_layers'], nonlinearity=c['nonlin'], z_dim=c['z']).cuda()
    cdae = ardae.MLPGradCARDAE(input_dim=c['z'], context_dim=c['z'], std=1., h_dim=c['cdae_h'], num_hidden_layers=c['cdae_L'], nonlinearity='softplus').cuda()
    mopt = ardae.Adam(model.parameters(), lr=1e-4, betas=(0.5, 0.999)); copt = ardae.RMSprop(cdae.parameters(), lr=1e-4, momentum=0.5)
    step = ardae.TrainStep(model, cdae, mopt, copt, nz_cdae=c['nz'], graph=True)
    x = (torch.rand(B, 784, device='cuda') < 0.13).float()
    for i in range(5):
        o = step(x, x, beta=1.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10):
        o = step(x, x, beta=1.0)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print('B=%d per GPU: %.3f ms/step, %.0f samples/s, cdae_loss %.4f model_loss %.2f, peak mem %.1f GB' % (
        B, ms, B / ms * 1e3, o['cdae_loss'].item(), o['model_loss'].item(), torch.cuda.max_memory_allocated() / 1e9))
    del step, model, cdae, mopt, copt
    torch.cuda.empty_cache()
