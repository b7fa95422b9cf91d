"""Profiling driver: two CDAE train calls at config-2 size (66 GEMM launches each)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'pytorch-ardae-vae_b200'))
import torch, ardae
torch.manual_seed(0)
B, S, d, H, L = 512, 256, 32, 256, 5
if len(sys.argv) > 1:
    L = int(sys.argv[1])
c = ardae.MLPGradCARDAE(input_dim=d, context_dim=d, std=1., h_dim=H, num_hidden_layers=L, nonlinearity='softplus').cuda()
x = torch.randn(B, S, d, device='cuda') * 100; ctx = torch.randn(B, 1, d, device='cuda'); std = torch.randn(B, S, 1, device='cuda') * 10
for i in range(2):
    _, loss = c(x, ctx, std=std, seed=1)
torch.cuda.synchronize()
print('loss', loss.item())
