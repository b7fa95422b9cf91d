"""Config 5 (BASELINE.json configs[4]) throughput probe: IWS log-likelihood with 5000 importance samples per image on
MNIST-shape images (MNISTIPVAE z=32 h=300 n=100), images/s on one GPU; under torchrun the images are sharded by rank
(ardae.evaluate_iws allreduces the partial sums).  Prints one JSON line."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'pytorch-ardae-vae_b200')); sys.path.insert(0, ROOT)
import torch, ardae
from bench import CFG as c

S = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
n_img = int(sys.argv[2]) if len(sys.argv) > 2 else 256
bs = int(sys.argv[3]) if len(sys.argv) > 3 else 16
torch.manual_seed(1234)
model = ardae.MNISTIPVAE(input_dim=c['D'], noise_dim=c['n'], h_dim=c['h'], num_hidden_layers=c['model_layers'],
                         nonlinearity=c['nonlin'], z_dim=c['z']).cuda()
x = (torch.rand(n_img, c['D'], device='cuda') < 0.13).float()
v = ardae.evaluate_iws(x[:2 * bs], model, S, batch_size=bs)  # warm-up (plans)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
v = ardae.evaluate_iws(x, model, S, batch_size=bs)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
rows = n_img * S
dec_macs = c['z'] * c['h'] + 2 * c['h'] * c['h'] + c['h'] * c['D']
enc_macs = c['n'] * c['h'] + c['h'] * c['z']
print(json.dumps(dict(metric='iws_images_per_sec', value=n_img / ms * 1e3, unit='images/s', iws_samples=S, images=n_img,
                      images_per_call=bs, ms=ms, logprob=float(v.item()),
                      algorithmic_tflops=2.0 * rows * (dec_macs + enc_macs) / ms * 1e-9,
                      reference_cpu_images_per_sec=1 / 0.076,
                      note='10k images at this rate: %.0f s (reference CPU loop: ~760 s, SURVEY 8a-9)' % (10000 / (n_img / ms * 1e3)))))
