#!/usr/bin/env python
"""Headline benchmark: AR-DAE VAE train samples/sec (BASELINE.json) on config 2
(dbMNIST-shape MNISTIPVAE z=32 h=300 n=100 + mlp-grad CDAE h=256 L=5, batch 512 per GPU, nz-cdae 256).

  python bench.py --gpus N --steps K --warmup W                (N > 1: launched by torch.distributed.run)
  python bench.py --impl reference ...                         (the reference's CPU path, host cores)

A step = num_cdae_updates (1) CDAE updates + 1 model update = the unit the reference logs as ms/step
(ivae_ardae.py:859,876); samples/sec = global train batch x steps/sec.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, 'pytorch-ardae-vae_b200'), os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

CFG = dict(  # BASELINE.json configs[1]; hyper-parameters from run_vae_dbmnist.sh:37
    kind='mnist', D=784, n=100, h=300, z=32, model_layers=2, nonlin='softplus',
    cdae_h=256, cdae_L=5, B=512, nz=256, nstd=1, nz_model=1, std_scale=10000., delta=0.1, beta=1.0,
    m_lr=1e-4, m_beta1=0.5, d_lr=1e-4, d_momentum=0.5)
WORKLOAD = 'configs[1] dbMNIST ivae_ardae MLP z=32 + mlp-grad CDAE nz 256, batch 512 per GPU'  # (< 120 chars)
WORKLOAD_DETAIL = ('MNISTIPVAE mlp-concat D=784 h=300 n=100 z=32 (2 hidden layers, softplus) + MLPGradCARDAE h=256 L=5, '
                   'train-nz-cdae 256, nstd 1, std-scale 1e4, delta 0.1, Adam(model, b1 0.5) + RMSprop(cdae, momentum 0.5), '
                   'lr 1e-4 (run_vae_dbmnist.sh:37)')
CFG1 = dict(  # BASELINE.json configs[0]: 25gaussians README command (run_vae_25gaussians.sh)
    kind='toy', D=2, n=10, h=256, z=2, model_layers=2, nonlin='relu',
    cdae_h=256, cdae_L=3, B=512, nz=256, nstd=1, nz_model=1, std_scale=10000., delta=0.1, beta=1.0,
    m_lr=1e-4, m_beta1=0.5, d_lr=1e-4, d_momentum=0.5)
CFG4 = dict(  # BASELINE.json configs[3]: dbMNIST conv implicit encoder / decoder (run_vae_dbmnist.sh:31), 8192 / 8 GPUs
    kind='conv', D=784, n=100, h=800, z=32, model_layers=0, nonlin='softplus',
    cdae_h=256, cdae_L=5, B=1024, nz=256, nstd=1, nz_model=1, std_scale=10000., delta=0.1, beta=1.0,
    m_lr=1e-4, m_beta1=0.5, d_lr=1e-4, d_momentum=0.5)


CFG_AUX = dict(  # run_vae_dbmnist.sh:28,34,40: hierarchical MNISTAuxIPVAE + the 'hidden1a' CDAE context
    kind='auxmnist', D=784, n=100, h=300, z=32, model_layers=2, nonlin='softplus', ctx_type='hidden1a', ctx_dim=600,
    cdae_h=256, cdae_L=5, B=512, nz=256, nstd=1, nz_model=1, std_scale=10000., delta=0.1, beta=1.0,
    m_lr=1e-4, m_beta1=0.5, d_lr=1e-4, d_momentum=0.5)


def cdae_alg_flops(B, nz, d, c, H, L):
    """6 sweeps x 2 x N x G + context branch on the B distinct rows."""
    N = B * nz
    G = (2 * L - 1) * H * H + d * H + H
    ctx = c * H + (L - 1) * H * H + H * H
    return 6 * 2 * N * G + 3 * 2 * B * ctx


def step_alg_flops(c, B):
    """SURVEY 8d whole-step figure: CDAE update + N-row encoder sampling (as executed) + model update."""
    cd = cdae_alg_flops(B, c['nz'] * c['nstd'], c['z'], c['z'], c['cdae_h'], c['cdae_L'])
    if c['kind'] == 'mnist':
        enc = 2 * B * c['nz'] * ((c['h'] + c['n']) * c['h'] + c['h'] * c['z'])
        mod = 3 * 2 * B * (c['D'] * c['h'] + 3 * c['h'] * c['h'] + (c['h'] + c['n']) * c['h'] + c['h'] * c['z']
                           + c['z'] * c['h'] + 2 * c['h'] * c['h'] + c['h'] * c['D'])
    else:
        enc, mod = 0, 0
    return cd + enc + mod


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
                  'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([v.strip() for v in out.strip().split(',')])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = max([int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()] or [0])
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith('active')})
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=mx or None, reasons=reasons, samples=len(sm))


def synth_batch(gen, p, B, device):
    import torch
    return torch.bernoulli(p.expand(B, -1), generator=gen).to(device)


def cpu_reference_leg(steps, warmup, B_sample=int(os.environ.get('ARDAE_BENCH_CPU_ROWS', '128'))):
    """The reference's CPU path on the host cores, bounded sample of the same workload.
    kind 'reference': the reference's own modules (when its tree is reachable: build container);
    kind 'port': the numpy oracle restatement (GPU box: the Python reference cannot travel)."""
    import numpy as np
    import torch
    import ref_harness as rh
    c = CFG
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    try:  # numpy's BLAS pool too (a launcher may have pinned it to one thread)
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=cores)
    except Exception:
        pass
    hp = dict(std_scale=c['std_scale'], delta=c['delta'], nz_cdae=c['nz'], nstd=c['nstd'], nz_model=c['nz_model'],
              beta=c['beta'], m_lr=c['m_lr'], m_beta1=c['m_beta1'], d_lr=c['d_lr'], d_momentum=c['d_momentum'])
    g = torch.Generator().manual_seed(1234)
    pix = torch.rand(1, c['D'], generator=g) * 0.26
    B, N = B_sample, B_sample * c['nz']

    def draw():
        return dict(enc_cdae=torch.randn(N, c['n'], generator=g), xi=torch.randn(B, c['nz'] * c['nstd'], 1, generator=g),
                    eps_cdae=torch.randn(B, c['nz'] * c['nstd'], c['z'], generator=g),
                    enc_model=torch.randn(B * c['nz_model'], c['n'], generator=g))
    times = []
    if rh.find_reference() is not None:
        kind = 'reference'
        model, cdae = rh.build_reference('mnist', dict(input_dim=c['D'], noise_dim=c['n'], h_dim=c['h'],
                                                       num_hidden_layers=c['model_layers'], nonlinearity=c['nonlin'],
                                                       z_dim=c['z']),
                                         dict(input_dim=c['z'], context_dim=c['z'], h_dim=c['cdae_h'],
                                              num_hidden_layers=c['cdae_L'], nonlinearity='softplus'), seed=1234)
        mopt, copt = rh.build_optimizers(model, cdae, hp)
        for i in range(warmup + steps):
            xc, xm, nz_ = torch.bernoulli(pix.expand(B, -1), generator=g), torch.bernoulli(pix.expand(B, -1), generator=g), draw()
            t0 = time.perf_counter()
            rh.ref_train_step(model, cdae, mopt, copt, xc, xm, nz_, hp, do_step=True)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    else:
        kind = 'port'
        import ardae_oracle as orc
        spec = orc.ModelSpec('mnist', c['D'], c['n'], c['h'], c['z'], c['model_layers'], c['nonlin'])
        cs = orc.CdaeSpec(c['z'], c['z'], c['cdae_h'], c['cdae_L'])
        rng = np.random.RandomState(1234)

        def lin(o, i):
            k = 1.0 / np.sqrt(i)
            return (rng.uniform(-k, k, (o, i)).astype(np.float32), rng.uniform(-k, k, (o,)).astype(np.float32))
        Pm, Pc = {}, {}
        dims_m = [('encode.inp_encode.layers.0', 300, 784), ('encode.inp_encode.layers.1', 300, 300),
                  ('encode.inp_encode.layers.2', 300, 300), ('encode.inp_encode.fc', 300, 300),
                  ('encode.fc.layers.0', 300, 400), ('encode.fc.fc', 32, 300), ('decode.main.layers.0', 300, 32),
                  ('decode.main.layers.1', 300, 300), ('decode.main.fc', 300, 300), ('decode.reparam.logit_fn', 784, 300)]
        for k, o, i in dims_m:
            Pm[k + '.weight'], Pm[k + '.bias'] = lin(o, i)
        H, L_, d = c['cdae_h'], c['cdae_L'], c['z']
        for pre, first in (('ctx_encode', d), ('inp_encode', d)):
            for k in orc.mlp_keys(pre, L_ - 1):
                Pc[k + '.weight'], Pc[k + '.bias'] = lin(H, first if k.endswith('layers.0') else H)
        for k in orc.mlp_keys('neglogprob', L_):
            o, i = (1, H) if k.endswith('.fc') else (H, 2 * H + 1 if k.endswith('layers.0') else H)
            Pc[k + '.weight'], Pc[k + '.bias'] = lin(o, i)
        state = {}
        for i in range(warmup + steps):
            xc = torch.bernoulli(pix.expand(B, -1), generator=g).numpy()
            xm = torch.bernoulli(pix.expand(B, -1), generator=g).numpy()
            nz_ = {k: v.numpy() for k, v in draw().items()}
            t0 = time.perf_counter()
            orc.train_step(spec, cs, Pm, Pc, xc, xm, nz_, hp, opt_state=state)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    sample = ('%d of 512 data rows x nz=256 (N=%d CDAE rows) per step, fp32, %s' % (
        B_sample, N, 'reference modules via oracle/ref_harness.py' if kind == 'reference' else 'numpy oracle port'))
    return dict(value=B_sample / sec, unit='samples/s', cores=cores, kind=kind, sample=sample), sec


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to its workers: give the CPU arm every host core back (before numpy / torch
    # load their BLAS) -- "all the host threads it can use"
    cores = str(os.cpu_count() or 1)
    for k in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[k] = cores
    # every step is a bounded sample (128 of the 512 data rows): ~3 s of host work, so the driver's own --steps /
    # --warmup are honoured as given (25 steps ~ 1.5 minutes)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    cb, sec = cpu_reference_leg(steps, warmup)
    line = dict(metric='train_samples_per_sec', value=cb['value'], unit='samples/s', impl='reference',
                n_gpus=args.gpus, steps=steps, warmup=warmup, ms_per_step=sec * 1e3, higher_is_better=True,
                scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
                config=dict(workload=WORKLOAD, detail=WORKLOAD_DETAIL, note='CPU, bounded sample: ' + cb['sample']),
                cpu_baseline=cb,
                e2e=dict(value=cb['value'], unit='samples/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def build_models(c, dev, cdae_kind='grad'):
    import ardae
    if c['kind'] == 'toy':
        model = ardae.ToyIPVAE(input_dim=c['D'], noise_dim=c['n'], h_dim=c['h'], num_hidden_layers=c['model_layers'],
                               nonlinearity=c['nonlin'], enc_type='concat', z_dim=c['z']).to(dev)
    elif c['kind'] == 'conv':
        model = ardae.ConvIPVAE(input_height=28, input_channels=1, z_dim=c['z'], noise_dim=c['n'],
                                nonlinearity=c['nonlin']).to(dev)
    elif c['kind'] == 'auxmnist':
        model = ardae.MNISTAuxIPVAE(input_dim=c['D'], noise_dim=c['n'], h_dim=c['h'], num_hidden_layers=c['model_layers'],
                                    nonlinearity=c['nonlin'], enc_type='simple', z_dim=c['z']).to(dev)
    else:
        model = ardae.MNISTIPVAE(input_dim=c['D'], noise_dim=c['n'], h_dim=c['h'], num_hidden_layers=c['model_layers'],
                                 nonlinearity=c['nonlin'], enc_type='concat', z_dim=c['z']).to(dev)
    cdae_cls = ardae.MLPGradCARDAE if cdae_kind == 'grad' else ardae.MLPResCARDAE
    cdae = cdae_cls(input_dim=c['z'], context_dim=c.get('ctx_dim', c['z']), std=1., h_dim=c['cdae_h'],
                    num_hidden_layers=c['cdae_L'], nonlinearity='softplus').to(dev)
    mopt = ardae.Adam(model.parameters(), lr=c['m_lr'], betas=(c['m_beta1'], 0.999))
    copt = ardae.RMSprop(cdae.parameters(), lr=c['d_lr'], momentum=c['d_momentum'])
    return model, cdae, mopt, copt


class Harness(object):
    """One training configuration on this rank: models, fused step, synthetic resident / pinned minibatches."""

    def __init__(self, c, B, dev, world, rank, graph=True, cdae_kind='grad', nb=8):
        import torch
        import torch.distributed as dist
        import ardae
        self.c, self.B, self.dev, self.world, self.rank = c, B, dev, world, rank
        torch.manual_seed(1234)  # same weights on every rank (TrainStep also broadcasts rank 0's replica)
        self.model, self.cdae, self.mopt, self.copt = build_models(c, dev, cdae_kind)
        self.step = ardae.TrainStep(self.model, self.cdae, self.mopt, self.copt, std_scale=c['std_scale'],
                                    delta=c['delta'], nz_cdae=c['nz'], nstd=c['nstd'], nz_model=c['nz_model'],
                                    process_group=dist.group.WORLD if world > 1 else None, seed=1234, graph=graph,
                                    ctx_type=c.get('ctx_type', 'lt0'))
        gen = torch.Generator().manual_seed(999 + rank)
        if c['kind'] == 'toy':   # 25-Gaussians mixture, generated on the device (ardae.toy_exp4)
            data, _ = ardae.toy_exp4(num_data=50000, seed=1 + rank, device=dev)
            idx = [torch.randint(0, data.size(0), (B,), generator=gen) for _ in range(2 * nb)]
            self.host = [(data[idx[2 * i].to(dev)].cpu().pin_memory(), data[idx[2 * i + 1].to(dev)].cpu().pin_memory())
                         for i in range(nb)]
        else:                    # dbMNIST-shape: x ~ Bernoulli(p), fixed per-pixel p (mean ink fraction ~0.13)
            pix = torch.rand(1, c['D'], generator=torch.Generator().manual_seed(5)) * 0.26
            self.host = [(torch.bernoulli(pix.expand(B, -1), generator=gen).pin_memory(),
                          torch.bernoulli(pix.expand(B, -1), generator=gen).pin_memory()) for _ in range(nb)]
        self.resident = [(a.to(dev), b.to(dev)) for a, b in self.host]
        self.nb = nb

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        import torch
        import torch.distributed as dist
        t = torch.tensor([ms], device=self.dev)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def time_resident(self, W, K):
        """K iterations on device-resident minibatches (CUDA-graph replays once captured); ms total, max over ranks."""
        import torch
        c = self.c
        for i in range(max(W, 4)):  # >= 2 eager iterations + capture + 1 replay
            self.step(*self.resident[i % self.nb], beta=c['beta'])
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        for i in range(K):
            out = self.step(*self.resident[i % self.nb], beta=c['beta'])
        e1.record()
        self.barrier()
        ms = self.max_over_ranks(e0.elapsed_time(e1))
        return ms, {k: float(out[k].item()) for k in ('cdae_loss', 'model_loss')}

    def time_e2e(self, K):
        """The same iterations fed from pinned host memory through the public call: every step copies both minibatches
        host -> device (TrainStep.stage: a copy stream, double-buffered, so the copy of step i+1 runs underneath step i)
        and reads the four losses device -> host, synchronised every step (the caller reads the losses)."""
        import torch
        c = self.c
        hloss = torch.empty(4, pin_memory=True)
        for i in range(2):
            self.step(*self.resident[i % self.nb], beta=c['beta'])
        self.barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        self.step.stage(*self.host[0])
        for i in range(K):
            o = self.step(beta=c['beta'])
            if i + 1 < K:
                self.step.stage(*self.host[(i + 1) % self.nb])
            hloss.copy_(o['losses'], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        f1.record()
        self.barrier()
        h2d = 2 * self.resident[0][0].numel() * 4
        return self.max_over_ranks(f0.elapsed_time(f1)), h2d

    def check_comm(self):
        """Raises if a fused data-parallel step ever timed out waiting for a peer (the numbers would be void)."""
        comm = getattr(self.step, '_comm', None)
        if comm is not None and getattr(self.step, 'dp_fused', False):
            comm.check()

    def close(self):
        import gc
        import torch
        self.step = self.model = self.cdae = self.mopt = self.copt = None
        self.resident = self.host = None
        gc.collect()
        torch.cuda.empty_cache()


def sub_record(c, B, dev, world, rank, W, K, label, with_e2e=True):
    """A secondary configuration measured the same way as the headline (device-resident value + e2e)."""
    h = Harness(c, B, dev, world, rank)
    ms, losses = h.time_resident(W, K)
    rec = dict(workload=label, per_gpu_batch=B, global_batch=B * world, steps=K, warmup=max(W, 4),
               value=B * world * K / (ms * 1e-3), unit='samples/s', ms_per_step=ms / K, final_losses=losses,
               graph=bool(h.step._g is not None), gpu_launches_per_step=h.step.count_launches(B))
    if with_e2e:
        e2e_ms, h2d = h.time_e2e(K)
        rec['e2e'] = dict(value=B * world * K / (e2e_ms * 1e-3), unit='samples/s', ms_per_step=e2e_ms / K,
                          h2d_bytes_per_step=h2d, d2h_bytes_per_step=16)
    h.check_comm()
    h.close()
    return rec


def iws_record(dev, world, rank, n_images=10000, S=5000):
    """BASELINE.json configs[4]: IWS log-likelihood, 5000 importance samples x 10k synthetic MNIST-shape images, images
    sharded by rank (ardae.evaluate_iws allreduces the partial sums); images/s over the whole job."""
    import torch
    import torch.distributed as dist
    import ardae
    c = CFG
    torch.manual_seed(1234)
    model = ardae.MNISTIPVAE(input_dim=c['D'], noise_dim=c['n'], h_dim=c['h'], num_hidden_layers=c['model_layers'],
                             nonlinearity=c['nonlin'], z_dim=c['z']).to(dev)
    with torch.no_grad():
        model.encode.fc.fc.weight.mul_(0.05)  # a posterior of trained-model width (the N(0,1) init gives |z| ~ 20)
    per = n_images // world
    g = torch.Generator(device=dev).manual_seed(77 + rank)
    x = (torch.rand(per, c['D'], device=dev, generator=g) < 0.13).float()
    pg = dist.group.WORLD if world > 1 else None
    ardae.evaluate_iws(x[:256], model, S, process_group=pg)  # warm-up: plans
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    v = ardae.evaluate_iws(x, model, S, process_group=pg)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    dec_macs = c['z'] * c['h'] + 2 * c['h'] * c['h'] + c['h'] * c['D']
    enc_macs = c['n'] * c['h'] + c['h'] * c['z']
    rows = per * world * S
    del model, x
    torch.cuda.empty_cache()
    return dict(workload='configs[4] IWS 5000 importance samples x 10k synthetic MNIST-shape images, sharded by rank',
                metric='iws_images_per_sec', value=per * world / ms * 1e3, unit='images/s', images=per * world,
                images_per_gpu=per, iws_samples=S, ms=ms, logprob_nats=float(v.item()),
                algorithmic_tflops=2.0 * rows * (dec_macs + enc_macs) / ms * 1e-9,
                note='algorithmic FLOPs: decoder + noise half of the encoder fc on every (image, sample) row (SURVEY 8d)')


def measure_tf32_peak(dev, index):
    """tf32 tensor peak the way MEASURED_PEAKS.json measures bf16: torch.matmul 8192^3, best of 10, CUDA events, with the
    SM clock sampled around the probe (MEASURED_PEAKS.json itself holds bf16 only)."""
    import torch
    torch.backends.cuda.matmul.allow_tf32 = True
    a = torch.randn(8192, 8192, device=dev)
    b = torch.randn(8192, 8192, device=dev)
    sampler = ClockSampler(index)
    sampler.start()
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(); a @ b; g1.record(); torch.cuda.synchronize()
        best = min(best, g0.elapsed_time(g1))
    t_end = time.time() + 0.5   # a few more back-to-back products so the 0.1 s sampler sees the probe under load
    while time.time() < t_end:
        a @ b
    torch.cuda.synchronize()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    del a, b
    return 2 * 8192 ** 3 / best * 1e-9, sampler.summary()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel eagerly (no CUDA-graph replay of the step)')
    ap.add_argument('--no-extra', action='store_true', help='skip the secondary records (strong scaling, configs 1 / 4 / 5)')
    ap.add_argument('--cdae', default='grad', choices=['grad', 'res'],
                    help="grad = mlp-grad (BASELINE.json's config, default); res = mlp-res residual CDAE (extra measurement)")
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    # keep stdout clean for the ONE JSON line (NCCL prints its version banner to stdout): everything else -> stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    import ardae
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    c = CFG
    W, K = max(3, args.warmup), max(1, args.steps)
    B = c['B']

    # roofline denominator first, on the idle GPU (a burst figure, like MEASURED_PEAKS.json's bf16_tflops): measured
    # after minutes of load the same probe reads 4-5 % lower (power cap), which would flatter the fraction
    tf32_peak, tf32_clocks = measure_tf32_peak(dev, local) if rank == 0 else (0.0, {})
    if world > 1:
        dist.barrier()

    # ================= headline: configs[1], 512 rows per GPU (weak scaling across --gpus)
    h = Harness(c, B, dev, world, rank, graph=not args.no_graph, cdae_kind=args.cdae)
    step, model, cdae = h.step, h.model, h.cdae
    dp_fused = bool(getattr(step, 'dp_fused', False))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total, final_losses = h.time_resident(W, K)
    graph_used = bool(step.graph and step._g is not None)

    # ---------------- profiled pass (eager launches: events cannot sit between the nodes of a graph replay):
    # K more iterations of the same workload with CUDA events around every segment, and around every launch of the
    # CDAE update in the last one -> segments_ms and the per-kernel rooflines
    from ardae import _lib
    h_train = cdae._plan(B, c['nz'] * c['nstd'], True)
    step.profile = []
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h.barrier()
    p0.record()
    for i in range(K):
        if i == K - 1:
            _lib.check(_lib.lib().ardae_cdae_set_profile(h_train, 1))
        step(*h.resident[i % h.nb], beta=c['beta'])
    p1.record()
    h.barrier()
    eager_ms = p0.elapsed_time(p1) / K
    kernel_ms = _lib.read_cdae_profile(h_train)  # per-launch times of the last profiled iteration
    _lib.check(_lib.lib().ardae_cdae_set_profile(h_train, 0))
    prof, step.profile = step.profile, None
    seg = {}
    for name, a, b in prof:
        seg.setdefault(name, []).append(a.elapsed_time(b))
    seg_ms = {k: sum(v) / K for k, v in seg.items()}

    # ---------------- end-to-end: pinned host inputs -> H2D every step, losses D2H every step
    e2e_ms, h2d_bytes = h.time_e2e(K)
    h.check_comm()
    sampler.stop_flag = True  # clocks / throttle reasons sampled across both timed regions (value and e2e)
    launches_per_step = step.count_launches(B)
    n_cdae = sum(p.numel() for p in cdae.parameters())
    opt_bytes_c = 28.0 * cdae._arena.total

    # ---------------- optimizer kernels on an HBM-resident arena (the 26 MB arenas of the step are
    # L2-resident and launch-bound; the same kernels over 256 M parameters = 7.2 GB show their HBM rate)
    opt_big = None
    if rank == 0:
        import ctypes
        nbig = 256 * 1024 * 1024
        bufs = [torch.zeros(nbig, device=dev) for _ in range(4)]
        bufs[1].fill_(1e-3)
        L = _lib.lib()
        opt_big = {}
        for name in ('adam', 'rmsprop'):
            def call():
                if name == 'adam':
                    _lib.check(L.ardae_adam_step(_lib.ptr(bufs[0]), _lib.ptr(bufs[1]), _lib.ptr(bufs[2]), _lib.ptr(bufs[3]),
                                                 ctypes.c_size_t(nbig), 1e-4, 0.5, 0.999, 1e-8, 1, 1.0, _lib.stream_ptr()))
                else:
                    _lib.check(L.ardae_rmsprop_step(_lib.ptr(bufs[0]), _lib.ptr(bufs[1]), _lib.ptr(bufs[2]), _lib.ptr(bufs[3]),
                                                    ctypes.c_size_t(nbig), 1e-4, 0.99, 1e-8, 0.5, 1.0, _lib.stream_ptr()))
            for _ in range(3):
                call()
            torch.cuda.synchronize()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(5):
                call()
            g1.record()
            torch.cuda.synchronize()
            opt_big[name] = g0.elapsed_time(g1) / 5
        del bufs
        torch.cuda.empty_cache()
    h.close()
    del step, model, cdae

    # ================= secondary records (every rank takes part; rank 0 reports)
    extra = {}
    if not args.no_extra and args.cdae == 'grad':
        Ks, Ws = max(5, min(K, 10)), 3
        # configs[2]: sbMNIST-shape, GLOBAL batch 4096 fixed, split over the ranks (strong scaling)
        if 4096 % world == 0:
            extra['strong_scaling'] = sub_record(c, 4096 // world, dev, world, rank, Ws, Ks,
                                                 'configs[2] sbMNIST-shape MLP z=32, global batch 4096 fixed (per GPU 4096 / n_gpus)',
                                                 with_e2e=False)
            extra['strong_scaling']['scaling'] = 'strong'
        extra['config1_25gaussians'] = sub_record(CFG1, CFG1['B'], dev, world, rank, Ws, Ks,
                                                  'configs[0] 25gaussians ToyIPVAE z=2 h=256 relu + mlp-grad CDAE h=256 L=3, batch 512 per GPU, nz 256')
        extra['config4_conv'] = sub_record(CFG4, CFG4['B'], dev, world, rank, Ws, Ks,
                                           'configs[3] dbMNIST ConvIPVAE 28x28 z=32 + mlp-grad CDAE h=256 L=5, 1024 rows per GPU (8192 over 8), nz 256')
        extra['aux_hidden1a'] = sub_record(CFG_AUX, CFG_AUX['B'], dev, world, rank, Ws, Ks,
                                           'run_vae_dbmnist.sh:28 family: MNISTAuxIPVAE 784/300/100/32 + mlp-grad CDAE h=256 L=5 on the hidden1a context (600 wide), batch 512 per GPU, nz 256',
                                           with_e2e=False)
        extra['config5_iws'] = iws_record(dev, world, rank)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline denominators
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    hbm_src = 'MEASURED_PEAKS.json (hbm_gbs)' if 'hbm_gbs' in peaks else 'fallback (B200_PROFILING.md)'
    bf16_peak = float(peaks.get('bf16_tflops', 2 * tf32_peak))
    tf32_src = ('torch.matmul tf32 8192^3, best of 10, measured at the start of this run (idle GPU) with its own clock record '
                '(MEASURED_PEAKS.json holds bf16 only: %s burst; half of it = %.0f)' % (peaks.get('bf16_tflops'), bf16_peak / 2))

    flops = cdae_alg_flops(B, c['nz'] * c['nstd'], c['z'], c['z'], c['cdae_h'], c['cdae_L'])
    ct = seg_ms['cdae_train']
    achieved = flops / ct * 1e-9
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))
    except Exception:
        pass
    # ---- per-kernel view of the CDAE update (CUDA events around each launch of the last profiled step).
    # FLOPs are algorithmic.  The [N, H] activation arrays a kernel streams are SPILL: by SURVEY 8d they are
    # overhead, not algorithmic bytes -- reported as overhead_bytes with the HBM rate they are moved at.
    N_, H_, L_, d_ = B * c['nz'] * c['nstd'], c['cdae_h'], c['cdae_L'], c['z']
    kp = (d_ + 31) // 32 * 32
    spill16 = args.cdae == 'grad'
    arr = N_ * H_ * (2.0 if spill16 else 4.0)  # one [N, H] spill array (bf16 in the mlp-grad plan)
    nl = 2 * L_ - 1
    sweep_flops = 2.0 * N_ * (nl * H_ * H_ + kp * H_)
    by_tag = {}
    for tag, t_ms in kernel_ms:
        by_tag.setdefault(tag, []).append(t_ms)
    info = {
        'chain_softplus3': dict(flops=sweep_flops, overhead=(nl + 1) * arr, executed=3.0,
                                note='primal forward, all 2L layers, 3xTF32 (three tensor products per algorithmic product)'),
        'chain_mul_sig': dict(flops=sweep_flops, overhead=(2 * nl + 1) * arr, executed=1.0, note='score backward incl. g = delta a_1 . A_1'),
        'chain_tangent': dict(flops=sweep_flops, overhead=4 * (nl + 1) * arr, executed=1.0, note='tangent forward (+ t spill)'),
        'chain_adjoint': dict(flops=2.0 * N_ * nl * H_ * H_, overhead=(3 * nl + 1) * arr, executed=1.0, note='adjoint backward, in place over t'),
        'gemm_tn16': dict(flops=2 * 2.0 * N_ * H_ * H_, overhead=4 * arr, executed=1.0, peak=bf16_peak,
                          note='weight gradient of one [H,H] layer: two bf16 contractions over the N rows'),
        'gemm_tn16_multi': dict(flops=2 * (L_ - 1) * 2 * 2.0 * N_ * H_ * H_, overhead=2 * (L_ - 1) * 4 * arr, executed=1.0,
                                peak=bf16_peak, note='weight gradients of the 2(L-1) [H,H] layers in one launch (bf16 x bf16 -> fp32)'),
        'gemm_tn': dict(flops=2 * 2.0 * N_ * H_ * H_, overhead=4 * arr, executed=1.0, note='weight gradient (fp32 spill plans)'),
    }
    kernels = []
    for tag in ('chain_softplus3', 'chain_tangent', 'chain_adjoint', 'chain_mul_sig', 'gemm_tn16_multi', 'gemm_tn16', 'gemm_tn'):
        ts = sorted(t for t in by_tag.get(tag, []) if t > 0.02)  # the N-row launches only
        if not ts:
            continue
        med = ts[len(ts) // 2]
        ts = [t for t in ts if 0.7 * med <= t <= 1.4 * med]  # the [H,H] contractions (not the d-wide / odd-pitch ones)
        avg = sum(ts) / len(ts)
        i = info[tag]
        pk = i.get('peak', tf32_peak)
        tfl = i['flops'] / avg * 1e-9
        gbs = i['overhead'] / avg * 1e-6
        kernels.append(dict(kernel=tag, launches=len(ts), ms_per_launch=avg, algorithmic_flops_per_launch=i['flops'],
                            achieved=tfl, unit='TFLOP/s', peak=pk, frac=tfl / pk, executed_frac=i['executed'] * tfl / pk,
                            algorithmic_bytes_per_launch=0.0, overhead_bytes_per_launch=i['overhead'],
                            overhead_gbs=gbs, overhead_hbm_frac=gbs / hbm_peak,
                            bound=('tensor' if i['executed'] * tfl / pk >= gbs / hbm_peak else 'hbm (activation spill: overhead bytes)'),
                            traffic=traffic.get(tag), note=i['note']))
    plan_overhead = sum(k['overhead_bytes_per_launch'] * k['launches'] for k in kernels)
    plan_traffic = None
    if traffic.get('cdae_train_total'):
        plan_traffic = traffic['cdae_train_total']
    step_flops = step_alg_flops(c, B)
    ms_step = ms_total / K
    line = dict(
        metric='train_samples_per_sec', value=B * world * K / (ms_total * 1e-3), unit='samples/s', n_gpus=world,
        steps=K, warmup=W, ms_per_step=ms_step, higher_is_better=True, scaling='weak', vs_baseline=None,
        dtype='tf32', data='synthetic',
        config=dict(workload=WORKLOAD if args.cdae == 'grad' else 'configs[1] shapes with --cdae mlp-res (NOT the BASELINE config)',
                    detail=WORKLOAD_DETAIL,
                    global_batch=B * world, per_gpu_batch=B, cdae_rows_per_gpu=B * c['nz'],
                    parallelism='dp%d' % world,
                    arithmetic='tf32 tensor-core operands, fp32 accumulate; primal forward 3xTF32; [N,H] activation spill stored bf16; weight gradients bf16 x bf16 -> fp32',
                    l2='no flush needed: per-step working set (activation spill) ~3 GB >> 126 MB L2',
                    launch=(('one CUDA-graph replay per step (%d kernels on 3 streams captured)' % launches_per_step
                             if (world == 1 or dp_fused) else
                             '3 CUDA-graph replays + 2 eager NCCL allreduces per step (%d kernels captured)' % launches_per_step)
                            if graph_used else 'eager launches'),
                    dp_exchange=('none (1 GPU)' if world == 1 else
                                 ('fused peer-memory kernel: gradient slices pushed over NVLink, reduce + optimizer update '
                                  '+ parameter push in one launch per arena (csrc/dp_fused.cuh)' if dp_fused else
                                  'NCCL SUM allreduce of the two flat gradient arenas + optimizer launch')),
                    ms_per_step_eager_profiled=eager_ms,
                    noise='in-kernel Philox', final_losses=final_losses),
        clocks=sampler.summary(),
        e2e=dict(value=B * world * K / (e2e_ms * 1e-3), unit='samples/s', ms_per_step=e2e_ms / K,
                 h2d_bytes_per_step=h2d_bytes, d2h_bytes_per_step=16),
        gpu_launches=launches_per_step * K,
        segments_ms=dict({k: round(v, 4) for k, v in seg_ms.items()},
                         note='eager profiled pass (CUDA events per segment); the timed region replays one graph'),
        # the north-star figure: algorithmic FLOPs of the dominant piece of the step, the CDAE update
        # (4 chain launches + 2L weight-gradient contractions + prologue / loss / context branch), over its duration
        roofline=dict(bound='tensor', kernel='cdae_train (whole CDAE update: 4 fused chain launches + %d weight-gradient contractions)' % (2 * L_),
                      achieved=achieved, peak=tf32_peak, unit='TFLOP/s', frac=achieved / tf32_peak,
                      frac_vs_half_bf16_burst=achieved / (bf16_peak / 2),
                      algorithmic_flops_per_launch=flops, ms_per_launch=ct, peak_source=tf32_src, peak_probe_clocks=tf32_clocks,
                      traffic=plan_traffic,
                      note='SURVEY 8d: 6 x 2 x N x G algorithmic FLOPs; executed tensor work is higher (3xTF32 primal sweep)'),
        roofline_step=dict(bound='tensor', kernel='whole training step', achieved=step_flops / ms_step * 1e-9, peak=tf32_peak,
                           unit='TFLOP/s', frac=step_flops / ms_step * 1e-9 / tf32_peak, algorithmic_flops_per_step=step_flops),
        roofline_kernels=kernels,
        roofline_spill=dict(bound='hbm', kernel='activation spill of the CDAE update (bf16 [N,H] arrays between the sweeps)',
                            algorithmic_bytes=0.0, overhead_bytes=plan_overhead,
                            achieved=plan_overhead / ct * 1e-6, peak=hbm_peak, unit='GB/s',
                            frac=plan_overhead / ct * 1e-6 / hbm_peak, peak_source=hbm_src,
                            note='by SURVEY 8d per-row activation I/O is overhead, not algorithmic bytes; listed so the remaining gap is visible'),
        roofline_hbm=dict(bound='hbm', kernel='rmsprop_kernel (flat CDAE arena, %d params)' % n_cdae,
                          achieved=opt_bytes_c / seg_ms['cdae_opt'] * 1e-6, peak=hbm_peak, unit='GB/s',
                          frac=opt_bytes_c / seg_ms['cdae_opt'] * 1e-6 / hbm_peak, peak_source=hbm_src,
                          algorithmic_bytes_per_launch=opt_bytes_c,
                          note='28 B/param; the 26 MB arena of the step is L2-resident and launch-latency bound (SURVEY 7.2-8)',
                          hbm_resident_arena=(dict(params=256 * 1024 * 1024, bytes=28.0 * 256 * 1024 * 1024,
                                                   adam_ms=opt_big['adam'], rmsprop_ms=opt_big['rmsprop'],
                                                   adam_gbs=28.0 * 256 * 1024 * 1024 / opt_big['adam'] * 1e-6,
                                                   rmsprop_gbs=28.0 * 256 * 1024 * 1024 / opt_big['rmsprop'] * 1e-6,
                                                   adam_frac=28.0 * 256 * 1024 * 1024 / opt_big['adam'] * 1e-6 / hbm_peak,
                                                   rmsprop_frac=28.0 * 256 * 1024 * 1024 / opt_big['rmsprop'] * 1e-6 / hbm_peak)
                                              if opt_big else None)))
    line.update(extra)
    if not args.no_cpu_baseline and world == 1:
        cb, _ = cpu_reference_leg(steps=3, warmup=1)  # ~10 s of host work: 4 iterations on 128 of the 512 data rows
        line['cpu_baseline'] = cb
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        os.dup2(2, 1)
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
