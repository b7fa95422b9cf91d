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
WORKLOAD = ('configs[1]: dbMNIST-shape ivae_ardae, MNISTIPVAE mlp-concat D=784 h=300 n=100 z=32 + mlp-grad CDAE '
            'h=256 L=5, batch 512 per GPU, train-nz-cdae 256, Adam(model)+RMSprop(cdae)')


def cdae_alg_flops(B, nz, d, c, H, L):
    """6 sweeps x 2 x N x G + context branch on the B distinct rows."""
    N = B * nz
    G = (2 * L - 1) * H * H + d * H + H
    ctx = c * H + (L - 1) * H * H + H * H
    return 6 * 2 * N * G + 3 * 2 * B * ctx


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
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    cb, sec = cpu_reference_leg(steps, warmup)
    line = dict(metric='train_samples_per_sec', value=cb['value'], unit='samples/s', impl='reference',
                n_gpus=args.gpus, steps=steps, warmup=warmup, ms_per_step=sec * 1e3, higher_is_better=True,
                scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
                config=dict(workload=WORKLOAD, note='CPU, bounded sample: ' + cb['sample']),
                cpu_baseline=cb,
                e2e=dict(value=cb['value'], unit='samples/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel eagerly (no CUDA-graph replay of the step)')
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

    torch.manual_seed(1234)  # same weights on every rank (replicated parameters)
    model = ardae.MNISTIPVAE(input_dim=c['D'], noise_dim=c['n'], h_dim=c['h'], num_hidden_layers=c['model_layers'],
                             nonlinearity=c['nonlin'], enc_type='concat', z_dim=c['z']).to(dev)
    cdae_cls = ardae.MLPGradCARDAE if args.cdae == 'grad' else ardae.MLPResCARDAE
    cdae = cdae_cls(input_dim=c['z'], context_dim=c['z'], std=1., h_dim=c['cdae_h'],
                    num_hidden_layers=c['cdae_L'], nonlinearity='softplus').to(dev)
    mopt = ardae.Adam(model.parameters(), lr=c['m_lr'], betas=(c['m_beta1'], 0.999))
    copt = ardae.RMSprop(cdae.parameters(), lr=c['d_lr'], momentum=c['d_momentum'])
    step = ardae.TrainStep(model, cdae, mopt, copt, std_scale=c['std_scale'], delta=c['delta'], nz_cdae=c['nz'],
                           nstd=c['nstd'], nz_model=c['nz_model'],
                           process_group=dist.group.WORLD if world > 1 else None, seed=1234,
                           graph=not args.no_graph)
    B = c['B']
    gen = torch.Generator().manual_seed(999 + rank)
    pix = torch.rand(1, c['D'], generator=torch.Generator().manual_seed(5)) * 0.26  # mean ink fraction ~0.13
    nb = 8
    host = [(torch.bernoulli(pix.expand(B, -1), generator=gen).pin_memory(),
             torch.bernoulli(pix.expand(B, -1), generator=gen).pin_memory()) for _ in range(nb)]
    resident = [(a.to(dev), b.to(dev)) for a, b in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (`value`): K iterations (CUDA-graph replays unless --no-graph)
    for i in range(max(W, 4)):  # >= 2 eager iterations + capture + 1 replay
        step(*resident[i % nb], beta=c['beta'])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K):
        out = step(*resident[i % nb], beta=c['beta'])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    graph_used = bool(step.graph and step._g is not None)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = t.item()
    final_losses = {k: float(out[k].item()) for k in ('cdae_loss', 'model_loss')}

    # ---------------- profiled pass (eager launches: events cannot sit between the nodes of a graph replay):
    # K more iterations of the same workload with CUDA events around every segment, and around every launch of the
    # CDAE update in the last one -> segments_ms and the per-kernel rooflines
    from ardae import _lib
    h_train = cdae._plan(B, c['nz'] * c['nstd'], True)
    step.profile = []
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    p0.record()
    for i in range(K):
        if i == K - 1:
            _lib.check(_lib.lib().ardae_cdae_set_profile(h_train, 1))
        step(*resident[i % nb], beta=c['beta'])
    p1.record()
    barrier()
    eager_ms = p0.elapsed_time(p1) / K
    kernel_ms = _lib.read_cdae_profile(h_train)  # per-launch times of the last profiled iteration
    _lib.check(_lib.lib().ardae_cdae_set_profile(h_train, 0))
    prof, step.profile = step.profile, None
    seg = {}
    for name, a, b in prof:
        seg.setdefault(name, []).append(a.elapsed_time(b))
    seg_ms = {k: sum(v) / K for k, v in seg.items()}

    # ---------------- end-to-end: pinned host inputs -> H2D every step, losses D2H every step
    xbuf = [torch.empty(B, c['D'], device=dev) for _ in range(2)]
    hloss = torch.empty(4, pin_memory=True)
    for i in range(2):
        step(*resident[i % nb], beta=c['beta'])
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(K):
        xbuf[0].copy_(host[i % nb][0], non_blocking=True)
        xbuf[1].copy_(host[i % nb][1], non_blocking=True)
        o = step(xbuf[0], xbuf[1], beta=c['beta'])
        hloss.copy_(torch.cat([o['cdae_loss'], o['model_loss'], o['recon'], o['prior']]), non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller reads the losses every step
    f1.record()
    barrier()
    sampler.stop_flag = True  # clocks / throttle reasons sampled across both timed regions (value and e2e)
    t2 = torch.tensor([f0.elapsed_time(f1)], device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_ms = t2.item()

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
    hbm_src = 'MEASURED_PEAKS.json' if 'hbm_gbs' in peaks else 'fallback (B200_PROFILING.md)'
    torch.backends.cuda.matmul.allow_tf32 = True
    a = torch.randn(8192, 8192, device=dev)
    b = torch.randn(8192, 8192, device=dev)
    for _ in range(2):
        a @ b
    best = 1e9
    for _ in range(5):
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(); a @ b; g1.record(); torch.cuda.synchronize()
        best = min(best, g0.elapsed_time(g1))
    tf32_peak = 2 * 8192 ** 3 / best * 1e-9  # TFLOP/s
    del a, b

    flops = cdae_alg_flops(B, c['nz'] * c['nstd'], c['z'], c['z'], c['cdae_h'], c['cdae_L'])
    ct = seg_ms['cdae_train']
    achieved = flops / ct * 1e-9
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))
    except Exception:
        pass
    # ---- per-kernel rooflines of the CDAE update (CUDA events around each launch, last timed step)
    N_, H_, L_ = B * c['nz'] * c['nstd'], c['cdae_h'], c['cdae_L']
    arr = N_ * H_ * 4.0  # one [N, H] fp32 activation array
    by_tag = {}
    for tag, t_ms in kernel_ms:
        by_tag.setdefault(tag, []).append(t_ms)
    nlay = 2 * L_ - 1  # fused H->H layers per sweep
    # algorithmic HBM bytes per launch: aux reads + spill writes per fused layer, + the initial activation
    nbatch = 2 * (L_ - 1)  # [H,H] weight-gradient contractions batched into one launch (grid.z = layer)
    alg_bytes = {'chain_mul_sig': (2 * nlay + 1) * arr, 'chain_tangent': (4 * nlay + 1) * arr,
                 'chain_adjoint': (3 * nlay + 1) * arr, 'gemm_tn': 4 * arr,
                 'gemm_tn_batch': nbatch * (4 if args.cdae == 'grad' else 2) * arr}
    kernels = []
    for tag in ('chain_tangent', 'chain_adjoint', 'chain_mul_sig', 'gemm_tn_batch', 'gemm_tn'):
        ts = sorted(t for t in by_tag.get(tag, []) if t > 0.02)  # the big (N-row) launches only
        if not ts:
            continue
        med = ts[len(ts) // 2]
        ts = [t for t in ts if 0.8 * med <= t <= 1.25 * med]  # gemm_tn: the 2L-1 [H,H] contractions (not d-wide / two-pass ones)
        avg = sum(ts) / len(ts)
        a = alg_bytes[tag] / avg * 1e-6
        kernels.append(dict(kernel=tag, bound='hbm', launches=len(ts), ms_per_launch=avg, achieved=a, peak=hbm_peak,
                            unit='GB/s', frac=a / hbm_peak, algorithmic_bytes_per_launch=alg_bytes[tag],
                            traffic=(traffic.get(tag) if tag != 'gemm_tn_batch' else
                                     (traffic.get('gemm_tn') * nbatch if traffic.get('gemm_tn') else None))))
    ts3 = [t for t in by_tag.get('chain_softplus3', []) if t > 0.02]
    if ts3:
        # the two 3xTF32 chains together run nlay layers: algorithmic 2*N*H*H per layer (executed: 3x)
        tot = sum(ts3)
        a = 2.0 * N_ * H_ * H_ * nlay / tot * 1e-9
        kernels.append(dict(kernel='chain_softplus3', bound='tensor', launches=len(ts3), ms_per_launch=tot / len(ts3),
                            achieved=a, peak=tf32_peak, unit='TFLOP/s', frac=a / tf32_peak, executed_tflops=3 * a,
                            executed_frac=3 * a / tf32_peak,
                            note='3xTF32: three tensor-core products per algorithmic product (fp32-accurate forward)',
                            traffic=traffic.get('chain_softplus3')))
    dominant = max(kernels, key=lambda k: k['ms_per_launch'] * k['launches']) if kernels else None
    n_cdae = sum(p.numel() for p in cdae.parameters())
    n_model = sum(p.numel() for p in model.parameters())
    opt_bytes_c, opt_bytes_m = 28.0 * cdae._arena.total, 28.0 * model._arena.total
    roof_opt = dict(bound='hbm', kernel='rmsprop_kernel (flat CDAE arena, %d params)' % n_cdae,
                    achieved=opt_bytes_c / seg_ms['cdae_opt'] * 1e-6, peak=hbm_peak, unit='GB/s',
                    frac=opt_bytes_c / seg_ms['cdae_opt'] * 1e-6 / hbm_peak, peak_source=hbm_src,
                    note='28 B/param; 26 MB arena is L2-resident and launch-latency bound (SURVEY 7.2-8)')
    line = dict(
        metric='train_samples_per_sec', value=B * world * K / (ms_total * 1e-3), unit='samples/s', n_gpus=world,
        steps=K, warmup=W, ms_per_step=ms_total / K, higher_is_better=True, scaling='weak', vs_baseline=None,
        dtype='tf32', data='synthetic',
        config=dict(workload=WORKLOAD if args.cdae == 'grad' else WORKLOAD.replace('mlp-grad', 'mlp-res (residual, NOT the BASELINE config)'),
                    global_batch=B * world, per_gpu_batch=B, cdae_rows_per_gpu=B * c['nz'],
                    parallelism='dp%d' % world, arithmetic='tf32 tensor-core operands, fp32 accumulate; forward sweeps 3xTF32',
                    l2='no flush needed: per-step working set (activation spill) ~7 GB >> 126 MB L2',
                    launch=(('one CUDA-graph replay per step (%d kernels on 3 streams captured)' % step.count_launches(B)
                             if world == 1 else
                             '3 CUDA-graph replays + 2 eager NCCL allreduces per step (%d kernels captured)' % step.count_launches(B))
                            if graph_used else 'eager launches'),
                    ms_per_step_eager_profiled=eager_ms,
                    noise='in-kernel Philox', final_losses=final_losses),
        clocks=sampler.summary(),
        e2e=dict(value=B * world * K / (e2e_ms * 1e-3), unit='samples/s', ms_per_step=e2e_ms / K,
                 h2d_bytes_per_step=2 * B * c['D'] * 4, d2h_bytes_per_step=16),
        gpu_launches=step.count_launches(B) * K,
        segments_ms=dict({k: round(v, 4) for k, v in seg_ms.items()},
                         note='eager profiled pass (CUDA events per segment); the timed region replays one graph'),
        roofline=(dict(dominant, peak_source=(hbm_src + ' (hbm_gbs)' if dominant['bound'] == 'hbm' else
                                              'torch.matmul tf32 8192^3 measured in this run (MEASURED_PEAKS.json holds bf16 only)'),
                       note='dominant kernel of the step by total time; every kernel of the CDAE update is in roofline_kernels')
                  if dominant else None),
        roofline_kernels=kernels,
        roofline_plan=dict(bound='tensor', kernel='whole cdae_train plan (4 fused chains + %d weight-gradient contractions + first/last layers)' % (2 * c['cdae_L']),
                           achieved=achieved, peak=tf32_peak, unit='TFLOP/s', frac=achieved / tf32_peak,
                           peak_source='torch.matmul tf32 8192^3 measured in this run (MEASURED_PEAKS.json holds bf16 only: %s burst)' % peaks.get('bf16_tflops'),
                           algorithmic_flops_per_launch=flops, ms_per_launch=ct,
                           note='the plan is HBM-bound by its activation spill (see roofline_kernels); this is the north-star tensor figure'),
        roofline_hbm=roof_opt)
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
