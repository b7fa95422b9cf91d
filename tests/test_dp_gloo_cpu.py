"""world_size-2 gloo test (CPU) of the data-parallel design: batch shards, losses and
gradients normalised by the GLOBAL counts (ardae.step.dp_scales), one SUM allreduce of the flat gradient
buffer per network == the single-process result on the concatenated batch.  The numpy oracle stands in for
the CUDA kernels as the per-rank compute (the kernels themselves need a B200)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ardae_oracle as orc
from golden_util import load_case, rel_err, sub


def _specs(meta):
    m, c = meta['model'], meta['cdae']
    spec = orc.ModelSpec(meta['kind'], m['input_dim'], m['noise_dim'], m['h_dim'], m['z_dim'], m['num_hidden_layers'],
                         m['nonlinearity'])
    return spec, orc.CdaeSpec(c['input_dim'], c['context_dim'], c['h_dim'], c['num_hidden_layers'])


def _worker(rank, world, port, q):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (here, os.path.join(here, '..', 'oracle'), os.path.join(here, '..', 'pytorch-ardae-vae_b200')):
        sys.path.insert(0, p)
    from ardae.step import dp_scales
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    z, meta = load_case('toy_small')
    spec, cs = _specs(meta)
    hp = meta['hp']
    Pm, Pc = sub(z, 'm0/'), sub(z, 'c0/')
    B = z['s0/x_cdae'].shape[0]
    Bl = B // world
    sl = slice(rank * Bl, (rank + 1) * Bl)
    nz, nstd, d = hp['nz_cdae'], hp['nstd'], spec.z_dim
    sc = dp_scales(Bl, nz, nstd, d, hp['nz_model'], world)
    # ---- CDAE shard: oracle loss/grads use LOCAL means; rescale to the global normalisation
    noise = sub(z, 's0/noise/')
    x = z['s0/x_cdae'][sl]
    zbar, _ = orc.encoder_forward(spec, Pm, x, np.zeros((Bl, spec.noise_dim)), 1)
    zz, _ = orc.encoder_forward(spec, Pm, x, noise['enc_cdae'].reshape(B, nz, -1)[sl].reshape(Bl * nz, -1), nz)
    lsm, std = orc.sigma_schedule(zz, zbar, hp['std_scale'], hp['delta'])
    loss, g, G = orc.cdae_loss_and_grads(cs, Pc, np.repeat(lsm, nstd, 1), zbar, std * noise['xi'][sl], noise['eps_cdae'][sl])
    local_to_global = sc['cdae_inv_count'] * (Bl * nz * nstd * d)
    keys = sorted(G)
    flat = torch.from_numpy(np.concatenate([G[k].ravel() for k in keys]) * local_to_global)
    lt = torch.tensor([loss * local_to_global])
    dist.all_reduce(flat)          # the one collective of the CDAE update
    dist.all_reduce(lt)
    if rank == 0:
        q.put((keys, flat.numpy(), float(lt)))
    dist.destroy_process_group()


def test_dp2_sum_allreduce_reproduces_full_batch():
    z, meta = load_case('toy_small')
    spec, cs = _specs(meta)
    hp = meta['hp']
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    keys, flat, loss = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference on the whole batch = the reference fixture itself
    ref = sub(z, 's0/cdae_grads/')
    ref_flat = np.concatenate([ref[k].ravel() for k in keys])
    assert rel_err(flat, ref_flat) < 1e-6
    assert abs(loss - float(z['s0/cdae_loss'])) < 1e-7 * abs(loss)


def test_dp_scales():
    from ardae.step import dp_scales
    s1 = dp_scales(512, 256, 1, 32, 1, 1)
    s8 = dp_scales(64, 256, 1, 32, 1, 8)
    assert s1 == s8  # strong scaling: same global batch -> identical normalisation
    assert s1['cdae_inv_count'] == 1.0 / (512 * 256 * 32) and s1['model_inv_rows'] == 1.0 / 512
