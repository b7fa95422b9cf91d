"""Robustness check of the data-parallel set-up: two ranks that each SEE ONLY THEIR OWN GPU (CUDA_VISIBLE_DEVICES = one
device per process).  Either the peer mapping works across the restricted visibility or every rank must fall back to
the NCCL exchange together -- never a hang, never a split decision.   python scripts/dp_fallback_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker(rank, world, port):
    os.environ['CUDA_VISIBLE_DEVICES'] = str(rank)
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    sys.path[:0] = [ROOT, os.path.join(ROOT, 'pytorch-ardae-vae_b200')]
    import torch
    import ardae
    import bench
    torch.cuda.set_device(0)
    torch.distributed.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', 0))
    c = bench.CFG
    torch.manual_seed(1)
    model, cdae, mopt, copt = bench.build_models(c, torch.device('cuda', 0))
    step = ardae.TrainStep(model, cdae, mopt, copt, nz_cdae=32, process_group=torch.distributed.group.WORLD, graph=True)
    x = (torch.rand(64, 784, device='cuda') < 0.13).float()
    for _ in range(6):
        out = step(x, x, beta=1.0)
    torch.cuda.synchronize()
    flat = torch.cat([model._arena.flat, cdae._arena.flat])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    torch.distributed.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], g) for g in gathered[1:])
    if step.dp_fused:
        step._comm.check()
    print('rank %d: dp_fused=%s replicas identical=%s loss=%.4f' % (rank, step.dp_fused, same, out['cdae_loss'].item()))
    assert same
    torch.distributed.destroy_process_group()


if __name__ == '__main__':
    import multiprocessing as mp
    mp.set_start_method('spawn')
    ps = [mp.Process(target=worker, args=(r, 2, 29577)) for r in range(2)]
    [p.start() for p in ps]
    [p.join() for p in ps]
    sys.exit(max(p.exitcode for p in ps))
