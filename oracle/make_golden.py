"""Generate tests/golden/*.npz by executing the REFERENCE modules (fp64, CPU, injected noise).

Run in the build container only (needs /root/reference):  python oracle/make_golden.py
The fixtures pin oracle/ardae_oracle.py (tests/test_oracle_golden.py) and, through it, the CUDA path.
Sizes are tiny so the files stay small; the architecture (layer structure, concat pattern,
activation, heads, optimizers) is the one ivae_ardae.py builds for configs 1 and 2.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')

CASES = {
    # 25gaussians-shaped (run_vae_25gaussians.sh): relu ToyIPVAE, CDAE L=3
    'toy_small': dict(kind='toy',
                      model=dict(input_dim=2, noise_dim=3, h_dim=16, num_hidden_layers=2, nonlinearity='relu', z_dim=2),
                      cdae=dict(input_dim=2, context_dim=2, h_dim=16, num_hidden_layers=3, nonlinearity='softplus'),
                      B=6, hp=dict(std_scale=10000., delta=0.1, nz_cdae=8, nstd=1, nz_model=1, beta=1.0,
                                   m_lr=1e-4, m_beta1=0.5, d_lr=1e-4, d_momentum=0.5), wscale=1.0),
    # dbMNIST-shaped (run_vae_dbmnist.sh:37): softplus MNISTIPVAE, CDAE L=5
    'mnist_small': dict(kind='mnist',
                        model=dict(input_dim=20, noise_dim=5, h_dim=12, num_hidden_layers=2, nonlinearity='softplus', z_dim=4),
                        cdae=dict(input_dim=4, context_dim=4, h_dim=16, num_hidden_layers=5, nonlinearity='softplus'),
                        B=5, hp=dict(std_scale=10000., delta=0.1, nz_cdae=6, nstd=2, nz_model=1, beta=0.7,
                                     m_lr=1e-4, m_beta1=0.5, d_lr=1e-4, d_momentum=0.5), wscale=1.0),
    # same, every matrix scaled x3 -> softplus / sigmoid in their saturated regions
    'mnist_small_x3': dict(kind='mnist',
                           model=dict(input_dim=20, noise_dim=5, h_dim=12, num_hidden_layers=2, nonlinearity='softplus', z_dim=4),
                           cdae=dict(input_dim=4, context_dim=4, h_dim=16, num_hidden_layers=5, nonlinearity='softplus'),
                           B=5, hp=dict(std_scale=100., delta=0.1, nz_cdae=6, nstd=1, nz_model=2, beta=1.0,
                                        m_lr=1e-3, m_beta1=0.9, d_lr=1e-3, d_momentum=0.9), wscale=3.0),
    # residual CDAE (--cdae mlp-res, run_vae_dbmnist.sh:25: the network outputs the score, plain back-prop), std_scale 100
    'mnist_small_res': dict(kind='mnist', cdae_kind='res',
                            model=dict(input_dim=20, noise_dim=5, h_dim=12, num_hidden_layers=2, nonlinearity='softplus', z_dim=4),
                            cdae=dict(input_dim=4, context_dim=4, h_dim=16, num_hidden_layers=5, nonlinearity='softplus'),
                            B=5, hp=dict(std_scale=100., delta=0.1, nz_cdae=6, nstd=1, nz_model=1, beta=1.0,
                                         m_lr=1e-3, m_beta1=0.9, d_lr=1e-4, d_momentum=0.9), wscale=1.0),
    # --cdae-ctx-type data: the CDAE is conditioned on the input image (2x-1) instead of the mean code
    'mnist_small_datactx': dict(kind='mnist', ctx_type='data',
                                model=dict(input_dim=20, noise_dim=5, h_dim=12, num_hidden_layers=2, nonlinearity='softplus', z_dim=4),
                                cdae=dict(input_dim=4, context_dim=20, h_dim=16, num_hidden_layers=3, nonlinearity='softplus'),
                                B=5, hp=dict(std_scale=10000., delta=0.1, nz_cdae=6, nstd=1, nz_model=1, beta=1.0,
                                             m_lr=1e-4, m_beta1=0.5, d_lr=1e-4, d_momentum=0.5), wscale=1.0),
    # conv implicit VAE (run_vae_dbmnist.sh:31 shape family, 12x12 images): `lite` = one step, weights
    # fp32-representable and stored as float32 (the reference hard-codes the 800 / 300 wide fc layers)
    'conv_small': dict(kind='conv', lite=True,
                       model=dict(input_height=12, input_channels=1, z_dim=4, noise_dim=5, nonlinearity='softplus'),
                       cdae=dict(input_dim=4, context_dim=4, h_dim=16, num_hidden_layers=3, nonlinearity='softplus'),
                       B=4, hp=dict(std_scale=10000., delta=0.1, nz_cdae=6, nstd=1, nz_model=1, beta=1.0,
                                    m_lr=1e-4, m_beta1=0.5, d_lr=1e-4, d_momentum=0.5), wscale=1.0),
    # hierarchical encoder family (--model auxmnist, run_vae_dbmnist.sh:40) with --cdae-ctx-type hidden1a: the CDAE
    # context is cat(h0, h) of the encoder at std = 0 (2 h_dim wide); noise packs eps0 | eps as [R, noise_dim + z_dim]
    'auxmnist_small': dict(kind='auxmnist', ctx_type='hidden1a',
                           model=dict(input_dim=20, noise_dim=5, h_dim=12, num_hidden_layers=2, nonlinearity='softplus', z_dim=4),
                           cdae=dict(input_dim=4, context_dim=24, h_dim=16, num_hidden_layers=3, nonlinearity='softplus'),
                           B=5, hp=dict(std_scale=10000., delta=0.1, nz_cdae=6, nstd=1, nz_model=1, beta=1.0,
                                        m_lr=1e-4, m_beta1=0.5, d_lr=1e-4, d_momentum=0.5), wscale=1.0),
    # config 4's model at its real geometry: ConvIPVAE 28x28, z = 32, noise 100 (run_vae_dbmnist.sh:31), two full steps.
    # `sampled`: tensors with more than 16384 elements are stored (gradients, post-step weights) at 8192 fixed random
    # positions ('sample_idx/<name>'); the initial weights are stored in full as float32
    'conv28': dict(kind='conv', lite=True, sampled=True, steps=2,
                   model=dict(input_height=28, input_channels=1, z_dim=32, noise_dim=100, nonlinearity='softplus'),
                   cdae=dict(input_dim=32, context_dim=32, h_dim=16, num_hidden_layers=3, nonlinearity='softplus'),
                   B=4, hp=dict(std_scale=10000., delta=0.1, nz_cdae=6, nstd=1, nz_model=1, beta=1.0,
                                m_lr=1e-4, m_beta1=0.5, d_lr=1e-4, d_momentum=0.5), wscale=1.0),
}


def t2n(d):
    return {k: (v.detach().numpy().copy() if v is not None else None) for k, v in d.items()}


def make_case(name, c):
    import zlib
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) % (2 ** 31))
    dt = torch.float64
    model, cdae = rh.build_reference(c['kind'], c['model'], c['cdae'], dtype=dt, seed=1234,
                                     cdae_kind=c.get('cdae_kind', 'grad'))
    if c['wscale'] != 1.0:
        with torch.no_grad():
            for m in (model, cdae):
                for k, p in m.named_parameters():
                    if p.dim() == 2:
                        p.mul_(c['wscale'])
    hp = c['hp']
    lite = c.get('lite', False)
    if lite:  # make every weight exactly fp32-representable
        with torch.no_grad():
            for m in (model, cdae):
                for p in m.parameters():
                    p.copy_(p.float().double())
    mopt, copt = rh.build_optimizers(model, cdae, hp)
    B, n, d = c['B'], c['model']['noise_dim'], c['model']['z_dim']
    if c['kind'] == 'auxmnist':
        n = n + d  # eps0 | eps packed
    D = c['model']['input_dim'] if 'input_dim' in c['model'] else c['model']['input_channels'] * c['model']['input_height'] ** 2
    f32 = (lambda a: a.astype(np.float32)) if lite else (lambda a: a)
    nz, nstd, nzm = hp['nz_cdae'], hp['nstd'], hp['nz_model']
    arrays = {}
    sampled = c.get('sampled', False)
    nsteps = c.get('steps', 1 if lite else 2)
    if sampled:
        rs = np.random.RandomState(zlib.crc32(name.encode()) % (2 ** 31))
        for k, v in model.state_dict().items():
            if v.numel() > 16384:
                arrays['sample_idx/' + k] = np.sort(rs.choice(v.numel(), 8192, replace=False)).astype(np.int64)

    def pick(k, a):  # the stored view of a model tensor: sampled positions for big tensors
        return a.ravel()[arrays['sample_idx/' + k]] if ('sample_idx/' + k) in arrays else a
    for k, v in model.state_dict().items():
        arrays['m0/' + k] = f32(v.numpy().copy())
    for k, v in cdae.state_dict().items():
        arrays['c0/' + k] = f32(v.numpy().copy())
    for step in range(nsteps):
        if c['kind'] in ('mnist', 'conv', 'auxmnist'):
            xc = torch.bernoulli(torch.full((B, D), 0.3, dtype=dt), generator=g)
            xm = torch.bernoulli(torch.full((B, D), 0.3, dtype=dt), generator=g)
        else:
            xc = torch.randn(B, D, dtype=dt, generator=g) * 2
            xm = torch.randn(B, D, dtype=dt, generator=g) * 2
        noise = dict(enc_cdae=torch.randn(B * nz, n, dtype=dt, generator=g),
                     xi=torch.randn(B, nz * nstd, 1, dtype=dt, generator=g),
                     eps_cdae=torch.randn(B, nz * nstd, d, dtype=dt, generator=g),
                     enc_model=torch.randn(B * nzm, n, dtype=dt, generator=g))
        out = rh.ref_train_step(model, cdae, mopt, copt, xc, xm, noise, hp, do_step=True,
                                ctx_type=c.get('ctx_type', 'lt0'), mnist_like=c['kind'] in ('mnist', 'conv', 'auxmnist'))
        p = 's%d/' % step
        arrays[p + 'x_cdae'], arrays[p + 'x_model'] = xc.numpy(), xm.numpy()
        for k, v in noise.items():
            arrays[p + 'noise/' + k] = v.numpy()
        for k in ('zbar', 'z_cdae', 'std', 'cdae_loss', 'cdae_score', 'model_loss', 'recon', 'prior', 'z_model', 'entropy_grad'):
            arrays[p + k] = out[k].detach().numpy().copy()
        for k, v in t2n(out['cdae_grads']).items():
            if v is not None:
                arrays[p + 'cdae_grads/' + k] = v
        for k, v in t2n(out['model_grads']).items():
            arrays[p + 'model_grads/' + k] = pick(k, v) if sampled else f32(v)
        if name == 'mnist_small' and step == 0:
            # a REAL reference checkpoint pair after the first iteration, written by the reference's own
            # utils.save_checkpoint with the dict layout of ivae_ardae.py:1116-1137 (checkpoint-interchange test:
            # resume on the B200 path, iteration 2 must reproduce s1/*_after)
            import types
            utils, _ = rh.import_reference()
            ck = os.path.join(OUT, 'ref_ckpt_' + name)
            os.makedirs(ck, exist_ok=True)
            o = types.SimpleNamespace(path=ck)
            common = dict(epoch=1, batch_idx=1, train_num_iters_per_epoch=10, best_val_loss=float('inf'), scheduler=None)
            utils.save_checkpoint(dict(common, model='mnist-concat', state_dict=model.state_dict(), optimizer=mopt.state_dict()),
                                  o, is_best=False, filename='model-checkpoint.pth.tar')
            utils.save_checkpoint(dict(common, cdae='mlp-grad', state_dict=cdae.state_dict(), optimizer=copt.state_dict()),
                                  o, is_best=False, filename='cdae-checkpoint.pth.tar')
        if sampled:
            for k, v in model.state_dict().items():
                arrays[p + 'm_after/' + k] = pick(k, v.numpy().copy())
            for k, v in cdae.state_dict().items():
                arrays[p + 'c_after/' + k] = v.numpy().copy()
        if not lite:
            for k, v in model.state_dict().items():
                arrays[p + 'm_after/' + k] = v.numpy().copy()
            for k, v in cdae.state_dict().items():
                arrays[p + 'c_after/' + k] = v.numpy().copy()
    if c.get('cdae_kind', 'grad') == 'grad':
        assert out['cdae_grads']['neglogprob.fc.bias'] is None
    else:
        assert all(v is not None for v in out['cdae_grads'].values())  # residual CDAE: plain back-prop
    # IWS (evaluate_iws) on the stepped weights
    b, S = 3, max(16, 2 * d)
    if lite:  # IWS on the INITIAL weights (the stepped ones are not stored)
        with torch.no_grad():
            model.load_state_dict({k: torch.from_numpy(arrays['m0/' + k]).double() for k in model.state_dict()})
    if c['kind'] in ('mnist', 'conv', 'auxmnist'):
        xe = torch.bernoulli(torch.full((b, D), 0.3, dtype=dt), generator=g)
    else:
        xe = torch.randn(b, D, dtype=dt, generator=g) * 2
    enc_noise = torch.randn(b, S, n, dtype=dt, generator=g)
    eta = torch.randn(b, S, d, dtype=dt, generator=g)
    arrays['iws/x'], arrays['iws/enc_noise'], arrays['iws/eta'] = xe.numpy(), enc_noise.numpy(), eta.numpy()
    arrays['iws/logprob'] = rh.ref_iws(model, xe, enc_noise, eta).numpy()
    meta = dict(kind=c['kind'], model=c['model'], cdae=c['cdae'], B=B, hp=hp, wscale=c['wscale'], iws=dict(b=b, S=S),
                cdae_kind=c.get('cdae_kind', 'grad'), ctx_type=c.get('ctx_type', 'lt0'), steps=nsteps, sampled=sampled,
                lite=lite)
    arrays['meta'] = np.array(json.dumps(meta))
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **arrays)
    last = 's%d/' % (nsteps - 1)
    print('%s: %d arrays, %.1f KB, cdae_loss=%.6g model_loss=%.6g iws=%.6g' % (
        name, len(arrays), os.path.getsize(path) / 1024., float(arrays[last + 'cdae_loss']),
        float(arrays[last + 'model_loss']), float(arrays['iws/logprob'])))


if __name__ == '__main__':
    only = sys.argv[1:]
    for name, c in CASES.items():
        if not only or name in only:
            make_case(name, c)
