"""CPU-only checks of the host-side mirror of the reference interface: constructor
signatures, state_dict keys / shapes identical to the reference's (taken from the reference-generated
fixtures), init semantics, error behaviour, and that nothing silently falls back to CPU compute."""
import numpy as np
import pytest
import torch

from golden_util import load_case, sub


def build(meta):
    import ardae
    import golden_util
    c = meta['cdae']
    model = golden_util.build_model(meta)
    cdae = golden_util.build_cdae(meta)
    return model, cdae


@pytest.mark.parametrize('name', ['toy_small', 'mnist_small', 'conv_small', 'mnist_small_res'])
def test_state_dict_layout_matches_reference(name):
    z, meta = load_case(name)
    model, cdae = build(meta)
    for mod, pref in ((model, 'm0/'), (cdae, 'c0/')):
        ref = sub(z, pref)
        sd = mod.state_dict()
        assert list(sd.keys()) == [k[len(pref):] for k in z.files if k.startswith(pref)]  # same order too
        for k, v in sd.items():
            assert tuple(v.shape) == tuple(ref[k].shape), k
        # a reference checkpoint loads without key or shape errors
        mod.load_state_dict({k: torch.from_numpy(np.asarray(v)).float() for k, v in ref.items()})


def test_full_size_parameter_counts():
    """config-2 model 1 062 816 params, CDAE 938 241; config-1 271 386 / 528 129."""
    import ardae
    m2 = ardae.MNISTIPVAE(input_dim=784, noise_dim=100, h_dim=300, num_hidden_layers=2, nonlinearity='softplus',
                          enc_type='concat', z_dim=32)
    c2 = ardae.MLPGradCARDAE(input_dim=32, context_dim=32, std=1., h_dim=256, num_hidden_layers=5, nonlinearity='softplus')
    m1 = ardae.ToyIPVAE(input_dim=2, noise_dim=10, h_dim=256, num_hidden_layers=2, nonlinearity='relu',
                        enc_type='concat', z_dim=2)
    c1 = ardae.MLPGradCARDAE(input_dim=2, context_dim=2, std=1., h_dim=256, num_hidden_layers=3, nonlinearity='softplus')
    n = lambda mod: sum(p.numel() for p in mod.parameters())
    assert (n(m2), n(c2), n(m1), n(c1)) == (1062816, 938241, 271386, 528129)
    m4 = ardae.ConvIPVAE(input_height=28, input_channels=1, z_dim=32, noise_dim=100, nonlinearity='softplus')
    assert n(m4) == 757773


def test_init_semantics():
    """toy.py:189-190,719-720 / mnist.py:20-25,158-159: normal_ on encode.fc.fc.weight (and toy mean_fn),
    xavier + zero bias on the whole MNIST decoder."""
    import ardae
    torch.manual_seed(0)
    m = ardae.MNISTIPVAE(input_dim=784, noise_dim=100, h_dim=300, num_hidden_layers=2, nonlinearity='softplus', z_dim=32)
    assert abs(m.encode.fc.fc.weight.std().item() - 1.0) < 0.05
    assert all(float(l.bias.detach().abs().max()) == 0.0 for l in m.decode.main.linears())
    assert float(m.decode.reparam.logit_fn.bias.detach().abs().max()) == 0.0
    t = ardae.ToyIPVAE(input_dim=2, noise_dim=10, h_dim=256, num_hidden_layers=2, nonlinearity='relu', z_dim=2)
    assert abs(t.decode.reparam.mean_fn.weight.std().item() - 1.0) < 0.15


def test_unsupported_configurations_raise():
    import ardae
    with pytest.raises(NotImplementedError):
        ardae.MLPGradCARDAE(input_dim=2, context_dim=2, std=1., h_dim=16, num_hidden_layers=3, nonlinearity='tanh')
    with pytest.raises(NotImplementedError):
        ardae.ToyIPVAE(input_dim=2, noise_dim=2, h_dim=16, num_hidden_layers=2, nonlinearity='relu', enc_type='scale', z_dim=2)
    t = ardae.ToyIPVAE(input_dim=2, noise_dim=2, h_dim=16, num_hidden_layers=2, nonlinearity='relu', z_dim=2)
    with pytest.raises(NotImplementedError):  # same as the reference for lmbd > 0 (toy.py:845-846)
        t.forward(torch.zeros(2, 2), lmbd=1.0)


def test_cpu_tensors_fail_loudly():
    """No CPU fallback: the product path refuses to compute without the CUDA library / device."""
    import ardae
    z, meta = load_case('toy_small')
    model, cdae = build(meta)
    with pytest.raises(RuntimeError):
        cdae(torch.zeros(2, 3, 2), torch.zeros(2, 1, 2), std=torch.zeros(2, 3, 1))
    with pytest.raises(RuntimeError):
        model(torch.zeros(2, 2))
    with pytest.raises(RuntimeError):
        model.encode(torch.zeros(2, 2), std=0)


def test_product_package_never_imports_oracle():
    import os
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'pytorch-ardae-vae_b200')
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dp, f)).read()
                assert 'ardae_oracle' not in src and 'ref_harness' not in src, f


def test_optimizer_api():
    import ardae
    z, meta = load_case('toy_small')
    model, cdae = build(meta)
    mo = ardae.Adam(model.parameters(), lr=1e-4, betas=(0.5, 0.999))
    co = ardae.RMSprop(cdae.parameters(), lr=1e-4, momentum=0.5)
    assert mo.param_groups[0]['lr'] == 1e-4 and co.param_groups[0]['momentum'] == 0.5
    assert 'state' in mo.state_dict() and 'param_groups' in co.state_dict()
    with pytest.raises(NotImplementedError):
        ardae.Adam(model.parameters(), amsgrad=True)


def test_reference_checkpoint_loads_and_annealing():
    """Checkpoint interchange: the files under tests/golden/ref_ckpt_mnist_small were written by the
    reference's own utils.save_checkpoint after one iteration (oracle/make_golden.py)."""
    import os
    import ardae
    from golden_util import GOLDEN_DIR
    z, meta = load_case('mnist_small')
    model, cdae = build(meta)
    hp = meta['hp']
    mo = ardae.Adam(model.parameters(), lr=hp['m_lr'], betas=(hp['m_beta1'], 0.999))
    co = ardae.RMSprop(cdae.parameters(), lr=hp['d_lr'], momentum=hp['d_momentum'])
    ck_dir = os.path.join(GOLDEN_DIR, 'ref_ckpt_mnist_small')

    class Opt(object):
        path = ck_dir
    o = Opt()
    ck = ardae.load_checkpoint(model, mo, o, filename='model-checkpoint.pth.tar')
    assert ck is not None and (o.start_epoch, o.start_batch_idx, o.train_num_iters_per_epoch) == (1, 1, 10)
    ardae.load_checkpoint(cdae, co, ck_dir, filename='cdae-checkpoint.pth.tar')
    for k, v in model.state_dict().items():
        assert v.dtype == torch.float32
        assert np.allclose(v.numpy(), z['s0/m_after/' + k], rtol=1e-6, atol=1e-7), k
    for k, v in cdae.state_dict().items():
        assert np.allclose(v.numpy(), z['s0/c_after/' + k], rtol=1e-6, atol=1e-7), k
    p0 = next(iter(model.parameters()))
    assert set(mo.state[p0]) >= {'step', 'exp_avg', 'exp_avg_sq'} and mo.state[p0]['step'] == 1
    assert mo.state[p0]['exp_avg'].dtype == torch.float32
    q0 = next(iter(cdae.parameters()))
    assert set(co.state[q0]) >= {'step', 'square_avg', 'momentum_buffer'} and co.state[q0]['step'] == 1
    assert list(cdae.parameters())[-1] not in co.state  # neglogprob.fc.bias never had a gradient: no state in the reference
    assert ardae.load_checkpoint(model, None, ck_dir, filename='missing.pth.tar') is None
    # beta annealing (utils/msc.py:53-55)
    assert ardae.annealing_func(1e-4, 1.0, 50000, 0) == pytest.approx(1e-4)
    assert ardae.annealing_func(1e-4, 1.0, 50000, 25000) == pytest.approx(1e-4 + (1.0 - 1e-4) * 0.5)
    assert ardae.annealing_func(1e-4, 1.0, 50000, 10 ** 9) == pytest.approx(1.0)
    assert ardae.annealing_func(1e-4, 1.0, None, 3) == 1.0


def test_checkpoint_written_here_loads_in_the_reference(tmp_path):
    """The other direction of the interchange: a checkpoint saved by ardae.save_checkpoint (state that came from the
    reference's files) is read back by the REFERENCE's own utils.load_checkpoint into reference modules / optimizers.
    Needs the reference tree (build container only)."""
    import os
    import sys
    import types
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle'))
    import ref_harness as rh
    if rh.find_reference() is None:
        pytest.skip('reference tree not available')
    import ardae
    from golden_util import GOLDEN_DIR
    z, meta = load_case('mnist_small')
    hp = meta['hp']
    model, cdae = build(meta)
    mo = ardae.Adam(model.parameters(), lr=hp['m_lr'], betas=(hp['m_beta1'], 0.999))
    co = ardae.RMSprop(cdae.parameters(), lr=hp['d_lr'], momentum=hp['d_momentum'])
    ck_dir = os.path.join(GOLDEN_DIR, 'ref_ckpt_mnist_small')
    ardae.load_checkpoint(model, mo, ck_dir, filename='model-checkpoint.pth.tar')
    ardae.load_checkpoint(cdae, co, ck_dir, filename='cdae-checkpoint.pth.tar')
    out = str(tmp_path)
    common = dict(epoch=3, batch_idx=7, train_num_iters_per_epoch=10, best_val_loss=1.5, scheduler=None)
    ardae.save_checkpoint(dict(common, model='mnist-concat', state_dict=model.state_dict(), optimizer=mo.state_dict()),
                          out, filename='model-checkpoint.pth.tar')
    ardae.save_checkpoint(dict(common, cdae='mlp-grad', state_dict=cdae.state_dict(), optimizer=co.state_dict()),
                          out, filename='cdae-checkpoint.pth.tar')
    utils, _ = rh.import_reference()
    rmodel, rcdae = rh.build_reference('mnist', meta['model'], meta['cdae'], seed=99)
    rmo, rco = rh.build_optimizers(rmodel, rcdae, hp)
    o = types.SimpleNamespace(path=out)
    utils.load_checkpoint(rmodel, rmo, o, filename='model-checkpoint.pth.tar', verbose=False)
    utils.load_checkpoint(rcdae, rco, o, filename='cdae-checkpoint.pth.tar', verbose=False)
    assert (o.start_epoch, o.start_batch_idx, o.best_val_loss) == (3, 7, 1.5)
    for k, v in rmodel.state_dict().items():
        assert np.allclose(v.numpy(), z['s0/m_after/' + k], rtol=1e-6, atol=1e-7), k
    for k, v in rcdae.state_dict().items():
        assert np.allclose(v.numpy(), z['s0/c_after/' + k], rtol=1e-6, atol=1e-7), k
    p0 = next(iter(rmodel.parameters()))
    assert int(rmo.state[p0]['step']) == 1 and rmo.state[p0]['exp_avg'].abs().sum() > 0
    q0 = next(iter(rcdae.parameters()))
    assert rco.state[q0]['square_avg'].abs().sum() > 0 and rco.state[q0]['momentum_buffer'].abs().sum() > 0
