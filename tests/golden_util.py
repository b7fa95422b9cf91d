"""Helpers shared by the golden-fixture tests."""
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = ['toy_small', 'mnist_small', 'mnist_small_x3', 'conv_small', 'mnist_small_res', 'mnist_small_datactx', 'conv28', 'auxmnist_small']


def model_dims(meta):
    """(input_dim, noise_dim, h_dim, z_dim, num_hidden_layers, nonlinearity) for any model kind."""
    m = meta['model']
    if meta['kind'] == 'conv':
        return (m['input_channels'] * m['input_height'] ** 2, m['noise_dim'], 800, m['z_dim'], 0, m['nonlinearity'])
    return (m['input_dim'], m['noise_dim'], m['h_dim'], m['z_dim'], m['num_hidden_layers'], m['nonlinearity'])


def build_model(meta):
    """The drop-in model class of the product package for a fixture's meta."""
    import ardae
    m = meta['model']
    if meta['kind'] == 'conv':
        return ardae.ConvIPVAE(input_height=m['input_height'], input_channels=m['input_channels'], z_dim=m['z_dim'],
                               noise_dim=m['noise_dim'], nonlinearity=m['nonlinearity'])
    if meta['kind'] == 'auxmnist':
        return ardae.MNISTAuxIPVAE(input_dim=m['input_dim'], noise_dim=m['noise_dim'], h_dim=m['h_dim'],
                                   num_hidden_layers=m['num_hidden_layers'], nonlinearity=m['nonlinearity'],
                                   enc_type='simple', z_dim=m['z_dim'])
    cls = ardae.ToyIPVAE if meta['kind'] == 'toy' else ardae.MNISTIPVAE
    return cls(input_dim=m['input_dim'], noise_dim=m['noise_dim'], h_dim=m['h_dim'],
               num_hidden_layers=m['num_hidden_layers'], nonlinearity=m['nonlinearity'], enc_type='concat',
               z_dim=m['z_dim'])


def cdae_spec(meta):
    """oracle CdaeSpec for a fixture (meta['cdae_kind']: 'grad' = mlp-grad, 'res' = mlp-res)."""
    import ardae_oracle as orc
    c = meta['cdae']
    return orc.CdaeSpec(c['input_dim'], c['context_dim'], c['h_dim'], c['num_hidden_layers'],
                        kind=meta.get('cdae_kind', 'grad'))


def build_cdae(meta):
    """The drop-in CDAE class of the product package for a fixture's meta."""
    import ardae
    c = meta['cdae']
    cls = ardae.MLPGradCARDAE if meta.get('cdae_kind', 'grad') == 'grad' else ardae.MLPResCARDAE
    return cls(input_dim=c['input_dim'], context_dim=c['context_dim'], std=1., h_dim=c['h_dim'],
               num_hidden_layers=c['num_hidden_layers'], nonlinearity=c['nonlinearity'], noise_type='gaussian',
               enc_ctx=True, enc_input=True)


def hp_of(meta):
    """Hyper-parameters of a fixture incl. the CDAE context type (ivae_ardae.py:729-741)."""
    return dict(meta['hp'], ctx_type=meta.get('ctx_type', 'lt0'))


def load_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    meta = json.loads(str(z['meta']))
    return z, meta


def is_lite(name, meta):
    """lite fixtures store float32 weights / gradients (looser oracle tolerance); older files carry no flag."""
    return bool(meta.get('lite', name == 'conv_small'))


def num_steps(name, meta):
    return int(meta.get('steps', 1 if is_lite(name, meta) else 2))


def pick(z, name, a):
    """The stored view of a model tensor: `sampled` fixtures keep big tensors at fixed random positions."""
    k = 'sample_idx/' + name
    return np.asarray(a).ravel()[z[k]] if k in z.files else np.asarray(a)


def sub(z, prefix):
    return {k[len(prefix):]: z[k] for k in z.files if k.startswith(prefix)}


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def cosine(a, b):
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    return float(a @ b / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-300))
