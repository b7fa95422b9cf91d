"""CPU-only checks of the C ABI: the shared library loads, exports every symbol include/ardae.h declares,
and its host-side argument validation returns error codes + messages (no compute calls: no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'pytorch-ardae-vae_b200', 'ardae', 'libardae.so')
HDR = os.path.join(ROOT, 'include', 'ardae.h')


@pytest.fixture(scope='module')
def lib():
    if not os.path.isfile(LIB):
        import subprocess
        subprocess.check_call(['make', '-C', os.path.join(ROOT, 'pytorch-ardae-vae_b200', 'csrc')])
    return ctypes.CDLL(LIB)


def declared_symbols():
    src = open(HDR).read()
    return sorted(set(re.findall(r'ARDAE_API\s+[\w\s\*]+?\b(ardae_\w+)\s*\(', src)))


def test_header_declares_expected_entry_points():
    names = declared_symbols()
    for must in ('ardae_cdae_train', 'ardae_cdae_score', 'ardae_model_encode', 'ardae_model_forward',
                 'ardae_model_backward', 'ardae_adam_step', 'ardae_rmsprop_step', 'ardae_sigma_schedule'):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_version_and_error_reporting(lib):
    lib.ardae_last_error.restype = ctypes.c_char_p
    assert lib.ardae_version() == 100

    class Cfg(ctypes.Structure):
        _fields_ = [(k, ctypes.c_int) for k in ('input_dim', 'context_dim', 'h_dim', 'num_hidden_layers', 'batch',
                                                'samples', 'train', 'kind')]
    n = ctypes.c_size_t(0)
    # a valid config: pure host-side dry build of the plan
    assert lib.ardae_cdae_workspace_bytes(ctypes.byref(Cfg(32, 32, 256, 5, 512, 256, 1)), ctypes.byref(n)) == 0
    arr16 = 131072 * 256 * 2  # one [N, H] bf16 spill array
    assert 8 * 5 * arr16 < n.value < 10 * 5 * arr16  # 8 L bf16 spill arrays (u, v, delta x2, tangent x2, t/adj x2) + small change
    train_bytes = n.value
    assert lib.ardae_cdae_workspace_bytes(ctypes.byref(Cfg(32, 32, 256, 5, 512, 256, 0)), ctypes.byref(n)) == 0
    assert 0 < n.value < 2 * train_bytes  # score-only plan: 4 L fp32 arrays
    # invalid configs -> negative code + message, no crash
    assert lib.ardae_cdae_workspace_bytes(ctypes.byref(Cfg(32, 32, 256, 1, 512, 256, 1)), ctypes.byref(n)) < 0
    assert b'num_hidden_layers' in lib.ardae_last_error()
    assert lib.ardae_cdae_workspace_bytes(ctypes.byref(Cfg(32, 32, 250, 5, 512, 256, 1)), ctypes.byref(n)) < 0
    assert b'multiple of 4' in lib.ardae_last_error()
    assert lib.ardae_cdae_workspace_bytes(None, ctypes.byref(n)) < 0
    # residual CDAE (kind 1): fp32 spill, but no tangent / adjoint arrays
    assert lib.ardae_cdae_workspace_bytes(ctypes.byref(Cfg(32, 32, 256, 5, 512, 256, 1, 1)), ctypes.byref(n)) == 0
    assert 0 < n.value < 2 * train_bytes
    assert lib.ardae_cdae_workspace_bytes(ctypes.byref(Cfg(32, 32, 256, 5, 512, 256, 1, 2)), ctypes.byref(n)) < 0


def test_model_workspace_query(lib):
    class MCfg(ctypes.Structure):
        _fields_ = [(k, ctypes.c_int) for k in ('kind', 'input_dim', 'noise_dim', 'h_dim', 'z_dim', 'n_inp', 'n_fc',
                                                'n_dec', 'act', 'batch', 'nz', 'mode')]
    n = ctypes.c_size_t(0)
    assert lib.ardae_model_workspace_bytes(ctypes.byref(MCfg(1, 784, 100, 300, 32, 4, 1, 3, 1, 512, 1, 1)), ctypes.byref(n)) == 0
    assert n.value > 0
    assert lib.ardae_model_workspace_bytes(ctypes.byref(MCfg(0, 2, 10, 256, 2, 2, 2, 2, 0, 512, 256, 0)), ctypes.byref(n)) == 0
    assert lib.ardae_model_workspace_bytes(ctypes.byref(MCfg(7, 2, 10, 256, 2, 2, 2, 2, 0, 512, 256, 0)), ctypes.byref(n)) < 0
    # empty batches are refused (the reference would fail inside torch.std over an empty dimension)
    assert lib.ardae_model_workspace_bytes(ctypes.byref(MCfg(0, 2, 10, 256, 2, 2, 2, 2, 0, 0, 256, 0)), ctypes.byref(n)) < 0
    assert lib.ardae_model_workspace_bytes(ctypes.byref(MCfg(0, 2, 10, 256, 2, 2, 2, 2, 0, 512, 0, 0)), ctypes.byref(n)) < 0


def test_hierarchical_and_conv_kinds_plan_on_the_host(lib):
    """kind 2 (ConvIPVAE) and kind 3 (MNISTAuxIPVAE): the dry build of every supported mode sizes a workspace; bad modes
    are refused with a message (no GPU needed)."""
    class MCfg(ctypes.Structure):
        _fields_ = [(k, ctypes.c_int) for k in ('kind', 'input_dim', 'noise_dim', 'h_dim', 'z_dim', 'n_inp', 'n_fc',
                                                'n_dec', 'act', 'batch', 'nz', 'mode', 'img_h', 'img_c')]
    lib.ardae_last_error.restype = ctypes.c_char_p
    n = ctypes.c_size_t(0)
    sizes = {}
    for mode, nz in ((0, 256), (1, 1), (2, 64), (3, 1)):
        cfg = MCfg(3, 784, 100, 300, 32, 2, 2, 2, 1, 64, nz, mode, 0, 0)
        assert lib.ardae_model_workspace_bytes(ctypes.byref(cfg), ctypes.byref(n)) == 0, (mode, lib.ardae_last_error())
        sizes[mode] = n.value
    assert sizes[0] > sizes[3] > 0 and sizes[1] > sizes[3]
    assert lib.ardae_model_workspace_bytes(ctypes.byref(MCfg(3, 784, 100, 300, 32, 2, 2, 2, 1, 64, 1, 4, 0, 0)), ctypes.byref(n)) < 0
    assert b'auxmnist' in lib.ardae_last_error()
    assert lib.ardae_model_workspace_bytes(ctypes.byref(MCfg(3, 784, 100, 300, 32, 2, 2, 2, 0, 64, 1, 1, 0, 0)), ctypes.byref(n)) < 0
    assert b'softplus' in lib.ardae_last_error()
    for mode, nz in ((0, 256), (1, 1), (2, 64), (3, 1), (4, 1), (5, 1)):
        cfg = MCfg(2, 784, 100, 800, 32, 3, 1, 2, 1, 32, nz, mode, 28, 1)
        assert lib.ardae_model_workspace_bytes(ctypes.byref(cfg), ctypes.byref(n)) == 0, (mode, lib.ardae_last_error())
    assert lib.ardae_model_workspace_bytes(ctypes.byref(MCfg(2, 784, 100, 800, 32, 3, 1, 2, 1, 32, 1, 1, 27, 1)), ctypes.byref(n)) < 0


def test_data_parallel_entry_points_validate_arguments(lib):
    lib.ardae_last_error.restype = ctypes.c_char_p
    n = ctypes.c_size_t(0)
    assert lib.ardae_dp_xchg_bytes(ctypes.c_size_t(1 << 20), 8, ctypes.byref(n)) == 0
    # header + 8 slots of one slice + the parameter area = 2 x the arena + change
    assert 2 * 4 * (1 << 20) <= n.value <= 2 * 4 * (1 << 20) + 8192
    assert lib.ardae_dp_xchg_bytes(ctypes.c_size_t(1 << 20), 17, ctypes.byref(n)) < 0
    assert lib.ardae_dp_xchg_bytes(ctypes.c_size_t(1001), 2, ctypes.byref(n)) < 0
    lib.ardae_dp_fused_step.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 4 + [ctypes.c_size_t, ctypes.c_void_p,
                                                                                     ctypes.c_void_p] + [ctypes.c_float] * 5 + [
        ctypes.c_int, ctypes.c_float, ctypes.c_void_p]
    assert lib.ardae_dp_fused_step(0, 0, 2, None, None, None, None, 16, None, None, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, 1.0, None) < 0
    assert b'null' in lib.ardae_last_error()
    lib.ardae_ipc_export.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_size_t)]
    assert lib.ardae_ipc_export(None, ctypes.create_string_buffer(64), ctypes.byref(n)) < 0


def test_no_gpu_means_loud_failure(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    assert lib.ardae_check_device(0) != 0  # CUDA error code, not a crash and not a silent fallback


def header_struct_fields(name):
    """Field names of `typedef struct { ... } name;` in include/ardae.h (all members are `int`)."""
    src = open(HDR).read()
    m = re.search(r'typedef\s+struct\s*\{([^{}]*)\}\s*' + name + r'\s*;', src, re.S)
    assert m, name
    body = re.sub(r'/\*.*?\*/', '', m.group(1), flags=re.S)
    fields = []
    for decl in body.split(';'):
        decl = decl.strip()
        if not decl:
            continue
        assert decl.startswith('int '), decl
        fields += [f.strip() for f in decl[4:].split(',')]
    return fields


def test_ctypes_structs_match_the_header():
    """The ctypes mirrors (product binding, INTEGRATION.md stub) must have exactly the header's fields, in order: a short
    struct makes the library read past the caller's memory (round-1 INTEGRATION.md declared 7 of the 8 ints)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, 'pytorch-ardae-vae_b200'))
    from ardae import _lib
    cfields = header_struct_fields('ardae_cdae_config')
    mfields = header_struct_fields('ardae_model_config')
    assert [f for f, _ in _lib.CdaeConfig._fields_] == cfields
    assert [f for f, _ in _lib.ModelConfig._fields_] == mfields
    assert all(t is ctypes.c_int for _, t in _lib.CdaeConfig._fields_ + _lib.ModelConfig._fields_)
    assert ctypes.sizeof(_lib.CdaeConfig) == 4 * len(cfields) and ctypes.sizeof(_lib.ModelConfig) == 4 * len(mfields)
    # the stub a reference maintainer would paste
    doc = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    m = re.search(r'class CdaeCfg\(ctypes\.Structure\):\s*_fields_ = \[\(k, ctypes\.c_int\) for k in\s*\((.*?)\)\]', doc, re.S)
    assert m, 'INTEGRATION.md: CdaeCfg stub not found'
    assert re.findall(r"'(\w+)'", m.group(1)) == cfields


def test_load_end_iter(tmp_path):
    """utils/msc.py:98-110 on a checkpoint with the reference's dict layout (ivae_ardae.py:1116-1137)."""
    import sys
    import torch
    sys.path.insert(0, os.path.join(ROOT, 'pytorch-ardae-vae_b200'))
    import ardae
    torch.save(dict(epoch=3, batch_idx=17, train_num_iters_per_epoch=100, best_val_loss=1.0, state_dict={}, optimizer={}),
               os.path.join(str(tmp_path), 'best-model-checkpoint.pth.tar'))
    end = ardae.load_end_iter(str(tmp_path), filename='best-model-checkpoint.pth.tar')
    assert end == (3 - 1) * 100 + 17 - 1
    assert not ardae.final_mode_should_stop(end - 1, end) and ardae.final_mode_should_stop(end, end)
    assert not ardae.final_mode_should_stop(10 ** 9, end, train_mode='train')
    with pytest.raises(ValueError):
        ardae.load_end_iter(str(tmp_path), filename='missing.pth.tar')
    assert issubclass(ardae.EndIterError, Exception)
