"""ctypes binding of libardae.so (C ABI declared in include/ardae.h).

There is deliberately no fallback: if the CUDA library is missing or the device is not a B200
every compute call raises.  PyTorch is used only for device memory, streams and autograd plumbing.
"""
import ctypes
import os

import torch

_LIB = None
_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libardae.so')

c_float_p = ctypes.POINTER(ctypes.c_float)


class CdaeConfig(ctypes.Structure):
    _fields_ = [('input_dim', ctypes.c_int), ('context_dim', ctypes.c_int), ('h_dim', ctypes.c_int),
                ('num_hidden_layers', ctypes.c_int), ('batch', ctypes.c_int), ('samples', ctypes.c_int),
                ('train', ctypes.c_int), ('kind', ctypes.c_int)]


class ModelConfig(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int) for k in ('kind', 'input_dim', 'noise_dim', 'h_dim', 'z_dim', 'n_inp', 'n_fc',
                                            'n_dec', 'act', 'batch', 'nz', 'mode', 'img_h', 'img_c')]


def _declare(lib):
    vp, i, sz, u64, u32, f = (ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_uint64,
                              ctypes.c_uint32, ctypes.c_float)
    lib.ardae_version.restype = i
    lib.ardae_last_error.restype = ctypes.c_char_p
    lib.ardae_check_device.argtypes = [i]
    lib.ardae_cdae_workspace_bytes.argtypes = [ctypes.POINTER(CdaeConfig), ctypes.POINTER(sz)]
    lib.ardae_cdae_create.argtypes = [ctypes.POINTER(CdaeConfig), ctypes.POINTER(vp), ctypes.POINTER(vp), i, vp, sz,
                                      ctypes.POINTER(vp)]
    lib.ardae_cdae_destroy.argtypes = [vp]
    lib.ardae_cdae_destroy.restype = None
    lib.ardae_cdae_num_launches.argtypes = [vp]
    lib.ardae_cdae_set_profile.argtypes = [vp, i]
    lib.ardae_cdae_read_profile.argtypes = [vp, i, ctypes.c_char_p, ctypes.POINTER(f), ctypes.POINTER(i)]
    lib.ardae_cdae_train.argtypes = [vp, vp, vp, vp, vp, i, u64, f, vp, vp, vp]
    lib.ardae_cdae_score.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.ardae_randn.argtypes = [vp, sz, u64, u32, vp]
    lib.ardae_model_workspace_bytes.argtypes = [ctypes.POINTER(ModelConfig), ctypes.POINTER(sz)]
    lib.ardae_model_create.argtypes = [ctypes.POINTER(ModelConfig), ctypes.POINTER(vp), ctypes.POINTER(vp), i, vp,
                                       sz, ctypes.POINTER(vp)]
    lib.ardae_model_destroy.argtypes = [vp]
    lib.ardae_model_destroy.restype = None
    lib.ardae_model_num_launches.argtypes = [vp, i]
    lib.ardae_model_encode.argtypes = [vp, vp, vp, vp, vp]
    lib.ardae_model_encode_with_mean.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.ardae_model_encode_hidden.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.ardae_model_forward.argtypes = [vp, vp, vp, f, f, vp, vp, vp, vp]
    lib.ardae_model_backward.argtypes = [vp, f, vp, f, vp]
    lib.ardae_model_set_beta_device.argtypes = [vp, vp]
    lib.ardae_model_decode.argtypes = [vp, vp, vp, vp]
    lib.ardae_model_forward_inp.argtypes = [vp, vp, vp, vp]
    lib.ardae_model_forward_all.argtypes = [vp, vp, vp, vp, vp]
    lib.ardae_model_backward_decoder.argtypes = [vp, f, vp]
    lib.ardae_model_backward_encoder.argtypes = [vp, f, vp, f, vp]
    lib.ardae_model_iws.argtypes = [vp, vp, vp, vp, u64, vp, vp, vp, vp]
    lib.ardae_sigma_schedule.argtypes = [vp, vp, i, i, i, i, f, f, vp, u64, vp, vp, vp, vp]
    lib.ardae_scaled_diff.argtypes = [vp, vp, i, i, i, f, vp, vp]
    lib.ardae_adam_step.argtypes = [vp, vp, vp, vp, sz, f, f, f, f, i, f, vp]
    lib.ardae_bernoulli.argtypes = [vp, vp, sz, u64, vp]
    lib.ardae_set_replay_counter.argtypes = [vp]
    lib.ardae_bump_replay_counter.argtypes = [vp, vp]
    lib.ardae_rmsprop_step.argtypes = [vp, vp, vp, vp, sz, f, f, f, f, f, vp]
    lib.ardae_ipc_export.argtypes = [vp, ctypes.c_char_p, ctypes.POINTER(sz)]
    lib.ardae_ipc_import.argtypes = [ctypes.c_char_p, sz, ctypes.POINTER(vp)]
    lib.ardae_dp_xchg_bytes.argtypes = [sz, i, ctypes.POINTER(sz)]
    lib.ardae_dp_fused_step.argtypes = [i, i, i, vp, vp, vp, vp, sz, ctypes.POINTER(vp), vp, f, f, f, f, f, i, f, vp]
    return lib


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.isfile(_LIB_PATH):
            raise RuntimeError('libardae.so not built (%s). Run `python -c "import __graft_entry__ as g; g.build()"` '
                               'or `make -C pytorch-ardae-vae_b200/csrc`. There is no CPU fallback.' % _LIB_PATH)
        _LIB = _declare(ctypes.CDLL(_LIB_PATH))
    return _LIB


def check(rc):
    if rc != 0:
        raise RuntimeError('libardae error %d: %s' % (rc, lib().ardae_last_error().decode()))


def require_cuda(t, name):
    if not (torch.is_tensor(t) and t.is_cuda and t.dtype == torch.float32):
        raise RuntimeError('%s must be a float32 CUDA tensor (the AR-DAE path has no CPU fallback)' % name)
    return t.contiguous()


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for k, t in enumerate(tensors):
        arr[k] = t.data_ptr()
    return arr


def read_cdae_profile(handle, max_ops=512):
    """[(tag, ms)] for every launch of the last call of a CDAE plan (ardae_cdae_set_profile must be on)."""
    tags = ctypes.create_string_buffer(16 * max_ops)
    ms = (ctypes.c_float * max_ops)()
    n = ctypes.c_int(0)
    check(lib().ardae_cdae_read_profile(handle, max_ops, tags, ms, ctypes.byref(n)))
    raw = tags.raw
    return [(raw[16 * k:16 * k + 16].split(b'\0')[0].decode(), float(ms[k])) for k in range(n.value)]
