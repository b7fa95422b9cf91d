"""Checkpoint interchange and schedule helpers of the training driver.

`save_checkpoint` / `load_checkpoint` read and write the reference's files (`utils/msc.py:67-96`, dict layout of
`ivae_ardae.py:1116-1137`: epoch, batch_idx, train_num_iters_per_epoch, state_dict, best_val_loss, optimizer,
scheduler), so an experiment started with the reference resumes on the B200 path and vice versa: the drop-in modules
keep the reference's state_dict keys and the flat optimizers keep its per-parameter state names (`step`, `exp_avg`,
`exp_avg_sq` for utils.Adam; `step`, `square_avg`, `momentum_buffer` for torch RMSprop).
"""
import os
import shutil

import torch


def annealing_func(val_init, val_fin, val_annealing, step):
    """utils/msc.py:53-55 (beta annealing, ivae_ardae.py:632-641): linear ramp over `val_annealing` steps."""
    if val_annealing is None:
        return float(val_fin)
    return float(val_init + (val_fin - val_init) / float(val_annealing) * float(min(val_annealing, step)))


def _path(opt_or_path):
    return opt_or_path if isinstance(opt_or_path, str) else opt_or_path.path


def save_checkpoint(state, opt, is_best=False, filename='checkpoint.pth.tar'):
    """utils/msc.py:67-72.  `opt`: the reference's options object (uses `.path`) or a directory path."""
    filename = os.path.join(_path(opt), filename)
    torch.save(state, filename)
    if is_best:
        shutil.copyfile(filename, 'model_best.pth.tar')
    return filename


def _cast_like(value, ref):
    return value.to(dtype=ref.dtype) if torch.is_tensor(value) and value.is_floating_point() else value


def load_checkpoint(model, optimizer, opt, filename='checkpoint.pth.tar', verbose=False, device=None, scheduler=None):
    """utils/msc.py:74-96.  Returns the checkpoint dict (None when the file does not exist) and, when `opt` is an
    options object, sets start_epoch / start_batch_idx / best_val_loss / train_num_iters_per_epoch on it like the
    reference.  Tensors are cast to the dtype of the receiving module (the reference may have trained in fp64)."""
    path = os.path.join(_path(opt), filename)
    if not os.path.isfile(path):
        if verbose:
            print("=> no checkpoint found at '{}'".format(path))
        return None
    ck = torch.load(path, map_location=device if device is not None else 'cpu')
    if not isinstance(opt, str):
        opt.start_epoch = ck['epoch']
        opt.start_batch_idx = ck['batch_idx']
        opt.best_val_loss = ck['best_val_loss']
        if 'train_num_iters_per_epoch' in ck:
            opt.train_num_iters_per_epoch = ck['train_num_iters_per_epoch']
        if 'start_std' in ck:
            opt.start_std = ck['start_std']
    if model is not None:
        ref = model.state_dict()
        model.load_state_dict({k: _cast_like(v, ref[k]) for k, v in ck['state_dict'].items()})
    if optimizer is not None:
        sd = ck['optimizer']
        p0 = optimizer.param_groups[0]['params'][0]
        state = {k: {n: (_cast_like(v, p0) if torch.is_tensor(v) and v.dim() > 0 else
                         (int(v) if n == 'step' else v)) for n, v in st.items()}
                 for k, st in sd['state'].items()}
        optimizer.load_state_dict(dict(state=state, param_groups=sd['param_groups']))
    if scheduler is not None and ck.get('scheduler') is not None:
        scheduler.load_state_dict(ck['scheduler'])
    return ck


def load_end_iter(opt, filename='best-checkpoint.pth.tar', verbose=False, device=None):
    """utils/msc.py:98-110: "final"-mode replay (ivae_ardae.py:284-285,699-700,1142,1167).  The iteration index at
    which the best validation checkpoint of the train/val run was written, i.e. where the re-run on train+val data
    stops: (epoch - 1) * train_num_iters_per_epoch + batch_idx - 1."""
    path = os.path.join(_path(opt), filename)
    if not os.path.isfile(path):
        raise ValueError("=> no checkpoint found at '{}'".format(path))
    if verbose:
        print("=> loading checkpoint '{}'".format(path))
    ck = torch.load(path, map_location=device if device is not None else 'cpu')
    i_ep = (ck['epoch'] - 1) * ck['train_num_iters_per_epoch'] + ck['batch_idx']
    return i_ep - 1


class EndIterError(Exception):
    """utils/msc.py:112-113: raised by the training loop when the "final"-mode run reaches `end_iter`."""
    pass


def final_mode_should_stop(i_ep, end_iter, train_mode='final'):
    """The stop test of the reference's loop (ivae_ardae.py:699-700): `i_ep` is the 0-based iteration about to run."""
    return train_mode == 'final' and end_iter is not None and (i_ep + 1) > end_iter
