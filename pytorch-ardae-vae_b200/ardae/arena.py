"""Flat parameter / gradient arenas.

All parameters of a module live in ONE contiguous fp32 device buffer (each tensor starting on a
256-byte boundary, as TMA and the vectorised optimizer want); `.data` of every nn.Parameter is a view
into it, `.grad` a view into a second buffer of the same layout.  This is what lets the optimizer be
a single elementwise pass (28 B/param) and the data-parallel gradient exchange a single allreduce.
"""
import torch

ALIGN = 64  # floats


class ParamArena(object):
    def __init__(self, module):
        self.module = module
        self.names, self.params = zip(*list(module.named_parameters()))
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        self.total = off
        import weakref
        for p in self.params:
            p._ardae_owner = weakref.ref(module)
        self.flat = None
        self.grad_flat = None
        self.stage_flat = None

    def device_ok(self):
        p0 = self.params[0]
        if self.flat is None or self.flat.device != p0.device:
            return False
        base = self.flat.data_ptr()
        return all(p.data_ptr() == base + 4 * o for p, o in zip(self.params, self.offsets))

    def ensure(self):
        """(Re)build the arenas if parameters were moved / re-materialised (e.g. by .to())."""
        if self.device_ok():
            return
        dev = self.params[0].device
        if dev.type != 'cuda':
            raise RuntimeError('AR-DAE modules must live on a CUDA device (no CPU fallback)')
        flat = torch.zeros(self.total, dtype=torch.float32, device=dev)
        for p, o in zip(self.params, self.offsets):
            if p.dtype != torch.float32:
                raise RuntimeError('AR-DAE modules are float32 only')
            v = flat[o:o + p.numel()].view(p.shape)
            v.copy_(p.data)
            p.data = v
        self.flat = flat
        self.grad_flat = torch.zeros_like(flat)
        self.stage_flat = torch.zeros_like(flat)
        for p in self.params:
            p.grad = None

    def view(self, flat, k):
        p, o = self.params[k], self.offsets[k]
        return flat[o:o + p.numel()].view(p.shape)

    def views(self, flat):
        return [self.view(flat, k) for k in range(len(self.params))]

    def grads_aliased(self):
        base = self.grad_flat.data_ptr()
        return all(p.grad is not None and p.grad.data_ptr() == base + 4 * o
                   for p, o in zip(self.params, self.offsets))

    def accumulate_staged(self, scale, skip=()):
        """grad += scale * staged, with .grad views (re)attached to the grad arena."""
        none = [p.grad is None for p in self.params]
        if all(none):
            self.grad_flat.zero_()
        for k, p in enumerate(self.params):
            if k in skip:
                continue
            g = self.view(self.grad_flat, k)
            if p.grad is None:
                if not all(none):
                    g.zero_()
                p.grad = g
            elif p.grad.data_ptr() != g.data_ptr():
                g.copy_(p.grad)
                p.grad = g
        if torch.is_tensor(scale):
            self.grad_flat.addcmul_(self.stage_flat, scale.reshape(1).expand_as(self.stage_flat))
        else:
            self.grad_flat.add_(self.stage_flat, alpha=float(scale))
