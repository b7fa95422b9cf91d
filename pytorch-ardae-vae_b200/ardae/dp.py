"""Peer-memory communicator of the data-parallel step (one process per GPU, one node).

`PeerComm` owns this rank's exchange buffer, opens every peer's buffer through CUDA IPC (handles travel through
torch.distributed) and launches `ardae_dp_fused_step`: gradient exchange + optimizer update as ONE kernel per arena
(csrc/dp_fused.cuh) instead of an NCCL allreduce followed by an optimizer launch.  No collective library call remains
inside the iteration, so the whole iteration is one CUDA graph on every rank.

Parameters stay replicated bit for bit; optimizer state is advanced only on the rank that owns a slice, so
`gather_state` must run before the optimizer state is saved or inspected."""
import ctypes
import os

import torch

from . import _lib


_COMMS = {}      # (process-group id, device index) -> PeerComm: one exchange buffer per process, shared by all TrainSteps
_IMPORTED = {}   # IPC handle bytes -> base pointer in this process (an allocation is opened once)


class PeerCommUnavailable(RuntimeError):
    """Raised on EVERY rank when some rank could not map its peers (the caller then uses the NCCL path)."""


class PeerComm(object):
    @staticmethod
    def get(process_group, device, max_floats):
        """The communicator of (process_group, device), created on first use or when a larger arena shows up.
        Collective: every rank calls it with the same sizes in the same order."""
        key = (id(process_group), torch.device(device).index)
        c = _COMMS.get(key)
        if c is None or c.max_floats < max_floats:
            c = PeerComm(process_group, device, max_floats)
            _COMMS[key] = c          # (a replaced communicator stays mapped: peers may still hold its handle)
        return c

    def __init__(self, process_group, device, max_floats):
        dist = torch.distributed
        self.pg = process_group
        self.world = dist.get_world_size(process_group)
        self.rank = dist.get_rank(process_group)
        self.device = device
        if self.world > 16:
            raise PeerCommUnavailable('PeerComm: at most 16 ranks (one NVSwitch domain)')
        L = _lib.lib()
        if max_floats % 4 != 0:
            raise PeerCommUnavailable('PeerComm: arena sizes must be multiples of 4 floats')
        self.max_floats = int(max_floats)
        nbytes = ctypes.c_size_t(0)
        _lib.check(L.ardae_dp_xchg_bytes(ctypes.c_size_t(self.max_floats), self.world, ctypes.byref(nbytes)))
        err = None
        with torch.cuda.device(device):
            self.xchg = torch.zeros(nbytes.value + 256, dtype=torch.uint8, device=device)
            self.ebs = torch.zeros(4, dtype=torch.int64, device=device)  # epoch, grid barrier, status, pad
            base = self.xchg.data_ptr() + (-self.xchg.data_ptr()) % 256
            torch.cuda.synchronize(device)
            handle = ctypes.create_string_buffer(64)
            off = ctypes.c_size_t(0)
            try:
                if os.environ.get('ARDAE_DP_FUSED_FAIL_RANK') == str(self.rank):  # test hook: one rank cannot export
                    raise RuntimeError('simulated export failure (ARDAE_DP_FUSED_FAIL_RANK)')
                _lib.check(L.ardae_ipc_export(ctypes.c_void_p(base), handle, ctypes.byref(off)))
                mine = (bytes(handle.raw), int(off.value))
            except RuntimeError as e:
                err, mine = e, None
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine, group=process_group)
            self.ptrs = (ctypes.c_void_p * self.world)()
            for r, item in enumerate(everyone):
                if r == self.rank:
                    self.ptrs[r] = base
                    continue
                try:
                    if item is None:
                        raise RuntimeError('rank %d could not export its exchange buffer' % r)
                    hb, o = item
                    if hb not in _IMPORTED:
                        out = ctypes.c_void_p(0)
                        _lib.check(L.ardae_ipc_import(hb, ctypes.c_size_t(0), ctypes.byref(out)))
                        _IMPORTED[hb] = out.value
                    self.ptrs[r] = _IMPORTED[hb] + o
                except RuntimeError as e:
                    err = err or e
            torch.cuda.synchronize(device)
            # collective verdict (doubles as the barrier: every rank has zeroed and mapped everything before step one)
            ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=process_group)
        if int(ok.item()) == 0:
            raise PeerCommUnavailable('peer memory could not be mapped on every rank (%s)' % (err,))

    def fused_step(self, kind, p, g, s1, s2, n, lr, b1, b2_or_alpha, eps, momentum, step, gscale):
        if n > self.max_floats:
            raise RuntimeError('PeerComm: arena larger than the exchange buffer')
        _lib.check(_lib.lib().ardae_dp_fused_step(
            int(kind), self.rank, self.world, _lib.ptr(p), _lib.ptr(g), _lib.ptr(s1), _lib.ptr(s2), ctypes.c_size_t(int(n)),
            self.ptrs, _lib.ptr(self.ebs), ctypes.c_float(lr), ctypes.c_float(b1), ctypes.c_float(b2_or_alpha),
            ctypes.c_float(eps), ctypes.c_float(momentum), int(step), ctypes.c_float(gscale), _lib.stream_ptr()))

    def check(self):
        """Host-side check (synchronises): raises if a peer flag never arrived in some fused step."""
        if int(self.ebs[2].item()) != 0:
            raise RuntimeError('PeerComm: a peer rank did not answer within the spin limit')

    def gather_state(self, opt):
        """Make `opt`'s flat state buffers complete on every rank (each rank holds the current values of its slice only)."""
        dist = torch.distributed
        ar = opt._setup()
        n4 = ar.total // 4
        q = (n4 + self.world - 1) // self.world
        lo, hi = min(self.rank * q, n4) * 4, min((self.rank + 1) * q, n4) * 4
        for buf in opt._bufs:
            own = torch.zeros_like(buf)
            own[lo:hi] = buf[lo:hi]
            dist.all_reduce(own, op=dist.ReduceOp.SUM, group=self.pg)
            buf.copy_(own)
        opt._state_sharded = False
