"""Fused AR-DAE training step: the body of train() (ivae_ardae.py:707-846, cdae_ctx_type='lt0') as
one stream-ordered sequence of libardae calls -- no autograd graph, no host sync, no `.item()`.

Differences from the reference's execution (not from its results):
  * `model.encode(x, std=0)` is evaluated once per minibatch instead of twice (:735/:748, :813/:826
    are pairwise identical);
  * the ELBO backward (:804) and the entropy-gradient injection (:834) are one backward pass with the
    summed upstream gradient on z;
  * noise is drawn on the device (Philox) unless injected through `noise=` (parity runs);
  * under data parallelism each rank holds a batch shard, losses/gradients are normalised by the
    GLOBAL row counts and the two flat gradient buffers are summed with one NCCL allreduce each;
  * `graph=True`: after two eager iterations the whole iteration (~150 launches on three streams) is captured
    into ONE CUDA graph and replayed; per-replay variation (Philox seeds, Adam's step) comes from a device-resident
    counter (ardae_set_replay_counter), beta from a device scalar (annealing does not re-capture), inputs are
    copied into static buffers.  Optimizer hyper-parameters (lr, betas, momentum, ...) are part of the capture
    signature: changing them re-captures.  In graph mode the returned tensors are STATIC buffers that the next
    replay overwrites -- clone them if they are read later than the next call.

Noise: draw k of iteration t uses the Philox seed  base + (64 t + k) * golden-ratio  (k < 64), on the host for eager
launches and through the device counter for replays, so no two draws of a run share a seed.
"""
import ctypes

import torch

from . import _lib


def dp_scales(B_local, nz, nstd, d, nz_model, world):
    """Normalisation constants of one rank under batch-sharded data parallelism: every rank
    scales its loss / gradient contributions by the GLOBAL counts, so that a plain SUM allreduce of the
    flat gradient buffers reproduces the reference's means over the whole batch
    (mse_loss mean over N*d, graddae/mlp.py:397; loss.mean() over B, ivae/mnist.py:249; 1/(B*nz_model) at
    ivae_ardae.py:834)."""
    n_local = B_local * nz * nstd
    return dict(cdae_inv_count=1.0 / float(n_local * world * d),
                model_inv_rows=1.0 / float(B_local * nz_model * world))


class TrainStep(object):
    def __init__(self, model, cdae, model_opt, cdae_opt, std_scale=10000., delta=0.1, nz_cdae=256, nstd=1,
                 nz_model=1, num_cdae_updates=1, process_group=None, seed=1234, graph=False, ctx_type='lt0',
                 data_ctx_affine=None):
        self.model, self.cdae, self.mopt, self.copt = model, cdae, model_opt, cdae_opt
        self.S, self.delta = float(std_scale), float(delta)
        self.nz, self.nstd, self.nzm, self.ncu = int(nz_cdae), int(nstd), int(nz_model), int(num_cdae_updates)
        # --cdae-ctx-type (ivae_ardae.py:729-741): 'lt0' = the mean code encode(x, std=0); 'data' = the input itself,
        # mapped to 2x-1 for the MNIST-like models ("if 'mnist' in opt.dataset"), as is for the toy model
        # 'hidden1a' = model.encode.forward_hidden(x, std=0), the two hidden codes of the hierarchical encoders (:739-741)
        if ctx_type not in ('lt0', 'data', 'hidden1a'):
            raise NotImplementedError("cdae_ctx_type must be 'lt0', 'data' or 'hidden1a'")
        if ctx_type == 'hidden1a' and getattr(model, 'KIND', None) != 'auxmnist':
            raise ValueError("cdae_ctx_type 'hidden1a' needs a hierarchical model (ivae_ardae.py:572-575)")
        self.ctx_type = ctx_type
        if data_ctx_affine is None:
            data_ctx_affine = (1.0, 0.0) if getattr(model, 'KIND', 'mnist') == 'toy' else (2.0, -1.0)
        self.data_ctx_affine = (float(data_ctx_affine[0]), float(data_ctx_affine[1]))
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        self.rank = torch.distributed.get_rank(process_group) if process_group is not None else 0
        self.seed = int(seed) * 1000003 + self.rank * 7919
        self.iter = 0      # iteration index t (advanced by __call__)
        self.draw = 0      # draw index k inside the iteration
        self.launches = 0
        # keep_noise: remember the device tensors of the noise this iteration drew (parity tests feed them to the
        # oracle); in graph mode they are static buffers overwritten by the next replay
        self.keep_noise = False
        self.last_noise = {}
        self._beta_dev = None   # device scalar holding beta (graph mode)
        if process_group is not None:
            self._sync_replicas()
        self.last_std = None
        self.profile = None  # set to a list to collect (name, start_event, end_event) per segment
        self.overlap = True
        self._side = None
        # CUDA-graph replay of the iteration (opt-in; eager whenever noise is injected or profiling is on)
        # Under data parallelism the two NCCL allreduces stay eager (capturing them hung on the 2-GPU box): the
        # iteration is captured as graph segments around them (three replays + two collectives per iteration).
        self.graph = bool(graph)
        self._cap = None        # segment recorder while capturing under DP
        self._g = None          # (CUDAGraph, static x_cdae list, static x_model, outputs, beta, shapes)
        self._g_eager_calls = 0
        self._g_ctr = None
        self._stg = None        # input staging (stage() / __call__() without inputs)
        import os
        self.dp_one_graph = os.environ.get('ARDAE_DP_ONE_GRAPH', '0') == '1'  # capture the collectives too (opt-in)
        # data parallel, opt-in (ARDAE_DP_FUSED=1): gradient exchange + optimizer as one peer-memory kernel per arena
        # (ardae/dp.py) instead of NCCL allreduce + optimizer launch; the iteration then holds no collective call and is
        # ONE graph per rank.  Measured +2-3 % at 8 GPUs; the NCCL path is the default because one of four 8-GPU runs
        # of the fused path did not finish within its time limit and could not be diagnosed
        self.dp_fused = self.world > 1 and os.environ.get('ARDAE_DP_FUSED', '0') == '1' and \
            torch.distributed.get_backend(process_group) == 'nccl'
        self._comm = None
        if self.dp_fused:
            self._init_dp_fused()

    class _Seg(object):
        def __init__(self, owner, name):
            self.o, self.name = owner, name

        def __enter__(self):
            if self.o.profile is not None:
                self.e0 = torch.cuda.Event(enable_timing=True)
                self.e1 = torch.cuda.Event(enable_timing=True)
                self.e0.record()

        def __exit__(self, *a):
            if self.o.profile is not None:
                self.e1.record()
                self.o.profile.append((self.name, self.e0, self.e1))

    def _seg(self, name):
        return TrainStep._Seg(self, name)

    def count_launches(self, B):
        """Kernel launches of one iteration (for bench.py's gpu_launches)."""
        L = _lib.lib()
        m, c = self.model, self.cdae
        n = 0
        n += L.ardae_model_num_launches(m._plans[m._plan(B, self.nz, 0)][0], 2)      # zbar (cdae minibatch): mean-code branch
        n += L.ardae_model_num_launches(m._plans[m._plan(B, 1, 0, 1)][0], 0)         # zbar (model minibatch)
        if self.ctx_type == 'hidden1a':                                              # + its std=0 branch (hidden context)
            n += L.ardae_model_num_launches(m._plans[m._plan(B, 1, 0, 1)][0], 2)
        n += L.ardae_model_num_launches(m._plans[m._plan(B, self.nz, 0)][0], 0)      # z samples
        n += L.ardae_cdae_num_launches(c._plan(B, self.nz * self.nstd, True))
        n += L.ardae_cdae_num_launches(c._plan(B, self.nzm, False))
        k = m._plan(B, self.nzm, 1)
        n += L.ardae_model_num_launches(m._plans[k][0], 0) + L.ardae_model_num_launches(m._plans[k][0], 1)
        n += 2 + 1 + 1 + 2 + 2  # randn x2, sigma schedule, scaled diff, optimizers, stage memsets
        return n

    SEED_STRIDE = 64  # == kReplaySeedStride (csrc/kernels.cuh)

    def _next_seed(self):
        self.draw += 1
        if self.draw >= self.SEED_STRIDE:  # sub-steps driven by hand (drop-in loops) without __call__
            self.iter += 1
            self.draw = 1
        k = self.iter * self.SEED_STRIDE + self.draw
        return ctypes.c_uint64((self.seed + k * 0x9E3779B97F4A7C15) & ((1 << 64) - 1))

    def _sync_replicas(self):
        """Data parallelism assumes bit-identical replicas: broadcast rank 0's parameters and optimizer state."""
        dist = torch.distributed
        src = dist.get_global_rank(self.pg, 0) if hasattr(dist, 'get_global_rank') else 0
        for mod, opt in ((self.model, self.mopt), (self.cdae, self.copt)):
            ar = mod._ensure()
            dist.broadcast(ar.flat, src=src, group=self.pg)
            opt._setup()
            for buf in opt._bufs:
                dist.broadcast(buf, src=src, group=self.pg)
            steps = torch.tensor([float(opt.state[p]['step']) for p in ar.params], device=ar.flat.device)
            dist.broadcast(steps, src=src, group=self.pg)
            for p, v in zip(ar.params, steps.tolist()):
                opt.state[p]['step'] = int(v)

    def _set_beta(self, beta, dev):
        if self._beta_dev is None:
            self._beta_dev = torch.empty(1, dtype=torch.float32, device=dev)
        self._beta_dev.fill_(float(beta))

    def _randn(self, rows, cols, dev):
        out = torch.empty(rows, cols, dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().ardae_randn(_lib.ptr(out), ctypes.c_size_t(out.numel()), self._next_seed(), 3,
                                          _lib.stream_ptr()))
        return out

    def _context(self, xs, zbar, hidden=None):
        """CDAE context rows [B, c] for a minibatch (xs: [B, D] view of the inputs)."""
        if self.ctx_type == 'hidden1a':
            return hidden
        if self.ctx_type == 'data':
            a, s = self.data_ctx_affine
            return (xs * a + s) if (a != 1.0 or s != 0.0) else xs.contiguous()
        return zbar

    def _peer_comm(self):
        if self._comm is None:
            from .dp import PeerComm
            mar, car = self.model._ensure(), self.cdae._ensure()
            self._comm = PeerComm.get(self.pg, mar.flat.device, max(mar.total, car.total))
        return self._comm

    def _init_dp_fused(self):
        """Decide, collectively, whether the peer-memory exchange is available (else: NCCL allreduce)."""
        from .dp import PeerCommUnavailable
        try:
            self._peer_comm()
        except PeerCommUnavailable as e:
            import warnings
            warnings.warn('ardae.TrainStep: %s; using NCCL allreduce' % (e,))
            self.dp_fused = False

    def gather_optimizer_state(self):
        """With the fused data-parallel update every rank advances only its slice of the optimizer state: call this
        before saving / inspecting optimizer state (no-op otherwise)."""
        if self._comm is not None and self.dp_fused:
            self._comm.gather_state(self.copt)
            self._comm.gather_state(self.mopt)

    def _allreduce(self, flat):
        if self.world <= 1:
            return
        cap = self._cap
        if cap is not None:
            # segment boundary: join the forward side stream, close the graph captured so far, note the collective,
            # open the next segment (nothing executes during capture, the collective included)
            main = torch.cuda.current_stream()
            if cap['side_forked'] and not cap['side_joined']:
                main.wait_stream(self._side)
                cap['side_joined'] = True
            cap['graph'].capture_end()
            cap['ops'].append(cap['graph'])
            cap['ops'].append(flat)
            cap['graph'] = torch.cuda.CUDAGraph()
            cap['graph'].capture_begin(pool=cap['pool'])
            return
        torch.distributed.all_reduce(flat, op=torch.distributed.ReduceOp.SUM, group=self.pg)

    def cdae_update(self, x, noise=None):
        L = _lib.lib()
        m, c = self.model, self.cdae
        B = x.size(0)
        d, n = m.z_dim, m._noise_width
        dev = x.device
        xs = _lib.require_cuda(x, 'x').view(B, -1)
        hid = None
        with self._seg('encode'):
            enc = noise['enc_cdae'] if noise is not None else self._randn(B * self.nz, n, dev)
            if self.ctx_type == 'hidden1a':
                z, zbar, hid = m._encode_hidden(xs, enc, self.nz)                        # :740 next to :748,:749
            else:
                z, zbar = m._encode_with_mean(xs, enc, self.nz)                          # :735,:748 and :749 in one pass
        N = B * self.nz * self.nstd
        xc = torch.empty(N, d, dtype=torch.float32, device=dev)
        sigma = torch.empty(N, dtype=torch.float32, device=dev)
        std = torch.empty(B, dtype=torch.float32, device=dev)
        xi = _lib.require_cuda(noise['xi'], 'xi').reshape(-1) if noise is not None else None
        _lib.check(L.ardae_sigma_schedule(_lib.ptr(z), _lib.ptr(zbar), B, self.nz, d, self.nstd,
                                          ctypes.c_float(self.S), ctypes.c_float(self.delta), _lib.ptr(xi),
                                          self._next_seed(), _lib.ptr(xc), _lib.ptr(sigma), _lib.ptr(std),
                                          _lib.stream_ptr()))                           # :753-767
        ar = c._ensure()
        h = c._plan(B, self.nz * self.nstd, True)
        ar.stage_flat.zero_()
        c._stage_gen += 1  # a pending drop-in loss.backward() of this module is stale from here on
        if noise is not None:
            eps = _lib.require_cuda(noise['eps_cdae'], 'eps').reshape(N, d).clone()
            gen = 0
        else:
            eps = torch.empty(N, d, dtype=torch.float32, device=dev)
            gen = 1
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        inv = dp_scales(B, self.nz, self.nstd, d, self.nzm, self.world)['cdae_inv_count']
        with self._seg('cdae_train'):
            _lib.check(L.ardae_cdae_train(h, _lib.ptr(xc), _lib.ptr(self._context(xs, zbar, hid)), _lib.ptr(sigma), _lib.ptr(eps), gen,
                                          self._next_seed(), ctypes.c_float(inv), _lib.ptr(loss), None,
                                          _lib.stream_ptr()))                           # :768-771
        if self.dp_fused:
            with self._seg('cdae_opt'):
                self.copt.step_flat_dp(ar.stage_flat, self._peer_comm(), skip=c.no_grad_params)   # :779, all ranks
        else:
            with self._seg('cdae_allreduce'):
                self._allreduce(ar.stage_flat)
            with self._seg('cdae_opt'):
                self.copt.step_flat(ar.stage_flat, skip=c.no_grad_params)           # :779
        self.last_std = std
        if self.keep_noise:
            self.last_noise.update(enc_cdae=enc, sigma=sigma, std=std, eps_cdae=eps, z_cdae=z, zbar_cdae=zbar)
        return loss

    def model_forward(self, x, beta, noise=None):
        """Everything of the model update that does not depend on the CDAE update of this iteration:
        ELBO forward (:801), zbar = encode(x, std=0) (:813,:826) and S*(z - zbar) (:827)."""
        L = _lib.lib()
        m = self.model
        B = x.size(0)
        d, n = m.z_dim, m._noise_width
        dev = x.device
        xs = _lib.require_cuda(x, 'x').view(B, -1)
        R = B * self.nzm
        enc = noise['enc_model'] if noise is not None else self._randn(R, n, dev)
        m._ensure()
        key = m._plan(B, self.nzm, 1)
        hm = m._plans[key][0]
        z = torch.empty(R, d, dtype=torch.float32, device=dev)
        sums = torch.empty(3, dtype=torch.float32, device=dev)
        inv_rows = dp_scales(B, self.nz, self.nstd, d, self.nzm, self.world)['model_inv_rows']
        if self.keep_noise:
            self.last_noise.update(enc_model=enc)
        if self.graph and self._beta_dev is None:  # sub-step driven by hand before the first __call__
            self._set_beta(beta, dev)
        m._plan_gen[key] = m._plan_gen.get(key, 0) + 1
        _lib.check(L.ardae_model_set_beta_device(hm, _lib.ptr(self._beta_dev) if self.graph else None))
        with self._seg('model_fwd'):
            _lib.check(L.ardae_model_forward(hm, _lib.ptr(xs), _lib.ptr(enc), ctypes.c_float(beta),
                                             ctypes.c_float(inv_rows), _lib.ptr(z), _lib.ptr(sums), None,
                                             _lib.stream_ptr()))                        # :801
            hid = None
            if self.ctx_type == 'hidden1a':
                _, zbar, hid = m._encode_hidden(xs, None, 1, slot=1)                     # :816 next to :826
            else:
                zbar = m._encode(xs, None, 1, slot=1)                                    # :813,:826
            if self.keep_noise:
                self.last_noise.update(zbar_model=zbar)
            xsd = torch.empty(R, d, dtype=torch.float32, device=dev)
            _lib.check(L.ardae_scaled_diff(_lib.ptr(z), _lib.ptr(zbar), R, self.nzm, d, ctypes.c_float(self.S),
                                           _lib.ptr(xsd), _lib.stream_ptr()))           # :827
            # decoder + prior half of the backward (:804): needs only the ELBO forward, so it runs here, underneath
            # the CDAE update; the encoder half waits for the entropy gradient in model_backward
            ar = m._ensure()
            ar.stage_flat.zero_()
            _lib.check(L.ardae_model_backward_decoder(hm, ctypes.c_float(1.0), _lib.stream_ptr()))
        return dict(hm=hm, z=z, sums=sums, zbar=zbar, xsd=xsd, inv_rows=inv_rows, B=B, R=R, enc=enc, xs=xs,
                    ctx=self._context(xs, zbar, hid))

    def model_backward(self, f, beta):
        """Entropy-gradient estimate with the UPDATED cdae (:829), one backward for :804 + :834, Adam (:846)."""
        L = _lib.lib()
        m, c = self.model, self.cdae
        d = m.z_dim
        dev = f['z'].device
        ar = m._ensure()
        c._ensure()
        hs = c._plan(f['B'], self.nzm, False)
        zero_sigma = torch.zeros(f['R'], dtype=torch.float32, device=dev)
        g = torch.empty(f['R'], d, dtype=torch.float32, device=dev)
        with self._seg('score'):
            _lib.check(L.ardae_cdae_score(hs, _lib.ptr(f['xsd']), _lib.ptr(f['ctx']), _lib.ptr(zero_sigma),
                                          _lib.ptr(g), _lib.stream_ptr()))              # :829
        # :834 (with beta in a device scalar the kernels multiply it in)
        gz_scale = self.S * f['inv_rows'] * (1.0 if self.graph else beta)
        with self._seg('model_bwd'):
            _lib.check(L.ardae_model_backward_encoder(f['hm'], ctypes.c_float(1.0), _lib.ptr(g),
                                                      ctypes.c_float(gz_scale), _lib.stream_ptr()))  # :804 + :834 (encoder half)
        if self.dp_fused:
            with self._seg('model_opt'):
                self.mopt.step_flat_dp(ar.stage_flat, self._peer_comm())                 # :846, all ranks
        else:
            with self._seg('model_allreduce'):
                self._allreduce(ar.stage_flat)
            with self._seg('model_opt'):
                self.mopt.step_flat(ar.stage_flat)                                       # :846
        return f['sums'], g, f['z']

    def model_update(self, x, beta, noise=None):
        return self.model_backward(self.model_forward(x, beta, noise), beta)

    def stage(self, x_cdae_host, x_model_host):
        """Start copying the NEXT iteration's minibatches from pinned host memory on a copy stream (double-buffered
        device staging).  The following `__call__()` without inputs consumes them, so the transfer runs underneath
        the current iteration instead of in front of the next one -- what the reference gets from its DataLoader
        workers + `.to(device)` (ivae_ardae.py:707-712).  x_cdae_host: one tensor (shared by all CDAE updates)."""
        dev = next(self.model.parameters()).device
        if self._stg is None:
            self._stg = dict(stream=torch.cuda.Stream(device=dev), k=0, bufs=[None, None], ready=[None, None],
                             consumed=[None, None], pending=None)
        st = self._stg
        k = st['k']
        st['k'] = 1 - k
        shapes = (tuple(x_cdae_host.shape), tuple(x_model_host.shape))
        if st['bufs'][k] is None or st['bufs'][k][2] != shapes:
            st['bufs'][k] = (torch.empty(shapes[0], dtype=torch.float32, device=dev),
                             torch.empty(shapes[1], dtype=torch.float32, device=dev), shapes)
        if st['consumed'][k] is not None:      # the iteration that read this pair last must have copied it out
            st['stream'].wait_event(st['consumed'][k])
        with torch.cuda.stream(st['stream']):
            st['bufs'][k][0].copy_(x_cdae_host, non_blocking=True)
            st['bufs'][k][1].copy_(x_model_host, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(st['stream'])
        st['ready'][k] = ev
        st['pending'] = k

    def _take_staged(self):
        st = self._stg
        if st is None or st['pending'] is None:
            raise RuntimeError('TrainStep(): no inputs given and nothing staged (call stage(x_cdae_host, x_model_host) first)')
        k = st['pending']
        st['pending'] = None
        torch.cuda.current_stream().wait_event(st['ready'][k])
        return k, st['bufs'][k][0], st['bufs'][k][1]

    def __call__(self, x_cdae=None, x_model=None, beta=1.0, noise=None):
        """One iteration.  x_cdae: the minibatch (or list of num_cdae_updates minibatches) for the CDAE
        update(s); x_model: the minibatch of the model update; both None: the pair handed to `stage()`.
        Returns device tensors (no sync)."""
        staged = None
        if x_cdae is None and x_model is None:
            staged, x_cdae, x_model = self._take_staged()
        out = self._iterate(x_cdae, x_model, beta, noise)
        if staged is not None:
            # graph mode copied the pair into the static buffers before the replay; eager mode read it in place
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self._stg['consumed'][staged] = ev
        return out

    def _iterate(self, x_cdae, x_model, beta, noise):
        self.iter += 1
        self.draw = 0
        if self.graph:
            x0 = x_cdae[0] if isinstance(x_cdae, (list, tuple)) else x_cdae
            self._set_beta(beta, x0.device)
        if self.graph and noise is None and self.profile is None:
            return self._call_graph(x_cdae, x_model, beta)
        out = self._call_eager(x_cdae, x_model, beta, noise)
        if self._g is not None:  # keep the captured graph's step counter in line with eager iterations in between
            _lib.check(_lib.lib().ardae_bump_replay_counter(ctypes.c_void_p(self._g_ctr.data_ptr()), _lib.stream_ptr()))
        return out

    # ------------------------------------------------------------------ CUDA-graph replay
    def _bump_host_steps(self):
        """What the optimizers' host-side bookkeeping would have done in an eager iteration."""
        for opt, skip in ((self.copt, self.cdae.no_grad_params), (self.mopt, ())):
            ar = opt._setup()
            for k, p in enumerate(ar.params):
                if k not in skip:
                    opt.state[p]['step'] += 1

    class _Segments(object):
        """A captured iteration under DP: CUDA graphs alternating with the flat buffers to allreduce eagerly."""

        def __init__(self, owner, ops, out):
            self.owner, self.ops, self.out = owner, ops, out

        def replay(self):
            for op in self.ops:
                if torch.is_tensor(op):
                    torch.distributed.all_reduce(op, op=torch.distributed.ReduceOp.SUM, group=self.owner.pg)
                else:
                    op.replay()

    def _capture_segments(self, gxs, gxm, beta):
        L = _lib.lib()
        cs = torch.cuda.Stream()
        cs.wait_stream(torch.cuda.current_stream())
        pool = torch.cuda.graph_pool_handle()
        cap = dict(graph=torch.cuda.CUDAGraph(), ops=[], pool=pool, side_forked=False, side_joined=False)
        with torch.cuda.stream(cs):
            cap['graph'].capture_begin(pool=pool)
            self._cap = cap
            try:
                out = self._call_eager(gxs, gxm, beta, None)
                _lib.check(L.ardae_bump_replay_counter(ctypes.c_void_p(self._g_ctr.data_ptr()), _lib.stream_ptr()))
            finally:
                self._cap = None
                cap['graph'].capture_end()
            cap['ops'].append(cap['graph'])
        torch.cuda.current_stream().wait_stream(cs)
        return TrainStep._Segments(self, cap['ops'], out)

    def _call_graph(self, x_cdae, x_model, beta):
        xs = list(x_cdae) if isinstance(x_cdae, (list, tuple)) else [x_cdae] * self.ncu
        hyper = tuple(tuple(sorted((k, v) for k, v in opt.param_groups[0].items() if k != 'params'))
                      for opt in (self.copt, self.mopt))
        sig = (tuple(tuple(x.shape) for x in xs), tuple(x_model.shape), xs[0].device, hyper, self.S, self.delta,
               self.keep_noise)
        if self._g is not None and self._g[4] != sig:
            self._g = None  # shapes / optimizer hyper-parameters changed: capture again (beta lives in a device scalar)
            self._g_eager_calls = 0
        if self._g is None:
            if self._g_eager_calls < 2:  # plans, side streams and allocator pools come to life eagerly
                self._g_eager_calls += 1
                return self._call_eager(x_cdae, x_model, beta, None)
            L = _lib.lib()
            dev = xs[0].device
            if self._g_ctr is None:
                self._g_ctr = torch.zeros(1, dtype=torch.int64, device=dev)
            self._g_ctr.zero_()
            same = all(x is xs[0] for x in xs)
            gx0 = torch.empty_like(xs[0])
            gxs = [gx0] * len(xs) if same else [torch.empty_like(x) for x in xs]
            gxm = torch.empty_like(x_model)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            _lib.check(L.ardae_set_replay_counter(ctypes.c_void_p(self._g_ctr.data_ptr())))
            steps_before = [[opt.state[p]['step'] for p in opt._setup().params] for opt in (self.copt, self.mopt)]
            try:
                if self.world > 1 and not (self.dp_one_graph or self.dp_fused):
                    g = self._capture_segments(gxs, gxm, beta)
                    out = g.out
                else:
                  # world > 1 here: the two NCCL allreduces are captured with the rest (thread-local capture mode: the
                  # process group's watchdog thread keeps making CUDA calls)
                  if self.dp_fused:
                      self._peer_comm()  # IPC setup is a collective: not inside the capture
                  with torch.cuda.graph(g, capture_error_mode='thread_local' if self.world > 1 else 'global'):
                    out = self._call_eager(gxs, gxm, beta, None)
                    _lib.check(L.ardae_bump_replay_counter(ctypes.c_void_p(self._g_ctr.data_ptr()), _lib.stream_ptr()))
            except Exception as e:  # e.g. a collective that cannot be captured: stay eager, loudly
                import warnings
                warnings.warn('ardae.TrainStep: CUDA-graph capture failed (%s); continuing eagerly' % (e,))
                _lib.check(L.ardae_set_replay_counter(None))
                for opt, before in zip((self.copt, self.mopt), steps_before):
                    for p, v in zip(opt._setup().params, before):
                        opt.state[p]['step'] = v
                self.graph = False
                torch.cuda.synchronize()
                return self._call_eager(x_cdae, x_model, beta, None)
            finally:
                _lib.check(L.ardae_set_replay_counter(None))
            # the captured call advanced the host-side step counters once; undo, replay() re-applies per replay
            for opt, skip in ((self.copt, self.cdae.no_grad_params), (self.mopt, ())):
                ar = opt._setup()
                for k, p in enumerate(ar.params):
                    if k not in skip:
                        opt.state[p]['step'] -= 1
            # the captured kernels carry step0 = (host step + 1); the counter starts at 0 for the first replay
            self._g = (g, gxs, gxm, out, sig, same)
        g, gxs, gxm, out, _, same = self._g
        if same:
            if xs[0].data_ptr() != gxs[0].data_ptr():
                gxs[0].copy_(xs[0], non_blocking=True)
        else:
            for dst, src in zip(gxs, xs):
                if src.data_ptr() != dst.data_ptr():
                    dst.copy_(src, non_blocking=True)
        if x_model.data_ptr() != gxm.data_ptr():
            gxm.copy_(x_model, non_blocking=True)
        g.replay()
        self._bump_host_steps()
        self.launches = 1
        return out

    def _call_eager(self, x_cdae, x_model, beta=1.0, noise=None):
        xs = x_cdae if isinstance(x_cdae, (list, tuple)) else [x_cdae] * self.ncu
        closs = None
        main = torch.cuda.current_stream()
        if self.overlap:
            # the ELBO forward of the model update does not depend on this iteration's CDAE update:
            # run its (latency-bound, B-row) kernels on a side stream underneath the CDAE sweeps
            if self._side is None:
                self._side = torch.cuda.Stream()
            self._side.wait_stream(main)
            if self._cap is not None:
                self._cap['side_forked'] = True
            with torch.cuda.stream(self._side):
                fwd = self.model_forward(x_model, beta, noise)
                for k in ('z', 'sums', 'zbar', 'xsd', 'enc', 'ctx'):
                    if torch.is_tensor(fwd[k]):
                        fwd[k].record_stream(main)
        for i in range(self.ncu):
            closs = self.cdae_update(xs[i], noise)
        if self.overlap:
            if self._cap is None or not self._cap['side_joined']:
                main.wait_stream(self._side)
        else:
            fwd = self.model_forward(x_model, beta, noise)
        sums, g, z = self.model_backward(fwd, beta)
        # `losses` = (cdae_loss, model_loss, recon, prior) in one tensor: one device->host read per logged iteration
        return dict(cdae_loss=closs, model_loss=sums[0:1], recon=sums[1:2], prior=sums[2:3], std=self.last_std,
                    entropy_grad=g, z_model=z, losses=torch.cat([closs.reshape(1), sums[0:3]]))
