"""Symmetric peer memory over NVLink and the one-launch all-reduce on top of it (csrc/sgc_peer.cu).

SURVEY.md section 8e: the exchange steps of view sharding (partial sums / counts, score maxima, partial-softmax sums, the
backward's normaliser dot and query gradient) and the weight-gradient average of scene-batch data parallelism.  NCCL calls
could not be captured into the step's CUDA graph in this stack (and cost a CPU launch gap after every replay); here every
exchange is ONE kernel launch: the ranks meet at flags in each other's memory, every rank pulls the peers' partials through
16-byte peer loads and reduces them in rank order (bit-identical results on all ranks, which the replicated voxel chain and
its deterministic top-k rely on).

``torch.distributed`` is used once, at construction, to exchange the 64-byte CUDA IPC handles of the allocations.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib

F32 = torch.float32


class _Raw:
    """``__cuda_array_interface__`` view of raw device memory (the allocation is owned by ``PeerMemory``)."""

    def __init__(self, ptr: int, nbytes: int, owner):
        self.__cuda_array_interface__ = {'shape': (nbytes,), 'typestr': '|u1', 'data': (ptr, False), 'version': 2}
        self._owner = owner


class PeerMemory:
    """One symmetric allocation per rank of ``group`` (all ranks on one node): ``nbytes`` of data behind one signal pad.

    ``view(shape, offset)`` -> fp32 tensor aliasing this rank's data region (producers write their partial straight into
    it); ``all_reduce(n, out, op, scale, offset)`` reduces the first ``n`` floats at ``offset`` over the ranks into the
    ordinary tensor ``out`` on the current stream.  Collectives of one PeerMemory must be issued in the same order on
    every rank and on one stream at a time (they share the signal pad); use one PeerMemory per concurrent stream."""

    def __init__(self, nbytes: int, group: Optional[dist.ProcessGroup] = None, device: Optional[torch.device] = None):
        solo = not (dist.is_available() and dist.is_initialized())   # no process group: a single rank, nothing to exchange
        world = 1 if solo else dist.get_world_size(group)
        rank = 0 if solo else dist.get_rank(group)
        if world > 8:
            raise ValueError('sgcdet_b200.peer: at most 8 ranks (one NVSwitch domain)')
        device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        lib = _lib.load()
        sig_bytes = int(lib.sgc_peer_sig_bytes())
        nbytes = (int(nbytes) + 255) // 256 * 256
        base = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        with torch.cuda.device(device):
            _lib.call('sgc_peer_alloc', sig_bytes + nbytes, ctypes.byref(base), handle)
            handles = [None] * world
            if not solo:
                dist.all_gather_object(handles, bytes(handle.raw), group=group)
            bases = []
            for r in range(world):
                if r == rank:
                    bases.append(int(base.value))
                    continue
                p = ctypes.c_void_p()
                _lib.call('sgc_peer_open', ctypes.create_string_buffer(handles[r], 64), ctypes.byref(p))
                bases.append(int(p.value))
        self._setup(rank, world, bases, nbytes, device, group, owns=[rank], mapped=[r for r in range(world) if r != rank])
        if not solo:
            dist.barrier(group=group)   # every rank has mapped every peer before the first collective can spin on a flag

    def _setup(self, rank, world, bases, nbytes, device, group, owns, mapped):
        lib = _lib.load()
        self.rank, self.world, self.bases, self.nbytes, self.device, self.group = rank, world, list(bases), nbytes, device, group
        self.sig_bytes = int(lib.sgc_peer_sig_bytes())
        self._owns, self._mapped = owns, mapped
        self._sigs = (ctypes.c_void_p * world)(*self.bases)
        self._pad = torch.as_tensor(_Raw(self.bases[rank], self.sig_bytes, self), device=device)
        self._local = torch.as_tensor(_Raw(self.bases[rank] + self.sig_bytes, nbytes, self), device=device)
        self._status = int(lib.sgc_peer_status_offset())
        self._closed = False

    @classmethod
    def simulate(cls, world: int, nbytes: int, device=None):
        """``world`` ranks inside ONE process (no process group, plain allocations on one device): the single-GPU tests run
        the ranks' launches on ``world`` streams, where they meet at the flags exactly like ranks on different GPUs do."""
        device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        lib = _lib.load()
        nbytes = (int(nbytes) + 255) // 256 * 256
        bases = []
        with torch.cuda.device(device):
            for _ in range(world):
                base = ctypes.c_void_p()
                _lib.call('sgc_peer_alloc', int(lib.sgc_peer_sig_bytes()) + nbytes, ctypes.byref(base), ctypes.create_string_buffer(64))
                bases.append(int(base.value))
        out = []
        for r in range(world):
            m = cls.__new__(cls)
            m._setup(r, world, bases, nbytes, device, None, owns=[r], mapped=[])
            out.append(m)
        return out

    def check(self):
        """Raises if a barrier of an earlier collective gave up waiting for a peer (synchronises the device)."""
        if int(self._pad[self._status:self._status + 4].view(torch.int32).item()) != 0:
            raise RuntimeError('sgcdet_b200.peer: a peer did not arrive at a collective within ~4 s (results are invalid)')

    def view(self, shape, offset_bytes: int = 0) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= int(s)
        if offset_bytes % 16 or offset_bytes + 4 * n > self.nbytes:
            raise ValueError('sgcdet_b200.peer: view outside the symmetric allocation (or not 16-byte aligned)')
        return self._local[offset_bytes:offset_bytes + 4 * n].view(F32).view(*shape)

    def all_reduce(self, n: int, out: torch.Tensor, op: str = 'sum', scale: float = 1.0, offset_bytes: int = 0,
                   max_blocks: int = 0, block_threads: int = 0) -> torch.Tensor:
        """out[:n] = scale * sum | max over the ranks of the n floats at ``offset_bytes`` of every rank's data region.
        ``max_blocks`` (0 = up to 128) and ``block_threads`` (0 = 512) -- the same on every rank -- bound the footprint of a
        collective that overlaps other kernels."""
        if offset_bytes % 16 or offset_bytes + 4 * n > self.nbytes or out.numel() < n or out.dtype != F32:
            raise ValueError('sgcdet_b200.peer: bad all_reduce arguments')
        bufs = (ctypes.c_void_p * self.world)(*[b + self.sig_bytes + offset_bytes for b in self.bases])
        with torch.cuda.device(self.device):
            _lib.call('sgc_peer_allreduce', bufs, self._sigs, self.rank, self.world, int(n), 1 if op == 'max' else 0,
                      float(scale), _lib.ptr(out), int(max_blocks), int(block_threads), _lib.stream(self.device))
        return out

    def all_reduce_tensors(self, tensors, scale: float = 1.0) -> None:
        """tensors[k] <- scale * sum over the ranks, in place (contiguous fp32, at most 64 per launch, every rank the same
        sizes in the same order): gather, barrier, reduce, scatter in ONE launch (``sgc_peer_allreduce_tensors``)."""
        bufs = (ctypes.c_void_p * self.world)(*[b + self.sig_bytes for b in self.bases])
        with torch.cuda.device(self.device):
            for i in range(0, len(tensors), 64):
                ts = tensors[i:i + 64]
                if sum((t.numel() + 3) // 4 * 16 for t in ts) > self.nbytes or any(t.dtype != F32 or not t.is_contiguous() for t in ts):
                    raise ValueError('sgcdet_b200.peer: bad all_reduce_tensors arguments')
                a = (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
                n = (ctypes.c_longlong * len(ts))(*[t.numel() for t in ts])
                _lib.call('sgc_peer_allreduce_tensors', bufs, self._sigs, self.rank, self.world, a, n, len(ts), float(scale),
                          _lib.stream(self.device))

    def close(self):
        """Unmap the peers' allocations and free the own one (collective: every rank calls it)."""
        if self._closed:
            return
        self._closed = True
        torch.cuda.synchronize(self.device)
        lib = _lib.load()
        if self._mapped:
            dist.barrier(group=self.group)
        with torch.cuda.device(self.device):
            for r in self._mapped:
                lib.sgc_peer_close(ctypes.c_void_p(self.bases[r]))
            if self._mapped:
                dist.barrier(group=self.group)
            self._local = self._pad = None
            for r in self._owns:
                lib.sgc_peer_free(ctypes.c_void_p(self.bases[r]))


class GradAverager:
    """Scene-batch data parallelism: average the path's weight gradients over the ranks with peer all-reduce launches that live
    INSIDE the step's CUDA graph and overlap the end of the backward (replaces cat -> ncclAllReduce -> div -> copy issued by
    the CPU after every replay, which sat entirely behind the last kernel of the step).

    The path hands every parameter group to autograd through ``functional.OnStream`` on that group's weight-gradient stream;
    its backward node runs exactly when the group's gradients are final.  While a GradAverager is active that node records an
    event and returns PLACEHOLDER tensors to autograd (which adopts them as ``p.grad``: no kernel); when the last group of the
    step has reported, ONE all-reduce over all of them is issued on a communication stream -- it waits for the recorded
    events, gathers the gradients into the symmetric buffer, reduces, and scatters the averages into the placeholders --
    while the large projection-gradient kernels of the finest level are still running.  Only the gradients those kernels
    produce (the four projection layers of every level, ~1.2 MB) are averaged after the backward, in ``finish_step()``.
    Eight separate collectives (one per group) were measured first: 995 vs 1 137 volumes/s on two GPUs -- they queue in
    autograd's issue order, finest level first, and end up serialised behind the finest level's weight gradients.

    Per step:  ``begin_step()`` before the forward, ``finish_step()`` after ``backward()``.  The number of groups per step is
    learnt in the first step (which therefore reduces everything in ``finish_step``)."""

    def __init__(self, params, group=None, device=None, overlap_blocks: int = 64):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        self.n = (n + 3) // 4 * 4 + 4 * len(self.params)
        self.mem = PeerMemory(4 * self.n, group, device)
        self.device = self.mem.device
        self.flat_in = self.mem.view((self.n,))
        self.flat_in.zero_()
        self.scale = 1.0 / self.mem.world
        # high priority: the block scheduler hands freed SM slots to its (small) CTAs first -- at default priority the launches
        # waited until the step's large grids had dispatched all their CTAs, i.e. until the end of the step
        self.comm = torch.cuda.Stream(device=self.device, priority=-1)
        self.overlap_blocks = overlap_blocks
        self.small_group_bytes = 64 << 10
        self._ids = {id(p) for p in self.params}
        self._active, self._expected = False, None
        self._pending, self._done, self._adopted = [], set(), []
        self.groups_last_step, self.copied_last_step = 0, 0

    def _copy_segments(self, srcs, dsts, max_blocks):
        """dsts[k][:] = srcs[k][:] (contiguous fp32 tensors of equal sizes) as one small-footprint launch per 64 tensors."""
        k = len(srcs)
        a = (ctypes.c_void_p * k)(*[t.data_ptr() for t in srcs])
        b = (ctypes.c_void_p * k)(*[t.data_ptr() for t in dsts])
        n = (ctypes.c_longlong * k)(*[t.numel() for t in srcs])
        with torch.cuda.device(self.device):
            _lib.call('sgc_peer_copy_segments', a, b, n, k, int(max_blocks), _lib.stream(self.device))

    def _reduce_into(self, grads, outs, overlapped: bool = False):
        """outs[i] = average over the ranks of grads[i] (final on the current stream): gather into the symmetric buffer, one
        all-reduce, scatter -- three own launches.  Every collective of this object runs on the communication stream (or after
        it has been joined), one after the other, and ends with a barrier behind the peers' last read: the symmetric buffer is
        reused from offset 0 every time.  ``overlapped``: 128-thread CTAs, at most ``overlap_blocks`` of them, so that the
        launches find room beside the persistent tcgen05 kernels of the step's tail (whose CTAs leave < 20 K registers per SM)."""
        sizes = [(g.numel() + 3) // 4 * 4 for g in grads]      # every tensor starts 16-byte aligned in the flat buffer
        nb = sum(sizes)
        if nb > self.n:
            raise RuntimeError('sgcdet_b200.peer.GradAverager: more gradient elements than the buffer holds')
        offs = [0]
        for z in sizes[:-1]:
            offs.append(offs[-1] + z)
        gs = [g if g.is_contiguous() else g.contiguous() for g in grads]
        mb, bt = (self.overlap_blocks, 128) if overlapped else (0, 0)
        self._copy_segments(gs, [self.flat_in[o:o + g.numel()] for o, g in zip(offs, gs)], mb)
        red = torch.empty(nb, device=self.device, dtype=F32)
        self.mem.all_reduce(nb, red, 'sum', self.scale, max_blocks=mb, block_threads=bt)
        self._copy_segments([red[o:o + g.numel()] for o, g in zip(offs, gs)], outs, mb)

    def _flush(self, overlapped: bool = True):
        """All reported groups in one collective on the communication stream.  Issued from the backward it runs beside the
        step's largest kernels and has ~0.3 ms to finish: a few CTAs only (a spinning 512-thread CTA takes half an SM's
        registers away from ``lift_bwd`` or a persistent tcgen05 kernel, which measurably delayed the step with 128 CTAs)."""
        if not self._pending:
            return
        grads, outs = [], []
        for evs, gs, os_ in self._pending:
            for ev in evs:
                self.comm.wait_event(ev)
            grads += gs
            outs += os_
        with torch.cuda.stream(self.comm):
            for t in grads + outs:
                t.record_stream(self.comm)
            self._reduce_into(grads, outs, overlapped)
        self._pending = []

    def _on_group(self, params, grads):
        """``functional.GRAD_REDUCER``: called from OnStream.backward on the group's weight-gradient stream."""
        if not self._active or any(g is None for g in grads) or any(id(p) not in self._ids for p in params):
            return grads
        if sum(g.numel() for g in grads) * 4 < self.small_group_bytes:
            # e.g. the occupancy heads (C + 1 floats): their gradient kernels trail the step on their stream, and waiting for
            # them would hold the one large collective back until the very end -- they join the tail in finish_step instead
            return grads
        from . import functional as SF
        cur = torch.cuda.current_stream(self.device)
        gs = [g.contiguous() for g in grads]
        outs = [torch.empty_like(g) for g in gs]
        # readiness: the event the producing backward recorded right behind its launch (functional.mark_grads_ready); only
        # when a gradient carries none, an event on this node's stream (which may have unrelated late work queued)
        evs = [SF.GRAD_READY.pop(g.data_ptr(), None) for g in grads]
        if any(e is None for e in evs):
            ev = torch.cuda.Event()
            ev.record(cur)
            evs = [e for e in evs if e is not None] + [ev]
        ev = list({id(e): e for e in evs}.values())
        # autograd adopts a returned tensor as p.grad only while nobody else holds it: keep storage ALIASES (.data: another
        # tensor object on the same memory) for the collective to write through; finish_step verifies the adoption
        aliases = [o.data for o in outs]
        self._pending.append((ev, gs, aliases))
        self._adopted += list(zip(params, aliases))
        self._done.update(id(p) for p in params)
        self.groups_last_step += 1
        if self._expected is not None and self.groups_last_step == self._expected:
            self._flush()
        return tuple(outs)

    def begin_step(self):
        from . import functional as SF
        SF.GRAD_REDUCER = self._on_group
        self._pending, self._done, self._active, self.groups_last_step, self._adopted = [], set(), True, 0, []
        SF.GRAD_READY.clear()

    def finish_step(self):
        """After ``backward()``: flush groups that were not flushed from the backward (first step), average in place every
        gradient no OnStream node has seen (on the current stream, which autograd has joined with all gradient streams), and
        make the current stream wait for the communication stream."""
        self._active = False
        main = torch.cuda.current_stream(self.device)
        if self._pending:
            self.comm.wait_stream(main)
            self._flush(overlapped=False)
        if self._expected is None and self.groups_last_step:
            self._expected = self.groups_last_step
        main.wait_stream(self.comm)
        # a gradient autograd copied instead of adopting (another holder, or an accumulation into an existing .grad) did not
        # see the collective's result: hand it over explicitly
        stale = [(p.grad, a) for p, a in self._adopted if p.grad is not None and p.grad.data_ptr() != a.data_ptr()]
        if stale:
            torch._foreach_copy_([g for g, _ in stale], [a for _, a in stale])
        self.copied_last_step = len(stale)
        self._adopted = []
        rest = [p.grad for p in self.params if id(p) not in self._done and p.grad is not None]
        if rest and all(g.is_contiguous() for g in rest):
            self.mem.all_reduce_tensors(rest, self.scale)      # one launch: this is the part behind the step's last kernel
        elif rest:
            self._reduce_into(rest, rest)

    def __call__(self):
        """Everything at once on the current stream, after the backward (no overlap): the round-1 placement."""
        self._pending, self._done, self._adopted = [], set(), []
        self.finish_step()

    def close(self):
        from . import functional as SF
        if SF.GRAD_REDUCER == self._on_group:
            SF.GRAD_REDUCER = None
        self.mem.close()
