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

    def all_reduce(self, n: int, out: torch.Tensor, op: str = 'sum', scale: float = 1.0, offset_bytes: int = 0) -> torch.Tensor:
        """out[:n] = scale * sum | max over the ranks of the n floats at ``offset_bytes`` of every rank's data region."""
        if offset_bytes % 16 or offset_bytes + 4 * n > self.nbytes or out.numel() < n or out.dtype != F32:
            raise ValueError('sgcdet_b200.peer: bad all_reduce arguments')
        bufs = (ctypes.c_void_p * self.world)(*[b + self.sig_bytes + offset_bytes for b in self.bases])
        with torch.cuda.device(self.device):
            _lib.call('sgc_peer_allreduce', bufs, self._sigs, self.rank, self.world, int(n), 1 if op == 'max' else 0,
                      float(scale), _lib.ptr(out), _lib.stream(self.device))
        return out

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
    its backward node runs exactly when the group's gradients are final.  While a GradAverager is active that node averages
    them over the ranks -- on ONE communication stream, in autograd's (deterministic, rank-independent) execution order, each
    group in its own slice of the symmetric buffer -- and returns the averaged tensors to autograd, so ``p.grad`` never holds
    an unreduced value.  The large projection-gradient kernels of the finest level keep running meanwhile; only the gradients
    those kernels produce (the four projection layers of every level, ~1.2 MB) are averaged after the backward, in
    ``finish_step()``.

    Per step:  ``begin_step()`` before the forward, ``finish_step()`` after ``backward()``."""

    def __init__(self, params, group=None, device=None):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        self.n = (n + 3) // 4 * 4
        self.mem = PeerMemory(4 * self.n, group, device)
        self.device = self.mem.device
        self.flat_in = self.mem.view((self.n,))
        self.flat_in.zero_()
        self.scale = 1.0 / self.mem.world
        self.comm = torch.cuda.Stream(device=self.device, priority=-1)
        self._ids = {id(p) for p in self.params}
        self._done, self._active = set(), False
        self.groups_last_step = 0

    def _reduce(self, grads):
        """grads (final, on the current stream) -> averaged copies.  Every collective of this object runs on the communication
        stream (or after it has been joined), one after the other, and ends with a barrier behind the peers' last read: the
        symmetric buffer can be reused from offset 0 every time."""
        sizes = [g.numel() for g in grads]
        nb = sum(sizes)
        torch._foreach_copy_(list(self.flat_in[:nb].split(sizes)), [g.reshape(-1) for g in grads])
        out = torch.empty(nb, device=self.device, dtype=F32)
        self.mem.all_reduce(nb, out, 'sum', self.scale)
        return [o.view(g.shape) for o, g in zip(out.split(sizes), grads)]

    def _on_group(self, params, grads):
        """``functional.GRAD_REDUCER``: called from OnStream.backward on the group's weight-gradient stream."""
        if not self._active or any(g is None for g in grads) or any(id(p) not in self._ids for p in params):
            return grads
        cur = torch.cuda.current_stream(self.device)
        ev = torch.cuda.Event()
        ev.record(cur)
        self.comm.wait_event(ev)
        with torch.cuda.stream(self.comm):
            for g in grads:
                g.record_stream(self.comm)
            outs = self._reduce([g.contiguous() for g in grads])
        cur.wait_stream(self.comm)
        for o in outs:
            o.record_stream(cur)
        self._done.update(id(p) for p in params)
        self.groups_last_step += 1
        return tuple(outs)

    def begin_step(self):
        from . import functional as SF
        SF.GRAD_REDUCER = self._on_group
        self._done, self._active, self.groups_last_step = set(), True, 0

    def finish_step(self):
        """After ``backward()``: average, in place, every gradient no OnStream node has seen (on the current stream, which
        autograd has joined with all gradient streams), and make the current stream wait for the communication stream."""
        self._active = False
        main = torch.cuda.current_stream(self.device)
        main.wait_stream(self.comm)
        rest = [p for p in self.params if id(p) not in self._done and p.grad is not None]
        if rest:
            grads = [p.grad for p in rest]
            torch._foreach_copy_(grads, self._reduce(grads))

    def __call__(self):
        """Everything at once on the current stream, after the backward (no overlap): the round-1 placement."""
        self._done = set()
        self.finish_step()

    def close(self):
        from . import functional as SF
        if SF.GRAD_REDUCER == self._on_group:
            SF.GRAD_REDUCER = None
        self.mem.close()
