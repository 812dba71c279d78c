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

    The parameters are bucketed in the order their gradients become final (recorded during the first, eager step through
    ``post_accumulate_grad`` hooks, like DDP's reverse-order buckets): every bucket is averaged on a communication stream as
    soon as its last gradient is accumulated, while the large projection-gradient kernels of the finest level are still
    running; only the small tail bucket (the gradients those kernels produce) remains behind the backward.

    Per step:  ``begin_step()`` before the forward, ``finish_step()`` after ``backward()`` (joins the communication stream;
    in the first step it averages everything at once)."""

    def __init__(self, params, group=None, device=None, bucket_bytes: int = 4 << 20, tail_bytes: int = 1 << 20):
        self.params = [p for p in params if p.requires_grad]
        self.sizes = [p.numel() for p in self.params]
        n = sum(self.sizes)
        self.n = (n + 3) // 4 * 4 + 4 * len(self.params)        # every bucket starts 16-byte aligned
        self.mem = PeerMemory(4 * self.n, group, device)
        self.device = self.mem.device
        self.flat_in = self.mem.view((self.n,))
        self.flat_in.zero_()
        self.flat_out = torch.empty(self.n, device=self.device, dtype=F32)
        self.scale = 1.0 / self.mem.world
        self.bucket_bytes, self.tail_bytes = bucket_bytes, tail_bytes
        self.comm = torch.cuda.Stream(device=self.device, priority=-1)
        self.order = []            # parameter indices in the order their gradients became final (first step)
        self.buckets = None        # list of (param indices, offset, n) once the order is known
        self._index = {id(p): i for i, p in enumerate(self.params)}
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self._pending, self._events, self._main = None, None, None

    # ---- bucket plan ------------------------------------------------------------------------------------------------------
    def _plan(self):
        order = self.order + [i for i in range(len(self.params)) if i not in set(self.order)]
        # tail bucket: the gradients that arrive last, up to tail_bytes; the rest in chunks of bucket_bytes
        tail, acc = [], 0
        for i in reversed(order):
            if acc + 4 * self.sizes[i] > self.tail_bytes and tail:
                break
            tail.insert(0, i)
            acc += 4 * self.sizes[i]
        head = order[:len(order) - len(tail)]
        groups, cur, acc = [], [], 0
        for i in head:
            cur.append(i)
            acc += 4 * self.sizes[i]
            if acc >= self.bucket_bytes:
                groups.append(cur)
                cur, acc = [], 0
        if cur:
            groups.append(cur)
        groups.append(tail)
        self.buckets, off = [], 0
        for gidx in groups:
            nb = sum(self.sizes[i] for i in gidx)
            self.buckets.append((gidx, off, nb))
            off += (nb + 3) // 4 * 4
        self._bucket_of = {i: b for b, (gidx, _, _) in enumerate(self.buckets) for i in gidx}

    def _reduce(self, idxs, off, nb):
        grads = [self.params[i].grad.view(-1) for i in idxs]
        sizes = [self.sizes[i] for i in idxs]
        torch._foreach_copy_(list(self.flat_in[off:off + nb].split(sizes)), grads)
        self.mem.all_reduce(nb, self.flat_out[off:off + nb], 'sum', self.scale, offset_bytes=4 * off)
        torch._foreach_copy_(grads, list(self.flat_out[off:off + nb].split(sizes)))

    # ---- per step -----------------------------------------------------------------------------------------------------------
    def begin_step(self):
        self._main = torch.cuda.current_stream(self.device)
        if self.buckets is None:
            self._pending = None
            if self.order:           # second step: the first one recorded the order
                self._plan()
        if self.buckets is not None:
            self._pending = [len(g) for g, _, _ in self.buckets]
            self._events = [[] for _ in self.buckets]

    def _on_grad(self, p):
        i = self._index[id(p)]
        if self.buckets is None:
            if self._pending is None and i not in self.order:
                self.order.append(i)
            return
        if self._pending is None:
            return
        b = self._bucket_of[i]
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))     # the stream this gradient was accumulated on
        self._events[b].append(ev)
        self._pending[b] -= 1
        if self._pending[b] == 0:
            for e in self._events[b]:
                self.comm.wait_event(e)
            idxs, off, nb = self.buckets[b]
            with torch.cuda.stream(self.comm):
                for j in idxs:
                    self.params[j].grad.record_stream(self.comm)
                self._reduce(idxs, off, nb)

    def finish_step(self):
        main = torch.cuda.current_stream(self.device)
        if self.buckets is None or self._pending is None:
            self._reduce(list(range(len(self.params))), 0, sum(self.sizes))      # first step: everything at once
            return
        for b, left in enumerate(self._pending):
            if left:                 # a bucket with a parameter that received no gradient this step: reduce it here
                idxs, off, nb = self.buckets[b]
                idxs = [j for j in idxs if self.params[j].grad is not None]
                if len(idxs) != len(self.buckets[b][0]):
                    raise RuntimeError('sgcdet_b200.peer.GradAverager: a parameter received no gradient in this step')
                main.wait_stream(self.comm)
                self._reduce(idxs, off, nb)
        main.wait_stream(self.comm)
        self._pending = None

    def __call__(self):
        """Everything at once on the current stream (no overlap): the round-1 placement, kept for comparison."""
        self._reduce(list(range(len(self.params))), 0, sum(self.sizes))

    def close(self):
        for h in self._hooks:
            h.remove()
        self.mem.close()
