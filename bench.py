#!/usr/bin/env python
"""bench.py -- view-transform throughput (voxel volumes/s, forward+backward) of sgcdet_b200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|reference-gpu]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One *step* = one synthetic scene (one voxel volume) through all three levels of AdaptiveSparseHead, forward
plus backward from a fixed upstream gradient and the occupancy loss.  Workload at N=1:
``SGCDet_ScanNet`` at the train shape V=40 (BASELINE.json configs[1]).  For N>1 every rank owns its own
scene (scene-batch data parallel, weak scaling) and the path's weight gradients are all-reduced with NCCL
inside the step.  Rank 0 prints ONE JSON line.

Arms:
  ours          the CUDA path (CUDA-graph replay of fwd+bwd), device-timed with CUDA events, max over ranks
  reference     the CPU restatement of the reference (oracle/, kind "port": the reference has no CPU
                implementation and its plugin cannot be imported without mmcv) on all host cores, bounded sample
  reference-gpu (extra, not part of the driver contract) the restated reference glue driving the reference's
                own DFA3D kernels compiled unmodified for sm_100a (oracle/_ref) on the same B200
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = 'voxel volumes/sec (fwd+bwd view transform)'
UNIT = 'volumes/s'


# -------------------------------------------------------------------------------------------------
# clocks
# -------------------------------------------------------------------------------------------------

class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-i', str(self.idx), '-lms', '100'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            f = [x.strip() for x in r.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    power_w_max=max(pw) if pw else None, samples=len(sm), reasons=sorted(reasons))


# -------------------------------------------------------------------------------------------------
# kernel accounting
# -------------------------------------------------------------------------------------------------

# kernels launched per C-ABI entry point (csrc/*.cu)
LAUNCHES = dict(sgc_project_compact=2, sgc_lift_fwd=1, sgc_lift_bwd=1, sgc_crossview_mean_fwd=1,
                sgc_crossview_attn_fwd=1, sgc_crossview_attn_bwd_qt=1, sgc_crossview_attn_bwd_slots=1,
                sgc_upsample2x_occ_fwd=1, sgc_upsample2x_occ_bwd=4, sgc_dropout_masks=1, sgc_topk_select=1, sgc_scatter_add_rows=1,
                sgc_gather_rows=1, sgc_split_bf16x3=1, sgc_colsum=1, sgc_pack_weight_tc=1, sgc_project_tc_bwd_data=1, sgc_project_tc_wgrad=2, sgc_split_rows_colsum=1, sgc_project_tc_fwd=1, sgc_rows_gemm_tc=1,
                sgc_prepare_weights=1, sgc_rows_wgrad_tc=2, sgc_topk_select_mc=9, sgc_rows_wgrad_group_tc=2,
                sgc_lift_bwd_tiles=7, sgc_topk_select_grid=1, sgc_fold_wcat=1, sgc_unfold_wcat_grad=1,
                sgc_occ_loss_fwd=1, sgc_occ_loss_bwd=1)


class CallRecorder:
    """Wraps sgcdet_b200._lib.call: counts launches; optionally brackets every call with CUDA events on the
    launching stream (instrumented pass only)."""

    def __init__(self):
        from sgcdet_b200 import _lib, functional
        self._lib, self._fn = _lib, functional
        self.orig = _lib.call
        self.count = 0
        self.timed = False
        self.events = []

    def __enter__(self):
        def call(name, *args):
            self.count += LAUNCHES.get(name, 1)
            if self.timed:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                self.orig(name, *args)
                b.record()
                self.events.append((name, args, a, b))
            else:
                self.orig(name, *args)
        self._lib.call = call
        self._fn.call = call
        return self

    def __exit__(self, *exc):
        self._lib.call = self.orig
        self._fn.call = self.orig


def kernel_algorithmic_bytes(name: str, args, n_pairs_by_q: dict) -> float:
    """Unique bytes a launch has to move (DESIGN.md "Kernels"): inputs read once + outputs written once."""
    f = 4.0
    if name in ('sgc_lift_fwd', 'sgc_lift_bwd', 'sgc_lift_bwd_tiles'):
        if name == 'sgc_lift_fwd':
            ldv, S, H, W, D, Q, C = args[1], args[11], args[12], args[13], args[14], args[15], args[16]
        elif name == 'sgc_lift_bwd_tiles':
            ldv, S, H, W, D, Q, C = args[1], args[13], args[14], args[15], args[16], args[17], args[18]
        else:
            ldv, S, H, W, D, Q, C = args[1], args[12], args[13], args[14], args[15], args[16], args[17]
        V = n_pairs_by_q[Q][1]
        P = n_pairs_by_q[Q][0]
        maps = f * V * S * (ldv + D)          # value + G + depth maps
        pairs = f * P * (C + 128) + P * 16    # slots + samp rows, pair id + ref point
        return maps + pairs if name == 'sgc_lift_fwd' else 2 * maps + pairs
    if name in ('sgc_rowop_fwd', 'sgc_rowop_bwd'):
        a = args[0]._obj   # ctypes.byref(struct)
        rn = float(a.R) * a.N
        if name == 'sgc_rowop_fwd':
            n32 = sum(1 for p in (a.x, a.residual, a.y, a.pre) if p)
            return rn * (4 * n32 + (1 if a.mask else 0) + (6 if a.ysplit else 0))
        n32 = sum(1 for p in (a.g, a.g2, a.pre, a.gate, a.gpre, a.gx) if p)
        return rn * (4 * n32 + (1 if a.mask else 0) + (6 if a.gxsplit else 0))
    if name == 'sgc_layernorm_bwd':
        return 3 * f * args[5] * args[6]
    if name.startswith('sgc_crossview'):
        split = name.endswith('_split')
        if split:
            name = name[:-len('_split')]
            extra = 6.0 * args[4 if name == 'sgc_crossview_mean_fwd' else 5] * \
                (args[3] if name == 'sgc_crossview_mean_fwd' else 8 * args[4])
            return extra + kernel_algorithmic_bytes(name, args, n_pairs_by_q)
        V, Q, C = (args[2], args[3], args[4]) if name == 'sgc_crossview_mean_fwd' else \
                  (args[3], args[4], args[5]) if name == 'sgc_crossview_attn_fwd' else \
                  (args[3], args[4], args[5]) if name == 'sgc_crossview_attn_bwd_qt' else (args[4], args[5], args[6])
        P = n_pairs_by_q[Q][0]
        if name == 'sgc_crossview_mean_fwd':
            return f * (P * C + Q * C) + 4 * V * Q
        if name == 'sgc_crossview_attn_fwd':
            return f * (P * C + 16 * Q * C + 8 * P) + 4 * V * Q
        if name == 'sgc_crossview_attn_bwd_qt':
            return f * (P * C + 16 * Q * C + 16 * P) + 4 * V * Q
        return f * (P * C + 17 * Q * C + 16 * P) + 4 * V * Q
    if name == 'sgc_upsample2x_occ_fwd':
        X, Y, Z, C = args[1:5]
        return f * (X * Y * Z * C + 8 * X * Y * Z * (C + 1))
    if name == 'sgc_upsample2x_occ_bwd':
        X, Y, Z, C = args[1:5]
        return f * (2 * X * Y * Z * C + 8 * X * Y * Z * (C + 3))
    if name == 'sgc_project_compact':
        V, Q = args[3], args[4]
        return V * Q * (12 + 1 + 4 + 1) + 4 * n_pairs_by_q.get(Q, (0, V))[0]
    if name == 'sgc_topk_select':
        return 5.0 * args[1] * 4
    if name == 'sgc_split_bf16x3':
        return float(args[1]) * args[2] * (4 + 6)   # fp32 read once, three bf16 slots written
    if name == 'sgc_colsum':
        return 4.0 * args[1] * args[2]
    if name == 'sgc_project_tc_fwd':
        V, C, S, N = args[3], args[4], args[5], args[7]
        return 4.0 * V * S * (C + N) + 4.0 * N * C
    if name == 'sgc_rows_wgrad_tc':
        M, N, R, B = args[3], args[7], args[8], args[9]
        return 4.0 * B * R * (M + N) + 4.0 * B * M * N
    if name == 'sgc_rows_gemm_tc':
        R, K, B, N = args[3], args[4], args[5], args[12]
        return 4.0 * B * R * (K + N) + 4.0 * B * N * K
    if name in ('sgc_scatter_add_rows', 'sgc_gather_rows'):
        return f * 3 * args[3] * args[4]
    return 0.0


# -------------------------------------------------------------------------------------------------
# our arm
# -------------------------------------------------------------------------------------------------

def bind_to_gpu_numa_node(local: int):
    """Pin this rank's host threads (and, by the kernel's local-allocation policy plus an explicit preferred-node policy,
    its pinned staging buffers) to the NUMA node the GPU's PCIe root hangs off.  Round 1's e2e scaled 3.4x on 8 GPUs because
    every rank's 270 MB of pinned inputs lived on node 0.  Returns the node or None (single-node host / not discoverable)."""
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = f'{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0'
        node = int(open(f'/sys/bus/pci/devices/{bdf}/numa_node').read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f'/sys/devices/system/node/node{node}/cpulist').read().strip().split(','):
            lo, _, hi = part.partition('-')
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
        try:   # set_mempolicy(MPOL_PREFERRED, {node}): pinned allocations made from now on land on the GPU's node
            import ctypes
            mask = ctypes.c_ulong(1 << node)
            ctypes.CDLL(None, use_errno=True).syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(64))
        except Exception:
            pass
        return node
    except Exception:
        return None


def gpu_numa_node(local: int):
    """NUMA node the GPU's PCIe root hangs off, or None."""
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = f'{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0'
        node = int(open(f'/sys/bus/pci/devices/{bdf}/numa_node').read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def prefer_numa_node(node) -> bool:
    """set_mempolicy(MPOL_PREFERRED, {node}) for this thread (node None: back to MPOL_DEFAULT).  Memory only -- the CPU
    affinity stays as it is, so the CPU baseline of a single-rank run keeps all host cores."""
    try:
        import ctypes
        libc = ctypes.CDLL(None, use_errno=True)
        if node is None:
            return libc.syscall(238, 0, None, ctypes.c_ulong(0)) == 0
        mask = ctypes.c_ulong(1 << node)
        return libc.syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(64)) == 0
    except Exception:
        return False


def run_ours(args):
    import torch.distributed as dist
    from sgcdet_b200 import plugin, synthetic as syn

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py (ours): no CUDA device -- sgcdet_b200 has no CPU path')
    if world == 1:
        from sgcdet_b200 import build
        build.build()   # no-op when the in-tree library is up to date (it normally travels with the snapshot)
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    numa_node = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    cfg = syn.CONFIGS[args.config]
    V = args.views
    B = max(1, args.scenes_per_gpu)
    head = plugin.build_voxel_head(cfg)
    head.load_state_dict(syn.make_state_dict(cfg), strict=True)
    head = head.to(dev)
    head.train(not args.eval_mode)
    params = [p for p in head.parameters()]
    from sgcdet_b200 import functional as SF
    scenes = []
    for b in range(B):
        sc_b = syn.make_scene(cfg, V, seed=1234 + rank * B + b, shift_origin=True).to(dev)
        sc_b.img_meta['sgc_projection'] = SF.compute_projection(sc_b.img_meta).to(dev)  # static buffer for graph replay
        scenes.append(dict(
            sc=sc_b,
            feats=[f.clone().requires_grad_(True) for f in sc_b.mlvl_feats[:cfg.num_levels]],
            dists=[d.clone().requires_grad_(True) for d in sc_b.mlvl_dpt_dists[:cfg.num_levels]],
            # channels_last_3d, like the returned volume
            gvol=sc_b.grad_volume.permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3),
            stream=torch.cuda.Stream(device=dev) if B > 1 else None))
    sc, feats, dists, gvol = scenes[0]['sc'], scenes[0]['feats'], scenes[0]['dists'], scenes[0]['gvol']
    all_inputs = [t for s in scenes for t in s['feats'] + s['dists']]
    loss_buf = torch.zeros(1, device=dev)

    loss_stream = torch.cuda.Stream(device=dev)
    LOSS_ON_SIDE = os.environ.get('SGC_LOSS_ON_SIDE', '1') != '0'

    def scene_fwd(s):
        """Forward of one scene.  loss = sum(volume * G) + occ_loss.  The gradient of the first term
        w.r.t. the volume is G itself, so the backward is seeded with G directly and the VALUE of the first term (needed
        only for the loss read-back) is evaluated beside the backward instead of in front of it."""
        vol, valid, occ = head(s['feats'], s['sc'].img_meta, s['dists'])
        # the occupancy loss (value and gradient) on the loss stream: it is not on the path from the volume to its gradient
        return vol, head.occ_loss(occ, None, s['sc'].geo_occ, stream=loss_stream if LOSS_ON_SIDE else None)['loss_occ']

    def step():
        main = torch.cuda.current_stream()
        if averager is not None:
            averager.begin_step()
        if B == 1:
            outs = [scene_fwd(scenes[0])]
        else:
            # B independent scenes in flight on B streams: the latency-bound per-voxel chains of different
            # scenes overlap; one backward over all of them (autograd replays every node on its own stream)
            outs = []
            for s in scenes:
                s['stream'].wait_stream(main)
                with torch.cuda.stream(s['stream']):
                    outs.append(scene_fwd(s))
            for s in scenes:
                main.wait_stream(s['stream'])
        def loss_value():
            # loss value on its own stream, concurrently with the backward
            with torch.cuda.stream(loss_stream):
                with torch.no_grad():
                    val = outs[0][1].detach().new_zeros(())
                    for (vol, l_occ), s in zip(outs, scenes):
                        vol.record_stream(loss_stream)
                        val = val + (vol * s['gvol']).sum() + l_occ
                    loss_buf.copy_(val.view(1))

        if LOSS_ON_SIDE:
            # the loss stream has joined the end of the forward inside occ_loss(stream=...); the backward is issued first, so that
            # the occupancy loss's gradient kernels (replayed by autograd on the loss stream) are not queued behind the
            # evaluation of the loss VALUE, which nothing in the step waits for
            torch.autograd.backward([t for vol, l_occ in outs for t in (vol, l_occ)],
                                    [g for s in scenes for g in (s['gvol'], None)])
            loss_value()
        else:
            loss_stream.wait_stream(main)
            loss_value()
            torch.autograd.backward([t for vol, l_occ in outs for t in (vol, l_occ)],
                                    [g for s in scenes for g in (s['gvol'], None)])
        main.wait_stream(loss_stream)
        if averager is not None:
            # the peer all-reduces of the parameter groups were issued from the backward as their gradients became final
            # (captured with the step); this averages the remaining tail and joins the communication stream
            averager.finish_step()

    # scene-batch DP: the path's weight gradients (~8 MB) are averaged across ranks every step.
    #   peer (default): own all-reduce kernels over NVLink peer memory (sgcdet_b200/peer.py, csrc/sgc_peer.cu) captured INTO the
    #                   step's CUDA graph, bucketed in the order the gradients become final and overlapped with the end of the
    #                   backward (only a small tail bucket stays behind the last kernel);
    #   nccl:           cat -> ncclAllReduce(avg) -> copy, issued by the CPU after every graph replay (round 1).
    ar_mode = 'none' if (world == 1 or args.no_grad_allreduce) else args.grad_allreduce
    averager = None
    if ar_mode == 'peer':
        from sgcdet_b200 import peer
        try:
            averager = peer.GradAverager(params, device=dev)
        except Exception as e:   # no peer access between the GPUs of this box: say so and use NCCL
            print(f'[bench] peer memory unavailable ({type(e).__name__}: {e}); falling back to NCCL', file=sys.stderr)
            ar_mode = 'nccl'
        ok = torch.tensor([1 if ar_mode == 'peer' else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            ar_mode, averager = 'nccl', None
    ar_in_graph = ar_mode == 'peer' and not args.no_graph
    avg_op = dist.ReduceOp.AVG if world > 1 else None

    def allreduce_grads():
        if ar_mode == 'nccl':
            flat = torch.cat([p.grad.reshape(-1) for p in params])
            dist.all_reduce(flat, op=avg_op)
            torch._foreach_copy_([p.grad.view(-1) for p in params], list(flat.split([p.numel() for p in params])))

    def zero_grads():
        for t in params + all_inputs:
            t.grad = None

    # warm-up (eager) on a side stream, then capture fwd+bwd into one CUDA graph
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(max(1, min(args.warmup, 3))):
            zero_grads()
            step()
            allreduce_grads()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    rec = CallRecorder()
    graph = None
    launches_per_step = 0
    if not args.no_graph:
        zero_grads()
        graph = torch.cuda.CUDAGraph()
        with rec:
            with torch.cuda.graph(graph):
                step()
        launches_per_step = rec.count

        def run_step():
            graph.replay()
            allreduce_grads()
    else:
        def run_step():
            zero_grads()
            step()
            allreduce_grads()
        with rec:
            run_step()
        launches_per_step = rec.count
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        run_step()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    total_ms = timed(run_step, args.steps)
    loss_val = float(loss_buf.item())

    # ---- end to end: host buffers -> H2D -> step -> D2H of the loss, every step -------------------------
    # single rank: the pinned staging buffers are placed on the GPU's NUMA node (memory policy only; multi-rank runs bound the
    # whole process in bind_to_gpu_numa_node): a remote node costs ~6 % of the PCIe rate on these boxes
    e2e_node = numa_node
    if world == 1 and not args.skip_e2e:
        e2e_node = gpu_numa_node(local)
        if e2e_node is not None and not prefer_numa_node(e2e_node):
            e2e_node = None
    host_in = [t.detach().cpu().pin_memory() for t in all_inputs]
    if world == 1 and e2e_node is not None:
        prefer_numa_node(None)
    dev_in = all_inputs
    h2d = sum(t.numel() * t.element_size() for t in host_in)
    loss_host = torch.zeros(1).pin_memory()

    # Double-buffered pipeline: the H2D copy of step k+1 (copy stream, pinned host -> device staging) overlaps the
    # compute of step k; every step still moves its own 270 MB of inputs and reads its own loss back.
    copy_stream = torch.cuda.Stream(device=dev)
    staging = [[torch.empty_like(t) for t in dev_in] for _ in range(2)]
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    losses_host = [torch.zeros(1).pin_memory() for _ in range(2)]

    def issue_copy(k):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[k % 2])          # staging slot free again
            for h, s in zip(host_in, staging[k % 2]):
                s.copy_(h, non_blocking=True)
            copied[k % 2].record(copy_stream)

    def e2e_run(n):
        main = torch.cuda.current_stream()
        for e in consumed:
            e.record(main)
        issue_copy(0)
        for k in range(n):
            if k + 1 < n:
                issue_copy(k + 1)
            main.wait_event(copied[k % 2])
            with torch.no_grad():
                torch._foreach_copy_(dev_in, staging[k % 2])   # device-side hand-off into the graph's static inputs
            consumed[k % 2].record(main)
            run_step()
            losses_host[k % 2].copy_(loss_buf, non_blocking=True)
        main.synchronize()

    e2e_steps = max(4, min(args.steps, 12))
    if args.skip_e2e:
        e2e_ms = float('nan')
    else:
        e2e_run(3)
        e2e_ms = timed(lambda: e2e_run(e2e_steps), 1)
    clk = clocks.stop() if rank == 0 else None

    # ---- instrumented eager pass: per-kernel device time with CUDA events (not part of `value`) --------
    kernels, roof, path_roof = {}, None, None
    ab = syn.algorithmic_bytes(cfg, V)
    if rank == 0 and not args.skip_e2e:
        rec2 = CallRecorder()
        rec2.timed = True
        n_inst = 3
        # every kernel timed ALONE: the side streams / priorities of the product path are switched off for this pass
        serial = {k: '0' for k in ('SGC_SIDE_LIFT_BWD', 'SGC_WSTREAM', 'SGC_SIDE_PREPARE', 'SGC_CHAIN_PRIORITY',
                                   'SGC_SIDE_GRADS')}
        saved_env = {k: os.environ.get(k) for k in serial}
        os.environ.update(serial)
        try:
            with rec2:
                for _ in range(n_inst):
                    zero_grads()
                    vol, valid, occ, its = head(feats, sc.img_meta, dists, return_intermediates=True)
                    ((vol * gvol).sum() + head.occ_loss(occ, None, sc.geo_occ)['loss_occ']).backward()
            torch.cuda.synchronize()
        finally:
            for k, v in saved_env.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        pairs_by_q = {it['pairs'].Q: (int(it['pairs'].view_offsets[-1]), V) for it in its}
        agg = {}
        for name, a, e0, e1 in rec2.events:
            key = name
            if name.startswith(('sgc_lift', 'sgc_crossview')) or name == 'sgc_project_compact':
                q = {'sgc_lift_fwd': 15, 'sgc_lift_bwd': 16, 'sgc_lift_bwd_tiles': 17, 'sgc_crossview_mean_fwd': 3, 'sgc_crossview_attn_fwd': 4,
                     'sgc_crossview_attn_bwd_qt': 4, 'sgc_crossview_attn_bwd_slots': 5, 'sgc_project_compact': 4,
                     'sgc_crossview_mean_fwd_split': 3, 'sgc_crossview_attn_fwd_split': 4,
                     'sgc_crossview_attn_bwd_qt_split': 4}[name]
                key = f'{name}[Q={a[q]}]'
            elif name.startswith('sgc_upsample'):
                key = f'{name}[{a[1]}x{a[2]}x{a[3]}]'
            elif name == 'sgc_project_tc_fwd':
                key = f'{name}[V={a[3]},C={a[4]},S={a[5]},N={a[7]}]'
            elif name in ('sgc_split_bf16x3', 'sgc_colsum'):
                key = f'{name}[{a[1]}x{a[2]}]'
            elif name == 'sgc_rows_wgrad_tc':
                key = f'{name}[B={a[9]},R={a[8]},M={a[3]},N={a[7]}]'
            elif name == 'sgc_rows_gemm_tc':
                key = f'{name}[B={a[5]},R={a[3]},K={a[4]},N={a[12]}]'
            elif name in ('sgc_rowop_fwd', 'sgc_rowop_bwd'):
                key = f'{name}[{a[0]._obj.R}x{a[0]._obj.N}]'
            d = agg.setdefault(key, dict(ms=0.0, n=0, bytes=kernel_algorithmic_bytes(name, a, pairs_by_q)))
            d['ms'] += e0.elapsed_time(e1)
            d['n'] += 1
        mine_ms = sum(d['ms'] for d in agg.values()) / n_inst
        # event brackets around sub-10us launches mostly measure launch latency: rank by total time, and take
        # the dominant kernel among launches that move at least 1 MB
        top = sorted(agg.items(), key=lambda kv: -kv[1]['ms'])
        top = [kv for kv in top if kv[1]['bytes'] >= 1e6] + [kv for kv in top if kv[1]['bytes'] < 1e6]
        peaks = load_peaks()
        for k, d in top[:8]:
            avg = d['ms'] / d['n']
            kernels[k] = dict(avg_ms=round(avg, 4), gbs=round(d['bytes'] / (avg * 1e-3) / 1e9, 1) if avg > 0 else None)
        k0, d0 = top[0]
        avg0 = d0['ms'] / d0['n']
        ach = d0['bytes'] / (avg0 * 1e-3) / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum of the same kernel at the same shape from the committed
        # `ncu --set full` capture (profiles/r2_ncu_traffic.json, written by tools/summarize_ncu.py); None if absent
        traffic = None
        tpath = os.path.join(ROOT, 'profiles', 'r2_ncu_traffic.json')
        if os.path.exists(tpath) and args.config == 'SGCDet_ScanNet' and V == 40:
            traffic = json.load(open(tpath)).get(k0)
        roof = dict(bound='hbm', kernel=k0, achieved=round(ach, 1), peak=peaks['hbm_gbs'], unit='GB/s',
                    frac=round(ach / peaks['hbm_gbs'], 4), traffic=traffic, peak_source=peaks['source'],
                    algorithmic_bytes_per_launch=int(d0['bytes']), avg_launch_ms=round(avg0, 4),
                    timing='CUDA events around each launch in an eager instrumented pass after the timed region')
        step_ms = total_ms / args.steps
        path_roof = dict(algorithmic_bytes_fwd_bwd=int(ab['fwd_bwd']), achieved_gbs=round(B * ab['fwd_bwd'] / (step_ms * 1e-3) / 1e9, 1),
                         frac_of_hbm=round(B * ab['fwd_bwd'] / (step_ms * 1e-3) / 1e9 / peaks['hbm_gbs'], 4),
                         own_kernels_ms_per_step=round(mine_ms, 3), pairs_per_level=[pairs_by_q[q][0] for q in sorted(pairs_by_q)])

    # ---- CPU baseline (rank 0, N == 1 only): the oracle port on the host cores, bounded sample --------
    cpu = None
    loss_check = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.config, V, volumes=1, want_masks=True)
        # one eval-mode, teacher-forced step of the product on the SAME scene the oracle just evaluated: the loss of the
        # line's workload against the oracle's (the timed steps run in train mode, whose FFN dropout makes the loss random)
        o_loss, o_masks = cpu.pop('_loss'), cpu.pop('_masks')
        was_training = head.training
        head.eval()
        with torch.no_grad():
            forced = [None] + [torch.nonzero(o_masks[i].view(-1) > 0).view(-1).to(dev, torch.int32)
                               for i in range(1, cfg.num_levels)]
            sc0 = syn.make_scene(cfg, V, shift_origin=True).to(dev)   # cpu_baseline's scene (default seed)
            vol_e, _, occ_e = head(sc0.mlvl_feats[:cfg.num_levels], sc0.img_meta, sc0.mlvl_dpt_dists[:cfg.num_levels],
                                   forced_selection=forced)
            p_loss = float((vol_e * sc0.grad_volume).sum() + head.occ_loss(occ_e, None, sc0.geo_occ)['loss_occ'])
        head.train(was_training)
        loss_check = dict(product_eval_loss=round(p_loss, 4), oracle_loss=round(o_loss, 4),
                          loss_vs_oracle_rel=abs(p_loss - o_loss) / max(abs(o_loss), 1e-12),
                          how='eval mode, teacher-forced with the oracle selection, same scene and weights')

    # ---- the reference's own kernels on the same GPU (rank 0, N == 1, after the timed region) ---------------
    ref_gpu, op_bench, depth_bench = None, None, None
    if rank == 0 and world == 1 and not args.no_reference_gpu:
        try:
            from oracle import gpu_ref
            r = gpu_ref.bench(args.config, V, 3, 1)
            ref_gpu = dict(value=r['value'], unit=UNIT, ms_per_step=r['ms_per_step'], kind='reference kernels',
                           what=r['config']['what'])
        except Exception as e:   # the checker must never break the product line
            ref_gpu = dict(unavailable=f'{type(e).__name__}: {e}'[:200])
        try:
            op_bench = operator_bench(dev)
        except Exception as e:
            op_bench = dict(unavailable=f'{type(e).__name__}: {e}'[:200])
        try:
            depth_bench = depth_producer_bench(dev)
        except Exception as e:
            depth_bench = dict(unavailable=f'{type(e).__name__}: {e}'[:200])

    # ---- config 5: view-sharded aggregation of ONE scene over the N ranks (every rank takes part) ---------
    vs = None
    if not args.no_view_sharded:
        try:
            vs = view_sharded_leg(args, rank, world, dev)
        except Exception as e:   # an optional leg must never break the headline line
            vs = dict(unavailable=f'{type(e).__name__}: {e}'[:300])

    # ---- config 3: full ARKit-shaped train step around the view transform, scene-batch DP -----------------
    ts = None
    if not args.no_train_step:
        try:
            ts = train_step_leg(args, rank, world, dev)
        except Exception as e:
            ts = dict(unavailable=f'{type(e).__name__}: {e}'[:300])

    if rank == 0:
        step_ms = total_ms / args.steps
        line = {
            'metric': METRIC, 'value': round(world * B * args.steps / (total_ms * 1e-3), 2), 'unit': UNIT,
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': round(step_ms, 4),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(cfg, V, B, world),
            'implementation': {
                'grad_average': 'none' if ar_mode == 'none' else (
                    'own peer-memory all-reduce kernels over NVLink, issued from the backward and overlapped with it'
                    + (' inside the CUDA graph' if ar_in_graph else '') if ar_mode == 'peer' else 'NCCL all-reduce after the replay'),
                'loss': 'sum(volume*G) + occ_loss every step; backward seeded with G (the gradient of the first term); the loss value '
                        'and the occupancy loss (forward and, through autograd, backward) run on a side stream beside the backward',
                'mode': 'eval' if args.eval_mode else 'train (FFN dropout 0.1 active; keep-masks of all levels drawn by one own Philox '
                                                     'launch per step, fresh on every graph replay)',
                'cuda_graph': not args.no_graph,
                'gemm': 'all GEMMs of the path are own tcgen05/TMEM/TMA kernels with the bf16 hi/lo split in shared memory and fp32 '
                        'accumulation: feature-map projection (forward, data gradient, weight gradient) and the voxel-count layers '
                        '(forward, data and weight gradients, 16- and 32-wide heads)',
                'streams': 'per-voxel chain on a high-priority stream; projections, lift backward and weight gradients on side '
                           'streams; one CUDA graph'},
            'e2e': {'value': None if args.skip_e2e else round(world * B * e2e_steps / (e2e_ms * 1e-3), 2), 'unit': UNIT, 'h2d_bytes_per_step': int(h2d),
                    'd2h_bytes_per_step': 4, 'steps': e2e_steps, 'numa_node_rank0': e2e_node,
                    'pipeline': 'double-buffered: H2D of step k+1 (copy stream) overlaps compute of step k'},
            'gpu_launches': int(launches_per_step * args.steps),
            'gpu_launches_per_step': int(launches_per_step),
            'clocks': clk, 'roofline': roof, 'path_roofline': path_roof, 'kernels': kernels, 'cpu_baseline': cpu,
            'loss': round(loss_val, 4), 'loss_check': loss_check,
            'loss_vs_oracle_rel': None if loss_check is None else loss_check['loss_vs_oracle_rel'],
            'reference_gpu': ref_gpu, 'operator_bench': op_bench, 'depth_producer_bench': depth_bench, 'view_sharded': vs, 'train_step': ts,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def train_step_leg(args, rank, world, dev):
    """BASELINE.json configs[2]: ``SGCDet_ARKit`` FULL train step with this view transform, one scene per rank (scene-batch
    DP): images -> ResNet-50 + FPN -> depth distribution -> ``AdaptiveSparseHead`` (the path) -> 3-D neck -> detection-head
    losses + ``occ_loss`` -> backward -> NCCL gradient all-reduce -> AdamW + OneCycleLR (``sgcdet_b200/harness.py``; everything
    but the view transform is a stand-in with the reference's tensor shapes, see its docstring).  Eager, train mode,
    device-timed, max over ranks; value = scenes per second over all ranks."""
    import torch.distributed as dist
    from sgcdet_b200 import harness, synthetic as syn
    cfg = syn.CONFIGS['SGCDet_ARKit']
    V = 40
    torch.manual_seed(1234)                       # identical initial weights on every rank
    model = harness.SGCDetShaped(cfg).to(dev).train()
    model.voxel_head.load_state_dict(syn.make_state_dict(cfg), strict=True)
    batch = harness.make_batch(cfg, V, dev, seed=1234 + rank)
    opt, sched = harness.configure_optimizers(model, total_steps=1000)
    params = [p for p in model.parameters() if p.requires_grad]
    warm, steps = 3, 5
    for _ in range(warm):
        loss = harness.train_step(model, batch, opt, sched, params, world)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = harness.train_step(model, batch, opt, sched, params, world)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.barrier()
    if rank != 0:
        return None
    n_par = sum(p.numel() for p in params)
    return dict(config=cfg.name, views=V, n_gpus=world, ms_per_step=round(float(ms.item()), 3),
                value=round(world * 1e3 / float(ms.item()), 3), unit='scenes/s', loss=round(float(loss), 4),
                trainable_parameters=int(n_par), grad_allreduce_bytes=int(4 * n_par) if world > 1 else 0,
                parallelism=f'scene-batch dp{world}' + (' + NCCL gradient all-reduce' if world > 1 else ''),
                mode='eager, train, images -> losses -> backward -> all-reduce -> AdamW/OneCycleLR step; stand-in backbone / '
                     'depth / neck / head around the real view transform')


def view_sharded_leg(args, rank, world, dev):
    """BASELINE.json configs[4]: ``SGCDet_large_ARKit`` with the V views of ONE scene split over the N ranks: projection / lift
    local to a view's owner, the cross-view statistics (sum + count, score max, partial-softmax sums -- the log-sum-exp merge --
    and in the backward the softmax-normaliser dot and the query gradient) exchanged as single kernel launches over NVLink peer
    memory (``sgcdet_b200/peer.py``) inside the fused level, the voxel chain replicated; the parameter gradients that are
    partial sums over views are reduced the same way.  The whole step (forward + backward + all exchanges) is ONE CUDA graph.
    Eval mode (the replicated chain must be identical on every rank), device-timed, max over ranks.  Rank 0 also times the
    unsharded step of the same scene (one CUDA graph as well) on its GPU and compares volume and gradients."""
    import torch.distributed as dist
    from sgcdet_b200 import functional as SF, parallel, plugin, synthetic as syn
    cfg = syn.CONFIGS[args.view_sharded_config]
    V = args.view_sharded_views
    sc = syn.make_scene(cfg, V, shift_origin=True).to(dev)
    head = plugin.build_voxel_head(cfg)
    head.load_state_dict(syn.make_state_dict(cfg), strict=True)
    head = head.to(dev).eval()
    views = parallel.shard_views(V, world, rank)
    f, m, d = parallel.shard_scene_inputs(sc.mlvl_feats, sc.img_meta, sc.mlvl_dpt_dists, views)
    m['sgc_projection'] = SF.compute_projection(m).to(dev)    # static device buffer: no H2D copy inside the step
    f = [t.requires_grad_(True) for t in f[:cfg.num_levels]]
    d = [t for t in d[:cfg.num_levels]]
    gvol = sc.grad_volume.permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3)
    xch = parallel.ViewShardExchange(head, device=dev)
    params = list(head.parameters())

    def make_step(feats, meta, dists, shard, forced=None, want_sel=False):
        def step():
            out = head(feats, meta, dists, view_shard=shard, forced_selection=forced, return_intermediates=want_sel)
            vol, valid, occ = out[:3]
            loss = (vol * gvol).sum() + head.occ_loss(occ, None, sc.geo_occ)['loss_occ']
            loss.backward()
            if shard is not None:
                shard.reduce_gradients(head)
            return (vol, valid, out[3]) if want_sel else (vol, valid)
        return step

    def capture(step, inputs):
        """Eager warm-up on a side stream, then the step as one CUDA graph; returns (replay, outputs, 'graph' | 'eager: why')."""
        def zero():
            for p in params + inputs:
                p.grad = None
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                zero()
                out = step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if args.no_graph:
            return (lambda: (zero(), step())[1]), out, 'eager (--no-graph)'
        try:
            zero()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = step()
            return (lambda: (g.replay(), out)[1]), out, 'one CUDA graph'
        except Exception as e:   # say so and time the eager step instead
            torch.cuda.synchronize()
            return (lambda: (zero(), step())[1]), out, f'eager (capture failed: {type(e).__name__}: {e})'[:200]

    def timed(fn, warm, steps, sync_ranks):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        if sync_ranks and world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        if sync_ranks and world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- parity first (untimed, eager): the unsharded step of the whole scene on rank 0 fixes the voxel selection; every
    # rank then runs the sharded step teacher-forced with it (a free-running top-k may legitimately pick different voxels
    # among near-ties, which makes volumes incomparable), and rank 0 compares volume and every parameter gradient
    feats = [t.clone().requires_grad_(True) for t in sc.mlvl_feats[:cfg.num_levels]]
    meta_full = dict(sc.img_meta)
    meta_full['sgc_projection'] = SF.compute_projection(sc.img_meta).to(dev)
    dists_full = sc.mlvl_dpt_dists[:cfg.num_levels]
    sels = [torch.empty(min(k, h.num_voxels), device=dev, dtype=torch.int32) for k, h in zip(head.topk_list, head.base_heads[1:])]
    vol_r = gref = None
    if rank == 0:
        vol_r, _, its = make_step(feats, meta_full, dists_full, None, want_sel=True)()
        gref = torch.cat([p.grad.reshape(-1) for p in params if p.grad is not None]).clone()
        vol_r = vol_r.detach().clone()
        for dst, it in zip(sels, its[1:]):
            dst.copy_(it['sel'])
    if world > 1:
        for t in sels:
            dist.broadcast(t, src=0)
    for p in params + f:
        p.grad = None
    vol_f, valid_f = make_step(f, m, d, xch, forced=[None] + sels)()
    torch.cuda.synchronize()
    xch.mem.check()
    gflat = torch.cat([p.grad.reshape(-1) for p in params if p.grad is not None]).clone()
    n_grads = sum(1 for p in params if p.grad is not None)
    parity = None
    if rank == 0:
        parity = dict(volume_max_abs_diff_vs_unsharded=float((vol_f.detach() - vol_r).abs().max()),
                      volume_rel_diff_vs_unsharded=float((vol_f.detach() - vol_r).norm() / vol_r.norm()),
                      param_grad_rel_diff_vs_unsharded=float((gflat - gref).norm() / gref.norm().clamp(min=1e-30)),
                      param_grads_compared=n_grads, how='eager, teacher-forced with the unsharded selection')

    # ---- timed: free-running selection, the whole step one CUDA graph ------------------------------------------------------
    replay, (vol, valid), mode = capture(make_step(f, m, d, xch), f)
    ms = timed(replay, 3, 20, True)
    xch.mem.check()
    # replicated chain: volume / valid mask AND (after reduce_gradients) every parameter gradient bit-identical on all ranks
    gflat = torch.cat([p.grad.reshape(-1) for p in params if p.grad is not None])
    chk = torch.stack([vol.detach().double().sum(), vol.detach().double().abs().sum(), valid.double().sum(),
                       gflat.double().sum(), gflat.double().abs().sum()])
    lo, hi = chk.clone(), chk.clone()
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    out = None
    if rank == 0:
        replay1, _, mode1 = capture(make_step(feats, meta_full, dists_full, None), feats)
        ms1 = timed(replay1, 3, 20, False)
        out = dict(config=cfg.name, views=V, n_gpus=world, views_per_rank=len(views), ms_per_step=round(ms, 3),
                   value=round(1e3 / ms, 2), unit='volumes/s', mode=mode + ', eval, fwd+bwd, device-timed, max over ranks',
                   unsharded_1gpu_ms=round(ms1, 3), unsharded_mode=mode1, speedup_vs_unsharded=round(ms1 / ms, 3),
                   replicas_identical=dict(volume=bool(torch.equal(lo[:3], hi[:3])), gradients=bool(torch.equal(lo[3:], hi[3:]))),
                   parity=parity,
                   collective='none (1 rank)' if world == 1 else
                   'own one-launch all-reduce over NVLink peer memory (sum, max), 5 exchanges per level + 1 for the parameter '
                   'gradients, inside the graph',
                   exchange_bytes_per_rank_per_step=int(4 * sum(2 * q * (cfg.embed_dims + 8) + q * (cfg.embed_dims + 1) + 2 * q * 8
                                                                for q in [math.prod(cfg.n_voxels_list[0])] + list(cfg.topk_list))))
    if world > 1:
        dist.barrier()
    xch.close()
    return out


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=float(d['hbm_gbs']), source='MEASURED_PEAKS.json (of measured)')
    return dict(hbm_gbs=6650.0, source='B200_PROFILING.md fallback (of fallback)')


# -------------------------------------------------------------------------------------------------
# CPU arms
# -------------------------------------------------------------------------------------------------

def cpu_baseline(config: str, V: int, volumes: int = 1, warm: bool = False, want_masks: bool = False):
    """The oracle port (torch CPU fp32, every host thread) on a bounded sample of the same workload.  With ``want_masks``
    the dict also carries ``_loss`` / ``_masks`` of the (eval-mode) oracle pass for the loss cross-check of the line."""
    from oracle import path_ref
    from sgcdet_b200 import synthetic as syn
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = syn.CONFIGS[config]
    sc = syn.make_scene(cfg, V, shift_origin=True)
    sd = syn.make_state_dict(cfg)
    sdg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'ref_3d' not in k else v) for k, v in sd.items()}
    feats = [f.clone().requires_grad_(True) for f in sc.mlvl_feats]
    dists = [d.clone().requires_grad_(True) for d in sc.mlvl_dpt_dists]

    keep = {}

    def one():
        vol, valid, occ, inter = path_ref.adaptive_sparse_head_forward(sdg, feats, sc.img_meta, dists, cfg,
                                                                     return_intermediates=True)
        loss = (vol * sc.grad_volume).sum() + path_ref.occ_loss(occ, sc.geo_occ)
        loss.backward()
        keep['loss'], keep['masks'] = float(loss.detach()), inter['masks']
        return keep['loss']

    if warm:
        one()
    t0 = time.perf_counter()
    for _ in range(volumes):
        one()
    dt = time.perf_counter() - t0
    out = dict(value=round(volumes / dt, 4), unit=UNIT, cores=cores, kind='port',
               sample=f'{volumes} volume(s) fwd+bwd of {cfg.name} V={V} on the CPU oracle port '
                      f'(torch {torch.__version__} CPU fp32, {torch.get_num_threads()} threads), {dt:.1f} s')
    if want_masks:
        out['_loss'], out['_masks'] = keep['loss'], keep['masks']
    return out


def operator_bench(dev, iters: int = 5):
    """DFA3D operator boundary (B2) at the reference's own unit-test shapes (unittest_DFA3D.py:43-56): the reference's
    two-stage kernels (oracle/_ref, unmodified, sm_100a) vs this library's fused one-stage kernels, forward and backward,
    CUDA events, same tensors.  GB/s = algorithmic bytes (value + depth maps read once, per-query tensors once) / time."""
    from oracle import build_ref
    import sgcdet_b200
    ext = build_ref.load()
    if ext is None:
        return {'unavailable': 'oracle/_ref/dfa3d_ref_ext.so not built'}
    sgcdet_b200.install_dropin()
    from dfa3D import ext_loader
    mine = ext_loader.load_ext('_ext', ['wms_deform_attn_forward'])
    B, M, Cm, D, Q, P = 6, 8, 32, 112, 9502, 8
    shapes = [(116, 200), (58, 100), (29, 50), (15, 25)]
    g = torch.Generator().manual_seed(5)
    s3 = torch.tensor([[h, w, D] for h, w in shapes], dtype=torch.long)
    sizes = s3[:, 0] * s3[:, 1]
    lsi = torch.cat([sizes.new_zeros(1), sizes.cumsum(0)[:-1]]).to(dev)
    S, L = int(sizes.sum()), len(shapes)
    s3 = s3.to(dev)
    value = torch.randn(B, S, M, Cm, generator=g).to(dev)
    dist = torch.randn(B, S, M, D, generator=g).softmax(-1).to(dev)
    loc = ((torch.rand(B, Q, M, L, P, 3, generator=g) - 0.5) * 1.2 + 0.5).to(dev)
    attn = torch.rand(B, Q, M, L * P, generator=g).softmax(-1).view(B, Q, M, L, P).to(dev)
    gout = torch.randn(B, Q, M * Cm, generator=g).to(dev)
    s2, loc2 = s3[..., :2].contiguous(), loc[..., :2].contiguous()

    def ref_fwd():
        ds = ext.ms_depth_score_sample_forward(dist, s3, lsi, loc, im2col_step=64)
        return ext.wms_deform_attn_forward(value, s2, lsi, loc2, attn, ds, im2col_step=64), ds

    def ref_bwd(ds):
        gv, gl2, ga, gds = torch.zeros_like(value), torch.zeros_like(loc2), torch.zeros_like(attn), torch.zeros_like(ds)
        ext.wms_deform_attn_backward(value, s2, lsi, loc2, attn, ds, gout, gv, gl2, ga, gds, im2col_step=64)
        gd, gl = torch.zeros_like(dist), torch.zeros_like(loc)
        ext.ms_depth_score_sample_backward(dist, s3, lsi, loc, gds, gd, gl, im2col_step=64)
        return gv, gd

    def my_fwd():
        return mine.dfa3d_fused_forward(value, dist, s3, lsi, loc, attn)

    def my_bwd(_):
        gv, gd, gl, ga = torch.zeros_like(value), torch.zeros_like(dist), torch.zeros_like(loc), torch.zeros_like(attn)
        mine.dfa3d_fused_backward(value, dist, s3, lsi, loc, attn, gout, gv, gd, gl, ga)
        return gv, gd

    def t(fn, *a):
        fn(*a)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            r = fn(*a)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters, r

    f = 4.0
    maps = f * B * S * M * (Cm + D)
    perq = f * B * Q * M * (Cm + L * P * 8)
    r_f, (out_r, ds_r) = t(ref_fwd)
    m_f, (out_m, ds_m) = t(my_fwd)
    r_b, (gv_r, gd_r) = t(ref_bwd, ds_r)
    m_b, (gv_m, gd_m) = t(my_bwd, None)
    err = float((out_r - out_m).abs().max()), float((gv_r - gv_m).abs().max()), float((gd_r - gd_m).abs().max())
    gbs = lambda ms, nbytes: round(nbytes / (ms * 1e-3) / 1e9, 1)
    return {'shape': f'unittest_DFA3D.py:43-56  B={B} S={S} (4 levels) M={M} Cm={Cm} D={D} Q={Q} P={P}',
            'fwd': {'reference_two_stage_ms': round(r_f, 3), 'ours_fused_ms': round(m_f, 3),
                    'reference_gbs': gbs(r_f, maps + perq), 'ours_gbs': gbs(m_f, maps + perq)},
            'bwd': {'reference_two_stage_ms': round(r_b, 3), 'ours_fused_ms': round(m_b, 3),
                    'reference_gbs': gbs(r_b, 2 * maps + perq), 'ours_gbs': gbs(m_b, 2 * maps + perq),
                    'note': 'both sides include the zero-fill of their gradient buffers (caller-zeroed contract, F3D:319-339)'},
            'max_abs_diff': {'out': err[0], 'grad_value': err[1], 'grad_dist': err[2]}}


def depth_producer_bench(dev, iters: int = 5):
    """SURVEY.md 8f rank 1 at the SGCDet_ScanNet shape (V=40 frames, 128 matching channels on the 60x80 map, K=2 neighbours,
    D=12 depth bins): the fused plane-sweep kernels + the softmax / pyramid kernel against the reference FORMULATION in eager
    PyTorch on the same GPU (oracle/depth_ref.py: homography grids + F.grid_sample + the [V,C,D,H,W] product, which is what
    depth_est_fusion.py:85-126,218-232 executes).  Algorithmic bytes = feature map read + correlation written (forward),
    + gradient map written and correlation gradient read (backward)."""
    from oracle import depth_ref
    from sgcdet_b200 import depth as SD, functional as SF, synthetic as syn
    cfg = syn.CONFIGS['SGCDet_ScanNet']
    V, C, H, W, K = 40, 128, 60, 80, 2
    g = torch.Generator().manual_seed(99)
    meta = syn.make_img_meta(cfg, V, g)
    base = meta['lidar2img']['extrinsic'][0]
    ext = []
    for i in range(V):           # a video-like trajectory: consecutive frames overlap
        T = np.eye(4, dtype=np.float32)
        T[0, 3], T[2, 3] = 0.04 * i, -0.02 * i
        ext.append((T @ base).astype(np.float32))
    meta['lidar2img']['extrinsic'] = ext
    depth = SD.depth_bin_centers(cfg.dbound)
    D = int(depth.shape[0])
    f = torch.randn(V, C, H, W, generator=g).to(dev)
    gc = torch.randn(V, D, H, W, generator=g).to(dev)
    intr = depth_ref.feature_intrinsic(torch.tensor(meta['lidar2img']['intrinsic']), meta['img_shape'], meta['ori_shape'], 4).to(dev)
    w2c = torch.tensor(np.stack(ext)).to(dev)
    dv = torch.tensor(depth).to(dev)

    def t_ms(fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    fo = f.clone().requires_grad_(True)
    fr = f.clone().requires_grad_(True)
    ours = SD.plane_sweep_correlation(fo, meta, 4, K, depth)
    ref = depth_ref.plane_sweep_correlation(fr, w2c, intr, dv, K)
    ours.backward(gc, retain_graph=True)
    ref.backward(gc, retain_graph=True)
    diff = dict(correlation=float((ours - ref).abs().max()), grad_feat=float((fo.grad - fr.grad).abs().max()),
                inside_fraction=float((ref != 0).float().mean()))
    with torch.no_grad():
        ms_f = t_ms(lambda: SD.plane_sweep_correlation(f, meta, 4, K, depth))
        ms_fr = t_ms(lambda: depth_ref.plane_sweep_correlation(f, w2c, intr, dv, K))

    def bwd(x, y):
        x.grad = None
        y.backward(gc, retain_graph=True)
    ms_b = t_ms(lambda: bwd(fo, ours))
    ms_br = t_ms(lambda: bwd(fr, ref))
    bf = 4.0 * (V * C * H * W + V * D * H * W)
    logits = torch.randn(V, D, H, W, generator=g).to(dev)
    crops = tuple(SD.pyramid_crops(meta))
    with torch.no_grad():
        ms_p = t_ms(lambda: SF.DepthPyramid.apply(logits, crops))
        ms_pr = t_ms(lambda: depth_ref.depth_pyramid(logits, crops))
    bp = 4.0 * V * D * (2 * H * W + sum(h * w for h, w in crops))
    return dict(shape=f'V={V} C={C} HxW={H}x{W} K={K} D={D} (configs/SGCDet_ScanNet.py:3,91)',
                plane_sweep_fwd=dict(reference_formulation_ms=round(ms_fr, 3), ours_fused_ms=round(ms_f, 3),
                                     ours_gbs=round(bf / ms_f / 1e6, 1)),
                plane_sweep_bwd=dict(reference_formulation_ms=round(ms_br, 3), ours_fused_ms=round(ms_b, 3),
                                     ours_gbs=round(2 * bf / ms_b / 1e6, 1)),
                softmax_pyramid_fwd=dict(reference_formulation_ms=round(ms_pr, 3), ours_fused_ms=round(ms_p, 3),
                                         ours_gbs=round(bp / ms_p / 1e6, 1)),
                max_abs_diff=diff)


def workload_config(cfg, V: int, B: int, world: int) -> dict:
    """The line's ``config``: names the workload only, identical for the product arm and the reference arm."""
    return {'workload': f'{cfg.name} view-transform fwd+bwd, V={V} views, {B} scene(s) per GPU per step',
            'scenes_per_gpu': B, 'embed_dims': cfg.embed_dims, 'n_voxels': list(cfg.n_voxels_list[-1]),
            'topk': list(cfg.topk_list), 'parallelism': f'scene-batch dp{world}',
            'l2': 'inputs larger than L2 (>= 0.33 GB of maps per step, no flush)'}


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if rank != 0:
        return
    from sgcdet_b200 import synthetic as syn
    cfg = syn.CONFIGS[args.config]
    steps = max(1, min(args.steps, 3))
    warm = max(0, min(args.warmup, 1))
    t0 = time.perf_counter()
    cpu = cpu_baseline(args.config, args.views, volumes=steps, warm=warm > 0)
    v = cpu['value']
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': world, 'steps': steps, 'warmup': warm,
        'ms_per_step': round(1e3 / v, 2), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': workload_config(cfg, args.views, 1, world),
        'implementation': {'what': 'CPU restatement of the reference path (oracle/path_ref.py, torch CPU fp32, all host threads); the '
                                   'reference has no CPU implementation: DFA3D is CUDA only'},
        'cpu_baseline': cpu,
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'note': f'steps/warmup clamped to {steps}/{warm} so the run ends within minutes; rank 0 only',
    }
    print(json.dumps(line))


def run_reference_gpu(args):
    """Extra arm: the reference's own CUDA kernels (oracle/_ref, unmodified, sm_100a) under the restated
    reference glue (per-view Python loops, padded rebatch, torch MHA) on the same GPU."""
    try:
        from oracle import gpu_ref
    except ImportError as e:
        print(json.dumps({'impl': 'reference-gpu', 'unavailable': str(e)}))
        return
    line = gpu_ref.bench(args.config, args.views, args.steps, args.warmup)
    line.update({'impl': 'reference-gpu', 'metric': METRIC, 'unit': UNIT})
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference', 'reference-gpu'])
    ap.add_argument('--config', default='SGCDet_ScanNet')
    ap.add_argument('--views', type=int, default=40)
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--scenes-per-gpu', type=int, default=1, help='independent scenes in flight per GPU per step')
    ap.add_argument('--eval-mode', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-grad-allreduce', action='store_true')
    ap.add_argument('--grad-allreduce', default=os.environ.get('SGC_GRAD_ALLREDUCE', 'peer'), choices=['peer', 'nccl'],
                    help='N > 1: weight-gradient average as one peer-memory kernel inside the CUDA graph, or NCCL after the replay')
    ap.add_argument('--no-view-sharded', action='store_true', help='skip the view-sharded leg (config 5)')
    ap.add_argument('--no-train-step', action='store_true', help='skip the full train-step leg (config 3)')
    ap.add_argument('--view-sharded-views', type=int, default=40)
    ap.add_argument('--view-sharded-config', default='SGCDet_large_ARKit')
    ap.add_argument('--no-reference-gpu', action='store_true', help='skip the reference-kernel leg and the operator micro-bench')
    ap.add_argument('--skip-e2e', action='store_true', help='profiling runs only: skip the e2e and instrumented passes')
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'ours':
        args.warmup = 3
    # stdout carries exactly ONE JSON line: libraries that write to file descriptor 1 themselves (NCCL prints its version
    # there when NCCL_DEBUG is set in the environment) are sent to stderr for the duration of the run
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    captured = []
    builtin_print = print

    def emit(*a, **k):
        captured.append(' '.join(str(x) for x in a))
    import builtins
    builtins.print = emit
    try:
        if args.impl == 'reference':
            run_reference(args)
        elif args.impl == 'reference-gpu':
            run_reference_gpu(args)
        else:
            run_ours(args)
    finally:
        builtins.print = builtin_print
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    for line in captured:
        print(line, flush=True)


if __name__ == '__main__':
    main()
