"""Peer-memory collectives (csrc/sgc_peer.cu, sgcdet_b200/peer.py) and the view-sharded product path built on them
(SURVEY.md 8e, config 5), checked on ONE GPU: the "ranks" are simulated inside the process -- every rank has its own
symmetric allocation and issues its launches on its own stream, where the kernels meet at the flags exactly like ranks on
different GPUs do (the spin loops give up after ~4 s, so a failure cannot hang the device).  The real multi-GPU run of the
same code is bench.py's view-sharded leg / tools/check_view_sharded.py (torchrun, 2+ GPUs)."""

import pytest
import torch

from sgcdet_b200 import parallel, peer, plugin, synthetic as syn

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-3, 1e-4
DEV = 'cuda'


@pytest.mark.parametrize('world,n', [(1, 1000), (2, 4), (2, 100003), (4, 65536 + 7), (8, 40000 + 3), (2, 140000 + 1),
                                     (3, 70001), (8, 80000 + 1), (4, 280000 + 2)])
def test_peer_allreduce_simulated_ranks(cuda_lib, world, n):
    """One-shot (<= 2 ranks or < 256 KB) and two-shot (reduce-scatter + all-gather) variants, sum / max, odd sizes.  The
    sizes keep world x CTAs-per-launch below the SM count: simulated ranks share ONE GPU, and all their spinning CTAs have to
    be resident at the same time (on real ranks every GPU holds only its own launch)."""
    mems = peer.PeerMemory.simulate(world, 4 * n)
    try:
        g = torch.Generator(device='cpu').manual_seed(7)
        parts = [torch.randn(n, generator=g).to(DEV) for _ in range(world)]
        streams = [torch.cuda.Stream() for _ in range(world)]
        for op in ('sum', 'max', 'sum'):       # the same pad is reused by consecutive collectives
            outs = []
            torch.cuda.synchronize()
            for r in range(world):
                with torch.cuda.stream(streams[r]):
                    mems[r].view((n,)).copy_(parts[r])
                    outs.append(mems[r].all_reduce(n, torch.empty(n, device=DEV), op, 0.5 if op == 'sum' else 1.0))
            torch.cuda.synchronize()
            for m in mems:
                m.check()
            ref = parts[0].clone()
            for p in parts[1:]:               # rank order: bit-exact
                ref = torch.maximum(ref, p) if op == 'max' else ref + p
            if op == 'sum':
                ref = ref * 0.5
            for o in outs:
                assert torch.equal(o, ref)
    finally:
        for m in mems:
            m.close()


def _run_unsharded(head, sc):
    feats = [f.clone().requires_grad_(True) for f in sc.mlvl_feats]
    dists = [d.clone().requires_grad_(True) for d in sc.mlvl_dpt_dists]
    vol, valid, occ, its = head(feats, sc.img_meta, dists, return_intermediates=True)
    loss = (vol * sc.grad_volume).sum() + head.occ_loss(occ, None, sc.geo_occ)['loss_occ']
    loss.backward()
    grads = {k: p.grad.clone() for k, p in head.named_parameters()}
    head.zero_grad(set_to_none=True)
    return vol.detach(), valid, occ.detach(), [None] + [it['sel'] for it in its[1:]], grads, [f.grad for f in feats], \
        [d.grad for d in dists]


def _close(name, a, b):
    # see tests/test_gpu_view_sharded.py: a handful of ReLU gates flip for pre-activations within round-off of zero, so the
    # gradients are compared norm-wise plus a loose bound on the worst entry
    scale = b.abs().max().item() + 1e-12
    rel = ((a - b).norm() / (b.norm() + 1e-12)).item()
    assert rel < 5e-3, f'{name}: relative Frobenius error {rel}'
    assert ((a - b).abs().max() / scale).item() < 2e-2, name


@pytest.mark.parametrize('cfg_name,V', [('tiny', 12), ('tiny256', 12)])
def test_view_sharded_product_path_single_rank(cuda_lib, cfg_name, V):
    """AdaptiveSparseHead.forward(view_shard=...) with ONE rank owning all views: every exchange runs (over the rank's own
    buffer), so the sharded kernels, the finishing steps and the hand-written backward are checked against the unsharded
    path deterministically inside this process."""
    cfg = syn.CONFIGS[cfg_name]
    sc = syn.make_scene(cfg, V, shift_origin=True).to(DEV)
    head = plugin.build_voxel_head(cfg)
    head.load_state_dict(syn.make_state_dict(cfg))
    head = head.to(DEV).eval()
    vol, valid, occ, forced, ref_grads, ref_gfeat, ref_gdist = _run_unsharded(head, sc)
    xch = parallel.ViewShardExchange(head, mem=peer.PeerMemory.simulate(1, parallel.ViewShardExchange.required_bytes(head))[0])
    try:
        f = [t.clone().requires_grad_(True) for t in sc.mlvl_feats]
        d = [t.clone().requires_grad_(True) for t in sc.mlvl_dpt_dists]
        vol_s, valid_s, occ_s = head(f, sc.img_meta, d, forced_selection=forced, view_shard=xch)
        ((vol_s * sc.grad_volume).sum() + head.occ_loss(occ_s, None, sc.geo_occ)['loss_occ']).backward()
        xch.reduce_gradients(head)
        torch.cuda.synchronize()
        xch.mem.check()
        assert torch.equal(valid_s, valid)
        torch.testing.assert_close(occ_s, occ, rtol=RTOL, atol=ATOL)
        torch.testing.assert_close(vol_s.detach(), vol, rtol=RTOL, atol=ATOL)
        for k, p in head.named_parameters():
            _close(k, p.grad, ref_grads[k])
        for lvl in range(cfg.num_levels):
            _close(f'feat{lvl}', f[lvl].grad, ref_gfeat[lvl])
            _close(f'dist{lvl}', d[lvl].grad, ref_gdist[lvl])
    finally:
        xch.close()


def _exclusive_compute_mode() -> bool:
    try:
        import pynvml
        pynvml.nvmlInit()
        mode = pynvml.nvmlDeviceGetComputeMode(pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device()))
        return mode != pynvml.NVML_COMPUTEMODE_DEFAULT
    except Exception:
        return False


@pytest.mark.parametrize('cfg_name,V,world', [('tiny', 12, 2), ('tiny256', 13, 2), ('tiny', 9, 3)])
def test_view_sharded_product_path_two_processes(cuda_lib, cfg_name, V, world, tmp_path):
    """The real thing on one GPU: `world` PROCESSES (gloo group for the handle exchange, CUDA IPC mappings of each other's
    symmetric buffers) share cuda:0 by time slicing; every rank runs AdaptiveSparseHead.forward(view_shard=...) + backward +
    reduce_gradients on its views and compares with the unsharded path of the whole scene (tests/_peer_worker.py)."""
    if _exclusive_compute_mode():
        pytest.skip('GPU is in an exclusive compute mode: several processes cannot share it')
    import subprocess
    import sys
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    store = tmp_path / 'rdzv'
    procs = [subprocess.Popen([sys.executable, os.path.join(root, 'tests', '_peer_worker.py'), cfg_name, str(V), str(world), str(r),
                               str(store)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, cwd=root)
             for r in range(world)]
    outs = []
    for p_ in procs:
        try:
            out, _ = p_.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            p_.kill()
            out, _ = p_.communicate()
            out += '\n[timeout]'
        outs.append(out)
    for r, (p_, out) in enumerate(zip(procs, outs)):
        assert p_.returncode == 0 and 'PEER_WORKER_OK' in out, f'rank {r} failed:\n{out[-3000:]}'


def test_grad_averager_groups_and_graph_capture(cuda_lib):
    """GradAverager with a single rank (the average of one rank is the identity): every parameter group is reduced from its
    OnStream backward node on the communication stream, the projection layers' gradients in finish_step, and the whole step
    is capturable into a CUDA graph.  Gradients must equal those of a step without the averager."""
    from sgcdet_b200 import functional as SF
    cfg = syn.CONFIGS['tiny']
    sc = syn.make_scene(cfg, 8, shift_origin=True).to(DEV)
    sc.img_meta['sgc_projection'] = SF.compute_projection(sc.img_meta).to(DEV)   # static buffer: no H2D copy under capture
    head = plugin.build_voxel_head(cfg)
    head.load_state_dict(syn.make_state_dict(cfg))
    head = head.to(DEV).eval()
    params = list(head.parameters())
    feats = [f.clone().requires_grad_(True) for f in sc.mlvl_feats]

    def fwd_bwd():
        for p in params + feats:
            p.grad = None
        vol, valid, occ = head(feats, sc.img_meta, sc.mlvl_dpt_dists)
        ((vol * sc.grad_volume).sum() + head.occ_loss(occ, None, sc.geo_occ)['loss_occ']).backward()

    fwd_bwd()
    torch.cuda.synchronize()
    ref = [p.grad.clone() for p in params]
    avg = peer.GradAverager(params)
    try:
        def step():
            avg.begin_step()
            fwd_bwd()
            avg.finish_step()

        def check(what):
            torch.cuda.synchronize()
            avg.mem.check()
            for k, (p, r) in enumerate(zip(params, ref)):
                _close(f'{what}: param {k}', p.grad, r)

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()
            check('first step (everything reduced after the backward)')
            step()
            check('second step (one collective issued from the backward)')
            assert avg.copied_last_step == 0      # autograd adopted every placeholder the collective wrote into
        torch.cuda.current_stream().wait_stream(side)
        assert avg.groups_last_step == 2 * cfg.num_levels      # attention + FFN blocks (the tiny occupancy heads join the tail)
        graph = torch.cuda.CUDAGraph()
        for p in params + feats:
            p.grad = None
        with torch.cuda.graph(graph):
            step()
        graph.replay()
        check('graph replay')
        graph.replay()
        check('second replay')
        # without begin_step the path is untouched
        fwd_bwd()
        check('inactive averager')
    finally:
        avg.close()


@pytest.mark.parametrize('world', [1, 2, 3])
def test_peer_allreduce_tensor_list_in_place(cuda_lib, world):
    """sgc_peer_allreduce_tensors: a list of tensors of odd sizes (one of them a misaligned view) averaged in place by one
    launch per rank (gather, barrier, rank-ordered reduce, scatter)."""
    sizes = [5, 1024, 3, 98304 + 1, 256 * 384, 17]
    mems = peer.PeerMemory.simulate(world, 4 * (sum(sizes) + 4 * len(sizes)))
    try:
        g = torch.Generator().manual_seed(11)
        base = [[torch.randn(n + 1, generator=g).to(DEV) for n in sizes] for _ in range(world)]
        tens = [[b[1:] if i == 3 else b[:-1] for i, b in enumerate(bs)] for bs in base]     # tensor 3: 4-byte aligned only
        ref = []
        for i in range(len(sizes)):
            a = tens[0][i].clone()
            for r in range(1, world):
                a = a + tens[r][i]
            ref.append(a * 0.25)
        streams = [torch.cuda.Stream() for _ in range(world)]
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                mems[r].all_reduce_tensors([t for t in tens[r]], 0.25)
        torch.cuda.synchronize()
        for m in mems:
            m.check()
        for r in range(world):
            for t, a in zip(tens[r], ref):
                assert torch.equal(t, a)
    finally:
        for m in mems:
            m.close()
