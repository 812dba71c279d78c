"""One rank of the multi-process view-sharding tests (argv: config, V, world, rank, store[, mode]).

mode ``unsharded`` (default; tests/test_gpu_peer.py::test_view_sharded_product_path_two_processes): forward + backward +
reduce_gradients against the unsharded product path of the whole scene.
mode ``refkernels`` (tests/test_gpu_full_shape_parity.py::test_view_sharded_large_arkit_forward_matches_reference_kernels):
forward against the reference's own DFA3D kernels under the restated glue (oracle/gpu_ref.py) at the north-star tolerance."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from sgcdet_b200 import parallel, plugin, synthetic as syn  # noqa: E402

RTOL, ATOL = 1e-3, 1e-4


def close(name, a, b):
    scale = b.abs().max().item() + 1e-12
    rel = ((a - b).norm() / (b.norm() + 1e-12)).item()
    assert rel < 5e-3, f'{name}: relative Frobenius error {rel}'
    assert ((a - b).abs().max() / scale).item() < 2e-2, name


def against_reference_kernels(cfg, V, world, rank, dev):
    from oracle import gpu_ref
    torch.backends.cuda.matmul.allow_tf32 = False
    gpu_ref.PINNED_PROJECTION = True
    sc = syn.make_scene(cfg, V, shift_origin=True).to(dev)
    sd = syn.make_state_dict(cfg)
    sdg = {k: v.to(dev) for k, v in sd.items()}
    with torch.no_grad():     # every rank evaluates the (deterministic) reference itself
        vol_r, valid_r, occ_r, masks = gpu_ref.head_forward_gpu(sdg, sc.mlvl_feats, sc.img_meta, sc.mlvl_dpt_dists, cfg,
                                                                training=False, return_masks=True)
    forced = [None] + [torch.nonzero(masks[i].view(-1) > 0).view(-1).to(dev, torch.int32) for i in range(1, cfg.num_levels)]
    head = plugin.build_voxel_head(cfg)
    head.load_state_dict(sd, strict=True)
    head = head.to(dev).eval()
    torch.cuda.synchronize()
    xch = parallel.ViewShardExchange(head, device=dev)
    views = parallel.shard_views(V, world, rank)
    f, m, d = parallel.shard_scene_inputs(sc.mlvl_feats, sc.img_meta, sc.mlvl_dpt_dists, views)
    with torch.no_grad():
        vol, valid, occ = head(f, m, d, forced_selection=forced, view_shard=xch)
    torch.cuda.synchronize()
    xch.mem.check()
    assert torch.equal(valid, valid_r)
    torch.testing.assert_close(occ, occ_r, rtol=RTOL, atol=ATOL)
    # rtol 1e-3 / atol 1e-4 element by element, except for at most 2 in a million elements within atol 1e-3 (see
    # tests/test_gpu_full_shape_parity.py::_close_but_for_outliers)
    bad = ((vol - vol_r).abs() > (ATOL + RTOL * vol_r.abs()))
    assert int(bad.sum()) <= 2e-6 * bad.numel(), f'{int(bad.sum())} of {bad.numel()} volume elements outside the tolerance'
    torch.testing.assert_close(vol, vol_r, rtol=RTOL, atol=1e-3)
    xch.close()


def main():
    cfg_name, V, world, rank, store = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    mode = sys.argv[6] if len(sys.argv) > 6 else 'unsharded'
    torch.cuda.set_device(0)
    dev = torch.device('cuda', 0)
    dist.init_process_group('gloo', init_method=f'file://{store}', rank=rank, world_size=world)
    cfg = syn.CONFIGS[cfg_name]
    if mode == 'refkernels':
        against_reference_kernels(cfg, V, world, rank, dev)
        dist.destroy_process_group()
        print('PEER_WORKER_OK', rank)
        return
    sc = syn.make_scene(cfg, V, shift_origin=True).to(dev)
    head = plugin.build_voxel_head(cfg)
    head.load_state_dict(syn.make_state_dict(cfg))
    head = head.to(dev).eval()
    # the unsharded reference of the whole scene (every rank computes its own copy, before any exchange is in flight)
    feats = [f.clone().requires_grad_(True) for f in sc.mlvl_feats]
    dists = [d.clone().requires_grad_(True) for d in sc.mlvl_dpt_dists]
    vol, valid, occ, its = head(feats, sc.img_meta, dists, return_intermediates=True)
    ((vol * sc.grad_volume).sum() + head.occ_loss(occ, None, sc.geo_occ)['loss_occ']).backward()
    ref_grads = {k: p.grad.clone() for k, p in head.named_parameters()}
    forced = [None] + [it['sel'] for it in its[1:]]
    head.zero_grad(set_to_none=True)
    torch.cuda.synchronize()

    xch = parallel.ViewShardExchange(head, device=dev)     # gloo exchanges the IPC handles; barrier inside
    views = parallel.shard_views(V, world, rank)
    f, m, d = parallel.shard_scene_inputs(sc.mlvl_feats, sc.img_meta, sc.mlvl_dpt_dists, views)
    f = [t.requires_grad_(True) for t in f]
    d = [t.requires_grad_(True) for t in d]
    vol_s, valid_s, occ_s = head(f, m, d, forced_selection=forced, view_shard=xch)
    ((vol_s * sc.grad_volume).sum() + head.occ_loss(occ_s, None, sc.geo_occ)['loss_occ']).backward()
    xch.reduce_gradients(head)
    torch.cuda.synchronize()
    xch.mem.check()
    assert torch.equal(valid_s, valid)
    torch.testing.assert_close(occ_s, occ, rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(vol_s.detach(), vol.detach(), rtol=RTOL, atol=ATOL)
    for k, p in head.named_parameters():
        close(k, p.grad, ref_grads[k])
    for lvl in range(cfg.num_levels):
        close(f'feat{lvl}', f[lvl].grad, feats[lvl].grad[:, views.start:views.stop])
        close(f'dist{lvl}', d[lvl].grad, dists[lvl].grad[:, views.start:views.stop])
    # replicated chain: bit-identical volumes on the ranks
    digest = torch.stack([vol_s.detach().double().sum(), occ_s.detach().double().sum()]).cpu()
    gathered = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(gathered, digest)
    assert all(torch.equal(g, gathered[0]) for g in gathered), gathered
    xch.close()
    dist.destroy_process_group()
    print('PEER_WORKER_OK', rank)


if __name__ == '__main__':
    main()
