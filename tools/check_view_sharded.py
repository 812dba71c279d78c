"""Real multi-GPU check of view sharding (NCCL): torchrun --nproc-per-node G tools/check_view_sharded.py [config] [views]
Every rank builds the same synthetic scene, keeps its slice of the views, runs forward+backward through
parallel.forward_view_sharded; rank 0 also runs the unsharded path and prints the differences + timings."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from sgcdet_b200 import parallel, plugin, synthetic as syn  # noqa: E402

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
cfg = syn.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else 'SGCDet_large_ARKit']
V = int(sys.argv[2]) if len(sys.argv) > 2 else 40
sc = syn.make_scene(cfg, V, shift_origin=True).to(dev)
head = plugin.build_voxel_head(cfg)
head.load_state_dict(syn.make_state_dict(cfg))
head = head.to(dev).eval()  # replicated compute must be identical on every rank (no per-rank dropout noise)
views = parallel.shard_views(V, world, rank)
f, m, d = parallel.shard_scene_inputs(sc.mlvl_feats, sc.img_meta, sc.mlvl_dpt_dists, views)
f = [t.requires_grad_(True) for t in f[:3]]
gvol = sc.grad_volume.permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3)


def step():
    for p in list(head.parameters()) + f:
        p.grad = None
    vol, valid, occ = parallel.forward_view_sharded(head, [(f, m, d)])
    loss = (vol * gvol).sum() + head.occ_loss(occ, None, sc.geo_occ)['loss_occ']
    loss.backward()
    parallel.allreduce_view_sharded_gradients(head)
    return vol, valid, occ, loss


for _ in range(3):
    vol, valid, occ, loss = step()
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K = 5
for _ in range(K):
    vol, valid, occ, loss = step()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / K], device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
# all ranks must hold the same volume / selection
chk = torch.stack([vol.detach().double().sum(), valid.double().sum(), occ.detach().double().sum()])
lo, hi = chk.clone(), chk.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN)
dist.all_reduce(hi, op=dist.ReduceOp.MAX)
if rank == 0:
    grads_s = {k: p.grad.clone() for k, p in head.named_parameters()}
    feats = [t.clone().requires_grad_(True) for t in sc.mlvl_feats[:3]]
    for p in head.parameters():
        p.grad = None
    vol_r, valid_r, occ_r = head(feats, sc.img_meta, sc.mlvl_dpt_dists)
    ((vol_r * gvol).sum() + head.occ_loss(occ_r, None, sc.geo_occ)['loss_occ']).backward()
    worst = 0.0
    for k, p in head.named_parameters():
        s = p.grad.abs().max().item() + 1e-12
        worst = max(worst, ((grads_s[k] - p.grad).abs().max() / s).item())
    gf = max(((f[i].grad - feats[i].grad[:, views.start:views.stop]).abs().max() / (feats[i].grad.abs().max() + 1e-12)).item()
             for i in range(3))
    print(json.dumps(dict(config=cfg.name, views=V, world=world, ms_per_step_eager=round(ms.item(), 3),
                          replicas_identical=bool(torch.allclose(lo, hi, rtol=1e-6)),
                          selection_overlap=float((valid & valid_r).sum() / valid_r.sum()),
                          volume_max_abs_diff=float((vol - vol_r).abs().max()), occ_max_abs_diff=float((occ - occ_r).abs().max()),
                          param_grad_max_rel_diff=worst, feat_grad_max_rel_diff=gf)))
dist.destroy_process_group()
