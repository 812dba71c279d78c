"""Kernel-level breakdown of one fwd+bwd step with torch.profiler (CUPTI): python tools/profile_step.py [config] [views]."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from sgcdet_b200 import plugin, synthetic as syn  # noqa: E402

cfg = syn.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else 'SGCDet_ScanNet']
V = int(sys.argv[2]) if len(sys.argv) > 2 else 40
dev = 'cuda'
sc = syn.make_scene(cfg, V, shift_origin=True).to(dev)
head = plugin.build_voxel_head(cfg)
head.load_state_dict(syn.make_state_dict(cfg))
head = head.to(dev).train()
feats = [f.clone().requires_grad_(True) for f in sc.mlvl_feats[:3]]
dists = [d.clone().requires_grad_(True) for d in sc.mlvl_dpt_dists[:3]]
gvol = sc.grad_volume.permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3)


avg = None
if os.environ.get('SGC_PROFILE_AVERAGER'):
    # the gradient averager of scene-batch DP with a single rank (no process group): same launches, stream structure and SM
    # footprint as on N ranks, without the link latency
    from sgcdet_b200 import peer
    avg = peer.GradAverager(list(head.parameters()))


loss_stream = torch.cuda.Stream()


def step():
    for t in list(head.parameters()) + feats + dists:
        t.grad = None
    if avg is not None:
        avg.begin_step()
    vol, valid, occ = head(feats, sc.img_meta, dists)
    # as bench.py: the backward is seeded with G (the gradient of sum(volume*G)); the occupancy loss (value and gradient) lives
    # on the loss stream, off the path from the volume to its gradient
    torch.autograd.backward([vol, head.occ_loss(occ, None, sc.geo_occ, stream=loss_stream)['loss_occ']], [gvol, None])
    torch.cuda.current_stream().wait_stream(loss_stream)
    if avg is not None:
        avg.finish_step()


for _ in range(3):
    step()
torch.cuda.synchronize()
N = 3
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(N):
        step()
    torch.cuda.synchronize()
evs = [e for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == 'CUDA']
tot = sum(e.device_time_total for e in evs)
print(f'GPU kernel time per step: {tot / N / 1e3:.3f} ms in {sum(e.count for e in evs) / N:.0f} launches')
for e in sorted(evs, key=lambda e: -e.device_time_total)[:45]:
    print(f'{e.device_time_total / N / 1e3:8.3f} ms {e.count / N:6.1f}x  {e.key[:110]}')

if os.environ.get('SGC_TRACE'):
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof2:
        step()
        torch.cuda.synchronize()
    prof2.export_chrome_trace(os.environ['SGC_TRACE'])

if os.environ.get('SGC_GRAPH_TRACE'):
    from sgcdet_b200 import functional as SF
    sc.img_meta['sgc_projection'] = SF.compute_projection(sc.img_meta).to(dev)
    # timeline of ONE CUDA-graph replay (true stream concurrency, no CPU launch gaps): feed to tools/graph_timeline.py
    for t in list(head.parameters()) + feats + dists:
        t.grad = None
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    for t in list(head.parameters()) + feats + dists:
        t.grad = None
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof3:
        g.replay()
        torch.cuda.synchronize()
    prof3.export_chrome_trace(os.environ['SGC_GRAPH_TRACE'])
