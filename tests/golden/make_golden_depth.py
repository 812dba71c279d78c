"""Generates tests/golden/depth_producer.pt by executing the reference's OWN functions of
/root/reference/mmdet3d_plugin/models/im2voxel/depth_utils/depth_est_fusion.py on the CPU.

    python tests/golden/make_golden_depth.py          # in the build container (needs /root/reference)

The module itself cannot be imported here (mmdet / mmcv / the ResNetFPN extractor are not installed), so the three
function DEFINITIONS this path needs -- get_closest_frame_ids, collect_proj, homo_warping -- are cut out of the
reference's source with ``ast`` at run time and executed unmodified in a namespace holding torch / F / np (nothing is
copied into this repository).  The neighbour loop of DepthNet_Fusion.forward (:218-232, three statements inside a method
that needs the whole network) is applied to the reference functions' outputs below."""
import ast
import math
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SRC = '/root/reference/mmdet3d_plugin/models/im2voxel/depth_utils/depth_est_fusion.py'
WANT = ('get_closest_frame_ids', 'collect_proj', 'homo_warping')


def reference_functions():
    tree = ast.parse(open(SRC).read())
    ns = {'torch': torch, 'F': F, 'np': np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in WANT:
            exec(compile(ast.Module([node], []), SRC, 'exec'), ns)
    return [ns[n] for n in WANT]


def smooth_extrinsics(V, seed):
    """world->cam poses along a short arc looking at a common point: neighbouring frames overlap, as in a video."""
    g = np.random.default_rng(seed)
    out = []
    for i in range(V):
        ang = 0.12 * i + 0.02 * g.standard_normal()
        pos = np.array([2.2 * math.cos(ang), 2.2 * math.sin(ang), 1.3 + 0.05 * g.standard_normal()])
        f = np.array([0.2, 0.1, 0.6]) - pos
        f /= np.linalg.norm(f)
        r = np.cross(f, np.array([0.0, 0.0, 1.0]))
        r /= np.linalg.norm(r)
        d = np.cross(f, r)
        R = np.stack([r, d, f])
        E = np.eye(4)
        E[:3, :3] = R
        E[:3, 3] = -R @ pos
        out.append(E.astype(np.float32))
    return np.stack(out)


def main():
    get_ids, collect_proj, homo_warping = reference_functions()
    cases = {}
    for name, V, C, H, W, k, seed in (('a', 6, 16, 12, 16, 2, 11), ('b', 4, 8, 10, 14, 2, 12), ('c', 7, 20, 9, 12, 4, 13)):
        g = torch.Generator().manual_seed(seed)
        f = torch.randn(V, C, H, W, generator=g)
        w2c = torch.tensor(smooth_extrinsics(V, seed))
        K = torch.eye(4)
        K[0, 0], K[1, 1], K[0, 2], K[1, 2] = 0.9 * W, 0.9 * W, W / 2 - 0.3, H / 2 + 0.2      # feature-map intrinsic
        depth = torch.tensor(np.arange(0.4, 4.0, 0.6, dtype=np.float32) + 0.3)
        kk = min(k, V - 1)
        ids = get_ids(V, kk)
        ref_proj, nei_projs = collect_proj(w2c, K, ids)
        nei_feats = torch.unbind(f[ids.view(-1)].view(V, kk, C, H, W), dim=1)
        dv = depth.unsqueeze(0).repeat(V, 1)
        corr = torch.zeros(V, depth.numel(), H, W)
        warped0 = None
        for nf, npj in zip(nei_feats, nei_projs):               # depth_est_fusion.py:222-231
            warped = homo_warping(nf, npj, ref_proj, dv)
            warped0 = warped if warped0 is None else warped0
            corr = corr + (warped * f.unsqueeze(2)).sum(dim=1) / torch.sqrt(torch.tensor(C).float())
        corr = corr / kk
        frac_in = float((warped0.abs().sum(1) > 0).float().mean())
        print(name, 'correlation', tuple(corr.shape), 'samples inside the neighbour image:', round(frac_in, 3))
        cases[name] = dict(f_mvs=f, w2c=w2c, intr_feat=K, depth=depth, k=k, ids=ids, warped0=warped0[:, ::4].contiguous(),
                           correlation=corr)
    torch.save(cases, os.path.join(ROOT, 'tests', 'golden', 'depth_producer.pt'))


if __name__ == '__main__':
    sys.path.insert(0, ROOT)
    main()
