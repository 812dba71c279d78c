"""The gather-style backward of the lift (csrc/sgc_lift_tiles.cu: taps binned by 4x8 pixel tile, value rows staged in shared
memory, every grad_vg row written once) against the scatter kernel of round 1 (csrc/sgc_lift.cu, REDs into a zero-filled
grad_vg), which is itself checked against the oracle and the reference's own kernels at module level.  Same inputs, every
output: grad of the value / folded maps, of the depth distribution and of the two biases."""
import pytest
import torch

from sgcdet_b200 import functional as SF
from sgcdet_b200 import plugin, synthetic as syn

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _level_inputs(cfg_name, V, level, seed=5):
    cfg = syn.CONFIGS[cfg_name]
    sc = syn.make_scene(cfg, V, seed=seed, shift_origin=True).to(DEV)
    head = plugin.build_voxel_head(cfg)
    head.load_state_dict(syn.make_state_dict(cfg))
    head = head.to(DEV).eval()
    nl = cfg.num_levels
    fi = nl - 1 - level
    h = sc.img_meta['img_shape'][0] // (4 * 2 ** (nl - 1 - level))
    w = sc.img_meta['img_shape'][1] // (4 * 2 ** (nl - 1 - level))
    dh = head.base_heads[level]
    with torch.no_grad():
        pre = dh.prepare(sc.mlvl_feats[fi], sc.mlvl_dpt_dists[fi], (h, w))
    g = torch.Generator().manual_seed(seed)
    N = dh.num_voxels
    k = cfg.topk_list[level - 1] if level > 0 else N
    sel = None if level == 0 else torch.sort(torch.randperm(N, generator=g)[:k]).values.to(DEV, torch.int32)
    proj = SF.compute_projection(sc.img_meta).to(DEV)
    pl = SF.project_compact(proj, dh.ref_3d, sel, sc.img_meta, cfg.dbound)
    return pre, pl, h, w, g


@pytest.mark.parametrize('cfg_name,V,level', [
    ('tiny', 9, 2), ('tiny256', 12, 1),
    ('SGCDet_ScanNet', 40, 2),             # 59 x 80: the last tile row has 3 of 4 pixel rows
    ('SGCDet_ScanNet', 40, 0),             # 14 x 20: partial tiles in both directions
    ('SGCDet_ARKit', 12, 1),               # 30 x 40
    ('SGCDet_large_ScanNet200', 6, 2),     # C = 128 (16-wide heads), 51 200 voxels
])
def test_tile_backward_matches_scatter_backward(cuda_lib, monkeypatch, cfg_name, V, level):
    pre, pl, h, w, g = _level_inputs(cfg_name, V, level)
    n = int(pl.view_offsets[-1])
    assert n > 0
    outs = {}
    gs = None
    for tiles in (False, True):
        monkeypatch.setattr(SF, 'LIFT_TILES', tiles)
        leaves = [pre[k].detach().clone().requires_grad_(True) for k in ('vg', 'dist', 'vbias', 'gbias')]
        slots, samp = SF.Lift.apply(*leaves, pl, h, w)
        if gs is None:
            gs = torch.randn(slots.shape, generator=g).to(DEV)
            gs[n:] = 0      # rows beyond the pair count are never read by either kernel
        slots.backward(gs)
        torch.cuda.synchronize()
        outs[tiles] = [t.grad.clone() for t in leaves]
    for name, a, b in zip(('grad_vg', 'grad_dist', 'grad_vbias', 'grad_gbias'), outs[False], outs[True]):
        assert torch.isfinite(b).all(), name
        scale = a.abs().max().item() + 1e-20
        err = (a - b).abs().max().item() / scale
        assert err < 2e-5, f'{name}: {err}'       # same fp32 terms, different summation order


def test_tile_backward_overwrites_every_row(cuda_lib, monkeypatch):
    """The tile kernel needs no zero fill: poison the output allocation's memory first (via the caching allocator) and check
    that pixels no tap touches come back as exact zeros."""
    monkeypatch.setattr(SF, 'LIFT_TILES', True)
    pre, pl, h, w, g = _level_inputs('tiny', 3, 2, seed=9)
    leaves = [pre[k].detach().clone().requires_grad_(True) for k in ('vg', 'dist', 'vbias', 'gbias')]
    poison = torch.full_like(leaves[0], float('nan'))
    del poison                                   # back to the allocator: the next empty_like of this size reuses it
    slots, samp = SF.Lift.apply(*leaves, pl, h, w)
    n = int(pl.view_offsets[-1])
    gs = torch.zeros_like(slots)
    gs[:n] = torch.randn(n, slots.shape[1], generator=g).to(DEV)
    slots.backward(gs)
    gv = leaves[0].grad
    assert torch.isfinite(gv).all()
    assert (gv == 0).any()                       # untouched pixels exist in a 3-view scene and are exact zeros
