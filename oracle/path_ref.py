"""CPU restatement of the SGCDet view-transform modules (test infrastructure only).

Functional (no nn.Module state): every function takes the weights as a flat dict keyed by the
reference's state-dict names (prefix ``voxel_head.`` stripped), so the same dict drives the oracle
and the product.  Third-party arithmetic on the path (``nn.MultiheadAttention``, ``LayerNorm``,
``F.interpolate``, ``softmax``, ``BCELoss``) is done with the identical torch functions the
reference calls, on the CPU.

Reference files restated here:
  * ``models/im2voxel/transformer_utils/encoder.py:168-223``  (projection + visibility mask)
  * ``.../deformable_cross_attention.py:67-116``              (Grid_Sample_3D_Feature)
  * ``.../deformable_cross_attention.py:364-496``             (MSDeformableAttention3D_DFA3D.forward)
  * ``.../deformable_cross_attention.py:705-837``             (DeformCrossAttention_DFA3D.forward)
  * ``.../encoder.py:262-340`` + ``custom_base_transformer_layer.py:72-156`` (cross_attn, norm, ffn, norm)
  * ``.../transformer.py:118-185``                            (get_vox_features: flatten to channel-last)
  * ``models/im2voxel/DenseHead.py:32-84``                    (voxel centres, compaction, scatter)
  * ``models/im2voxel/AdaptiveSparseHead.py:9-13,43-103``     (coarse-to-fine loop, top-k, occ loss)

Fixed contracts the oracle DEFINES (the reference leaves them unspecified):
  * projection: explicit fp32 operation order, one rounding per operation, no FMA
    (``((P0*x + P1*y) + P2*z) + P3``); the reference uses a batched ``torch.matmul`` (encoder.py:203)
    whose accumulation order is unspecified.  ``point_sampling_matmul`` keeps the literal form for a
    tolerance cross-check.
  * top-k: value descending, ties broken by ascending index (``torch.topk`` tie order is unspecified,
    AdaptiveSparseHead.py:10).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from . import dfa3d_ref

EPS = 1e-5


# ---------------------------------------------------------------- projection (encoder.py:168-223)

def compute_projection(img_meta: dict, stride: int = 1) -> torch.Tensor:
    """encoder.py:168-177 -- on the CPU, in fp32, with the same torch ops."""
    intrinsic = torch.tensor(np.asarray(img_meta['lidar2img']['intrinsic'])[:3, :3])
    ratio = img_meta['ori_shape'][0] / (img_meta['img_shape'][0] / stride)
    intrinsic[:2] /= ratio
    proj = [intrinsic @ torch.tensor(np.asarray(e))[:3] for e in img_meta['lidar2img']['extrinsic']]
    return torch.stack(proj)  # [V,3,4]


def point_sampling(ref_3d: torch.Tensor, img_meta: dict, dbound: Sequence[float]):
    """ref_3d [Q,3] fp32 -> (ref_cam [V,Q,3] = (u,v,d), mask [V,Q] bool); pinned op order."""
    assert ref_3d.dtype == torch.float32
    ogfH, ogfW = img_meta['img_shape'][0], img_meta['img_shape'][1]
    origin = torch.tensor(np.asarray(img_meta['lidar2img']['origin'], dtype=np.float32))
    p = ref_3d + origin  # encoder.py:187-188
    P = compute_projection(img_meta, 1)  # [V,3,4]
    px, py, pz = p[:, 0][None], p[:, 1][None], p[:, 2][None]  # [1,Q]

    def row(i):
        a = P[:, i, 0:1] * px
        a = a + P[:, i, 1:2] * py
        a = a + P[:, i, 2:3] * pz
        return a + P[:, i, 3:4]

    x, y, z = row(0), row(1), row(2)  # [V,Q]
    zc = torch.maximum(z, torch.full_like(z, EPS))  # encoder.py:206
    u = (x / zc) / float(ogfW)  # encoder.py:209
    v = (y / zc) / float(ogfH)  # encoder.py:210
    d = (z - float(dbound[0])) / float(dbound[1] - dbound[0])  # encoder.py:211
    mask = (z > EPS) & (u > EPS) & (u < (1.0 - EPS)) & (v > EPS) & (v < (1.0 - EPS))  # :213-219
    return torch.stack([u, v, d], dim=-1), mask


def point_sampling_matmul(ref_3d: torch.Tensor, img_meta: dict, dbound: Sequence[float]):
    """The literal tensor program of encoder.py:179-223 (torch.matmul); tolerance check only."""
    ogfH, ogfW = img_meta['img_shape'][0], img_meta['img_shape'][1]
    origin = torch.tensor(np.asarray(img_meta['lidar2img']['origin'], dtype=np.float32))
    reference_points = ref_3d.view(1, 1, -1, 3) + origin
    reference_points = reference_points.permute(1, 0, 2, 3)
    D, B, num_query = reference_points.size()[:3]
    projection = compute_projection(img_meta, 1)
    num_cam = projection.shape[0]
    reference_points = reference_points.view(D, B, 1, num_query, 3).repeat(1, 1, num_cam, 1, 1)
    reference_points = torch.cat(
        (reference_points, torch.ones(*reference_points.shape[:-1], 1).type_as(reference_points)), dim=-1)
    projection = projection.unsqueeze(0).unsqueeze(0)
    cam = torch.matmul(projection, reference_points.permute(0, 1, 2, 4, 3)).permute(0, 1, 2, 4, 3)
    points_d = cam[..., 2:3]
    cam[..., 0:2] = cam[..., 0:2] / torch.maximum(points_d, torch.ones_like(cam[..., 2:3]) * EPS)
    cam[..., 0] /= ogfW
    cam[..., 1] /= ogfH
    cam[..., 2] = (cam[..., 2] - dbound[0]) / (dbound[1] - dbound[0])
    m = (points_d > EPS)
    m = (m & (cam[..., 0:1] > EPS) & (cam[..., 0:1] < (1.0 - EPS))
         & (cam[..., 1:2] > EPS) & (cam[..., 1:2] < (1.0 - EPS)))
    return cam[0, 0], m[0, 0, :, :, 0]  # [V,Q,3], [V,Q]


# ---------------------------------------------------------------- top-k (AdaptiveSparseHead.py:9-13)

def topk_mask(occ: torch.Tensor, k: int) -> torch.Tensor:
    """occ [N] -> {0,1} float mask of the k largest; ties -> lower index first."""
    order = torch.sort(occ, descending=True, stable=True).indices[:k]
    mask = torch.zeros_like(occ)
    mask[order] = 1.0
    return mask


# ---------------------------------------------------------------- one DenseHead level

def _p(sd: Dict[str, torch.Tensor], level: int, name: str) -> torch.Tensor:
    return sd[f'base_heads.{level}.cross_transformer.encoder.layers.0.{name}']


def dense_head_forward(sd, level: int, feat, dpt_dist, img_meta, proposal, cfg, *,
                       training: bool = False, return_intermediates: bool = False):
    """DenseHead.forward (DenseHead.py:50-84) for one level.

    feat [1,V,C,h,w], dpt_dist [1,V,D,h,w] (already cropped), proposal [N] float mask or None.
    Returns the dense volume [1,C,X,Y,Z] (and optionally a dict of intermediates).
    """
    assert feat.shape[0] == 1
    _, V, C, h, w = feat.shape
    Dd = dpt_dist.shape[2]
    M, Pn = cfg.num_heads, cfg.num_points
    Cm = C // M
    n_vox = cfg.n_voxels_list[level]
    N = int(np.prod(n_vox))
    ref_3d_all = sd[f'base_heads.{level}.ref_3d']
    if proposal is None:
        proposal = torch.ones(N)
    sel = torch.nonzero(proposal > 0).view(-1)  # DenseHead.py:66 (ascending)
    Q = sel.numel()
    ref_3d = ref_3d_all[sel]

    # transformer.py:151-170: channel-last flatten
    value_raw = feat[0].flatten(2).permute(0, 2, 1).contiguous()  # [V,S,C]
    dist = dpt_dist[0].flatten(2).permute(0, 2, 1).contiguous()  # [V,S,D]
    S = h * w
    shapes3d = torch.tensor([[h, w, Dd]], dtype=torch.long)
    lsi = torch.zeros(1, dtype=torch.long)

    ref_cam, mask = point_sampling(ref_3d, img_meta, cfg.dbound)  # [V,Q,3], [V,Q]

    # DCA:758-773 -- per-view visible lists; the zero padding to max_len only produces rows that are
    # discarded at DCA:815-818, so the restatement works on the ragged lists directly.
    da = 'attentions.0.deformable_attention.'
    slots = torch.zeros(V, Q, C, dtype=feat.dtype)
    inter = dict(ref_cam=ref_cam, mask=mask, sel=sel, pairs=[])
    for v in range(V):
        idx = torch.nonzero(mask[v]).view(-1)
        if idx.numel() == 0:
            continue
        n = idx.numel()
        rp = ref_cam[v, idx]  # [n,3]
        # Grid_Sample_3D_Feature (DCA:67-116): M=1, L=1, P=1, weights = 1, on the RAW features
        loc1 = rp.view(1, n, 1, 1, 1, 3)
        q_img, _ = dfa3d_ref.dfa3d_forward(value_raw[v].view(1, S, 1, C), dist[v].view(1, S, 1, Dd),
                                           shapes3d, lsi, loc1, torch.ones(1, n, 1, 1, 1, dtype=feat.dtype))
        q_img = q_img.view(n, C)
        # MSDeformableAttention3D_DFA3D.forward (DCA:417-489)
        val = F.linear(value_raw[v], _p(sd, level, da + 'value_proj.weight'), _p(sd, level, da + 'value_proj.bias'))
        val = val.view(1, S, M, Cm)
        dist_m = dist[v].view(1, S, 1, Dd).repeat(1, 1, M, 1)
        off_uv = F.linear(q_img, _p(sd, level, da + 'sampling_offsets.weight'),
                          _p(sd, level, da + 'sampling_offsets.bias')).view(1, n, M, 1, Pn, 2)
        off_d = F.linear(q_img, _p(sd, level, da + 'sampling_offsets_depth.weight'),
                         _p(sd, level, da + 'sampling_offsets_depth.bias')).view(1, n, M, 1, Pn, 1)
        off = torch.cat([off_uv, off_d], dim=-1)
        aw = F.linear(q_img, _p(sd, level, da + 'attention_weights.weight'),
                      _p(sd, level, da + 'attention_weights.bias')).view(1, n, M, Pn)
        aw = aw.softmax(-1).view(1, n, M, 1, Pn)
        normalizer = torch.tensor([w, h, Dd], dtype=feat.dtype)  # DCA:445-446 (W,H,D)
        loc = rp.view(1, n, 1, 1, 1, 3) + off / normalizer
        out, ds = dfa3d_ref.dfa3d_forward(val, dist_m, shapes3d, lsi, loc, aw)
        slots[v, idx] = out.view(n, C)  # DCA:815-818
        if return_intermediates:
            inter['pairs'].append(dict(view=v, idx=idx, q_img=q_img, loc=loc.view(n, M, Pn, 3),
                                       attn=aw.view(n, M, Pn), out=out.view(n, C)))

    # DCA:819-837
    count = mask.sum(0)  # [Q]
    valid_index = torch.nonzero(count).view(-1)
    output = torch.zeros(Q, C, dtype=feat.dtype)
    if valid_index.numel() > 0:
        valid_slots = slots[:, valid_index]  # [V,L,C]
        valid_mask = mask[:, valid_index]  # [V,L]
        slots_mean = (valid_slots * valid_mask.unsqueeze(-1)).sum(0) / count[valid_index].unsqueeze(-1)
        slots_mean = F.linear(slots_mean, _p(sd, level, 'attentions.0.output_proj.weight'),
                              _p(sd, level, 'attentions.0.output_proj.bias'))
        pooled, _ = F.multi_head_attention_forward(
            slots_mean.unsqueeze(0), valid_slots, valid_slots, C, 8,
            _p(sd, level, 'attentions.0.attention_pooling.in_proj_weight'),
            _p(sd, level, 'attentions.0.attention_pooling.in_proj_bias'),
            None, None, False, 0.0,
            _p(sd, level, 'attentions.0.attention_pooling.out_proj.weight'),
            _p(sd, level, 'attentions.0.attention_pooling.out_proj.bias'),
            training=training, key_padding_mask=~valid_mask.transpose(0, 1), need_weights=False)
        output[valid_index] = pooled[0]
    # + inp_residual (zero queries), dropout p=0   (DCA:837, config dropout=0)
    inter['attn_out'] = output
    # encoder.py:310-338: norm, ffn (mmcv FFN: x + W2 relu(W1 x + b1) + b2; dropout in train mode), norm
    x = F.layer_norm(output, (C,), _p(sd, level, 'norms.0.weight'), _p(sd, level, 'norms.0.bias'))
    hdn = F.relu(F.linear(x, _p(sd, level, 'ffns.0.layers.0.0.weight'), _p(sd, level, 'ffns.0.layers.0.0.bias')))
    x = x + F.linear(hdn, _p(sd, level, 'ffns.0.layers.1.weight'), _p(sd, level, 'ffns.0.layers.1.bias'))
    x = F.layer_norm(x, (C,), _p(sd, level, 'norms.1.weight'), _p(sd, level, 'norms.1.bias'))
    inter['seed_feats'] = x
    # DenseHead.py:80-83
    vol = torch.zeros(N, C, dtype=feat.dtype)
    vol[sel] = x
    vol = vol.reshape(*n_vox, C).permute(3, 0, 1, 2).unsqueeze(0)
    if return_intermediates:
        return vol, inter
    return vol


def adaptive_sparse_head_forward(sd, mlvl_feats, img_meta, mlvl_dpt_dists, cfg, *,
                                 forced_proposals: Optional[List[Optional[torch.Tensor]]] = None,
                                 return_intermediates: bool = False):
    """AdaptiveSparseHead.forward (AdaptiveSparseHead.py:43-93) -> (volume, valid, occ_preds).

    ``forced_proposals[i]`` (if given) replaces the top-k mask of level i (teacher forcing used by the
    parity tests: occupancy is a float that only matches to 1e-3, the selection is integer-exact).
    """
    nl = cfg.num_levels
    volumes: List[Optional[torch.Tensor]] = [None] * nl
    occ_list = []
    masks: List[Optional[torch.Tensor]] = [None] * nl
    inters = []
    for i in range(nl):
        ds = 4 * 2 ** (nl - 1 - i)
        height = img_meta['img_shape'][0] // ds
        width = img_meta['img_shape'][1] // ds
        fi = nl - 1 - i
        feat = mlvl_feats[fi][:, :, :, :height, :width]
        dist = mlvl_dpt_dists[fi][:, :, :, :height, :width]
        if i == 0:
            r = dense_head_forward(sd, i, feat, dist, img_meta, None, cfg,
                                   return_intermediates=return_intermediates)
            volumes[i], it = r if return_intermediates else (r, None)
        else:
            up = F.interpolate(volumes[i - 1], scale_factor=2, mode='trilinear', align_corners=False)
            occ = torch.sigmoid(F.linear(up.permute(0, 2, 3, 4, 1), sd[f'occ_pred_heads.{i - 1}.0.weight'],
                                         sd[f'occ_pred_heads.{i - 1}.0.bias'])).reshape(1, -1)
            occ_list.append(occ)
            if forced_proposals is not None and forced_proposals[i] is not None:
                masks[i] = forced_proposals[i]
            else:
                masks[i] = topk_mask(occ[0].detach(), cfg.topk_list[i - 1])
            r = dense_head_forward(sd, i, feat, dist, img_meta, masks[i], cfg,
                                   return_intermediates=return_intermediates)
            dv, it = r if return_intermediates else (r, None)
            volumes[i] = up + dv
        inters.append(it)
    volume_out = volumes[-1]
    occ_preds = torch.cat(occ_list[::-1], dim=1)
    X, Y, Z = cfg.n_voxels_list[-1]
    valid = masks[-1].view(X, Y, Z).bool().long().unsqueeze(0).unsqueeze(0)
    if return_intermediates:
        return volume_out, valid, occ_preds, dict(volumes=volumes, masks=masks, levels=inters)
    return volume_out, valid, occ_preds


def occ_loss(occ_pred: torch.Tensor, geo_occ_gt: torch.Tensor) -> torch.Tensor:
    """AdaptiveSparseHead.py:100-103."""
    N = occ_pred.shape[1]
    return F.binary_cross_entropy(occ_pred, geo_occ_gt[:, 0:N].to(occ_pred.dtype)).mean() * 0.5


def path_loss(sd, scene, *, forced_proposals=None):
    """Scalar used for gradient parity: sum(volume * G) + occ_loss."""
    vol, valid, occ = adaptive_sparse_head_forward(sd, scene.mlvl_feats, scene.img_meta,
                                                   scene.mlvl_dpt_dists, scene.cfg,
                                                   forced_proposals=forced_proposals)
    return (vol * scene.grad_volume).sum() + occ_loss(occ, scene.geo_occ), (vol, valid, occ)
