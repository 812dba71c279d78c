"""GPU-side reference arm (test / bench infrastructure only; never imported by sgcdet_b200/).

"The reference on the same B200": the reference's OWN DFA3D CUDA kernels, compiled unmodified for sm_100a into
oracle/_ref by oracle/build_ref.py, driven by a restatement of the reference's Python glue in its original
structure -- per-view ``nonzero`` lists (host syncs), zero-padded rebatch to ``max_len``, two batched DFA3D calls
with B = V, fp32 ``value_proj`` over every pixel, the 8x ``repeat`` of the depth distribution, Python loops
scattering into ``slots``, ``nn.MultiheadAttention`` over views, LayerNorm / FFN, ``F.interpolate`` and ``torch.topk``
(deformable_cross_attention.py:67-116, 364-496, 705-837; encoder.py:179-223, 262-340; DenseHead.py:50-84;
AdaptiveSparseHead.py:43-93; multi_scale_3ddeformable_attn_function.py:275-351).  The plugin itself cannot be imported
on the GPU box (no mmcv, no /root/reference), hence the restatement; the kernels are the reference's.
"""
from __future__ import annotations

import json
import os
import time

import numpy as np
import torch
import torch.nn.functional as F

from . import build_ref, path_ref

_EXT = None


def ext():
    global _EXT
    if _EXT is None:
        _EXT = build_ref.load()
        if _EXT is None:
            raise RuntimeError('oracle/_ref/dfa3d_ref_ext.so is missing (build it with oracle/build_ref.py where /root/reference exists)')
    return _EXT


class RefDFA3D(torch.autograd.Function):
    """multi_scale_3ddeformable_attn_function.py:275-351 on the reference extension."""

    @staticmethod
    def forward(ctx, value, dist, shapes3d, lsi, loc, attn, im2col_step):
        e = ext()
        ds = e.ms_depth_score_sample_forward(dist, shapes3d, lsi, loc, im2col_step=im2col_step)
        out = e.wms_deform_attn_forward(value, shapes3d[..., :2].contiguous(), lsi, loc[..., :2].contiguous(), attn, ds,
                                        im2col_step=im2col_step)
        ctx.save_for_backward(value, dist, shapes3d, lsi, loc, attn, ds)
        ctx.step = im2col_step
        return out, ds

    @staticmethod
    def backward(ctx, gout, gds):
        e = ext()
        value, dist, shapes3d, lsi, loc, attn, ds = ctx.saved_tensors
        if gds.sum() != 0.0:  # F3D:314 (host sync in every backward)
            raise NotImplementedError
        g_value = torch.zeros_like(value)
        g_loc2 = torch.zeros([*loc.shape[:-1], 2], dtype=loc.dtype, device=loc.device)
        g_attn = torch.zeros_like(attn)
        g_ds = torch.zeros_like(ds)
        e.wms_deform_attn_backward(value, shapes3d[..., :2].contiguous(), lsi, loc[..., :2].contiguous(), attn, ds,
                                   gout.contiguous(), g_value, g_loc2, g_attn, g_ds, im2col_step=ctx.step)
        g_dist = torch.zeros_like(dist)
        g_loc = torch.zeros_like(loc)
        e.ms_depth_score_sample_backward(dist, shapes3d, lsi, loc, g_ds.contiguous(), g_dist, g_loc, im2col_step=ctx.step)
        g_loc[..., :2] = g_loc[..., :2] + g_loc2
        return g_value, g_dist, None, None, g_loc, g_attn, None


def _p(sd, level, name):
    return sd[f'base_heads.{level}.cross_transformer.encoder.layers.0.{name}']


def point_sampling_gpu(ref_3d, img_meta, dbound, device):
    """encoder.py:179-223 as written (projection built on the CPU and uploaded every call, batched matmul)."""
    eps = 1e-5
    ogfH, ogfW = img_meta['img_shape'][0], img_meta['img_shape'][1]
    origin = torch.tensor(np.asarray(img_meta['lidar2img']['origin'], dtype=np.float32)).to(device)
    rp = ref_3d.view(1, 1, -1, 3) + origin
    rp = rp.permute(1, 0, 2, 3)
    D, B, nq = rp.size()[:3]
    projection = path_ref.compute_projection(img_meta, 1).to(device)
    V = projection.shape[0]
    rp = rp.view(D, B, 1, nq, 3).repeat(1, 1, V, 1, 1)
    rp = torch.cat((rp, torch.ones(*rp.shape[:-1], 1).type_as(rp)), dim=-1)
    cam = torch.matmul(projection.unsqueeze(0).unsqueeze(0), rp.permute(0, 1, 2, 4, 3)).permute(0, 1, 2, 4, 3)
    pd = cam[..., 2:3]
    cam[..., 0:2] = cam[..., 0:2] / torch.maximum(pd, torch.ones_like(pd) * eps)
    cam[..., 0] /= ogfW
    cam[..., 1] /= ogfH
    cam[..., 2] = (cam[..., 2] - dbound[0]) / (dbound[1] - dbound[0])
    m = (pd > eps) & (cam[..., 0:1] > eps) & (cam[..., 0:1] < (1.0 - eps)) & (cam[..., 1:2] > eps) & (cam[..., 1:2] < (1.0 - eps))
    return cam.permute(2, 1, 3, 0, 4), m.permute(2, 1, 3, 0, 4).squeeze(-1)  # [V,B,nq,D,3], [V,B,nq,D]


def point_sampling_pinned_gpu(ref_3d, img_meta, dbound, device):
    """``path_ref.point_sampling`` (the projection contract: explicit fp32 operation order, one rounding per operation) on
    the device, in the layout of ``point_sampling_gpu``.  The reference's batched matmul leaves the accumulation order to the
    library, so a handful of (view, voxel) pairs within round-off of a visibility bound flip between implementations; the
    full-size parity tests pin the projection to compare everything downstream of it."""
    ogfH, ogfW = img_meta['img_shape'][0], img_meta['img_shape'][1]
    origin = torch.tensor(np.asarray(img_meta['lidar2img']['origin'], dtype=np.float32)).to(device)
    p = ref_3d + origin
    P = path_ref.compute_projection(img_meta, 1).to(device)
    px, py, pz = p[:, 0][None], p[:, 1][None], p[:, 2][None]

    def row(i):
        a = P[:, i, 0:1] * px
        a = a + P[:, i, 1:2] * py
        a = a + P[:, i, 2:3] * pz
        return a + P[:, i, 3:4]

    x, y, z = row(0), row(1), row(2)
    zc = torch.maximum(z, torch.full_like(z, path_ref.EPS))
    u = (x / zc) / float(ogfW)
    v = (y / zc) / float(ogfH)
    d = (z - float(dbound[0])) / float(dbound[1] - dbound[0])
    mask = (z > path_ref.EPS) & (u > path_ref.EPS) & (u < (1.0 - path_ref.EPS)) & (v > path_ref.EPS) & (v < (1.0 - path_ref.EPS))
    cam = torch.stack([u, v, d], dim=-1)                      # [V,Q,3]
    return cam[:, None, :, None, :], mask[:, None, :, None]   # [V,1,Q,1,3], [V,1,Q,1]


PINNED_PROJECTION = False


def dense_head_forward_gpu(sd, level, feat, dpt_dist, img_meta, proposal, cfg, training):
    dev = feat.device
    _, V, C, h, w = feat.shape
    Dd = dpt_dist.shape[2]
    M, Pn = cfg.num_heads, cfg.num_points
    n_vox = cfg.n_voxels_list[level]
    N = int(np.prod(n_vox))
    if proposal is None:
        proposal = torch.ones(N, device=dev)
    sel = torch.nonzero(proposal > 0).view(-1)                        # DenseHead.py:66 (host sync)
    Q = sel.numel()
    ref_3d = sd[f'base_heads.{level}.ref_3d'][sel]
    query = torch.zeros(1, Q, C, device=dev)
    # transformer.py:151-170
    value = feat[0].flatten(2).permute(0, 2, 1).contiguous()            # [V,S,C]  (NCHW -> channel-last copy)
    dist = dpt_dist[0].flatten(2).permute(0, 2, 1).contiguous()         # [V,S,D]
    S = h * w
    shapes = torch.as_tensor([[h, w]], dtype=torch.long, device=dev)
    shapes3d = torch.cat([shapes, shapes.new_ones(1, 1) * Dd], dim=-1).contiguous()
    lsi = torch.zeros(1, dtype=torch.long, device=dev)
    sampler = point_sampling_pinned_gpu if PINNED_PROJECTION else point_sampling_gpu
    ref_cam, bev_mask = sampler(ref_3d, img_meta, cfg.dbound, dev)   # [V,1,Q,1,3], [V,1,Q,1]
    # DCA:758-773: per-view index lists + padded rebatch
    indexes = [bev_mask[i][0].sum(-1).nonzero().squeeze(-1) for i in range(V)]   # V host syncs
    max_len = max(len(e) for e in indexes)
    ref_rebatch = ref_cam.new_zeros([1, V, max_len, 1, 3])
    for i in range(V):
        ref_rebatch[0, i, :len(indexes[i])] = ref_cam[i, 0, indexes[i]]
    ref_rebatch = ref_rebatch.view(V, max_len, 1, 3)
    # Grid_Sample_3D_Feature (DCA:67-116)
    loc1 = ref_rebatch[:, :, None, None, None, :, :].view(V, max_len, 1, 1, 1, 3).contiguous()
    q_img, _ = RefDFA3D.apply(value.view(V, S, 1, C), dist.view(V, S, 1, Dd).contiguous(), shapes3d, lsi, loc1,
                              torch.ones(V, max_len, 1, 1, 1, device=dev), 128)
    # MSDeformableAttention3D_DFA3D.forward (DCA:417-489)
    da = 'attentions.0.deformable_attention.'
    val = F.linear(value, _p(sd, level, da + 'value_proj.weight'), _p(sd, level, da + 'value_proj.bias')).view(V, S, M, C // M)
    dist_m = dist.view(V, S, 1, Dd).repeat(1, 1, M, 1)
    off_uv = F.linear(q_img, _p(sd, level, da + 'sampling_offsets.weight'), _p(sd, level, da + 'sampling_offsets.bias')).view(V, max_len, M, 1, Pn, 2)
    off_d = F.linear(q_img, _p(sd, level, da + 'sampling_offsets_depth.weight'), _p(sd, level, da + 'sampling_offsets_depth.bias')).view(V, max_len, M, 1, Pn, 1)
    off = torch.cat([off_uv, off_d], dim=-1)
    aw = F.linear(q_img, _p(sd, level, da + 'attention_weights.weight'), _p(sd, level, da + 'attention_weights.bias')).view(V, max_len, M, Pn)
    aw = aw.softmax(-1).view(V, max_len, M, 1, Pn)
    normalizer = torch.stack([shapes3d[..., 1], shapes3d[..., 0], shapes3d[..., 2]], -1)
    loc = (ref_rebatch[:, :, None, None, None, :, :] + (off / normalizer[None, None, None, :, None, :]).view(V, max_len, M, 1, Pn, 1, 3))
    loc = loc.view(V, max_len, M, 1, Pn, 3).contiguous()
    queries, ds = RefDFA3D.apply(val, dist_m, shapes3d, lsi, loc, aw.contiguous(), 128)
    _ = (ds.mean(dim=-1) * aw).flatten(-2).sum(dim=-1, keepdim=True)   # weight_update, computed and discarded (DCA:492)
    # DCA:815-837
    slots = torch.zeros([V, 1, Q, C], device=dev)
    for i in range(V):
        slots[i, 0, indexes[i]] = queries[i, :len(indexes[i])]
    count = (bev_mask.sum(-1) > 0).permute(1, 2, 0).sum(-1)
    valid_index = count.nonzero()[:, 1]                                   # host sync
    valid_num = count[:, valid_index]
    valid_slots = slots[:, :, valid_index, :]
    valid_mask = bev_mask[:, :, valid_index, :]
    slots_mean = (valid_slots * valid_mask).sum(dim=0) / valid_num[..., None]
    slots_mean = F.linear(slots_mean, _p(sd, level, 'attentions.0.output_proj.weight'), _p(sd, level, 'attentions.0.output_proj.bias'))
    vs = valid_slots.squeeze(1)
    key_padding = ~valid_mask.squeeze(3).squeeze(1).transpose(1, 0)
    pooled, _ = F.multi_head_attention_forward(
        slots_mean.squeeze(0).unsqueeze(0), vs, vs, C, 8,
        _p(sd, level, 'attentions.0.attention_pooling.in_proj_weight'), _p(sd, level, 'attentions.0.attention_pooling.in_proj_bias'),
        None, None, False, 0.0, _p(sd, level, 'attentions.0.attention_pooling.out_proj.weight'),
        _p(sd, level, 'attentions.0.attention_pooling.out_proj.bias'), training=training, key_padding_mask=key_padding,
        need_weights=False)
    output = torch.zeros([1, Q, C], device=dev)
    output[:, valid_index, :] = pooled
    x = output + query
    x = F.layer_norm(x, (C,), _p(sd, level, 'norms.0.weight'), _p(sd, level, 'norms.0.bias'))
    hdn = F.dropout(F.relu(F.linear(x, _p(sd, level, 'ffns.0.layers.0.0.weight'), _p(sd, level, 'ffns.0.layers.0.0.bias'))), 0.1, training)
    x = x + F.dropout(F.linear(hdn, _p(sd, level, 'ffns.0.layers.1.weight'), _p(sd, level, 'ffns.0.layers.1.bias')), 0.1, training)
    x = F.layer_norm(x, (C,), _p(sd, level, 'norms.1.weight'), _p(sd, level, 'norms.1.bias'))
    vol = torch.zeros(N, C, device=dev)
    vol[sel, :] = x[0]
    return vol.reshape(*n_vox, C).permute(3, 0, 1, 2).unsqueeze(0)


def head_forward_gpu(sd, mlvl_feats, img_meta, mlvl_dpt_dists, cfg, training=True, return_masks=False):
    """``return_masks``: also return the per-level {0,1} proposal masks (what the parity tests teacher-force the product with)."""
    nl = cfg.num_levels
    volumes, occ_list, masks = [None] * nl, [], [None] * nl
    for i in range(nl):
        ds = 4 * 2 ** (nl - 1 - i)
        hh, ww = img_meta['img_shape'][0] // ds, img_meta['img_shape'][1] // ds
        fi = nl - 1 - i
        feat = mlvl_feats[fi][:, :, :, :hh, :ww]
        dist = mlvl_dpt_dists[fi][:, :, :, :hh, :ww]
        if i == 0:
            volumes[i] = dense_head_forward_gpu(sd, i, feat, dist, img_meta, None, cfg, training)
        else:
            up = F.interpolate(volumes[i - 1], scale_factor=2, mode='trilinear', align_corners=False)
            occ = torch.sigmoid(F.linear(up.permute(0, 2, 3, 4, 1), sd[f'occ_pred_heads.{i - 1}.0.weight'],
                                         sd[f'occ_pred_heads.{i - 1}.0.bias'])).reshape(1, -1)
            occ_list.append(occ)
            _, idx = torch.topk(occ, k=cfg.topk_list[i - 1], dim=1)          # AdaptiveSparseHead.py:9-13
            m = torch.zeros_like(occ)
            m.scatter_(1, idx, 1.0)
            masks[i] = m.squeeze(0)
            volumes[i] = up + dense_head_forward_gpu(sd, i, feat, dist, img_meta, masks[i], cfg, training)
    occ_preds = torch.cat(occ_list[::-1], dim=1)
    X, Y, Z = cfg.n_voxels_list[-1]
    valid = masks[-1].view(X, Y, Z).bool().long().unsqueeze(0).unsqueeze(0)
    if return_masks:
        return volumes[-1], valid, occ_preds, masks
    return volumes[-1], valid, occ_preds


def bench(config: str, V: int, steps: int, warmup: int) -> dict:
    from sgcdet_b200 import synthetic as syn
    cfg = syn.CONFIGS[config]
    dev = torch.device('cuda', 0)
    torch.backends.cuda.matmul.allow_tf32 = False   # fp32 path, like the reference on torch 1.10 with TF32 off
    sc = syn.make_scene(cfg, V, shift_origin=True).to(dev)
    sd = {k: (v.to(dev).requires_grad_(True) if v.is_floating_point() and 'ref_3d' not in k else v.to(dev))
          for k, v in syn.make_state_dict(cfg).items()}
    feats = [f.clone().requires_grad_(True) for f in sc.mlvl_feats[:3]]
    dists = [d.clone().requires_grad_(True) for d in sc.mlvl_dpt_dists[:3]]
    leaves = [t for t in sd.values() if t.requires_grad] + feats + dists

    def step():
        for t in leaves:
            t.grad = None
        vol, valid, occ = head_forward_gpu(sd, feats, sc.img_meta, dists, cfg, training=True)
        loss = (vol * sc.grad_volume).sum() + path_ref.occ_loss(occ, sc.geo_occ)
        loss.backward()
        return float(loss)

    for _ in range(max(1, warmup)):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {'value': round(1e3 / ms, 2), 'ms_per_step': round(ms, 3), 'n_gpus': 1, 'steps': steps, 'warmup': warmup,
            'higher_is_better': True, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'{cfg.name} view-transform fwd+bwd, V={V} views, 1 scene per step',
                       'what': "reference DFA3D CUDA kernels (unmodified, sm_100a) under the restated reference glue, eager PyTorch"}}


if __name__ == '__main__':
    print(json.dumps(bench(os.environ.get('SGC_CONFIG', 'SGCDet_ScanNet'), int(os.environ.get('SGC_VIEWS', '40')), 5, 2)))
