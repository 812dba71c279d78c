"""Train-step harness for BASELINE.json configs[2] ("SGCDet_ARKit full train step with the new view-transform, scene-batch
data-parallel"): a detector-SHAPED module around the real view-transform head, so that a full step -- images in, losses,
backward, gradient all-reduce, AdamW + OneCycleLR -- can be timed without mmcv / mmdet / Lightning (none of which is installed
here).  SURVEY.md section 8(f)-4.

What is REAL: ``voxel_head`` = ``sgcdet_b200.plugin.AdaptiveSparseHead`` (the path of this repository) and its ``occ_loss``;
the data flow of ``SGCDet.build_volume`` / ``forward_train`` (mmdet3d_plugin/models/detectors/SGCDet.py:61-113): images
``[B, N, 3, H, W]`` -> backbone -> FPN -> reshape to ``[B, N, C, h, w]`` -> depth distribution at stride 4 and its nearest
/2, /4 pyramid (``:83-85``) -> voxel head -> 3-D neck -> detection head losses (+ ``occ_loss``); the optimiser set-up of
``LightningTools/pl_model.py:92-142`` (AdamW, backbone parameters at 0.1x the learning rate, OneCycleLR stepped per
iteration); one scene per rank and a gradient all-reduce (what Lightning's DDP strategy does, ``main.py:63-86``).

What is a STAND-IN with the reference's tensor shapes (out of scope of this repository): the backbone is
torchvision's ResNet-50 with random weights (the config loads ``torchvision://resnet50``; stage 1 frozen, BN in eval mode as in
``configs/SGCDet_ARKit.py:76-86``), the FPN is torchvision's, ``DepthNet_Fusion`` is a 3x3 convolution + softmax over the 12
depth bins, ``FastIndoorImVoxelNeck`` is three ``BasicBlock3dV2``-shaped residual blocks (necks/imvoxelnet.py:146-173) giving
128 channels at three scales, and the detection head is one 3x3x3 convolution per scale with centerness / box / class outputs
trained against fixed synthetic targets inside ``valid``.  No number produced with this module says anything about detection
quality; it exists to time the step around the view transform.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import plugin, synthetic as syn


class _Block3d(nn.Module):
    """``BasicBlock3dV2`` shape (necks/imvoxelnet.py:146-173): conv-bn-relu-conv-bn (+ strided 1x1 identity), relu."""

    def __init__(self, cin: int, cout: int, stride: int = 1):
        super().__init__()
        self.conv1 = nn.Conv3d(cin, cout, 3, stride, 1, bias=False)
        self.norm1 = nn.BatchNorm3d(cout)
        self.conv2 = nn.Conv3d(cout, cout, 3, 1, 1, bias=False)
        self.norm2 = nn.BatchNorm3d(cout)
        self.down = None
        if stride != 1 or cin != cout:
            self.down = nn.Sequential(nn.Conv3d(cin, cout, 1, stride, bias=False), nn.BatchNorm3d(cout))

    def forward(self, x):
        idt = x if self.down is None else self.down(x)
        out = F.relu(self.norm1(self.conv1(x)), inplace=True)
        out = self.norm2(self.conv2(out))
        return F.relu(out + idt, inplace=True)


class SGCDetShaped(nn.Module):
    """See the module docstring.  ``forward_train(batch) -> dict of losses`` like ``SGCDet.forward_train``."""

    def __init__(self, cfg: syn.PathConfig, n_classes: int = 17, pretrained_backbone: bool = False):
        super().__init__()
        import torchvision
        from torchvision.ops import FeaturePyramidNetwork
        C = cfg.embed_dims
        self.cfg = cfg
        r = torchvision.models.resnet50(weights=None)
        self.backbone = nn.ModuleDict(dict(stem=nn.Sequential(r.conv1, r.bn1, r.relu, r.maxpool), layer1=r.layer1, layer2=r.layer2,
                                           layer3=r.layer3, layer4=r.layer4))
        for m in (self.backbone['stem'], self.backbone['layer1']):        # frozen_stages=1
            for p in m.parameters():
                p.requires_grad_(False)
        self.neck = FeaturePyramidNetwork([256, 512, 1024, 2048], C)
        self.depth_head = nn.Conv2d(C, cfg.depth_bins, 3, padding=1)      # stand-in for DepthNet_Fusion
        self.voxel_head = plugin.build_voxel_head(cfg)
        self.neck_3d = nn.ModuleList([_Block3d(C, 128), _Block3d(128, 128, 2), _Block3d(128, 128, 2)])
        self.bbox_head = nn.ModuleList([nn.Conv3d(128, 1 + 7 + n_classes, 3, padding=1) for _ in range(3)])
        self.n_classes = n_classes

    def train(self, mode: bool = True):
        super().train(mode)
        for m in self.backbone.modules():                                  # norm_eval=True
            if isinstance(m, nn.BatchNorm2d):
                m.eval()
        return self

    def build_volume(self, batch):
        img = batch['img']
        B, N = img.shape[:2]
        x = img.reshape(B * N, *img.shape[2:])
        x = self.backbone['stem'](x)
        feats = {}
        for i, name in enumerate(('layer1', 'layer2', 'layer3', 'layer4')):
            x = self.backbone[name](x)
            feats[str(i)] = x
        pyr = list(self.neck(feats).values())
        pyr = [f.reshape(B, N, *f.shape[1:]) for f in pyr]                 # [B, N, C, h, w]
        dpt = self.depth_head(pyr[0].flatten(0, 1)).softmax(1)
        dpt = dpt.reshape(B, N, *dpt.shape[1:])                            # [B, N, D, h, w]
        mlvl_dpt = [dpt, F.interpolate(dpt, scale_factor=(1, 0.5, 0.5), mode='nearest'),
                    F.interpolate(dpt, scale_factor=(1, 0.25, 0.25), mode='nearest')]
        volume, valid, occ = self.voxel_head([f.contiguous() for f in pyr], batch['img_metas'][0], mlvl_dpt)
        return volume, valid, occ

    def forward_train(self, batch) -> Dict[str, torch.Tensor]:
        volume, valid, occ = self.build_volume(batch)
        x, outs = volume, []
        for blk in self.neck_3d:
            x = blk(x)
            outs.append(x)
        losses = {}
        v = valid.float()
        cen = box = cls = 0.0
        for i, (f, head) in enumerate(zip(outs, self.bbox_head)):
            o = head(f)
            m = v if i == 0 else F.max_pool3d(v, 2 ** i)
            t = batch['det_targets'][i]
            w = m / m.sum().clamp(min=1.0)
            cen = cen + (F.binary_cross_entropy_with_logits(o[:, :1], t[:, :1], reduction='none') * w).sum()
            box = box + (F.smooth_l1_loss(o[:, 1:8], t[:, 1:8], reduction='none') * w).sum() / 7
            cls = cls + (F.binary_cross_entropy_with_logits(o[:, 8:], t[:, 8:], reduction='none') * w).sum() / self.n_classes
        losses.update(loss_centerness=cen, loss_bbox=box, loss_cls=cls)
        losses.update(self.voxel_head.occ_loss(occ, None, batch['geo_occ']))
        return losses


def make_batch(cfg: syn.PathConfig, num_views: int, device, seed: int = 1234, n_classes: int = 17) -> dict:
    """One synthetic scene in the layout ``SGCDet.forward_train`` consumes: normalised images, image meta, targets."""
    g = torch.Generator().manual_seed(seed)
    meta = syn.make_img_meta(cfg, num_views, g, shift_origin=True)
    H, W = cfg.img_shape
    H = (H + 31) // 32 * 32 if H % 8 else H                               # Pad(size=(240, 320)) of the pipeline
    img = torch.randn(1, num_views, 3, H, W, generator=g)
    X, Y, Z = cfg.n_voxels_list[-1]
    tg = [torch.rand(1, 1 + 7 + n_classes, X >> i, Y >> i, Z >> i, generator=g).round_() for i in range(3)]
    n_occ = sum(int(torch.tensor(n).prod()) for n in cfg.n_voxels_list[1:])
    geo = (torch.rand(1, n_occ, generator=g) < 0.2).float()
    return dict(img=img.to(device), img_metas=[meta], det_targets=[t.to(device) for t in tg], geo_occ=geo.to(device))


def configure_optimizers(model: nn.Module, lr: float = 2e-4, weight_decay: float = 1e-4, total_steps: int = 1000):
    """LightningTools/pl_model.py:92-142: AdamW with the backbone at 0.1x, OneCycleLR stepped per iteration."""
    groups = [dict(params=[p for n, p in model.named_parameters() if p.requires_grad and 'backbone' in n], lr=lr * 0.1,
                   weight_decay=weight_decay, name='backbone'),
              dict(params=[p for n, p in model.named_parameters() if p.requires_grad and 'backbone' not in n], lr=lr,
                   weight_decay=weight_decay, name='others')]
    opt = torch.optim.AdamW(groups)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=[lr * 0.1, lr], total_steps=total_steps, pct_start=0.05,
                                                cycle_momentum=False, anneal_strategy='cos', final_div_factor=1e2)
    return opt, sched


def train_step(model: SGCDetShaped, batch: dict, opt, sched, params: List[nn.Parameter], world: int = 1) -> torch.Tensor:
    """``pl_model.training_step`` + what the trainer does around it: sum of the losses, backward, gradient averaging over the
    ranks (one flattened NCCL all-reduce), optimiser and scheduler step."""
    from . import parallel
    losses = model.forward_train(batch)
    loss = sum(losses.values())
    opt.zero_grad(set_to_none=True)
    loss.backward()
    if world > 1:
        parallel.allreduce_gradients(params)
    opt.step()
    sched.step()
    return loss.detach()
