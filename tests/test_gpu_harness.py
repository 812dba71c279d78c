"""The train-step harness of BASELINE.json configs[2] (sgcdet_b200/harness.py): a detector-shaped module around the real view
transform takes optimiser steps -- losses finite, every trainable parameter of the view-transform head receives a gradient
and moves, the frozen backbone stage does not."""
import pytest
import torch

from sgcdet_b200 import harness, synthetic as syn

pytestmark = pytest.mark.gpu


def test_train_steps_update_the_view_transform(cuda_lib):
    cfg = syn.CONFIGS['SGCDet_ARKit']
    torch.manual_seed(0)
    model = harness.SGCDetShaped(cfg).cuda().train()
    model.voxel_head.load_state_dict(syn.make_state_dict(cfg))
    batch = harness.make_batch(cfg, 6, 'cuda')
    opt, sched = harness.configure_optimizers(model, total_steps=10)
    params = [p for p in model.parameters() if p.requires_grad]
    before = {k: p.detach().clone() for k, p in model.voxel_head.named_parameters()}
    frozen = model.backbone['layer1'][0].conv1.weight.detach().clone()
    losses = [float(harness.train_step(model, batch, opt, sched, params)) for _ in range(3)]
    assert all(l == l and abs(l) < 1e6 for l in losses), losses
    moved = [k for k, p in model.voxel_head.named_parameters() if not torch.equal(p.detach(), before[k])]
    assert len(moved) == len(before), set(before) - set(moved)
    assert torch.equal(model.backbone['layer1'][0].conv1.weight, frozen)
    assert [g['name'] for g in opt.param_groups] == ['backbone', 'others']
    assert opt.param_groups[0]['lr'] < opt.param_groups[1]['lr']
