"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares
(no compute calls), the Python bindings mirror the header, modules keep the reference's names / ctor kwargs /
state-dict keys, and the product refuses to run without CUDA (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

import sgcdet_b200
from sgcdet_b200 import _lib, plugin, synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_decls():
    text = open(os.path.join(ROOT, 'include', 'sgcdet_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return {m.group(1): m.group(2) for m in re.finditer(r'\b(?:int|long long)\s+(\w+)\s*\(([^;]*?)\)\s*;', text, flags=re.S)}


def test_library_exports_every_header_symbol():
    from sgcdet_b200 import build
    build.build()
    lib = ctypes.CDLL(str(_lib.lib_path()))
    decls = _header_decls()
    assert len(decls) >= 60   # grows with the library; every declared symbol is checked below
    for name in decls:
        assert hasattr(lib, name), f'{name} declared in include/sgcdet_b200.h but not exported'


def test_python_bindings_mirror_header():
    decls = _header_decls()
    assert set(decls) == set(_lib.SIGNATURES), set(decls) ^ set(_lib.SIGNATURES)
    for name, args in decls.items():
        params = [a.strip() for a in args.split(',') if a.strip() and a.strip() != 'void']
        sig = _lib.SIGNATURES[name]
        assert len(params) == len(sig), name
        for p, t in zip(params, sig):
            if '*' in p:
                assert t is ctypes.c_void_p, (name, p)
            elif p.startswith('unsigned'):
                assert False, (name, p)
            elif p.startswith('float'):
                assert t is ctypes.c_float, (name, p)
            elif p.startswith('long long'):
                assert t is ctypes.c_longlong, (name, p)
            else:
                assert t is ctypes.c_int, (name, p)


def test_no_cpu_fallback():
    cfg = syn.CONFIGS['tiny']
    head = plugin.build_voxel_head(cfg)
    sc = syn.make_scene(cfg, 2)
    with pytest.raises(RuntimeError):
        head(sc.mlvl_feats, sc.img_meta, sc.mlvl_dpt_dists)
    sgcdet_b200.install_dropin()
    from dfa3D import ext_loader
    ext = ext_loader.load_ext('_ext', ['wms_deform_attn_backward', 'wms_deform_attn_forward',
                                       'ms_depth_score_sample_forward', 'ms_depth_score_sample_backward'])
    with pytest.raises(RuntimeError):
        ext.ms_depth_score_sample_forward(torch.rand(1, 4, 1, 2), torch.tensor([[2, 2, 2]]), torch.zeros(1, dtype=torch.long),
                                          torch.rand(1, 3, 1, 1, 1, 3), im2col_step=64)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'sgcdet_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f


@pytest.mark.parametrize('name', ['SGCDet_ScanNet', 'SGCDet_ARKit', 'SGCDet_large_ScanNet200', 'SGCDet_large_ARKit'])
def test_state_dict_keys_and_shapes(name):
    """Checkpoint compatibility: same keys and shapes as the reference modules."""
    cfg = syn.CONFIGS[name]
    head = plugin.build_voxel_head(cfg)
    sd = head.state_dict()
    C = cfg.embed_dims
    for i in range(3):
        pre = f'base_heads.{i}.cross_transformer.encoder.layers.0.'
        for k, shp in syn.level_param_shapes(C).items():
            assert tuple(sd[pre + k].shape) == shp, pre + k
        assert tuple(sd[f'base_heads.{i}.vox_coords'].shape) == (int(torch.tensor(cfg.n_voxels_list[i]).prod()), 4)
        assert sd[f'base_heads.{i}.vox_coords'].dtype == torch.int64
    assert tuple(sd['occ_pred_heads.0.0.weight'].shape) == (1, C) and 'occ_pred_heads.1.0.bias' in sd
    assert len(sd) == 3 * (len(syn.level_param_shapes(C)) + 2) + 4
    head.load_state_dict(syn.make_state_dict(cfg), strict=True)
    # voxel 'centres' are lower corners (DenseHead.py:44-45)
    n, vs = torch.tensor(cfg.n_voxels_list[2]), torch.tensor(cfg.voxel_size_list[2])
    torch.testing.assert_close(sd['base_heads.2.ref_3d'][0], -n / 2. * vs)


def test_builds_from_reference_style_config_dict():
    """The dict a SGCDet_*.py config hands to build_head (configs/SGCDet_ScanNet.py:17-68,113-119)."""
    embed_dims = 256
    cross_transformer = dict(
        type='PerceptionTransformer_DFA3D', embed_dims=embed_dims,
        encoder=dict(type='VoxFormerEncoder_DFA3D', num_layers=1, return_intermediate=False, dbound=[0.2, 5, 0.4],
                     transformerlayers=dict(
                         type='VoxFormerLayer',
                         attn_cfgs=[dict(type='DeformCrossAttention_DFA3D',
                                         deformable_attention=dict(type='MSDeformableAttention3D_DFA3D', embed_dims=embed_dims,
                                                                   num_heads=8, num_points=4, num_levels=1, im2col_step=128),
                                         embed_dims=embed_dims, inter_view_aggregation='attn', dropout=0)],
                         ffn_cfgs=dict(type='FFN', embed_dims=256, feedforward_channels=embed_dims * 2, num_fcs=2,
                                       ffn_drop=0.1, act_cfg=dict(type='ReLU', inplace=True)),
                         operation_order=('cross_attn', 'norm', 'ffn', 'norm'))))
    vs = [(.64, .64, .8), (.32, .32, .4), (.16, .16, .2)]
    nv = [(10, 10, 4), (20, 20, 8), (40, 40, 16)]
    cfg = dict(type='AdaptiveSparseHead', embed_dims=embed_dims, topk_list=[800, 6400], voxel_size_list=vs, n_voxels_list=nv,
               base_head_configs=[dict(type='DenseHead', voxel_size=vs[i], n_voxels=nv[i], embed_dims=embed_dims,
                                       cross_transformer=cross_transformer) for i in range(3)])
    head = plugin.build_voxel_head(cfg)
    assert isinstance(head, plugin.AdaptiveSparseHead) and len(head.base_heads) == 3
    da = head.base_heads[0].cross_transformer.encoder.layers[0].attentions[0].deformable_attention
    # reference default init (DCA:194-212,351-362): zero offset weights, ring biases
    assert da.sampling_offsets.weight.abs().max() == 0 and da.attention_weights.weight.abs().max() == 0
    b = da.sampling_offsets.bias.view(8, 1, 4, 2)
    torch.testing.assert_close(b[0, 0, :, 0], torch.tensor([1., 2., 3., 4.]))
    torch.testing.assert_close(da.sampling_offsets_depth.bias.view(8, 4)[0], torch.tensor([.5, 1., 1.5, 2.]))


def test_folded_weights_layout():
    """Wcat rows: value_proj then [m][p][ox,oy,od,logit] (csrc/sgc_lift.cu)."""
    da = plugin.MSDeformableAttention3D_DFA3D(embed_dims=128, num_heads=8, num_levels=1, num_points=4)
    with torch.no_grad():
        for p in da.parameters():
            p.copy_(torch.randn_like(p))
    wcat, vb, gb = da.folded_weights()
    assert wcat.shape == (128 + 128, 128) and gb.shape == (128,)
    x = torch.randn(5, 128)
    g = x @ wcat[128:].t() + gb
    off = da.sampling_offsets(x).view(5, 8, 4, 2)
    offd = da.sampling_offsets_depth(x).view(5, 8, 4)
    logit = da.attention_weights(x).view(5, 8, 4)
    g = g.view(5, 8, 4, 4)
    torch.testing.assert_close(g[..., :2], off)
    torch.testing.assert_close(g[..., 2], offd)
    torch.testing.assert_close(g[..., 3], logit)


def test_algorithmic_bytes_match_baseline_table():
    """BASELINE.md section 4."""
    b = syn.algorithmic_bytes('SGCDet_ScanNet', 40)
    assert abs(b['fwd'] / 1e9 - 0.84) < 0.01 and abs(b['fwd_bwd'] / 1e9 - 2.20) < 0.01
    b = syn.algorithmic_bytes('SGCDet_large_ScanNet200', 40)
    assert abs(b['fwd'] / 1e9 - 0.63) < 0.01 and abs(b['fwd_bwd'] / 1e9 - 1.53) < 0.01


def test_host_side_planning_helpers():
    """Host-only entry points of the C ABI (no device work): column split of the voxel-count GEMM kernel, scratch sizes of
    the weight-gradient kernel and of the many-CTA top-k, and the argument checks that precede any launch."""
    lib = _lib.load()
    # the widest column part of {256,128,64,32} dividing N that still yields enough work items to spread over the SMs
    assert lib.sgc_rows_gemm_tc_auto_ncta(6400, 256, 1) == 128      # 50 row tiles x 2 parts
    assert lib.sgc_rows_gemm_tc_auto_ncta(6400, 512, 1) == 256      # 50 x 2
    assert lib.sgc_rows_gemm_tc_auto_ncta(400, 256, 1) == 64        # 4 row tiles (coarsest level): 16 work items
    assert lib.sgc_rows_gemm_tc_auto_ncta(400, 512, 1) == 128
    assert lib.sgc_rows_gemm_tc_auto_ncta(800, 256, 1) == 32        # 7 row tiles: split as far as possible
    assert lib.sgc_rows_gemm_tc_auto_ncta(6400, 256, 8) == 256      # 8 heads: 400 work items already
    assert lib.sgc_rows_gemm_tc_auto_ncta(6400, 32, 8) == 32        # per-head output of 32 columns
    assert lib.sgc_rows_gemm_tc_auto_ncta(100, 96, 1) == 32         # 96 = 3 x 32
    assert lib.sgc_rows_gemm_tc_auto_ncta(100, 48, 1) == 0          # N % 32 != 0 -> rejected
    # split-K partials [kch][B][M][N] + column sums; kch <= 148 / tiles and >= 128 rows per CTA
    assert lib.sgc_rows_wgrad_tc_scratch_floats(256, 256, 6400, 1) == 50 * (256 * 256 + 256)
    assert lib.sgc_rows_wgrad_tc_scratch_floats(256, 256, 400, 1) == 3 * (256 * 256 + 256)
    assert lib.sgc_rows_wgrad_tc_scratch_floats(256, 32, 6400, 8) == 9 * 8 * (256 * 32 + 256)
    assert lib.sgc_rows_wgrad_tc_scratch_floats(100, 256, 6400, 1) == 0   # M % 128 != 0
    assert lib.sgc_topk_scratch_ints(204800) == 2052 + 2 * 100
    assert lib.sgc_topk_scratch_ints(1) == 2052 + 2
    # argument errors are reported before anything is launched (cudaErrorInvalidValue = 1)
    assert lib.sgc_rows_gemm_tc(None, 256, 0, 10, 256, 1, None, 256, 0, 0, None, 0, 256, None, 256, 0, 0, None) == 1
    assert lib.sgc_rows_wgrad_tc(None, 256, 0, 256, None, 256, 0, 256, 10, 1, None, 0, 256, 1, 1.0, None, 0, None, None) == 1
    assert lib.sgc_topk_select_mc(None, 10, 11, None, None, None, None) == 1
