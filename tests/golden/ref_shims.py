"""Minimal stand-ins for mmcv / mmdet / dfa3D so that the reference's OWN view-transform modules can be imported
from /root/reference in this container (no mmcv, no GPU) and run on the CPU to produce golden vectors.

Used only by tests/golden/make_golden_plugin.py.  What is stood in, and how it follows the third-party code
(mmcv-full 1.5.3 / mmdet 2.25.1, docs/install.md:5-8 of the reference; not vendored there):

  * Registry / build_from_cfg: ``cls(**cfg_without_type)``.
  * BaseModule / ModuleList / Sequential: torch.nn equivalents (init_cfg ignored).
  * FFN: ``layers = Sequential(Sequential(Linear, act, Dropout), Linear, Dropout)``, ``forward = identity + layers(x)``
    (mmcv/cnn/bricks/transformer.py FFN, num_fcs=2, add_identity=True, dropout_layer=None).
  * build_norm_layer(dict(type='LN'), C) -> ('ln', nn.LayerNorm(C)); xavier_init / constant_init as in mmcv.cnn.
  * TransformerLayerSequence: builds ``num_layers`` layers from ``transformerlayers`` into ``self.layers``.
  * force_fp32 / auto_fp16 / deprecated_api_warning: identity decorators (fp16_enabled is False everywhere).
  * dfa3D._ext: the reference's CUDA-only extension is replaced by the CPU oracle kernels (oracle/dfa3d_ref.py,
    themselves pinned against the real extension on the B200, tests/golden/dfa3d_ref_*.pt).

The reference's two `if torch.cuda.is_available() and value.is_cuda:` guards
(deformable_cross_attention.py:108,482) have no CPU branch; the loader rewrites exactly those two lines to
`if True:` when executing the module (nothing else is modified, nothing is copied into this repository).
"""
from __future__ import annotations

import copy
import importlib
import importlib.util
import sys
import types

import torch
import torch.nn as nn

REF = '/root/reference'
I2V = REF + '/mmdet3d_plugin/models/im2voxel'


class Registry:
    def __init__(self, name):
        self.name, self.module_dict = name, {}

    def register_module(self, name=None, force=False, module=None):
        def _r(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls
        return _r(module) if module is not None else _r

    def build(self, cfg, **kw):
        cfg = dict(copy.deepcopy(cfg))
        cfg.update(kw)
        typ = cfg.pop('type')
        return self.module_dict[typ](**cfg)


class ConfigDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    __setattr__ = dict.__setitem__


class BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self._is_init = False
        self.init_cfg = init_cfg

    def init_weights(self):
        pass


class FFN(BaseModule):
    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2, act_cfg=dict(type='ReLU', inplace=True),
                 ffn_drop=0., dropout_layer=None, add_identity=True, init_cfg=None, **kwargs):
        super().__init__(init_cfg)
        assert num_fcs == 2 and dropout_layer is None
        self.embed_dims = embed_dims
        self.layers = nn.Sequential(
            nn.Sequential(nn.Linear(embed_dims, feedforward_channels), nn.ReLU(inplace=True), nn.Dropout(ffn_drop)),
            nn.Linear(feedforward_channels, embed_dims), nn.Dropout(ffn_drop))
        self.add_identity = add_identity

    def forward(self, x, identity=None):
        out = self.layers(x)
        if not self.add_identity:
            return out
        return (x if identity is None else identity) + out


def _identity_decorator(*dargs, **dkw):
    if len(dargs) == 1 and callable(dargs[0]) and not dkw:
        return dargs[0]
    return lambda f: f


def xavier_init(module, gain=1, bias=0, distribution='normal'):
    if hasattr(module, 'weight') and module.weight is not None:
        (nn.init.xavier_uniform_ if distribution == 'uniform' else nn.init.xavier_normal_)(module.weight, gain=gain)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def constant_init(module, val, bias=0):
    if hasattr(module, 'weight') and module.weight is not None:
        nn.init.constant_(module.weight, val)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def install():
    """Register the stand-in modules in sys.modules and return the imported reference modules."""
    from oracle import dfa3d_ref

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    regs = {n: Registry(n) for n in ('ATTENTION', 'TRANSFORMER_LAYER', 'TRANSFORMER_LAYER_SEQUENCE',
                                     'FEEDFORWARD_NETWORK', 'POSITIONAL_ENCODING', 'HEADS', 'TRANSFORMER')}
    regs['FEEDFORWARD_NETWORK'].register_module(name='FFN', module=FFN)

    class TransformerLayerSequence(BaseModule):
        def __init__(self, transformerlayers=None, num_layers=None, init_cfg=None):
            super().__init__(init_cfg)
            if isinstance(transformerlayers, dict):
                transformerlayers = [copy.deepcopy(transformerlayers) for _ in range(num_layers)]
            self.num_layers = num_layers
            self.layers = nn.ModuleList([regs['TRANSFORMER_LAYER'].build(c) for c in transformerlayers])
            self.embed_dims = self.layers[0].embed_dims
            self.pre_norm = self.layers[0].pre_norm

    ext_dummy = types.SimpleNamespace(ms_deform_attn_forward=None, ms_deform_attn_backward=None)
    mmcv = mod('mmcv', ConfigDict=ConfigDict, deprecated_api_warning=_identity_decorator)
    mmcv.__path__ = []
    mod('mmcv.ops').__path__ = []
    mod('mmcv.ops.multi_scale_deform_attn', multi_scale_deformable_attn_pytorch=None)
    mod('mmcv.cnn', xavier_init=xavier_init, constant_init=constant_init, Linear=nn.Linear,
        build_activation_layer=lambda cfg: nn.ReLU(inplace=cfg.get('inplace', False)),
        build_norm_layer=lambda cfg, n: ('ln', nn.LayerNorm(n))).__path__ = []
    mod('mmcv.cnn.bricks').__path__ = []
    mod('mmcv.cnn.bricks.registry', **{k: regs[k] for k in ('ATTENTION', 'TRANSFORMER_LAYER', 'TRANSFORMER_LAYER_SEQUENCE',
                                                            'FEEDFORWARD_NETWORK', 'POSITIONAL_ENCODING')})
    mod('mmcv.cnn.bricks.transformer', build_attention=regs['ATTENTION'].build,
        build_transformer_layer=regs['TRANSFORMER_LAYER'].build,
        build_transformer_layer_sequence=regs['TRANSFORMER_LAYER_SEQUENCE'].build,
        build_feedforward_network=regs['FEEDFORWARD_NETWORK'].build, TransformerLayerSequence=TransformerLayerSequence)
    mod('mmcv.runner', force_fp32=_identity_decorator, auto_fp16=_identity_decorator).__path__ = []
    mod('mmcv.runner.base_module', BaseModule=BaseModule, ModuleList=nn.ModuleList, Sequential=nn.Sequential)
    mod('mmcv.utils', ext_loader=types.SimpleNamespace(load_ext=lambda name, funcs: ext_dummy),
        TORCH_VERSION=torch.__version__, digit_version=lambda v: tuple(int(x) for x in v.split('+')[0].split('.')[:3]))
    mmdet = mod('mmdet')
    mmdet.__path__ = []
    mod('mmdet.models', HEADS=regs['HEADS'], build_head=regs['HEADS'].build).__path__ = []
    mod('mmdet.models.utils', build_transformer=regs['TRANSFORMER'].build).__path__ = []
    mod('mmdet.models.utils.builder', TRANSFORMER=regs['TRANSFORMER'])

    # dfa3D._ext on the CPU = the oracle kernels, with the reference extension's calling convention
    def ms_depth_score_sample_forward(value, shapes, lsi, loc, im2col_step):
        return dfa3d_ref.depth_score_sample_forward(value, shapes, lsi, loc)

    def wms_deform_attn_forward(value, shapes, lsi, loc, attn, ds, im2col_step):
        return dfa3d_ref.wms_deform_attn_forward(value, shapes, lsi, loc, attn, ds)

    def wms_deform_attn_backward(value, shapes, lsi, loc, attn, ds, grad_output, grad_value, grad_loc, grad_attn,
                                 grad_ds, im2col_step):
        with torch.enable_grad():  # the reference calls this under once_differentiable (no-grad mode)
            gv, gl, ga, gd = dfa3d_ref.wms_deform_attn_backward(value, shapes, lsi, loc, attn, ds, grad_output)
        grad_value.add_(gv); grad_loc.copy_(gl); grad_attn.copy_(ga); grad_ds.copy_(gd)

    def ms_depth_score_sample_backward(value, shapes, lsi, loc, grad_output, grad_value, grad_loc, im2col_step):
        with torch.enable_grad():
            gv, gl = dfa3d_ref.depth_score_sample_backward(value, shapes, lsi, loc, grad_output)
        grad_value.add_(gv); grad_loc.copy_(gl)

    ext = mod('dfa3D._ext', ms_depth_score_sample_forward=ms_depth_score_sample_forward,
              wms_deform_attn_forward=wms_deform_attn_forward, wms_deform_attn_backward=wms_deform_attn_backward,
              ms_depth_score_sample_backward=ms_depth_score_sample_backward)
    dfa = mod('dfa3D', ext_loader=types.SimpleNamespace(load_ext=lambda name, funcs: ext))
    dfa.__path__ = []
    mod('dfa3D.ext_loader', load_ext=lambda name, funcs: ext)

    # the reference package skeleton, without executing its __init__ files (they import datasets, heads, ...)
    pkg = mod('ref_i2v')
    pkg.__path__ = [I2V]
    tu = mod('ref_i2v.transformer_utils')
    tu.__path__ = [I2V + '/transformer_utils']

    def load(name, path, patches=()):
        src = open(path).read()
        for old, new in patches:
            assert src.count(old) >= 1, (path, old)
            src = src.replace(old, new)
        spec = importlib.util.spec_from_loader(name, loader=None, origin=path)
        m = importlib.util.module_from_spec(spec)
        m.__file__ = path
        m.__package__ = name.rsplit('.', 1)[0]
        sys.modules[name] = m
        exec(compile(src, path, 'exec'), m.__dict__)
        return m

    T = I2V + '/transformer_utils/'
    load('ref_i2v.transformer_utils.multi_scale_deformable_attn_function', T + 'multi_scale_deformable_attn_function.py')
    load('ref_i2v.transformer_utils.multi_scale_3ddeformable_attn_function', T + 'multi_scale_3ddeformable_attn_function.py')
    load('ref_i2v.transformer_utils.custom_base_transformer_layer', T + 'custom_base_transformer_layer.py')
    dca = load('ref_i2v.transformer_utils.deformable_cross_attention', T + 'deformable_cross_attention.py',
               patches=[('if torch.cuda.is_available() and value.is_cuda:', 'if True:')])
    enc = load('ref_i2v.transformer_utils.encoder', T + 'encoder.py')
    trf = load('ref_i2v.transformer_utils.transformer', T + 'transformer.py')
    dh = load('ref_i2v.DenseHead', I2V + '/DenseHead.py')
    ash = load('ref_i2v.AdaptiveSparseHead', I2V + '/AdaptiveSparseHead.py')
    return dict(dca=dca, encoder=enc, transformer=trf, DenseHead=dh, AdaptiveSparseHead=ash, registries=regs)
