"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (mv2d_b200/).

Stand-ins for the handful of mmcv 1.6.1 / mmdet 2.25.1 / mmdet3d 1.0.0 names that the
reference's hot-path files import at module scope, so that the files under
``/root/reference/mmdet3d_plugin/models`` import and run UNMODIFIED on CPU in this container
(mmcv/mmdet/mmdet3d are not installed and cannot be: no network).  This is the "strongest
oracle" of SURVEY.md section 8c / App. F: the reference's own Python produces the golden
vectors committed under ``tests/golden/`` (see ``oracle/make_golden.py``).

Semantics follow SURVEY.md App. A (third-party behaviour, restated from the pinned versions
named at /root/reference/README.md:12).  ``/root/reference`` exists only in the build
container; nothing that runs on the GPU box imports this file.
"""
import copy
import functools
import sys
import types
import warnings

import torch
import torch.nn as nn
import torch.nn.functional as F

REFERENCE_ROOT = '/root/reference'


# --------------------------------------------------------------------------- registry
class Registry:
    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def _reg(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls
        if module is not None:
            return _reg(module)
        return _reg

    def get(self, key):
        return self.module_dict.get(key)

    def build(self, cfg, default_args=None):
        return build_from_cfg(cfg, self, default_args)


class ConfigDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def build_from_cfg(cfg, registry, default_args=None):
    args = dict(cfg)
    if default_args:
        for k, v in default_args.items():
            args.setdefault(k, v)
    t = args.pop('type')
    cls = registry.get(t) if isinstance(t, str) else t
    if cls is None:
        raise KeyError(f'{t} is not in the {registry.name} registry')
    return cls(**args)


# --------------------------------------------------------------------------- mmcv.runner
class BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self._is_init = False
        self.init_cfg = copy.deepcopy(init_cfg)

    def init_weights(self):
        pass


class ModuleList(BaseModule, nn.ModuleList):
    def __init__(self, modules=None, init_cfg=None):
        BaseModule.__init__(self, init_cfg)
        nn.ModuleList.__init__(self, modules)


class Sequential(BaseModule, nn.Sequential):
    def __init__(self, *args, init_cfg=None):
        BaseModule.__init__(self, init_cfg)
        nn.Sequential.__init__(self, *args)


def _noop_decorator_factory(*dargs, **dkwargs):
    def deco(fn):
        return fn
    return deco


def deprecated_api_warning(name_dict, cls_name=None):
    def deco(fn):
        @functools.wraps(fn)
        def wrapped(*args, **kwargs):
            for old, new in name_dict.items():
                if old in kwargs:
                    kwargs[new] = kwargs.pop(old)
            return fn(*args, **kwargs)
        return wrapped
    return deco


def to_2tuple(x):
    return (x, x) if not isinstance(x, (tuple, list)) else tuple(x)


# --------------------------------------------------------------------------- mmcv.cnn
ATTENTION = Registry('attention')
TRANSFORMER_LAYER = Registry('transformerLayer')
TRANSFORMER_LAYER_SEQUENCE = Registry('transformer-layers sequence')
POSITIONAL_ENCODING = Registry('position encoding')
FEEDFORWARD_NETWORK = Registry('feed-forward Network')
DROPOUT_LAYERS = Registry('drop out layers')
TRANSFORMER = Registry('Transformer')
HEADS = Registry('head')
LOSSES = Registry('loss')
ROI_EXTRACTORS = Registry('roi_extractor')
BBOX_CODERS = Registry('bbox_coder')
BBOX_ASSIGNERS = Registry('bbox_assigner')
BBOX_SAMPLERS = Registry('bbox_sampler')


class Dropout(nn.Dropout):
    def __init__(self, drop_prob=0.5, inplace=False):
        super().__init__(p=drop_prob, inplace=inplace)


DROPOUT_LAYERS.register_module(module=Dropout)


def build_dropout(cfg, default_args=None):
    return build_from_cfg(cfg, DROPOUT_LAYERS, default_args)


def build_activation_layer(cfg):
    cfg = dict(cfg)
    t = cfg.pop('type')
    return {'ReLU': nn.ReLU, 'GELU': nn.GELU, 'Sigmoid': nn.Sigmoid}[t](**cfg)


def build_norm_layer(cfg, num_features, postfix=''):
    cfg = dict(cfg)
    t = cfg.pop('type')
    cfg.pop('requires_grad', None)
    assert t == 'LN', t
    cfg.setdefault('eps', 1e-5)
    return 'ln' + str(postfix), nn.LayerNorm(num_features, **cfg)


def build_conv_layer(cfg, *args, **kwargs):
    assert cfg is None or cfg.get('type', 'Conv2d') in ('Conv2d', 'Conv')
    return nn.Conv2d(*args, **kwargs)


def xavier_init(module, gain=1, bias=0, distribution='normal'):
    if hasattr(module, 'weight') and module.weight is not None:
        if distribution == 'uniform':
            nn.init.xavier_uniform_(module.weight, gain=gain)
        else:
            nn.init.xavier_normal_(module.weight, gain=gain)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def bias_init_with_prob(prior_prob):
    import math
    return float(-math.log((1 - prior_prob) / prior_prob))


class ConvModule(nn.Module):
    """mmcv ConvModule with norm_cfg=None: Conv2d(bias=True) + ReLU(inplace)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0,
                 conv_cfg=None, norm_cfg=None, act_cfg=dict(type='ReLU'), **kwargs):
        super().__init__()
        assert norm_cfg is None
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride,
                              padding=padding, bias=True)
        self.activate = nn.ReLU(inplace=True) if act_cfg is not None else None

    def forward(self, x):
        x = self.conv(x)
        if self.activate is not None:
            x = self.activate(x)
        return x


def build_attention(cfg, default_args=None):
    return build_from_cfg(cfg, ATTENTION, default_args)


def build_feedforward_network(cfg, default_args=None):
    return build_from_cfg(cfg, FEEDFORWARD_NETWORK, default_args)


def build_positional_encoding(cfg, default_args=None):
    return build_from_cfg(cfg, POSITIONAL_ENCODING, default_args)


def build_transformer_layer(cfg, default_args=None):
    return build_from_cfg(cfg, TRANSFORMER_LAYER, default_args)


def build_transformer_layer_sequence(cfg, default_args=None):
    return build_from_cfg(cfg, TRANSFORMER_LAYER_SEQUENCE, default_args)


class MultiheadAttention(BaseModule):
    """mmcv.cnn.bricks.transformer.MultiheadAttention (constructor; forward is overridden
    by the reference's FlattenMHSelfAttention)."""

    def __init__(self, embed_dims, num_heads, attn_drop=0., proj_drop=0.,
                 dropout_layer=dict(type='Dropout', drop_prob=0.), init_cfg=None,
                 batch_first=False, **kwargs):
        super().__init__(init_cfg)
        dropout_layer = dict(dropout_layer)
        if 'dropout' in kwargs:
            attn_drop = kwargs['dropout']
            dropout_layer['drop_prob'] = kwargs.pop('dropout')
        self.embed_dims = embed_dims
        self.num_heads = num_heads
        self.batch_first = batch_first
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, attn_drop, **kwargs)
        self.proj_drop = nn.Dropout(proj_drop)
        self.dropout_layer = build_dropout(dropout_layer) if dropout_layer else nn.Identity()


ATTENTION.register_module(module=MultiheadAttention)


class FFN(BaseModule):
    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2,
                 act_cfg=dict(type='ReLU', inplace=True), ffn_drop=0., dropout_layer=None,
                 add_identity=True, init_cfg=None, **kwargs):
        super().__init__(init_cfg)
        assert num_fcs >= 2
        self.embed_dims = embed_dims
        self.feedforward_channels = feedforward_channels
        self.num_fcs = num_fcs
        self.activate = build_activation_layer(act_cfg)
        layers = []
        in_channels = embed_dims
        for _ in range(num_fcs - 1):
            layers.append(Sequential(nn.Linear(in_channels, feedforward_channels), self.activate,
                                     nn.Dropout(ffn_drop)))
            in_channels = feedforward_channels
        layers.append(nn.Linear(feedforward_channels, embed_dims))
        layers.append(nn.Dropout(ffn_drop))
        self.layers = Sequential(*layers)
        self.dropout_layer = build_dropout(dropout_layer) if dropout_layer else nn.Identity()
        self.add_identity = add_identity

    def forward(self, x, identity=None):
        out = self.layers(x)
        if not self.add_identity:
            return self.dropout_layer(out)
        if identity is None:
            identity = x
        return identity + self.dropout_layer(out)


FEEDFORWARD_NETWORK.register_module(module=FFN)


class BaseTransformerLayer(BaseModule):
    def __init__(self, attn_cfgs=None,
                 ffn_cfgs=dict(type='FFN', embed_dims=256, feedforward_channels=1024, num_fcs=2,
                               ffn_drop=0., act_cfg=dict(type='ReLU', inplace=True)),
                 operation_order=None, norm_cfg=dict(type='LN'), init_cfg=None,
                 batch_first=False, **kwargs):
        deprecated_args = dict(feedforward_channels='feedforward_channels',
                               ffn_dropout='ffn_drop', ffn_num_fcs='num_fcs')
        ffn_cfgs = copy.deepcopy(ffn_cfgs)
        for ori_name, new_name in deprecated_args.items():
            if ori_name in kwargs:
                ffn_cfgs[new_name] = kwargs[ori_name]
        super().__init__(init_cfg)
        self.batch_first = batch_first
        num_attn = operation_order.count('self_attn') + operation_order.count('cross_attn')
        if isinstance(attn_cfgs, dict):
            attn_cfgs = [copy.deepcopy(attn_cfgs) for _ in range(num_attn)]
        else:
            assert num_attn == len(attn_cfgs)
        self.num_attn = num_attn
        self.operation_order = operation_order
        self.norm_cfg = norm_cfg
        self.pre_norm = operation_order[0] == 'norm'
        self.attentions = ModuleList()
        index = 0
        for op in operation_order:
            if op in ('self_attn', 'cross_attn'):
                cfg = dict(attn_cfgs[index])
                if 'batch_first' in cfg:
                    assert self.batch_first == cfg['batch_first']
                else:
                    cfg['batch_first'] = self.batch_first
                attention = build_attention(cfg)
                attention.operation_name = op
                self.attentions.append(attention)
                index += 1
        self.embed_dims = self.attentions[0].embed_dims
        self.ffns = ModuleList()
        num_ffns = operation_order.count('ffn')
        if isinstance(ffn_cfgs, dict):
            ffn_cfgs = [copy.deepcopy(ffn_cfgs) for _ in range(num_ffns)]
        for i in range(num_ffns):
            if 'embed_dims' not in ffn_cfgs[i]:
                ffn_cfgs[i]['embed_dims'] = self.embed_dims
            else:
                assert ffn_cfgs[i]['embed_dims'] == self.embed_dims
            self.ffns.append(build_feedforward_network(ffn_cfgs[i], dict(type='FFN')))
        self.norms = ModuleList()
        for _ in range(operation_order.count('norm')):
            self.norms.append(build_norm_layer(norm_cfg, self.embed_dims)[1])

    def forward(self, query, key=None, value=None, query_pos=None, key_pos=None,
                attn_masks=None, query_key_padding_mask=None, key_padding_mask=None, **kwargs):
        norm_index = attn_index = ffn_index = 0
        identity = query
        if attn_masks is None:
            attn_masks = [None for _ in range(self.num_attn)]
        elif isinstance(attn_masks, torch.Tensor):
            attn_masks = [copy.deepcopy(attn_masks) for _ in range(self.num_attn)]
        else:
            assert len(attn_masks) == self.num_attn
        for layer in self.operation_order:
            if layer == 'self_attn':
                temp_key = temp_value = query
                query = self.attentions[attn_index](
                    query, temp_key, temp_value, identity if self.pre_norm else None,
                    query_pos=query_pos, key_pos=query_pos, attn_mask=attn_masks[attn_index],
                    key_padding_mask=query_key_padding_mask, **kwargs)
                attn_index += 1
                identity = query
            elif layer == 'norm':
                query = self.norms[norm_index](query)
                norm_index += 1
            elif layer == 'cross_attn':
                query = self.attentions[attn_index](
                    query, key, value, identity if self.pre_norm else None,
                    query_pos=query_pos, key_pos=key_pos, attn_mask=attn_masks[attn_index],
                    key_padding_mask=key_padding_mask, **kwargs)
                attn_index += 1
                identity = query
            elif layer == 'ffn':
                query = self.ffns[ffn_index](query, identity if self.pre_norm else None)
                ffn_index += 1
        return query


class TransformerLayerSequence(BaseModule):
    def __init__(self, transformerlayers=None, num_layers=None, init_cfg=None):
        super().__init__(init_cfg)
        if isinstance(transformerlayers, dict):
            transformerlayers = [copy.deepcopy(transformerlayers) for _ in range(num_layers)]
        else:
            assert isinstance(transformerlayers, list) and len(transformerlayers) == num_layers
        self.num_layers = num_layers
        self.layers = ModuleList()
        for i in range(num_layers):
            self.layers.append(build_transformer_layer(transformerlayers[i]))
        self.embed_dims = self.layers[0].embed_dims
        self.pre_norm = self.layers[0].pre_norm

    def forward(self, query, key, value, query_pos=None, key_pos=None, attn_masks=None,
                query_key_padding_mask=None, key_padding_mask=None, **kwargs):
        for layer in self.layers:
            query = layer(query, key, value, query_pos=query_pos, key_pos=key_pos,
                          attn_masks=attn_masks, query_key_padding_mask=query_key_padding_mask,
                          key_padding_mask=key_padding_mask, **kwargs)
        return query


# --------------------------------------------------------------------------- mmdet
def inverse_sigmoid(x, eps=1e-5):
    x = x.clamp(min=0, max=1)
    x1 = x.clamp(min=eps)
    x2 = (1 - x).clamp(min=eps)
    return torch.log(x1 / x2)


def bbox2roi(bbox_list):
    rois_list = []
    for img_id, bboxes in enumerate(bbox_list):
        if bboxes.size(0) > 0:
            img_inds = bboxes.new_full((bboxes.size(0), 1), img_id)
            rois = torch.cat([img_inds, bboxes[:, :4]], dim=-1)
        else:
            rois = bboxes.new_zeros((0, 5))
        rois_list.append(rois)
    return torch.cat(rois_list, 0)


class _LossBag(nn.Module):
    """Losses are not exercised by the forward hot path; only ``use_sigmoid`` is read
    (query_generator.py:90)."""

    def __init__(self, use_sigmoid=False, **kwargs):
        super().__init__()
        self.use_sigmoid = use_sigmoid
        self.cfg = kwargs


for _n in ('FocalLoss', 'L1Loss', 'CrossEntropyLoss', 'SmoothL1Loss'):
    LOSSES.register_module(name=_n, module=_LossBag)


def build_loss(cfg):
    return build_from_cfg(cfg, LOSSES)


class RoIAlign(nn.Module):
    """mmcv.ops.RoIAlign (avg, aligned=True) through torchvision, whose kernel follows the
    same adaptive-grid rule for sampling_ratio<=0."""

    def __init__(self, output_size, spatial_scale=1.0, sampling_ratio=0, pool_mode='avg',
                 aligned=True, use_torchvision=False):
        super().__init__()
        self.output_size = to_2tuple(output_size)
        self.spatial_scale = float(spatial_scale)
        self.sampling_ratio = int(sampling_ratio)
        self.aligned = aligned
        assert pool_mode == 'avg'

    def forward(self, x, rois):
        from torchvision.ops import roi_align
        return roi_align(x, rois, self.output_size, self.spatial_scale,
                         max(self.sampling_ratio, 0), self.aligned)


class SingleRoIExtractor(BaseModule):
    def __init__(self, roi_layer, out_channels, featmap_strides, finest_scale=56, init_cfg=None):
        super().__init__(init_cfg)
        cfg = dict(roi_layer)
        assert cfg.pop('type') == 'RoIAlign'
        self.roi_layers = nn.ModuleList([RoIAlign(spatial_scale=1.0 / s, **cfg)
                                         for s in featmap_strides])
        self.out_channels = out_channels
        self.featmap_strides = featmap_strides

    @property
    def num_inputs(self):
        return len(self.featmap_strides)

    def forward(self, feats, rois, roi_scale_factor=None):
        assert len(feats) == 1
        if len(rois) == 0:
            return feats[0].new_zeros(0, self.out_channels, *self.roi_layers[0].output_size)
        return self.roi_layers[0](feats[0], rois)


ROI_EXTRACTORS.register_module(module=SingleRoIExtractor)


class BaseRoIHead(BaseModule):
    def __init__(self, bbox_roi_extractor=None, bbox_head=None, mask_roi_extractor=None,
                 mask_head=None, shared_head=None, train_cfg=None, test_cfg=None,
                 pretrained=None, init_cfg=None):
        super().__init__(init_cfg)
        self.train_cfg = train_cfg
        self.test_cfg = test_cfg
        if bbox_head is not None:
            self.init_bbox_head(bbox_roi_extractor, bbox_head)
        self.init_assigner_sampler()

    @property
    def with_bbox(self):
        return hasattr(self, 'bbox_head') and self.bbox_head is not None


class BBoxTestMixin:
    pass


class MaskTestMixin:
    pass


class BaseBBoxCoder:
    def __init__(self, **kwargs):
        pass


def multi_apply(func, *args, **kwargs):
    pfunc = functools.partial(func, **kwargs) if kwargs else func
    return tuple(map(list, zip(*map(pfunc, *args))))


def reduce_mean(t):
    return t


def build_linear_layer(cfg, *args, **kwargs):
    cfg = dict(cfg or dict(type='Linear'))
    assert cfg.pop('type') == 'Linear'
    return nn.Linear(*args, **kwargs, **cfg)


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_installed = False


def install():
    """Put the stand-ins into sys.modules, then make the reference's packages importable
    WITHOUT executing mmdet3d_plugin/__init__.py (it pulls in nuscenes-devkit)."""
    global _installed
    if _installed:
        return
    import os
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f'{REFERENCE_ROOT} not present: the reference runs only in the '
                           f'build container')
    _mod('mmcv')
    _mod('mmcv.runner', BaseModule=BaseModule, auto_fp16=_noop_decorator_factory,
         force_fp32=_noop_decorator_factory)
    _mod('mmcv.runner.base_module', BaseModule=BaseModule, ModuleList=ModuleList,
         Sequential=Sequential)
    _mod('mmcv.utils', ConfigDict=ConfigDict, build_from_cfg=build_from_cfg,
         deprecated_api_warning=deprecated_api_warning, to_2tuple=to_2tuple)
    _mod('mmcv.cnn', ConvModule=ConvModule, Conv2d=nn.Conv2d, Linear=nn.Linear,
         build_activation_layer=build_activation_layer, build_norm_layer=build_norm_layer,
         build_conv_layer=build_conv_layer, xavier_init=xavier_init,
         bias_init_with_prob=bias_init_with_prob)
    _mod('mmcv.cnn.bricks')
    _mod('mmcv.cnn.bricks.transformer', BaseTransformerLayer=BaseTransformerLayer,
         TransformerLayerSequence=TransformerLayerSequence, MultiheadAttention=MultiheadAttention,
         FFN=FFN, build_transformer_layer_sequence=build_transformer_layer_sequence,
         build_attention=build_attention, build_positional_encoding=build_positional_encoding,
         POSITIONAL_ENCODING=POSITIONAL_ENCODING)
    _mod('mmcv.cnn.bricks.drop', build_dropout=build_dropout)
    _mod('mmcv.cnn.bricks.registry', ATTENTION=ATTENTION, TRANSFORMER_LAYER=TRANSFORMER_LAYER,
         TRANSFORMER_LAYER_SEQUENCE=TRANSFORMER_LAYER_SEQUENCE)
    _mod('mmdet')
    _mod('mmdet.core', bbox2roi=bbox2roi,
         build_bbox_coder=lambda cfg: build_from_cfg(cfg, BBOX_CODERS),
         build_assigner=lambda cfg: None, build_sampler=lambda cfg, **kw: None,
         multi_apply=multi_apply, reduce_mean=reduce_mean)
    _mod('mmdet.core.bbox', BaseBBoxCoder=BaseBBoxCoder)
    _mod('mmdet.core.bbox.builder', BBOX_CODERS=BBOX_CODERS)
    _mod('mmdet.models')
    _mod('mmdet.models.builder', HEADS=HEADS, build_loss=build_loss,
         build_head=lambda cfg: build_from_cfg(cfg, HEADS),
         build_roi_extractor=lambda cfg: build_from_cfg(cfg, ROI_EXTRACTORS))
    _mod('mmdet.models.utils', build_linear_layer=build_linear_layer,
         build_transformer=lambda cfg: build_from_cfg(cfg, TRANSFORMER))
    _mod('mmdet.models.utils.builder', TRANSFORMER=TRANSFORMER)
    _mod('mmdet.models.utils.transformer', inverse_sigmoid=inverse_sigmoid)
    _mod('mmdet.models.losses', accuracy=lambda *a, **k: None)
    _mod('mmdet.models.roi_heads')
    _mod('mmdet.models.roi_heads.base_roi_head', BaseRoIHead=BaseRoIHead)
    _mod('mmdet.models.roi_heads.test_mixins', BBoxTestMixin=BBoxTestMixin,
         MaskTestMixin=MaskTestMixin)
    _mod('mmdet3d')
    _mod('mmdet3d.models')
    _mod('mmdet3d.models.builder', HEADS=HEADS, build_loss=build_loss)
    # the reference's own packages, as namespace shells pointing into /root/reference
    for pkg in ('mmdet3d_plugin', 'mmdet3d_plugin.core', 'mmdet3d_plugin.core.bbox',
                'mmdet3d_plugin.core.bbox.coders', 'mmdet3d_plugin.models'):
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(REFERENCE_ROOT, *pkg.split('.'))]
        sys.modules[pkg] = m
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        import mmdet3d_plugin.models.utils  # noqa: F401  (registers transformer bricks)
        import mmdet3d_plugin.models.roi_heads  # noqa: F401
        import mmdet3d_plugin.core.bbox.coders.nms_free_coder  # noqa: F401
    _installed = True


def load_reference_config(path):
    """Minimal mmcv.Config.fromfile: ``_base_`` (str|list, relative), recursive dict merge,
    ``_delete_=True`` replacement.  Returns a plain nested dict."""
    import os

    def _load(p):
        ns = {}
        with open(p) as f:
            exec(compile(f.read(), p, 'exec'), ns)
        cfg = {k: v for k, v in ns.items() if not k.startswith('__') and
               not isinstance(v, types.ModuleType) and not callable(v)}
        base = cfg.pop('_base_', [])
        if isinstance(base, str):
            base = [base]
        merged = {}
        for b in base:
            bcfg = _load(os.path.normpath(os.path.join(os.path.dirname(p), b)))
            dup = set(merged) & set(bcfg)
            assert not dup, f'duplicate keys in bases: {dup}'
            merged.update(bcfg)
        return _merge(cfg, merged)

    def _merge(a, b):
        b = copy.deepcopy(b)
        for k, v in a.items():
            if isinstance(v, dict) and k in b and isinstance(b[k], dict) and \
                    not v.get('_delete_', False):
                b[k] = _merge(v, b[k])
            else:
                if isinstance(v, dict):
                    v = {kk: vv for kk, vv in v.items() if kk != '_delete_'}
                b[k] = copy.deepcopy(v)
        return b

    return _load(path)


def build_reference_head(cfg_path):
    """HEADS.build(cfg.model.roi_head) exactly as MV2D.__init__ does (mv2d.py:34-38)."""
    install()
    cfg = load_reference_config(cfg_path)
    roi_head = copy.deepcopy(cfg['model']['roi_head'])
    roi_head.update(train_cfg=None, test_cfg=ConfigDict(cfg['model']['test_cfg']['rcnn']))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        head = build_from_cfg(roi_head, HEADS)
    return head.eval(), cfg


# ---------------------------------------------------------------------------------------------------------------
# Training targets / losses.  The reference's own code for this row --
# HungarianAssigner3D.assign (core/bbox/assigners/hungarian_assigner_3d.py:66-150), BBox3DL1Cost
# (core/bbox/match_costs/match_cost.py:6-26), normalize_bbox (core/bbox/util.py:38-58) and
# CrossAttentionBoxHead.loss_single / get_targets / dn_loss_single (cross_attention_head.py:244-343,379-538) --
# runs unmodified; what it imports from mmdet 2.25.1 is restated below from that version's published behaviour
# (unpinned by any reference test, like the rest of this shim).
MATCH_COST = Registry('match_cost')


@MATCH_COST.register_module()
class FocalLossCost:
    """mmdet/core/bbox/match_costs/match_cost.py FocalLossCost._focal_loss_cost (2.25.1)."""

    def __init__(self, weight=1., alpha=0.25, gamma=2, eps=1e-12, binary_input=False):
        self.weight, self.alpha, self.gamma, self.eps = weight, alpha, gamma, eps

    def __call__(self, cls_pred, gt_labels):
        cls_pred = cls_pred.sigmoid()
        neg_cost = -(1 - cls_pred + self.eps).log() * (1 - self.alpha) * cls_pred.pow(self.gamma)
        pos_cost = -(cls_pred + self.eps).log() * self.alpha * (1 - cls_pred).pow(self.gamma)
        cls_cost = pos_cost[:, gt_labels] - neg_cost[:, gt_labels]
        return cls_cost * self.weight


@MATCH_COST.register_module()
class IoUCost:
    def __init__(self, iou_mode='giou', weight=1.):
        self.weight = weight

    def __call__(self, *a, **k):
        raise NotImplementedError('IoUCost weight is 0 in the reference configs and never called (hungarian_assigner_3d.py:119-127)')


def build_match_cost(cfg):
    return build_from_cfg(cfg, MATCH_COST)


class AssignResult:
    """mmdet/core/bbox/assigners/assign_result.py (fields only)."""

    def __init__(self, num_gts, gt_inds, max_overlaps, labels=None):
        self.num_gts, self.gt_inds, self.max_overlaps, self.labels = num_gts, gt_inds, max_overlaps, labels


class BaseAssigner:
    pass


class SamplingResult:
    """mmdet/core/bbox/samplers/sampling_result.py."""

    def __init__(self, pos_inds, neg_inds, bboxes, gt_bboxes, assign_result, gt_flags):
        self.pos_inds, self.neg_inds = pos_inds, neg_inds
        self.pos_bboxes, self.neg_bboxes = bboxes[pos_inds], bboxes[neg_inds]
        self.num_gts = gt_bboxes.shape[0]
        self.pos_assigned_gt_inds = assign_result.gt_inds[pos_inds] - 1
        if gt_bboxes.numel() == 0:
            self.pos_gt_bboxes = torch.empty_like(gt_bboxes).view(-1, 4)
        else:
            self.pos_gt_bboxes = gt_bboxes[self.pos_assigned_gt_inds.long(), :]


class PseudoSampler:
    """mmdet/core/bbox/samplers/pseudo_sampler.py."""

    def __init__(self, **kwargs):
        pass

    def sample(self, assign_result, bboxes, gt_bboxes, *args, **kwargs):
        pos_inds = torch.nonzero(assign_result.gt_inds > 0, as_tuple=False).squeeze(-1).unique()
        neg_inds = torch.nonzero(assign_result.gt_inds == 0, as_tuple=False).squeeze(-1).unique()
        gt_flags = bboxes.new_zeros(bboxes.shape[0], dtype=torch.uint8)
        return SamplingResult(pos_inds, neg_inds, bboxes, gt_bboxes, assign_result, gt_flags)


def _weight_reduce_loss(loss, weight=None, reduction='mean', avg_factor=None):
    """mmdet/models/losses/utils.py weight_reduce_loss."""
    if weight is not None:
        loss = loss * weight
    if avg_factor is None:
        return loss.mean() if reduction == 'mean' else (loss.sum() if reduction == 'sum' else loss)
    if reduction == 'mean':
        eps = torch.finfo(torch.float32).eps
        return loss.sum() / (avg_factor + eps)
    assert reduction == 'none'
    return loss


class FocalLoss(nn.Module):
    """mmdet/models/losses/focal_loss.py, use_sigmoid=True, the py_sigmoid_focal_loss path (the mmcv CUDA op
    computes the same -alpha_t (1 - p_t)^gamma log p_t)."""

    def __init__(self, use_sigmoid=True, gamma=2.0, alpha=0.25, reduction='mean', loss_weight=1.0, activated=False):
        super().__init__()
        assert use_sigmoid and not activated
        self.use_sigmoid, self.gamma, self.alpha, self.reduction, self.loss_weight = use_sigmoid, gamma, alpha, reduction, loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None):
        num_classes = pred.size(1)
        target = F.one_hot(target, num_classes=num_classes + 1)[:, :num_classes].type_as(pred)
        p = pred.sigmoid()
        pt = (1 - p) * target + p * (1 - target)
        focal_weight = (self.alpha * target + (1 - self.alpha) * (1 - target)) * pt.pow(self.gamma)
        loss = F.binary_cross_entropy_with_logits(pred, target, reduction='none') * focal_weight
        if weight is not None and weight.shape != loss.shape:
            weight = weight.view(-1, 1) if weight.size(0) == loss.size(0) else weight.view(loss.size(0), -1)
        return self.loss_weight * _weight_reduce_loss(loss, weight, self.reduction, avg_factor)


class L1Loss(nn.Module):
    """mmdet/models/losses/smooth_l1_loss.py L1Loss."""

    def __init__(self, reduction='mean', loss_weight=1.0):
        super().__init__()
        self.reduction, self.loss_weight = reduction, loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None):
        if target.numel() == 0:
            return self.loss_weight * pred.sum() * 0
        return self.loss_weight * _weight_reduce_loss(torch.abs(pred - target), weight, self.reduction, avg_factor)


def install_loss_support():
    """Stand-ins the reference's assigner / match-cost modules import, then the reference modules themselves."""
    install()
    import os
    _mod('mmdet.core.bbox.assigners', AssignResult=AssignResult, BaseAssigner=BaseAssigner)
    _mod('mmdet.core.bbox.match_costs', build_match_cost=build_match_cost)
    _mod('mmdet.core.bbox.match_costs.builder', MATCH_COST=MATCH_COST)
    _mod('mmdet.core.bbox.iou_calculators', bbox_overlaps=None)
    sys.modules['mmdet.core.bbox.builder'].BBOX_ASSIGNERS = BBOX_ASSIGNERS
    for pkg in ('mmdet3d_plugin.core.bbox.assigners', 'mmdet3d_plugin.core.bbox.match_costs'):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = [os.path.join(REFERENCE_ROOT, *pkg.split('.'))]
            sys.modules[pkg] = m
    import importlib
    mc = importlib.import_module('mmdet3d_plugin.core.bbox.match_costs.match_cost')      # registers BBox3DL1Cost
    asg = importlib.import_module('mmdet3d_plugin.core.bbox.assigners.hungarian_assigner_3d')
    return asg.HungarianAssigner3D, mc.BBox3DL1Cost
