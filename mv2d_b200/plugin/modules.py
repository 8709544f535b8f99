"""Parameter-holding modules with the reference's names/kwargs; forwards call libmv2d_b200.

Every class cites the reference class it stands in for.  Training-only paths (losses, DN query
preparation, assigners) are out of scope this round and raise.
"""
import copy

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..engine import HotPath
from ..registry import (ATTENTION, BBOX_ASSIGNERS, BBOX_CODERS, BBOX_SAMPLERS, MATCH_COST, DETECTORS, HEADS, LOSSES, NECKS, POSITIONAL_ENCODING,
                        ROI_EXTRACTORS, TRANSFORMER, TRANSFORMER_LAYER, TRANSFORMER_LAYER_SEQUENCE,
                        build_from_cfg)


# ------------------------------------------------------------------ small config holders
class _LossCfg(nn.Module):
    """Loss configs are carried (``use_sigmoid`` is read by QueryGenerator, query_generator.py:90);
    the losses themselves belong to the training rows that are out of scope."""

    def __init__(self, use_sigmoid=False, **kwargs):
        super().__init__()
        self.use_sigmoid = use_sigmoid
        self.cfg = dict(kwargs)


for _n in ('FocalLoss', 'L1Loss', 'CrossEntropyLoss', 'SmoothL1Loss'):
    LOSSES.register_module(name=_n, module=_LossCfg)


@ROI_EXTRACTORS.register_module()
class SingleRoIExtractor(nn.Module):
    """mmdet SingleRoIExtractor with one RoIAlign level (exp/...single_frame...:46-50).  Holds the
    configuration; the pooling itself is ``roi_align_tokens_kernel`` inside ``mv2d_roi_align_qg``."""

    def __init__(self, roi_layer, out_channels, featmap_strides, **kwargs):
        super().__init__()
        assert roi_layer['type'] == 'RoIAlign' and len(featmap_strides) == 1
        assert roi_layer.get('output_size', 7) == 7, 'libmv2d_b200 is specialised for 7x7 RoI tokens'
        assert roi_layer.get('sampling_ratio', -1) <= 0, 'adaptive sampling grid only (sampling_ratio=-1)'
        self.roi_layer = dict(roi_layer)
        self.out_channels = out_channels
        self.featmap_strides = list(featmap_strides)

    @property
    def num_inputs(self):
        return len(self.featmap_strides)


@BBOX_CODERS.register_module()
class NMSFreeCoder:
    """core/bbox/coders/nms_free_coder.py:18-123; decode runs ``mv2d_nms_free_decode``."""

    def __init__(self, pc_range, post_center_range=None, max_num=100, score_threshold=None, num_classes=10):
        assert score_threshold is None, 'score_threshold is not used by the MV2D configs'
        self.pc_range = pc_range
        self.post_center_range = post_center_range
        self.max_num = max_num
        self.num_classes = num_classes


@POSITIONAL_ENCODING.register_module()
class SinePositionalEncoding3D(nn.Module):
    """models/utils/positional_encoding.py:14-106 (normalize=True path is what PE uses)."""

    def __init__(self, num_feats, temperature=10000, normalize=False, scale=2 * np.pi, eps=1e-6, offset=0.,
                 init_cfg=None):
        super().__init__()
        assert num_feats == 128 and normalize and temperature == 10000 and offset == 0.
        self.num_feats, self.temperature, self.normalize = num_feats, temperature, normalize
        self.scale, self.eps, self.offset = scale, eps, offset


class SELayer(nn.Module):
    """models/utils/pe.py:36-48 (parameters only)."""

    def __init__(self, channels):
        super().__init__()
        self.conv_reduce = nn.Conv2d(channels, channels, 1, bias=True)
        self.conv_expand = nn.Conv2d(channels, channels, 1, bias=True)


class PE(nn.Module):
    """models/utils/pe.py:51-169.  ``forward(mlvl_feats, img_metas)`` runs ``mv2d_pe3d``."""

    def __init__(self, positional_encoding, strides, position_range, depth_num, depth_start=1, LID=True,
                 embed_dims=256, with_fpe=False, adapt_pos3d=True, no_sin_enc=False):
        super().__init__()
        assert LID and with_fpe and adapt_pos3d and not no_sin_enc and embed_dims == 256 and len(strides) == 1, \
            'libmv2d_b200 implements the configuration the MV2D configs use (LID, with_fpe, adapt_pos3d)'
        self.strides, self.position_range = strides, position_range
        self.depth_num, self.depth_start, self.embed_dims = depth_num, depth_start, embed_dims
        self.position_encoder = nn.Sequential(nn.Conv2d(3 * depth_num, embed_dims * 4, 1), nn.ReLU(),
                                              nn.Conv2d(embed_dims * 4, embed_dims, 1))
        self.adapt_pos3d = nn.Sequential(nn.Conv2d(embed_dims * 3 // 2, embed_dims * 4, 1), nn.ReLU(),
                                         nn.Conv2d(embed_dims * 4, embed_dims, 1))
        self.positional_encoding = build_from_cfg(positional_encoding, POSITIONAL_ENCODING)
        self.fpe = SELayer(embed_dims)
        self._owner = None   # the head that owns the engine

    def forward(self, mlvl_feats, img_metas):
        assert self._owner is not None, 'PE must be owned by an MV2D head'
        eng = self._owner.engine()
        out = []
        for x in mlvl_feats:
            feat, feat_tf32 = eng.to_nhwc(x.float().contiguous())
            i2l, _ = eng.geom_prep(eng._upload_cams(img_metas))
            pe, _ = eng.pe3d(feat, i2l, img_metas, feat_tf32)
            out.append(pe.permute(0, 3, 1, 2))   # NCHW view, as the reference returns
        return out


class QueryGenerator(nn.Module):
    """roi_heads/utils/query_generator.py:19-405, the configuration the MV2D configs use:
    conv3x3 -> avg-pool -> FC(1024) -> cat intrinsics(16) -> MLP(512,256) -> fc_center(3)."""

    def __init__(self, return_cfg=dict(), wich_cp=False, with_avg_pool=True, with_center=True, roi_feat_size=7,
                 in_channels=256, num_classes=10, extra_encoding=None, num_shared_convs=1, num_shared_fcs=1,
                 conv_out_channels=256, fc_out_channels=1024, loss_cls=None, **kwargs):
        super().__init__()
        extra_encoding = extra_encoding or dict(num_layers=2, feat_channels=[512, 256],
                                                features=[dict(type='intrinsic', in_channels=16)])
        assert with_avg_pool and with_center and num_shared_convs == 1 and num_shared_fcs == 1
        assert roi_feat_size == 7 and in_channels == 256 and conv_out_channels == 256 and fc_out_channels == 1024
        assert list(extra_encoding['feat_channels']) == [512, 256]
        assert [f['type'] for f in extra_encoding['features']] == ['intrinsic']
        conv = nn.Module()
        conv.conv = nn.Conv2d(in_channels, conv_out_channels, 3, padding=1)
        self.shared_convs = nn.ModuleList([conv])
        self.shared_fcs = nn.ModuleList([nn.Linear(conv_out_channels, fc_out_channels)])
        self.extra_enc = nn.Sequential(nn.Linear(fc_out_channels + 16, 512), nn.ReLU(inplace=True),
                                       nn.Linear(512, 256), nn.ReLU(inplace=True))
        self.fc_center = nn.Linear(256, 3)
        self.loss_cls = loss_cls
        self.return_cfg = dict(return_cfg or {})
        self._owner = None

    @torch.no_grad()
    def forward(self, x, intrinsics, extrinsics, extra_feats=dict()):
        """utils/query_generator.py:343-350: x [N,256,7,7] RoI features, intrinsics [N,4,4] (per-RoI K' of
        MV2DHead.get_box_params), extrinsics [N,4,4], extra_feats['intrinsic'] [N,16] -> (center_lidar [N,3],
        return_feats).  One call into ``mv2d_roi_align_qg`` (phase 3: conv3x3 as a 3xTF32 tcgen05 implicit GEMM, FC chain,
        fc_center, center2lidar)."""
        assert self._owner is not None, 'QueryGenerator must be owned by an MV2D head'
        out = self._owner.engine().query_generator(x, intrinsics, extrinsics, extra_feats['intrinsic'])
        return_feats = dict()
        if self.return_cfg.get('enc', False):
            return_feats['enc'] = out['enc']
        return out['center_lidar'], return_feats


class BoxCorrelation(nn.Module):
    """roi_heads/utils/box_correlation.py:11-398 (``topk_matched`` mode).  The gen_* methods run
    ``mv2d_box_corr`` and return the reference's tensors."""

    def __init__(self, sample_size=4, num_depth=8, depth_start=0.5, depth_end=70, correlation_mode=None,
                 LID=True, expand_stride=0, force_cpu=False):
        super().__init__()
        assert LID and correlation_mode and correlation_mode.startswith('topk_matched'), correlation_mode
        info = correlation_mode.split(':')
        self.topk, self.iou_thr, self.ratio = int(info[1]), float(info[2]), float(info[3])
        self.sample_size, self.num_depth = sample_size, num_depth
        self.depth_start, self.depth_end, self.expand_stride = depth_start, depth_end, expand_stride
        self.correlation_mode = correlation_mode
        self._owner = None

    def engine_cfg(self):
        return dict(sample_size=self.sample_size, corr_num_depth=self.num_depth, corr_depth_start=self.depth_start,
                    corr_depth_end=float(self.depth_end), topk=self.topk, iou_thr=self.iou_thr, ratio=self.ratio,
                    expand_stride=self.expand_stride)

    def _run(self, rois, num_proposals_per_img, img_metas, h, w):
        eng = self._owner.engine()
        V = len(img_metas)
        starts = np.concatenate([[0], np.cumsum(num_proposals_per_img)]).astype(np.int32)
        roi_start = torch.from_numpy(starts).to(eng.device)
        _, trans = eng.geom_prep(eng._upload_cams(img_metas))
        return eng.box_corr(rois.float().contiguous(), roi_start, trans, rois.shape[0], V, img_metas, h, w)

    @torch.no_grad()
    def gen_box_roi_correlation(self, rois, num_proposals_per_img, img_metas):
        """-> (corr int64 [N,M] padded with 0, mask bool [N,M]) as box_correlation.py:165-193."""
        if rois.numel() == 0:
            return rois.new_zeros((0, 0), dtype=torch.int64), rois.new_zeros((0, 0), dtype=torch.bool)
        out = self._run(rois, num_proposals_per_img, img_metas, 1, 1)
        cnt = out['match_cnt'].long()
        M = int(cnt.max())
        mask = torch.arange(M, device=rois.device)[None] < cnt[:, None]
        corr = out['match'][:, :M].long() * mask
        return corr, mask

    @torch.no_grad()
    def gen_box_correlation(self, rois, num_proposals_per_img, img_metas, feat, stride):
        """-> bool [N,V,h,w] as box_correlation.py:95-162 (padding mask already removed)."""
        _, _, h, w = feat.shape
        out = self._run(rois, num_proposals_per_img, img_metas, h, w)
        V, N = len(img_metas), rois.shape[0]
        words = out['keymask']
        bits = (words[:, :, None] >> torch.arange(32, device=words.device, dtype=torch.int32)) & 1
        return bits.view(N, -1)[:, :V * h * w].bool().view(N, V, h, w)


@ATTENTION.register_module()
class FlattenMHSelfAttention(nn.Module):
    """models/utils/petr_transformer.py:314-370 (over mmcv MultiheadAttention): parameters
    ``attn.in_proj_weight/bias``, ``attn.out_proj``."""

    def __init__(self, embed_dims, num_heads, attn_drop=0., proj_drop=0., dropout_layer=None, init_cfg=None,
                 batch_first=False, **kwargs):
        super().__init__()
        assert embed_dims == 256 and num_heads == 8
        self.embed_dims, self.num_heads, self.batch_first = embed_dims, num_heads, batch_first
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, kwargs.get('dropout', attn_drop))


@ATTENTION.register_module()
class PETRMultiheadAttention(FlattenMHSelfAttention):
    """models/utils/petr_transformer.py:373-513."""


class _FFN(nn.Module):
    """mmcv FFN parameter layout: layers.0.0 = Linear(256,2048), layers.1 = Linear(2048,256)."""

    def __init__(self, embed_dims, feedforward_channels, ffn_drop=0.):
        super().__init__()
        self.layers = nn.Sequential(
            nn.Sequential(nn.Linear(embed_dims, feedforward_channels), nn.ReLU(inplace=True), nn.Dropout(ffn_drop)),
            nn.Linear(feedforward_channels, embed_dims), nn.Dropout(ffn_drop))


@TRANSFORMER_LAYER.register_module()
class PETRTransformerDecoderLayer(nn.Module):
    """models/utils/petr_transformer.py:194-311 over mmcv BaseTransformerLayer."""

    def __init__(self, attn_cfgs, feedforward_channels, ffn_dropout=0.0, operation_order=None, act_cfg=None,
                 norm_cfg=None, ffn_num_fcs=2, with_cp=True, **kwargs):
        super().__init__()
        assert tuple(operation_order) == ('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm'), \
            'libmv2d_b200 implements the operation order of the MV2D configs'
        assert feedforward_channels == 2048 and ffn_num_fcs == 2
        self.operation_order = tuple(operation_order)
        self.attentions = nn.ModuleList([build_from_cfg(c, ATTENTION) for c in attn_cfgs])
        self.embed_dims = self.attentions[0].embed_dims
        self.ffns = nn.ModuleList([_FFN(self.embed_dims, feedforward_channels, ffn_dropout)])
        self.norms = nn.ModuleList([nn.LayerNorm(self.embed_dims) for _ in range(3)])
        self.use_checkpoint = with_cp
        self.pre_norm = False


@TRANSFORMER_LAYER_SEQUENCE.register_module()
class PETRTransformerDecoder(nn.Module):
    """models/utils/petr_transformer.py:546-593."""

    def __init__(self, transformerlayers=None, num_layers=None, post_norm_cfg=dict(type='LN'),
                 return_intermediate=False, init_cfg=None):
        super().__init__()
        assert return_intermediate and post_norm_cfg is not None
        self.num_layers = num_layers
        self.layers = nn.ModuleList([build_from_cfg(copy.deepcopy(transformerlayers), TRANSFORMER_LAYER)
                                     for _ in range(num_layers)])
        self.embed_dims = self.layers[0].embed_dims
        self.post_norm = nn.LayerNorm(self.embed_dims)
        self.return_intermediate = return_intermediate


@TRANSFORMER.register_module()
class MV2DTransformer(nn.Module):
    """roi_heads/bbox_heads/cross_attention_head.py:22-49."""

    def __init__(self, encoder=None, decoder=None, init_cfg=None, cross=False):
        super().__init__()
        assert encoder is None
        self.decoder = build_from_cfg(decoder, TRANSFORMER_LAYER_SEQUENCE)
        self.embed_dims = self.decoder.embed_dims
        self._owner = None      # the MV2D head whose engine runs the decoder

    @torch.no_grad()
    def forward(self, x, mask, query_embed, pos_embed, attn_mask=None, cross_attn_mask=None, **kwargs):
        """cross_attention_head.py:22-49: x / pos_embed [bs,n,c,h,w], mask [bs,n,h,w], query_embed [bs,n_query,c],
        attn_mask [n_query,n_query], cross_attn_mask [n_query,n,h,w] -> (out_dec [num_layers,bs,n_query,c], memory).
        The six decoder layers run in ``mv2d_decoder`` (``HotPath.transformer`` converts the dense interface: gathered
        RoI blocks become match lists, the dense bool mask a bit-packed key mask)."""
        assert self._owner is not None, 'MV2DTransformer must be owned by an MV2D head'
        outs, _, _ = self._owner.engine().transformer(x, mask, query_embed, pos_embed, None, attn_mask, cross_attn_mask)
        bs, nq = query_embed.shape[0], query_embed.shape[1]
        return outs.view(outs.shape[0], bs, nq, -1), x


class _Cfg:
    """Plain holder of constructor kwargs (assigner / match costs / sampler: their arithmetic runs inside
    mv2d_loss, csrc/loss.cu)."""

    def __init__(self, **kwargs):
        self.cfg = dict(kwargs)
        self.weight = kwargs.get('weight', 1.0)


@BBOX_ASSIGNERS.register_module()
class HungarianAssigner3D(_Cfg):
    """core/bbox/assigners/hungarian_assigner_3d.py:29-64: cls_cost / reg_cost / iou_cost configs, pc_range."""

    def __init__(self, cls_cost=None, reg_cost=None, iou_cost=None, pc_range=None):
        super().__init__(pc_range=pc_range)
        self.cls_cost = build_from_cfg(cls_cost or dict(type='FocalLossCost', weight=1.0), MATCH_COST)
        self.reg_cost = build_from_cfg(reg_cost or dict(type='BBox3DL1Cost', weight=1.0), MATCH_COST)
        self.iou_cost = build_from_cfg(iou_cost or dict(type='IoUCost', weight=0.0), MATCH_COST)


for _n in ('FocalLossCost', 'BBox3DL1Cost', 'IoUCost', 'ClassificationCost', 'BBoxL1Cost'):
    MATCH_COST.register_module(name=_n, module=type(_n, (_Cfg,), {}))
BBOX_SAMPLERS.register_module(name='PseudoSampler', module=type('PseudoSampler', (_Cfg,), {}))


@HEADS.register_module()
class CrossAttentionBoxHead(nn.Module):
    """roi_heads/bbox_heads/cross_attention_head.py:87-242 (forward + get_bboxes)."""

    def __init__(self, num_classes, transformer, pc_range, embed_dims=256, num_reg_fcs=2,
                 group_reg_dims=(2, 2, 1, 1, 2, 2), use_reg_layer=False, pre_embed=False, loss_cls=None,
                 loss_bbox=None, bbox_coder=None, sync_cls_avg_factor=False, train_cfg=None, test_cfg=None,
                 **kwargs):
        super().__init__()
        assert num_classes == 10 and embed_dims == 256 and num_reg_fcs == 2 and not use_reg_layer and not pre_embed
        assert sum(group_reg_dims) == 10
        self.loss_cls = build_from_cfg(loss_cls, LOSSES) if loss_cls else _LossCfg()
        self.loss_bbox = build_from_cfg(loss_bbox, LOSSES) if loss_bbox else _LossCfg()
        self.transformer = build_from_cfg(transformer, TRANSFORMER)
        self.pc_range, self.embed_dims, self.num_classes = pc_range, embed_dims, num_classes
        self.num_pred = transformer['decoder']['num_layers']
        self.query_embedding = nn.Sequential(nn.Linear(embed_dims * 3 // 2, embed_dims), nn.ReLU(),
                                             nn.Linear(embed_dims, embed_dims))
        cls_branch = []
        for _ in range(num_reg_fcs):
            cls_branch += [nn.Linear(embed_dims, embed_dims), nn.LayerNorm(embed_dims), nn.ReLU(inplace=True)]
        cls_branch.append(nn.Linear(embed_dims, num_classes))
        reg_branch = []
        for _ in range(num_reg_fcs):
            reg_branch += [nn.Linear(embed_dims, embed_dims), nn.ReLU()]
        reg_branch.append(nn.Linear(embed_dims, sum(group_reg_dims)))
        self.cls_branches = nn.ModuleList([copy.deepcopy(nn.Sequential(*cls_branch)) for _ in range(self.num_pred)])
        self.reg_branches = nn.ModuleList([copy.deepcopy(nn.Sequential(*reg_branch)) for _ in range(self.num_pred)])
        self.bbox_coder = build_from_cfg(bbox_coder, BBOX_CODERS) if bbox_coder else None
        code_weights = kwargs.get('code_weights', [1.0] * 8 + [0.2, 0.2])[:kwargs.get('code_size', 10)]
        self.code_weights = nn.Parameter(torch.tensor(code_weights), requires_grad=False)
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self.assigner = build_from_cfg(train_cfg['assigner'], BBOX_ASSIGNERS) if (train_cfg and train_cfg.get('assigner')) else None
        self._owner = None

    @torch.no_grad()
    def position_embedding(self, query_pos):
        """cross_attention_head.py:199-200: query_embedding(pos2posemb3d(reference points)) -- ``mv2d_query_embedding``."""
        q = self._owner.engine().query_embedding(query_pos)
        return q.view(*query_pos.shape[:-1], self.embed_dims)

    @torch.no_grad()
    def forward(self, reference_points, x, masks, pos_embed, attn_mask=None, cross_attn_mask=None, force_fp32=False,
                query_embeds=None, return_query_feats=False, **kwargs):
        """cross_attention_head.py:202-242: reference_points [bs,n_query,3] (normalised), x / pos_embed [bs,n,c,h,w],
        masks [bs,n,h,w] -> (all_cls_scores, all_bbox_preds) [num_layers,bs,n_query,10]: query embedding, the decoder,
        the per-layer cls / reg branches and the reference-point refinement in one ``mv2d_decoder`` call."""
        eng = self._owner.engine()
        query_embeds = self.position_embedding(reference_points)
        outs, cls, box = eng.transformer(x, masks, query_embeds, pos_embed, reference_points, attn_mask, cross_attn_mask)
        L, bs, nq = cls.shape[0], reference_points.shape[0], reference_points.shape[1]
        cls, box = cls.view(L, bs, nq, -1), box.view(L, bs, nq, -1)
        if return_query_feats:
            return cls, box, outs[-1].view(bs, nq, -1)
        return cls, box

    def get_bboxes(self, preds_dicts, img_metas, rescale=False):
        """cross_attention_head.py:356-377: NMS-free decode (top-k, denormalise, range filter) + z-shift."""
        eng = self._owner.engine()
        ret = []
        for i, (cls, box) in enumerate(zip(preds_dicts['cls_scores'], preds_dicts['bbox_preds'])):
            boxes, scores, labels = eng.decode(cls, box, self.bbox_coder.max_num)
            bt = img_metas[i].get('box_type_3d') if i < len(img_metas) else None
            ret.append([bt(boxes, boxes.size(-1)) if bt is not None else boxes, scores, labels])
        return ret

    # ---- next row f3: targets + losses, forward values (no backward kernels yet) -- csrc/loss.cu
    def _loss_kwargs(self):
        kw = dict(code_weights=[float(x) for x in self.code_weights.tolist()])
        lc, lb = self.loss_cls.cfg, self.loss_bbox.cfg
        kw.update(cls_loss_weight=lc.get('loss_weight', 1.0), focal_gamma=lc.get('gamma', 2.0), focal_alpha=lc.get('alpha', 0.25),
                  bbox_loss_weight=lb.get('loss_weight', 1.0))
        if self.assigner is not None:
            kw.update(cls_cost_weight=self.assigner.cls_cost.weight, reg_cost_weight=self.assigner.reg_cost.weight)
        return kw

    @staticmethod
    def _gt_tensor(gt):
        """LiDARInstance3DBoxes-like (gravity_center, tensor) or a plain [G,9] tensor (cross_attention_head.py:450-452)."""
        if hasattr(gt, 'gravity_center'):
            return torch.cat((gt.gravity_center, gt.tensor[:, 3:]), dim=1)
        return gt

    @torch.no_grad()
    def loss(self, gt_bboxes_3d_list, gt_labels_3d_list, preds_dicts, cls_reg_targets=None, gt_bboxes_ignore=None):
        """cross_attention_head.py:436-462 for one sample: {'loss_cls', 'loss_bbox'} of the given layer's predictions."""
        assert gt_bboxes_ignore is None and len(gt_bboxes_3d_list) == 1
        cls, box = preds_dicts['cls_scores'][0], preds_dicts['bbox_preds'][0]
        out = self._owner.engine().loss(cls.view(1, -1, 10), box.view(1, -1, 10), self._gt_tensor(gt_bboxes_3d_list[0]),
                                        gt_labels_3d_list[0], **self._loss_kwargs())
        return dict(loss_cls=out['loss_cls'][0], loss_bbox=out['loss_bbox'][0])

    @torch.no_grad()
    def dn_loss_single(self, cls_scores, bbox_preds, known_bboxs, known_labels, num_total_pos, pc_range, split,
                       neg_bbox_loss=False):
        """cross_attention_head.py:475-538.  known_bboxs [pad,9] repeats the G ground-truth boxes (query i <- box i % G)."""
        eng = self._owner.engine()
        pad = known_labels.numel()
        G = max(pad // max(eng.cfg['denoise_scalar'], 1), 1)
        assert pad % G == 0 and torch.equal(known_bboxs[:G].repeat(pad // G, 1), known_bboxs), 'known_bboxs must tile the GT boxes'
        dummy = torch.zeros((1, 0, 10), device=cls_scores.device)
        saved = eng.cfg['denoise_split']
        eng.cfg['denoise_split'] = split
        try:
            out = eng.loss(dummy, dummy, known_bboxs[:G], torch.zeros(G, dtype=torch.int32), dn_cls=cls_scores.view(1, -1, 10),
                           dn_box=bbox_preds.view(1, -1, 10), dn_labels=known_labels, neg_bbox_loss=neg_bbox_loss,
                           **self._loss_kwargs())
        finally:
            eng.cfg['denoise_split'] = saved
        return out['dn_loss_cls'][0], out['dn_loss_bbox'][0]


# ------------------------------------------------------------------ RoI heads
@HEADS.register_module()
class MV2DHead(nn.Module):
    """roi_heads/mv2d_head.py:19-267.  ``simple_test`` / ``_bbox_forward`` run the five C-ABI stages
    (the dense-feature-map decoder of the base class is what MV2DTHead uses)."""
    MODE = 'T'

    def __init__(self, bbox_roi_extractor, bbox_head, query_generator, pe, box_correlation, pc_range,
                 intrins_feat_scale=0.1, feat_lvl=0, force_fp32=False, train_cfg=None, test_cfg=None, **kwargs):
        super().__init__()
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self.bbox_roi_extractor = build_from_cfg(bbox_roi_extractor, ROI_EXTRACTORS)
        bbox_head = dict(bbox_head, train_cfg=train_cfg, test_cfg=test_cfg)
        self.bbox_head = build_from_cfg(bbox_head, HEADS)
        self.query_generator = QueryGenerator(**dict(query_generator, loss_cls=self.bbox_head.loss_cls))
        self.position_encoding = PE(**pe)
        self.box_corr_module = BoxCorrelation(**box_correlation)
        self.pc_range, self.intrins_feat_scale, self.feat_lvl, self.force_fp32 = pc_range, intrins_feat_scale, feat_lvl, force_fp32
        self.roi_size = [7, 7]
        self.stage_loss_weights = train_cfg.get('stage_loss_weights') if train_cfg else None
        for m in (self.position_encoding, self.box_corr_module, self.bbox_head, self.query_generator, self.bbox_head.transformer):
            object.__setattr__(m, '_owner', self)
        self._engine = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: setattr(module, '_engine', None))

    @property
    def strides(self):
        return self.position_encoding.strides

    @property
    def num_classes(self):
        return self.bbox_head.num_classes

    @property
    def with_bbox(self):
        return True

    def engine(self):
        """The packed-weight engine, rebuilt after load_state_dict()."""
        if self._engine is None:
            pe = self.position_encoding
            cfg = dict(pc_range=list(self.pc_range), position_range=list(pe.position_range), depth_num=pe.depth_num,
                       depth_start=float(pe.depth_start), stride=pe.strides[0],
                       intrins_feat_scale=self.intrins_feat_scale,
                       num_views_per_frame=getattr(self, 'num_views', 6))
            cfg.update(self.box_corr_module.engine_cfg())
            cfg.update(self._denoise_cfg())
            dev = next(self.parameters()).device
            self._engine = HotPath(self.state_dict(), mode=self.MODE, device=dev, **cfg)
        return self._engine

    def _denoise_cfg(self):
        return {}

    def _dn_inputs(self, img_metas):
        """Denoising-query inputs of the training-mode forward, or None (eval / use_denoise off)."""
        return None

    def _results(self, out, dn=None):
        L = out['cls_scores'].shape[0]
        N = out['N']
        mask_dict = None
        if dn is not None:
            # the dict MV2DSHead.prepare_for_dn returns (mv2d_s_head.py:106-113) + output_known_lbs_bboxes (:193-196)
            G, pad, scalar = dn['gt_labels'].numel(), out['dn_pad'], self.denoise_scalar
            dev = out['cls_scores'].device
            idx = torch.arange(G, device=dev)
            mask_dict = dict(
                known_indice=idx.repeat(scalar), batch_idx=torch.zeros(G, dtype=torch.long, device=dev),
                map_known_indice=torch.cat([idx + G * i for i in range(scalar)]) if scalar else idx[:0],
                known_lbs_bboxes=(out['dn_labels'].long(), dn['gt_boxes'].to(dev).repeat(scalar, 1)),
                know_idx=[torch.ones_like(dn['gt_labels']).to(dev)], pad_size=pad)
            if pad > 0:
                mask_dict['output_known_lbs_bboxes'] = (out['dn_cls_scores'][:, None], out['dn_bbox_preds'][:, None])
        return dict(cls_scores=[out['cls_scores'][l] for l in range(L)],
                    bbox_preds=[out['bbox_preds'][l] for l in range(L)],
                    bbox_feats=out['tok_feat'].view(N, 7, 7, 256).permute(0, 3, 1, 2), return_feats=dict(),
                    intrinsics=out['roi_intrinsics'].view(N, 4, 4), extrinsics=None, rois=out['rois'],
                    dn_mask_dict=mask_dict)

    @torch.no_grad()
    def _bbox_forward(self, x, proposal_list, img_metas):
        """x: list with the P4 feature [V,256,h,w] (the PE is computed inside the fused path; a
        reference-style [V,512,h,w] feat||pe tensor is accepted and its first half is used).
        In training mode with ``use_denoise`` the GT boxes of ``img_metas[0]`` spawn the denoising queries
        (mv2d_s_head.py:158-180, mv2d_t_head.py:91-98); forward only -- no autograd through the kernels."""
        feat = x[self.feat_lvl]
        if feat.shape[1] == 2 * self.position_encoding.embed_dims:
            feat = feat[:, :self.position_encoding.embed_dims]
        dn = self._dn_inputs(img_metas)
        # a channels-last map presented in the reference's [V,C,h,w] shape (what the FPN neck module returns after
        # .permute(0, 3, 1, 2)) is consumed in place: no transpose back and forth
        cl = feat.is_cuda and feat.dim() == 4 and feat.stride(1) == 1 and feat.permute(0, 2, 3, 1).is_contiguous()
        if cl:
            out = self.engine().forward(feat.permute(0, 2, 3, 1), proposal_list, img_metas, feat_is_nhwc=True, dn=dn)
        else:
            out = self.engine().forward(feat, proposal_list, img_metas, dn=dn)
        return self._results(out, dn)

    @torch.no_grad()
    def simple_test(self, x, proposal_list, img_metas, rescale=False):
        assert len(img_metas) // img_metas[0]['num_views'] == 1   # mv2d_head.py:251
        res = self._bbox_forward(x, proposal_list, img_metas)
        return self.bbox_head.get_bboxes({'cls_scores': [res['cls_scores'][-1]], 'bbox_preds': [res['bbox_preds'][-1]]},
                                         img_metas)

    def trainer(self):
        """The flat-buffer trainer over THIS module's parameters (mv2d_b200/train.py).  On first use every hot-path
        Parameter's storage is re-pointed at its slice of the trainer's flat parameter buffer, so a torch optimizer
        stepping the module's Parameters updates the buffer the kernels read -- no copy in either direction."""
        if getattr(self, '_trainer', None) is None:
            from ..train import HotPathTrainer
            pe = self.position_encoding
            dev = next(self.parameters()).device
            kw = self.bbox_head._loss_kwargs()
            ecfg = dict(pc_range=list(self.pc_range), position_range=list(pe.position_range), depth_num=pe.depth_num,
                        depth_start=float(pe.depth_start), stride=pe.strides[0], intrins_feat_scale=self.intrins_feat_scale)
            ecfg.update(self.box_corr_module.engine_cfg())
            ecfg.update(self._denoise_cfg())
            ecfg['num_views_per_frame'] = getattr(self, 'num_views', 6)
            tr = HotPathTrainer(self.state_dict(), device=dev, stage_loss_weights=self.stage_loss_weights,
                                pc_range=list(self.pc_range), engine_cfg=ecfg, mode=self.MODE,
                                use_denoise=bool(getattr(self, 'use_denoise', False)),
                                denoise_weight=float(getattr(self, 'denoise_weight', 1.0)),
                                neg_bbox_loss=bool(getattr(self, 'neg_bbox_loss', False)), **kw)
            params = dict(self.named_parameters())
            self._train_params = []
            for name in tr.table:
                prm = params[name]
                prm.data = tr.sd_view(name)                 # a view of the flat buffer in the Parameter's own shape
                self._train_params.append((name, prm))
            object.__setattr__(self, '_trainer', tr)
        return self._trainer

    def forward_train(self, x, img_metas, proposal_list, gt_bboxes, gt_labels, gt_bboxes_3d, gt_labels_3d,
                      ori_gt_bboxes_3d, ori_gt_labels_3d, attr_labels=None, gt_bboxes_ignore=None, gt_masks=None, **kwargs):
        """MV2DSHead.forward_train (roi_heads/mv2d_s_head.py:236-307): the loss dict ``l{i}.loss_cls`` / ``l{i}.loss_bbox``
        (each times stage_loss_weights[i]) of one sample.  The values come from the CUDA forward; their SUM carries
        the autograd edge: calling ``.backward()`` on it (what mmdet's ``_parse_losses`` + the runner do) runs
        ``mv2d_decoder_train_backward`` + ``mv2d_front_train_backward``, accumulates into every hot-path
        Parameter's ``.grad`` and hands d loss / d feat back to autograd, so the torch backbone trains through it."""
        assert len(img_metas) // img_metas[0]['num_views'] == 1      # mv2d_s_head.py:250
        if getattr(self, 'use_denoise', False) and self.MODE != 'T':
            raise NotImplementedError('denoising queries are trained with the two-frame head (the reference trains MV2D-S '
                                      'without them: configs/mv2d/exp/*single_frame*:44); their forward exists for both heads')
        feat = x[self.feat_lvl]
        if feat.shape[1] == 2 * self.position_encoding.embed_dims:
            feat = feat[:, :self.position_encoding.embed_dims]
        boxes = self.bbox_head._gt_tensor(ori_gt_bboxes_3d[0])
        tr = self.trainer()
        self._engine = None           # the packed inference weights go stale as soon as the optimizer steps
        total, loss_cls, loss_bbox, dn_cls, dn_bbox = _TrainStep.apply(feat, self, [p[:, :6] for p in proposal_list], img_metas, boxes,
                                                                       ori_gt_labels_3d[0], *[prm for _, prm in self._train_params])
        w = tr.stage_loss_weights
        losses = {}
        for i in range(tr.L):
            if tr.mode == 'T' and tr.use_denoise and boxes.shape[0] > 0:      # mv2d_s_head.py:288-298
                losses[f'l{i}.dn_loss_cls'] = dn_cls[i] * tr.denoise_weight * w[i]
                losses[f'l{i}.dn_loss_bbox'] = dn_bbox[i] * tr.denoise_weight * w[i]
            losses[f'l{i}.loss_cls'] = loss_cls[i] * w[i]
            losses[f'l{i}.loss_bbox'] = loss_bbox[i] * w[i]
        # the entries above are plain values; the autograd edge rides on the first one (the sum of the dict is `total`)
        losses['l0.loss_cls'] = losses['l0.loss_cls'] + (total - total.detach())
        return losses


class _TrainStep(torch.autograd.Function):
    """One sample's hot-path training step as an autograd node: forward = the CUDA forward with saved activations +
    targets / losses, backward = the CUDA backward.  EVERY hot-path Parameter is an input of the node and gets its slice
    of the flat gradient buffer back from ``backward``, so autograd's AccumulateGrad nodes -- and with them the hooks of
    (MM)DistributedDataParallel's bucketed all-reduce -- see all of them (find_unused_parameters=False works)."""

    @staticmethod
    def forward(ctx, feat, head, proposal_list, img_metas, gt_boxes, gt_labels, *params):
        tr = head.trainer()
        tr.zero_grad()
        out = tr.forward(feat.detach(), proposal_list, img_metas, gt_boxes, gt_labels)
        ctx.head, ctx.num_pos = head, out['num_pos']
        loss_cls, loss_bbox = out['loss_cls'].clone(), out['loss_bbox'].clone()
        dn_cls = out['dn_loss_cls'].clone() if 'dn_loss_cls' in out else torch.zeros_like(loss_cls)
        dn_bbox = out['dn_loss_bbox'].clone() if 'dn_loss_bbox' in out else torch.zeros_like(loss_cls)
        ctx.mark_non_differentiable(loss_cls, loss_bbox, dn_cls, dn_bbox)
        return out['loss'].clone(), loss_cls, loss_bbox, dn_cls, dn_bbox

    @staticmethod
    def backward(ctx, g_total, g_cls, g_bbox, g_dn_cls, g_dn_bbox):
        import torch.distributed as dist
        from ..train import DecoderTrainer
        head = ctx.head
        tr = head.trainer()
        factor = None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            # the reference's loss_bbox normaliser is the mean positive count over ranks (cross_attention_head.py:419-420)
            factor = DecoderTrainer.global_bbox_avg_factor([ctx.num_pos])
        gin = tr.backward(factor)
        if float(g_total) != 1.0:      # a scaled loss (gradient accumulation, AMP scaler): scale what this step produced
            tr.grads.mul_(g_total)
        grads = [tr.grad(name) for name, _ in head._train_params]      # views of the flat buffer; autograd copies them
        return (gin['d_feat'] * g_total, None, None, None, None, None, *grads)


@HEADS.register_module()
class MV2DSHead(MV2DHead):
    """roi_heads/mv2d_s_head.py:19-305 (eval forward: RoI-token keys of self + epipolar matches)."""
    MODE = 'S'

    def __init__(self, use_denoise=False, neg_bbox_loss=False, denoise_scalar=10, denoise_noise_scale=1.0,
                 denoise_noise_trans=0.0, denoise_weight=1.0, denoise_split=0.75, **kwargs):
        super().__init__(**kwargs)
        self.use_denoise, self.neg_bbox_loss = use_denoise, neg_bbox_loss
        self.denoise_scalar, self.denoise_noise_scale, self.denoise_noise_trans = denoise_scalar, denoise_noise_scale, denoise_noise_trans
        self.denoise_weight, self.denoise_split = denoise_weight, denoise_split

    def _denoise_cfg(self):
        return dict(denoise_scalar=self.denoise_scalar, denoise_noise_scale=self.denoise_noise_scale,
                    denoise_noise_trans=self.denoise_noise_trans, denoise_split=self.denoise_split,
                    num_classes=self.num_classes)

    def _dn_inputs(self, img_metas, rand=None):
        if not (self.training and self.use_denoise):
            return None
        gt = img_metas[0]['gt_bboxes_3d']       # mv2d_s_head.py:41-44: gravity centre + tensor[:, 3:]
        boxes = torch.cat((gt.gravity_center, gt.tensor[:, 3:]), dim=1)
        return dict(gt_boxes=boxes, gt_labels=img_metas[0]['gt_labels_3d'], rand=rand)


@HEADS.register_module()
class MV2DTHead(MV2DSHead):
    """roi_heads/mv2d_t_head.py:19-142 (two frames: dense feature-map keys + per-query mask, velocity / dt)."""
    MODE = 'T'

    def __init__(self, num_views=6, **kwargs):
        self.num_views = num_views
        super().__init__(**kwargs)


# ------------------------------------------------------------------ detector shells
class _ConvHolder(nn.Module):
    """mmcv ConvModule without norm / activation: parameter names ``conv.weight`` / ``conv.bias``."""

    def __init__(self, cin, cout, k):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, padding=k // 2)


@NECKS.register_module()
class FPN(nn.Module):
    """The MV2D neck (next row f4): mmdet FPN restricted to what configs/mv2d/exp/*.py:32-39 build -- ONE level
    (start_level = end_level, num_outs = 1), 256 -> 256: ``lateral_convs.0.conv`` (1x1) and ``fpn_convs.0.conv`` (3x3).
    Holds the parameters under mmdet's names; ``forward`` runs mv2d_fpn_neck and returns the channels-last map the
    roi_head consumes (``.permute(0, 3, 1, 2)`` is the reference layout as a view)."""

    def __init__(self, in_channels, out_channels, num_outs, start_level=0, end_level=-1, **kwargs):
        super().__init__()
        end = len(in_channels) - 1 if end_level == -1 else end_level
        assert num_outs == 1 and end == start_level and in_channels[start_level] == 256 and out_channels == 256, \
            'libmv2d_b200 implements the one-level 256 -> 256 neck of the MV2D configs'
        self.start_level = start_level
        self.lateral_convs = nn.ModuleList([_ConvHolder(256, 256, 1)])
        self.fpn_convs = nn.ModuleList([_ConvHolder(256, 256, 3)])
        self._packed = None

    def forward(self, inputs, engine):
        from ..pack import PackedNeck
        x = inputs[self.start_level] if isinstance(inputs, (list, tuple)) else inputs
        lat, out = self.lateral_convs[0].conv, self.fpn_convs[0].conv
        if torch.is_grad_enabled() and (x.requires_grad or lat.weight.requires_grad or out.weight.requires_grad):
            # training: the neck's backward is torch's -- two convolutions under autograd, so that
            # d loss / d feat of the hot path reaches the neck's parameters and the backbone.  Channels-last VIEW of the
            # NCHW result: ``.permute(0, 3, 1, 2)`` of it is the contiguous reference layout again.
            y = F.conv2d(F.conv2d(x, lat.weight, lat.bias), out.weight, out.bias, padding=1)
            return y.permute(0, 2, 3, 1), None
        if self._packed is None or self._packed[0] != x.device or self._packed[2] != self._version_key():
            self._packed = (x.device, PackedNeck(self.state_dict(), x.device), self._version_key())
        return engine.neck(x, self._packed[1])

    def _version_key(self):
        """Re-pack after an optimizer step changed the parameters in place."""
        return tuple(int(p._version) for p in self.parameters())


@DETECTORS.register_module()
class MV2D(nn.Module):
    """detectors/mv2d.py:18-293, thin shell: the 2D detector + FPN stay torch (north star) and are
    injected as ``base_detector`` (any callable ``(img[V,3,H,W], img_metas) -> (feat[V,256,h,w], detections)``);
    ``simple_test`` hands their outputs to the roi_head exactly as mv2d.py:251-261 does."""

    def __init__(self, base_detector=None, neck=None, roi_head=None, train_cfg=None, test_cfg=None,
                 use_grid_mask=None, **kwargs):
        super().__init__()
        roi_head = dict(roi_head, train_cfg=(train_cfg or {}).get('rcnn') if train_cfg else None,
                        test_cfg=(test_cfg or {}).get('rcnn') if test_cfg else None)
        self.roi_head = build_from_cfg(roi_head, HEADS)
        self.base_detector = base_detector if callable(base_detector) else None
        self.neck_cfg, self.train_cfg, self.test_cfg = neck, train_cfg, test_cfg
        self.neck = build_from_cfg(neck, NECKS) if isinstance(neck, dict) and neck.get('type') in NECKS else None

    @property
    def with_neck(self):
        return self.neck is not None

    def process_detector_feat(self, detector_feat):
        """detectors/mv2d.py:122-127.  Returns (feat, is_channels_last): with the neck the feature map comes out of
        mv2d_fpn_neck channels-last, which is what the roi_head's kernels read."""
        if self.with_neck:
            feat, _ = self.neck(detector_feat, self.roi_head.engine())
            return feat, True
        x = detector_feat[0] if isinstance(detector_feat, (list, tuple)) else detector_feat
        return x, False

    def process_2d_detections(self, results, device):
        """detectors/mv2d.py:60-86 (next row f2): per view, the 2D detector's per-class box arrays
        [n_c, 5] -> one [n, 6] tensor (x1, y1, x2, y2, score, label), boxes smaller than
        ``detection_proposal.min_bbox_size`` dropped.  Accepts numpy arrays (the reference's bbox2result
        format) or tensors that already live on the device (no host round trip in that case)."""
        cfg = self.train_cfg if self.train_cfg is not None else (self.test_cfg or {})
        min_size = (cfg.get('detection_proposal') or {}).get('min_bbox_size', 0)
        out = []
        for res in results:
            per_cls = []
            for label_id, boxes in enumerate(res):
                b = torch.as_tensor(boxes, dtype=torch.float32).reshape(-1, 5).to(device)
                per_cls.append(torch.cat([b, b.new_full((b.shape[0], 1), float(label_id))], dim=1))
            det = torch.cat(per_cls, dim=0) if per_cls else torch.zeros((0, 6), device=device)
            if min_size > 0:
                wh = det[:, 2:4] - det[:, 0:2]
                det = det[(wh >= min_size).all(dim=1)]
            out.append(det)
        return out

    @staticmethod
    def process_2d_gt(gt_bboxes, gt_labels, device):
        """detectors/mv2d.py:47-58: per view [m,4] boxes + [m] labels -> [m,6] (x1, y1, x2, y2, 1, label)."""
        return [torch.cat([b.to(device).float(), torch.ones((len(l), 1), device=device), l.to(device).float()[:, None]], dim=-1)
                for b, l in zip(gt_bboxes, gt_labels)]

    def complement_2d_gt(self, detections, gts, thr=0.35):
        """detectors/mv2d.py:104-117 for ONE view, on the device (``mv2d_handoff_2d``)."""
        min_size = ((self.train_cfg or {}).get('detection_proposal') or {}).get('min_bbox_size', 0)
        if len(gts) == 0:
            return detections
        if len(detections) == 0:
            return gts
        # the detections come size-filtered from process_2d_detections, so the kernel's filter leaves them untouched
        return self.roi_head.engine().handoff_2d([detections], [gts], min_size, thr)[0]

    def handoff(self, results, gt_bboxes=None, gt_labels=None, device=None):
        """process_2d_detections + complement_2d_gt for all views in ONE device call (detectors/mv2d.py:196-202):
        ``results`` = per view either the detector's per-class [n_c,5] arrays or an [n,6] tensor."""
        device = device or next(self.roi_head.parameters()).device
        cfg = self.train_cfg if self.train_cfg is not None else (self.test_cfg or {})
        min_size = (cfg.get('detection_proposal') or {}).get('min_bbox_size', 0)
        thr = (self.train_cfg or {}).get('complement_2d_gt', -1) if gt_bboxes is not None else -1
        dets = []
        for res in results:
            if torch.is_tensor(res):
                dets.append(res.to(device).float().reshape(-1, 6))
                continue
            per_cls = [torch.cat([torch.as_tensor(b, dtype=torch.float32).reshape(-1, 5).to(device),
                                  torch.full((len(b), 1), float(i), device=device)], dim=1) for i, b in enumerate(res)]
            dets.append(torch.cat(per_cls, 0) if per_cls else torch.zeros((0, 6), device=device))
        gts = self.process_2d_gt(gt_bboxes, gt_labels, device) if thr > 0 else None
        return self.roi_head.engine().handoff_2d(dets, gts, min_size, thr)

    def forward_train(self, img, img_metas, gt_bboxes_2d, gt_labels_2d, gt_bboxes_2d_to_3d, gt_bboxes_3d, gt_labels_3d,
                      attr_labels=None, gt_bboxes_ignore=None, detector_out=None):
        """detectors/mv2d.py:129-213, the shell around the hot path: per-view metas and ground truth, the 2D detector's
        losses and detections (``base_detector``: a torch callable returning (feat, results, losses); or pass
        ``detector_out`` = that triple), the hand-off on the device, the neck, ``roi_head.forward_train``.
        One sample per call, as the reference asserts (:143); batches of samples go through
        ``HotPath.forward_batch`` / ``TrainStep``."""
        batch_size, num_views = img.shape[0], img.shape[1]
        assert batch_size == 1, 'only support batch_size 1 now'          # mv2d.py:143
        img = img.view(batch_size * num_views, *img.shape[2:])
        ori_img_metas, ori_gt_bboxes_3d, ori_gt_labels_3d = img_metas, gt_bboxes_3d, gt_labels_3d
        views_meta, gt_bboxes, gt_labels, v_boxes_3d, v_labels_3d = [], [], [], [], []
        for i in range(batch_size):
            for j in range(num_views):                                   # mv2d.py:157-166
                m = dict(num_views=num_views)
                for k, v in ori_img_metas[i].items():
                    m[k] = v[j] if isinstance(v, list) else (v[:3] if k == 'ori_shape' else v)
                views_meta.append(m)
            lab3d = ori_gt_labels_3d[i]
            box3d = ori_gt_bboxes_3d[i]
            for j in range(num_views):                                   # mv2d.py:170-174
                ids = gt_bboxes_2d_to_3d[i][j].unique()
                sel = ids[ids > -1].long()
                v_boxes_3d.append(box3d[sel])
                v_labels_3d.append(lab3d[sel])
            gt_bboxes.extend(gt_bboxes_2d[i])
            gt_labels.extend(gt_labels_2d[i])
        losses = dict()
        if detector_out is None:
            assert self.base_detector is not None, 'inject a torch 2D detector or pass detector_out'
            detector_out = self.base_detector(img, views_meta, gt_bboxes=gt_bboxes, gt_labels=gt_labels)
        detector_feat, results, det_losses = detector_out
        for k, v in (det_losses or {}).items():
            losses['det_' + k] = v
        detections = self.handoff(results, gt_bboxes, gt_labels, device=img.device)
        if self.with_neck:
            cl, _ = self.process_detector_feat(detector_feat)
            feat = cl.permute(0, 3, 1, 2)
        else:
            feat = detector_feat[0] if isinstance(detector_feat, (list, tuple)) else detector_feat
        losses.update(self.roi_head.forward_train([feat], views_meta, detections, gt_bboxes, gt_labels, v_boxes_3d, v_labels_3d,
                                                  ori_gt_bboxes_3d, ori_gt_labels_3d, attr_labels, None))
        return losses

    @torch.no_grad()
    def simple_test(self, img, img_metas, detections=None, feat=None):
        if feat is None or detections is None:
            assert self.base_detector is not None, 'inject a torch 2D detector or pass feat/detections'
            feat, detections = self.base_detector(img, img_metas)
            if self.with_neck:      # mv2d.py:122-127: the detector hands over its FPN outputs, the neck makes P4'
                cl, _ = self.process_detector_feat(feat)
                feat = cl.permute(0, 3, 1, 2)
        outs = self.roi_head.simple_test([feat], detections, img_metas)
        # scene-level tail, mv2d.py:262-293: box3d_multiclass_nms (score_thr / nms_thr / max_per_scene from
        # test_cfg.rcnn) and bbox3d2result (results leave the device)
        cfg = (self.test_cfg or {}).get('rcnn') or {}
        nms = cfg.get('nms') or {}
        bt = img_metas[0].get('box_type_3d')
        results = []
        for boxes, scores, labels in outs:
            raw = boxes.tensor if hasattr(boxes, 'tensor') else boxes
            b, s, l = self.roi_head.engine().scene_nms(raw, scores, labels, cfg.get('score_thr', 0.0), nms.get('nms_thr', 1.0),
                                                       cfg.get('max_per_scene', 300))
            b = b.cpu()
            results.append(dict(pts_bbox=dict(boxes_3d=bt(b, b.shape[-1]) if bt is not None else b, scores_3d=s.cpu(),
                                              labels_3d=l.cpu())))
        return results


@DETECTORS.register_module()
class MV2DT(MV2D):
    """detectors/mv2d_t.py:17-136 (same shell; the T head consumes 12 views)."""
