"""Training step of the MV2D-S decoder slice on one GPU (BASELINE configs[3]).

``DecoderTrainer`` owns ONE flat fp32 parameter buffer and ONE flat gradient buffer for the parameters of
``bbox_head`` (query embedding, six decoder layers, post norm, cls / reg branches) laid out as
``mv2d_train_param_info`` says, and drives ``mv2d_decoder_train_forward`` / ``mv2d_decoder_train_backward``
(csrc/train.cu): forward with saved activations, Hungarian targets and losses, then the gradients of every
parameter and of the slice's inputs (reference points, RoI key tokens, RoI value tokens).  What the reference
does with torch autograd over ``CrossAttentionBoxHead.forward`` + ``loss`` (cross_attention_head.py:199-242,
379-434) as ``MV2DSHead.forward_train`` sums them (mv2d_s_head.py:262-307).

Data parallelism: samples are sharded over ranks, the only collective is the sum all-reduce of the
flat gradient buffer (``all_reduce_grads``: one NCCL call), followed by a fused AdamW pass (``adamw_step``).

There is no CPU fallback: the library must be present and the tensors on a CUDA device.
"""
import ctypes as C

import torch

from . import lib
from .dist import all_reduce_sum

_BH = 'bbox_head.'
GLOBAL_NAMES = [_BH + n for n in (
    'query_embedding.0.weight', 'query_embedding.0.bias', 'query_embedding.2.weight', 'query_embedding.2.bias',
    'transformer.decoder.post_norm.weight', 'transformer.decoder.post_norm.bias')]
_DEC = 'transformer.decoder.layers.{l}.'
LAYER_NAMES = [_BH + n for n in (
    _DEC + 'attentions.0.attn.in_proj_weight', _DEC + 'attentions.0.attn.in_proj_bias',
    _DEC + 'attentions.0.attn.out_proj.weight', _DEC + 'attentions.0.attn.out_proj.bias',
    _DEC + 'attentions.1.attn.in_proj_weight', _DEC + 'attentions.1.attn.in_proj_bias',
    _DEC + 'attentions.1.attn.out_proj.weight', _DEC + 'attentions.1.attn.out_proj.bias',
    _DEC + 'ffns.0.layers.0.0.weight', _DEC + 'ffns.0.layers.0.0.bias',
    _DEC + 'ffns.0.layers.1.weight', _DEC + 'ffns.0.layers.1.bias',
    _DEC + 'norms.0.weight', _DEC + 'norms.0.bias', _DEC + 'norms.1.weight', _DEC + 'norms.1.bias',
    _DEC + 'norms.2.weight', _DEC + 'norms.2.bias',
    'cls_branches.{l}.0.weight', 'cls_branches.{l}.0.bias', 'cls_branches.{l}.1.weight', 'cls_branches.{l}.1.bias',
    'cls_branches.{l}.3.weight', 'cls_branches.{l}.3.bias', 'cls_branches.{l}.4.weight', 'cls_branches.{l}.4.bias',
    'cls_branches.{l}.6.weight', 'cls_branches.{l}.6.bias',
    'reg_branches.{l}.0.weight', 'reg_branches.{l}.0.bias', 'reg_branches.{l}.2.weight', 'reg_branches.{l}.2.bias',
    'reg_branches.{l}.4.weight', 'reg_branches.{l}.4.bias')]
_PE, _QG = 'position_encoding.', 'query_generator.'
CONV_W = _QG + 'shared_convs.0.conv.weight'    # state_dict [c_out, c_in, ky, kx]; flat buffer [c_out, ky, kx, c_in]
FRONT_NAMES = [
    _PE + 'position_encoder.0.weight', _PE + 'position_encoder.0.bias', _PE + 'position_encoder.2.weight', _PE + 'position_encoder.2.bias',
    _PE + 'adapt_pos3d.0.weight', _PE + 'adapt_pos3d.0.bias', _PE + 'adapt_pos3d.2.weight', _PE + 'adapt_pos3d.2.bias',
    _PE + 'fpe.conv_reduce.weight', _PE + 'fpe.conv_reduce.bias', _PE + 'fpe.conv_expand.weight', _PE + 'fpe.conv_expand.bias',
    CONV_W, _QG + 'shared_convs.0.conv.bias', _QG + 'shared_fcs.0.weight', _QG + 'shared_fcs.0.bias',
    _QG + 'extra_enc.0.weight', _QG + 'extra_enc.0.bias', _QG + 'extra_enc.2.weight', _QG + 'extra_enc.2.bias',
    _QG + 'fc_center.weight', _QG + 'fc_center.bias']
# MV2D_TRAIN_GLOBAL_TENSORS / MV2D_TRAIN_LAYER_TENSORS / MV2D_TRAIN_FRONT_TENSORS
assert len(GLOBAL_NAMES) == 6 and len(LAYER_NAMES) == 34 and len(FRONT_NAMES) == 22

LOSS_DEFAULTS = dict(   # configs/mv2d/exp/mv2d_r50_frcnn_single_frame_roi_1408x512_ep72.py:87-95,132-137
    code_weights=[1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.5, 1.5, 2.0, 2.0],
    cls_loss_weight=2.0, focal_gamma=2.0, focal_alpha=0.25, bbox_loss_weight=0.25,
    cls_cost_weight=2.0, reg_cost_weight=0.25)
PC_RANGE = [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]


def set_tensor_cores(on):
    """1 (default) = the GPU-filling contractions of the training step run as 3xTF32 on the tcgen05 tensor cores,
    0 = fp32 FFMA everywhere (``mv2d_train_set_tensor_cores``).  Returns the previous mode."""
    return int(lib.load().mv2d_train_set_tensor_cores(int(bool(on))))


def param_table(num_layers):
    """state_dict name (relative to ``roi_head.``) -> (offset, numel) in the flat buffers, from the library itself."""
    h = lib.load()
    table = {}
    off, num = C.c_longlong(), C.c_longlong()
    names = list(GLOBAL_NAMES) + [n.format(l=l) for l in range(num_layers) for n in LAYER_NAMES] + list(FRONT_NAMES)
    for tid, name in enumerate(names):
        lib.check(h.mv2d_train_param_info(num_layers, tid, C.byref(off), C.byref(num)), 'mv2d_train_param_info')
        table[name] = (off.value, num.value)
    return table, int(h.mv2d_train_param_total(num_layers))


class DecoderTrainer:
    """Flat parameter / gradient buffers of the whole hot path plus the training step of the decoder slice
    (``forward`` / ``backward`` on given slice inputs).  ``HotPathTrainer`` below adds the front end.

    state_dict: the reference's names (with or without the ``roi_head.`` prefix).  stage_loss_weights: one weight per
    decoder layer (train_cfg.rcnn.stage_loss_weights).  loss_cfg: overrides of ``LOSS_DEFAULTS``."""

    def __init__(self, state_dict, device='cuda', num_layers=None, stage_loss_weights=None, pc_range=None, **loss_cfg):
        self.lib = lib.load()
        # the flat buffers (layout, state_dict round trip, gradient all-reduce) also work on a CPU device, which is
        # what the gloo tests use; forward / backward / adamw_step need CUDA -- there is no CPU fallback
        self.device = torch.device(device)
        sd = {k[len('roi_head.'):] if k.startswith('roi_head.') else k: v for k, v in state_dict.items()}
        if num_layers is None:
            num_layers = 1 + max(int(k.split('.')[4]) for k in sd if k.startswith(_BH + 'transformer.decoder.layers.'))
        assert 1 <= num_layers <= lib.MAX_LAYERS
        self.L = num_layers
        self.table, self.total = param_table(num_layers)
        self.shapes = {}
        self.params = torch.zeros(self.total, dtype=torch.float32, device=self.device)
        self.grads = torch.zeros_like(self.params)
        self.exp_avg = self.exp_avg_sq = None
        self.step_count = 0
        self.load_state_dict(sd)
        # train_cfg.rcnn.stage_loss_weights (configs/mv2d/exp/*single_frame*.py:131: 0.1 for each of the six layers)
        self.stage_loss_weights = list(stage_loss_weights) if stage_loss_weights is not None else [0.1] * num_layers
        assert len(self.stage_loss_weights) == num_layers
        self.loss_cfg = dict(LOSS_DEFAULTS, **loss_cfg)
        self.pc_range = list(pc_range or PC_RANGE)
        dim_t = torch.arange(128, dtype=torch.float32)
        self.dim_t = (10000 ** (2 * (dim_t // 2) / 128)).to(self.device)       # utils/pe.py:24-25
        self._ws = None
        self._keep = None
        self._p = None

    # ------------------------------------------------------------------ parameters
    def view(self, name, buf=None):
        off, num = self.table[name]
        t = (self.params if buf is None else buf)[off:off + num]
        return t.view(self.shapes[name]) if name in self.shapes else t

    def load_state_dict(self, sd):
        for name in self.table:
            src = sd[name].detach().to(torch.float32)
            if name == CONV_W:
                src = src.permute(0, 2, 3, 1).contiguous()      # K order of the im2col GEMM: (ky, kx, c_in)
            self.shapes[name] = tuple(src.shape)
            assert src.numel() == self.table[name][1], f'{name}: {tuple(src.shape)} does not match the library layout'
            self.view(name).copy_(src.to(self.device))

    def _sd_layout(self, name, t):
        return t.permute(0, 3, 1, 2).contiguous() if name == CONV_W else t.clone()

    def state_dict(self):
        """The reference's names and shapes (relative to ``roi_head.``)."""
        return {n: self._sd_layout(n, self.view(n)) for n in self.table}

    def sd_view(self, name, buf=None):
        """A view of one tensor's slice of the flat parameter (or another flat) buffer in the reference's state_dict
        shape (a permuted, non-contiguous view for the re-laid-out conv weight)."""
        t = self.view(name, buf)
        return t.permute(0, 3, 1, 2) if name == CONV_W else t

    def grad(self, name):
        """Gradient of one tensor in the reference's state_dict layout: a view into the flat gradient buffer."""
        return self.sd_view(name, self.grads)

    def named_grads(self):
        return {n: self.grad(n) for n in self.table}

    def zero_grad(self):
        self.grads.zero_()

    # ------------------------------------------------------------------ one sample
    def _stage_weights(self):
        """stage_loss_weights on the device, uploaded once (a per-call torch.tensor(list, device=...) is a pageable
        host-to-device copy that blocks the host until the stream has drained: no running ahead, no lane overlap)."""
        key = tuple(float(x) for x in self.stage_loss_weights)
        if getattr(self, '_stage_w_key', None) != key:
            self._stage_w_key, self._stage_w_dev = key, torch.tensor(key, dtype=torch.float32, device=self.device)
        return self._stage_w_dev

    def _need_cuda(self):
        if self.device.type != 'cuda' or not torch.cuda.is_available():
            raise RuntimeError('mv2d_b200.DecoderTrainer: the training step runs on a CUDA device only (no CPU fallback)')

    def _params(self, ref, tok_kin, tok_mem, match, match_cnt, gt_boxes, gt_labels):
        dev = self.device
        N, M = match.shape
        assert ref.shape == (N, 3) and tok_kin.shape == (N, 49, 256) and tok_mem.shape == (N, 49, 256)
        f32 = dict(device=dev, dtype=torch.float32)
        ref, tok_kin, tok_mem = (t.to(**f32).contiguous() for t in (ref, tok_kin, tok_mem))
        match = match.to(dev, torch.int32).contiguous()
        match_cnt = match_cnt.to(dev, torch.int32).contiguous()
        gt = gt_boxes.to(**f32).contiguous().view(-1, 9)
        lab = gt_labels.to(dev).to(torch.int32).contiguous()
        G, L = gt.shape[0], self.L
        out = dict(cls_scores=torch.empty((L, N, 10), **f32), bbox_preds=torch.empty((L, N, 10), **f32),
                   assigned=torch.empty((L, N), device=dev, dtype=torch.int32), losses=torch.empty((L, 4), **f32),
                   num_pos=torch.empty((L,), **f32),
                   d_ref=torch.empty((N, 3), **f32), d_tok_kin=torch.empty((N, 49, 256), **f32),
                   d_tok_mem=torch.empty((N, 49, 256), **f32))
        ws_bytes = int(self.lib.mv2d_decoder_train_workspace_bytes(N, L, M, G))
        if self._ws is None or self._ws.numel() * 4 < ws_bytes:
            self._ws = torch.empty(ws_bytes // 4 + 64, **f32)
        p = lib.TrainParams()
        p.N, p.L, p.max_match, p.G, p.num_classes = N, L, M, G, 10
        p.pc_range = (C.c_float * 6)(*self.pc_range)
        c = self.loss_cfg
        for k in ('cls_cost_weight', 'reg_cost_weight', 'cls_loss_weight', 'bbox_loss_weight', 'focal_alpha', 'focal_gamma'):
            setattr(p, k, c[k])
        p.code_weights = (C.c_float * 10)(*c['code_weights'])
        p.stage_loss_weights = (C.c_float * lib.MAX_LAYERS)(*(self.stage_loss_weights + [0.0] * (lib.MAX_LAYERS - L)))
        p.params, p.grads, p.dim_t = self.params.data_ptr(), self.grads.data_ptr(), self.dim_t.data_ptr()
        p.ref, p.tok_kin, p.tok_mem = ref.data_ptr(), tok_kin.data_ptr(), tok_mem.data_ptr()
        p.match, p.match_cnt = match.data_ptr(), match_cnt.data_ptr()
        p.gt_boxes, p.gt_labels = (gt.data_ptr(), lab.data_ptr()) if G > 0 else (None, None)
        for k, t in out.items():
            setattr(p, k, t.data_ptr())
        p.workspace, p.workspace_bytes = self._ws.data_ptr(), ws_bytes
        self._keep = (ref, tok_kin, tok_mem, match, match_cnt, gt, lab)
        return p, out

    @torch.no_grad()
    def forward(self, ref, tok_kin, tok_mem, match, match_cnt, gt_boxes, gt_labels):
        """Training-mode forward of one sample.  ref [N,3], tok_kin / tok_mem [N,49,256], match [N,M] int,
        match_cnt [N] (``HotPath`` stage tensors), gt_boxes [G,9], gt_labels [G].  Returns cls_scores / bbox_preds
        [L,N,10], assigned [L,N], loss_cls / loss_bbox [L] (unweighted) and ``loss`` = the weighted total
        sum_l stage_loss_weights[l] * (loss_cls[l] + loss_bbox[l])."""
        self._need_cuda()
        self._p, out = self._params(ref, tok_kin, tok_mem, match, match_cnt, gt_boxes, gt_labels)
        lib.check(self.lib.mv2d_decoder_train_forward(C.byref(self._p), lib.stream_ptr()), 'mv2d_decoder_train_forward')
        self._out = out
        w = self._stage_weights()
        res = dict(cls_scores=out['cls_scores'], bbox_preds=out['bbox_preds'], assigned=out['assigned'],
                   loss_cls=out['losses'][:, 0], loss_bbox=out['losses'][:, 1], num_pos=out['num_pos'])
        res['loss'] = (w * (out['losses'][:, 0] + out['losses'][:, 1])).sum()
        return res

    @torch.no_grad()
    def backward(self, bbox_avg_factor=None):
        """Gradient of the last forward's ``loss``: accumulates into the flat gradient buffer and returns the input
        gradients d_ref [N,3], d_tok_kin [N,49,256] (gradient w.r.t. the key input feat + pe tokens) and d_tok_mem
        [N,49,256] (gradient w.r.t. the value input).
        bbox_avg_factor [L] (device): the reference divides loss_bbox by clamp(reduce_mean(num_total_pos), min=1) taken
        across ranks (cross_attention_head.py:419-420); ``global_bbox_avg_factor`` makes it from the forwards' ``num_pos``.
        With it the gradient (and the stored loss_bbox values) use that factor instead of the sample's own count."""
        self._need_cuda()
        assert self._p is not None, 'backward() needs a forward() first'
        if bbox_avg_factor is not None:
            self._baf = bbox_avg_factor.to(self.device, torch.float32).contiguous()
            self._p.bbox_avg_factor = self._baf.data_ptr()
        else:
            self._p.bbox_avg_factor = None
        lib.check(self.lib.mv2d_decoder_train_backward(C.byref(self._p), lib.stream_ptr()), 'mv2d_decoder_train_backward')
        out = self._out
        return dict(d_ref=out['d_ref'], d_tok_kin=out['d_tok_kin'], d_tok_mem=out['d_tok_mem'])

    @staticmethod
    @torch.no_grad()
    def global_bbox_avg_factor(num_pos_list, group=None):
        """clamp(mean over ALL samples of the global batch of num_pos, min=1) per decoder layer: the local sums are
        all-reduced in one tiny collective (the reference: reduce_mean over ranks with one sample per rank,
        cross_attention_head.py:419-420).  num_pos_list: the ``num_pos`` [L] of this rank's forwards."""
        s = torch.stack(list(num_pos_list), 0).sum(0)
        n = torch.tensor([float(len(num_pos_list))], device=s.device)
        buf = torch.cat([s, n])
        all_reduce_sum(buf, group)
        return torch.clamp(buf[:-1] / buf[-1], min=1.0)

    # ------------------------------------------------------------------ data parallel + optimizer
    def all_reduce_grads(self, group=None):
        """The one collective of the data-parallel step: sum of the flat gradient buffer over ranks (NCCL on GPUs)."""
        return all_reduce_sum(self.grads, group)

    @torch.no_grad()
    def clip_grad_norm_(self, max_norm=35.0, grad_scale=1.0):
        """mmcv OptimizerHook's grad_clip (configs/mv2d/exp/*.py:172-175: max_norm 35, L2) over the hot-path parameters:
        with the flat buffer the total norm is ONE reduction.  ``grad_scale`` is applied first (1 / (world * samples)
        after a sum all-reduce), then the buffer is scaled in place by min(1, max_norm / (norm + 1e-6)) -- torch's
        clip_grad_norm_ semantics, no host sync.  Returns the norm before clipping (a device scalar).  Plain torch
        ops on the flat buffer (plumbing); pass grad_scale=1 to ``adamw_step`` afterwards."""
        if grad_scale != 1.0:
            self.grads.mul_(grad_scale)
        norm = torch.linalg.vector_norm(self.grads, dtype=torch.float64).float()      # fp64 accumulation over 14 M entries
        self.grads.mul_(torch.clamp(max_norm / (norm + 1e-6), max=1.0))
        return norm

    @torch.no_grad()
    def adamw_step(self, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, grad_scale=1.0):
        """torch.optim.AdamW semantics over the flat buffers in one launch (exp configs: AdamW lr 2e-4, wd 0.01)."""
        self._need_cuda()
        if self.exp_avg is None:
            self.exp_avg, self.exp_avg_sq = torch.zeros_like(self.params), torch.zeros_like(self.params)
        self.step_count += 1
        lib.check(self.lib.mv2d_adamw_step(self.params.data_ptr(), self.grads.data_ptr(), self.exp_avg.data_ptr(),
                                           self.exp_avg_sq.data_ptr(), self.total, lr, betas[0], betas[1], eps, weight_decay,
                                           self.step_count, grad_scale, lib.stream_ptr()), 'mv2d_adamw_step')


class HotPathTrainer(DecoderTrainer):
    """The whole hot-path training step of MV2D-S on one GPU: position encoding -> RoIAlign -> query generator ->
    decoder -> targets / losses, and back to the gradient of EVERY hot-path parameter (14.0 M, one flat buffer) and of
    the FPN feature map (which the torch backbone continues from).  The stages without parameters or gradients
    (camera geometry, per-RoI intrinsics, box correlation) come from the inference engine.
    Reference: MV2DSHead.forward_train (roi_heads/mv2d_s_head.py:236-307) + torch autograd."""

    def __init__(self, state_dict, device='cuda', engine_cfg=None, mode='S', use_denoise=None, denoise_weight=1.0,
                 neg_bbox_loss=None, **kw):
        super().__init__(state_dict, device=device, **kw)
        self._need_cuda()
        from .engine import HotPath
        sd = {k[len('roi_head.'):] if k.startswith('roi_head.') else k: v for k, v in state_dict.items()}
        self.mode = mode
        self.engine = HotPath(sd, mode=mode, device=self.device, **(engine_cfg or {}))
        # the reference trains MV2D-T with denoising queries and neg_bbox_loss, MV2D-S without either
        # (configs/mv2d/exp/*two_frames*:44-47, *single_frame*:44)
        self.use_denoise = (mode == 'T') if use_denoise is None else use_denoise
        self.neg_bbox_loss = (mode == 'T') if neg_bbox_loss is None else neg_bbox_loss
        self.denoise_weight = float(denoise_weight)
        assert mode == 'T' or not self.use_denoise, 'denoising queries are trained with the two-frame head (mode T)'
        self._front_ws = None
        self._fp = None

    @torch.no_grad()
    def forward(self, feat, proposal_list, img_metas, gt_boxes, gt_labels, rand=None, uploaded=None):
        """feat [V,256,h,w] fp32 NCHW (as the FPN emits it), proposal_list: V tensors [n_v, >=4], img_metas: V dicts.
        Two-frame head: ``rand`` [scalar*G,3] = the uniform noise of the denoising queries (None: torch.rand).
        uploaded: the tuple ``engine._upload_meta`` returned for this sample (CUDA-graph replays upload the metadata
        outside the captured region)."""
        if self.mode == 'T':
            return self._forward_t(feat, proposal_list, img_metas, gt_boxes, gt_labels, rand)
        eng, dev = self.engine, self.device
        f32 = dict(device=dev, dtype=torch.float32)
        feat = feat.to(**f32).contiguous()
        cams, rois, roi_start, counts, N = uploaded if uploaded is not None else eng._upload_meta(proposal_list, img_metas)
        i2l, trans = eng.geom_prep(cams)
        feat_nhwc, _ = eng.to_nhwc(feat)
        V, h, w, _ = feat_nhwc.shape
        corr = eng.box_corr(rois, roi_start, trans, N, V, img_metas, h, w)
        k_roi = eng.roi_align_qg(rois, cams, feat_nhwc, None, N, phase=4)['roi_intrinsics']      # K' only (weight independent)
        _, _, not_mask, _ = eng._masks(img_metas, h, w)
        c = eng.cfg
        out = dict(tok_mem=torch.empty((N, 49, 256), **f32), tok_kin=torch.empty((N, 49, 256), **f32),
                   ref=torch.empty((N, 3), **f32), d_feat=torch.empty((V, h, w, 256), **f32))
        ws_bytes = int(self.lib.mv2d_front_train_workspace_bytes(N, V, h, w))
        if self._front_ws is None or self._front_ws.numel() * 4 < ws_bytes:
            self._front_ws = torch.empty(ws_bytes // 4 + 64, **f32)
        p = lib.FrontTrainParams()
        p.N, p.V, p.h, p.w, p.L, p.stride = N, V, h, w, self.L, c['stride']
        p.depth_num, p.pad_h, p.pad_w = c['depth_num'], int(img_metas[0]['pad_shape'][0]), int(img_metas[0]['pad_shape'][1])
        p.depth_start = c['depth_start']
        p.position_range = (C.c_double * 6)(*c['position_range'])
        p.pc_range = (C.c_float * 6)(*self.pc_range)
        p.intrins_feat_scale = c['intrins_feat_scale']
        p.params, p.grads = self.params.data_ptr(), self.grads.data_ptr()
        p.rois, p.roi_intrinsics, p.extrinsics, p.img2lidar = rois.data_ptr(), k_roi.data_ptr(), cams[2].data_ptr(), i2l.data_ptr()
        p.not_mask, p.dim_t, p.feat = not_mask.data_ptr(), self.dim_t.data_ptr(), feat_nhwc.data_ptr()
        p.tok_mem, p.tok_kin, p.ref, p.d_feat = (out[k].data_ptr() for k in ('tok_mem', 'tok_kin', 'ref', 'd_feat'))
        p.workspace, p.workspace_bytes = self._front_ws.data_ptr(), ws_bytes
        lib.check(self.lib.mv2d_front_train_forward(C.byref(p), lib.stream_ptr()), 'mv2d_front_train_forward')
        self._fp, self._fout = p, out
        self._fkeep = (feat_nhwc, rois, k_roi, cams, i2l, not_mask)
        res = super().forward(out['ref'], out['tok_kin'], out['tok_mem'], corr['match'], corr['match_cnt'], gt_boxes, gt_labels)
        res.update(ref=out['ref'], tok_mem=out['tok_mem'], tok_kin=out['tok_kin'], match=corr['match'], match_cnt=corr['match_cnt'],
                   rois=rois, N=N)
        return res

    # ------------------------------------------------------------------ two-frame head (+ denoising queries)
    def _front_params(self, N, V, h, w, rois, k_roi, cams, i2l, not_mask, feat_nhwc, img_metas, out):
        c = self.engine.cfg
        f32 = dict(device=self.device, dtype=torch.float32)
        ws_bytes = int(self.lib.mv2d_front_train_workspace_bytes(N, V, h, w))
        if self._front_ws is None or self._front_ws.numel() * 4 < ws_bytes:
            self._front_ws = torch.empty(ws_bytes // 4 + 64, **f32)
        p = lib.FrontTrainParams()
        p.N, p.V, p.h, p.w, p.L, p.stride = N, V, h, w, self.L, c['stride']
        p.depth_num, p.pad_h, p.pad_w = c['depth_num'], int(img_metas[0]['pad_shape'][0]), int(img_metas[0]['pad_shape'][1])
        p.depth_start = c['depth_start']
        p.position_range = (C.c_double * 6)(*c['position_range'])
        p.pc_range = (C.c_float * 6)(*self.pc_range)
        p.intrins_feat_scale = c['intrins_feat_scale']
        p.params, p.grads = self.params.data_ptr(), self.grads.data_ptr()
        p.rois, p.roi_intrinsics, p.extrinsics, p.img2lidar = rois.data_ptr(), k_roi.data_ptr(), cams[2].data_ptr(), i2l.data_ptr()
        p.not_mask, p.dim_t, p.feat = not_mask.data_ptr(), self.dim_t.data_ptr(), feat_nhwc.data_ptr()
        p.tok_mem, p.tok_kin, p.ref, p.d_feat = (out[k].data_ptr() for k in ('tok_mem', 'tok_kin', 'ref', 'd_feat'))
        p.workspace, p.workspace_bytes = self._front_ws.data_ptr(), ws_bytes
        return p

    @torch.no_grad()
    def _forward_t(self, feat, proposal_list, img_metas, gt_boxes, gt_labels, rand=None):
        """MV2DTHead training forward (roi_heads/mv2d_t_head.py:26-142 under mv2d_s_head.py:236-307): front end with saved
        activations, key masks from the box correlation, denoising queries prepended (``mv2d_dn_prepare``), the decoder
        over pad + N rows with the feature cells as keys, Hungarian + denoising losses."""
        eng, dev = self.engine, self.device
        f32 = dict(device=dev, dtype=torch.float32)
        feat = feat.to(**f32).contiguous()
        cams, rois, roi_start, counts, N = eng._upload_meta(proposal_list, img_metas)
        i2l, trans = eng.geom_prep(cams)
        feat_nhwc, _ = eng.to_nhwc(feat)
        V, h, w, _ = feat_nhwc.shape
        R = V * h * w
        corr = eng.box_corr(rois, roi_start, trans, N, V, img_metas, h, w)
        k_roi = eng.roi_align_qg(rois, cams, feat_nhwc, None, N, phase=4)['roi_intrinsics']
        _, _, not_mask, _ = eng._masks(img_metas, h, w)
        out = dict(tok_mem=torch.empty((N, 49, 256), **f32), tok_kin=torch.empty((N, 49, 256), **f32),
                   ref=torch.empty((N, 3), **f32), d_feat=torch.empty((V, h, w, 256), **f32),
                   kin_map=torch.empty((R, 256), **f32))
        p = self._front_params(N, V, h, w, rois, k_roi, cams, i2l, not_mask, feat_nhwc, img_metas, out)
        p.kin_out = out['kin_map'].data_ptr()
        lib.check(self.lib.mv2d_front_train_forward(C.byref(p), lib.stream_ptr()), 'mv2d_front_train_forward')
        self._fp, self._fout = p, out
        self._fkeep = (feat_nhwc, rois, k_roi, cams, i2l, not_mask)
        # denoising queries + the training-mode key masks (a query without any key gets key 0: mv2d_t_head.py:80-82)
        gt = gt_boxes.to(**f32).contiguous().view(-1, 9)
        lab = gt_labels.to(dev).to(torch.int32).contiguous()
        G = gt.shape[0] if self.use_denoise else 0
        dn = dict(gt_boxes=gt[:G], gt_labels=lab[:G], rand=rand if G > 0 else torch.zeros((0, 3), **f32))
        qg = dict(ref=out['ref'], query_pos=out['ref'])        # query_pos is recomputed from the live parameters below
        qg_d, corr_d, T, pad, extra = eng.dn_prepare(qg, corr, N, dn)
        L, M = self.L, 1
        res_buf = dict(cls_scores=torch.empty((L, T, 10), **f32), bbox_preds=torch.empty((L, T, 10), **f32),
                       assigned=torch.empty((L, N), device=dev, dtype=torch.int32), losses=torch.empty((L, 4), **f32),
                       num_pos=torch.empty((L,), **f32), d_ref=torch.empty((T, 3), **f32),
                       d_kin_map=torch.empty((R, 256), **f32), d_mem_map=torch.empty((R, 256), **f32))
        q = lib.TrainParams()
        q.N, q.L, q.max_match, q.G, q.num_classes = N, L, M, gt.shape[0], 10
        q.mode, q.pad, q.num_rows, q.mask_words = 1, pad, R, corr['mask_words']
        q.neg_bbox_loss = int(self.neg_bbox_loss)
        q.vel_dt, q.dn_split, q.denoise_weight = eng._vel_dt(img_metas), eng.cfg['denoise_split'], self.denoise_weight
        q.pc_range = (C.c_float * 6)(*self.pc_range)
        c = self.loss_cfg
        for k in ('cls_cost_weight', 'reg_cost_weight', 'cls_loss_weight', 'bbox_loss_weight', 'focal_alpha', 'focal_gamma'):
            setattr(q, k, c[k])
        q.code_weights = (C.c_float * 10)(*c['code_weights'])
        q.stage_loss_weights = (C.c_float * lib.MAX_LAYERS)(*(self.stage_loss_weights + [0.0] * (lib.MAX_LAYERS - L)))
        q.params, q.grads, q.dim_t = self.params.data_ptr(), self.grads.data_ptr(), self.dim_t.data_ptr()
        ref_all = qg_d['ref'].contiguous()
        q.ref = ref_all.data_ptr()
        q.kin_map, q.mem_map = out['kin_map'].data_ptr(), feat_nhwc.data_ptr()
        q.keymask, q.key_list, q.key_cnt = corr_d['keymask'].data_ptr(), corr_d['key_list'].data_ptr(), corr_d['key_cnt'].data_ptr()
        q.self_attn_mask = extra['dn_attn_mask'].data_ptr() if pad > 0 else None
        q.dn_labels = extra['dn_labels'].data_ptr() if pad > 0 else None
        q.gt_boxes, q.gt_labels = (gt.data_ptr(), lab.data_ptr()) if gt.shape[0] > 0 else (None, None)
        for k in ('cls_scores', 'bbox_preds', 'assigned', 'losses', 'num_pos', 'd_ref', 'd_kin_map', 'd_mem_map'):
            setattr(q, k, res_buf[k].data_ptr())
        ws_bytes = int(self.lib.mv2d_decoder_train_workspace_bytes_p(C.byref(q)))
        if self._ws is None or self._ws.numel() * 4 < ws_bytes:
            self._ws = torch.empty(ws_bytes // 4 + 64, **f32)
        q.workspace, q.workspace_bytes = self._ws.data_ptr(), ws_bytes
        self._p, self._out = q, res_buf
        self._keep = (ref_all, gt, lab, corr_d, extra, qg_d)
        lib.check(self.lib.mv2d_decoder_train_forward(C.byref(q), lib.stream_ptr()), 'mv2d_decoder_train_forward')
        wts = self._stage_weights()
        ls = res_buf['losses']
        res = dict(cls_scores=res_buf['cls_scores'][:, pad:], bbox_preds=res_buf['bbox_preds'][:, pad:],
                   dn_cls_scores=res_buf['cls_scores'][:, :pad], dn_bbox_preds=res_buf['bbox_preds'][:, :pad],
                   assigned=res_buf['assigned'], loss_cls=ls[:, 0], loss_bbox=ls[:, 1], dn_loss_cls=ls[:, 2], dn_loss_bbox=ls[:, 3],
                   num_pos=res_buf['num_pos'], ref=out['ref'], rois=rois, N=N, dn_pad=pad, dn_labels=extra.get('dn_labels'))
        res['loss'] = (wts * (ls[:, 0] + ls[:, 1] + self.denoise_weight * (ls[:, 2] + ls[:, 3]))).sum()
        return res

    @torch.no_grad()
    def _backward_t(self, bbox_avg_factor=None):
        q, out = self._p, self._out
        if bbox_avg_factor is not None:
            self._baf = bbox_avg_factor.to(self.device, torch.float32).contiguous()
            q.bbox_avg_factor = self._baf.data_ptr()
        else:
            q.bbox_avg_factor = None
        lib.check(self.lib.mv2d_decoder_train_backward(C.byref(q), lib.stream_ptr()), 'mv2d_decoder_train_backward')
        p, pad, N = self._fp, int(q.pad), int(q.N)
        d_ref = out['d_ref'][pad:].contiguous()
        zeros = torch.zeros((N, 49, 256), device=self.device)
        p.d_ref, p.d_tok_kin, p.d_tok_mem = d_ref.data_ptr(), zeros.data_ptr(), zeros.data_ptr()
        p.d_pe_extra, p.d_feat_extra, p.d_feat_extra2 = out['d_kin_map'].data_ptr(), out['d_kin_map'].data_ptr(), out['d_mem_map'].data_ptr()
        self._bkeep = (d_ref, zeros)
        lib.check(self.lib.mv2d_front_train_backward(C.byref(p), lib.stream_ptr()), 'mv2d_front_train_backward')
        return dict(d_ref=d_ref, d_feat=self._fout['d_feat'].permute(0, 3, 1, 2))

    @torch.no_grad()
    def backward(self, bbox_avg_factor=None):
        """Accumulates every parameter gradient into the flat buffer; returns d loss / d feat as [V,256,h,w]."""
        if self.mode == 'T':
            return self._backward_t(bbox_avg_factor)
        self.backward_decoder(bbox_avg_factor)
        return self.backward_front()

    # The single-frame backward in its two halves: the decoder half fills the decoder slice of the flat gradient buffer
    # ([0, decoder_grad_end): 80 % of it) and is done before the front end starts, so a data-parallel step can all-reduce
    # that slice while the front-end half (PE MLPs, RoIAlign, query generator: ~60 % of the backward) still runs.
    @torch.no_grad()
    def backward_decoder(self, bbox_avg_factor=None):
        assert self.mode == 'S'
        self._gin = super().backward(bbox_avg_factor)
        return self._gin

    @torch.no_grad()
    def backward_front(self):
        gin, p = self._gin, self._fp
        p.d_ref, p.d_tok_kin, p.d_tok_mem = (gin[k].data_ptr() for k in ('d_ref', 'd_tok_kin', 'd_tok_mem'))
        lib.check(self.lib.mv2d_front_train_backward(C.byref(p), lib.stream_ptr()), 'mv2d_front_train_backward')
        return dict(gin, d_feat=self._fout['d_feat'].permute(0, 3, 1, 2))

    @property
    def decoder_grad_end(self):
        """End (in floats) of the decoder's slice of the flat buffers: every ``bbox_head.*`` tensor precedes the front end's."""
        front = min(off for name, (off, n) in self.table.items() if not name.startswith('bbox_head.'))
        assert all(off + n <= front for name, (off, n) in self.table.items() if name.startswith('bbox_head.'))
        return front


class TrainStep:
    """One data-parallel optimisation step over this rank's samples with several samples IN FLIGHT: ``lanes`` trainers
    share the flat parameter buffer, each has its own stream, workspaces and gradient buffer, and the samples are dealt
    round-robin.  One sample's step is a chain of ~650 mostly small kernels that leaves most of the GPU idle (DESIGN.md
    section 8), so two chains side by side nearly overlap; the lanes' gradient buffers are summed into lane 0's before
    the all-reduce.  lanes=1 is the plain sequential step.  Same results as running the samples one after the other
    up to fp32 summation order."""

    def __init__(self, state_dict, device='cuda', lanes=2, sync_bbox_avg_factor=True, use_graphs=None, overlap_all_reduce=True, **kw):
        assert lanes >= 1
        self._sd, self._kw = state_dict, dict(kw)
        # CUDA-graph replay of every sample's forward and backward (single-frame head; keyed by the sample's shapes): the
        # step is ~650 launches per sample, which one Python thread cannot enqueue as fast as the GPU retires them
        if use_graphs is None:
            import os
            use_graphs = os.environ.get('MV2D_TRAIN_GRAPHS', '1') != '0'
        self.use_graphs = bool(use_graphs) and kw.get('mode', 'S') == 'S'
        self._graphs = {}
        self.timing = None
        self.main = HotPathTrainer(state_dict, device=device, **kw)
        self.lanes = [self.main]
        self.streams = [torch.cuda.Stream(device=self.main.device)]
        # True (the reference's objective, cross_attention_head.py:419-420): loss_bbox is divided by the mean positive count
        # over all samples of the GLOBAL batch; False: by each sample's own count (no collective before the backward)
        self.sync_bbox_avg_factor = sync_bbox_avg_factor
        # the decoder slice of the gradient buffer is all-reduced under the front-end half of the backward (single-frame head)
        self.overlap_all_reduce = overlap_all_reduce
        self._comm = torch.cuda.Stream(device=self.main.device)
        self._ev_dec = [torch.cuda.Event()]
        for _ in range(lanes - 1):
            self._add_lane()
        self.total = self.main.total

    def _graph_entry(self, k, smp):
        """The captured forward / backward of lane k for this sample's shapes; the sample's tensors are copied into the
        graph's static inputs and its metadata uploaded (outside the graphs) before the replay."""
        lane = self.lanes[k]
        feat, boxes, metas, gt_boxes, gt_labels = smp[:5]
        eng, dev = lane.engine, lane.device
        N = max(sum(int(b.shape[0]) for b in boxes), 1)
        key = (k, N, int(gt_boxes.shape[0]), tuple(feat.shape), eng._masks(metas, feat.shape[2], feat.shape[3])[0])
        ent = self._graphs.get(key)
        if ent is None:
            st = dict(feat=torch.empty(tuple(feat.shape), dtype=torch.float32, device=dev),
                      gt_boxes=torch.empty((gt_boxes.shape[0], 9), dtype=torch.float32, device=dev),
                      gt_labels=torch.empty((gt_boxes.shape[0],), dtype=torch.int32, device=dev),
                      factor=torch.ones((lane.L,), dtype=torch.float32, device=dev))
            st['feat'].copy_(feat); st['gt_boxes'].copy_(gt_boxes.view(-1, 9)); st['gt_labels'].copy_(gt_labels)
            up = eng._upload_meta(boxes, metas)
            keep_grads = lane.grads.clone()
            lane.forward(st['feat'], boxes, metas, st['gt_boxes'], st['gt_labels'], uploaded=up)     # warm-up: sizes every buffer
            lane.backward(st['factor'])
            torch.cuda.synchronize()
            gf, gb, gb2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(gf):
                out = lane.forward(st['feat'], boxes, metas, st['gt_boxes'], st['gt_labels'], uploaded=up)
            with torch.cuda.graph(gb, pool=gf.pool()):
                lane.backward_decoder(st['factor'])
            with torch.cuda.graph(gb2, pool=gf.pool()):
                lane.backward_front()
            lane.grads.copy_(keep_grads)         # warm-up and capture do not count
            torch.cuda.synchronize()
            ent = dict(fwd=gf, bwd=gb, bwd_front=gb2, out=out,
                       keep=(lane._p, lane._out, lane._keep, lane._fp, lane._fout, lane._fkeep, lane._gin), **st)
            self._graphs[key] = ent
            torch.cuda.current_stream().wait_stream(torch.cuda.default_stream(dev))
        ent['feat'].copy_(feat, non_blocking=True)
        ent['gt_boxes'].copy_(gt_boxes.view(-1, 9), non_blocking=True)
        ent['gt_labels'].copy_(gt_labels, non_blocking=True)
        eng._upload_meta(boxes, metas)
        return ent

    def _add_lane(self):
        t = HotPathTrainer(self._sd, device=self.main.device, **self._kw)
        t.params = self.main.params            # ONE set of weights; own gradients, workspaces, engine buffers
        self.lanes.append(t)
        self.streams.append(torch.cuda.Stream(device=self.main.device))
        self._ev_dec.append(torch.cuda.Event())

    @torch.no_grad()
    def step(self, samples, lr=2e-4, weight_decay=0.01, world=1, optimize=True, max_grad_norm=None):
        """samples: list of (feat, proposal_list, img_metas, gt_boxes, gt_labels).  Returns the mean weighted loss
        (a device scalar).  Gradients end up summed in ``self.main.grads`` (all-reduced over ranks) and are scaled by
        1 / (ranks of the process group x samples); ``world`` is only checked against the group.  max_grad_norm: the
        reference's grad_clip (35 in the exp configs) over the hot-path parameters; None = off."""
        if not samples:
            return torch.zeros((), device=self.main.device)
        while len(self.lanes) < len(samples) and self.sync_bbox_avg_factor:
            self._add_lane()          # every sample keeps its activations until the global loss normaliser is known
        cur = torch.cuda.current_stream()
        for t in self.lanes:
            t.zero_grad()
        outs = []
        split, dec_end = False, 0
        for s in self.streams:
            s.wait_stream(cur)
        if not self.sync_bbox_avg_factor:      # each sample normalised by its own count: forward + backward back to back
            for i, smp in enumerate(samples):
                k = i % len(self.lanes)
                with torch.cuda.stream(self.streams[k]):
                    outs.append(self.lanes[k].forward(*smp))
                    self.lanes[k].backward()
        else:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record(cur)
            ents = []
            for i, smp in enumerate(samples):
                with torch.cuda.stream(self.streams[i]):
                    if self.use_graphs:
                        ent = self._graph_entry(i, smp)
                        ents.append(ent)
                        ent['fwd'].replay()
                        outs.append(ent['out'])
                    else:
                        outs.append(self.lanes[i].forward(*smp))
            for s in self.streams:
                cur.wait_stream(s)
            # loss_bbox / its gradient are divided by the mean positive count over the GLOBAL batch (one tiny all-reduce)
            factor = DecoderTrainer.global_bbox_avg_factor([o['num_pos'] for o in outs])
            for s in self.streams:
                s.wait_stream(cur)
            split = self.main.mode == 'S' and self.overlap_all_reduce
            dec_end = self.main.decoder_grad_end if split else 0
            for i in range(len(samples)):
                with torch.cuda.stream(self.streams[i]):
                    if self.use_graphs:
                        ents[i]['factor'].copy_(factor)
                        ents[i]['bwd'].replay()
                    elif split:
                        self.lanes[i].backward_decoder(factor)
                    else:
                        self.lanes[i].backward(factor)
                    if split:
                        self._ev_dec[i].record(self.streams[i])
            if split:
                # decoder gradients of every lane are complete: fold them and start their all-reduce (80 % of the bytes) on
                # the communication stream while the lanes run the front-end half of the backward
                with torch.cuda.stream(self._comm):
                    for i in range(len(samples)):
                        self._comm.wait_event(self._ev_dec[i])
                    for t in self.lanes[1:len(samples)]:
                        self.main.grads[:dec_end].add_(t.grads[:dec_end])
                    all_reduce_sum(self.main.grads[:dec_end])
                for i in range(len(samples)):
                    with torch.cuda.stream(self.streams[i]):
                        if self.use_graphs:
                            ents[i]['bwd_front'].replay()
                        else:
                            self.lanes[i].backward_front()
            elif self.use_graphs:
                for i in range(len(samples)):
                    with torch.cuda.stream(self.streams[i]):
                        ents[i]['bwd_front'].replay()
        for s in self.streams:
            cur.wait_stream(s)
        if self.sync_bbox_avg_factor:
            ev[1].record(cur)
        if self.sync_bbox_avg_factor and split:
            for t in self.lanes[1:len(samples)]:
                self.main.grads[dec_end:].add_(t.grads[dec_end:])
            nranks = all_reduce_sum(self.main.grads[dec_end:])
            cur.wait_stream(self._comm)
        else:
            for t in self.lanes[1:]:
                self.main.grads.add_(t.grads)
            nranks = self.main.all_reduce_grads()
        if self.sync_bbox_avg_factor:
            ev[2].record(cur)
        assert world in (1, nranks), f'world={world} but the process group has {nranks} ranks'
        scale = 1.0 / (nranks * len(samples))
        if optimize and max_grad_norm is not None:
            self.main.clip_grad_norm_(max_grad_norm, grad_scale=scale)
            scale = 1.0
        if optimize:
            self.main.adamw_step(lr=lr, weight_decay=weight_decay, grad_scale=scale)
        if self.sync_bbox_avg_factor:
            ev[3].record(cur)
            self.timing = ev        # fwd + bwd: ev[0] -> ev[1]; gradient all-reduce: ev[1] -> ev[2]; clip + AdamW: ev[2] -> ev[3]
        w = self.main._stage_weights()
        losses = [(w * (o['loss_cls'] + o['loss_bbox'])).sum() for o in outs]      # after the backward: rescaled loss_bbox
        return torch.stack(losses).mean()
