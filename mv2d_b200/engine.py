"""Host-side driver of the hot path: owns the device buffers, marshals the per-sample metadata
and enqueues the five C-ABI stages (geom_prep, pe3d, roi_align_qg, box_corr, decoder) on the
current CUDA stream.  No arithmetic of the path happens in Python/torch; torch supplies device
memory, streams and (optionally) CUDA-graph capture.

Mirrors ``MV2DHead.simple_test`` minus decode
(reference roi_heads/mv2d_head.py:249-261 -> mv2d_s_head.py:122-211 / mv2d_t_head.py:26-142).
"""
import contextlib
import ctypes as C
import math
import os

import numpy as np
import torch
import torch.nn.functional as F

from . import lib
from .pack import PackedNeck, PackedWeights

DEFAULTS = dict(
    pc_range=[-51.2, -51.2, -5.0, 51.2, 51.2, 3.0],
    position_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0],
    depth_num=64, depth_start=1.0, stride=16, intrins_feat_scale=0.1,
    sample_size=4, corr_num_depth=8, corr_depth_start=0.5, corr_depth_end=70.0,
    topk=1, iou_thr=0.0, ratio=0.0, expand_stride=0, num_views_per_frame=6,
    # denoising queries of the training-mode forward: MV2DSHead constructor defaults (mv2d_s_head.py:20-27)
    denoise_scalar=10, denoise_noise_scale=1.0, denoise_noise_trans=0.0, denoise_split=0.75, num_classes=10,
)


def feat_pad_mask(img_metas, h, w):
    """[V,h,w] uint8, 1 = cell lies in the padded border (utils/pe.py:146-155,
    mv2d_t_head.py:68-76): ones outside img_shape, nearest F.interpolate to (h, w).  The mask
    is separable, so the interpolation runs on two tiny 1-D tensors."""
    pad_h, pad_w, _ = img_metas[0]['pad_shape']
    out = np.zeros((len(img_metas), h, w), dtype=np.uint8)
    for v, m in enumerate(img_metas):
        ih, iw, _ = m['img_shape']
        rows = torch.ones(1, 1, pad_h, 1)
        rows[:, :, :ih] = 0
        cols = torch.ones(1, 1, 1, pad_w)
        cols[..., :iw] = 0
        r = F.interpolate(rows, size=(h, 1)).bool().view(h, 1).numpy()
        c = F.interpolate(cols, size=(1, w)).bool().view(1, w).numpy()
        out[v] = (r | c).astype(np.uint8)
    return out


class HotPath:
    """MV2D-S ('S') / MV2D-T ('T') decoder hot path on one GPU."""

    def __init__(self, state_dict, mode='S', device='cuda', cache_sine_branch=False, overlap=True,
                 persistent_decoder=None, fold_first_self_attn=True, xa_form=None, weights=None, **cfg):
        if not torch.cuda.is_available():
            raise RuntimeError('mv2d_b200.HotPath needs a CUDA device (there is no CPU fallback)')
        self.lib = lib.load()
        self.mode = mode
        self.device = torch.device(device)
        self.cfg = dict(DEFAULTS)
        if mode == 'T':
            # exp/mv2d_r50_frcnn_two_frames_1408x512_ep72.py:44-47, 121-124
            self.cfg.update(topk=20, expand_stride=2, denoise_noise_scale=1.25, denoise_split=0.6)
        self.cfg.update(cfg)
        # fold_first_self_attn: layer 0's self-attention output is a packed constant (pack.first_layer_self_attn_const)
        # weights: an already packed set to share (Pipeline lanes); otherwise packed from the state_dict here
        self.w = weights if weights is not None else PackedWeights(state_dict, self.device,
                                                                   fold_first_self_attn=fold_first_self_attn)
        self.L = self.w.num_layers
        self.cache_sine_branch = cache_sine_branch
        self._sine_cache = {}
        self._mask_cache = {}
        self._buf = {}
        self._pin = {}
        self._graphs = {}
        self._graphs_b = {}
        self._batch_grid = None
        self.overlap = overlap
        if persistent_decoder is None:
            # measured on B200 (profiles/r01_persistent_decoder.md): the single-launch persistent decoder is
            # correct but ~30 % slower than one launch per stage at N = 300, so it is opt-in
            persistent_decoder = os.environ.get('MV2D_DECODER', 'staged') == 'persistent'
        self.persistent_decoder = persistent_decoder
        if xa_form is None:
            # two-frame head: 1 = key-stationary cross-attention over projected K/V tiles (csrc/xa_tile.cuh),
            #                 0 = query-stationary absorbed form (the S head's formulation applied to ~2000 keys/query)
            xa_form = int(os.environ.get('MV2D_XA_FORM', '1'))
        self.xa_form = xa_form if (mode == 'T' and not persistent_decoder) else 0
        # the K/V projections are GPU-filling GEMMs, the decoder layers they overlap with are chains of small
        # latency-bound kernels: the decoder runs on a high-priority stream so its CTAs are placed first
        self._kv = torch.cuda.Stream(device=self.device)
        self._hi = torch.cuda.Stream(device=self.device, priority=-1)
        self._ev_hi0, self._ev_hi1 = torch.cuda.Event(), torch.cuda.Event()
        self._ev_pe = torch.cuda.Event()
        self._ev_kv = [torch.cuda.Event() for _ in range(self.L)]
        self._mem_split = None        # (map pointer, hi, lo) when to_nhwc wrote the TF32 halves of the map beside it
        self._kv_row_live = None      # live-tile flags the last K/V projection honoured (pointer), None = projected everything
        self._side = torch.cuda.Stream(device=self.device)
        self._side2 = torch.cuda.Stream(device=self.device)
        self._copy = torch.cuda.Stream(device=self.device)
        self._ev_fork, self._ev_join, self._ev_join2 = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
        self.graph_launches = 0
        c = self.cfg
        S, Dn = c['sample_size'], c['corr_num_depth']
        idx = torch.arange(Dn).float()
        bin_size = (c['corr_depth_end'] - c['corr_depth_start']) / (Dn * (1 + Dn))
        # box_correlation.py:198, 221-225 -- same torch CPU ops, so the tables are bit-identical
        self.lin = torch.linspace(0, 1, S).to(self.device)
        self.depths = (c['corr_depth_start'] + bin_size * idx * (idx + 1)).to(self.device)

    def launch_count(self):
        """Kernel launches of libmv2d_b200 in this process: counted inside the library for eager
        calls, plus (kernels captured in a graph) x (replays) for CUDA-graph replays."""
        return int(self.lib.mv2d_launch_count()) + self.graph_launches

    # ------------------------------------------------------------------ buffers
    def _get(self, name, shape, dtype=torch.float32):
        n = int(np.prod(shape))
        t = self._buf.get(name)
        if t is None or t.numel() < n or t.dtype != dtype:
            t = torch.empty(max(n, 1), dtype=dtype, device=self.device)
            self._buf[name] = t
        return t[:n].view(*shape) if n > 0 else t[:0].view(*shape)

    def _masks(self, img_metas, h, w):
        key = (tuple(tuple(m['img_shape']) for m in img_metas), tuple(img_metas[0]['pad_shape']), h, w)
        ent = self._mask_cache.get(key)
        if ent is None:
            pad = feat_pad_mask(img_metas, h, w)
            ent = (key, torch.from_numpy(pad).to(self.device), torch.from_numpy(1 - pad).to(self.device),
                   bool(pad.any()))
            self._mask_cache[key] = ent
        return ent

    # ------------------------------------------------------------------ stages
    def _upload_meta(self, proposal_list, img_metas):
        """Per-sample metadata in ONE pinned staging buffer and ONE H2D copy:
        [lidar2img | intrinsics | extrinsics] (3*V*16 f64), rois (N*5 f32), roi_start ((V+1) i32).
        rois follow mmdet bbox2roi: (view, x1, y1, x2, y2); an empty detection set becomes the
        reference's dummy box (mv2d_s_head.py:124-127)."""
        V = len(img_metas)
        if sum(len(p) for p in proposal_list) == 0:
            p0 = torch.tensor([[0, 50, 50, 100, 100, 0]], dtype=torch.float32, device=proposal_list[0].device)
            proposal_list = [p0] + list(proposal_list[1:])
        counts = [int(p.shape[0]) for p in proposal_list]
        N = sum(counts)
        if max(counts) > 256:       # CORR_MAXR of csrc/roi.cu: the IoU table of box_corr_kernel
            raise RuntimeError('box correlation handles at most 256 detections per view (include/mv2d_b200.h, limits)')
        on_device = proposal_list[0].is_cuda
        cam_b, roi_b, st_b = 3 * V * 16 * 8, N * 5 * 4, (V + 1) * 4
        roi_off = cam_b
        st_off = (roi_off + roi_b + 7) // 8 * 8
        total = st_off + st_b
        pin = self._pin.get('meta')
        if pin is None:   # fixed capacity: captured graphs bake the device address in, it must never move
            pin = torch.empty(128 * 1024, dtype=torch.uint8).pin_memory()
            self._pin['meta'] = pin
            self._pin['meta_ev'] = None
        elif self._pin['meta_ev'] is not None:
            # the previous call's asynchronous copy reads this staging buffer when its stream gets there, and the host
            # may be several samples ahead of the device (Pipeline, TrainStep): wait for THAT copy before overwriting
            # the buffer.  It sits at the head of the previous sample's work, so this returns at once unless the host
            # is more than a whole sample ahead on this engine.
            self._pin['meta_ev'].synchronize()
        assert total <= pin.numel(), f'too many RoIs/views for the metadata buffer ({total} B)'
        dev = self._get('meta', (pin.numel(),), torch.uint8)
        host = pin.numpy()
        cams = host[:cam_b].view(np.float64).reshape(3, V, 16)
        for v, m in enumerate(img_metas):
            cams[0, v] = np.asarray(m['lidar2img'], dtype=np.float64).reshape(16)
            cams[1, v] = np.asarray(m['intrinsics'], dtype=np.float64).reshape(16)
            cams[2, v] = np.asarray(m['extrinsics'], dtype=np.float64).reshape(16)
        if not on_device:
            rh = host[roi_off:roi_off + roi_b].view(np.float32).reshape(N, 5)
            o = 0
            for v, p in enumerate(proposal_list):
                n = counts[v]
                rh[o:o + n, 0] = v
                rh[o:o + n, 1:] = p[:, :4].float().numpy()
                o += n
        host[st_off:st_off + st_b].view(np.int32)[:] = np.concatenate([[0], np.cumsum(counts)])
        dev[:total].copy_(pin[:total], non_blocking=True)
        if self._pin['meta_ev'] is None:
            self._pin['meta_ev'] = torch.cuda.Event()
        self._pin['meta_ev'].record()
        d_cams = dev[:cam_b].view(torch.float64).view(3, V, 16)
        d_rois = dev[roi_off:roi_off + roi_b].view(torch.float32).view(N, 5)
        d_start = dev[st_off:st_off + st_b].view(torch.int32)
        if on_device:   # detections already live on the GPU: bbox2roi there
            view = torch.repeat_interleave(torch.arange(V, device=self.device, dtype=torch.float32),
                                           torch.tensor(counts, device=self.device))
            d_rois[:, 0] = view
            d_rois[:, 1:] = torch.cat([p[:, :4].float() for p in proposal_list], 0)
        return d_cams, d_rois, d_start, counts, N

    def _upload_cams(self, img_metas):
        """Camera matrices only (stage-level entry points of the plugin modules)."""
        V = len(img_metas)
        cams = np.empty((3, V, 16), dtype=np.float64)
        for v, m in enumerate(img_metas):
            cams[0, v] = np.asarray(m['lidar2img'], dtype=np.float64).reshape(16)
            cams[1, v] = np.asarray(m['intrinsics'], dtype=np.float64).reshape(16)
            cams[2, v] = np.asarray(m['extrinsics'], dtype=np.float64).reshape(16)
        d = self._get('cams_only', (3, V, 16), torch.float64)
        d.copy_(torch.from_numpy(cams), non_blocking=True)
        return d

    def _upload_rois(self, proposal_list):
        raise RuntimeError('use _upload_meta')

    def geom_prep(self, cams):
        V = cams.shape[1]
        i2l = self._get('img2lidar', (V, 16), torch.float64)
        trans = self._get('trans', (V, V, 16), torch.float64)
        lib.check(self.lib.mv2d_geom_prep(cams[0].data_ptr(), V, i2l.data_ptr(), trans.data_ptr(),
                                          lib.stream_ptr()), 'mv2d_geom_prep')
        return i2l, trans

    def to_nhwc(self, feat_nchw):
        V, Cc, h, w = feat_nchw.shape
        out = self._get('feat_nhwc', (V, h, w, Cc))
        out_tf32 = self._get('feat_tf32', (V, h, w, Cc))
        self._mem_split = None
        if self.mode == 'T' and self.xa_form == 1 and Cc == 256:
            # two-frame head: the value rows of the K/V projection are this map; its TF32 hi half is feat_tf32, the lo half
            # comes out of the same pass (kv_project then skips its mv2d_split_tf32 over the map)
            lo = self._get('mem_lo', (V * h * w, 256))
            lib.check(self.lib.mv2d_nchw_to_nhwc_split(lib.ptr(feat_nchw), out.data_ptr(), out_tf32.data_ptr(), lo.data_ptr(), V, Cc,
                                                       h * w, lib.stream_ptr()), 'mv2d_nchw_to_nhwc_split')
            self._mem_split = (out.data_ptr(), out_tf32, lo)
            return out, out_tf32
        lib.check(self.lib.mv2d_nchw_to_nhwc(lib.ptr(feat_nchw), out.data_ptr(), out_tf32.data_ptr(), V, Cc, h * w,
                                             lib.stream_ptr()), 'mv2d_nchw_to_nhwc')
        return out, out_tf32

    def neck(self, x, neck_weights, in_is_nhwc=False):
        """Next row f4: the MV2D neck (one-level FPN: 1x1 lateral + 3x3 output conv, detectors/mv2d.py:122-127) on the
        2D detector's P4 ``x`` [V,256,h,w] -> channels-last feature map [V,h,w,256] (+ TF32 copy): feed it to
        ``forward(..., feat_is_nhwc=True)``.  ``neck_weights``: a ``PackedNeck`` or a ``neck.*`` state_dict."""
        if not isinstance(neck_weights, PackedNeck):
            key = id(neck_weights)
            if getattr(self, '_neck_cache', (None, None))[0] != key:
                self._neck_cache = (key, PackedNeck(neck_weights, self.device))
            neck_weights = self._neck_cache[1]
        x = x.to(self.device, torch.float32).contiguous()
        V, h, w = (x.shape[0], x.shape[1], x.shape[2]) if in_is_nhwc else (x.shape[0], x.shape[2], x.shape[3])
        out = self._get('feat_nhwc', (V, h, w, 256))
        out_tf32 = self._get('feat_tf32', (V, h, w, 256))
        ws_bytes = self.lib.mv2d_fpn_neck_workspace_bytes(V, h, w)
        ws = self._get('neck_ws', (ws_bytes // 4 + 1,))
        p = lib.NeckParams()
        p.V, p.h, p.w, p.in_is_nhwc = V, h, w, int(in_is_nhwc)
        p.x = x.data_ptr()
        for k in ('lat_w', 'lat_w_lo', 'lat_b', 'fpn_w', 'fpn_w_lo', 'fpn_b'):
            setattr(p, k, neck_weights.p(k))
        p.feat, p.feat_tf32 = out.data_ptr(), out_tf32.data_ptr()
        p.workspace, p.workspace_bytes = ws.data_ptr(), ws_bytes
        lib.check(self.lib.mv2d_fpn_neck(C.byref(p), lib.stream_ptr()), 'mv2d_fpn_neck')
        return out, out_tf32

    def pe3d(self, feat_nhwc, img2lidar, img_metas, feat_tf32=None, phase=0, dims=None, batch=None):
        """PE.forward (utils/pe.py:137-169) -> (pe [V,h,w,256], kin = feat + pe or None).
        phase 1 = the feature-independent part only (feat_nhwc may still be in flight), 2 = the rest.
        batch: the descriptor ``_upload_meta_batch`` returns (V then counts the views of all samples)."""
        V, h, w = dims if dims is not None else feat_nhwc.shape[:3]
        self._last_grid = (h, w)
        c, W = self.cfg, self.w
        key, pad_mask, not_mask, has_pad = self._masks(img_metas, h, w) if batch is None else batch['masks']
        pe = self._get('pe', (V, h, w, 256))
        kin = self._get('kin', (V, h, w, 256)) if self.mode == 'T' else None
        ws_bytes = self.lib.mv2d_pe3d_workspace_bytes(V, h, w, c['depth_num'])
        ws = self._get('pe_ws', (ws_bytes // 4,))
        p = lib.PeParams()
        p.V, p.h, p.w, p.depth_num = V, h, w, c['depth_num']
        p.phase = phase
        # no padded cells (img_shape == pad_shape in every view): the sine branch's first layer is separable
        p.sine_separable = int((not has_pad) and os.environ.get('MV2D_SINE_SEPARABLE', '1') != '0')
        if batch is not None:
            p.views_per_sample = batch['Vs']
            p.sine_shared = int(batch['same_masks'] and os.environ.get('MV2D_SINE_SHARED', '1') != '0')
            img_metas = batch['metas'][0]
        p.pad_h, p.pad_w = int(img_metas[0]['pad_shape'][0]), int(img_metas[0]['pad_shape'][1])
        p.stride = c['stride']
        p.unfused_mlp = int(getattr(self, 'pe_unfused', False))
        p.depth_start = c['depth_start']
        p.position_range = (C.c_double * 6)(*c['position_range'])
        p.feat_tf32 = feat_tf32.data_ptr() if feat_tf32 is not None else None
        p.feat, p.img2lidar, p.not_mask, p.dim_t = feat_nhwc.data_ptr(), img2lidar.data_ptr(), not_mask.data_ptr(), W.p('dim_t')
        for f in ('w_pos0', 'b_pos0', 'w_pos2', 'b_pos2', 'w_adapt0', 'b_adapt0', 'w_adapt2', 'b_adapt2',
                  'w_se_reduce', 'b_se_reduce', 'w_se_expand', 'b_se_expand'):
            setattr(p, f, W.p(f))
        cached = self._sine_cache.get(key) if (self.cache_sine_branch and phase == 0 and batch is None) else None
        if cached is not None:
            p.sine_branch_cached = cached.data_ptr()
        elif self.cache_sine_branch and phase == 0 and batch is None:
            self._sine_cache[key] = torch.empty((V * h * w, 256), device=self.device)
            p.sine_branch_out = self._sine_cache[key].data_ptr()
        p.pe = pe.data_ptr()
        p.kin = kin.data_ptr() if kin is not None else None
        p.workspace, p.workspace_bytes = ws.data_ptr(), ws_bytes
        lib.check(self.lib.mv2d_pe3d(C.byref(p), lib.stream_ptr()), 'mv2d_pe3d')
        return pe, kin

    def roi_align_qg(self, rois, cams, feat_nhwc, pe_nhwc, N, phase=0):
        V, h, w, _ = feat_nhwc.shape
        c, W = self.cfg, self.w
        tok_feat = self._get('tok_feat', (N, 49, 256))
        tok_kin = self._get('tok_kin', (N, 49, 256)) if self.mode == 'S' else None
        kroi = self._get('roi_intrinsics', (N, 16), torch.float64)
        center = self._get('center_lidar', (N, 3))
        ref = self._get('ref', (N, 3))
        qpos = self._get('query_pos', (N, 256))
        ws_bytes = self.lib.mv2d_roi_align_qg_workspace_bytes(N)
        ws = self._get('qg_ws', (ws_bytes // 4,))
        p = lib.QgParams()
        p.N, p.V, p.h, p.w, p.stride, p.phase = N, V, h, w, c['stride'], phase
        p.pc_range = (C.c_float * 6)(*c['pc_range'])
        p.intrins_feat_scale = c['intrins_feat_scale']
        p.rois, p.intrinsics, p.extrinsics = rois.data_ptr(), cams[1].data_ptr(), cams[2].data_ptr()
        p.feat = feat_nhwc.data_ptr()
        p.pe = pe_nhwc.data_ptr() if (tok_kin is not None and pe_nhwc is not None) else None
        p.dim_t = W.p('dim_t')
        for f in ('w_conv', 'b_conv', 'w_conv_lo', 'w_fc', 'b_fc', 'w_enc0', 'b_enc0', 'w_enc2', 'b_enc2', 'w_center',
                  'b_center', 'w_qe0', 'b_qe0', 'w_qe2', 'b_qe2', 'w_fc_hi', 'w_fc_lo', 'w_enc0_hi', 'w_enc0_lo', 'w_enc2_hi',
                  'w_enc2_lo', 'w_qe0_hi', 'w_qe0_lo', 'w_qe2_hi', 'w_qe2_lo'):
            setattr(p, f, W.p(f))
        p.tok_feat = tok_feat.data_ptr()
        p.tok_kin = tok_kin.data_ptr() if tok_kin is not None else None
        p.roi_intrinsics, p.center_lidar = kroi.data_ptr(), center.data_ptr()
        p.ref, p.query_pos = ref.data_ptr(), qpos.data_ptr()
        p.workspace, p.workspace_bytes = ws.data_ptr(), ws_bytes
        lib.check(self.lib.mv2d_roi_align_qg(C.byref(p), lib.stream_ptr()), 'mv2d_roi_align_qg')
        return dict(tok_feat=tok_feat, tok_kin=tok_kin, roi_intrinsics=kroi, center_lidar=center, ref=ref,
                    query_pos=qpos)

    def box_corr(self, rois, roi_start, trans, N, V, img_metas, h, w, batch=None):
        """V = views of ONE sample; with ``batch`` N = B * Np rows and roi_start is [B, V+1]."""
        c = self.cfg
        max_match = 1 + (V - 1) * c['topk']
        match = self._get('match', (N, max_match), torch.int32)
        cnt = self._get('match_cnt', (N,), torch.int32)
        p = lib.CorrParams()
        p.N, p.V = N, V
        B = 1
        if batch is not None:
            B = p.batch = batch['B']
            p.rows_per_sample = batch['Np']
            img_metas = batch['metas'][0]
        p.img_h, p.img_w = int(img_metas[0]['pad_shape'][0]), int(img_metas[0]['pad_shape'][1])
        p.topk, p.sample_size, p.num_depth, p.max_match = c['topk'], c['sample_size'], c['corr_num_depth'], max_match
        p.ratio, p.iou_thr, p.depth_start = c['ratio'], c['iou_thr'], c['corr_depth_start']
        p.rois, p.roi_start, p.trans = rois.data_ptr(), roi_start.data_ptr(), trans.data_ptr()
        p.lin, p.depths = self.lin.data_ptr(), self.depths.data_ptr()
        p.match, p.match_cnt = match.data_ptr(), cnt.data_ptr()
        out = dict(match=match, match_cnt=cnt, max_match=max_match)
        if self.mode == 'T':
            words = (V * h * w + 31) // 32
            keymask = self._get('keymask', (N, words), torch.int32)
            key_cnt = self._get('key_cnt', (N,), torch.int32)
            _, pad_mask, _, has_pad = self._masks(img_metas, h, w) if batch is None else batch['masks']
            p.h, p.w, p.stride, p.expand_stride = h, w, c['stride'], c['expand_stride']
            p.pad_mask = pad_mask.data_ptr() if has_pad else None
            # the compacted key lists feed the query-stationary form only (and the denoising prepare builds its own)
            key_list = self._get('key_list', (N, words * 32), torch.int16) if batch is None else None
            p.keymask, p.key_cnt = keymask.data_ptr(), key_cnt.data_ptr()
            p.key_list = key_list.data_ptr() if key_list is not None else None
            out.update(keymask=keymask, key_cnt=key_cnt, key_list=key_list, mask_words=words)
        lib.check(self.lib.mv2d_box_corr(C.byref(p), lib.stream_ptr()), 'mv2d_box_corr')
        if self.mode == 'T' and self.xa_form == 1 and N > 0:
            # tile / record lists of the key-stationary cross-attention depend on the masks only: build them here,
            # beside the position embedding, instead of at the head of the decoder
            d = lib.DecoderParams()
            d.N, d.num_rows, d.grid_h, d.grid_w = N, B * V * h * w, h, w
            d.keymask, d.mask_words = out['keymask'].data_ptr(), out['mask_words']
            if batch is not None:
                d.batch, d.rows_per_sample = B, batch['Np']
            xa_bytes = self.lib.mv2d_xa_tile_workspace_bytes_batch(B, N // B, V, h, w)
            xa_ws = self._get('xa_ws', (xa_bytes,), torch.uint8)
            d.xa_workspace, d.xa_workspace_bytes = xa_ws.data_ptr(), xa_bytes
            live_ok = B == 1 or (V * h * w) % 128 == 0      # a 128-row tile of the projection must lie inside one sample
            row_live = self._get('kv_row_live', ((B * V * h * w + 127) // 128,), torch.uint8) if live_ok else None
            d.row_tile_live = row_live.data_ptr() if live_ok else None
            lib.check(self.lib.mv2d_xa_tile_prepare(C.byref(d), lib.stream_ptr()), 'mv2d_xa_tile_prepare')
            out['xa_prepared_for'] = (out['keymask'].data_ptr(), N, xa_ws.data_ptr())
            out['row_tile_live'] = row_live     # 128-row tiles of the K/V projection some query has a key in
        return out

    def dn_prepare(self, qg, corr, N, dn):
        """Row a20 (training-mode forward): prepend scalar*G denoising queries built from the GT boxes.
        dn = dict(gt_boxes [G,9] = (gravity centre, w, l, h, yaw, vx, vy), gt_labels [G], rand [scalar*G,3] or
        None -> torch.rand).  Returns (qg', corr', T, pad, extras) for ``decoder``
        (mv2d_s_head.py:39-120,158-180; mv2d_t_head.py:79-98)."""
        c, W = self.cfg, self.w
        dev = self.device
        gt = dn['gt_boxes'].to(dev, torch.float32).contiguous().view(-1, 9)
        G, scalar = gt.shape[0], c['denoise_scalar']
        labels = dn['gt_labels'].to(dev).to(torch.int32).contiguous()
        pad = G * scalar
        rand = dn.get('rand')
        rand = torch.rand(pad, 3, device=dev) if rand is None else rand.to(dev, torch.float32).contiguous()
        assert tuple(rand.shape) == (pad, 3) and labels.numel() == G
        T = pad + N
        p = lib.DnParams()
        p.N, p.G, p.scalar, p.num_classes = N, G, scalar, c['num_classes']
        p.mode = 0 if self.mode == 'S' else 1
        p.noise_scale, p.noise_trans = c['denoise_noise_scale'], c['denoise_noise_trans']
        p.split, p.eps = c['denoise_split'], 1e-4
        p.pc_range = (C.c_float * 6)(*c['pc_range'])
        p.gt_boxes, p.gt_labels, p.rand, p.ref = gt.data_ptr(), labels.data_ptr(), rand.data_ptr(), qg['ref'].data_ptr()
        for k in ('w_qe0', 'b_qe0', 'w_qe2', 'b_qe2', 'dim_t'):
            setattr(p, k, W.p(k))
        ref_all = self._get('dn_ref', (T, 3))
        dn_labels = self._get('dn_labels', (pad,), torch.int32)
        attn_mask = self._get('dn_attn_mask', (T, T), torch.uint8)
        qpos_all = self._get('dn_query_pos', (T, 256))
        p.ref_all, p.dn_labels, p.attn_mask, p.query_pos_all = (ref_all.data_ptr(), dn_labels.data_ptr(),
                                                                attn_mask.data_ptr(), qpos_all.data_ptr())
        corr2 = dict(corr)
        words = 0
        if self.mode == 'S':
            mm = max(N, corr['max_match'])
            match_all = self._get('dn_match', (T, mm), torch.int32)
            cnt_all = self._get('dn_match_cnt', (T,), torch.int32)
            p.match, p.match_cnt, p.max_match, p.max_match_all = (corr['match'].data_ptr(), corr['match_cnt'].data_ptr(),
                                                                  corr['max_match'], mm)
            p.match_all, p.match_cnt_all = match_all.data_ptr(), cnt_all.data_ptr()
            corr2.update(match=match_all, match_cnt=cnt_all, max_match=mm)
        else:
            words = corr['mask_words']
            km_all = self._get('dn_keymask', (T, words), torch.int32)
            kl_all = self._get('dn_key_list', (T, words * 32), torch.int16)
            kc_all = self._get('dn_key_cnt', (T,), torch.int32)
            p.keymask, p.key_cnt, p.mask_words, p.train_unmask = corr['keymask'].data_ptr(), corr['key_cnt'].data_ptr(), words, 1
            p.keymask_all, p.key_list_all, p.key_cnt_all = km_all.data_ptr(), kl_all.data_ptr(), kc_all.data_ptr()
            corr2.update(keymask=km_all, key_list=kl_all, key_cnt=kc_all)
        ws_bytes = self.lib.mv2d_dn_workspace_bytes(T, words)
        ws = self._get('dn_ws', (ws_bytes // 4 + 1,))
        p.workspace, p.workspace_bytes = ws.data_ptr(), ws_bytes
        lib.check(self.lib.mv2d_dn_prepare(C.byref(p), lib.stream_ptr()), 'mv2d_dn_prepare')
        qg2 = dict(qg, query_pos=qpos_all, ref=ref_all)
        return qg2, corr2, T, pad, dict(dn_labels=dn_labels, dn_attn_mask=attn_mask, dn_ref=ref_all[:pad])

    def kv_project(self, kin_rows, mem_rows, record_events=False, row_live=None):
        """Two-frame head, xa_form 1: K_l = (mem + pos) Wk_l^T and V_l = mem Wv_l^T for every decoder layer
        (3xTF32 tcgen05 GEMMs over all V*h*w cells), on the current stream.  With record_events the per-layer
        events ``_ev_kv[l]`` are recorded so the decoder layers on another stream can start as soon as their
        projection is done."""
        L, R = self.L, kin_rows.shape[0]
        n = R * 256
        kp, vp = self._get('kp', (L, R, 256)), self._get('vp', (L, R, 256))
        st = lib.stream_ptr()
        p = lib.KvParams()
        p.num_rows, p.L = R, L
        if os.environ.get('MV2D_KV_RAW', '0') == '1':
            # plain fp32 rows: the GEMM splits rows and weights into TF32 hi/lo in shared memory.  Bit-identical to
            # the pre-split call and half the L2->SM bytes, but measured SLOWER on B200 (672 vs 528 us for the 12
            # projections): with 3 x 64 KB stages the split sits on the TMA -> MMA latency chain.  Opt-in.
            p.kin_hi, p.mem_hi = kin_rows.data_ptr(), mem_rows.data_ptr()
        else:
            kin_hi, kin_lo = self._get('kin_hi', (R, 256)), self._get('kin_lo', (R, 256))
            lib.check(self.lib.mv2d_split_tf32(kin_rows.data_ptr(), kin_hi.data_ptr(), kin_lo.data_ptr(), n, st), 'mv2d_split_tf32')
            ms = getattr(self, '_mem_split', None)
            if ms is not None and ms[0] == mem_rows.data_ptr() and ms[2].shape[0] == R:
                mem_hi, mem_lo = ms[1], ms[2]           # written by to_nhwc beside the map
            else:
                mem_hi, mem_lo = self._get('mem_hi', (R, 256)), self._get('mem_lo', (R, 256))
                lib.check(self.lib.mv2d_split_tf32(mem_rows.data_ptr(), mem_hi.data_ptr(), mem_lo.data_ptr(), n, st), 'mv2d_split_tf32')
            p.kin_hi, p.kin_lo, p.mem_hi, p.mem_lo = kin_hi.data_ptr(), kin_lo.data_ptr(), mem_hi.data_ptr(), mem_lo.data_ptr()
        p.layers = self.w.layers_ptr()
        p.kp, p.vp = kp.data_ptr(), vp.data_ptr()
        self._kv_row_live = None
        if row_live is not None and os.environ.get('MV2D_KV_SKIP', '1') != '0':
            p.row_tile_live = self._kv_row_live = row_live.data_ptr()
        # one persistent launch for all layers (csrc/kvproj.cu; needs the pre-split rows): it holds every SM anyway, so
        # the decoder layers could not overlap it; MV2D_KV_PER_LAYER=1 keeps one call (and one event) per layer
        per_layer = os.environ.get('MV2D_KV_PER_LAYER', '0') == '1' or os.environ.get('MV2D_KV_RAW', '0') == '1' \
            or os.environ.get('MV2D_KV_PERSISTENT', '1') == '0'
        for l in (range(L) if per_layer else (0,)):
            p.layer_begin, p.layer_end = (l, l + 1) if per_layer else (0, L)
            lib.check(self.lib.mv2d_kv_project(C.byref(p), st), 'mv2d_kv_project')
            if record_events:
                for e in (self._ev_kv[l:l + 1] if per_layer else self._ev_kv):
                    e.record(torch.cuda.current_stream())
        return kp, vp

    def _decoder_params(self, qg, corr, kin_rows, mem_rows, N, vel_dt=0.0, self_attn_mask=None, vel_row_start=0,
                        kv=None, grid=None, batch=None):
        c, W, L = self.cfg, self.w, self.L
        cls = self._get('cls_scores', (L, N, 10))
        box = self._get('bbox_preds', (L, N, 10))
        outs = self._get('outs_dec', (L, N, 256))
        ws_bytes = self.lib.mv2d_decoder_workspace_bytes(N, L)
        ws = self._get('dec_ws', (ws_bytes // 4,))
        p = lib.DecoderParams()
        p.N, p.L = N, L
        p.persistent = 1 if self.persistent_decoder else 0
        p.mode = 0 if self.mode == 'S' else 1
        p.num_rows = kin_rows.shape[0]
        p.pc_range = (C.c_float * 6)(*c['pc_range'])
        p.vel_dt, p.vel_row_start = vel_dt, vel_row_start
        p.query_pos, p.ref = qg['query_pos'].data_ptr(), qg['ref'].data_ptr()
        p.kin_rows, p.mem_rows = kin_rows.data_ptr(), mem_rows.data_ptr()
        if self.mode == 'S':
            p.match, p.match_cnt, p.max_match = corr['match'].data_ptr(), corr['match_cnt'].data_ptr(), corr['max_match']
        else:
            p.keymask, p.mask_words = corr['keymask'].data_ptr(), corr['mask_words']
            if corr.get('key_list') is not None:
                p.key_list, p.key_cnt = corr['key_list'].data_ptr(), corr['key_cnt'].data_ptr()
        p.self_attn_mask = self_attn_mask.data_ptr() if self_attn_mask is not None else None
        p.layers, p.branches = W.layers_ptr(), W.branches_ptr()
        p.cls_scores, p.bbox_preds, p.outs_dec = cls.data_ptr(), box.data_ptr(), outs.data_ptr()
        p.workspace, p.workspace_bytes = ws.data_ptr(), ws_bytes
        B = 1
        if batch is not None:
            B = p.batch = batch['B']
            p.rows_per_sample = batch['Np']
            p.n_real = batch['n_real'].data_ptr()
            if batch.get('vel_dt') is not None:
                p.vel_dt_batch = batch['vel_dt'].data_ptr()
        if kv is not None:      # two-frame head, key-stationary cross-attention over the projected K/V
            V = kin_rows.shape[0] // (B * grid[0] * grid[1])
            xa_bytes = self.lib.mv2d_xa_tile_workspace_bytes_batch(B, N // B, V, grid[0], grid[1])
            xa_ws = self._get('xa_ws', (xa_bytes,), torch.uint8)
            p.xa_form, p.grid_h, p.grid_w = 1, grid[0], grid[1]
            p.kp, p.vp = kv[0].data_ptr(), kv[1].data_ptr()
            p.xa_workspace, p.xa_workspace_bytes = xa_ws.data_ptr(), xa_bytes
            p.xa_prepared = int(corr.get('xa_prepared_for') == (corr['keymask'].data_ptr(), N, xa_ws.data_ptr()))
            # the tensor-core tile attention reads whole 8x8 tiles: it must know which 128-row tiles the projection skipped
            p.row_tile_live = self._kv_row_live
        return p, cls, box, outs

    def decoder(self, qg, corr, kin_rows, mem_rows, N, vel_dt=0.0, self_attn_mask=None, vel_row_start=0,
                kv=None, grid=None, wait_kv_events=False, batch=None):
        L = self.L
        if self.xa_form == 1 and kv is None:    # stage-level call: project on this stream, then decode
            grid = grid or self._last_grid
            kv = self.kv_project(kin_rows, mem_rows)
        p, cls, box, outs = self._decoder_params(qg, corr, kin_rows, mem_rows, N, vel_dt, self_attn_mask, vel_row_start,
                                                 kv, grid, batch)
        if kv is not None and wait_kv_events:
            main = torch.cuda.current_stream()
            for l in range(L):          # layer l starts when its projection (on the kv stream) is done
                main.wait_event(self._ev_kv[l])
                p.layer_begin, p.layer_end = l, l + 1
                lib.check(self.lib.mv2d_decoder(C.byref(p), lib.stream_ptr()), 'mv2d_decoder')
        else:
            lib.check(self.lib.mv2d_decoder(C.byref(p), lib.stream_ptr()), 'mv2d_decoder')
        return cls, box, outs

    def cross_attention_core(self, qg, corr, kin_rows, mem_rows, N, q, layer=0, kv=None, grid=None, batch=None):
        """The sparse cross-attention core of one decoder layer alone (``mv2d_cross_attention_core``): S head
        q [N,2048] absorbed queries -> ctx [N,2048]; T head (xa_form 1, after ``box_corr`` prepared the key tiles)
        q [N,256] -> ctx [N,256].  Used by the attention-only microbenchmark (tools/attention_sweep.py)."""
        p, _, _, _ = self._decoder_params(qg, corr, kin_rows, mem_rows, N, kv=kv, grid=grid, batch=batch)
        ctx = self._get('xa_core_ctx', tuple(q.shape))
        lib.check(self.lib.mv2d_cross_attention_core(C.byref(p), layer, q.data_ptr(), ctx.data_ptr(), None, lib.stream_ptr()),
                  'mv2d_cross_attention_core')
        return ctx

    # ------------------------------------------------------------------ whole path
    def _vel_dt(self, img_metas):
        nvf = self.cfg['num_views_per_frame']
        if self.mode != 'T' or len(img_metas) <= nvf:
            return 0.0
        ts = np.array([m['timestamp'] for m in img_metas], dtype=np.float64)   # mv2d_t_head.py:131-132
        return float(ts[nvf:].mean() - ts[:nvf].mean())

    def _enqueue_pre(self, cams, img_metas, dims):
        """Everything that does not read the feature map: camera geometry, frustum coordinates, position MLP
        and sine branch.  With host-resident inputs this overlaps the H2D copy of the feature map."""
        i2l, trans = self.geom_prep(cams)
        feat_buf = self._get('feat_nhwc', (dims[0], dims[1], dims[2], 256))
        self.pe3d(feat_buf, i2l, img_metas, None, phase=1, dims=dims)
        return i2l, trans

    def _enqueue_post(self, feat_in, feat_is_nhwc, cams, rois, roi_start, counts, N, img_metas, i2l, trans, pe_phase=2,
                      dn=None, batch=None):
        V = len(img_metas)          # views of ONE sample (a batch stacks B*V views in feat_in / cams)
        assert batch is None or dn is None, 'denoising queries run one sample at a time'
        feat_tf32 = None
        if feat_is_nhwc:
            feat = feat_in
            self._mem_split = None      # a map that did not come through to_nhwc has no lo half beside it
            _, h, w, _ = feat.shape
            own = self._buf.get('feat_nhwc')
            if own is not None and 'feat_tf32' in self._buf and feat.data_ptr() == own.data_ptr():
                feat_tf32 = self._get('feat_tf32', tuple(feat.shape))    # written beside it by ``neck``
        else:
            _, _, h, w = feat_in.shape
            feat, feat_tf32 = self.to_nhwc(feat_in)
        # A batch of the single-frame head runs the front end on ONE stream: PE first, then RoIAlign pools feat and
        # feat + pe in one pass (phase 0) instead of a second pass over the RoIs after the join -- every kernel fills the
        # GPU at B x 300 RoIs, so the fork buys nothing and the second pass costs 160 us per 8 samples
        # (tools/batch_sweep.py --overlap 0/1: 2669 vs 2582 samples/s resident, 2578 vs 2529 end to end at B = 8 x 4 lanes)
        if self.overlap and not (batch is not None and self.mode == 'S'):
            # fork: everything of the query generator that does not need the position embedding (RoIAlign
            # of the image feature, 3x3 conv, FC chain, reference points, query embedding) runs on a side
            # stream and the box correlation on a second one, concurrently with the position MLPs / SE gate /
            # PE combine on the main stream
            main = torch.cuda.current_stream()
            self._ev_fork.record(main)
            with torch.cuda.stream(self._side):
                self._side.wait_event(self._ev_fork)
                qg = self.roi_align_qg(rois, cams, feat, None, N, phase=1)
                self._ev_join.record(self._side)
            with torch.cuda.stream(self._side2):
                self._side2.wait_event(self._ev_fork)
                corr = self.box_corr(rois, roi_start, trans, N, V, img_metas, h, w, batch=batch)
                self._ev_join2.record(self._side2)
            pe, kin = self.pe3d(feat, i2l, img_metas, feat_tf32, phase=pe_phase, batch=batch)
            if self.xa_form == 1:
                # K/V projections of all layers on their own stream; decoder layer l waits for projection l only
                self._ev_pe.record(main)
                with torch.cuda.stream(self._kv):
                    self._kv.wait_event(self._ev_pe)
                    self._kv.wait_event(self._ev_join2)        # the live-tile flags come from the box correlation stream
                    # denoising queries may add keys (train_unmask gives key 0 to a query without any): project everything
                    kv = self.kv_project(kin.view(-1, 256), feat.view(-1, 256), record_events=True,
                                         row_live=corr.get('row_tile_live') if dn is None else None)
            main.wait_event(self._ev_join)
            main.wait_event(self._ev_join2)
            if self.mode == 'S':
                self.roi_align_qg(rois, cams, feat, pe, N, phase=2)      # tok_kin = tok_feat + RoIAlign(pe)
        else:
            pe, kin = self.pe3d(feat, i2l, img_metas, feat_tf32, phase=pe_phase, batch=batch)
            qg = self.roi_align_qg(rois, cams, feat, pe, N)
            corr = self.box_corr(rois, roi_start, trans, N, V, img_metas, h, w, batch=batch)
            if self.xa_form == 1:
                kv = self.kv_project(kin.view(-1, 256), feat.view(-1, 256))
        qg_d, corr_d, T, pad, extra, sa_mask = qg, corr, N, 0, {}, None
        if dn is not None:      # training-mode forward: denoising queries are prepended (row a20)
            qg_d, corr_d, T, pad, extra = self.dn_prepare(qg, corr, N, dn)
            sa_mask = extra['dn_attn_mask']
        if self.mode == 'S':
            cls, box, outs = self.decoder(qg_d, corr_d, qg['tok_kin'].view(-1, 256), qg['tok_feat'].view(-1, 256), T,
                                          self_attn_mask=sa_mask, batch=batch)
        else:
            piped = self.xa_form == 1 and self.overlap
            run_on = self._hi if (piped and os.environ.get('MV2D_DEC_PRIO', '1') != '0') else None
            if run_on is not None:
                cur = torch.cuda.current_stream()
                self._ev_hi0.record(cur)
                run_on.wait_event(self._ev_hi0)
            with torch.cuda.stream(run_on) if run_on is not None else contextlib.nullcontext():
                cls, box, outs = self.decoder(qg_d, corr_d, kin.view(-1, 256), feat.view(-1, 256), T,
                                              vel_dt=self._vel_dt(img_metas), self_attn_mask=sa_mask, vel_row_start=pad,
                                              kv=kv if self.xa_form == 1 else None, grid=(h, w), wait_kv_events=piped,
                                              batch=batch)
                if run_on is not None:
                    self._ev_hi1.record(run_on)
            if run_on is not None:
                cur.wait_event(self._ev_hi1)
        out = dict(cls_scores=cls[:, pad:], bbox_preds=box[:, pad:], outs_dec=outs[:, pad:], rois=rois, pe=pe,
                   feat_nhwc=feat, N=N, num_per_view=counts)
        out.update(qg)
        out.update(corr)
        if dn is not None:
            out.update(extra, dn_cls_scores=cls[:, :pad], dn_bbox_preds=box[:, :pad], dn_pad=pad,
                       query_pos=qg_d['query_pos'][pad:])
        return out

    def _enqueue(self, feat_in, feat_is_nhwc, cams, rois, roi_start, counts, N, img_metas, dn=None):
        """Enqueue every stage of the path on the current stream (capturable: no host sync)."""
        dims = (feat_in.shape[0], feat_in.shape[1], feat_in.shape[2]) if feat_is_nhwc else \
               (feat_in.shape[0], feat_in.shape[2], feat_in.shape[3])
        del dims
        # device-resident input: the whole PE runs on the main stream beside the query generator (side stream)
        i2l, trans = self.geom_prep(cams)
        return self._enqueue_post(feat_in, feat_is_nhwc, cams, rois, roi_start, counts, N, img_metas, i2l, trans,
                                  pe_phase=0, dn=dn)

    @torch.no_grad()
    def forward(self, feat, proposal_list, img_metas, feat_is_nhwc=False, use_graph=False, dn=None, bucket=1):
        """feat [V,256,h,w] fp32 (NCHW as the FPN emits it; device, or pinned host memory),
        proposal_list: V tensors [n_v, >=4] (device or host), img_metas: V dicts.  Returns a dict
        with cls_scores / bbox_preds [L,N,10] and the stage tensors (views of reused buffers).

        use_graph=True replays a CUDA graph of the whole path captured for this (N, V, h, w,
        masks) signature: inputs are copied into the graph's static buffers, then one launch.

        dn = dict(gt_boxes, gt_labels[, rand]) runs the training-mode forward with denoising queries
        (``dn_prepare``; eager only): the result additionally holds dn_cls_scores / dn_bbox_preds
        [L,pad,10], dn_labels, dn_attn_mask, dn_pad."""
        if bucket > 1:
            # a stream of samples whose detection count changes every time: the query rows are padded to a multiple of
            # ``bucket`` (padding rows carry a dummy box, the device-side count keeps them out of every real row's
            # result), so one captured graph serves a whole bucket of N instead of one graph per exact N
            assert dn is None and not feat_is_nhwc
            o = self.forward_batch(feat[None], [proposal_list], [img_metas], use_graph=use_graph, bucket=bucket)
            out = dict(o)
            out.update(o['samples'][0])
            return out
        if not use_graph:
            feat = feat.to(self.device, torch.float32, non_blocking=True).contiguous()
            cams, rois, roi_start, counts, N = self._upload_meta(proposal_list, img_metas)
            return self._enqueue(feat, feat_is_nhwc, cams, rois, roi_start, counts, N, img_metas, dn=dn)
        assert dn is None, 'the denoising (training-mode) forward is eager only'
        assert not feat_is_nhwc
        counts = [int(p.shape[0]) for p in proposal_list]
        N = max(sum(counts), 1)
        mkey = self._masks(img_metas, feat.shape[2], feat.shape[3])[0]
        host_input = not feat.is_cuda
        key = (N, tuple(feat.shape), mkey, self._vel_dt(img_metas), host_input)
        ent = self._graphs.get(key)
        if ent is None:
            # warm-up (eager) run: sizes every buffer and sets kernel attributes, then capture
            static_feat = torch.empty(tuple(feat.shape), dtype=torch.float32, device=self.device)
            static_feat.copy_(feat, non_blocking=True)
            cams, rois, roi_start, counts, N = self._upload_meta(proposal_list, img_metas)
            self._enqueue(static_feat, False, cams, rois, roi_start, counts, N, img_metas)
            torch.cuda.synchronize()
            dims = (feat.shape[0], feat.shape[2], feat.shape[3])
            launches0 = self.launch_count()
            if host_input:
                # two graphs: `pre` does not touch the feature map and runs while it is still being copied in
                g_pre, g_post = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_pre):
                    i2l, trans = self._enqueue_pre(cams, img_metas, dims)
                with torch.cuda.graph(g_post, pool=g_pre.pool()):
                    out = self._enqueue_post(static_feat, False, cams, rois, roi_start, counts, N, img_metas, i2l, trans)
                graphs = (g_pre, g_post)
            else:
                g_all = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_all):
                    out = self._enqueue(static_feat, False, cams, rois, roi_start, counts, N, img_metas)
                graphs = (g_all,)
            ent = dict(graphs=graphs, out=out, feat=static_feat, launches=self.launch_count() - launches0,
                       keep=dict(self._buf), ev_in=torch.cuda.Event(), ev_done=torch.cuda.Event())
            self._graphs[key] = ent
        main = torch.cuda.current_stream()
        if host_input:
            # H2D of the feature map on the copy stream, overlapped with the feature-independent graph
            # (the small metadata copy goes first: both copies share the H2D engine, and graph `pre` needs it)
            _, _, _, counts, _ = self._upload_meta(proposal_list, img_metas)
            self._copy.wait_event(ent['ev_done'])            # the previous step may still be reading the buffer
            self._copy.wait_stream(main)
            with torch.cuda.stream(self._copy):
                ent['feat'].copy_(feat, non_blocking=True)
                ent['ev_in'].record(self._copy)
            ent['graphs'][0].replay()
            main.wait_event(ent['ev_in'])
            ent['graphs'][1].replay()
            ent['ev_done'].record(main)
        else:
            ent['feat'].copy_(feat, non_blocking=True)
            _, _, _, counts, _ = self._upload_meta(proposal_list, img_metas)
            ent['graphs'][0].replay()
        self.graph_launches += ent['launches']
        out = dict(ent['out'])
        out['num_per_view'] = counts
        return out

    # ------------------------------------------------------------------ stage-level entries (the plugin modules' own forwards)
    def query_embedding(self, ref):
        """CrossAttentionBoxHead.position_embedding (cross_attention_head.py:199-200): ref [N,3] -> query_pos [N,256]."""
        ref = ref.to(self.device, torch.float32).contiguous().view(-1, 3)
        N, W = ref.shape[0], self.w
        qpos = self._get('se_query_pos', (N, 256))
        ws = self._get('se_qe_ws', (max(N, 1) * 640,))
        lib.check(self.lib.mv2d_query_embedding(ref.data_ptr(), N, W.p('w_qe0'), W.p('b_qe0'), W.p('w_qe2'), W.p('b_qe2'), W.p('dim_t'),
                                                qpos.data_ptr(), ws.data_ptr(), lib.stream_ptr()), 'mv2d_query_embedding')
        return qpos

    def to_tokens(self, x, x2=None, name='se_tok'):
        """[B,C,h,w] (+ a second map added on the way) -> channels-last rows [B,h*w,C] (``mv2d_nchw_add_to_nhwc``)."""
        x = x.to(self.device, torch.float32).contiguous()
        Bn, Cc, h, w = x.shape
        out = self._get(name, (Bn, h * w, Cc))
        x2p = x2.to(self.device, torch.float32).contiguous().data_ptr() if x2 is not None else None
        lib.check(self.lib.mv2d_nchw_add_to_nhwc(x.data_ptr(), x2p, out.data_ptr(), Bn, Cc, h * w, lib.stream_ptr()), 'mv2d_nchw_add_to_nhwc')
        return out

    def query_generator(self, x, intrinsics, extrinsics, intrins_feat):
        """QueryGenerator.forward on its own (utils/query_generator.py:343-405): x [N,256,7,7] RoI features, intrinsics
        [N,4,4] (K' of get_box_params, fp64), extrinsics [N,4,4], intrins_feat [N,16] -> dict(center_lidar [N,3], enc
        [N,256], ref [N,3], query_pos [N,256])."""
        N, W, c = x.shape[0], self.w, self.cfg
        tok = self.to_tokens(x, name='se_qg_tok')
        kroi = intrinsics.to(self.device, torch.float64).contiguous().view(N, 16)
        eroi = extrinsics.to(self.device, torch.float64).contiguous().view(N, 16)
        ifeat = intrins_feat.to(self.device, torch.float32).contiguous().view(N, 16)
        center, ref, qpos, enc = (self._get('se_center', (N, 3)), self._get('se_ref', (N, 3)), self._get('se_qpos', (N, 256)),
                                  self._get('se_enc', (N, 256)))
        ws_bytes = self.lib.mv2d_roi_align_qg_workspace_bytes(N)
        ws = self._get('qg_ws', (ws_bytes // 4,))
        p = lib.QgParams()
        p.N, p.V, p.h, p.w, p.stride, p.phase = N, 1, 1, 1, c['stride'], 3
        p.pc_range = (C.c_float * 6)(*c['pc_range'])
        p.intrins_feat_scale = c['intrins_feat_scale']
        p.dim_t = W.p('dim_t')
        for f in ('w_conv', 'b_conv', 'w_conv_lo', 'w_fc', 'b_fc', 'w_enc0', 'b_enc0', 'w_enc2', 'b_enc2', 'w_center',
                  'b_center', 'w_qe0', 'b_qe0', 'w_qe2', 'b_qe2', 'w_fc_hi', 'w_fc_lo', 'w_enc0_hi', 'w_enc0_lo', 'w_enc2_hi',
                  'w_enc2_lo', 'w_qe0_hi', 'w_qe0_lo', 'w_qe2_hi', 'w_qe2_lo'):
            setattr(p, f, W.p(f))
        p.tok_feat, p.roi_intrinsics, p.roi_extrinsics, p.intrins_feat = tok.data_ptr(), kroi.data_ptr(), eroi.data_ptr(), ifeat.data_ptr()
        p.center_lidar, p.ref, p.query_pos, p.enc_out = center.data_ptr(), ref.data_ptr(), qpos.data_ptr(), enc.data_ptr()
        p.workspace, p.workspace_bytes = ws.data_ptr(), ws_bytes
        lib.check(self.lib.mv2d_roi_align_qg(C.byref(p), lib.stream_ptr()), 'mv2d_roi_align_qg')
        return dict(center_lidar=center, enc=enc, ref=ref, query_pos=qpos)

    def transformer(self, x, mask, query_embed, pos_embed, ref=None, attn_mask=None, cross_attn_mask=None, vel_dt=0.0):
        """MV2DTransformer.forward / CrossAttentionBoxHead.forward on the reference's dense interface
        (cross_attention_head.py:22-49, 202-242).  x / pos_embed [bs,n,C,h,w], mask [bs,n,h,w] (True = padded key),
        query_embed [bs,nq,C].  Two layouts, as the two heads call it:
          S head: bs = number of queries, nq = 1, memory of query i = its n gathered RoI feature blocks (h = w = 7);
          T head: bs = 1, memory = the n feature views, cross_attn_mask [nq,n,h,w] (True = masked).
        Returns (outs_dec [L,nq_total,256], cls_scores [L,nq_total,10], bbox_preds [L,nq_total,10]); the branch outputs
        need ``ref`` [nq_total,3]."""
        bs, n, Cc, h, w = x.shape
        dev = self.device
        mem = self.to_tokens(x.reshape(bs * n, Cc, h, w), name='se_mem')                       # [bs*n, h*w, C]
        kin = self.to_tokens(x.reshape(bs * n, Cc, h, w), pos_embed.reshape(bs * n, Cc, h, w), name='se_kin')
        qpos = query_embed.to(dev, torch.float32).reshape(-1, 256).contiguous()
        N = qpos.shape[0]
        ref_t = (ref.to(dev, torch.float32).reshape(-1, 3).contiguous() if ref is not None else self._get('se_ref0', (N, 3)).fill_(0.5))
        qg = dict(query_pos=qpos, ref=ref_t)
        sa_mask = attn_mask.to(dev).to(torch.uint8).contiguous() if attn_mask is not None else None
        if self.mode == 'S':
            assert h * w == 49 and query_embed.shape[1] == 1, 'the single-frame head attends to gathered 7x7 RoI blocks, one query per batch entry'
            valid = ~mask.to(dev).reshape(bs, n, h * w).all(-1)                                  # [N, n] RoI slot is real
            cnt = valid.sum(1).to(torch.int32)
            order = torch.argsort((~valid).to(torch.int8), dim=1, stable=True).to(torch.int32)   # real slots first, in order
            match = (order + torch.arange(bs, device=dev, dtype=torch.int32)[:, None] * n).contiguous()
            corr = dict(match=match, match_cnt=cnt.contiguous(), max_match=n)
            out = self.decoder(qg, corr, kin.view(-1, 256), mem.view(-1, 256), N, self_attn_mask=sa_mask)
        else:
            assert bs == 1, 'the two-frame head calls the transformer with one batch entry'
            R = n * h * w
            words = (R + 31) // 32
            keep = torch.ones((N, R), dtype=torch.bool, device=dev)
            if cross_attn_mask is not None:
                keep &= ~cross_attn_mask.to(dev).reshape(N, R)
            keep &= ~mask.to(dev).reshape(1, R)
            bits = torch.zeros((N, words * 32), dtype=torch.int64, device=dev)
            bits[:, :R] = keep
            packed = (bits.view(N, words, 32) << torch.arange(32, device=dev, dtype=torch.int64)).sum(-1)
            keymask = packed.to(torch.int32).contiguous()                                       # bit c of word c/32 (wraps to the sign bit)
            corr = dict(keymask=keymask, mask_words=words, key_list=None, key_cnt=None)
            out = self.decoder(qg, corr, kin.view(-1, 256), mem.view(-1, 256), N, vel_dt=vel_dt, self_attn_mask=sa_mask, grid=(h, w))
        cls, box, outs = out
        return outs, cls, box

    # ------------------------------------------------------------------ batches (a segment dimension through ONE kernel chain)
    def _upload_meta_batch(self, proposal_lists, metas_list, bucket=1):
        """Metadata of B samples in one pinned staging buffer and one H2D copy.  Layout of the rows (C ABI, "Batches"):
        sample b owns query rows [b*Np, (b+1)*Np), Np = max_b n_b rounded up to ``bucket``; the first n_b are its
        detections (sorted by view, view index = b*V + v), the rest padding rows carrying the reference's dummy box
        (mv2d_s_head.py:124-127), whose results are dropped.  Returns (cams [3,B*V,16], rois [B*Np,5],
        roi_start [B,V+1], batch descriptor)."""
        B, V = len(metas_list), len(metas_list[0])
        assert all(len(m) == V for m in metas_list), 'every sample of a batch needs the same number of views'
        plists = []
        for pl in proposal_lists:
            if sum(len(p) for p in pl) == 0:
                p0 = torch.tensor([[0, 50, 50, 100, 100, 0]], dtype=torch.float32, device=pl[0].device)
                pl = [p0] + list(pl[1:])
            plists.append(pl)
        counts = [[int(p.shape[0]) for p in pl] for pl in plists]
        n_b = [sum(c) for c in counts]
        if max(max(c) for c in counts) > 256:
            raise RuntimeError('box correlation handles at most 256 detections per view (include/mv2d_b200.h, limits)')
        Np = (max(n_b) + bucket - 1) // bucket * bucket
        N = B * Np
        cam_b, roi_b, st_b, nr_b = 3 * B * V * 16 * 8, N * 5 * 4, B * (V + 1) * 4, B * 4
        roi_off = cam_b
        st_off = (roi_off + roi_b + 7) // 8 * 8
        nr_off = st_off + st_b
        vd_off = nr_off + nr_b
        total = vd_off + B * 4
        pin = self._pin.get('meta_b')
        if pin is None or pin.numel() < total:
            assert pin is None or not self._graphs_b, 'the batch metadata buffer cannot grow once graphs were captured'
            pin = torch.empty(max(total, 1 << 20), dtype=torch.uint8).pin_memory()
            self._pin['meta_b'] = pin
            self._pin['meta_b_ev'] = None
            self._buf.pop('meta_b', None)
        elif self._pin['meta_b_ev'] is not None:
            self._pin['meta_b_ev'].synchronize()       # the previous copy out of this staging buffer has run
        dev = self._get('meta_b', (pin.numel(),), torch.uint8)
        host = pin.numpy()
        cams = host[:cam_b].view(np.float64).reshape(3, B * V, 16)
        rh = host[roi_off:roi_off + roi_b].view(np.float32).reshape(N, 5)
        sh = host[st_off:st_off + st_b].view(np.int32).reshape(B, V + 1)
        on_device = plists[0][0].is_cuda
        for b, metas in enumerate(metas_list):
            for v, m in enumerate(metas):
                cams[0, b * V + v] = np.asarray(m['lidar2img'], dtype=np.float64).reshape(16)
                cams[1, b * V + v] = np.asarray(m['intrinsics'], dtype=np.float64).reshape(16)
                cams[2, b * V + v] = np.asarray(m['extrinsics'], dtype=np.float64).reshape(16)
            o = b * Np
            for v, p in enumerate(plists[b]):
                n = counts[b][v]
                rh[o:o + n, 0] = b * V + v
                if not on_device:
                    rh[o:o + n, 1:] = p[:, :4].float().numpy()
                o += n
            rh[o:(b + 1) * Np] = (b * V, 50, 50, 100, 100)          # padding rows
            sh[b] = b * Np + np.concatenate([[0], np.cumsum(counts[b])])
        host[nr_off:nr_off + nr_b].view(np.int32)[:] = n_b
        host[vd_off:vd_off + B * 4].view(np.float32)[:] = [self._vel_dt(m) for m in metas_list]
        dev[:total].copy_(pin[:total], non_blocking=True)
        if self._pin['meta_b_ev'] is None:
            self._pin['meta_b_ev'] = torch.cuda.Event()
        self._pin['meta_b_ev'].record()
        d_cams = dev[:cam_b].view(torch.float64).view(3, B * V, 16)
        d_rois = dev[roi_off:roi_off + roi_b].view(torch.float32).view(N, 5)
        d_start = dev[st_off:st_off + st_b].view(torch.int32)
        d_nreal = dev[nr_off:nr_off + nr_b].view(torch.int32)
        d_vel = dev[vd_off:vd_off + B * 4].view(torch.float32)
        if on_device:
            for b in range(B):
                d_rois[b * Np:b * Np + n_b[b], 1:] = torch.cat([p[:, :4].float() for p in plists[b]], 0)
        h, w = self._batch_grid
        keys = [self._masks(m, h, w)[0] for m in metas_list]
        same = all(k == keys[0] for k in keys)
        mk = tuple(keys)
        ent = self._mask_cache.get(('batch', mk))
        if ent is None:
            per = [self._masks(m, h, w) for m in metas_list]
            ent = (mk, torch.cat([e[1] for e in per], 0).contiguous(), torch.cat([e[2] for e in per], 0).contiguous(),
                   any(e[3] for e in per))
            self._mask_cache[('batch', mk)] = ent
        batch = dict(B=B, Np=Np, Vs=V, n_b=n_b, counts=counts, n_real=d_nreal, metas=metas_list, masks=ent, same_masks=same,
                     vel_dt=d_vel if self.mode == 'T' else None)
        return d_cams, d_rois, d_start, batch

    def _enqueue_batch(self, feat, cams, rois, roi_start, batch):
        """All stages of a batch on the current stream (capturable)."""
        B, V, Np = batch['B'], batch['Vs'], batch['Np']
        i2l = self._get('img2lidar', (B * V, 16), torch.float64)
        trans = self._get('trans', (B, V, V, 16), torch.float64)
        lib.check(self.lib.mv2d_geom_prep_batch(cams[0].data_ptr(), B, V, i2l.data_ptr(), trans.data_ptr(), lib.stream_ptr()),
                  'mv2d_geom_prep_batch')
        return self._enqueue_post(feat, False, cams, rois, roi_start, batch['counts'], B * Np, batch['metas'][0], i2l, trans,
                                  pe_phase=0, batch=batch)

    @torch.no_grad()
    def forward_batch(self, feats, proposal_lists, metas_list, use_graph=False, bucket=1):
        """B samples through ONE kernel chain (the reference asserts B == 1: detectors/mv2d.py:143,
        roi_heads/mv2d_head.py:210,251; here the batch is a segment dimension -- see "Batches" in include/mv2d_b200.h).
        feats [B,V,256,h,w] fp32 (device, or pinned host memory) or a list of B [V,256,h,w] tensors; proposal_lists: B lists
        of V tensors [n_v, >= 4]; metas_list: B lists of V dicts.  Returns a dict of batched tensors -- cls_scores /
        bbox_preds [L, B*Np, 10] with sample b in rows [b*Np, b*Np + n_b[b]) -- plus ``samples``: per-sample dicts of views
        (cls_scores [L,n_b,10], bbox_preds, outs_dec, ref, query_pos, rois, N).
        ``bucket`` rounds Np up (a CUDA graph is captured per (B, Np): see ``forward(..., bucket=)``)."""
        if isinstance(feats, (list, tuple)):
            feats = torch.stack(list(feats), 0)
        B, V, Cc, h, w = feats.shape
        assert B == len(proposal_lists) == len(metas_list) and Cc == 256
        self._batch_grid = (h, w)
        host_input = not feats.is_cuda

        def result(out, batch):
            Np, n_b = batch['Np'], batch['n_b']
            out = dict(out)
            out.update(B=B, Np=Np, n_b=n_b)
            keys = ('cls_scores', 'bbox_preds', 'outs_dec')
            out['samples'] = [dict({k: out[k][:, b * Np:b * Np + n_b[b]] for k in keys},
                                   **{k: out[k][b * Np:b * Np + n_b[b]] for k in ('ref', 'query_pos', 'rois', 'center_lidar',
                                                                                  'tok_feat', 'roi_intrinsics', 'match', 'match_cnt')
                                      if out.get(k) is not None},
                                   N=n_b[b], num_per_view=batch['counts'][b]) for b in range(B)]
            return out

        if not use_graph:
            feat = feats.to(self.device, torch.float32, non_blocking=True).contiguous().view(B * V, Cc, h, w)
            cams, rois, roi_start, batch = self._upload_meta_batch(proposal_lists, metas_list, bucket)
            return result(self._enqueue_batch(feat, cams, rois, roi_start, batch), batch)
        n_b = [max(sum(int(p.shape[0]) for p in pl), 1) for pl in proposal_lists]
        Np = (max(n_b) + bucket - 1) // bucket * bucket
        mkey = tuple(self._masks(m, h, w)[0] for m in metas_list)
        key = (B, Np, tuple(feats.shape), mkey, host_input)
        ent = self._graphs_b.get(key)
        if ent is None:
            static_feat = torch.empty((B * V, Cc, h, w), dtype=torch.float32, device=self.device)
            static_feat.copy_(feats.view(B * V, Cc, h, w), non_blocking=True)
            cams, rois, roi_start, batch = self._upload_meta_batch(proposal_lists, metas_list, bucket)
            self._enqueue_batch(static_feat, cams, rois, roi_start, batch)      # warm-up: sizes buffers, sets attributes
            torch.cuda.synchronize()
            launches0 = self.launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self._enqueue_batch(static_feat, cams, rois, roi_start, batch)
            ent = dict(graph=g, out=out, feat=static_feat, launches=self.launch_count() - launches0, keep=dict(self._buf),
                       ev_in=torch.cuda.Event(), ev_done=torch.cuda.Event())
            self._graphs_b[key] = ent
        main = torch.cuda.current_stream()
        _, _, _, batch = self._upload_meta_batch(proposal_lists, metas_list, bucket)
        if host_input:
            self._copy.wait_event(ent['ev_done'])
            self._copy.wait_stream(main)
            with torch.cuda.stream(self._copy):
                ent['feat'].copy_(feats.view(B * V, Cc, h, w), non_blocking=True)
                ent['ev_in'].record(self._copy)
            main.wait_event(ent['ev_in'])
        else:
            ent['feat'].copy_(feats.view(B * V, Cc, h, w), non_blocking=True)
        ent['graph'].replay()
        ent['ev_done'].record(main)
        self.graph_launches += ent['launches']
        return result(ent['out'], batch)

    LOSS_DEFAULTS = dict(   # configs/mv2d/exp/mv2d_r50_frcnn_single_frame_roi_1408x512_ep72.py:87-95,132-137
        code_weights=[1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.5, 1.5, 2.0, 2.0],
        cls_loss_weight=2.0, focal_gamma=2.0, focal_alpha=0.25, bbox_loss_weight=0.25,
        cls_cost_weight=2.0, reg_cost_weight=0.25)

    @torch.no_grad()
    def loss(self, cls_scores, bbox_preds, gt_boxes, gt_labels, dn_cls=None, dn_box=None, dn_labels=None,
             neg_bbox_loss=None, **over):
        """Next row f3: Hungarian targets + focal / L1 losses of every decoder layer (and the denoising losses) in
        one device call, no host round trip (mv2d_s_head.py:278-299 -> bbox_head.loss / dn_loss_single).
        cls_scores / bbox_preds [L,N,10] (views with a layer stride are fine), gt_boxes [G,9] = (gravity centre,
        w, l, h, yaw, vx, vy), gt_labels [G].  Returns dict(loss_cls [L], loss_bbox [L], dn_loss_cls [L],
        dn_loss_bbox [L], assigned [L,N] int32 with -1 = background).  Forward values only."""
        c = dict(self.LOSS_DEFAULTS, **over)
        dev = self.device
        L, N = cls_scores.shape[0], cls_scores.shape[1]

        def rows(t):        # [L, M, 10] with contiguous rows; the layer stride may be larger than M*10
            assert t.dim() == 3 and t.shape[2] == 10 and t.dtype == torch.float32 and t.is_cuda
            if t.shape[1] > 0 and (t.stride(2) != 1 or t.stride(1) != 10):
                t = t.contiguous()
            return t, (t.stride(0) if t.shape[0] > 1 else t.shape[1] * 10)
        cls_scores, ls = rows(cls_scores)
        bbox_preds, ls2 = rows(bbox_preds)
        if ls != ls2:       # one layer stride for both: fall back to packed copies
            (cls_scores, ls), (bbox_preds, ls2) = rows(cls_scores.contiguous()), rows(bbox_preds.contiguous())
        gt = gt_boxes.to(dev, torch.float32).contiguous().view(-1, 9)
        lab = gt_labels.to(dev).to(torch.int32).contiguous()
        G = gt.shape[0]
        p = lib.LossParams()
        p.N, p.G, p.L, p.num_classes = N, G, L, self.cfg['num_classes']
        p.layer_stride = ls
        for k in ('cls_cost_weight', 'reg_cost_weight', 'cls_loss_weight', 'bbox_loss_weight', 'focal_alpha', 'focal_gamma'):
            setattr(p, k, c[k])
        p.dn_split = self.cfg['denoise_split']
        # exp/mv2d_r50_frcnn_two_frames_1408x512_ep72.py:45 turns neg_bbox_loss on, the S head default is off
        p.neg_bbox_loss = int((self.mode == 'T') if neg_bbox_loss is None else neg_bbox_loss)
        p.code_weights = (C.c_float * 10)(*c['code_weights'])
        p.cls_scores, p.bbox_preds = cls_scores.data_ptr(), bbox_preds.data_ptr()
        p.gt_boxes, p.gt_labels = gt.data_ptr(), lab.data_ptr()
        keep = [gt, lab, cls_scores, bbox_preds]
        if dn_cls is not None and dn_cls.shape[1] > 0:
            dn_cls, ds = rows(dn_cls)
            dn_box, ds2 = rows(dn_box)
            if ds != ds2:
                (dn_cls, ds), (dn_box, ds2) = rows(dn_cls.contiguous()), rows(dn_box.contiguous())
            dl = dn_labels.to(dev).to(torch.int32).contiguous()
            p.pad, p.dn_layer_stride = dn_cls.shape[1], ds
            p.dn_cls, p.dn_box, p.dn_labels = dn_cls.data_ptr(), dn_box.data_ptr(), dl.data_ptr()
            keep += [dn_cls, dn_box, dl]
        assigned = torch.empty((L, N), dtype=torch.int32, device=dev)
        losses = torch.empty((L, 4), dtype=torch.float32, device=dev)
        ws_bytes = self.lib.mv2d_loss_workspace_bytes(N, G, L)
        ws = self._get('loss_ws', (ws_bytes // 4 + 1,))
        p.assigned, p.losses, p.workspace, p.workspace_bytes = assigned.data_ptr(), losses.data_ptr(), ws.data_ptr(), ws_bytes
        lib.check(self.lib.mv2d_loss(C.byref(p), lib.stream_ptr()), 'mv2d_loss')
        return dict(loss_cls=losses[:, 0], loss_bbox=losses[:, 1], dn_loss_cls=losses[:, 2], dn_loss_bbox=losses[:, 3],
                    assigned=assigned)

    @torch.no_grad()
    def forward_losses(self, feat, proposal_list, img_metas, gt_boxes, gt_labels, rand=None, use_denoise=None,
                       stage_loss_weights=None, denoise_weight=1.0):
        """Training-mode forward + the loss dict of MV2DSHead.forward_train (mv2d_s_head.py:262-307), forward values
        only (there are no backward kernels yet): the hot path with denoising queries when ``use_denoise`` (default:
        the two-frame head, as in the exp configs), then ``loss``.  Keys follow the reference: ``l{i}.loss_cls``,
        ``l{i}.loss_bbox``, ``l{i}.dn_loss_cls``, ``l{i}.dn_loss_bbox``, each times stage_loss_weights[i]
        (default 0.1 per layer, exp/...:131)."""
        use_denoise = (self.mode == 'T') if use_denoise is None else use_denoise
        dn = dict(gt_boxes=gt_boxes, gt_labels=gt_labels, rand=rand) if (use_denoise and gt_boxes.shape[0] > 0) else None
        out = self.forward(feat, proposal_list, img_metas, dn=dn)
        kw = {}
        if dn is not None:
            kw = dict(dn_cls=out['dn_cls_scores'], dn_box=out['dn_bbox_preds'], dn_labels=out['dn_labels'])
        l = self.loss(out['cls_scores'], out['bbox_preds'], gt_boxes, gt_labels, **kw)
        w = stage_loss_weights or [0.1] * self.L
        losses = {}
        for i in range(self.L):
            if dn is not None:
                losses[f'l{i}.dn_loss_cls'] = l['dn_loss_cls'][i] * denoise_weight * w[i]
                losses[f'l{i}.dn_loss_bbox'] = l['dn_loss_bbox'][i] * denoise_weight * w[i]
            losses[f'l{i}.loss_cls'] = l['loss_cls'][i] * w[i]
            losses[f'l{i}.loss_bbox'] = l['loss_bbox'][i] * w[i]
        return losses, out, l

    @torch.no_grad()
    def decode(self, cls_scores, bbox_preds, max_num=300):
        """NMSFreeCoder.decode_single + z-shift on the device (next row f1)."""
        N = cls_scores.shape[0]
        boxes = torch.empty((max_num, 9), device=self.device)
        scores = torch.empty((max_num,), device=self.device)
        labels = torch.empty((max_num,), dtype=torch.int32, device=self.device)
        valid = torch.empty((max_num,), dtype=torch.uint8, device=self.device)
        post = (C.c_float * 6)(*self.cfg['position_range'])
        lib.check(self.lib.mv2d_nms_free_decode(lib.ptr(cls_scores.contiguous()), lib.ptr(bbox_preds.contiguous()),
                                                N, max_num, post, boxes.data_ptr(), scores.data_ptr(),
                                                labels.data_ptr(), valid.data_ptr(), lib.stream_ptr()),
                  'mv2d_nms_free_decode')
        k = min(max_num, N * 10)
        m = valid[:k].bool()
        return boxes[:k][m], scores[:k][m], labels[:k][m].long()

    @torch.no_grad()
    def handoff_2d(self, detections, gts=None, min_bbox_size=0.0, complement_thr=-1.0):
        """Next row f2 on the device (``mv2d_handoff_2d``): per view, detections [n_v,6] filtered by the minimum box size,
        then (complement_thr > 0) the 2D ground truth ``gts`` [m_v,6] that no kept detection covers appended
        (detectors/mv2d.py:60-117).  Lists of per-view tensors in (device or host), list of per-view DEVICE tensors out;
        the boxes never leave the device -- only the V result counts are read back, where the reference's boolean
        indexing synchronises as well."""
        dev, V = self.device, len(detections)
        dets = [d.to(dev, torch.float32).reshape(-1, 6) for d in detections]
        det = torch.cat(dets, 0).contiguous() if V else torch.zeros((0, 6), device=dev)
        dstart = torch.tensor(np.concatenate([[0], np.cumsum([d.shape[0] for d in dets])]), dtype=torch.int32, device=dev)
        use_gt = complement_thr > 0 and gts is not None
        if use_gt:
            g = [x.to(dev, torch.float32).reshape(-1, 6) for x in gts]
            gt = torch.cat(g, 0).contiguous()
            gstart = torch.tensor(np.concatenate([[0], np.cumsum([x.shape[0] for x in g])]), dtype=torch.int32, device=dev)
            gcounts = [x.shape[0] for x in g]
        else:
            gt, gstart, gcounts = None, torch.zeros(V + 1, dtype=torch.int32, device=dev), [0] * V
        out = torch.empty((det.shape[0] + (gt.shape[0] if use_gt else 0) + 1, 6), device=dev)
        cnt = torch.empty((V,), dtype=torch.int32, device=dev)
        lib.check(self.lib.mv2d_handoff_2d(det.data_ptr() if det.numel() else None, dstart.data_ptr(),
                                           gt.data_ptr() if (use_gt and gt.numel()) else None, gstart.data_ptr(), V,
                                           float(min_bbox_size), float(complement_thr if use_gt else -1.0), out.data_ptr(),
                                           cnt.data_ptr(), lib.stream_ptr()), 'mv2d_handoff_2d')
        counts = cnt.tolist()
        res, o = [], 0
        for v in range(V):
            res.append(out[o:o + counts[v]])
            o += dets[v].shape[0] + gcounts[v]
        return res

    @torch.no_grad()
    def scene_nms(self, boxes, scores, labels, score_thr=0.0, nms_thr=1.0, max_num=300):
        """Scene-level tail of MV2D.simple_test (detectors/mv2d.py:266-282, mmdet3d box3d_multiclass_nms at the configs'
        nms_thr = 1.0): regroup the decoded boxes by class / descending score, cap at max_num.  Device in, device out."""
        n = int(boxes.shape[0])
        b = boxes.to(self.device, torch.float32).contiguous()
        s = scores.to(self.device, torch.float32).contiguous()
        l = labels.to(self.device).to(torch.int32).contiguous()
        ob = torch.empty((max_num, 9), device=self.device)
        os_ = torch.empty((max_num,), device=self.device)
        ol = torch.empty((max_num,), dtype=torch.int32, device=self.device)
        cnt = torch.zeros((1,), dtype=torch.int32, device=self.device)
        lib.check(self.lib.mv2d_scene_nms(b.data_ptr() if n else None, s.data_ptr() if n else None, l.data_ptr() if n else None,
                                          None, n, score_thr, nms_thr, max_num, ob.data_ptr(), os_.data_ptr(), ol.data_ptr(),
                                          cnt.data_ptr(), lib.stream_ptr()), 'mv2d_scene_nms')
        k = int(cnt.item())         # the result leaves the device here anyway (bbox3d2result, mv2d.py:284-287)
        return ob[:k], os_[:k], ol[:k].long()
