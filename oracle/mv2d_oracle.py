"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product path (mv2d_b200/).

CPU restatement, in plain torch (fp32; fp64 where the reference says ``.double()``), of the
MV2D decoder hot path.  One function per stage so that each
CUDA kernel has a stage-level checker.  Every function cites the reference lines it follows
(paths relative to /root/reference/mmdet3d_plugin/models/).

Parity pinning: the reference has no tests or golden vectors, so this
restatement is pinned against the reference's OWN Python run unmodified in the build
container through ``oracle/ref_shim.py``: ``oracle/make_golden.py`` wrote the outputs of that
run to ``tests/golden/*.npz`` and ``tests/test_oracle_golden.py`` checks this file against them
(and, when /root/reference is present, ``tests/test_reference_parity_cpu.py`` re-runs the
reference live).  Third-party semantics (mmcv RoIAlign / BaseTransformerLayer / FFN, mmdet
inverse_sigmoid / bbox2roi, torch MultiheadAttention) are restated per SURVEY.md App. A.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

DEFAULT_CFG = dict(
    pc_range=[-51.2, -51.2, -5.0, 51.2, 51.2, 3.0],
    position_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0],
    depth_num=64, depth_start=1.0, stride=16, roi_size=7, embed=256, heads=8,
    num_layers=6, intrins_feat_scale=0.1,
    # BoxCorrelation (roi_heads/utils/box_correlation.py:12-13 + exp configs)
    sample_size=4, corr_num_depth=8, corr_depth_start=0.5, corr_depth_end=70.0,
    topk=1, iou_thr=0.0, ratio=0.0, expand_stride=0,
    num_views_per_frame=6,
)


def make_cfg(mode='S', **over):
    cfg = dict(DEFAULT_CFG)
    if mode == 'T':  # exp/mv2d_r50_frcnn_two_frames_1408x512_ep72.py:121-124
        cfg.update(topk=20, expand_stride=2, denoise_noise_scale=1.25, denoise_split=0.6)   # :44-47
    cfg['mode'] = mode
    cfg.update(over)
    return cfg


# ----------------------------------------------------------------------------- small helpers
def inverse_sigmoid(x, eps=1e-5):
    """mmdet.models.utils.transformer.inverse_sigmoid."""
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def bbox2roi(proposal_list):
    """mmdet.core.bbox2roi: per view i with n>0 boxes -> [i, x1, y1, x2, y2]."""
    out = []
    for i, b in enumerate(proposal_list):
        if b.shape[0] > 0:
            out.append(torch.cat([b.new_full((b.shape[0], 1), i), b[:, :4]], dim=-1))
        else:
            out.append(b.new_zeros((0, 5)))
    return torch.cat(out, 0)


def guard_empty(proposal_list):
    """roi_heads/mv2d_s_head.py:124-127 -- inject one dummy box when nothing was detected."""
    if sum(len(p) for p in proposal_list) == 0:
        p0 = torch.tensor([[0, 50, 50, 100, 100, 0]], dtype=proposal_list[0].dtype)
        proposal_list = [p0] + list(proposal_list[1:])
    return proposal_list


def lid_depths(num, start, end, dtype):
    """LID depth bins (utils/pe.py:96-100; box_correlation.py:221-225)."""
    idx = torch.arange(num, dtype=dtype)
    bin_size = (end - start) / (num * (1 + num))
    return start + bin_size * idx * (idx + 1)


# ----------------------------------------------------------------------------- a1-a3: PE
def feat_masks(img_metas, h, w):
    """Padding mask at feature resolution (utils/pe.py:146-155): ones outside img_shape,
    nearest-neighbour F.interpolate to (h, w)."""
    V = len(img_metas)
    pad_h, pad_w, _ = img_metas[0]['pad_shape']
    masks = torch.ones((1, V, pad_h, pad_w))
    for v in range(V):
        ih, iw, _ = img_metas[v]['img_shape']
        masks[0, v, :ih, :iw] = 0
    return F.interpolate(masks, size=(h, w)).to(torch.bool)


def pe_coords(img_metas, h, w, cfg):
    """Frustum coordinates -> normalised -> inverse_sigmoid (utils/pe.py:84-130).
    Returns [V, 3*D, h, w] float32, channel = d*3 + xyz."""
    pr = cfg['position_range']
    D = cfg['depth_num']
    pad_h, pad_w, _ = img_metas[0]['pad_shape']
    V = len(img_metas)
    coords_h = (torch.arange(h, dtype=torch.float64) + 0.5) * pad_h / h - 0.5
    coords_w = (torch.arange(w, dtype=torch.float64) + 0.5) * pad_w / w - 0.5
    coords_d = lid_depths(D, cfg['depth_start'], pr[3], torch.float64)
    gw, gh, gd = torch.meshgrid(coords_w, coords_h, coords_d, indexing='ij')
    coords = torch.stack([gw, gh, gd, torch.ones_like(gw)], dim=-1)  # [W,H,D,4]
    coords[..., :2] = coords[..., :2] * torch.maximum(coords[..., 2:3],
                                                     torch.full_like(coords[..., 2:3], 1e-3))
    img2lidar = torch.from_numpy(np.asarray(
        [np.linalg.inv(m['lidar2img']) for m in img_metas])).double().to(coords.device)  # host inverse, pe.py:111
    c3 = torch.matmul(img2lidar.view(V, 1, 1, 1, 4, 4), coords.view(1, w, h, D, 4, 1))
    c3 = c3.squeeze(-1)[..., :3]  # [V,W,H,D,3]
    for i in range(3):
        c3[..., i] = (c3[..., i] - pr[i]) / (pr[i + 3] - pr[i])
    c3 = c3.permute(0, 3, 4, 2, 1).contiguous().view(V, D * 3, h, w)
    return inverse_sigmoid(c3).float()


def sine_pos_3d(mask, stride, num_feats=128, temperature=10000, scale=2 * math.pi, eps=1e-6):
    """SinePositionalEncoding3D.forward with normalize=True, offset=0
    (utils/positional_encoding.py:58-96).  mask [B,V,h,w] bool -> [B,V,3*num_feats,h,w]."""
    not_mask = 1 - mask.to(torch.int)
    n_embed = not_mask.cumsum(1, dtype=torch.float32)
    y_embed = not_mask.cumsum(2, dtype=torch.float32)
    x_embed = not_mask.cumsum(3, dtype=torch.float32)
    if stride > 0:
        y_embed = (y_embed - 0.5) * stride
        x_embed = (x_embed - 0.5) * stride
    n_embed = n_embed / (n_embed[:, -1:, :, :] + eps) * scale
    y_embed = y_embed / (y_embed[:, :, -1:, :] + eps) * scale
    x_embed = x_embed / (x_embed[:, :, :, -1:] + eps) * scale
    dim_t = torch.arange(num_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / num_feats)
    B, V, H, W = mask.shape

    def emb(e):
        p = e[..., None] / dim_t
        # NOTE stack at dim=4 (not -1): first half all sines, second half all cosines
        return torch.stack((p[..., 0::2].sin(), p[..., 1::2].cos()), dim=4).view(B, V, H, W, -1)

    pos = torch.cat((emb(n_embed), emb(y_embed), emb(x_embed)), dim=4)
    return pos.permute(0, 1, 4, 2, 3)


def _conv1x1(x, w, b):
    return F.conv2d(x, w, b)


def pe_forward(sd, feat, img_metas, cfg, return_parts=False):
    """PE.forward (utils/pe.py:137-169) with with_fpe=True, adapt_pos3d=True.
    feat [V,256,h,w] -> pos_embed [V,256,h,w]."""
    p = 'position_encoding.'
    V, C, h, w = feat.shape
    masks = feat_masks(img_metas, h, w)
    coords = pe_coords(img_metas, h, w, cfg)
    x = _conv1x1(coords, sd[p + 'position_encoder.0.weight'], sd[p + 'position_encoder.0.bias'])
    x = _conv1x1(F.relu(x), sd[p + 'position_encoder.2.weight'], sd[p + 'position_encoder.2.bias'])
    # SELayer (pe.py:44-48)
    g = _conv1x1(feat, sd[p + 'fpe.conv_reduce.weight'], sd[p + 'fpe.conv_reduce.bias'])
    g = _conv1x1(F.relu(g), sd[p + 'fpe.conv_expand.weight'], sd[p + 'fpe.conv_expand.bias'])
    x = x * torch.sigmoid(g)
    sin = sine_pos_3d(masks, cfg['stride']).flatten(0, 1)
    s = _conv1x1(sin, sd[p + 'adapt_pos3d.0.weight'], sd[p + 'adapt_pos3d.0.bias'])
    s = _conv1x1(F.relu(s), sd[p + 'adapt_pos3d.2.weight'], sd[p + 'adapt_pos3d.2.bias'])
    out = x + s
    if return_parts:
        return out, dict(coords=coords, sine=sin, sine_branch=s)
    return out


# ----------------------------------------------------------------------------- a4: RoIAlign
def roi_align(x, rois, out_size=7, spatial_scale=1.0 / 16):
    """mmcv RoIAlign avg, aligned=True, sampling_ratio<=0 (adaptive) -- SURVEY App. A.
    x [V,C,H,W], rois [N,5] (view,x1,y1,x2,y2 px) -> [N,C,out,out].  Plain loops over RoIs
    (oracle clarity over speed)."""
    V, C, H, W = x.shape
    N = rois.shape[0]
    out = x.new_zeros((N, C, out_size, out_size))
    for n in range(N):
        v = int(rois[n, 0])
        # the mmcv kernel works in float32; keep the same rounding for the bin geometry
        x1, y1, x2, y2 = [np.float32(np.float32(t) * np.float32(spatial_scale)) - np.float32(0.5)
                          for t in rois[n, 1:5].float().numpy()]
        rw, rh = np.float32(x2 - x1), np.float32(y2 - y1)
        bw, bh = np.float32(rw / np.float32(out_size)), np.float32(rh / np.float32(out_size))
        gh = int(math.ceil(float(rh) / out_size))
        gw = int(math.ceil(float(rw) / out_size))
        count = max(gh * gw, 1)
        ph = torch.arange(out_size, dtype=torch.float32)
        iy = torch.arange(gh, dtype=torch.float32)
        ix = torch.arange(gw, dtype=torch.float32)
        ys = float(y1) + ph[:, None] * float(bh) + (iy[None, :] + 0.5) * float(bh) / max(gh, 1)
        xs = float(x1) + ph[:, None] * float(bw) + (ix[None, :] + 0.5) * float(bw) / max(gw, 1)
        ys, xs = ys.reshape(-1), xs.reshape(-1)  # [7*gh], [7*gw]

        def axis(c, size):
            valid = (c >= -1.0) & (c <= size)
            c = c.clamp(min=0)
            lo = c.floor().long()
            hi_edge = lo >= size - 1
            lo = torch.where(hi_edge, torch.full_like(lo, size - 1), lo)
            hi = torch.where(hi_edge, lo, lo + 1)
            c = torch.where(hi_edge, lo.float(), c)
            l = c - lo.float()
            return lo, hi, l, valid

        ylo, yhi, ly, yv = axis(ys, H)
        xlo, xhi, lx, xv = axis(xs, W)
        fm = x[v]  # [C,H,W]
        hy, hx = 1 - ly, 1 - lx
        val = (fm[:, ylo][:, :, xlo] * (hy[:, None] * hx[None, :]) +
               fm[:, ylo][:, :, xhi] * (hy[:, None] * lx[None, :]) +
               fm[:, yhi][:, :, xlo] * (ly[:, None] * hx[None, :]) +
               fm[:, yhi][:, :, xhi] * (ly[:, None] * lx[None, :]))
        val = val * (yv[:, None] & xv[None, :]).float()
        val = val.view(C, out_size, gh, out_size, gw).sum(dim=(2, 4)) / count
        out[n] = val
    return out


# ----------------------------------------------------------------------------- a5: box params
def get_box_params(proposal_list, img_metas, roi_size=7):
    """MV2DHead.get_box_params (roi_heads/mv2d_head.py:51-72): per-RoI intrinsics K' in the
    RoI's 7x7 frame (fp64) and extrinsics."""
    Ks, Es = [], []
    for bbox, meta in zip(proposal_list, img_metas):
        n = bbox.shape[0]
        K = torch.from_numpy(np.asarray(meta['intrinsics'])).double().repeat(n, 1, 1)
        E = torch.from_numpy(np.asarray(meta['extrinsics'])).double().repeat(n, 1, 1)
        wh = bbox[:, 2:4] - bbox[:, :2]  # float32 arithmetic, as in the reference
        scale = wh.new_tensor([roi_size, roi_size])[None] / wh
        K[:, :2, 2] = K[:, :2, 2] - bbox[:, :2] - 0.5 / scale
        K[:, :2] = K[:, :2] * scale[..., None]
        Ks.append(K)
        Es.append(E)
    return torch.cat(Ks, 0), torch.cat(Es, 0)


def process_intrins_feat(rois, intrinsics, scale=0.1, min_size=4):
    """MV2DHead.process_intrins_feat (roi_heads/mv2d_head.py:95-101)."""
    f = intrinsics.view(intrinsics.shape[0], 16).clone().float() * scale
    wh = rois[:, 3:5] - rois[:, 1:3]
    f[(wh < min_size).any(1)] = 0
    return f


# ----------------------------------------------------------------------------- a6-a8: QG
def query_generator_feat(sd, roi_feat, intrins_feat):
    """QueryGenerator.get_roi_feat (roi_heads/utils/query_generator.py:352-374)."""
    p = 'query_generator.'
    x = F.relu(F.conv2d(roi_feat, sd[p + 'shared_convs.0.conv.weight'],
                        sd[p + 'shared_convs.0.conv.bias'], padding=1))
    x = F.avg_pool2d(x, roi_feat.shape[-1]).flatten(1)
    x = F.relu(F.linear(x, sd[p + 'shared_fcs.0.weight'], sd[p + 'shared_fcs.0.bias']))
    x = torch.cat([x, intrins_feat], dim=1).clamp(min=-5e3, max=5e3)
    x = F.relu(F.linear(x, sd[p + 'extra_enc.0.weight'], sd[p + 'extra_enc.0.bias']))
    x = F.relu(F.linear(x, sd[p + 'extra_enc.2.weight'], sd[p + 'extra_enc.2.bias']))
    return x


def center2lidar(center_pred, intrinsic, extrinsic):
    """QueryGenerator.center2lidar (query_generator.py:333-341)."""
    ci = torch.cat([center_pred[:, :2] * center_pred[:, 2:3], center_pred[:, 2:3]], dim=1)
    ch = torch.cat([ci, ci.new_ones((ci.shape[0], 1))], dim=1)
    lidar2img = torch.bmm(intrinsic, extrinsic.transpose(1, 2))
    img2lidar = torch.inverse(lidar2img).float()
    return torch.bmm(img2lidar, ch[..., None])[:, :3, 0]


def query_generator(sd, roi_feat, intrinsics, extrinsics, intrins_feat, cfg):
    """QueryGenerator.forward + reference-point normalisation
    (query_generator.py:343-405; roi_heads/mv2d_s_head.py:146-153 -- the trailing clamp is
    not in-place, i.e. a no-op)."""
    x = query_generator_feat(sd, roi_feat, intrins_feat)
    c = F.linear(x, sd['query_generator.fc_center.weight'], sd['query_generator.fc_center.bias'])
    xyz = center2lidar(c, intrinsics, extrinsics)
    pc = cfg['pc_range']
    ref = torch.stack([(xyz[:, i] - pc[i]) / (pc[i + 3] - pc[i]) for i in range(3)], dim=1)
    return ref, dict(enc=x, center_pred=c, center_lidar=xyz)


# ----------------------------------------------------------------------------- a9-a10: box corr
def _epipolar_points(rois, img_metas, cfg):
    """gen_sample_points_in_rois + gen_epipolar_in_each_view
    (roi_heads/utils/box_correlation.py:196-257).  Returns transformed points
    [N, V, S*S*Dn, 2] float32 and validity [N, V, S*S*Dn]."""
    S, Dn = cfg['sample_size'], cfg['corr_num_depth']
    H, W = img_metas[0]['pad_shape'][:2]
    V = len(img_metas)
    N = rois.shape[0]
    lidar2img = torch.from_numpy(np.stack([np.asarray(m['lidar2img']) for m in img_metas])).double()
    img2lidar = torch.inverse(lidar2img)
    trans = torch.matmul(lidar2img[None], img2lidar[:, None])  # [src, dst, 4, 4]
    lin = torch.linspace(0, 1, S)
    gy, gx = torch.meshgrid(lin, lin, indexing='ij')
    croi = torch.stack([gx, gy], dim=-1)  # [S,S,2]
    wh = rois[:, 3:5] - rois[:, 1:3]
    pts = (rois[:, None, None, 1:3] + wh[:, None, None] * croi[None]).reshape(N * S * S, 2)
    view_ids = rois[:, 0].long().repeat_interleave(S * S)
    idx = torch.arange(Dn).float()
    bin_size = (cfg['corr_depth_end'] - cfg['corr_depth_start']) / (Dn * (1 + Dn))
    depth = (cfg['corr_depth_start'] + bin_size * idx * (idx + 1))  # float32
    P = pts.shape[0]
    p2d = torch.cat([pts[:, None].expand(P, Dn, 2), depth[None, :, None].expand(P, Dn, 1)],
                    dim=-1).double()
    hom = torch.cat([p2d[..., :2] * p2d[..., 2:3], p2d[..., 2:3], p2d.new_ones((P, Dn, 1))], dim=-1)
    tm = trans[view_ids]  # [P, V, 4, 4]
    cam = torch.matmul(tm[:, :, None], hom[:, None, ..., None])[..., :3, 0]  # [P,V,Dn,3]
    tp = cam[..., :2] / cam[..., 2:3].clamp_min(1e-2)
    valid = cam[..., 2] >= cfg['corr_depth_start']
    valid &= (0 <= tp[..., 0]) & (tp[..., 0] <= W - 1) & (0 <= tp[..., 1]) & (tp[..., 1] <= H - 1)
    valid[torch.arange(P), view_ids] = False
    tp = tp.float()
    tp = tp.view(N, S * S, V, Dn, 2).permute(0, 2, 1, 3, 4).reshape(N, V, S * S * Dn, 2)
    valid = valid.view(N, S * S, V, Dn).permute(0, 2, 1, 3).reshape(N, V, S * S * Dn)
    return tp, valid


def _box_iou(a, b, eps=1e-4):
    """BoxCorrelation.box_iou (box_correlation.py:385-398); a [4], b [m,4]."""
    xy0 = torch.maximum(a[None, 0:2], b[:, 0:2])
    xy1 = torch.minimum(a[None, 2:4], b[:, 2:4])
    wh = (xy1 - xy0).clamp(min=0)
    inter = wh[:, 0] * wh[:, 1]
    area_a = (a[2] - a[0]) * (a[3] - a[1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    return inter / (area_a + area_b - inter + eps)


def box_match_lists(rois, num_per_view, img_metas, cfg):
    """epipolar_in_box, ``topk_matched`` mode (box_correlation.py:260-382), restated per RoI:
    for every other view whose RoIs are hit by a valid epipolar sample, the bounding box of
    the valid samples is IoU-matched to that view's RoIs; top-k by IoU, kept if IoU>0 (and
    > ratio*max or > iou_thr).  Returns a python list (per RoI) of matched global RoI ids in
    the reference's order: self first, then by ascending view, descending IoU."""
    N = rois.shape[0]
    if N == 0:
        return []
    V = len(img_metas)
    tp, valid = _epipolar_points(rois, img_metas, cfg)
    starts = np.concatenate([[0], np.cumsum(num_per_view)]).astype(int)
    topk, iou_thr, ratio = cfg['topk'], cfg['iou_thr'], cfg['ratio']
    max_rois = max(num_per_view)
    out = []
    for n in range(N):
        ids = [n]
        for v in range(V):
            nv = num_per_view[v]
            if nv == 0:
                continue
            vm = valid[n, v]
            if not vm.any():
                continue
            pts = tp[n, v]
            rv = rois[starts[v]:starts[v + 1], 1:5]
            hit = ((rv[:, None, 0] <= pts[None, :, 0]) & (pts[None, :, 0] <= rv[:, None, 2]) &
                   (rv[:, None, 1] <= pts[None, :, 1]) & (pts[None, :, 1] <= rv[:, None, 3]))
            hit = (hit & vm[None]).any(-1)
            if not hit.any():
                continue
            pv = pts[vm]
            t_roi = torch.cat([pv.min(0)[0], pv.max(0)[0]])
            iou = _box_iou(t_roi, rv)
            # padded slots of the reference (rois_pad rows of zeros) have iou forced to 0
            iou_pad = torch.zeros(max_rois)
            iou_pad[:nv] = iou
            order = torch.argsort(iou_pad, descending=True, stable=True)[:topk]
            top = iou_pad[order]
            keep = ((top > ratio * top.max()) | (top > iou_thr)) & (top > 0)
            for j, k in zip(order.tolist(), keep.tolist()):
                if k:
                    ids.append(int(starts[v] + j))
        out.append(ids)
    return out


def box_roi_correlation(rois, num_per_view, img_metas, cfg):
    """gen_box_roi_correlation (box_correlation.py:165-193): padded corr [N,M] int64 (pad 0)
    and mask [N,M] bool."""
    lists = box_match_lists(rois, num_per_view, img_metas, cfg)
    N = len(lists)
    if N == 0:
        return torch.zeros((0, 0), dtype=torch.int64), torch.zeros((0, 0), dtype=torch.bool)
    M = max(len(l) for l in lists)
    corr = torch.zeros((N, M), dtype=torch.int64)
    mask = torch.zeros((N, M), dtype=torch.bool)
    for n, l in enumerate(lists):
        corr[n, :len(l)] = torch.tensor(l)
        mask[n, :len(l)] = True
    return corr, mask


def feat_in_rois(rois, V, h, w, stride, expand_stride):
    """Own-view cell mask (box_correlation.py:102-115): cell centre +-(0.5+expand)*stride
    against the box."""
    ys = (torch.arange(h, dtype=torch.float32) + 0.5) * stride - 0.5
    xs = (torch.arange(w, dtype=torch.float32) + 0.5) * stride - 0.5
    m = 0.5 * stride + expand_stride * stride
    inx = (xs[None, :] + m >= rois[:, 1:2]) & (xs[None, :] - m <= rois[:, 3:4])  # [N,w]
    iny = (ys[None, :] + m >= rois[:, 2:3]) & (ys[None, :] - m <= rois[:, 4:5])  # [N,h]
    inb = iny[:, :, None] & inx[:, None, :]
    out = torch.zeros((rois.shape[0], V, h, w), dtype=torch.bool)
    out[torch.arange(rois.shape[0]), rois[:, 0].long()] = inb
    return out


def box_correlation_mask(rois, num_per_view, img_metas, h, w, cfg):
    """gen_box_correlation (box_correlation.py:95-162): per-query key mask [N,V,h,w] =
    OR of the own-view cell masks of all matched RoIs (self included)."""
    V = len(img_metas)
    own = feat_in_rois(rois, V, h, w, cfg['stride'], cfg['expand_stride'])
    lists = box_match_lists(rois, num_per_view, img_metas, cfg)
    out = torch.zeros_like(own)
    for n, l in enumerate(lists):
        out[n] = own[torch.tensor(l)].any(0)
    return out


# ----------------------------------------------------------------------------- a12: query embed
def pos2posemb3d(pos, num_pos_feats=128, temperature=10000):
    """utils/pe.py:21-33 -- interleaved sin/cos, concat order (y, x, z)."""
    pos = pos * (2 * math.pi)
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / num_pos_feats)

    def emb(p):
        p = p[..., None] / dim_t
        return torch.stack((p[..., 0::2].sin(), p[..., 1::2].cos()), dim=-1).flatten(-2)

    return torch.cat((emb(pos[..., 1]), emb(pos[..., 0]), emb(pos[..., 2])), dim=-1)


def query_embed(sd, ref):
    """CrossAttentionBoxHead.position_embedding
    (roi_heads/bbox_heads/cross_attention_head.py:199-200)."""
    p = 'bbox_head.query_embedding.'
    x = pos2posemb3d(ref, 128)
    x = F.relu(F.linear(x, sd[p + '0.weight'], sd[p + '0.bias']))
    return F.linear(x, sd[p + '2.weight'], sd[p + '2.bias'])


# ----------------------------------------------------------------------------- a14-a17: decoder
def _mha(sd, prefix, q, k, v, heads, attn_mask=None, key_padding_mask=None):
    """torch.nn.MultiheadAttention forward (seq-first), through the functional the module
    itself calls.  q [L,B,C], k/v [S,B,C]."""
    C = q.shape[-1]
    out, _ = F.multi_head_attention_forward(
        q, k, v, C, heads, sd[prefix + 'in_proj_weight'], sd[prefix + 'in_proj_bias'],
        None, None, False, 0.0, sd[prefix + 'out_proj.weight'], sd[prefix + 'out_proj.bias'],
        training=False, key_padding_mask=key_padding_mask, need_weights=True,
        attn_mask=attn_mask)
    return out


def decoder_layer(sd, l, query, query_pos, memory, key_pos, cfg, self_mask=None,
                  cross_mask=None, key_padding_mask=None):
    """PETRTransformerDecoderLayer, order (self_attn, norm, cross_attn, norm, ffn, norm)
    (utils/petr_transformer.py:194-311 over mmcv BaseTransformerLayer).
    query/query_pos [nq,bs,C]; memory/key_pos [nk,bs,C]."""
    p = f'bbox_head.transformer.decoder.layers.{l}.'
    heads, C = cfg['heads'], query.shape[-1]
    nq, bs, _ = query.shape
    # FlattenMHSelfAttention (petr_transformer.py:314-370): ONE sequence over all nq*bs queries
    qk = (query + query_pos).view(nq * bs, 1, C)
    sa = _mha(sd, p + 'attentions.0.attn.', qk, qk, query.view(nq * bs, 1, C), heads,
              attn_mask=self_mask)
    query = query + sa.view(nq, bs, C)
    query = F.layer_norm(query, (C,), sd[p + 'norms.0.weight'], sd[p + 'norms.0.bias'], 1e-5)
    # PETRMultiheadAttention (petr_transformer.py:373-513)
    ca = _mha(sd, p + 'attentions.1.attn.', query + query_pos, memory + key_pos, memory, heads,
              attn_mask=cross_mask, key_padding_mask=key_padding_mask)
    query = query + ca
    query = F.layer_norm(query, (C,), sd[p + 'norms.1.weight'], sd[p + 'norms.1.bias'], 1e-5)
    # mmcv FFN: x + W2 relu(W1 x)
    hdn = F.relu(F.linear(query, sd[p + 'ffns.0.layers.0.0.weight'], sd[p + 'ffns.0.layers.0.0.bias']))
    query = query + F.linear(hdn, sd[p + 'ffns.0.layers.1.weight'], sd[p + 'ffns.0.layers.1.bias'])
    query = F.layer_norm(query, (C,), sd[p + 'norms.2.weight'], sd[p + 'norms.2.bias'], 1e-5)
    return query


def decoder(sd, query_pos, memory, key_pos, cfg, self_mask=None, cross_mask=None,
            key_padding_mask=None):
    """MV2DTransformer + PETRTransformerDecoder with return_intermediate
    (cross_attention_head.py:22-49; petr_transformer.py:569-593): target=0, post_norm on
    every intermediate.  Returns [L, nq, bs, C]."""
    C = query_pos.shape[-1]
    query = torch.zeros_like(query_pos)
    inter = []
    for l in range(cfg['num_layers']):
        query = decoder_layer(sd, l, query, query_pos, memory, key_pos, cfg, self_mask,
                              cross_mask, key_padding_mask)
        inter.append(F.layer_norm(query, (C,), sd['bbox_head.transformer.decoder.post_norm.weight'],
                                  sd['bbox_head.transformer.decoder.post_norm.bias'], 1e-5))
    return torch.stack(inter)


# ----------------------------------------------------------------------------- a18: branches
def branches(sd, outs_dec, ref, cfg):
    """cls/reg branches + reference-point refinement (cross_attention_head.py:216-242).
    outs_dec [L, ..., C]; ref [..., 3] broadcastable.  Returns (cls [L,...,10], box [L,...,10])."""
    pc = cfg['pc_range']
    C = outs_dec.shape[-1]
    cls_all, box_all = [], []
    rinv = inverse_sigmoid(ref.clone())
    for l in range(outs_dec.shape[0]):
        x = outs_dec[l]
        p = f'bbox_head.cls_branches.{l}.'
        c = F.linear(x, sd[p + '0.weight'], sd[p + '0.bias'])
        c = F.relu(F.layer_norm(c, (C,), sd[p + '1.weight'], sd[p + '1.bias'], 1e-5))
        c = F.linear(c, sd[p + '3.weight'], sd[p + '3.bias'])
        c = F.relu(F.layer_norm(c, (C,), sd[p + '4.weight'], sd[p + '4.bias'], 1e-5))
        c = F.linear(c, sd[p + '6.weight'], sd[p + '6.bias'])
        p = f'bbox_head.reg_branches.{l}.'
        r = F.relu(F.linear(x, sd[p + '0.weight'], sd[p + '0.bias']))
        r = F.relu(F.linear(r, sd[p + '2.weight'], sd[p + '2.bias']))
        r = F.linear(r, sd[p + '4.weight'], sd[p + '4.bias']).clone()
        r[..., 0:2] = (r[..., 0:2] + rinv[..., 0:2]).sigmoid()
        r[..., 4:5] = (r[..., 4:5] + rinv[..., 2:3]).sigmoid()
        r[..., 0] = r[..., 0] * (pc[3] - pc[0]) + pc[0]
        r[..., 1] = r[..., 1] * (pc[4] - pc[1]) + pc[1]
        r[..., 4] = r[..., 4] * (pc[5] - pc[2]) + pc[2]
        cls_all.append(c)
        box_all.append(r)
    return torch.stack(cls_all), torch.stack(box_all)


# ----------------------------------------------------------------------------- full heads
def _prologue(sd, feat, pe, proposal_list, img_metas, cfg):
    proposal_list = guard_empty(proposal_list)
    rois = bbox2roi(proposal_list)
    K, E = get_box_params(proposal_list, img_metas, cfg['roi_size'])
    roi_feat = roi_align(feat, rois, cfg['roi_size'], 1.0 / cfg['stride'])
    ifeat = process_intrins_feat(rois, K, cfg['intrins_feat_scale'])
    ref, qg = query_generator(sd, roi_feat, K, E, ifeat, cfg)
    return proposal_list, rois, K, E, roi_feat, ifeat, ref, qg


def prepare_for_dn(ref, gt_boxes, gt_labels, rand, cfg, scalar=10, noise_scale=None, noise_trans=0.0, split=None,
                   num_classes=10, eps=1e-4):
    """MV2DSHead.prepare_for_dn, training branch, batch_size 1 (roi_heads/mv2d_s_head.py:39-120), with the
    uniform noise `rand` ([scalar*G,3] in [0,1), what torch.rand_like returns there) passed in.
    Returns (padded reference points [pad+N,3], self-attention mask [T,T] bool, known_labels, pad_size)."""
    pc = cfg['pc_range']
    # config values: two_frames exp config (1.25, 0.6); MV2DSHead constructor defaults otherwise (1.0, 0.75)
    noise_scale = cfg.get('denoise_noise_scale', 1.0) if noise_scale is None else noise_scale
    split = cfg.get('denoise_split', 0.75) if split is None else split
    G = gt_boxes.shape[0]
    centers = gt_boxes[:, :3].repeat(scalar, 1).clone()
    scale = gt_boxes[:, 3:6].repeat(scalar, 1)
    labels = gt_labels.repeat(scalar).clone()
    rand_prob = rand * 2 - 1.0
    centers = centers + rand_prob * (scale / 2 + noise_trans) * noise_scale
    for i in range(3):
        centers[:, i] = (centers[:, i] - pc[i]) / (pc[i + 3] - pc[i])
    centers = centers.clamp(min=eps, max=1.0 - eps)
    labels[torch.norm(rand_prob, 2, 1) > split] = num_classes
    pad = G * scalar
    padded = torch.cat([torch.zeros(pad, 3), ref], 0)
    idx = torch.cat([torch.arange(G) + G * i for i in range(scalar)])
    padded[idx] = centers
    T = pad + ref.shape[0]
    mask = torch.zeros(T, T, dtype=torch.bool)
    mask[pad:, :pad] = True                        # matching queries cannot see the denoising ones
    for i in range(scalar):                        # denoising groups cannot see each other
        mask[G * i:G * (i + 1), G * (i + 1):pad] = True
        mask[G * i:G * (i + 1), :G * i] = True
    return padded, mask, labels, pad


def mv2d_s_forward(sd, feat, proposal_list, img_metas, cfg=None, return_stages=False, dn=None):
    """MV2DHead.simple_test minus decode for MV2DSHead, eval mode
    (roi_heads/mv2d_head.py:249-261 -> mv2d_s_head.py:122-211).
    Returns (cls_scores [L,N,10], bbox_preds [L,N,10])."""
    cfg = cfg or make_cfg('S')
    pe = pe_forward(sd, feat, img_metas, cfg)
    proposal_list, rois, K, E, roi_feat, ifeat, ref, qg = _prologue(sd, feat, pe, proposal_list,
                                                                    img_metas, cfg)
    roi_pe = roi_align(pe, rois, cfg['roi_size'], 1.0 / cfg['stride'])
    num_per_view = [len(p) for p in proposal_list]
    corr, mask = box_roi_correlation(rois, num_per_view, img_metas, cfg)
    N, M = corr.shape
    C = feat.shape[1]
    # corr_feats [N,M,C,7,7] -> memory [M*49, N, C]  (cross_attention_head.py:26-31)
    mem = roi_feat[corr].permute(1, 3, 4, 0, 2).reshape(M * 49, N, C)
    pos = roi_pe[corr].permute(1, 3, 4, 0, 2).reshape(M * 49, N, C)
    kpm = (~mask)[:, :, None].expand(N, M, 49).reshape(N, M * 49)
    dn_out = None
    if dn is None:
        qpos = query_embed(sd, ref[:, None])  # [N,1,C] (bs=N, nq=1)
        outs = decoder(sd, qpos.permute(1, 0, 2), mem, pos, cfg, key_padding_mask=kpm)  # [L,1,N,C]
        outs = outs.transpose(1, 2)  # [L,N,1,C]
        cls, box = branches(sd, outs, ref[:, None], cfg)
        cls, box = cls.flatten(1, 2), box.flatten(1, 2)
    else:
        # training with denoising queries (mv2d_s_head.py:158-180): ONE batch entry, every RoI's tokens as
        # memory, a [T, N*49] cross mask (denoising rows see everything) and the group self-attention mask
        vis = torch.zeros(N, N + 1, dtype=torch.bool)
        vis.scatter_(1, torch.where(mask, corr, torch.full_like(corr, N)), True)
        cross = ~vis[:, :N, None].expand(N, N, 49).reshape(N, N * 49)
        ref_all, self_mask, dn_labels, pad = prepare_for_dn(ref, dn['gt_boxes'], dn['gt_labels'], dn['rand'], cfg)
        cross = torch.cat([torch.zeros(pad, N * 49, dtype=torch.bool), cross], 0)
        mem1 = roi_feat.permute(0, 2, 3, 1).reshape(N * 49, 1, C)
        pos1 = roi_pe.permute(0, 2, 3, 1).reshape(N * 49, 1, C)
        qall = query_embed(sd, ref_all[None])  # [1,T,C]
        outs = decoder(sd, qall.permute(1, 0, 2), mem1, pos1, cfg, self_mask=self_mask, cross_mask=cross)  # [L,T,1,C]
        cls, box = branches(sd, outs.transpose(1, 2), ref_all[None], cfg)   # [L,1,T,10]
        cls, box = cls.flatten(1, 2), box.flatten(1, 2)
        dn_out = dict(cls=cls[:, :pad], box=box[:, :pad], ref=ref_all[:pad], attn_mask=self_mask, labels=dn_labels)
        cls, box = cls[:, pad:], box[:, pad:]
        outs = outs[:, pad:]                    # [L,N,1,C]
        qpos = qall[0, pad:, None]
    if return_stages:
        return cls, box, dict(pe=pe, rois=rois, intrinsics=K, extrinsics=E, roi_feat=roi_feat,
                              roi_pe=roi_pe, intrins_feat=ifeat, ref=ref, corr=corr,
                              corr_mask=mask, query_pos=qpos[:, 0], outs_dec=outs[:, :, 0], dn=dn_out, **qg)
    return cls, box


def mv2d_t_forward(sd, feat, proposal_list, img_metas, cfg=None, return_stages=False, dn=None):
    """MV2DTHead eval forward (roi_heads/mv2d_t_head.py:26-142): dense feature-map keys
    compacted to the union of per-query masks, per-query bool cross mask, velocity / dt."""
    cfg = cfg or make_cfg('T')
    pe = pe_forward(sd, feat, img_metas, cfg)
    proposal_list, rois, K, E, roi_feat, ifeat, ref, qg = _prologue(sd, feat, pe, proposal_list,
                                                                    img_metas, cfg)
    V, C, h, w = feat.shape
    num_per_view = [len(p) for p in proposal_list]
    key_mask = box_correlation_mask(rois, num_per_view, img_metas, h, w, cfg)  # [N,V,h,w]
    pad_mask = feat_masks(img_metas, h, w)[0]  # [V,h,w]
    cross = ~key_mask
    if dn is not None:   # training: a query without any key gets key (0,0,0) un-masked (mv2d_t_head.py:80-82)
        invalid = cross.view(cross.shape[0], -1).all(1)
        cross[invalid, 0, 0, 0] = False
    roi_mask = (~cross).any(0)  # [V,h,w]
    mem = feat.permute(0, 2, 3, 1)[roi_mask]  # [Nk,C]
    pos = pe.permute(0, 2, 3, 1)[roi_mask]
    kpm = pad_mask[roi_mask][None]  # [1,Nk]
    cross = cross[:, roi_mask]  # [N,Nk]
    self_mask, pad, ref_all = None, 0, ref
    if dn is not None:   # denoising queries are prepended; they see every compacted key (mv2d_t_head.py:91-98)
        ref_all, self_mask, dn_labels, pad = prepare_for_dn(ref, dn['gt_boxes'], dn['gt_labels'], dn['rand'], cfg)
        cross = torch.cat([cross.all(dim=0)[None].repeat(pad, 1), cross], 0)
    qpos = query_embed(sd, ref_all[None])  # [1,T,C]
    outs = decoder(sd, qpos.permute(1, 0, 2), mem[:, None], pos[:, None], cfg, self_mask=self_mask,
                   cross_mask=cross, key_padding_mask=kpm)  # [L,T,1,C]
    outs = outs.transpose(1, 2)  # [L,1,T,C]
    cls, box = branches(sd, outs, ref_all[None], cfg)
    cls, box = cls.flatten(1, 2), box.flatten(1, 2)
    dn_out = None
    if dn is not None:
        dn_out = dict(cls=cls[:, :pad], box=box[:, :pad], ref=ref_all[:pad], attn_mask=self_mask, labels=dn_labels)
        cls, box, outs = cls[:, pad:], box[:, pad:], outs[:, :, pad:]
    nvf = cfg['num_views_per_frame']
    if len(img_metas) > nvf:
        ts = np.array([m['timestamp'] for m in img_metas])
        dt = ts[nvf:].mean() - ts[:nvf].mean()
        box = torch.cat([box[..., :8], box[..., 8:] / dt], dim=-1)
    if return_stages:
        return cls, box, dict(pe=pe, rois=rois, intrinsics=K, extrinsics=E, roi_feat=roi_feat,
                              intrins_feat=ifeat, ref=ref, key_mask=key_mask,
                              query_pos=qpos[0, pad:], outs_dec=outs[:, 0], dn=dn_out, **qg)
    return cls, box


# ----------------------------------------------------------------------------- f1: decode (next row)
def nms_free_decode(cls_scores, bbox_preds, cfg, max_num=300):
    """NMSFreeCoder.decode_single + get_bboxes z-shift
    (core/bbox/coders/nms_free_coder.py:49-102; cross_attention_head.py:372)."""
    post = torch.tensor(cfg['position_range'])
    k = min(max_num, cls_scores.numel())
    scores, idx = cls_scores.sigmoid().view(-1).topk(k)
    labels = idx % cls_scores.shape[-1]
    b = bbox_preds[idx // cls_scores.shape[-1]]
    rot = torch.atan2(b[:, 6:7], b[:, 7:8])
    boxes = torch.cat([b[:, 0:1], b[:, 1:2], b[:, 4:5], b[:, 2:3].exp(), b[:, 3:4].exp(),
                       b[:, 5:6].exp(), rot, b[:, 8:9], b[:, 9:10]], dim=-1)
    m = (boxes[:, :3] >= post[:3]).all(1) & (boxes[:, :3] <= post[3:]).all(1)
    boxes, scores, labels = boxes[m], scores[m], labels[m]
    boxes = boxes.clone()
    boxes[:, 2] = boxes[:, 2] - boxes[:, 5] * 0.5
    return boxes, scores, labels


# ============================================================================= training targets and losses
# SURVEY.md 8f rank 3 ("next" row f3): Hungarian targets + focal / L1 losses of one decoder layer, and the
# denoising loss.  Forward values only (no autograd here: the restatement is the checker of the CUDA kernels).
LOSS_CFG = dict(   # configs/mv2d/exp/mv2d_r50_frcnn_single_frame_roi_1408x512_ep72.py:87-95,132-137
    code_weights=[1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.5, 1.5, 2.0, 2.0],
    cls_loss_weight=2.0, focal_gamma=2.0, focal_alpha=0.25, bbox_loss_weight=0.25,
    cls_cost_weight=2.0, reg_cost_weight=0.25, num_classes=10, bg_cls_weight=0.0,
)


def normalize_bbox(b):
    """core/bbox/util.py:38-58: (cx, cy, cz, w, l, h, rot, vx, vy) -> (cx, cy, log w, log l, cz, log h, sin, cos, vx, vy)."""
    return torch.cat([b[..., 0:1], b[..., 1:2], b[..., 3:4].log(), b[..., 4:5].log(), b[..., 2:3], b[..., 5:6].log(),
                      b[..., 6:7].sin(), b[..., 6:7].cos(), b[..., 7:8], b[..., 8:9]], dim=-1)


def match_cost(cls_pred, bbox_pred, gt_boxes, gt_labels, lc=LOSS_CFG):
    """hungarian_assigner_3d.py:119-131: FocalLossCost (mmdet 2.25.1) + BBox3DL1Cost (match_cost.py:24-26) on the
    first 8 normalised box codes, then nan_to_num(100, 100, -100)."""
    a, g, eps = lc['focal_alpha'], lc['focal_gamma'], 1e-12
    p = cls_pred.sigmoid()
    neg = -(1 - p + eps).log() * (1 - a) * p.pow(g)
    pos = -(p + eps).log() * a * (1 - p).pow(g)
    cls_cost = (pos[:, gt_labels] - neg[:, gt_labels]) * lc['cls_cost_weight']
    reg_cost = torch.cdist(bbox_pred[:, :8], normalize_bbox(gt_boxes)[:, :8], p=1) * lc['reg_cost_weight']
    return torch.nan_to_num(cls_cost + reg_cost, nan=100.0, posinf=100.0, neginf=-100.0)


def hungarian_assign(cls_pred, bbox_pred, gt_boxes, gt_labels, lc=LOSS_CFG):
    """hungarian_assigner_3d.py:66-150 -> assigned gt index per query, -1 = background."""
    from scipy.optimize import linear_sum_assignment
    N, G = bbox_pred.shape[0], gt_boxes.shape[0]
    out = torch.full((N,), -1, dtype=torch.long)
    if N == 0 or G == 0:
        return out
    # the assigner runs under no_grad on the host (hungarian_assigner_3d.py:135-137: cost.detach().cpu())
    rows, cols = linear_sum_assignment(match_cost(cls_pred, bbox_pred, gt_boxes, gt_labels, lc).detach())
    out[torch.from_numpy(rows)] = torch.from_numpy(cols)
    return out


def _focal_loss_sum(pred, labels, lc):
    """mmdet FocalLoss (sigmoid) element sum; labels == num_classes is background."""
    C = pred.shape[1]
    t = F.one_hot(labels, num_classes=C + 1)[:, :C].type_as(pred)
    p = pred.sigmoid()
    pt = (1 - p) * t + p * (1 - t)
    fw = (lc['focal_alpha'] * t + (1 - lc['focal_alpha']) * (1 - t)) * pt.pow(lc['focal_gamma'])
    return (F.binary_cross_entropy_with_logits(pred, t, reduction='none') * fw).sum()


def loss_single(cls_scores, bbox_preds, gt_boxes, gt_labels, lc=LOSS_CFG):
    """cross_attention_head.py:379-434 for one sample and one decoder layer (as mv2d_s_head.py:281-286 calls it).
    gt_boxes [G,9] = (gravity centre, w, l, h, yaw, vx, vy).  Returns (loss_cls, loss_bbox, assigned [N])."""
    N = cls_scores.shape[0]
    eps = torch.finfo(torch.float32).eps
    assigned = hungarian_assign(cls_scores, bbox_preds, gt_boxes, gt_labels, lc)
    pos = assigned >= 0
    num_pos, num_neg = int(pos.sum()), int((~pos).sum())
    labels = torch.full((N,), lc['num_classes'], dtype=torch.long)
    labels[pos] = gt_labels[assigned[pos]]
    cls_avg = max(num_pos * 1.0 + num_neg * lc['bg_cls_weight'], 1)
    loss_cls = lc['cls_loss_weight'] * _focal_loss_sum(cls_scores, labels, lc) / (cls_avg + eps)
    targets = torch.zeros(N, gt_boxes.shape[1] if gt_boxes.numel() else 9)
    targets[pos] = gt_boxes[assigned[pos]]
    nt = normalize_bbox(targets)
    ok = torch.isfinite(nt).all(dim=-1)
    w = pos.float()[:, None] * torch.tensor(lc['code_weights'])
    if int(ok.sum()) == 0:
        loss_bbox = torch.zeros(())
    else:
        loss_bbox = lc['bbox_loss_weight'] * ((bbox_preds[ok, :10] - nt[ok, :10]).abs() * w[ok]).sum() / (max(num_pos, 1) + eps)
    return torch.nan_to_num(loss_cls), torch.nan_to_num(loss_bbox), assigned


def dn_loss_single(cls_scores, bbox_preds, known_boxes, known_labels, num_tgt, split, lc=LOSS_CFG, neg_bbox_loss=False):
    """cross_attention_head.py:475-538 (neg_bbox_loss False): cls_scores/bbox_preds [pad,10] of the denoising
    queries, known_boxes [pad,9] = the GT box each one was noised from, known_labels [pad] (num_classes = negative)."""
    eps = torch.finfo(torch.float32).eps
    cls_avg = max(num_tgt * 3.14159 / 6 * split * split * split, 1)
    loss_cls = lc['cls_loss_weight'] * _focal_loss_sum(cls_scores, known_labels.long(), lc) / (cls_avg + eps)
    kb = known_boxes.clone()
    if not neg_bbox_loss:       # cross_attention_head.py:521-523
        kb[known_labels == lc['num_classes']] = 0
    nt = normalize_bbox(kb)
    ok = torch.isfinite(nt).all(dim=-1)
    w = torch.tensor(lc['code_weights']).repeat(kb.shape[0], 1)
    w[:, 6:8] = 0
    if int(ok.sum()) == 0:
        loss_bbox = torch.zeros(())
    else:
        loss_bbox = lc['bbox_loss_weight'] * ((bbox_preds[ok, :10] - nt[ok, :10]).abs() * w[ok]).sum() / (max(num_tgt, 1) + eps)
    return torch.nan_to_num(loss_cls), torch.nan_to_num(loss_bbox)


def decoder_slice_loss(sd, ref, roi_feat, roi_pe, corr, mask, gt_boxes, gt_labels, cfg=None, stage_loss_weights=None, lc=LOSS_CFG):
    """Row e (training step) of the MV2D-S head, DIFFERENTIABLE: the slice mv2d_decoder_train_forward covers --
    bbox_head forward on the gathered RoI tokens as MV2DSHead._bbox_forward_denoise calls it without denoising
    (mv2d_s_head.py:184-196 -> cross_attention_head.py:202-242) and the per-layer losses summed with
    stage_loss_weights as forward_train does (mv2d_s_head.py:278-305).  ref [N,3], roi_feat / roi_pe [N,256,7,7],
    corr [N,M] int64, mask [N,M] bool.  Returns (total, cls [L,N,10], box [L,N,10], [(loss_cls, loss_bbox, assigned)])."""
    cfg = cfg or make_cfg('S')
    N, M = corr.shape
    C = roi_feat.shape[1]
    mem = roi_feat[corr].permute(1, 3, 4, 0, 2).reshape(M * 49, N, C)
    pos = roi_pe[corr].permute(1, 3, 4, 0, 2).reshape(M * 49, N, C)
    kpm = (~mask)[:, :, None].expand(N, M, 49).reshape(N, M * 49)
    qpos = query_embed(sd, ref[:, None])
    outs = decoder(sd, qpos.permute(1, 0, 2), mem, pos, cfg, key_padding_mask=kpm).transpose(1, 2)
    cls, box = branches(sd, outs, ref[:, None], cfg)
    cls, box = cls.flatten(1, 2), box.flatten(1, 2)
    w = stage_loss_weights or [0.1] * cls.shape[0]
    total, per = 0.0, []
    for l in range(cls.shape[0]):
        a, b, asg = loss_single(cls[l], box[l], gt_boxes, gt_labels, lc)
        total = total + w[l] * (a + b)
        per.append((a, b, asg))
    return total, cls, box, per


def hot_path_loss(sd, feat, proposal_list, img_metas, gt_boxes, gt_labels, cfg=None, stage_loss_weights=None, lc=LOSS_CFG):
    """The whole MV2D-S hot path as a DIFFERENTIABLE function of the weights and of the feature map: what
    MV2DSHead.forward_train computes (roi_heads/mv2d_s_head.py:236-307: position encoding, RoIAlign of cat(feat, pe),
    query generator, bbox_head on the gathered RoI tokens, per-layer losses times stage_loss_weights).  The box
    correlation and the per-RoI camera parameters are computed without gradient, as in the reference
    (@torch.no_grad, box_correlation.py:164, mv2d_head.py:51).  Returns (total, cls, box, per-layer tuples)."""
    cfg = cfg or make_cfg('S')
    pe = pe_forward(sd, feat, img_metas, cfg)
    with torch.no_grad():
        proposal_list = guard_empty(proposal_list)
        rois = bbox2roi(proposal_list)
        K, E = get_box_params(proposal_list, img_metas, cfg['roi_size'])
        ifeat = process_intrins_feat(rois, K, cfg['intrins_feat_scale'])
        corr, mask = box_roi_correlation(rois, [len(p) for p in proposal_list], img_metas, cfg)
    roi_feat = roi_align(feat, rois, cfg['roi_size'], 1.0 / cfg['stride'])
    roi_pe = roi_align(pe, rois, cfg['roi_size'], 1.0 / cfg['stride'])
    ref, _ = query_generator(sd, roi_feat, K, E, ifeat, cfg)
    return decoder_slice_loss(sd, ref, roi_feat, roi_pe, corr, mask, gt_boxes, gt_labels, cfg, stage_loss_weights, lc)


def hot_path_loss_t(sd, feat, proposal_list, img_metas, gt_boxes, gt_labels, rand, cfg=None, stage_loss_weights=None,
                    denoise_weight=1.0, neg_bbox_loss=True, lc=LOSS_CFG):
    """The two-frame head's training step WITH denoising queries as a DIFFERENTIABLE function (the configuration the
    reference trains MV2D-T with, exp/mv2d_r50_frcnn_two_frames_1408x512_ep*.py:44-47): MV2DSHead.forward_train as
    MV2DTHead inherits it (roi_heads/mv2d_s_head.py:236-307) -- per layer loss_cls / loss_bbox of the matching queries
    plus dn_loss_cls / dn_loss_bbox of the denoising queries times denoise_weight, each times stage_loss_weights.
    Returns (total, dict of the named weighted losses)."""
    cfg = cfg or make_cfg('T')
    cls, box, st = mv2d_t_forward(sd, feat, proposal_list, img_metas, cfg, return_stages=True,
                                  dn=dict(gt_boxes=gt_boxes, gt_labels=gt_labels, rand=rand))
    dn = st['dn']
    pad = dn['cls'].shape[1]
    known_boxes = gt_boxes.repeat(pad // max(gt_boxes.shape[0], 1), 1)
    w = stage_loss_weights or [0.1] * cls.shape[0]
    losses = {}
    for l in range(cls.shape[0]):
        a, b, _ = loss_single(cls[l], box[l], gt_boxes, gt_labels, lc)
        c, d = dn_loss_single(dn['cls'][l], dn['box'][l], known_boxes, dn['labels'], pad, cfg['denoise_split'], lc,
                              neg_bbox_loss=neg_bbox_loss)
        losses[f'l{l}.loss_cls'], losses[f'l{l}.loss_bbox'] = a * w[l], b * w[l]
        losses[f'l{l}.dn_loss_cls'], losses[f'l{l}.dn_loss_bbox'] = c * denoise_weight * w[l], d * denoise_weight * w[l]
    return sum(losses.values()), losses


# ============================================================================= neck
def fpn_neck(neck_sd, x):
    """The MV2D neck: mmdet 2.25.1 FPN with in_channels [256]*5, start_level = end_level = 2, num_outs = 1
    (configs/mv2d/exp/mv2d_r50_frcnn_single_frame_roi_1408x512_ep72.py:32-39; called in detectors/mv2d.py:122-127).
    One level => no top-down path: out = fpn_convs[0](lateral_convs[0](x)), ConvModules without norm / activation.
    mmdet is third-party and absent from /root/reference: restated from its published FPN.forward (unpinned)."""
    sd = {(k[len('neck.'):] if k.startswith('neck.') else k): v for k, v in neck_sd.items()}
    lat = F.conv2d(x, sd['lateral_convs.0.conv.weight'], sd['lateral_convs.0.conv.bias'])
    return F.conv2d(lat, sd['fpn_convs.0.conv.weight'], sd['fpn_convs.0.conv.bias'], padding=1)


def scene_nms(boxes, scores, labels, score_thr=0.0, max_num=300, num_classes=10):
    """detectors/mv2d.py:266-282 -> mmdet3d 1.0 box3d_multiclass_nms with the configs' nms_thr = 1.0 (exp/...:150-154):
    the rotated BEV NMS suppresses nothing, so per class (ascending) the boxes with score > score_thr in descending
    score order are concatenated; beyond max_num the best max_num by score are kept in descending score order.
    (mmdet3d is third-party and absent: restated from its published behaviour, unpinned.)"""
    ob, os_, ol = [], [], []
    for c in range(num_classes):
        m = (labels == c) & (scores > score_thr)
        if m.any():
            idx = torch.nonzero(m).squeeze(1)
            order = torch.argsort(scores[idx], descending=True, stable=True)
            ob.append(boxes[idx][order]); os_.append(scores[idx][order]); ol.append(labels[idx][order])
    if not ob:
        return boxes.new_zeros((0, boxes.shape[1])), scores.new_zeros((0,)), labels.new_zeros((0,))
    b, s, l = torch.cat(ob), torch.cat(os_), torch.cat(ol)
    if b.shape[0] > max_num:
        inds = torch.argsort(s, descending=True, stable=True)[:max_num]
        b, s, l = b[inds], s[inds], l[inds]
    return b, s, l


# ----------------------------------------------------------------------------- f2: detections hand-off (next row)
def process_2d_detections(results, min_bbox_size=0):
    """MV2D.process_2d_detections (detectors/mv2d.py:60-86): per view the 2D detector's per-class [n_c,5] arrays ->
    one [n,6] tensor (x1, y1, x2, y2, score, label); boxes with a side below ``min_bbox_size`` are dropped."""
    out = []
    for res in results:
        det = torch.cat([torch.cat([torch.as_tensor(b, dtype=torch.float32).reshape(-1, 5),
                                    torch.full((len(b), 1), float(i))], dim=1) for i, b in enumerate(res)], dim=0)
        if min_bbox_size > 0:
            wh = det[:, 2:4] - det[:, 0:2]
            det = det[(wh >= min_bbox_size).all(dim=1)]
        out.append(det)
    return out


def box_iou_2d(a, b, eps=1e-4):
    """MV2D.box_iou (detectors/mv2d.py:88-102): a [n,4], b [m,4] -> [n,m]; same fp32 operation order."""
    a, b = a[:, None, :], b[None, :, :]
    wh = torch.maximum(torch.minimum(a[..., 2:4], b[..., 2:4]) - torch.maximum(a[..., 0:2], b[..., 0:2]), a.new_tensor(0))
    inter = wh.prod(-1)
    union = (a[..., 2:4] - a[..., 0:2]).prod(-1) + (b[..., 2:4] - b[..., 0:2]).prod(-1) - inter
    return inter / (union + eps)


def complement_2d_gt(detections, gts, thr=0.35, min_bbox_size=0):
    """MV2D.complement_2d_gt (detectors/mv2d.py:104-117): append the 2D ground-truth boxes ([m,6], score 1) whose best
    IoU with the detections is below ``thr`` and whose sides reach ``min_bbox_size``.  Quirks kept: no ground truth ->
    the detections; no detections -> ALL ground-truth boxes, unfiltered."""
    if len(gts) == 0:
        return detections
    if len(detections) == 0:
        return gts
    max_iou = box_iou_2d(gts[:, :4], detections[:, :4]).max(-1)[0]
    wh = gts[:, 2:4] - gts[:, 0:2]
    keep = (max_iou < thr) & (wh >= min_bbox_size).all(dim=1)
    return torch.cat([detections, gts[keep]], dim=0)
