"""Seeded synthetic inputs and weights for the MV2D decoder hot path.

Everything is drawn from ``numpy.random.Generator(PCG64(seed))`` so the oracle, the CUDA path,
the golden fixtures and the bench see identical bits on any box with this image.

Shapes and names of the weights follow the reference ``state_dict`` (prefix ``roi_head.``
stripped) -- SURVEY.md App. B; checked against the reference constructors in
``tests/test_reference_parity_cpu.py`` when ``/root/reference`` is present.
"""
import math

import numpy as np
import torch

PC_RANGE = [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]
POST_RANGE = [-61.2, -61.2, -10.0, 61.2, 61.2, 10.0]
IMG_H, IMG_W = 512, 1408
STRIDE = 16
FEAT_H, FEAT_W = IMG_H // STRIDE, IMG_W // STRIDE
EMBED = 256
# loader order of nuScenes cameras (reference datasets/pipelines/loading.py:70)
CAM_YAW_DEG = [0.0, -55.0, 55.0, 180.0, 110.0, -110.0]


def _rz(deg):
    a = math.radians(deg)
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]], dtype=np.float64)


def make_cameras(num_views=6, ego_shift=(0.0, -4.0, 0.0), jitter=None):
    """nuScenes-like 6-camera rig at test-time image augmentation (resize .88, crop
    (0,280,1408,792)).  Views 6..11 (two-frame model) are the same rig moved by ``ego_shift``.
    Returns per-view dicts with ``intrinsics``, ``extrinsics`` (= lidar2cam transposed, as
    custom_nuscenes_dataset.py:141-150 stores it) and ``lidar2img`` = K @ extrinsics.T."""
    K = np.eye(4, dtype=np.float64)
    K[0, 0] = K[1, 1] = 1114.45
    K[0, 2], K[1, 2] = 718.3, 152.5
    cam_axes = np.array([[1.0, 0, 0], [0, 0, 1.0], [0, -1.0, 0]], dtype=np.float64)
    out = []
    for v in range(num_views):
        yaw = CAM_YAW_DEG[v % 6]
        if jitter is not None:
            yaw = yaw + float(jitter[v])
        R = _rz(yaw) @ cam_axes
        t = _rz(yaw) @ np.array([0.0, 0.5, 0.0]) + np.array([0.0, 0.0, -0.3])
        if v >= 6:
            t = t + np.asarray(ego_shift, dtype=np.float64)
        cam2lidar = np.eye(4, dtype=np.float64)
        cam2lidar[:3, :3] = R
        cam2lidar[:3, 3] = t
        lidar2cam = np.linalg.inv(cam2lidar)
        extr = np.ascontiguousarray(lidar2cam.T)
        out.append(dict(intrinsics=K.copy(), extrinsics=extr, lidar2img=K @ extr.T))
    return out


def make_img_metas(num_views=6, img_shape=(IMG_H, IMG_W, 3), pad_shape=(IMG_H, IMG_W, 3),
                   jitter=None):
    cams = make_cameras(num_views, jitter=jitter)
    metas = []
    for v, cam in enumerate(cams):
        m = dict(num_views=num_views, pad_shape=tuple(pad_shape), img_shape=tuple(img_shape),
                 timestamp=0.0 if v < 6 else 0.5)
        m.update(cam)
        metas.append(m)
    return metas


def make_boxes(rng, num_views, boxes_per_view, img_w=IMG_W, img_h=IMG_H):
    """Per view ``[n_v, 6]`` = (x1, y1, x2, y2, score, label) float32, min side 8 px."""
    if isinstance(boxes_per_view, int):
        boxes_per_view = [boxes_per_view] * num_views
    out = []
    for v in range(num_views):
        rows = []
        while len(rows) < boxes_per_view[v]:
            cx, cy = rng.uniform(0, img_w), rng.uniform(100, 412)
            w, h = rng.uniform(16, 320), rng.uniform(16, 240)
            x1, x2 = max(cx - w / 2, 0.0), min(cx + w / 2, img_w - 1.0)
            y1, y2 = max(cy - h / 2, 0.0), min(cy + h / 2, img_h - 1.0)
            if x2 - x1 < 8 or y2 - y1 < 8:
                continue
            rows.append([x1, y1, x2, y2, rng.uniform(0.05, 1.0), float(rng.integers(10))])
        out.append(torch.tensor(np.asarray(rows, dtype=np.float32).reshape(-1, 6)))
    return out


def make_feat(rng, num_views, channels=EMBED, h=FEAT_H, w=FEAT_W):
    return torch.from_numpy(rng.standard_normal((num_views, channels, h, w), dtype=np.float32))


def _xavier(rng, shape):
    fan_out, fan_in = shape[0], int(np.prod(shape[1:]))
    recept = 1
    if len(shape) > 2:
        recept = int(np.prod(shape[2:]))
        fan_in, fan_out = shape[1] * recept, shape[0] * recept
    a = math.sqrt(6.0 / (fan_in + fan_out))
    return torch.from_numpy(rng.uniform(-a, a, size=shape).astype(np.float32))


def _bias(rng, n, lo=-0.1, hi=0.1):
    return torch.from_numpy(rng.uniform(lo, hi, size=(n,)).astype(np.float32))


def make_neck_state_dict(seed=0, embed=EMBED):
    """``neck.*`` of the reference model: a one-level mmdet FPN (configs/mv2d/exp/*.py:32-39) = a 1x1 lateral conv and
    a 3x3 output conv, both 256 -> 256 with bias."""
    rng = np.random.Generator(np.random.PCG64(7000 + seed))
    return {'lateral_convs.0.conv.weight': _xavier(rng, (embed, embed, 1, 1)), 'lateral_convs.0.conv.bias': _bias(rng, embed),
            'fpn_convs.0.conv.weight': _xavier(rng, (embed, embed, 3, 3)), 'fpn_convs.0.conv.bias': _bias(rng, embed)}


def make_state_dict(seed=0, num_layers=6, embed=EMBED, ffn=2048, num_classes=10, code_size=10,
                    depth_num=64):
    """Hot-path ``state_dict`` with the reference's key names and shapes."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = {}

    def lin(name, out_c, in_c):
        sd[name + '.weight'] = _xavier(rng, (out_c, in_c))
        sd[name + '.bias'] = _bias(rng, out_c)

    def conv(name, out_c, in_c, k):
        sd[name + '.weight'] = _xavier(rng, (out_c, in_c, k, k))
        sd[name + '.bias'] = _bias(rng, out_c)

    def ln(name, n):
        sd[name + '.weight'] = torch.from_numpy(rng.uniform(0.8, 1.2, size=(n,)).astype(np.float32))
        sd[name + '.bias'] = _bias(rng, n)

    # PE (pe.py:64-82)
    conv('position_encoding.position_encoder.0', embed * 4, 3 * depth_num, 1)
    conv('position_encoding.position_encoder.2', embed, embed * 4, 1)
    conv('position_encoding.adapt_pos3d.0', embed * 4, embed * 3 // 2, 1)
    conv('position_encoding.adapt_pos3d.2', embed, embed * 4, 1)
    conv('position_encoding.fpe.conv_reduce', embed, embed, 1)
    conv('position_encoding.fpe.conv_expand', embed, embed, 1)
    # QueryGenerator (query_generator.py:175-234)
    conv('query_generator.shared_convs.0.conv', embed, embed, 3)
    lin('query_generator.shared_fcs.0', 1024, embed)
    lin('query_generator.extra_enc.0', 512, 1024 + 16)
    lin('query_generator.extra_enc.2', embed, 512)
    sd['query_generator.fc_center.weight'] = torch.from_numpy(
        (rng.standard_normal((3, embed)) * 0.05).astype(np.float32))
    sd['query_generator.fc_center.bias'] = torch.tensor([3.5, 3.5, 20.0])
    # bbox head (cross_attention_head.py:118-146,184-185)
    lin('bbox_head.query_embedding.0', embed, embed * 3 // 2)
    lin('bbox_head.query_embedding.2', embed, embed)
    for l in range(num_layers):
        p = f'bbox_head.transformer.decoder.layers.{l}.'
        for a in (0, 1):
            sd[p + f'attentions.{a}.attn.in_proj_weight'] = _xavier(rng, (3 * embed, embed))
            sd[p + f'attentions.{a}.attn.in_proj_bias'] = _bias(rng, 3 * embed)
            lin(p + f'attentions.{a}.attn.out_proj', embed, embed)
        lin(p + 'ffns.0.layers.0.0', ffn, embed)
        lin(p + 'ffns.0.layers.1', embed, ffn)
        for n in (0, 1, 2):
            ln(p + f'norms.{n}', embed)
    ln('bbox_head.transformer.decoder.post_norm', embed)
    for l in range(num_layers):
        p = f'bbox_head.cls_branches.{l}.'
        lin(p + '0', embed, embed)
        ln(p + '1', embed)
        lin(p + '3', embed, embed)
        ln(p + '4', embed)
        lin(p + '6', num_classes, embed)
        p = f'bbox_head.reg_branches.{l}.'
        lin(p + '0', embed, embed)
        lin(p + '2', embed, embed)
        lin(p + '4', code_size, embed)
    sd['bbox_head.code_weights'] = torch.tensor([1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.5, 1.5, 2.0, 2.0])
    return sd


def make_sample(seed=0, num_views=6, boxes_per_view=50, img_shape=(IMG_H, IMG_W, 3),
                pad_shape=(IMG_H, IMG_W, 3), cam_jitter_deg=0.0):
    """One synthetic sample: (feat [V,256,h,w] NCHW f32, proposal_list, img_metas)."""
    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    jitter = rng.uniform(-cam_jitter_deg, cam_jitter_deg, size=num_views) if cam_jitter_deg else None
    metas = make_img_metas(num_views, img_shape, pad_shape, jitter=jitter)
    h, w = pad_shape[0] // STRIDE, pad_shape[1] // STRIDE
    feat = make_feat(rng, num_views, EMBED, h, w)
    boxes = make_boxes(rng, num_views, boxes_per_view, img_w=img_shape[1], img_h=img_shape[0])
    return feat, boxes, metas


# ----------------------------------------------------------------------------- named cases
# Parity-test / golden cases.  ``extra_boxes`` are appended verbatim to the given view
# (edge cases: boxes below the 4-px intrinsics-feature threshold, image-border boxes).
CASES = {
    # config 1 of BASELINE.json: 50 queries, 1 decoder layer
    's_cfg1': dict(mode='S', seed=11, num_views=6, boxes_per_view=[9, 8, 8, 9, 8, 8], num_layers=1),
    # small S case with an empty view and a tiny (<4 px) box
    's_small': dict(mode='S', seed=1, num_views=6, boxes_per_view=[5, 0, 7, 3, 6, 4], num_layers=6,
                    extra_boxes={3: [[700.0, 200.0, 703.0, 260.0, 0.9, 2.0],
                                     [0.0, 0.0, 1407.0, 511.0, 0.8, 1.0]]}),
    # zero detections -> dummy box guard (mv2d_s_head.py:124-127)
    's_empty': dict(mode='S', seed=2, num_views=6, boxes_per_view=0, num_layers=2),
    # padded image: img_shape < pad_shape makes the padding masks non-trivial
    's_pad': dict(mode='S', seed=3, num_views=6, boxes_per_view=4, num_layers=2,
                  img_shape=(480, 1376, 3), pad_shape=(512, 1408, 3)),
    # one query only
    's_one': dict(mode='S', seed=4, num_views=6, boxes_per_view=[0, 0, 1, 0, 0, 0], num_layers=2),
    # config 2 of BASELINE.json (the bench workload): 300 queries, 6 layers
    's_cfg2': dict(mode='S', seed=0, num_views=6, boxes_per_view=50, num_layers=6),
    't_small': dict(mode='T', seed=5, num_views=12, boxes_per_view=[4, 3, 0, 5, 4, 3, 4, 0, 3, 5, 4, 4],
                    num_layers=6),
    't_pad': dict(mode='T', seed=6, num_views=12, boxes_per_view=3, num_layers=2,
                  img_shape=(480, 1376, 3), pad_shape=(512, 1408, 3)),
    # config 3 of BASELINE.json (one of its two samples): 12 views, 300 queries
    't_cfg3': dict(mode='T', seed=7, num_views=12, boxes_per_view=25, num_layers=6),
    's_dn': dict(mode='S', seed=9, num_views=6, boxes_per_view=[4, 3, 5, 2, 4, 3], num_layers=2,
                 dn=dict(num_gt=5, seed=81)),
    # row a20: training-mode forward with denoising queries (7 GT boxes x 10 noised copies prepended)
    't_dn': dict(mode='T', seed=8, num_views=12, boxes_per_view=[3, 2, 4, 3, 2, 3, 3, 2, 4, 3, 2, 3], num_layers=2,
                 dn=dict(num_gt=7, seed=80)),
}


def make_dn_inputs(dn_spec, scalar=10):
    """Synthetic GT for the denoising branch: boxes [G,9] = (gravity centre xyz, w, l, h, yaw, vx, vy),
    labels [G], and the uniform noise prepare_for_dn draws with torch.rand_like ([scalar*G, 3] in [0,1))."""
    rng = np.random.Generator(np.random.PCG64(dn_spec['seed']))
    G = dn_spec['num_gt']
    boxes = np.concatenate([rng.uniform(-40, 40, (G, 2)), rng.uniform(-3, 1, (G, 1)), rng.uniform(0.5, 5, (G, 3)),
                            rng.uniform(-3.1, 3.1, (G, 1)), rng.uniform(-2, 2, (G, 2))], 1).astype(np.float32)
    labels = rng.integers(0, 10, G).astype(np.int64)
    rand = rng.uniform(0, 1, (scalar * G, 3)).astype(np.float32)
    return torch.from_numpy(boxes), torch.from_numpy(labels), torch.from_numpy(rand)


def case_inputs(spec):
    """(feat, proposal_list, img_metas) for a CASES entry."""
    img_shape = tuple(spec.get('img_shape', (IMG_H, IMG_W, 3)))
    pad_shape = tuple(spec.get('pad_shape', (IMG_H, IMG_W, 3)))
    feat, boxes, metas = make_sample(spec['seed'], spec['num_views'], spec['boxes_per_view'],
                                     img_shape=img_shape, pad_shape=pad_shape,
                                     cam_jitter_deg=spec.get('cam_jitter_deg', 0.0))
    for v, rows in spec.get('extra_boxes', {}).items():
        extra = torch.tensor(rows, dtype=torch.float32).reshape(-1, 6)
        boxes[int(v)] = torch.cat([boxes[int(v)], extra], 0)
    return feat, boxes, metas
