"""Row f4 on the GPU: mv2d_fpn_neck (1x1 lateral + 3x3 output conv of the one-level FPN, both 3xTF32 tcgen05 GEMMs,
the 3x3 with 4-D TMA implicit im2col over the feature map) against torch conv2d in fp64 and, end to end, against the
oracle (neck -> hot path)."""
import pytest
import torch
import torch.nn.functional as F

from mv2d_b200 import synth

pytestmark = pytest.mark.gpu


def _ref64(nsd, x):
    lat = F.conv2d(x.double(), nsd['lateral_convs.0.conv.weight'].double(), nsd['lateral_convs.0.conv.bias'].double())
    return F.conv2d(lat, nsd['fpn_convs.0.conv.weight'].double(), nsd['fpn_convs.0.conv.bias'].double(), padding=1)


@pytest.mark.parametrize('V,h,w', [(6, 32, 88), (12, 32, 88), (2, 30, 85), (1, 5, 3)])
def test_fpn_neck_matches_conv2d_fp64(V, h, w, state_dicts):
    from mv2d_b200.engine import HotPath
    eng = HotPath(state_dicts(6), mode='S')
    nsd = synth.make_neck_state_dict(3)
    x = torch.randn(V, 256, h, w, generator=torch.Generator().manual_seed(V * h + w))
    feat, feat_tf32 = eng.neck(x.cuda(), nsd)
    torch.cuda.synchronize()
    ref = _ref64(nsd, x).permute(0, 2, 3, 1)
    d = (feat.cpu().double() - ref).abs()
    assert torch.isfinite(feat).all()
    # 3xTF32 products are fp32-grade (2^-21); the tensor core's fp32 accumulation over K = 256 + 2304 terms leaves
    # ~1e-4 absolute at O(1) outputs (same figure as the RoI conv, test_tcgen05_gemm_3xtf32_and_im2col)
    assert (d <= 3e-4 + 1e-4 * ref.abs()).all(), f'max |d| = {d.max().item():.3e}'
    # the TF32 copy: feat rounded to 10 mantissa bits
    assert ((feat_tf32 - feat).abs() <= feat.abs() * 2.0 ** -11 + 1e-30).all()
    # channels-last input takes the same path minus the transpose
    feat2, _ = eng.neck(x.permute(0, 2, 3, 1).contiguous().cuda(), nsd, in_is_nhwc=True)
    assert torch.equal(feat2, feat)


def test_neck_feeds_the_hot_path(state_dicts):
    """detector P4 -> neck -> decoder hot path, against the oracle's neck + forward."""
    from mv2d_b200.engine import HotPath
    from oracle import mv2d_oracle as O
    sd = state_dicts(6)
    eng = HotPath(sd, mode='S')
    nsd = synth.make_neck_state_dict(5)
    p4, boxes, metas = synth.make_sample(55, 6, 6)
    with torch.no_grad():
        feat_ref = O.fpn_neck(nsd, p4)
        cls, box = O.mv2d_s_forward(sd, feat_ref, boxes, metas, O.make_cfg('S'))
    feat, _ = eng.neck(p4.cuda(), nsd)
    out = eng.forward(feat, boxes, metas, feat_is_nhwc=True)
    torch.cuda.synchronize()
    bad_c = (out['cls_scores'].cpu() - cls).abs() > 1e-3 + 1e-3 * cls.abs()
    bad_b = (out['bbox_preds'].cpu() - box).abs() > 1e-3 + 1e-3 * box.abs()
    assert not bad_c.any() and not bad_b.any()
