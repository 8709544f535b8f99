"""CPU: host-side weight re-layout (the library's mv2d_pack_weights, called through mv2d_b200/pack.py).  The absorbed
cross-attention matrices must reproduce torch.nn.MultiheadAttention exactly in real arithmetic; checked here in fp32
against the oracle's MHA call on random inputs with a per-query key mask, and buffer by buffer against a torch fp64
restatement of the packing."""
import math

import torch
import torch.nn.functional as F

from mv2d_b200.pack import PackedNeck, PackedWeights, round_tf32
from oracle import mv2d_oracle as O


def absorb_cross_attention(in_w, in_b, out_w, out_b):
    """torch fp64 restatement of the absorbed matrices (checker for csrc/pack.cpp)."""
    in_w, in_b, out_w, out_b = [t.detach().double() for t in (in_w, in_b, out_w, out_b)]
    wq, wk, wv = in_w[:256], in_w[256:512], in_w[512:]
    bq, bv = in_b[:256], in_b[512:]
    scale = 1.0 / math.sqrt(32)
    qw, qb, ow = [], [], []
    for h in range(8):
        s = slice(h * 32, (h + 1) * 32)
        qw.append(scale * wk[s].T @ wq[s])
        qb.append(scale * wk[s].T @ bq[s])
        ow.append(out_w[:, s] @ wv[s])
    return torch.cat(qw, 0).float(), torch.cat(qb, 0).float(), torch.cat(ow, 1).float(), (out_w @ bv + out_b).float()


def first_layer_self_attn_const(in_proj_bias, out_w, out_b):
    return (out_w.double() @ in_proj_bias.double()[512:] + out_b.double()).float()


def _close(a, b, ulps=2):
    """equal up to `ulps` fp32 ulps of the larger magnitude (fp64 summation order differs between the two)"""
    tol = ulps * torch.finfo(torch.float32).eps * torch.maximum(a.abs(), b.abs()).clamp_min(1e-30)
    return bool(((a - b).abs() <= tol).all())


def test_pack_matches_torch_restatement(state_dicts):
    sd = state_dicts(2)
    w = PackedWeights(sd, 'cpu')
    assert w.num_layers == 2 and w.nbytes() == sum((v.numel() * 4 + 255) // 256 * 256 for k, v in w.t.items()
                                                   if not k.split('.')[-1] in ('xa_k_w', 'xa_k_w_lo', 'xa_v_w', 'xa_v_w_lo'))
    # the plain K / V projections of all layers are stacked into one operand (rows (2 l + side) * 256 ..)
    for l in range(2):
        for side, nm in enumerate(('xa_k_w', 'xa_v_w')):
            assert w.t[f'l{l}.{nm}'].data_ptr() == w.t['xa_kv_w'].data_ptr() + (2 * l + side) * 256 * 256 * 4
            assert w.t[f'l{l}.{nm}_lo'].data_ptr() == w.t['xa_kv_w_lo'].data_ptr() + (2 * l + side) * 256 * 256 * 4
            assert getattr(w.layers[l], nm) == w.t[f'l{l}.{nm}'].data_ptr()
    for l in range(2):
        p = f'bbox_head.transformer.decoder.layers.{l}.attentions.1.attn.'
        qw, qb, ow, ob = absorb_cross_attention(sd[p + 'in_proj_weight'], sd[p + 'in_proj_bias'], sd[p + 'out_proj.weight'],
                                                sd[p + 'out_proj.bias'])
        # the split is exact where it is formed: hi + lo == w to the last TF32 bit of lo
        for name, ref in ((f'l{l}.ca_q_w', qw), (f'l{l}.ca_o_w', ow)):
            hi, lo = w.t[name], w.t[name + '_lo']
            assert torch.equal(hi, round_tf32(hi)) and torch.equal(lo, round_tf32(lo))
            assert (hi + lo - ref).abs().max() <= 2e-7 * ref.abs().max()
        assert _close(w.t[f'l{l}.ca_q_b'], qb) and _close(w.t[f'l{l}.ca_o_b'], ob) and _close(w.t[f'l{l}.xa_o_b'], ob)
        in_w = sd[p + 'in_proj_weight']
        scale = 1.0 / math.sqrt(32)
        assert _close(w.t[f'l{l}.xa_q_w'], (scale * in_w[:256].double()).float(), 1)
        assert torch.equal(w.t[f'l{l}.xa_k_raw'], in_w[256:512]) and torch.equal(w.t[f'l{l}.xa_v_raw'], in_w[512:])
        assert torch.equal(w.t[f'l{l}.xa_k_w'], round_tf32(in_w[256:512]))
        assert torch.equal(w.t[f'l{l}.xa_k_w_lo'], round_tf32(in_w[256:512] - round_tf32(in_w[256:512])))
        assert torch.equal(w.t[f'l{l}.xa_o_w_hi'] + w.t[f'l{l}.xa_o_w_lo'], round_tf32(sd[p + 'out_proj.weight']) +
                           round_tf32(sd[p + 'out_proj.weight'] - round_tf32(sd[p + 'out_proj.weight'])))
        q = f'bbox_head.transformer.decoder.layers.{l}.'
        assert torch.equal(w.t[f'l{l}.sa_in_w'], sd[q + 'attentions.0.attn.in_proj_weight'])
        assert torch.equal(w.t[f'l{l}.ffn_b1'], sd[q + 'ffns.0.layers.0.0.bias'])
        assert torch.equal(w.t[f'l{l}.ffn_w1'], round_tf32(sd[q + 'ffns.0.layers.0.0.weight']))
        assert torch.equal(w.t[f'l{l}.ln_g2'], sd[q + 'norms.2.weight'])
    assert 'l0.sa_const' in w.t and 'l1.sa_const' not in w.t
    assert torch.equal(w.t['w_pos0'], round_tf32(sd['position_encoding.position_encoder.0.weight'].reshape(1024, 192)))
    assert torch.equal(w.t['w_enc0'][:, :1040], sd['query_generator.extra_enc.0.weight']) and not w.t['w_enc0'][:, 1040:].any()
    assert torch.equal(w.t['w_enc0_hi'], round_tf32(w.t['w_enc0']))
    assert torch.equal(w.t['br.reg_w1'][1], sd['bbox_head.reg_branches.1.2.weight'])
    assert torch.equal(w.t['br.cls_w0_hi'], round_tf32(w.t['br.cls_w0']))
    dim_t = torch.arange(128, dtype=torch.float32)
    assert torch.equal(w.t['dim_t'], 10000 ** (2 * (dim_t // 2) / 128))
    # the structs point into the arena
    base = w.arena.data_ptr()
    assert base <= w.layers[1].ca_o_w_lo < base + w.nbytes() and w.layers[1].sa_const is None
    assert w.layers[0].sa_const == w.t['l0.sa_const'].data_ptr() and w.branches.post_b == w.t['post_b'].data_ptr()
    assert PackedWeights(sd, 'cpu', fold_first_self_attn=False).layers[0].sa_const is None


def test_pack_reports_missing_and_misshapen_keys(state_dicts):
    import pytest
    sd = dict(state_dicts(1))
    bad = dict(sd)
    del bad['bbox_head.reg_branches.0.4.bias']
    with pytest.raises(RuntimeError, match='missing key .*reg_branches.0.4.bias'):
        PackedWeights(bad, 'cpu')
    bad = dict(sd)
    bad['query_generator.fc_center.weight'] = torch.zeros(4, 256)
    with pytest.raises(RuntimeError, match='fc_center.weight.* has 1024 elements, expected 768'):
        PackedWeights(bad, 'cpu')


def test_pack_neck():
    g = torch.Generator().manual_seed(3)
    sd = {'neck.lateral_convs.0.conv.weight': torch.randn(256, 256, 1, 1, generator=g), 'neck.lateral_convs.0.conv.bias': torch.randn(256, generator=g),
          'neck.fpn_convs.0.conv.weight': torch.randn(256, 256, 3, 3, generator=g), 'neck.fpn_convs.0.conv.bias': torch.randn(256, generator=g)}
    for d in (sd, {k[5:]: v for k, v in sd.items()}):
        n = PackedNeck(d, 'cpu')
        km = sd['neck.fpn_convs.0.conv.weight'].permute(0, 2, 3, 1).reshape(256, -1)
        assert torch.equal(n.t['fpn_w'], round_tf32(km)) and torch.equal(n.t['fpn_w_lo'], round_tf32(km - round_tf32(km)))
        assert torch.equal(n.t['lat_w'] + n.t['lat_w_lo'], round_tf32(sd['neck.lateral_convs.0.conv.weight'].reshape(256, 256)) +
                           round_tf32(sd['neck.lateral_convs.0.conv.weight'].reshape(256, 256) - n.t['lat_w']))
        assert torch.equal(n.t['fpn_b'], sd['neck.fpn_convs.0.conv.bias'])


def test_absorbed_cross_attention_equals_mha(state_dicts):
    sd = state_dicts(1)
    p = 'bbox_head.transformer.decoder.layers.0.attentions.1.attn.'
    g = torch.Generator().manual_seed(0)
    nq, nk = 9, 57
    x = torch.randn(nq, 256, generator=g)          # query + query_pos
    mem = torch.randn(nk, 256, generator=g)
    pos = torch.randn(nk, 256, generator=g)
    mask = torch.rand(nq, nk, generator=g) < 0.6    # True = masked
    mask[:, 0] = False
    ref = O._mha(sd, p, x[:, None], (mem + pos)[:, None], mem[:, None], 8, attn_mask=mask)[:, 0]
    w = PackedWeights(sd, 'cpu')
    qw, qb, ow, ob = w.t['l0.ca_q_w'] + w.t['l0.ca_q_w_lo'], w.t['l0.ca_q_b'], w.t['l0.ca_o_w'] + w.t['l0.ca_o_w_lo'], w.t['l0.ca_o_b']
    qt = (x @ qw.T + qb).view(nq, 8, 256)
    logits = torch.einsum('qhc,kc->qhk', qt, mem + pos).masked_fill(mask[:, None, :], float('-inf'))
    ctx = torch.einsum('qhk,kc->qhc', logits.softmax(-1), mem).reshape(nq, 2048)
    out = ctx @ ow.T + ob
    assert (out - ref).abs().max() < 2e-5


def test_conv_repack_matches_conv2d(state_dicts):
    sd = state_dicts(1)
    w = PackedWeights(sd, 'cpu')
    x = torch.randn(3, 256, 7, 7)
    ref = F.conv2d(x, sd['query_generator.shared_convs.0.conv.weight'],
                   sd['query_generator.shared_convs.0.conv.bias'], padding=1)
    tok = F.pad(x, (1, 1, 1, 1)).permute(0, 2, 3, 1)            # [n, 9, 9, c]
    cols = torch.stack([tok[:, ky:ky + 7, kx:kx + 7] for ky in range(3) for kx in range(3)], 3)  # [n,7,7,9,c]
    out = cols.reshape(3, 49, 9 * 256) @ (w.t['w_conv'] + w.t['w_conv_lo']).T + w.t['b_conv']   # hi + lo = w
    assert (out.view(3, 7, 7, 256).permute(0, 3, 1, 2) - ref).abs().max() < 2e-4
    assert w.num_layers == 1 and w.t['br.cls_w2'].shape == (1, 10, 256)


def test_first_layer_self_attention_is_a_constant(state_dicts):
    """The folding the staged decoder uses for layer 0: with target = 0 the FlattenMHSelfAttention output is
    out_proj(bv) + bo for every query, for any query_pos and any (non-degenerate) attention mask."""
    sd = state_dicts(2)
    p = 'bbox_head.transformer.decoder.layers.0.attentions.0.attn.'
    const = PackedWeights(sd, 'cpu').t['l0.sa_const']
    assert _close(const, first_layer_self_attn_const(sd[p + 'in_proj_bias'], sd[p + 'out_proj.weight'], sd[p + 'out_proj.bias']))
    g = torch.Generator().manual_seed(5)
    qpos = torch.randn(37, 1, 256, generator=g) * 3
    mask = torch.rand(37, 37, generator=g) < 0.5
    mask[torch.arange(37), torch.arange(37)] = False        # at least one visible key per row
    for m in (None, mask):
        sa = O._mha(sd, p, qpos, qpos, torch.zeros_like(qpos), 8, attn_mask=m)
        assert (sa[:, 0] - const[None]).abs().max() < 1e-6
