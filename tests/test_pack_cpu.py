"""CPU: host-side weight re-layout (mv2d_b200/pack.py).  The absorbed cross-attention matrices
must reproduce torch.nn.MultiheadAttention exactly in real arithmetic; checked here in fp32
against the oracle's MHA call on random inputs with a per-query key mask."""
import torch
import torch.nn.functional as F

from mv2d_b200.pack import PackedWeights, absorb_cross_attention, first_layer_self_attn_const
from oracle import mv2d_oracle as O


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
    qw, qb, ow, ob = absorb_cross_attention(sd[p + 'in_proj_weight'], sd[p + 'in_proj_bias'],
                                            sd[p + 'out_proj.weight'], sd[p + 'out_proj.bias'])
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
    const = first_layer_self_attn_const(sd[p + 'in_proj_bias'], sd[p + 'out_proj.weight'], sd[p + 'out_proj.bias'])
    g = torch.Generator().manual_seed(5)
    qpos = torch.randn(37, 1, 256, generator=g) * 3
    mask = torch.rand(37, 37, generator=g) < 0.5
    mask[torch.arange(37), torch.arange(37)] = False        # at least one visible key per row
    for m in (None, mask):
        sa = O._mha(sd, p, qpos, qpos, torch.zeros_like(qpos), 8, attn_mask=m)
        assert (sa[:, 0] - const[None]).abs().max() < 1e-6
