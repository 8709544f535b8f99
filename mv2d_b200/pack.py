"""One-time re-layout of the reference ``state_dict`` into the buffers libmv2d_b200 consumes.

Nothing here runs per sample.  Key names are the reference's (prefix
``roi_head.`` optional).  Derived matrices are formed in fp64 and rounded once to fp32:

* ``ca_q_w / ca_q_b``  -- cross-attention query side with the key projection absorbed:
  per head h,  scale * Wk_h^T Wq_h  ([256 x 256]) and  scale * Wk_h^T bq_h, stacked to
  [2048, 256] / [2048].  The key bias only adds a per-(query, head) constant to the logits,
  which softmax cancels (utils/petr_transformer.py:503-508 -> torch MultiheadAttention).
* ``ca_o_w / ca_o_b``  -- output side with the value projection absorbed: per head
  Wo[:, 32h:32h+32] Wv_h stacked along K to [256, 2048], bias Wo bv + bo (probabilities
  sum to one).
* the 3x3 conv of the query generator is stored K-major with K ordered (tap, c_in).
"""
import ctypes as C
import math

import torch

from . import lib

EMBED, HEADS, HD = 256, 8, 32


def _strip(sd):
    out = {}
    for k, v in sd.items():
        if k.startswith('roi_head.'):
            k = k[len('roi_head.'):]
        out[k] = v
    return out


def absorb_cross_attention(in_w, in_b, out_w, out_b):
    """fp64 construction of the absorbed cross-attention matrices (see module docstring)."""
    in_w, in_b, out_w, out_b = [t.detach().double().cpu() for t in (in_w, in_b, out_w, out_b)]
    wq, wk, wv = in_w[:EMBED], in_w[EMBED:2 * EMBED], in_w[2 * EMBED:]
    bq, bv = in_b[:EMBED], in_b[2 * EMBED:]
    scale = 1.0 / math.sqrt(HD)
    qw, qb, ow = [], [], []
    for h in range(HEADS):
        s = slice(h * HD, (h + 1) * HD)
        qw.append(scale * wk[s].T @ wq[s])          # [256(key dim), 256(x dim)]
        qb.append(scale * wk[s].T @ bq[s])          # [256]
        ow.append(out_w[:, s] @ wv[s])              # [256(out), 256(mem dim)]
    ca_q_w = torch.cat(qw, 0)                       # [2048, 256]
    ca_q_b = torch.cat(qb, 0)                       # [2048]
    ca_o_w = torch.cat(ow, 1)                       # [256, 2048]
    ca_o_b = out_w @ bv + out_b
    return ca_q_w.float(), ca_q_b.float(), ca_o_w.float(), ca_o_b.float()


def plain_cross_attention(in_w, in_b, out_w, out_b):
    """Per-role cross-attention projections for the key-stationary kernel: the 1/sqrt(head_dim) scale is
    folded into the query side, the key bias is dropped (it shifts all logits of a (query, head) by the same
    amount) and the value bias moves behind the softmax: out = Wo ctx + (Wo bv + bo).  fp64, rounded once."""
    in_w, in_b, out_w, out_b = [t.detach().double().cpu() for t in (in_w, in_b, out_w, out_b)]
    scale = 1.0 / math.sqrt(HD)
    wq, wk, wv = in_w[:EMBED], in_w[EMBED:2 * EMBED], in_w[2 * EMBED:]
    bq, bv = in_b[:EMBED], in_b[2 * EMBED:]
    return ((scale * wq).float(), (scale * bq).float(), wk.float(), wv.float(), out_w.float(),
            (out_w @ bv + out_b).float())


def first_layer_self_attn_const(in_proj_bias, out_w, out_b):
    """Self-attention output of decoder layer 0.  The target starts at zero (cross_attention_head.py:32,
    ``target = torch.zeros_like(query_embed)``), value = target, so every value row equals the value bias bv;
    a softmax-weighted mean of identical rows is that row, hence attn_out = out_proj(bv) + bo for every query
    regardless of q, k and masks (petr_transformer.py:314-370).  fp64, rounded once."""
    E = out_w.shape[0]
    bv = in_proj_bias.double()[2 * E:]
    return (out_w.double() @ bv + out_b.double()).float()


def round_tf32(t):
    """Round-to-nearest (ties away from zero) fp32 -> TF32, kept in fp32: what cvt.rna.tf32.f32 does."""
    bits = t.detach().float().contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


def split_tf32(t):
    """w = hi + lo with both parts exactly TF32-representable (operands of the 3xTF32 GEMM)."""
    t = t.detach().float()
    hi = round_tf32(t)
    return hi, round_tf32(t - hi)


class PackedNeck:
    """``neck.*`` (one-level FPN) as the operands of mv2d_fpn_neck: TF32 hi / lo splits, the 3x3 kernel K-major with K
    ordered (ky, kx, c_in).  Accepts keys with or without the ``neck.`` prefix."""

    def __init__(self, state_dict, device):
        sd = {(k[len('neck.'):] if k.startswith('neck.') else k): v for k, v in state_dict.items()}
        lat, fpn = sd['lateral_convs.0.conv.weight'], sd['fpn_convs.0.conv.weight']
        assert tuple(lat.shape) == (EMBED, EMBED, 1, 1) and tuple(fpn.shape) == (EMBED, EMBED, 3, 3), \
            'the MV2D neck is a one-level FPN, 256 -> 256 (configs/mv2d/exp/*.py:32-39)'
        self.t = {}

        def put(name, t):
            self.t[name] = t.detach().float().contiguous().to(device)

        hi, lo = split_tf32(lat.reshape(EMBED, EMBED))
        put('lat_w', hi); put('lat_w_lo', lo); put('lat_b', sd['lateral_convs.0.conv.bias'])
        hi, lo = split_tf32(fpn.permute(0, 2, 3, 1).reshape(EMBED, -1))
        put('fpn_w', hi); put('fpn_w_lo', lo); put('fpn_b', sd['fpn_convs.0.conv.bias'])

    def p(self, name):
        return self.t[name].data_ptr()


class PackedWeights:
    """Device-resident weights + the host-side ctypes structs that point at them."""

    def __init__(self, state_dict, device, num_layers=None, fold_first_self_attn=True):
        sd = _strip(state_dict)
        if num_layers is None:
            num_layers = 1 + max(int(k.split('.')[4]) for k in sd
                                 if k.startswith('bbox_head.transformer.decoder.layers.'))
        assert 1 <= num_layers <= lib.MAX_LAYERS
        self.num_layers = num_layers
        self.device = device
        self.t = {}   # name -> device tensor (keeps the storage alive)

        def put(name, tensor):
            self.t[name] = tensor.detach().float().contiguous().to(device)
            return self.t[name]

        def conv1x1(key):   # PE MLPs run as single-pass TF32 tensor-core GEMMs: weights pre-rounded
            return round_tf32(sd[key].reshape(sd[key].shape[0], -1))

        pe = 'position_encoding.'
        put('w_pos0', conv1x1(pe + 'position_encoder.0.weight')); put('b_pos0', sd[pe + 'position_encoder.0.bias'])
        put('w_pos2', conv1x1(pe + 'position_encoder.2.weight')); put('b_pos2', sd[pe + 'position_encoder.2.bias'])
        put('w_adapt0', conv1x1(pe + 'adapt_pos3d.0.weight')); put('b_adapt0', sd[pe + 'adapt_pos3d.0.bias'])
        put('w_adapt2', conv1x1(pe + 'adapt_pos3d.2.weight')); put('b_adapt2', sd[pe + 'adapt_pos3d.2.bias'])
        put('w_se_reduce', conv1x1(pe + 'fpe.conv_reduce.weight')); put('b_se_reduce', sd[pe + 'fpe.conv_reduce.bias'])
        put('w_se_expand', conv1x1(pe + 'fpe.conv_expand.weight')); put('b_se_expand', sd[pe + 'fpe.conv_expand.bias'])
        qg = 'query_generator.'
        wc = sd[qg + 'shared_convs.0.conv.weight']           # [co, ci, ky, kx]
        wc_hi, wc_lo = split_tf32(wc.permute(0, 2, 3, 1).reshape(wc.shape[0], -1))  # [co, (ky,kx,ci)]
        put('w_conv', wc_hi)
        put('w_conv_lo', wc_lo)
        put('b_conv', sd[qg + 'shared_convs.0.conv.bias'])
        put('w_fc', sd[qg + 'shared_fcs.0.weight']); put('b_fc', sd[qg + 'shared_fcs.0.bias'])
        put('w_enc0', torch.nn.functional.pad(sd[qg + 'extra_enc.0.weight'].float(), (0, 16)))   # K 1040 -> 1056
        put('b_enc0', sd[qg + 'extra_enc.0.bias'])
        put('w_enc2', sd[qg + 'extra_enc.2.weight']); put('b_enc2', sd[qg + 'extra_enc.2.bias'])
        put('w_center', sd[qg + 'fc_center.weight']); put('b_center', sd[qg + 'fc_center.bias'])
        bh = 'bbox_head.'
        put('w_qe0', sd[bh + 'query_embedding.0.weight']); put('b_qe0', sd[bh + 'query_embedding.0.bias'])
        put('w_qe2', sd[bh + 'query_embedding.2.weight']); put('b_qe2', sd[bh + 'query_embedding.2.bias'])

        for name in ('w_fc', 'w_enc0', 'w_enc2', 'w_qe0', 'w_qe2'):     # 3xTF32 operands of the FC chain for batches
            hi, lo = split_tf32(self.t[name])
            put(name + '_hi', hi); put(name + '_lo', lo)

        self.layers = (lib.LayerWeights * num_layers)()
        for l in range(num_layers):
            p = f'{bh}transformer.decoder.layers.{l}.'
            lw = self.layers[l]
            lw.sa_in_w = put(f'l{l}.sa_in_w', sd[p + 'attentions.0.attn.in_proj_weight']).data_ptr()
            lw.sa_in_b = put(f'l{l}.sa_in_b', sd[p + 'attentions.0.attn.in_proj_bias']).data_ptr()
            lw.sa_out_w = put(f'l{l}.sa_out_w', sd[p + 'attentions.0.attn.out_proj.weight']).data_ptr()
            lw.sa_out_b = put(f'l{l}.sa_out_b', sd[p + 'attentions.0.attn.out_proj.bias']).data_ptr()
            qw, qb, ow, ob = absorb_cross_attention(
                sd[p + 'attentions.1.attn.in_proj_weight'], sd[p + 'attentions.1.attn.in_proj_bias'],
                sd[p + 'attentions.1.attn.out_proj.weight'], sd[p + 'attentions.1.attn.out_proj.bias'])
            for field, mat in (('ca_q_w', qw), ('ca_o_w', ow), ('ffn_w1', sd[p + 'ffns.0.layers.0.0.weight']),
                               ('ffn_w2', sd[p + 'ffns.0.layers.1.weight'])):
                hi, lo = split_tf32(mat)     # 3xTF32 tcgen05 operands
                setattr(lw, field, put(f'l{l}.{field}', hi).data_ptr())
                setattr(lw, field + '_lo', put(f'l{l}.{field}_lo', lo).data_ptr())
            # plain projections for the key-stationary form of the two-frame head (xa_tile.cuh)
            xq_w, xq_b, xk_w, xv_w, xo_w, xo_b = plain_cross_attention(
                sd[p + 'attentions.1.attn.in_proj_weight'], sd[p + 'attentions.1.attn.in_proj_bias'],
                sd[p + 'attentions.1.attn.out_proj.weight'], sd[p + 'attentions.1.attn.out_proj.bias'])
            lw.xa_q_w = put(f'l{l}.xa_q_w', xq_w).data_ptr()
            lw.xa_q_b = put(f'l{l}.xa_q_b', xq_b).data_ptr()
            for field, mat in (('xa_k_w', xk_w), ('xa_v_w', xv_w)):
                hi, lo = split_tf32(mat)
                setattr(lw, field, put(f'l{l}.{field}', hi).data_ptr())
                setattr(lw, field + '_lo', put(f'l{l}.{field}_lo', lo).data_ptr())
            # hi / lo splits of the four [*,256] matrices the small-M path multiplies with FFMA: batches of more than
            # 512 query rows run them as 3xTF32 tensor-core GEMMs as well
            for field, mat in (('sa_in_w', sd[p + 'attentions.0.attn.in_proj_weight']),
                               ('sa_out_w', sd[p + 'attentions.0.attn.out_proj.weight']), ('xa_q_w', xq_w), ('xa_o_w', xo_w)):
                hi, lo = split_tf32(mat)
                setattr(lw, field + '_hi', put(f'l{l}.{field}_hi', hi).data_ptr())
                setattr(lw, field + '_lo', put(f'l{l}.{field}_lo', lo).data_ptr())
            lw.xa_k_raw = put(f'l{l}.xa_k_raw', xk_w).data_ptr()
            lw.xa_v_raw = put(f'l{l}.xa_v_raw', xv_w).data_ptr()
            lw.xa_o_w = put(f'l{l}.xa_o_w', xo_w).data_ptr()
            lw.xa_o_b = put(f'l{l}.xa_o_b', xo_b).data_ptr()
            lw.ca_q_b = put(f'l{l}.ca_q_b', qb).data_ptr()
            lw.ca_o_b = put(f'l{l}.ca_o_b', ob).data_ptr()
            lw.ffn_b1 = put(f'l{l}.ffn_b1', sd[p + 'ffns.0.layers.0.0.bias']).data_ptr()
            lw.ffn_b2 = put(f'l{l}.ffn_b2', sd[p + 'ffns.0.layers.1.bias']).data_ptr()
            for n in range(3):
                lw.ln_g[n] = put(f'l{l}.ln_g{n}', sd[p + f'norms.{n}.weight']).data_ptr()
                lw.ln_b[n] = put(f'l{l}.ln_b{n}', sd[p + f'norms.{n}.bias']).data_ptr()
            if l == 0 and fold_first_self_attn:
                lw.sa_const = put('l0.sa_const', first_layer_self_attn_const(
                    sd[p + 'attentions.0.attn.in_proj_bias'], sd[p + 'attentions.0.attn.out_proj.weight'],
                    sd[p + 'attentions.0.attn.out_proj.bias'])).data_ptr()

        def stack(fmt):
            return torch.stack([sd[fmt.format(l)] for l in range(num_layers)], 0)

        b = lib.BranchWeights()
        for field, fmt in [
                ('cls_w0', bh + 'cls_branches.{}.0.weight'), ('cls_b0', bh + 'cls_branches.{}.0.bias'),
                ('cls_g0', bh + 'cls_branches.{}.1.weight'), ('cls_be0', bh + 'cls_branches.{}.1.bias'),
                ('cls_w1', bh + 'cls_branches.{}.3.weight'), ('cls_b1', bh + 'cls_branches.{}.3.bias'),
                ('cls_g1', bh + 'cls_branches.{}.4.weight'), ('cls_be1', bh + 'cls_branches.{}.4.bias'),
                ('cls_w2', bh + 'cls_branches.{}.6.weight'), ('cls_b2', bh + 'cls_branches.{}.6.bias'),
                ('reg_w0', bh + 'reg_branches.{}.0.weight'), ('reg_b0', bh + 'reg_branches.{}.0.bias'),
                ('reg_w1', bh + 'reg_branches.{}.2.weight'), ('reg_b1', bh + 'reg_branches.{}.2.bias'),
                ('reg_w2', bh + 'reg_branches.{}.4.weight'), ('reg_b2', bh + 'reg_branches.{}.4.bias')]:
            setattr(b, field, put('br.' + field, stack(fmt)).data_ptr())
        for field in ('cls_w0', 'cls_w1', 'reg_w0', 'reg_w1'):
            hi, lo = split_tf32(self.t['br.' + field])
            setattr(b, field + '_hi', put(f'br.{field}_hi', hi).data_ptr())
            setattr(b, field + '_lo', put(f'br.{field}_lo', lo).data_ptr())
        b.post_g = put('post_g', sd[bh + 'transformer.decoder.post_norm.weight']).data_ptr()
        b.post_b = put('post_b', sd[bh + 'transformer.decoder.post_norm.bias']).data_ptr()
        self.branches = b

        # small constant tables, computed with the same torch CPU ops as the reference
        dim_t = torch.arange(128, dtype=torch.float32)
        put('dim_t', 10000 ** (2 * (dim_t // 2) / 128))          # pe.py:24-25, positional_encoding.py:78-80

    def p(self, name):
        return self.t[name].data_ptr()

    def nbytes(self):
        return sum(v.numel() * v.element_size() for v in self.t.values())

    def layers_ptr(self):
        return C.cast(self.layers, C.POINTER(lib.LayerWeights))

    def branches_ptr(self):
        return C.pointer(self.branches)
