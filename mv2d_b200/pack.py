"""One-time re-layout of the reference ``state_dict`` into the buffers libmv2d_b200 consumes: a thin caller of the
library's ``mv2d_pack_weights`` / ``mv2d_pack_neck`` (csrc/pack.cpp, host code -- the same entry a C++ host uses).

Nothing here runs per sample.  Key names are the reference's (prefix ``roi_head.`` optional).
The library forms the derived matrices in fp64 and rounds them once to fp32:

* ``ca_q_w / ca_q_b``  -- cross-attention query side with the key projection absorbed: per head h,
  scale * Wk_h^T Wq_h  ([256 x 256]) and  scale * Wk_h^T bq_h, stacked to [2048, 256] / [2048].  The key bias only adds a
  per-(query, head) constant to the logits, which softmax cancels (utils/petr_transformer.py:503-508 -> torch
  MultiheadAttention).
* ``ca_o_w / ca_o_b``  -- output side with the value projection absorbed: per head Wo[:, 32h:32h+32] Wv_h stacked
  along K to [256, 2048], bias Wo bv + bo (probabilities sum to one).
* ``xa_*`` -- plain per-role projections for the key-stationary kernel: 1/sqrt(head_dim) folded into the query side,
  the key bias dropped, the value bias moved behind the softmax: out = Wo ctx + (Wo bv + bo).
* ``l0.sa_const`` -- self-attention output of decoder layer 0.  The target starts at zero (cross_attention_head.py:32),
  value = target, so every value row equals the value bias bv; a softmax-weighted mean of identical rows is that row,
  hence attn_out = out_proj(bv) + bo for every query regardless of q, k and masks (petr_transformer.py:314-370).
* the 3x3 convolutions are stored K-major with K ordered (tap, c_in); tensor-core operands are TF32 hi / lo splits.
"""
import ctypes as C

import torch

from . import lib

EMBED, HEADS, HD = 256, 8, 32

_SHAPES = {     # views handed out by .t[name]; anything not listed is 1-D
    'w_pos0': (1024, 192), 'w_pos2': (256, 1024), 'w_adapt0': (1024, 384), 'w_adapt2': (256, 1024),
    'w_se_reduce': (256, 256), 'w_se_expand': (256, 256), 'w_conv': (256, 2304), 'w_conv_lo': (256, 2304),
    'w_fc': (1024, 256), 'w_enc0': (512, 1056), 'w_enc2': (256, 512), 'w_center': (3, 256), 'w_qe0': (256, 384),
    'w_qe2': (256, 256), 'sa_in_w': (768, 256), 'sa_out_w': (256, 256), 'ca_q_w': (2048, 256), 'ca_o_w': (256, 2048),
    'ffn_w1': (2048, 256), 'ffn_w2': (256, 2048), 'xa_q_w': (256, 256), 'xa_k_w': (256, 256), 'xa_v_w': (256, 256),
    'xa_o_w': (256, 256), 'xa_k_raw': (256, 256), 'xa_v_raw': (256, 256), 'lat_w': (256, 256), 'fpn_w': (256, 2304),
    'cls_w0': (-1, 256, 256), 'cls_w1': (-1, 256, 256), 'reg_w0': (-1, 256, 256), 'reg_w1': (-1, 256, 256),
    'cls_w2': (-1, 10, 256), 'reg_w2': (-1, 10, 256), 'cls_b2': (-1, 10), 'reg_b2': (-1, 10), 'cls_b0': (-1, 256),
    'cls_g0': (-1, 256), 'cls_be0': (-1, 256), 'cls_b1': (-1, 256), 'cls_g1': (-1, 256), 'cls_be1': (-1, 256),
    'reg_b0': (-1, 256), 'reg_b1': (-1, 256)}


def _shape_of(name):
    base = name.split('.')[-1]
    for suffix in ('_hi', '_lo'):
        if base.endswith(suffix) and base[:-3] in _SHAPES:
            base = base[:-3]
    return _SHAPES.get(base)


def round_tf32(t):
    """Round-to-nearest (ties away from zero) fp32 -> TF32, kept in fp32: what cvt.rna.tf32.f32 does."""
    bits = t.detach().float().contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


def _named_tensors(state_dict, extra=None):
    """HOST fp32 contiguous copies of the state_dict + the ctypes array describing them (keeps both alive)."""
    keep, arr = [], []
    for k, v in list(state_dict.items()) + list((extra or {}).items()):
        if not torch.is_tensor(v) or not v.is_floating_point():
            continue
        t = v.detach().to('cpu', torch.float32).contiguous()
        keep.append((k.encode(), t))
    arr = (lib.NamedTensor * len(keep))()
    for i, (k, t) in enumerate(keep):
        arr[i].name, arr[i].data, arr[i].numel = k, t.data_ptr(), t.numel()
    return keep, arr


def _upload(dl, device, nbytes, n_entries, fill):
    """Allocate the device arena, let `fill(host_ptr, nbytes, device_base, dir, cap, n_dir)` write the host image,
    copy it over, and return (arena, {name: fp32 view})."""
    arena = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
    skip = -arena.data_ptr() % 256            # cudaMalloc is 256-byte aligned already; a host arena (tests) may not be
    arena = arena[skip:skip + nbytes]
    on_host = arena.device.type == 'cpu'
    host = arena if on_host else torch.empty(nbytes, dtype=torch.uint8)
    host.zero_()
    directory = (lib.PackedEntry * n_entries)()
    n_dir = C.c_int(0)
    lib.check(fill(host.data_ptr(), nbytes, arena.data_ptr(), directory, n_entries, C.byref(n_dir)), 'mv2d_pack')
    if not on_host:
        arena.copy_(host)
    views = {}
    for e in directory[:n_dir.value]:
        name = e.name.decode()
        v = arena[e.offset:e.offset + 4 * e.numel].view(torch.float32)
        shape = _shape_of(name)
        views[name] = v.view(*shape) if shape else v
    return arena, views


class PackedNeck:
    """``neck.*`` (one-level FPN) as the operands of mv2d_fpn_neck: TF32 hi / lo splits, the 3x3 kernel K-major with K
    ordered (ky, kx, c_in).  Accepts keys with or without the ``neck.`` prefix."""

    def __init__(self, state_dict, device):
        dl = lib.load()
        keep, arr = _named_tensors({k: v for k, v in state_dict.items() if 'lateral_convs.0' in k or 'fpn_convs.0' in k})
        n = C.c_int(0)
        nbytes = dl.mv2d_pack_neck_bytes(C.byref(n))
        self.arena, self.t = _upload(dl, device, nbytes, n.value, lambda host, nb, base, d, cap, nd:
                                     dl.mv2d_pack_neck(arr, len(keep), host, nb, base, d, cap, nd))

    def p(self, name):
        return self.t[name].data_ptr()


class PackedWeights:
    """Device-resident weights (one arena) + the host-side ctypes structs that point into it."""

    def __init__(self, state_dict, device, num_layers=None, fold_first_self_attn=True):
        dl = lib.load()
        if num_layers is None:
            num_layers = 1 + max(int(k.split('decoder.layers.')[1].split('.')[0]) for k in state_dict
                                 if 'bbox_head.transformer.decoder.layers.' in k)
        assert 1 <= num_layers <= lib.MAX_LAYERS
        self.num_layers = num_layers
        self.device = device
        # the sine embeddings' frequency table, computed with the same torch CPU ops as the reference
        dim_t = torch.arange(128, dtype=torch.float32)
        dim_t = 10000 ** (2 * (dim_t // 2) / 128)                 # pe.py:24-25, positional_encoding.py:78-80
        keep, arr = _named_tensors(state_dict, {'dim_t': dim_t})
        n = C.c_int(0)
        nbytes = dl.mv2d_pack_weights_bytes(num_layers, int(fold_first_self_attn), C.byref(n))
        assert nbytes > 0
        self.layers = (lib.LayerWeights * num_layers)()
        self.branches = lib.BranchWeights()
        self.arena, self.t = _upload(dl, device, nbytes, n.value, lambda host, nb, base, d, cap, nd:
                                     dl.mv2d_pack_weights(arr, len(keep), num_layers, int(fold_first_self_attn), host, nb, base,
                                                          d, cap, nd, self.layers, C.byref(self.branches)))

    def p(self, name):
        return self.t[name].data_ptr()

    def nbytes(self):
        return self.arena.numel()

    def layers_ptr(self):
        return C.cast(self.layers, C.POINTER(lib.LayerWeights))

    def branches_ptr(self):
        return C.pointer(self.branches)
