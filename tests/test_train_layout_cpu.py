"""CPU: the flat parameter / gradient layout of the training step (mv2d_train_param_info) against the reference's
state_dict -- every hot-path tensor has exactly one slot, slots do not overlap, start 64-byte aligned, and the host
mirror round-trips a state_dict bit for bit (including the re-laid-out 3x3 conv weight)."""
import pytest
import torch

from mv2d_b200 import synth
from mv2d_b200.train import CONV_W, FRONT_NAMES, GLOBAL_NAMES, LAYER_NAMES, DecoderTrainer, param_table


@pytest.mark.parametrize('L', [1, 2, 6])
def test_layout_covers_the_state_dict_once(L):
    table, total = param_table(L)
    sd = synth.make_state_dict(0, num_layers=L)
    learnable = {k for k in sd if k != 'bbox_head.code_weights'}          # a constant buffer, not a Parameter
    assert set(table) == learnable
    assert len(table) == len(GLOBAL_NAMES) + L * len(LAYER_NAMES) + len(FRONT_NAMES)
    spans = sorted((off, off + n, k) for k, (off, n) in table.items())
    assert spans[0][0] == 0
    for (a0, a1, ka), (b0, b1, kb) in zip(spans, spans[1:]):
        assert a1 <= b0, f'{ka} overlaps {kb}'
        assert b0 % 16 == 0, f'{kb} starts at {b0}: not a multiple of 16 floats'
    assert spans[-1][1] <= total
    for k, (off, n) in table.items():
        assert n == sd[k].numel(), k
    assert sum(n for _, n in table.values()) == sum(sd[k].numel() for k in learnable)


def test_host_mirror_round_trips_and_grad_views_alias_the_flat_buffer():
    sd = synth.make_state_dict(3, num_layers=2)
    tr = DecoderTrainer(sd, device='cpu')
    back = tr.state_dict()
    assert all(torch.equal(back[k], sd[k].float()) for k in back)
    # the conv weight lives as [c_out, ky, kx, c_in] in the flat buffer and is presented in the state_dict's layout
    off, n = tr.table[CONV_W]
    flat = tr.params[off:off + n].view(256, 3, 3, 256)
    assert torch.equal(flat.permute(0, 3, 1, 2), sd[CONV_W].float())
    tr.grads.zero_()
    tr.grad(CONV_W)[5, 7, 1, 2] = 3.0
    tr.grad('bbox_head.cls_branches.1.6.bias')[4] = 2.0
    assert float(tr.grads.sum()) == 5.0 and float(tr.grads[off:off + n].view(256, 3, 3, 256)[5, 1, 2, 7]) == 3.0
    with pytest.raises(RuntimeError):
        tr.forward(None, None, None, torch.zeros(1, 1), None, None, None)        # no CPU fallback


def test_clip_grad_norm_matches_torch_on_the_flat_buffer():
    sd = synth.make_state_dict(0, num_layers=1)
    tr = DecoderTrainer(sd, device='cpu')
    g = torch.Generator().manual_seed(1)
    for scale, max_norm in ((1.0, 35.0), (0.25, 3.0)):
        tr.grads.zero_()                       # the padding between tensors is never written by the kernels: stays zero
        for k in tr.table:
            tr.grad(k).copy_(torch.randn(tr.grad(k).shape, generator=g) * 0.05)
        ref = [tr.grad(k).clone() * scale for k in tr.table]
        params = [torch.nn.Parameter(torch.zeros_like(r)) for r in ref]
        for p_, r in zip(params, ref):
            p_.grad = r.clone()
        # the padding between tensors holds zeros, so the flat norm equals the norm over the tensors
        want = torch.nn.utils.clip_grad_norm_(params, max_norm)
        got = tr.clip_grad_norm_(max_norm, grad_scale=scale)
        exact = float(torch.sqrt(sum((r.double() ** 2).sum() for r in ref)))
        assert abs(float(got) - exact) <= 1e-6 * exact and abs(float(want) - exact) <= 5e-4 * exact
        for k, p_ in zip(tr.table, params):
            assert torch.allclose(tr.grad(k), p_.grad, rtol=1e-3, atol=1e-8), k


def test_decoder_tensors_form_the_leading_slice_of_the_flat_buffers():
    """TrainStep all-reduces the decoder's gradient slice while the front-end half of the backward still runs: every
    ``bbox_head.*`` tensor has to end before the first front-end tensor starts."""
    from mv2d_b200.train import param_table
    table, total = param_table(6)
    front = min(off for name, (off, n) in table.items() if not name.startswith('bbox_head.'))
    assert all(off + n <= front for name, (off, n) in table.items() if name.startswith('bbox_head.'))
    assert 0 < front < total and front % 16 == 0
