"""ctypes binding of ``libmv2d_b200.so`` (C ABI declared in ``include/mv2d_b200.h``).

The product path has NO fallback: if the shared library is missing or a call fails this module
raises.  The structures below mirror the header field by field; ``_check_layout`` compares
their sizes with ``mv2d_sizeof`` at load time so a drifted mirror fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libmv2d_b200.so')
ABI_VERSION = 5
MAX_LAYERS = 8

c_f = C.c_void_p  # device pointers travel as integers (tensor.data_ptr())


class PeParams(C.Structure):
    _fields_ = [
        ('V', C.c_int), ('h', C.c_int), ('w', C.c_int), ('depth_num', C.c_int),
        ('pad_h', C.c_int), ('pad_w', C.c_int), ('stride', C.c_int), ('phase', C.c_int),
        ('depth_start', C.c_double), ('position_range', C.c_double * 6),
        ('feat', c_f), ('feat_tf32', c_f), ('img2lidar', c_f), ('not_mask', c_f), ('dim_t', c_f),
        ('w_pos0', c_f), ('b_pos0', c_f), ('w_pos2', c_f), ('b_pos2', c_f),
        ('w_adapt0', c_f), ('b_adapt0', c_f), ('w_adapt2', c_f), ('b_adapt2', c_f),
        ('w_se_reduce', c_f), ('b_se_reduce', c_f), ('w_se_expand', c_f), ('b_se_expand', c_f),
        ('sine_branch_cached', c_f), ('sine_branch_out', c_f),
        ('pe', c_f), ('kin', c_f), ('workspace', c_f), ('workspace_bytes', C.c_size_t),
        ('sine_separable', C.c_int), ('views_per_sample', C.c_int), ('sine_shared', C.c_int), ('unfused_mlp', C.c_int),
    ]


class QgParams(C.Structure):
    _fields_ = [
        ('N', C.c_int), ('V', C.c_int), ('h', C.c_int), ('w', C.c_int), ('stride', C.c_int),
        ('phase', C.c_int),
        ('pc_range', C.c_float * 6), ('intrins_feat_scale', C.c_float), ('reserved1', C.c_float),
        ('rois', c_f), ('intrinsics', c_f), ('extrinsics', c_f), ('feat', c_f), ('pe', c_f),
        ('dim_t', c_f),
        ('w_conv', c_f), ('b_conv', c_f), ('w_conv_lo', c_f), ('w_fc', c_f), ('b_fc', c_f),
        ('w_enc0', c_f), ('b_enc0', c_f), ('w_enc2', c_f), ('b_enc2', c_f),
        ('w_center', c_f), ('b_center', c_f), ('w_qe0', c_f), ('b_qe0', c_f),
        ('w_qe2', c_f), ('b_qe2', c_f),
        ('tok_feat', c_f), ('tok_kin', c_f), ('roi_intrinsics', c_f), ('center_lidar', c_f),
        ('ref', c_f), ('query_pos', c_f), ('workspace', c_f), ('workspace_bytes', C.c_size_t),
        ('w_fc_hi', c_f), ('w_fc_lo', c_f), ('w_enc0_hi', c_f), ('w_enc0_lo', c_f), ('w_enc2_hi', c_f), ('w_enc2_lo', c_f),
        ('w_qe0_hi', c_f), ('w_qe0_lo', c_f), ('w_qe2_hi', c_f), ('w_qe2_lo', c_f),
        ('roi_extrinsics', c_f), ('intrins_feat', c_f), ('enc_out', c_f),
    ]


class CorrParams(C.Structure):
    _fields_ = [
        ('N', C.c_int), ('V', C.c_int), ('img_h', C.c_int), ('img_w', C.c_int),
        ('topk', C.c_int), ('sample_size', C.c_int), ('num_depth', C.c_int), ('max_match', C.c_int),
        ('ratio', C.c_float), ('iou_thr', C.c_float), ('depth_start', C.c_float),
        ('reserved0', C.c_float),
        ('rois', c_f), ('roi_start', c_f), ('trans', c_f), ('lin', c_f), ('depths', c_f),
        ('match', c_f), ('match_cnt', c_f),
        ('h', C.c_int), ('w', C.c_int), ('stride', C.c_int), ('expand_stride', C.c_int),
        ('pad_mask', c_f), ('keymask', c_f), ('key_cnt', c_f), ('key_list', c_f),
        ('batch', C.c_int), ('rows_per_sample', C.c_int),
    ]


class LayerWeights(C.Structure):
    _fields_ = [
        ('sa_in_w', c_f), ('sa_in_b', c_f), ('sa_out_w', c_f), ('sa_out_b', c_f),
        ('ca_q_w', c_f), ('ca_q_w_lo', c_f), ('ca_q_b', c_f), ('ca_o_w', c_f), ('ca_o_w_lo', c_f), ('ca_o_b', c_f),
        ('ffn_w1', c_f), ('ffn_w1_lo', c_f), ('ffn_b1', c_f), ('ffn_w2', c_f), ('ffn_w2_lo', c_f), ('ffn_b2', c_f),
        ('ln_g', c_f * 3), ('ln_b', c_f * 3), ('sa_const', c_f),
        ('xa_q_w', c_f), ('xa_q_b', c_f), ('xa_k_w', c_f), ('xa_k_w_lo', c_f), ('xa_v_w', c_f), ('xa_v_w_lo', c_f),
        ('xa_o_w', c_f), ('xa_o_b', c_f), ('xa_k_raw', c_f), ('xa_v_raw', c_f),
        ('sa_in_w_hi', c_f), ('sa_in_w_lo', c_f), ('sa_out_w_hi', c_f), ('sa_out_w_lo', c_f),
        ('xa_q_w_hi', c_f), ('xa_q_w_lo', c_f), ('xa_o_w_hi', c_f), ('xa_o_w_lo', c_f),
    ]


class BranchWeights(C.Structure):
    _fields_ = [
        ('cls_w0', c_f), ('cls_b0', c_f), ('cls_g0', c_f), ('cls_be0', c_f),
        ('cls_w1', c_f), ('cls_b1', c_f), ('cls_g1', c_f), ('cls_be1', c_f),
        ('cls_w2', c_f), ('cls_b2', c_f),
        ('reg_w0', c_f), ('reg_b0', c_f), ('reg_w1', c_f), ('reg_b1', c_f),
        ('reg_w2', c_f), ('reg_b2', c_f),
        ('post_g', c_f), ('post_b', c_f),
        ('cls_w0_hi', c_f), ('cls_w0_lo', c_f), ('cls_w1_hi', c_f), ('cls_w1_lo', c_f),
        ('reg_w0_hi', c_f), ('reg_w0_lo', c_f), ('reg_w1_hi', c_f), ('reg_w1_lo', c_f),
    ]


class DecoderParams(C.Structure):
    _fields_ = [
        ('N', C.c_int), ('L', C.c_int), ('mode', C.c_int), ('num_rows', C.c_int),
        ('max_match', C.c_int), ('mask_words', C.c_int), ('persistent', C.c_int),
        ('vel_row_start', C.c_int),
        ('pc_range', C.c_float * 6), ('vel_dt', C.c_float), ('reserved2', C.c_float),
        ('query_pos', c_f), ('ref', c_f), ('kin_rows', c_f), ('mem_rows', c_f),
        ('match', c_f), ('match_cnt', c_f), ('keymask', c_f), ('key_list', c_f), ('key_cnt', c_f),
        ('self_attn_mask', c_f),
        ('layers', C.POINTER(LayerWeights)), ('branches', C.POINTER(BranchWeights)),
        ('cls_scores', c_f), ('bbox_preds', c_f), ('outs_dec', c_f),
        ('workspace', c_f), ('workspace_bytes', C.c_size_t),
        ('layer_begin', C.c_int), ('layer_end', C.c_int), ('xa_form', C.c_int), ('grid_h', C.c_int),
        ('grid_w', C.c_int), ('xa_prepared', C.c_int),
        ('kp', c_f), ('vp', c_f), ('xa_workspace', c_f), ('xa_workspace_bytes', C.c_size_t),
        ('row_tile_live', c_f),
        ('batch', C.c_int), ('rows_per_sample', C.c_int), ('n_real', c_f), ('vel_dt_batch', c_f),
    ]


class KvParams(C.Structure):
    _fields_ = [
        ('num_rows', C.c_int), ('L', C.c_int), ('layer_begin', C.c_int), ('layer_end', C.c_int),
        ('kin_hi', c_f), ('kin_lo', c_f), ('mem_hi', c_f), ('mem_lo', c_f),
        ('layers', C.POINTER(LayerWeights)), ('kp', c_f), ('vp', c_f), ('row_tile_live', c_f),
    ]


class DnParams(C.Structure):
    _fields_ = [
        ('N', C.c_int), ('G', C.c_int), ('scalar', C.c_int), ('num_classes', C.c_int),
        ('mode', C.c_int), ('max_match', C.c_int), ('max_match_all', C.c_int), ('mask_words', C.c_int),
        ('train_unmask', C.c_int), ('reserved0', C.c_int),
        ('noise_scale', C.c_float), ('noise_trans', C.c_float), ('split', C.c_float), ('eps', C.c_float),
        ('pc_range', C.c_float * 6),
        ('gt_boxes', c_f), ('gt_labels', c_f), ('rand', c_f), ('ref', c_f),
        ('match', c_f), ('match_cnt', c_f), ('keymask', c_f), ('key_cnt', c_f),
        ('w_qe0', c_f), ('b_qe0', c_f), ('w_qe2', c_f), ('b_qe2', c_f), ('dim_t', c_f),
        ('ref_all', c_f), ('dn_labels', c_f), ('attn_mask', c_f), ('query_pos_all', c_f),
        ('match_all', c_f), ('match_cnt_all', c_f),
        ('keymask_all', c_f), ('key_list_all', c_f), ('key_cnt_all', c_f),
        ('workspace', c_f), ('workspace_bytes', C.c_size_t),
    ]


class LossParams(C.Structure):
    _fields_ = [
        ('N', C.c_int), ('G', C.c_int), ('L', C.c_int), ('num_classes', C.c_int),
        ('pad', C.c_int), ('neg_bbox_loss', C.c_int),
        ('layer_stride', C.c_longlong), ('dn_layer_stride', C.c_longlong),
        ('cls_cost_weight', C.c_float), ('reg_cost_weight', C.c_float), ('cls_loss_weight', C.c_float),
        ('bbox_loss_weight', C.c_float),
        ('focal_alpha', C.c_float), ('focal_gamma', C.c_float), ('dn_split', C.c_float), ('reserved1', C.c_float),
        ('code_weights', C.c_float * 10),
        ('cls_scores', c_f), ('bbox_preds', c_f), ('gt_boxes', c_f), ('gt_labels', c_f),
        ('dn_cls', c_f), ('dn_box', c_f), ('dn_labels', c_f),
        ('assigned', c_f), ('losses', c_f), ('workspace', c_f), ('workspace_bytes', C.c_size_t),
        ('num_pos', c_f), ('bbox_avg_factor', c_f),
    ]


class NeckParams(C.Structure):
    _fields_ = [
        ('V', C.c_int), ('h', C.c_int), ('w', C.c_int), ('in_is_nhwc', C.c_int),
        ('x', c_f), ('lat_w', c_f), ('lat_w_lo', c_f), ('lat_b', c_f),
        ('fpn_w', c_f), ('fpn_w_lo', c_f), ('fpn_b', c_f),
        ('feat', c_f), ('feat_tf32', c_f), ('workspace', c_f), ('workspace_bytes', C.c_size_t),
    ]


class TrainParams(C.Structure):
    _fields_ = [
        ('N', C.c_int), ('L', C.c_int), ('max_match', C.c_int), ('G', C.c_int),
        ('num_classes', C.c_int), ('reserved0', C.c_int),
        ('pc_range', C.c_float * 6),
        ('cls_cost_weight', C.c_float), ('reg_cost_weight', C.c_float), ('cls_loss_weight', C.c_float),
        ('bbox_loss_weight', C.c_float), ('focal_alpha', C.c_float), ('focal_gamma', C.c_float),
        ('code_weights', C.c_float * 10),
        ('stage_loss_weights', C.c_float * MAX_LAYERS),
        ('params', c_f), ('grads', c_f), ('dim_t', c_f), ('ref', c_f), ('tok_kin', c_f), ('tok_mem', c_f),
        ('match', c_f), ('match_cnt', c_f), ('gt_boxes', c_f), ('gt_labels', c_f),
        ('cls_scores', c_f), ('bbox_preds', c_f), ('assigned', c_f), ('losses', c_f),
        ('d_ref', c_f), ('d_tok_kin', c_f), ('d_tok_mem', c_f),
        ('workspace', c_f), ('workspace_bytes', C.c_size_t),
        ('num_pos', c_f), ('bbox_avg_factor', c_f),
        ('mode', C.c_int), ('pad', C.c_int), ('num_rows', C.c_int), ('mask_words', C.c_int),
        ('neg_bbox_loss', C.c_int), ('reserved5', C.c_int),
        ('vel_dt', C.c_float), ('dn_split', C.c_float), ('denoise_weight', C.c_float), ('reserved6', C.c_float),
        ('kin_map', c_f), ('mem_map', c_f), ('keymask', c_f), ('key_list', c_f), ('key_cnt', c_f),
        ('self_attn_mask', c_f), ('dn_labels', c_f), ('d_kin_map', c_f), ('d_mem_map', c_f),
    ]


class FrontTrainParams(C.Structure):
    _fields_ = [
        ('N', C.c_int), ('V', C.c_int), ('h', C.c_int), ('w', C.c_int), ('L', C.c_int), ('stride', C.c_int),
        ('depth_num', C.c_int), ('pad_h', C.c_int), ('pad_w', C.c_int), ('reserved0', C.c_int),
        ('depth_start', C.c_double), ('position_range', C.c_double * 6),
        ('pc_range', C.c_float * 6), ('intrins_feat_scale', C.c_float), ('reserved1', C.c_float),
        ('params', c_f), ('grads', c_f), ('rois', c_f), ('roi_intrinsics', c_f), ('extrinsics', c_f), ('img2lidar', c_f),
        ('not_mask', c_f), ('dim_t', c_f), ('feat', c_f),
        ('tok_mem', c_f), ('tok_kin', c_f), ('ref', c_f), ('pe_out', c_f),
        ('d_ref', c_f), ('d_tok_kin', c_f), ('d_tok_mem', c_f), ('d_feat', c_f),
        ('workspace', c_f), ('workspace_bytes', C.c_size_t),
        ('d_pe_extra', c_f), ('d_feat_extra', c_f), ('d_feat_extra2', c_f), ('kin_out', c_f),
    ]


class NamedTensor(C.Structure):      # Mv2dNamedTensor: a HOST fp32 tensor under its reference state_dict key
    _fields_ = [('name', C.c_char_p), ('data', C.c_void_p), ('numel', C.c_int64)]


class PackedEntry(C.Structure):      # Mv2dPackedEntry: one buffer of the packed arena
    _fields_ = [('name', C.c_char * 48), ('offset', C.c_int64), ('numel', C.c_int64)]


_STRUCTS = [PeParams, QgParams, CorrParams, DecoderParams, LayerWeights, BranchWeights, DnParams, KvParams, LossParams, NeckParams,
            TrainParams, FrontTrainParams]

# every symbol include/mv2d_b200.h declares: (name, restype, argtypes)
SYMBOLS = [
    ('mv2d_abi_version', C.c_int, []),
    ('mv2d_last_error', C.c_char_p, []),
    ('mv2d_launch_count', C.c_ulonglong, []),
    ('mv2d_sizeof', C.c_size_t, [C.c_int]),
    ('mv2d_pack_weights_bytes', C.c_int64, [C.c_int, C.c_int, C.POINTER(C.c_int)]),
    ('mv2d_pack_weights', C.c_int, [C.POINTER(NamedTensor), C.c_int, C.c_int, C.c_int, c_f, C.c_int64, c_f, C.POINTER(PackedEntry),
                                    C.c_int, C.POINTER(C.c_int), C.POINTER(LayerWeights), C.POINTER(BranchWeights)]),
    ('mv2d_pack_neck_bytes', C.c_int64, [C.POINTER(C.c_int)]),
    ('mv2d_pack_neck', C.c_int, [C.POINTER(NamedTensor), C.c_int, c_f, C.c_int64, c_f, C.POINTER(PackedEntry), C.c_int,
                                 C.POINTER(C.c_int)]),
    ('mv2d_geom_prep', C.c_int, [c_f, C.c_int, c_f, c_f, c_f]),
    ('mv2d_geom_prep_batch', C.c_int, [c_f, C.c_int, C.c_int, c_f, c_f, c_f]),
    ('mv2d_nchw_to_nhwc', C.c_int, [c_f, c_f, c_f, C.c_int, C.c_int, C.c_int, c_f]),
    ('mv2d_nchw_to_nhwc_split', C.c_int, [c_f, c_f, c_f, c_f, C.c_int, C.c_int, C.c_int, c_f]),
    ('mv2d_nchw_add_to_nhwc', C.c_int, [c_f, c_f, c_f, C.c_int, C.c_int, C.c_int, c_f]),
    ('mv2d_query_embedding', C.c_int, [c_f, C.c_int, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f]),
    ('mv2d_split_tf32', C.c_int, [c_f, c_f, c_f, C.c_longlong, c_f]),
    ('mv2d_gemm_3xtf32', C.c_int, [c_f, c_f, C.c_int, c_f, c_f, C.c_int, c_f, c_f, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_int, c_f]),
    ('mv2d_pe3d_workspace_bytes', C.c_size_t, [C.c_int] * 4),
    ('mv2d_pe3d', C.c_int, [C.POINTER(PeParams), c_f]),
    ('mv2d_roi_align_qg_workspace_bytes', C.c_size_t, [C.c_int]),
    ('mv2d_roi_align_qg', C.c_int, [C.POINTER(QgParams), c_f]),
    ('mv2d_box_corr', C.c_int, [C.POINTER(CorrParams), c_f]),
    ('mv2d_handoff_2d', C.c_int, [c_f, c_f, c_f, c_f, C.c_int, C.c_float, C.c_float, c_f, c_f, c_f]),
    ('mv2d_dn_workspace_bytes', C.c_size_t, [C.c_int, C.c_int]),
    ('mv2d_dn_prepare', C.c_int, [C.POINTER(DnParams), c_f]),
    ('mv2d_decoder_workspace_bytes', C.c_size_t, [C.c_int, C.c_int]),
    ('mv2d_decoder', C.c_int, [C.POINTER(DecoderParams), c_f]),
    ('mv2d_cross_attention_core', C.c_int, [C.POINTER(DecoderParams), C.c_int, c_f, c_f, c_f, c_f]),
    ('mv2d_xa_tile_workspace_bytes', C.c_size_t, [C.c_int] * 4),
    ('mv2d_xa_tile_workspace_bytes_batch', C.c_size_t, [C.c_int] * 5),
    ('mv2d_kv_project', C.c_int, [C.POINTER(KvParams), c_f]),
    ('mv2d_loss_workspace_bytes', C.c_size_t, [C.c_int] * 3),
    ('mv2d_loss', C.c_int, [C.POINTER(LossParams), c_f]),
    ('mv2d_fpn_neck_workspace_bytes', C.c_size_t, [C.c_int] * 3),
    ('mv2d_fpn_neck', C.c_int, [C.POINTER(NeckParams), c_f]),
    ('mv2d_xa_tile_prepare', C.c_int, [C.POINTER(DecoderParams), c_f]),
    ('mv2d_train_param_total', C.c_longlong, [C.c_int]),
    ('mv2d_train_set_tensor_cores', C.c_int, [C.c_int]),
    ('mv2d_train_param_info', C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    ('mv2d_decoder_train_workspace_bytes', C.c_size_t, [C.c_int] * 4),
    ('mv2d_decoder_train_workspace_bytes_p', C.c_size_t, [C.POINTER(TrainParams)]),
    ('mv2d_decoder_train_forward', C.c_int, [C.POINTER(TrainParams), c_f]),
    ('mv2d_train_debug_offset', C.c_longlong, [C.c_int] * 6),
    ('mv2d_decoder_train_backward', C.c_int, [C.POINTER(TrainParams), c_f]),
    ('mv2d_front_train_workspace_bytes', C.c_size_t, [C.c_int] * 4),
    ('mv2d_front_train_forward', C.c_int, [C.POINTER(FrontTrainParams), c_f]),
    ('mv2d_front_train_backward', C.c_int, [C.POINTER(FrontTrainParams), c_f]),
    ('mv2d_adamw_step', C.c_int, [c_f, c_f, c_f, c_f, C.c_longlong, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                  C.c_int, C.c_float, c_f]),
    ('mv2d_gemm', C.c_int, [c_f, C.c_int, c_f, C.c_int, c_f, c_f, C.c_int, C.c_int, C.c_int, C.c_int,
                            C.c_int, c_f]),
    ('mv2d_debug_clock_probe', C.c_int, [C.c_longlong, c_f, c_f]),
    ('mv2d_scene_nms', C.c_int, [c_f, c_f, c_f, c_f, C.c_int, C.c_float, C.c_float, C.c_int, c_f, c_f, c_f, c_f, c_f]),
    ('mv2d_nms_free_decode', C.c_int, [c_f, c_f, C.c_int, C.c_int, C.POINTER(C.c_float), c_f, c_f, c_f,
                                       c_f, c_f]),
]

_lib = None


def load():
    """Load the shared library (once).  Raises if it is missing -- there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
            f'or `make -C mv2d_b200/csrc`.  mv2d_b200 has no CPU fallback.')
    lib = C.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.mv2d_abi_version() != ABI_VERSION:
        raise RuntimeError(f'libmv2d_b200 ABI {lib.mv2d_abi_version()} != binding {ABI_VERSION}')
    for i, st in enumerate(_STRUCTS):
        if lib.mv2d_sizeof(i) != C.sizeof(st):
            raise RuntimeError(f'{st.__name__}: python mirror is {C.sizeof(st)} bytes, '
                               f'library says {lib.mv2d_sizeof(i)}')
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().mv2d_last_error().decode(errors='replace')
        raise RuntimeError(f'{what} failed (rc={rc}): {msg}')


def ptr(t):
    """Device pointer of a contiguous torch tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_contiguous(), 'libmv2d_b200 needs contiguous tensors'
    return t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
