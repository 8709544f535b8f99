# Hot-path model settings shared by the single-frame and two-frame heads.  Field names and values
# are the ones the reference experiment configs feed to the registry
# (reference: configs/mv2d/exp/mv2d_r50_frcnn_*_1408x512_ep*.py); this file only describes the
# roi_head (the path libmv2d_b200 accelerates), not the 2D detector, data pipeline or schedule.
point_cloud_range = [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]
post_range = [-61.2, -61.2, -10.0, 61.2, 61.2, 10.0]
roi_size = 7
roi_strides = [16]

decoder_layer = dict(
    type='PETRTransformerDecoderLayer',
    attn_cfgs=[dict(type='FlattenMHSelfAttention', embed_dims=256, num_heads=8, dropout=0.1),
               dict(type='PETRMultiheadAttention', embed_dims=256, num_heads=8, dropout=0.1)],
    feedforward_channels=2048, ffn_dropout=0.1, with_cp=False,
    operation_order=('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm'))

roi_head = dict(
    pc_range=point_cloud_range,
    force_fp32=True,
    bbox_roi_extractor=dict(type='SingleRoIExtractor',
                            roi_layer=dict(type='RoIAlign', output_size=roi_size, sampling_ratio=-1),
                            featmap_strides=roi_strides, out_channels=512),
    bbox_head=dict(
        type='CrossAttentionBoxHead', num_classes=10, pc_range=point_cloud_range,
        transformer=dict(type='MV2DTransformer',
                         decoder=dict(type='PETRTransformerDecoder', return_intermediate=True, num_layers=6,
                                      transformerlayers=decoder_layer)),
        bbox_coder=dict(type='NMSFreeCoder', post_center_range=post_range, pc_range=point_cloud_range,
                        max_num=300, num_classes=10),
        code_weights=[1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.5, 1.5, 2.0, 2.0],
        loss_cls=dict(type='FocalLoss', use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=2.0),
        loss_bbox=dict(type='L1Loss', loss_weight=0.25)),
    query_generator=dict(with_avg_pool=True, num_shared_convs=1, num_shared_fcs=1, in_channels=256,
                         fc_out_channels=1024, roi_feat_size=roi_size,
                         extra_encoding=dict(num_layers=2, feat_channels=[512, 256],
                                             features=[dict(type='intrinsic', in_channels=16)])),
    pe=dict(positional_encoding=dict(type='SinePositionalEncoding3D', num_feats=128, normalize=True),
            strides=roi_strides, position_range=post_range, depth_num=64, with_fpe=True),
)
test_rcnn = dict(score_thr=0.0, nms=dict(nms_thr=1.0, use_rotate_nms=True), max_per_scene=300)

# training-side settings of the roi_head (reference: configs/mv2d/exp/mv2d_r50_frcnn_single_frame_roi_1408x512_ep72.py:130-141):
# consumed by the loss row (mv2d_loss: Hungarian cost weights, per-layer loss weights)
model = dict(
    # the MV2D neck: a one-level FPN on the 2D detector's P4 (reference configs/mv2d/exp/*.py:32-39)
    neck=dict(type='FPN', in_channels=[256, 256, 256, 256, 256], out_channels=256, start_level=2, end_level=2, num_outs=1),
    train_cfg=dict(
        rcnn=dict(
            stage_loss_weights=[0.1, 0.1, 0.1, 0.1, 0.1, 0.1],
            assigner=dict(type='HungarianAssigner3D',
                          cls_cost=dict(type='FocalLossCost', weight=2.0),
                          reg_cost=dict(type='BBox3DL1Cost', weight=0.25),
                          iou_cost=dict(type='IoUCost', weight=0.0),
                          pc_range=point_cloud_range),
            sampler_cfg=dict(type='PseudoSampler'),
            pos_weight=-1,
            debug=False)))
