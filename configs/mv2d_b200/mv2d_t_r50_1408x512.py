# MV2D-T (two frames, 12 views): dense feature-map keys restricted per query to the cells of its own
# and its top-20 epipolar-matched RoIs (expanded by 2 strides); velocities divided by the frame gap.
_base_ = ['./_roi_head_common.py']
model = dict(type='MV2DT',
             roi_head=dict(type='MV2DTHead', use_denoise=True, neg_bbox_loss=True, denoise_noise_scale=1.25,
                           denoise_split=0.6,
                           box_correlation=dict(expand_stride=2, correlation_mode='topk_matched:20:0.0:0.0')),
             test_cfg=dict(rcnn=dict(score_thr=0.0, max_per_scene=300)))
