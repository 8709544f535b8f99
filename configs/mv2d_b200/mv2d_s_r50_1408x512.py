# MV2D-S (single frame, 6 views): RoI-token keys of each query's own RoI + top-1 epipolar match per view.
_base_ = ['./_roi_head_common.py']
model = dict(type='MV2D',
             roi_head=dict(type='MV2DSHead', use_denoise=False,
                           box_correlation=dict(correlation_mode='topk_matched:1:0.0:0.0')),
             test_cfg=dict(rcnn=dict(score_thr=0.0, max_per_scene=300)))
