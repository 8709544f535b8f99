"""The reference's torch composite on a GPU ("reference-on-GPU" line of BASELINE.md section 3; test / bench
infrastructure like the rest of oracle/, never imported by the product).

Same stage functions as ``mv2d_oracle.mv2d_s_forward`` (each cites the reference lines it restates), placed the way a
torch user would place them: the dense tensor stages -- PE MLPs, RoIAlign, query generator, decoder, branches -- run
on the GPU through torch's own kernels (cuDNN / cuBLAS conv and linear, torchvision's CUDA RoIAlign, which has the
semantics of the mmcv op the reference calls: avg, aligned, adaptive sampling grid); the @torch.no_grad index stage
(box correlation: per-RoI Python loops over tiny tensors in this restatement) stays on the host.
fp32, no TF32 (``torch.backends.*.allow_tf32 = False``), eval mode."""
import time

import torch
import torchvision

from . import mv2d_oracle as O


def to_device(sd, device):
    return {k: v.to(device) for k, v in sd.items()}


@torch.no_grad()
def s_forward(sd_dev, feat_dev, proposal_list, img_metas, cfg=None, timings=None):
    """MV2D-S hot path: (cls_scores, bbox_preds) [L,N,10] on ``feat_dev.device``; proposal_list on the host."""
    cfg = cfg or O.make_cfg('S')
    dev = feat_dev.device
    t0 = time.perf_counter()
    proposal_list = O.guard_empty(proposal_list)
    rois = O.bbox2roi(proposal_list)
    K, E = O.get_box_params(proposal_list, img_metas, cfg['roi_size'])
    num_per_view = [len(p) for p in proposal_list]
    corr, mask = O.box_roi_correlation(rois, num_per_view, img_metas, cfg)        # host: index work
    t1 = time.perf_counter()
    with torch.device(dev):
        pe = O.pe_forward(sd_dev, feat_dev, img_metas, cfg)
        rois_d, K_d, E_d = rois.to(dev), K.to(dev), E.to(dev)
        roi_feat = torchvision.ops.roi_align(feat_dev, rois_d, cfg['roi_size'], 1.0 / cfg['stride'], sampling_ratio=0, aligned=True)
        roi_pe = torchvision.ops.roi_align(pe, rois_d, cfg['roi_size'], 1.0 / cfg['stride'], sampling_ratio=0, aligned=True)
        ifeat = O.process_intrins_feat(rois_d, K_d, cfg['intrins_feat_scale'])
        ref, _ = O.query_generator(sd_dev, roi_feat, K_d, E_d, ifeat, cfg)
        corr_d, mask_d = corr.to(dev), mask.to(dev)
        N, M = corr.shape
        C = feat_dev.shape[1]
        mem = roi_feat[corr_d].permute(1, 3, 4, 0, 2).reshape(M * 49, N, C)
        pos = roi_pe[corr_d].permute(1, 3, 4, 0, 2).reshape(M * 49, N, C)
        kpm = (~mask_d)[:, :, None].expand(N, M, 49).reshape(N, M * 49)
        qpos = O.query_embed(sd_dev, ref[:, None])
        outs = O.decoder(sd_dev, qpos.permute(1, 0, 2), mem, pos, cfg, key_padding_mask=kpm).transpose(1, 2)
        cls, box = O.branches(sd_dev, outs, ref[:, None], cfg)
    if timings is not None:
        timings['index_host_s'] = t1 - t0
    return cls.flatten(1, 2), box.flatten(1, 2)


def time_s(sd, samples, device, warmup=5, iters=20):
    """samples/s of ``s_forward`` on ``device``: host index stage by wall clock + dense stages by CUDA events (the two
    do not overlap in a plain torch script), after ``warmup`` untimed runs."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    sd_dev = to_device(sd, device)
    feats = [s[0].to(device) for s in samples]
    tot_ev, tot_host, tim = 0.0, 0.0, {}
    for i in range(warmup + iters):
        f, (_, boxes, metas) = feats[i % len(samples)], samples[i % len(samples)]
        torch.cuda.synchronize(device)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        s_forward(sd_dev, f, boxes, metas, timings=tim)
        b.record()
        torch.cuda.synchronize(device)
        wall = time.perf_counter() - t0
        if i >= warmup:
            tot_ev += wall
            tot_host += tim['index_host_s']
    ms = 1e3 * tot_ev / iters
    return dict(value=1e3 / ms, unit='samples/s', ms_per_sample=ms, index_host_ms=1e3 * tot_host / iters, warmup=warmup, iters=iters,
                kind='torch composite on the same GPU (oracle stage functions: cuDNN / cuBLAS fp32, torchvision RoIAlign; '
                     'box correlation on the host), eager, one sample at a time',
                torch=torch.__version__)
