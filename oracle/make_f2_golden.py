"""Golden vectors for next row f2 (detections hand-off), written by the REFERENCE's own code.

detectors/mv2d.py cannot be imported here (it needs mmdet3d / cv2 / the 2D detector), but the three methods of the
hand-off -- process_2d_detections, box_iou, complement_2d_gt (detectors/mv2d.py:60-117) -- are self-contained torch
code: their source is cut out of the reference file with ``ast`` and executed unmodified against a stub ``self`` that
carries train_cfg.  Run in the build container: python -m oracle.make_f2_golden"""
import ast
import os
import textwrap
import types

import numpy as np
import torch

REF = '/root/reference/mmdet3d_plugin/models/detectors/mv2d.py'
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def reference_methods():
    src = open(REF).read()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == 'MV2D')
    ns = dict(torch=torch, np=np)
    for fn in cls.body:
        if isinstance(fn, ast.FunctionDef) and fn.name in ('process_2d_detections', 'box_iou', 'complement_2d_gt', 'process_2d_gt'):
            exec(textwrap.dedent(ast.get_source_segment(src, fn)), ns)
    return ns


def cases():
    """(name, per-class detection arrays of one view, gt boxes [m,4], gt labels [m], min_bbox_size, thr)"""
    rng = np.random.Generator(np.random.PCG64(77))

    def boxes(n, lo=4, hi=300):
        c = rng.uniform(0, 1400, (n, 2)); wh = rng.uniform(lo, hi, (n, 2))
        return np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)

    def dets(ns):
        return [np.concatenate([boxes(n), rng.uniform(0.05, 1, (n, 1)).astype(np.float32)], 1) for n in ns]
    out = []
    g = boxes(9)
    d = dets([3, 0, 5, 2, 0, 0, 4, 1, 0, 6])
    d[2][:3, :4] = g[:3] + rng.uniform(-6, 6, (3, 4)).astype(np.float32)      # three detections sit on ground-truth boxes
    out.append(('mixed', d, g, rng.integers(0, 10, 9), 8.0, 0.4))
    out.append(('no_gt', dets([2, 3, 0, 0, 1, 0, 0, 0, 0, 2]), np.zeros((0, 4), np.float32), np.zeros((0,), np.int64), 8.0, 0.4))
    out.append(('no_det', dets([0] * 10), boxes(5, 2, 40), rng.integers(0, 10, 5), 8.0, 0.4))
    small = boxes(12, 2, 14)
    out.append(('small_boxes', dets([4, 4, 4, 0, 0, 0, 0, 0, 0, 0]), small, rng.integers(0, 10, 12), 8.0, 0.4))
    d2 = dets([6, 6, 0, 0, 0, 0, 0, 0, 0, 0])
    d2[0][:, 2:4] = d2[0][:, 0:2] + rng.uniform(1, 7.9, (6, 2)).astype(np.float32)    # all below the size threshold
    out.append(('tiny_dets', d2, boxes(4), rng.integers(0, 10, 4), 8.0, 0.35))
    return out


def main():
    ns = reference_methods()
    for name, per_cls, gt, gl, min_size, thr in cases():
        stub = types.SimpleNamespace(train_cfg=dict(detection_proposal=dict(min_bbox_size=min_size), complement_2d_gt=thr), test_cfg=None)
        stub.box_iou = ns['box_iou']            # a @staticmethod in the reference: called as self.box_iou(a, b)
        det = ns['process_2d_detections'](stub, [per_cls], 'cpu')[0]
        gts = ns['process_2d_gt'](stub, [torch.from_numpy(gt)], [torch.from_numpy(gl)], 'cpu')[0]
        out = ns['complement_2d_gt'](stub, det, gts, thr=thr)
        np.savez(os.path.join(OUT, f'f2_{name}.npz'), det_in=np.concatenate(
            [np.concatenate([b, np.full((len(b), 1), i, np.float32)], 1) for i, b in enumerate(per_cls)], 0).astype(np.float32),
            gt_boxes=gt, gt_labels=gl, min_size=np.float32(min_size), thr=np.float32(thr), det_filtered=det.numpy(), out=out.numpy())
        print(name, det.shape, gts.shape, '->', tuple(out.shape))


if __name__ == '__main__':
    main()
