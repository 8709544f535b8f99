"""Inter-sample pipelining of the hot path at batch size 1.

One sample's path is a chain of ~100 dependent kernels: the front end (position-embedding MLPs, RoIAlign + query
generator) fills the GPU for ~0.4 ms, the 6-layer decoder that follows is ~0.65 ms of small latency-bound kernels
that leave most SMs idle.  ``Pipeline`` keeps ``depth`` samples in flight: ``depth`` independent lanes (each a
``HotPath`` with its own buffers, streams and CUDA graphs, all sharing ONE set of packed weights) are fed round-robin,
so the front end of sample i+1 runs under the decoder of sample i.  Every sample is still processed on its own
(the reference asserts batch size 1, mv2d_head.py:106); only the schedule changes, the results are bit-identical
to ``HotPath.forward``.
"""
import torch

from .engine import HotPath


class Pipeline:
    def __init__(self, state_dict, mode='S', device='cuda', depth=2, **kw):
        assert depth >= 1
        first = HotPath(state_dict, mode=mode, device=device, **kw)
        self.lanes = [first] + [HotPath(None, mode=mode, device=device, weights=first.w, **kw)
                                for _ in range(depth - 1)]
        self.depth = depth
        self.device = first.device
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(depth)]
        self.done = [torch.cuda.Event() for _ in range(depth)]
        self._next = 0
        self._host = [None] * depth

    @property
    def L(self):
        return self.lanes[0].L

    def launch_count(self):
        # the library-wide eager counter is shared by all lanes: count it once, graph replays per lane
        return int(self.lanes[0].lib.mv2d_launch_count()) + sum(l.graph_launches for l in self.lanes)

    def submit(self, feat, proposal_list, img_metas, to_host=False):
        """Enqueue one sample on the next lane (ordered after the work already on the caller's stream).
        Returns (ticket, out): ``out`` holds views of that lane's buffers, valid until the lane is used again
        (``depth`` submissions later).  to_host=True also enqueues the D2H copy of cls_scores / bbox_preds into
        the lane's pinned host buffers (out['host_cls'], out['host_box'])."""
        k = self._next % self.depth
        self._next += 1
        s = self.streams[k]
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            out = self.lanes[k].forward(feat, proposal_list, img_metas, use_graph=True)
            if to_host:
                h = self._host[k]
                if h is None or h[0].shape != out['cls_scores'].shape:
                    h = (torch.empty(out['cls_scores'].shape, dtype=torch.float32).pin_memory(),
                         torch.empty(out['bbox_preds'].shape, dtype=torch.float32).pin_memory())
                    self._host[k] = h
                h[0].copy_(out['cls_scores'], non_blocking=True)
                h[1].copy_(out['bbox_preds'], non_blocking=True)
                out['host_cls'], out['host_box'] = h
            self.done[k].record(s)
        return k, out

    def submit_batch(self, feats, proposal_lists, metas_list, to_host=False, bucket=1):
        """Enqueue one BATCH (B samples through one kernel chain, ``HotPath.forward_batch``) on the next lane: the
        GPU-filling front end of batch i+1 runs under the decoder tail of batch i.  Same contract as ``submit``."""
        k = self._next % self.depth
        self._next += 1
        s = self.streams[k]
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            out = self.lanes[k].forward_batch(feats, proposal_lists, metas_list, use_graph=True, bucket=bucket)
            if to_host:
                h = self._host[k]
                if h is None or h[0].shape != out['cls_scores'].shape:
                    h = (torch.empty(out['cls_scores'].shape, dtype=torch.float32).pin_memory(),
                         torch.empty(out['bbox_preds'].shape, dtype=torch.float32).pin_memory())
                    self._host[k] = h
                h[0].copy_(out['cls_scores'], non_blocking=True)
                h[1].copy_(out['bbox_preds'], non_blocking=True)
                out['host_cls'], out['host_box'] = h
            self.done[k].record(s)
        return k, out

    def wait(self, ticket):
        """Make the caller's stream wait for the sample behind ``ticket``."""
        torch.cuda.current_stream().wait_event(self.done[ticket])

    def join(self):
        """Make the caller's stream wait for every lane."""
        cur = torch.cuda.current_stream()
        for s in self.streams:
            cur.wait_stream(s)
