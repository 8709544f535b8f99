"""BASELINE config 5: attention-only microbench.  One decoder cross-attention layer (absorbed query
projection GEMM + sparse attention core + output GEMM), N in {100,300,900} queries x V in {6,12}
views, against the HBM roofline with SURVEY.md 8d's algorithmic bytes.  Prints one JSON line per
point.  Replicas only across GPUs (no collective on this path), so multi-GPU = N x these numbers."""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mv2d_b200 import synth
from mv2d_b200.engine import HotPath

peaks = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json'))) \
    if os.path.exists(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')) else dict(hbm_gbs=6650.0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
sd1 = synth.make_state_dict(0, num_layers=1)


def time_fn(fn, reps=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts)


for mode, V in (('S', 6), ('T', 12)):
    for N in (100, 300, 900):
        per_view = N // V
        eng = HotPath(sd1, mode=mode)          # 1 decoder layer: the timed stage is exactly one layer + branches
        feat, boxes, metas = synth.make_sample(7, V, per_view)
        out = eng.forward(feat.cuda(), boxes, metas)
        torch.cuda.synchronize()
        n = out['N']
        qg = {k: out[k] for k in ('query_pos', 'ref', 'tok_feat', 'tok_kin')}
        C = 256
        if mode == 'S':
            corr = dict(match=out['match'], match_cnt=out['match_cnt'], max_match=out['max_match'])
            kin_rows, mem_rows = out['tok_kin'].view(-1, 256), out['tok_feat'].view(-1, 256)
            mc = float(out['match_cnt'].float().mean())
            n_k, mask_b, gathered = 49 * n, n * mc * 49, n * mc * 49
        else:
            corr = dict(keymask=out['keymask'], mask_words=out['mask_words'], key_list=out['key_list'], key_cnt=out['key_cnt'])
            mem_rows = out['feat_nhwc'].view(-1, 256)
            kin_rows = eng._buf['kin'][:mem_rows.numel()].view(-1, 256)
            km = out['keymask'].cpu().numpy().view(np.uint32)
            n_k = int(np.unpackbits(np.bitwise_or.reduce(km, axis=0).view(np.uint8)).sum())
            mask_b, gathered = n * n_k, float(out['key_cnt'].sum())
        alg = 4 * (2 * n_k * C + 2 * n * C + 4 * C * C + 4 * C) + mask_b
        run = lambda: eng.decoder(qg, corr, kin_rows, mem_rows, n, vel_dt=0.5 if mode == 'T' else 0.0)
        # replayed from a CUDA graph: no CPU launch gaps inside the timed region (the eager chain is launch bound)
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            run()
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                run()
        torch.cuda.synchronize()
        t = time_fn(graph.replay)
        # the sparse cross-attention kernel(s) alone (mv2d_cross_attention_core: xa_roi, or xt_attn_mma + xt_merge), L2 flushed
        core = {}
        try:
            if mode == 'S':
                q_core = torch.randn(n, 2048, device='cuda')
                fn = lambda: eng.cross_attention_core(qg, corr, kin_rows, mem_rows, n, q_core, layer=0)
                core_bytes = float(out['match_cnt'].sum()) * 49 * 2 * 1024        # 100 KB per (query, RoI) unit
            else:
                h_, w_ = out['feat_nhwc'].shape[1:3]
                kv = (eng._buf['kp'].view(eng.L, -1, 256)[:, :mem_rows.shape[0]], eng._buf['vp'].view(eng.L, -1, 256)[:, :mem_rows.shape[0]])
                corr_t = {k: out.get(k) for k in ('keymask', 'mask_words', 'key_list', 'key_cnt', 'xa_prepared_for', 'row_tile_live')}
                q_core = torch.randn(n, 256, device='cuda')
                fn = lambda: eng.cross_attention_core(qg, corr_t, kin_rows, mem_rows, n, q_core, layer=0, kv=kv, grid=(h_, w_))
                core_bytes = 4 * 2 * n_k * C                                        # projected K and V rows of the union of keys
            tc = time_fn(fn)
            core = dict(core_us=tc, core_MB=core_bytes / 1e6, core_GBs=core_bytes / tc / 1e3, core_frac_hbm=core_bytes / tc / 1e3 / peaks['hbm_gbs'])
        except Exception as e:
            core = dict(core_error=f'{type(e).__name__}: {e}'[:200])
        # rows actually gathered by the kernel (each query streams its own key rows, 2 KB per key)
        moved = gathered * 2 * C * 4
        print(json.dumps(dict(mode=mode, V=V, N=n, unique_keys=n_k, keys_gathered=gathered, us_layer=t,
                              algorithmic_MB=alg / 1e6, achieved_GBs=alg / t / 1e3, frac_hbm=alg / t / 1e3 / peaks['hbm_gbs'],
                              gathered_MB=moved / 1e6, gathered_GBs=moved / t / 1e3, **core,
                              note='one full decoder layer (self-attn, cross-attn incl. its K/V projection for the two-frame head, FFN) + branches, '
                                   'replayed from a CUDA graph, charged to the attention bytes')))
