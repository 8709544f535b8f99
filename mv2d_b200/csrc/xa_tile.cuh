// Two-frame head (MV2D-T) cross-attention, KEY-STATIONARY form.  Included by decoder.cu.
//   reference: roi_heads/mv2d_t_head.py:67-88 (per-query key mask over the [V,h,w] cells),
//              utils/petr_transformer.py:426-513 (PETRMultiheadAttention -> nn.MultiheadAttention)
//
// A query of the two-frame head attends to ~2 000 of the V*h*w feature cells (its own box plus the boxes of up
// to 20 matched RoIs in the other views); the union over the ~300 queries covers ~3/4 of all cells.  The
// query-stationary kernel (cross_attn_kernel, mode 1) streams every query's key rows separately: 300 x 1 921 x
// 2 KB = 1.2 GB from L2 per layer for 60 MB of distinct data.  Here the loop nest is turned around:
//   * the key / value projections K_l = (mem + pos) Wk_l^T, V_l = mem Wv_l^T are computed ONCE per layer for all
//     cells by the 3xTF32 tcgen05 GEMM (mv2d_kv_project);
//   * the cells are cut into 8x8 tiles (64 keys = 64 KB of K + 64 KB of V, fetched with 16 bulk copies into
//     shared memory); one CTA per tile walks the queries that have at least one key in it (list built once per
//     sample by xt_prep_kernel from the bit-packed mask) -- a warp per query, only the query's own keys
//     (64-bit tile mask) are touched -- and leaves an un-normalised (acc[256], m[8], l[8]) record per
//     (query, tile);
//   * xt_merge_kernel folds a query's records in ascending tile order (fixed order => bitwise reproducible).
// Every K/V row is read from HBM/L2 once per layer; the per-query re-reads hit shared memory.
#pragma once
#include "common.cuh"

namespace mv2d {

#define XT_TS 8                        // tile side, cells
#define XT_KEYS (XT_TS * XT_TS)
#define XT_REC 272                     // floats per record: acc[256], m[8], l[8]
#define XT_THREADS 512
#define XT_MERGE_MAXT 2048             // tiles per SAMPLE the lists are sized for (V*ceil(h/8)*ceil(w/8) <= 2048)
#define XT_ORDER_MAXT 8192             // tiles of a whole batch the heaviest-first order is sized for
#define XT_SMEM_BYTES (2 * XT_KEYS * MV2D_C * 4 + (XT_THREADS / 32) * XT_KEYS * 8 * 4 + (XT_THREADS / 32) * 80 + 64)

// Batch: B samples, each with Np query rows (row b*Np + i), V views and tiles_ps = V*tiles_y*tiles_x tiles; the tile
// id t = b*tiles_ps + local tile, N = B*Np, ntiles = B*tiles_ps.  A tile only ever meets the queries of its own sample.
struct XtGeom {
    int N, V, h, w, tiles_x, tiles_y, ntiles;
    int B, Np, tiles_ps;
};

struct XtPrepArgs {
    XtGeom g;
    const uint32_t* keymask; int mask_words;
    int* tile_cnt;                     // [ntiles]
    uint16_t* tile_q;                  // [ntiles, N] queries with a key in the tile, ascending
    unsigned long long* tile_mask;     // [ntiles, N] their 64-bit key masks (bit r*8+c = cell (ty*8+r, tx*8+c))
    short* slot_of;                    // [N, ntiles] position of the query in the tile's list, -1 = none
    int* tile_work;                    // [ntiles] sum over the tile's queries of their key counts
};

// grid = ntiles, 256 threads
__global__ void __launch_bounds__(256) xt_prep_kernel(XtPrepArgs a) {
    pdl_wait();
    pdl_trigger();
    const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sb = t / a.g.tiles_ps, tl = t - sb * a.g.tiles_ps, q0 = sb * a.g.Np;     // sample, local tile, first query row
    const int tx = tl % a.g.tiles_x, ty = (tl / a.g.tiles_x) % a.g.tiles_y, v = tl / (a.g.tiles_x * a.g.tiles_y);
    const int ncols = min(XT_TS, a.g.w - tx * XT_TS);
    __shared__ int wsum[8];
    __shared__ int work_s;
    if (tid == 0) work_s = 0;
    int running = 0, work = 0;
    for (int base = 0; base < a.g.Np; base += 256) {
        const int nl = base + tid, n = q0 + nl;
        unsigned long long m64 = 0ull;
        if (nl < a.g.Np) {
            const uint32_t* km = a.keymask + (long long)n * a.mask_words;
#pragma unroll
            for (int r = 0; r < XT_TS; ++r) {
                const int y = ty * XT_TS + r;
                if (y < a.g.h) {
                    const int c = (v * a.g.h + y) * a.g.w + tx * XT_TS;
                    const int word = c >> 5, sh = c & 31;
                    unsigned long long two = km[word];
                    if (sh + ncols > 32 && word + 1 < a.mask_words) two |= (unsigned long long)km[word + 1] << 32;
                    const unsigned bits = (unsigned)(two >> sh) & ((1u << ncols) - 1u);
                    m64 |= (unsigned long long)bits << (r * 8);
                }
            }
        }
        const bool active = m64 != 0ull;
        work += __popcll(m64);
        const unsigned bal = __ballot_sync(0xffffffffu, active);
        if (lane == 0) wsum[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { const int c = wsum[i]; if (i < warp) before += c; total += c; }
        const int pos = running + before + __popc(bal & ((1u << lane) - 1u));
        if (nl < a.g.Np) {
            if (active) {
                a.tile_q[(long long)t * a.g.Np + pos] = (uint16_t)n;
                a.tile_mask[(long long)t * a.g.Np + pos] = m64;
            }
            a.slot_of[(long long)n * a.g.tiles_ps + tl] = active ? (short)pos : (short)-1;
        }
        running += total;
        __syncthreads();
    }
    work = __reduce_add_sync(0xffffffffu, work);
    if (lane == 0) atomicAdd(&work_s, work);
    __syncthreads();
    if (tid == 0) { a.tile_cnt[t] = running; a.tile_work[t] = work_s; }
}

// ---- per-query record lists and a heaviest-first tile order, built once per sample
struct XtListArgs {
    XtGeom g;
    const short* slot_of;              // [N, ntiles]
    const int* tile_work;              // [ntiles] keys x queries of the tile
    int* qlist;                        // [N, ntiles] record ids (tile * N + slot) of the query, ascending tile order
    int* qcnt;                         // [N]
    int* order;                        // [ntiles] tiles sorted by work, heaviest first (ties: lower id first)
};

// grid = N + ceil(ntiles / 256), 256 threads.  Blocks 0..N-1: ordered compaction of slot_of[n, :]; the others: the tile
// order (each ranks 256 tiles of the batch against all of them).
__global__ void __launch_bounds__(256) xt_list_kernel(XtListArgs a) {
    pdl_wait();
    pdl_trigger();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if ((int)blockIdx.x >= a.g.N) {
        // rank by counting out of shared memory, once per batch
        __shared__ int work_s[XT_ORDER_MAXT];
        const int nall = a.g.ntiles;
        for (int t = tid; t < nall; t += 256) work_s[t] = a.tile_work[t];
        __syncthreads();
        const int t = ((int)blockIdx.x - a.g.N) * 256 + tid;
        if (t < nall) {
            const int w = work_s[t];
            int rank = 0;
#pragma unroll 8
            for (int u = 0; u < nall; ++u) {
                const int x = work_s[u];
                rank += (x > w) || (x == w && u < t);
            }
            a.order[rank] = t;
        }
        return;
    }
    const int nt = a.g.tiles_ps;                 // a query only meets the tiles of its own sample
    const int n = blockIdx.x;
    const int tile0 = (n / a.g.Np) * a.g.tiles_ps;
    __shared__ int chunk_cnt[64];      // ntiles <= 2048 => <= 64 chunks of 32
    const int nchunks = (nt + 31) >> 5;
    int sv[8];
    unsigned bal[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int c = warp + u * 8;
        sv[u] = -1; bal[u] = 0u;
        if (c < nchunks) {
            const int t = c * 32 + lane;
            sv[u] = t < nt ? (int)a.slot_of[(long long)n * nt + t] : -1;
            bal[u] = __ballot_sync(0xffffffffu, sv[u] >= 0);
            if (lane == 0) chunk_cnt[c] = __popc(bal[u]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int c = warp + u * 8;
        if (c < nchunks) {
            int base = 0;
            for (int i = 0; i < c; ++i) base += chunk_cnt[i];
            if (sv[u] >= 0) a.qlist[(long long)n * nt + base + __popc(bal[u] & ((1u << lane) - 1u))] = (tile0 + c * 32 + lane) * a.g.Np + sv[u];
        }
    }
    if (tid == 0) {
        int tot = 0;
        for (int i = 0; i < nchunks; ++i) tot += chunk_cnt[i];
        a.qcnt[n] = tot;
    }
}

// 128-row tiles of the K/V projection GEMM that hold at least one key of some query: grid = tiles, 128 threads
// Batch (rows_ps = cells of one sample, a multiple of 128; Np query rows per sample): tile t belongs to ONE sample.
__global__ void __launch_bounds__(128) xt_rowlive_kernel(const uint32_t* __restrict__ keymask, int mask_words, int N, int num_rows,
                                                         uint8_t* __restrict__ live, int rows_ps, int Np) {
    pdl_wait();
    pdl_trigger();
    const int t = blockIdx.x;
    int w0 = t * 4, nbeg = 0, nend = N;
    if (rows_ps > 0) {
        const int sb = (t * 128) / rows_ps;
        w0 = ((t * 128) - sb * rows_ps) >> 5;
        nbeg = sb * Np; nend = nbeg + Np;
    }
    int any = 0;
    for (int n = nbeg + threadIdx.x; n < nend; n += 128) {
        const uint32_t* km = keymask + (long long)n * mask_words;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (w0 + j < mask_words) any |= km[w0 + j] != 0u;
    }
    any = __syncthreads_or(any);
    if (threadIdx.x == 0) live[t] = any ? 1 : 0;
}

struct XtAttnArgs {
    XtGeom g;
    const float* q;                    // [N,256] projected queries, 1/sqrt(32) folded in
    const float* kp; const float* vp;  // [V*h*w,256] projected keys / values of this layer
    const int* tile_cnt; const uint16_t* tile_q; const unsigned long long* tile_mask;
    const int* order;                  // [ntiles] heaviest tile first
    float* rec;                        // [ntiles*N, XT_REC]
    int qsplit;                        // gridDim.y: the tile's query list is dealt round-robin to this many CTAs
    const uint8_t* row_live;           // nullable [ceil(rows/128)]: 0 = mv2d_kv_project skipped this 128-row tile (its kp / vp
                                       // rows are stale): the tensor-core kernel reads whole tiles, so it zero-fills those cells
};

// transpose-reduce 4 per-lane partials over the 4 lanes of a head: 3 shuffles instead of 8.
// result: lane holds in v[0] the full sum of index (lane & 3).
__device__ __forceinline__ void reduce4_in4(float (&v)[4], int lane) {
    const bool up2 = lane & 2, up1 = lane & 1;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = up2 ? v[i] : v[i + 2], keep = up2 ? v[i + 2] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    const float send = up1 ? v[0] : v[1], keep = up1 ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
}

__device__ __forceinline__ uint32_t xt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void xt_wait(uint64_t* bar) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(xt_smem_u32(bar)) : "memory");
}

// grid = (ntiles, qsplit), XT_THREADS threads, XT_SMEM_BYTES dynamic shared memory.
// Lane layout: lane = 4*head + part; the lane owns two 4-channel chunks of its head's 32 channels of q, of every K / V
// row and of the accumulator (two conflict-free LDS.128 per row).  A logit is the sum of 4 lanes' partial dots; four keys are reduced
// together with a 3-shuffle transpose, which leaves the logit of (head, key 4*grp + part) in lane 4*head + part -- the
// softmax statistics of a head therefore live in the registers of its 4 lanes.
__global__ void __launch_bounds__(XT_THREADS, 1) xt_attn_kernel(XtAttnArgs a) {
    pdl_wait();
    pdl_trigger();
    const int t = a.order[blockIdx.x];
    const int cnt = a.tile_cnt[t];
    constexpr int NW = XT_THREADS / 32;
    if ((int)blockIdx.y * NW >= cnt) return;
    extern __shared__ __align__(128) unsigned char xt_smem[];
    float* Ks = reinterpret_cast<float*>(xt_smem);
    float* Vs = Ks + XT_KEYS * MV2D_C;
    float* scb = Vs + XT_KEYS * MV2D_C;                          // [NW][64 slots][8 heads] probabilities
    unsigned char* klb = reinterpret_cast<unsigned char*>(scb + NW * XT_KEYS * 8);   // [NW][80] key ids of the query
    uint64_t* bar = reinterpret_cast<uint64_t*>(klb + NW * 80);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sb = t / a.g.tiles_ps, tl = t - sb * a.g.tiles_ps;
    const int tx = tl % a.g.tiles_x, ty = (tl / a.g.tiles_x) % a.g.tiles_y, v = sb * a.g.V + tl / (a.g.tiles_x * a.g.tiles_y);   // global view
    const int ncols = min(XT_TS, a.g.w - tx * XT_TS), nrows = min(XT_TS, a.g.h - ty * XT_TS);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(xt_smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t row_bytes = (uint32_t)ncols * MV2D_C * 4;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(xt_smem_u32(bar)), "r"(2u * nrows * row_bytes) : "memory");
        for (int r = 0; r < nrows; ++r) {
            const long long off = ((long long)(v * a.g.h + ty * XT_TS + r) * a.g.w + tx * XT_TS) * MV2D_C;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(xt_smem_u32(Ks + r * XT_TS * MV2D_C)), "l"(a.kp + off), "r"(row_bytes), "r"(xt_smem_u32(bar)) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(xt_smem_u32(Vs + r * XT_TS * MV2D_C)), "l"(a.vp + off), "r"(row_bytes), "r"(xt_smem_u32(bar)) : "memory");
        }
    }
    __syncthreads();            // barrier initialised before anyone polls it
    float* sc = scb + warp * XT_KEYS * 8;
    unsigned char* kl = klb + warp * 80;
    const uint32_t* kl4 = reinterpret_cast<const uint32_t*>(kl);
    bool waited = false;
    // the lane's two 16-byte chunks of its head's 128-byte slice: chunks (part, part+4), swapped for odd heads so that
    // the 8 lanes of a quarter warp cover all 32 banks in each of the two LDS.128
    const int hd = lane >> 2, part = lane & 3;
    const int offA = hd * 32 + ((part + 4 * (hd & 1)) & 7) * 4, offB = hd * 32 + (((part + 4 * (hd & 1)) & 7) ^ 4) * 4;
    const int step = a.qsplit * NW;
    int i = blockIdx.y * NW + warp;
    // software pipeline: the next query's list entry and q slice are in flight while this one is processed
    int n_nx = 0; unsigned long long m_nx = 0ull; float4 qa_nx = make_float4(0.f, 0.f, 0.f, 0.f), qb_nx = qa_nx;
    auto fetch = [&](int ii) {
        n_nx = a.tile_q[(long long)t * a.g.Np + ii];
        m_nx = a.tile_mask[(long long)t * a.g.Np + ii];
        qa_nx = __ldg(reinterpret_cast<const float4*>(a.q + (long long)n_nx * MV2D_C + offA));
        qb_nx = __ldg(reinterpret_cast<const float4*>(a.q + (long long)n_nx * MV2D_C + offB));
    };
    if (i < cnt) fetch(i);
    for (; i < cnt; i += step) {
        const unsigned long long m64 = m_nx;
        const float4 qa = qa_nx, qb = qb_nx;
        const unsigned mlo = (unsigned)m64, mhi = (unsigned)(m64 >> 32);
        const int nlo = __popc(mlo), nk = nlo + __popc(mhi);
        if (i + step < cnt) fetch(i + step);
        // ---- key ids of the query, ascending, padded with the first key up to the next multiple of 8
        __syncwarp();           // the previous query's loops are done with kl / sc
        {
            const unsigned lt = (1u << lane) - 1u;
            if ((mlo >> lane) & 1u) kl[__popc(mlo & lt)] = (unsigned char)lane;
            if ((mhi >> lane) & 1u) kl[nlo + __popc(mhi & lt)] = (unsigned char)(32 + lane);
            if (lane < 8) kl[nk + lane] = (unsigned char)(mlo ? __ffs(mlo) - 1 : 31 + __ffs(mhi));
        }
        __syncwarp();
        if (!waited) { xt_wait(bar); waited = true; }
        // ---- logits, 4 keys per step
        float sv[16];
#pragma unroll
        for (int grp = 0; grp < 16; ++grp) {
            sv[grp] = -INFINITY;
            if (grp * 4 < nk) {                             // warp-uniform
                const uint32_t kw = kl4[grp];
                float p[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float* row = Ks + ((kw >> (8 * j)) & 0xffu) * MV2D_C;
                    const float4 ka = *reinterpret_cast<const float4*>(row + offA);
                    const float4 kb = *reinterpret_cast<const float4*>(row + offB);
                    float x = qa.x * ka.x;
                    x = fmaf(qa.y, ka.y, x); x = fmaf(qa.z, ka.z, x); x = fmaf(qa.w, ka.w, x);
                    x = fmaf(qb.x, kb.x, x); x = fmaf(qb.y, kb.y, x); x = fmaf(qb.z, kb.z, x); x = fmaf(qb.w, kb.w, x);
                    p[j] = x;
                }
                reduce4_in4(p, lane);
                if (grp * 4 + part < nk) sv[grp] = p[0];
            }
        }
        // ---- softmax statistics in registers: over the lane's 16 values, then over the head's 4 lanes
        float mx = sv[0];
#pragma unroll
        for (int grp = 1; grp < 16; ++grp) mx = fmaxf(mx, sv[grp]);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        float l = 0.f;
#pragma unroll
        for (int grp = 0; grp < 16; ++grp) {
            if (grp * 4 < nk) {
                const float e = __expf(sv[grp] - mx);       // exp(-inf) = 0 for the padding slots
                l += e;
                sc[(grp * 4 + part) * 8 + hd] = e;
            }
        }
        l += __shfl_xor_sync(0xffffffffu, l, 1);
        l += __shfl_xor_sync(0xffffffffu, l, 2);
        float* r = a.rec + ((long long)t * a.g.Np + i) * XT_REC;
        if (part == 0) { r[256 + hd] = mx; r[264 + hd] = l; }
        __syncwarp();
        // ---- acc = sum_k p_k * V_k over the query's keys, 4 keys per step (padding slots carry p = 0)
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
        for (int s = 0; s < nk; s += 4) {
            const uint32_t kw = kl4[s >> 2];
            float4 va[4], vb[4];
            float w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float* row = Vs + ((kw >> (8 * j)) & 0xffu) * MV2D_C;
                va[j] = *reinterpret_cast<const float4*>(row + offA);
                vb[j] = *reinterpret_cast<const float4*>(row + offB);
                w[j] = sc[(s + j) * 8 + hd];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                a0.x = fmaf(w[j], va[j].x, a0.x); a0.y = fmaf(w[j], va[j].y, a0.y);
                a0.z = fmaf(w[j], va[j].z, a0.z); a0.w = fmaf(w[j], va[j].w, a0.w);
                a1.x = fmaf(w[j], vb[j].x, a1.x); a1.y = fmaf(w[j], vb[j].y, a1.y);
                a1.z = fmaf(w[j], vb[j].z, a1.z); a1.w = fmaf(w[j], vb[j].w, a1.w);
            }
        }
        *reinterpret_cast<float4*>(r + offA) = a0;
        *reinterpret_cast<float4*>(r + offB) = a1;
    }
    // a CTA must not exit while its bulk copies are in flight
    if (!waited) xt_wait(bar);
}

// ---------------------------------------------------------------------------------------------------------------
// Tensor-core form of the per-tile attention (default; MV2D_XT_MMA=0 keeps xt_attn_kernel).
// The FFMA kernel above re-reads a K and a V row (2 KB) from shared memory for every (query, key) pair: it is bound by
// shared-memory bandwidth.  Here the tile's queries are taken 16 at a time as the M rows of warp-level TF32 MMAs
// (mma.sync.m16n8k8), one warp per HEAD:
//     S_h [16 x 64] = Q_h [16 x 32] . K_h^T          8 key blocks x 4 channel steps
//     O_h [16 x 32] = P_h [16 x 64] . V_h            8 key steps x 4 channel blocks,  P = exp(S - rowmax) on the query's keys
// so a K / V element is read from shared memory once per 16 queries (and only its head's 128-byte slice per warp).
// Every product is error-compensated 3xTF32 (x = hi + lo, both TF32: hi*hi + lo*hi + hi*lo accumulated in fp32), the
// splits made in registers on the fly, so the result is fp32-grade like the FFMA kernel's.
// The C fragment of S (columns 2t, 2t+1 of a key block) is reused as the A fragment of the P.V product by reading V's
// rows in the matching order (MMA k index t <-> key 2t, t+4 <-> key 2t+1): no shuffles between the two products.
// Keys outside a query's 64-bit mask get p = 0; cells outside the feature map (edge tiles) are zero-filled.
// K / V rows sit at a pitch of 260 floats, which makes both B-fragment access patterns bank-conflict free.
// Record format, tile order and the merge kernel are unchanged.
#define XTM_THREADS 512                    // 16 warps: warp = (head, row-block parity)
#define XTM_PITCH 260
#define XTM_SMEM_BYTES (2 * XT_KEYS * XTM_PITCH * 4 + 64)       // K rows, V rows, two mbarriers

// x = hi + lo, both TF32, at 4 instructions: hi = x rounded to TF32 by adding half an ulp to the bit pattern and masking
// (round-half-away, no inf / nan handling: the operands are finite activations), lo = x - hi (exact) with the same half ulp
// added -- the tensor core reads the upper 19 bits of a .tf32 operand, so the mask of lo is the hardware's.
// Same accuracy class as two cvt.rna.tf32.f32 (hi*hi + lo*hi + hi*lo: ~2^-21 relative per term); cvt.rna.tf32.f32 is an
// ~7-instruction sequence on sm_100a and two of them per operand element were 80 % of the instruction stream of both
// mma.sync attention kernels (ncu source view): 37 -> 31 us (self-attention, 8 x 300), 97 -> 81 us (two-frame tile attention).
__device__ __forceinline__ void split_tf32_reg(float x, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi)) + 0x1000u;
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// grid = (ntiles, qsplit), XTM_THREADS threads (warp = head + 8 * row-block parity), XTM_SMEM_BYTES dynamic shared memory
__global__ void __launch_bounds__(XTM_THREADS, 1) xt_attn_mma_kernel(XtAttnArgs a) {
    pdl_wait();
    pdl_trigger();
    const int t = a.order[blockIdx.x];
    const int cnt = a.tile_cnt[t];
    const int nblk = (cnt + 15) >> 4;
    if ((int)blockIdx.y * 2 >= nblk) return;
    extern __shared__ __align__(128) unsigned char xt_smem[];
    float* Ks = reinterpret_cast<float*>(xt_smem);
    float* Vs = Ks + XT_KEYS * XTM_PITCH;
    uint64_t* bar = reinterpret_cast<uint64_t*>(Vs + XT_KEYS * XTM_PITCH);
    const int tid = threadIdx.x, hd = (tid >> 5) & 7, par = tid >> 8, lane = tid & 31, g = lane >> 2, tg = lane & 3;
    const int sb = t / a.g.tiles_ps, tl = t - sb * a.g.tiles_ps;
    const int tx = tl % a.g.tiles_x, ty = (tl / a.g.tiles_x) % a.g.tiles_y, v = sb * a.g.V + tl / (a.g.tiles_x * a.g.tiles_y);
    const int ncols = min(XT_TS, a.g.w - tx * XT_TS), nrows = min(XT_TS, a.g.h - ty * XT_TS);
    // one 1 KB bulk copy per cell and operand, spread over 64 threads.  Cells outside the map (edge tiles) and cells whose
    // 128-row tile the projection skipped take no copy and are zero-filled: p = 0 times stale data must stay 0.
    __shared__ unsigned long long use_mask;
    bool use = false;
    long long off = 0;
    if (tid < XT_KEYS) {
        const int r = tid >> 3, c = tid & 7;
        if (r < nrows && c < ncols) {
            const long long cell = (long long)(v * a.g.h + ty * XT_TS + r) * a.g.w + tx * XT_TS + c;
            off = cell * MV2D_C;
            use = a.row_live == nullptr || a.row_live[cell >> 7] != 0;
        }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, use);
    if (tid == 0) use_mask = 0ull;
    __syncthreads();
    if (tid < XT_KEYS && lane == 0 && bal) atomicOr(&use_mask, (unsigned long long)bal << (tid & 32));
    if (tid == 0) {      // bar[0]: the K rows have landed, bar[1]: the V rows
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(xt_smem_u32(bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(xt_smem_u32(bar + 1)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const unsigned long long um = use_mask;
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(xt_smem_u32(bar)), "r"((unsigned)__popcll(um) * MV2D_C * 4) : "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(xt_smem_u32(bar + 1)), "r"((unsigned)__popcll(um) * MV2D_C * 4) : "memory");
    }
    __syncthreads();            // barrier armed before the copies are issued and before anyone polls it
    if (use) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(xt_smem_u32(Ks + tid * XTM_PITCH)), "l"(a.kp + off), "r"(MV2D_C * 4), "r"(xt_smem_u32(bar)) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(xt_smem_u32(Vs + tid * XTM_PITCH)), "l"(a.vp + off), "r"(MV2D_C * 4), "r"(xt_smem_u32(bar + 1)) : "memory");
    }
    if (um != ~0ull) {
        for (int i = tid; i < XT_KEYS * MV2D_C; i += XTM_THREADS) {
            const int key = i >> 8;
            if (!((um >> key) & 1ull)) { Ks[key * XTM_PITCH + (i & 255)] = 0.f; Vs[key * XTM_PITCH + (i & 255)] = 0.f; }
        }
        __syncthreads();
    }
    bool waited = false, waited_v = false;
    const float* Kh = Ks + hd * 32;
    const float* Vh = Vs + hd * 32;
    for (int b = blockIdx.y * 2 + par; b < nblk; b += 2 * a.qsplit) {
        // ---- the block's 16 queries: rows g and g + 8 of this lane
        const int i0 = b * 16 + g, i1 = i0 + 8;
        const bool ok0 = i0 < cnt, ok1 = i1 < cnt;
        const long long e0 = (long long)t * a.g.Np + (ok0 ? i0 : b * 16), e1 = (long long)t * a.g.Np + (ok1 ? i1 : b * 16);
        const int n0 = a.tile_q[e0], n1 = a.tile_q[e1];
        const unsigned long long m0 = ok0 ? a.tile_mask[e0] : 0ull, m1 = ok1 ? a.tile_mask[e1] : 0ull;
        uint32_t qh[4][4], ql[4][4];
        {
            const float* q0 = a.q + (long long)n0 * MV2D_C + hd * 32 + tg;
            const float* q1 = a.q + (long long)n1 * MV2D_C + hd * 32 + tg;
#pragma unroll
            for (int kc = 0; kc < 4; ++kc) {
                split_tf32_reg(__ldg(q0 + kc * 8), qh[kc][0], ql[kc][0]);
                split_tf32_reg(__ldg(q1 + kc * 8), qh[kc][1], ql[kc][1]);
                split_tf32_reg(__ldg(q0 + kc * 8 + 4), qh[kc][2], ql[kc][2]);
                split_tf32_reg(__ldg(q1 + kc * 8 + 4), qh[kc][3], ql[kc][3]);
            }
        }
        if (!waited) { xt_wait(bar); waited = true; }
        // ---- S = Q K^T
        // (the MMAs are issued in groups of four independent accumulators: an mma.sync that reads the accumulator the
        // previous one wrote stalls the warp for the MMA latency, and `asm volatile` keeps the source order)
        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
#pragma unroll
            for (int jh = 0; jh < 2; ++jh) {
                uint32_t bh0[4], bl0[4], bh1[4], bl1[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const float* krow = Kh + ((jh * 4 + jj) * 8 + g) * XTM_PITCH + tg;
                    split_tf32_reg(krow[kc * 8], bh0[jj], bl0[jj]);
                    split_tf32_reg(krow[kc * 8 + 4], bh1[jj], bl1[jj]);
                }
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) mma_tf32_16x8x8(s[jh * 4 + jj], qh[kc], bh0[jj], bh1[jj]);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) mma_tf32_16x8x8(s[jh * 4 + jj], ql[kc], bh0[jj], bh1[jj]);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) mma_tf32_16x8x8(s[jh * 4 + jj], qh[kc], bl0[jj], bl1[jj]);
            }
        }
        // ---- masked softmax numerators over the tile's keys: rows g (values 0, 1) and g + 8 (values 2, 3)
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k0 = j * 8 + 2 * tg;
            if (!((m0 >> k0) & 1ull)) s[j][0] = -INFINITY;
            if (!((m0 >> (k0 + 1)) & 1ull)) s[j][1] = -INFINITY;
            if (!((m1 >> k0) & 1ull)) s[j][2] = -INFINITY;
            if (!((m1 >> (k0 + 1)) & 1ull)) s[j][3] = -INFINITY;
            mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
            mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float sub0 = ok0 ? mx0 : 0.f, sub1 = ok1 ? mx1 : 0.f;      // rows past the list: all keys masked, exp(-inf - 0) = 0
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j][0] = __expf(s[j][0] - sub0); s[j][1] = __expf(s[j][1] - sub0);
            s[j][2] = __expf(s[j][2] - sub1); s[j][3] = __expf(s[j][3] - sub1);
            l0 += s[j][0] + s[j][1];
            l1 += s[j][2] + s[j][3];
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        if (!waited_v) { xt_wait(bar + 1); waited_v = true; }
        // ---- O = P V : MMA k index tg <-> key 8 j + 2 tg, tg + 4 <-> key 8 j + 2 tg + 1
        float o[4][4];
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) o[nb][0] = o[nb][1] = o[nb][2] = o[nb][3] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint32_t ph[4], pl[4];
            split_tf32_reg(s[j][0], ph[0], pl[0]);      // (row g,     k = tg)
            split_tf32_reg(s[j][2], ph[1], pl[1]);      // (row g + 8, k = tg)
            split_tf32_reg(s[j][1], ph[2], pl[2]);      // (row g,     k = tg + 4)
            split_tf32_reg(s[j][3], ph[3], pl[3]);      // (row g + 8, k = tg + 4)
            const float* vrow = Vh + (j * 8 + 2 * tg) * XTM_PITCH + g;
            uint32_t bh0[4], bl0[4], bh1[4], bl1[4];
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                split_tf32_reg(vrow[nb * 8], bh0[nb], bl0[nb]);
                split_tf32_reg(vrow[XTM_PITCH + nb * 8], bh1[nb], bl1[nb]);
            }
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) mma_tf32_16x8x8(o[nb], ph, bh0[nb], bh1[nb]);
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) mma_tf32_16x8x8(o[nb], pl, bh0[nb], bh1[nb]);
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) mma_tf32_16x8x8(o[nb], ph, bl0[nb], bl1[nb]);
        }
        // ---- records
        if (ok0) {
            float* r = a.rec + e0 * XT_REC;
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) *reinterpret_cast<float2*>(r + hd * 32 + nb * 8 + 2 * tg) = make_float2(o[nb][0], o[nb][1]);
            if (tg == 0) { r[256 + hd] = mx0; r[264 + hd] = l0; }
        }
        if (ok1) {
            float* r = a.rec + e1 * XT_REC;
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) *reinterpret_cast<float2*>(r + hd * 32 + nb * 8 + 2 * tg) = make_float2(o[nb][2], o[nb][3]);
            if (tg == 0) { r[256 + hd] = mx1; r[264 + hd] = l1; }
        }
    }
    if (!waited) xt_wait(bar);      // a CTA must not exit while its bulk copies are in flight
    if (!waited_v) xt_wait(bar + 1);
}

struct XtMergeArgs {
    XtGeom g;
    const int* qlist;                  // [N, ntiles] record ids of the query (xt_list_kernel)
    const int* qcnt;                   // [N]
    const float* rec;
    float* ctx;                        // [N,256] softmax-weighted mean of the projected values (heads concatenated)
    float* ctx_lo;                     // nullable: ctx then holds the TF32 hi part, this the lo part (3xTF32 output projection)
};

#define XT_MERGE_THREADS 256

// grid = N, 256 threads, dynamic shared memory = cnt-independent: ntiles * (4 + 32) bytes.
// The query's record ids are staged in shared memory, the per-head maxima are found first, then the weight
// exp(m_r - M) of every (record, head) is computed once; the accumulation is nothing but independent, coalesced
// 16-byte loads: thread = (record phase 0..3, float4 of the 256 channels), records phase, phase+4, ... in ascending
// order, the four phases folded in a fixed order => bitwise reproducible.
__global__ void __launch_bounds__(XT_MERGE_THREADS) xt_merge_kernel(XtMergeArgs a) {
    pdl_wait();
    pdl_trigger();
    const int n = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    extern __shared__ __align__(16) unsigned char xm_smem[];
    int* list = reinterpret_cast<int*>(xm_smem);                                  // [tiles_ps]
    float* wgt = reinterpret_cast<float*>(xm_smem) + a.g.tiles_ps;                 // [tiles_ps][8]
    __shared__ float Ms[8][8], Mg[8], Lw[8][8];
    __shared__ float4 fold[3][64];
    const int cnt = a.qcnt[n];
    for (int r = tid; r < cnt; r += XT_MERGE_THREADS) list[r] = a.qlist[(long long)n * a.g.tiles_ps + r];
    __syncthreads();
    const int h = tid & 7, rsub = tid >> 3;             // (record phase 0..31, head)
    // ---- global max per head (the raw maxima are parked in wgt)
    {
        float mx = -INFINITY;
        for (int r = rsub; r < cnt; r += 32) {
            const float m = __ldcg(a.rec + (long long)list[r] * XT_REC + 256 + h);
            wgt[r * 8 + h] = m;
            mx = fmaxf(mx, m);
        }
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
        if (lane < 8) Ms[warp][lane] = mx;
    }
    __syncthreads();
    if (tid < 8) {
        float mx = Ms[0][tid];
#pragma unroll
        for (int w = 1; w < 8; ++w) mx = fmaxf(mx, Ms[w][tid]);
        Mg[tid] = mx;
    }
    __syncthreads();
    {   // ---- weights, and l = sum_r l_r * w_r (thread (rsub, h) owns records rsub, rsub+32, ...: fixed order)
        const float Mh = Mg[h];
        float lsum = 0.f;
        for (int r = rsub; r < cnt; r += 32) {
            const float w = __expf(wgt[r * 8 + h] - Mh);
            wgt[r * 8 + h] = w;
            lsum = fmaf(__ldcg(a.rec + (long long)list[r] * XT_REC + 264 + h), w, lsum);
        }
        lsum += __shfl_xor_sync(0xffffffffu, lsum, 8);
        lsum += __shfl_xor_sync(0xffffffffu, lsum, 16);
        if (lane < 8) Lw[warp][lane] = lsum;
    }
    __syncthreads();
    const int ph = tid >> 6, c4 = tid & 63, hc = c4 >> 3;       // record phase, float4 index, its head
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int r = ph;
    for (; r + 12 < cnt; r += 16) {
        float4 x[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = __ldcg(reinterpret_cast<const float4*>(a.rec + (long long)list[r + 4 * j] * XT_REC) + c4);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float w = wgt[(r + 4 * j) * 8 + hc];
            acc.x = fmaf(x[j].x, w, acc.x); acc.y = fmaf(x[j].y, w, acc.y); acc.z = fmaf(x[j].z, w, acc.z); acc.w = fmaf(x[j].w, w, acc.w);
        }
    }
    for (; r < cnt; r += 4) {
        const float4 x = __ldcg(reinterpret_cast<const float4*>(a.rec + (long long)list[r] * XT_REC) + c4);
        const float w = wgt[r * 8 + hc];
        acc.x = fmaf(x.x, w, acc.x); acc.y = fmaf(x.y, w, acc.y); acc.z = fmaf(x.z, w, acc.z); acc.w = fmaf(x.w, w, acc.w);
    }
    if (ph > 0) fold[ph - 1][c4] = acc;
    __syncthreads();
    if (ph == 0) {
#pragma unroll
        for (int w = 0; w < 3; ++w) { const float4 y = fold[w][c4]; acc.x += y.x; acc.y += y.y; acc.z += y.z; acc.w += y.w; }
        float l = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) l += Lw[w][hc];
        const float inv = l > 0.f ? 1.f / l : 0.f;
        float4 r = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
        if (a.ctx_lo) {
            const float4 hi = make_float4(round_tf32(r.x), round_tf32(r.y), round_tf32(r.z), round_tf32(r.w));
            reinterpret_cast<float4*>(a.ctx_lo + (long long)n * MV2D_C)[c4] =
                make_float4(round_tf32(r.x - hi.x), round_tf32(r.y - hi.y), round_tf32(r.z - hi.z), round_tf32(r.w - hi.w));
            r = hi;
        }
        reinterpret_cast<float4*>(a.ctx + (long long)n * MV2D_C)[c4] = r;
    }
}

}  // namespace mv2d
