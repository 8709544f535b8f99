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
#define XT_THREADS 256
#define XT_SMEM_BYTES (2 * XT_KEYS * MV2D_C * 4 + (XT_THREADS / 32) * XT_KEYS * 8 * 4 + 64)

struct XtGeom {
    int N, V, h, w, tiles_x, tiles_y, ntiles;
};

struct XtPrepArgs {
    XtGeom g;
    const uint32_t* keymask; int mask_words;
    int* tile_cnt;                     // [ntiles]
    uint16_t* tile_q;                  // [ntiles, N] queries with a key in the tile, ascending
    unsigned long long* tile_mask;     // [ntiles, N] their 64-bit key masks (bit r*8+c = cell (ty*8+r, tx*8+c))
    short* slot_of;                    // [N, ntiles] position of the query in the tile's list, -1 = none
};

// grid = ntiles, 256 threads
__global__ void __launch_bounds__(256) xt_prep_kernel(XtPrepArgs a) {
    pdl_wait();
    pdl_trigger();
    const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = t % a.g.tiles_x, ty = (t / a.g.tiles_x) % a.g.tiles_y, v = t / (a.g.tiles_x * a.g.tiles_y);
    const int ncols = min(XT_TS, a.g.w - tx * XT_TS);
    __shared__ int wsum[8];
    int running = 0;
    for (int base = 0; base < a.g.N; base += 256) {
        const int n = base + tid;
        unsigned long long m64 = 0ull;
        if (n < a.g.N) {
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
        const unsigned bal = __ballot_sync(0xffffffffu, active);
        if (lane == 0) wsum[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { const int c = wsum[i]; if (i < warp) before += c; total += c; }
        const int pos = running + before + __popc(bal & ((1u << lane) - 1u));
        if (n < a.g.N) {
            if (active) {
                a.tile_q[(long long)t * a.g.N + pos] = (uint16_t)n;
                a.tile_mask[(long long)t * a.g.N + pos] = m64;
            }
            a.slot_of[(long long)n * a.g.ntiles + t] = active ? (short)pos : (short)-1;
        }
        running += total;
        __syncthreads();
    }
    if (tid == 0) a.tile_cnt[t] = running;
}

struct XtAttnArgs {
    XtGeom g;
    const float* q;                    // [N,256] projected queries, 1/sqrt(32) folded in
    const float* kp; const float* vp;  // [V*h*w,256] projected keys / values of this layer
    const int* tile_cnt; const uint16_t* tile_q; const unsigned long long* tile_mask;
    float* rec;                        // [ntiles*N, XT_REC]
    int qsplit;                        // gridDim.y: the tile's query list is dealt round-robin to this many CTAs
};

// transpose-reduce 8 per-lane partials over the 8 lanes of a head group: 7 shuffles instead of 24.
// result: lane holds in v[0] the full sum of index (lane & 7).
__device__ __forceinline__ void reduce8_in8(float (&v)[8], int lane) {
    const bool up4 = lane & 4, up2 = lane & 2, up1 = lane & 1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = up4 ? v[i] : v[i + 4], keep = up4 ? v[i + 4] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = up2 ? v[i] : v[i + 2], keep = up2 ? v[i + 2] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    const float send = up1 ? v[0] : v[1], keep = up1 ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
}

__device__ __forceinline__ uint32_t xt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// grid = (ntiles, qsplit), 256 threads, XT_SMEM_BYTES dynamic shared memory
__global__ void __launch_bounds__(XT_THREADS, 1) xt_attn_kernel(XtAttnArgs a) {
    pdl_wait();
    pdl_trigger();
    const int t = blockIdx.x;
    const int cnt = a.tile_cnt[t];
    constexpr int NW = XT_THREADS / 32;
    if ((int)blockIdx.y * NW >= cnt) return;
    extern __shared__ __align__(128) unsigned char xt_smem[];
    float* Ks = reinterpret_cast<float*>(xt_smem);
    float* Vs = Ks + XT_KEYS * MV2D_C;
    float* scb = Vs + XT_KEYS * MV2D_C;                          // [NW][64 slots][8 heads]
    uint64_t* bar = reinterpret_cast<uint64_t*>(scb + NW * XT_KEYS * 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tx = t % a.g.tiles_x, ty = (t / a.g.tiles_x) % a.g.tiles_y, v = t / (a.g.tiles_x * a.g.tiles_y);
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
    bool waited = false;
    const int step = a.qsplit * NW;
    int i = blockIdx.y * NW + warp;
    // software pipeline: the next query's list entry and q slice are in flight while this one is processed
    int n_nx = 0; unsigned long long m_nx = 0ull; float4 q0_nx = make_float4(0.f, 0.f, 0.f, 0.f), q1_nx = q0_nx;
    auto fetch = [&](int ii) {
        n_nx = a.tile_q[(long long)t * a.g.N + ii];
        m_nx = a.tile_mask[(long long)t * a.g.N + ii];
        q0_nx = __ldg(reinterpret_cast<const float4*>(a.q + (long long)n_nx * MV2D_C + lane * 4));
        q1_nx = __ldg(reinterpret_cast<const float4*>(a.q + (long long)n_nx * MV2D_C + 128 + lane * 4));
    };
    if (i < cnt) fetch(i);
    for (; i < cnt; i += step) {
        // this lane's slice of q: channels 4*lane..+3 (head lane>>3) and 128+4*lane..+3 (head 4+(lane>>3))
        const unsigned long long m64 = m_nx;
        const float4 q0 = q0_nx, q1 = q1_nx;
        const int nk = __popcll(m64);
        if (i + step < cnt) fetch(i + step);
        if (!waited) {
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(xt_smem_u32(bar)) : "memory");
            waited = true;
        }
        // ---- logits of the query's keys in this tile, 8 keys per step
        unsigned long long mm = m64;
        for (int s0 = 0; s0 < nk; s0 += 8) {
            float p0[8], p1[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                p0[j] = 0.f; p1[j] = 0.f;
                if (mm) {                                   // warp-uniform
                    const int k = __ffsll((long long)mm) - 1;
                    mm &= mm - 1;
                    const float* row = Ks + k * MV2D_C;
                    const float4 k0 = *reinterpret_cast<const float4*>(row + lane * 4);
                    const float4 k1 = *reinterpret_cast<const float4*>(row + 128 + lane * 4);
                    p0[j] = fmaf(q0.w, k0.w, fmaf(q0.z, k0.z, fmaf(q0.y, k0.y, q0.x * k0.x)));
                    p1[j] = fmaf(q1.w, k1.w, fmaf(q1.z, k1.z, fmaf(q1.y, k1.y, q1.x * k1.x)));
                }
            }
            reduce8_in8(p0, lane);
            reduce8_in8(p1, lane);
            const int slot = s0 + (lane & 7);
            if (slot < nk) {
                sc[slot * 8 + (lane >> 3)] = p0[0];
                sc[slot * 8 + 4 + (lane >> 3)] = p1[0];
            }
        }
        __syncwarp();
        // ---- softmax statistics: lane = (head = lane & 7, phase = lane >> 3), slots phase, phase+4, ...
        float mx = -INFINITY;
        {
            const int h = lane & 7;
            for (int s = lane >> 3; s < nk; s += 4) mx = fmaxf(mx, sc[s * 8 + h]);
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
            float sum = 0.f;
            for (int s = lane >> 3; s < nk; s += 4) {
                const float p = __expf(sc[s * 8 + h] - mx);
                sc[s * 8 + h] = p;
                sum += p;
            }
            sum += __shfl_xor_sync(0xffffffffu, sum, 8);
            sum += __shfl_xor_sync(0xffffffffu, sum, 16);
            float* r = a.rec + ((long long)t * a.g.N + i) * XT_REC;
            if (lane < 8) { r[256 + lane] = mx; r[264 + lane] = sum; }
        }
        __syncwarp();
        // ---- acc = sum_k p_k * V_k over the query's keys
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
        mm = m64;
        for (int s = 0; s < nk; ++s) {
            const int k = __ffsll((long long)mm) - 1;
            mm &= mm - 1;
            const float* row = Vs + k * MV2D_C;
            const float4 v0 = *reinterpret_cast<const float4*>(row + lane * 4);
            const float4 v1 = *reinterpret_cast<const float4*>(row + 128 + lane * 4);
            const float w0 = sc[s * 8 + (lane >> 3)], w1 = sc[s * 8 + 4 + (lane >> 3)];
            a0.x = fmaf(w0, v0.x, a0.x); a0.y = fmaf(w0, v0.y, a0.y); a0.z = fmaf(w0, v0.z, a0.z); a0.w = fmaf(w0, v0.w, a0.w);
            a1.x = fmaf(w1, v1.x, a1.x); a1.y = fmaf(w1, v1.y, a1.y); a1.z = fmaf(w1, v1.z, a1.z); a1.w = fmaf(w1, v1.w, a1.w);
        }
        {
            float* r = a.rec + ((long long)t * a.g.N + i) * XT_REC;
            *reinterpret_cast<float4*>(r + lane * 4) = a0;
            *reinterpret_cast<float4*>(r + 128 + lane * 4) = a1;
        }
        __syncwarp();           // sc is rewritten by the next query of this warp
    }
    // a CTA must not exit while its bulk copies are in flight
    if (!waited) {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(xt_smem_u32(bar)) : "memory");
    }
}

struct XtMergeArgs {
    XtGeom g;
    const short* slot_of;              // [N, ntiles]
    const float* rec;
    float* ctx;                        // [N,256] softmax-weighted mean of the projected values (heads concatenated)
};

#define XT_MERGE_THREADS 128
#define XT_MERGE_MAXT 2048             // tiles a query's list is sized for (V*ceil(h/8)*ceil(w/8) <= 2048)

// grid = N, 128 threads
__global__ void __launch_bounds__(XT_MERGE_THREADS) xt_merge_kernel(XtMergeArgs a) {
    pdl_wait();
    pdl_trigger();
    const int n = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    __shared__ int list[XT_MERGE_MAXT];
    __shared__ int cnt_s;
    __shared__ float Ms[4][8], Mg[8], Ls[4][8];
    __shared__ float4 accs[3][64];
    if (warp == 0) {
        int running = 0;
        for (int base = 0; base < a.g.ntiles; base += 32) {
            const int t = base + lane;
            const int s = t < a.g.ntiles ? (int)a.slot_of[(long long)n * a.g.ntiles + t] : -1;
            const unsigned bal = __ballot_sync(0xffffffffu, s >= 0);
            if (s >= 0) list[running + __popc(bal & ((1u << lane) - 1u))] = t * a.g.N + s;
            running += __popc(bal);
        }
        if (lane == 0) cnt_s = running;
    }
    __syncthreads();
    const int cnt = cnt_s;
    // ---- global max per head
    {
        const int h = lane & 7;
        float mx = -INFINITY;
        for (int r = warp * 4 + (lane >> 3); r < cnt; r += 16) mx = fmaxf(mx, __ldcg(a.rec + (long long)list[r] * XT_REC + 256 + h));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
        if (lane < 8) Ms[warp][lane] = mx;
    }
    __syncthreads();
    if (tid < 8) Mg[tid] = fmaxf(fmaxf(Ms[0][tid], Ms[1][tid]), fmaxf(Ms[2][tid], Ms[3][tid]));
    __syncthreads();
    // ---- weighted sums: warp w takes records w, w+4, ... in ascending order
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    float lsum = 0.f;
    const float M0 = Mg[lane >> 3], M1 = Mg[4 + (lane >> 3)], Ml = Mg[lane & 7];
    for (int r = warp; r < cnt; r += 4) {
        const float* rp = a.rec + (long long)list[r] * XT_REC;
        const float4 x0 = __ldcg(reinterpret_cast<const float4*>(rp + lane * 4));
        const float4 x1 = __ldcg(reinterpret_cast<const float4*>(rp + 128 + lane * 4));
        const float w0 = __expf(__ldcg(rp + 256 + (lane >> 3)) - M0);
        const float w1 = __expf(__ldcg(rp + 260 + (lane >> 3)) - M1);
        a0.x = fmaf(x0.x, w0, a0.x); a0.y = fmaf(x0.y, w0, a0.y); a0.z = fmaf(x0.z, w0, a0.z); a0.w = fmaf(x0.w, w0, a0.w);
        a1.x = fmaf(x1.x, w1, a1.x); a1.y = fmaf(x1.y, w1, a1.y); a1.z = fmaf(x1.z, w1, a1.z); a1.w = fmaf(x1.w, w1, a1.w);
        if (lane < 8) lsum = fmaf(__ldcg(rp + 264 + lane), __expf(__ldcg(rp + 256 + lane) - Ml), lsum);
    }
    if (lane < 8) Ls[warp][lane] = lsum;
    if (warp > 0) { accs[warp - 1][lane] = a0; accs[warp - 1][32 + lane] = a1; }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int w = 0; w < 3; ++w) {
            const float4 y0 = accs[w][lane], y1 = accs[w][32 + lane];
            a0.x += y0.x; a0.y += y0.y; a0.z += y0.z; a0.w += y0.w;
            a1.x += y1.x; a1.y += y1.y; a1.z += y1.z; a1.w += y1.w;
        }
        const int h0 = lane >> 3;
        const float l0 = Ls[0][h0] + Ls[1][h0] + Ls[2][h0] + Ls[3][h0];
        const float l1 = Ls[0][4 + h0] + Ls[1][4 + h0] + Ls[2][4 + h0] + Ls[3][4 + h0];
        const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
        float* o = a.ctx + (long long)n * MV2D_C;
        *reinterpret_cast<float4*>(o + lane * 4) = make_float4(a0.x * i0, a0.y * i0, a0.z * i0, a0.w * i0);
        *reinterpret_cast<float4*>(o + 128 + lane * 4) = make_float4(a1.x * i1, a1.y * i1, a1.z * i1, a1.w * i1);
    }
}

}  // namespace mv2d
