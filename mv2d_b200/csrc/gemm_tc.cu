// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[M,N] = epi( A[M,K] . W[N,K]^T )
//   * operands are fp32 in HBM, K-contiguous; TMA (cp.async.bulk.tensor) stages 128-byte-swizzled
//     [rows x 32 floats] boxes into shared memory, one elected thread issues
//     tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=BN, K=8), the accumulator tile lives in TMEM
//     and is read back with tcgen05.ld by four epilogue warps.
//   * PASSES = 1: single-pass TF32 (operands pre-rounded to TF32 by their producers) -- used for
//     the PE MLPs only, where SURVEY.md App. E shows TF32 is inside the parity gate.
//   * PASSES = 3: error-compensated 3xTF32 (a = a_hi + a_lo, both exactly TF32-representable;
//     acc += a_hi*w_hi + a_hi*w_lo + a_lo*w_hi) -- fp32-grade accuracy for the query-generator
//     3x3 conv, whose result feeds the precision-critical reference points.
//   * IM2COL: the A operand is the implicit im2col of the [N,7,7,256] RoI tokens: a 4-D tensor
//     map with box (32 ch, 7, 7, 1) loaded at (c0, dx-1, dy-1, roi); TMA zero-fills the halo.
//     Two RoIs share one 128-row tile (rows 0..48 and 64..112).
//   * IM2COL = 3: the same operand, FIVE RoIs per 256-row tile: the five 49-row boxes land back to back (rows 0..244; the
//     128B swizzle is a function of the shared-memory address, so boxes need not start on a 1024-byte atom) and every
//     K step issues two M = 128 MMAs (rows 0..127, 128..255) on the same W tile into two TMEM accumulators.  245 of
//     256 MMA rows are real (98 of 128 before) and a W tile is fetched once per five RoIs instead of once per two.
//   * IM2COL = 2: the same trick on a whole feature map [V,h,w,256] (3x3 conv, padding 1, of the FPN neck): a
//     tile is 16 rows x 8 columns of pixels, the box (32 ch, 8, 16, 1) is loaded at (c0, x0+dx-1, y0+dy-1, v).
// Warp roles (192 threads): warp 0 TMA producer, warp 1 TMEM alloc + MMA issue, warps 2..5
// epilogue (TMEM lane quadrant = warp_idx % 4).  One output tile per CTA; several CTAs per SM
// overlap each other's prologue/epilogue.
#include <cuda.h>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "tc_ptx.cuh"

namespace mv2d {

static constexpr int TC_THREADS = 192;

struct TcArgs {
    float* C; float* C_lo; int ldc;
    int nkb_per_split; long long split_stride;
    const float* bias;
    int M, N, K;
    int flags;
    int n_rois;                                                        // IM2COL 1: number of RoIs
    int fm_h, fm_w;                                                    // IM2COL 2: feature grid (tiles of 16 x 8 pixels)
    const uint8_t* m_tile_live;                                        // nullable [m_tiles]: 0 = nobody reads this row tile, skip it
    const float* gx; const float* gs; const float* gfeat; float* kin;  // GEMM_GATE extras
    int gs_mod;                                                        // > 0: gs row = out row % gs_mod
    int groups; int group_rows;                                        // grouped launch: blockIdx.z = group (no split-K)
    int n_switch;                                                      // > 0: n tiles at columns >= n_switch load A through tmAlo's slot pair (see launch)
};

// RAW (PASSES == 3 only): the operands arrive as plain fp32 -- ONE copy of each tile crosses the L2->SM fabric instead
// of a pre-split hi and lo copy -- and the four epilogue warps, idle during the main loop, split every stage in
// shared memory (hi in place, lo beside it; element-wise, so the 128B swizzle TMA wrote is preserved) before the
// MMA warp consumes it: full_bar (TMA landed) -> split -> fence.proxy.async -> conv_bar -> tcgen05.mma.
template <int BN, int PASSES, int IM2COL, int STAGES, bool RAW = false>
__global__ void __launch_bounds__(TC_THREADS, (PASSES == 1 ? 2 : 1))
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
               const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmWlo,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmA2lo, TcArgs g) {
    static_assert(!RAW || (PASSES == 3 && !IM2COL), "RAW is the in-kernel split of the plain 3xTF32 GEMM");
    constexpr bool PACK5 = IM2COL == 3;
    constexpr int NACC = PACK5 ? 2 : 1;             // accumulators (M = 128 each) per CTA
    constexpr int A_BYTES = NACC * TC_BM * TC_BK * 4;      // 16 KB per 128 rows
    constexpr int W_BYTES = BN * TC_BK * 4;
    constexpr int NOP = PASSES == 3 ? 2 : 1;        // hi (+ lo) copies of each operand
    constexpr int STAGE_BYTES = NOP * (A_BYTES + W_BYTES);
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment is required by the 128B swizzle atoms
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint64_t* conv_bar = tmem_full_bar + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(conv_bar + STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tile = blockIdx.y, n0 = blockIdx.x * BN;
    if (g.m_tile_live != nullptr) {             // CTA-uniform: before any barrier / TMEM allocation
        pdl_wait();
        if (g.m_tile_live[m_tile] == 0) return;
    }
    const bool grouped = !IM2COL && g.groups > 1;
    const int grp = grouped ? blockIdx.z : 0;
    const int a_row0 = grp * g.group_rows;      // first A / C row of this group
    const int w_row0 = grp * g.N;               // first W row / bias element of this group
    const int nkb = g.nkb_per_split, kb0 = grouped ? 0 : blockIdx.z * g.nkb_per_split;   // this CTA's K range (split-K)
    // CTAs that share an operand tile walk K in rotated order, so they do not all hit the same L2 lines at once
    const int rot = IM2COL ? 0 : (int)((blockIdx.x + 3 * blockIdx.y) % nkb);
    constexpr int ROI_BYTES = MV2D_TOK * TC_BK * 4;        // one im2col box: 49 rows x 128 B

    // output columns >= n_switch read their A rows from the second operand (CTA-uniform)
    const bool second_a = !IM2COL && g.n_switch > 0 && n0 >= g.n_switch;
    const CUtensorMap* mapA = second_a ? &tmA2 : &tmA;
    const CUtensorMap* mapAlo = second_a ? &tmA2lo : &tmAlo;
    if (warp == 0 && lane == 0) {
        tmap_prefetch(mapA); tmap_prefetch(&tmW);
        if (PASSES == 3) { tmap_prefetch(mapAlo); tmap_prefetch(&tmWlo); }
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); mbar_init(&conv_bar[s], 128); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(NACC * BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    // everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the previous kernel
    pdl_wait();
    pdl_trigger();

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES, ph = (kb / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t* st = smem + s * STAGE_BYTES;
                // bytes TMA will deliver: an im2col box is 49 rows x 128 B per RoI, not a full 64-row half tile
                constexpr int A_TX = IM2COL == 1 ? 2 * ROI_BYTES : (PACK5 ? 5 * ROI_BYTES : A_BYTES);
                mbar_expect_tx(&full_bar[s], (RAW ? 1 : NOP) * (A_TX + W_BYTES));
                if (IM2COL == 2) {
                    const int kg = kb0 + kb;
                    const int tap = kg / (MV2D_C / TC_BK), c0 = (kg % (MV2D_C / TC_BK)) * TC_BK;
                    const int tiles_x = (g.fm_w + 7) >> 3, tiles_y = (g.fm_h + 15) >> 4;
                    const int tx = m_tile % tiles_x, ty = (m_tile / tiles_x) % tiles_y, v = m_tile / (tiles_x * tiles_y);
                    const int x0 = tx * 8 + tap % 3 - 1, y0 = ty * 16 + tap / 3 - 1;
                    tma_load_4d(&tmA, &full_bar[s], st, c0, x0, y0, v);
                    if (PASSES == 3) tma_load_4d(&tmAlo, &full_bar[s], st + A_BYTES, c0, x0, y0, v);
                } else if (PACK5) {
                    const int kg = kb0 + kb;
                    const int tap = kg / (MV2D_C / TC_BK), c0 = (kg % (MV2D_C / TC_BK)) * TC_BK;
                    const int dx = tap % 3 - 1, dy = tap / 3 - 1;
#pragma unroll
                    for (int r = 0; r < 5; ++r) {   // five RoIs back to back: rows 49 r .. 49 r + 48 (RoIs past the end: zero fill)
                        tma_load_4d(&tmA, &full_bar[s], st + r * ROI_BYTES, c0, dx, dy, m_tile * 5 + r);
                        tma_load_4d(&tmAlo, &full_bar[s], st + A_BYTES + r * ROI_BYTES, c0, dx, dy, m_tile * 5 + r);
                    }
                } else if (IM2COL == 1) {
                    const int kg = kb0 + kb;
                    const int tap = kg / (MV2D_C / TC_BK), c0 = (kg % (MV2D_C / TC_BK)) * TC_BK;
                    const int dx = tap % 3 - 1, dy = tap / 3 - 1;
#pragma unroll
                    for (int r = 0; r < 2; ++r) {   // two RoIs per 128-row tile, 64 rows apart
                        tma_load_4d(&tmA, &full_bar[s], st + r * (A_BYTES / 2), c0, dx, dy, m_tile * 2 + r);
                        if (PASSES == 3)
                            tma_load_4d(&tmAlo, &full_bar[s], st + A_BYTES + r * (A_BYTES / 2), c0, dx, dy, m_tile * 2 + r);
                    }
                } else {
                    tma_load_2d(mapA, &full_bar[s], st, (kb0 + (kb + rot) % nkb) * TC_BK, a_row0 + m_tile * TC_BM);
                    if (PASSES == 3 && !RAW) tma_load_2d(mapAlo, &full_bar[s], st + A_BYTES, (kb0 + (kb + rot) % nkb) * TC_BK, a_row0 + m_tile * TC_BM);
                }
                tma_load_2d(&tmW, &full_bar[s], st + NOP * A_BYTES, (kb0 + (kb + rot) % nkb) * TC_BK, w_row0 + n0);
                if (PASSES == 3 && !RAW) tma_load_2d(&tmWlo, &full_bar[s], st + NOP * A_BYTES + W_BYTES, (kb0 + (kb + rot) % nkb) * TC_BK, w_row0 + n0);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // instruction descriptor: c=F32 [4,6), a=b=TF32 [7,10)/[10,13), K-major both, N>>3 [17,23), M>>4 [24,29)
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES, ph = (kb / STAGES) & 1;
                mbar_wait(RAW ? &conv_bar[s] : &full_bar[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
                const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + A_BYTES);
                const uint64_t w_hi = make_desc(sa + NOP * A_BYTES), w_lo = make_desc(sa + NOP * A_BYTES + W_BYTES);
#pragma unroll
                for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
                    const uint64_t adv = (uint64_t)((k * TC_UMMA_K * 4) >> 4);   // advance start address inside the swizzle row
#pragma unroll
                    for (int acc = 0; acc < NACC; ++acc) {
                        const uint64_t ra = adv + (uint64_t)((acc * TC_BM * TC_BK * 4) >> 4);    // rows 128 acc .. of the A tile
                        umma_tf32(tmem_base + acc * BN, a_hi + ra, w_hi + adv, idesc, (kb | k) != 0);
                        if (PASSES == 3) {
                            umma_tf32(tmem_base + acc * BN, a_hi + ra, w_lo + adv, idesc, 1);
                            umma_tf32(tmem_base + acc * BN, a_lo + ra, w_hi + adv, idesc, 1);
                        }
                    }
                }
                umma_commit(&empty_bar[s]);               // frees the smem slot when these MMAs retire
            }
            umma_commit(tmem_full_bar);                   // accumulator complete
        }
    } else {
        // ================= epilogue: TMEM -> registers -> (smem transpose) -> global =================
        // tcgen05.ld hands each thread one ROW of the tile (32 consecutive columns per chunk).  Storing that
        // directly would write 16-byte pieces 4 KB apart; instead every warp transposes its 32x32 chunk through
        // a private, XOR-swizzled 4 KB staging tile (the pipeline stages are free by now) so that each store /
        // load instruction of the warp touches 4 rows x 128 contiguous bytes.
        if (RAW) {
            // ---- operand split, stage by stage, while the MMA warp works on the previous ones
            const int te = threadIdx.x - 64;              // 0..127
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES, ph = (kb / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                float4* a4 = reinterpret_cast<float4*>(smem + s * STAGE_BYTES);
                float4* w4 = reinterpret_cast<float4*>(smem + s * STAGE_BYTES + NOP * A_BYTES);
                // all loads first (independent LDS.128, one shared-memory latency), then split and store
                constexpr int NA = A_BYTES / 16 / 128, NWV = W_BYTES / 16 / 128;
                float4 va[NA], vw[NWV];
#pragma unroll
                for (int i = 0; i < NA; ++i) va[i] = a4[te + i * 128];
#pragma unroll
                for (int i = 0; i < NWV; ++i) vw[i] = w4[te + i * 128];
                auto split4 = [](const float4 v, float4* hi, float4* lo) {
                    float4 h, l;
                    h.x = round_tf32(v.x); h.y = round_tf32(v.y); h.z = round_tf32(v.z); h.w = round_tf32(v.w);
                    l.x = round_tf32(v.x - h.x); l.y = round_tf32(v.y - h.y); l.z = round_tf32(v.z - h.z); l.w = round_tf32(v.w - h.w);
                    *hi = h; *lo = l;
                };
#pragma unroll
                for (int i = 0; i < NA; ++i) split4(va[i], a4 + te + i * 128, a4 + A_BYTES / 16 + te + i * 128);
#pragma unroll
                for (int i = 0; i < NWV; ++i) split4(vw[i], w4 + te + i * 128, w4 + W_BYTES / 16 + te + i * 128);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA's async-proxy reads
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&conv_bar[s])) : "memory");
            }
        }
        mbar_wait(tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;                           // TMEM lane quadrant this warp may access
        float* stg = reinterpret_cast<float*>(smem) + q * 1024;
        int acc_row0 = 0;                                 // first tile row of the accumulator being drained
        auto out_row = [&](int r, long long& orow) -> bool {   // tile row -> output row
            if (PACK5) {
                orow = (long long)m_tile * 5 * MV2D_TOK + r;
                return r < 5 * MV2D_TOK && orow < (long long)g.n_rois * MV2D_TOK;
            }
            if (IM2COL == 1) {
                const int roi = m_tile * 2 + (r >> 6), tok = r & 63;
                orow = (long long)roi * MV2D_TOK + tok;
                return tok < MV2D_TOK && roi < g.n_rois;
            }
            if (IM2COL == 2) {
                const int tiles_x = (g.fm_w + 7) >> 3, tiles_y = (g.fm_h + 15) >> 4;
                const int tx = m_tile % tiles_x, ty = (m_tile / tiles_x) % tiles_y, v = m_tile / (tiles_x * tiles_y);
                const int y = ty * 16 + (r >> 3), x = tx * 8 + (r & 7);
                orow = ((long long)v * g.fm_h + y) * g.fm_w + x;
                return y < g.fm_h && x < g.fm_w;
            }
            orow = (long long)a_row0 + m_tile * TC_BM + r;      // rows of the tile beyond the group's M belong to the next group
            return m_tile * TC_BM + r < g.M;
        };
        // registers (my row, 32 cols) -> global [4 rows x 128 B per instruction]
        auto store_t = [&](float* __restrict__ dst, const float (&x)[32], int n) {
            __syncwarp();
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4)
                *reinterpret_cast<float4*>(stg + lane * 32 + ((j4 ^ (lane & 7)) << 2)) =
                    make_float4(x[j4 * 4], x[j4 * 4 + 1], x[j4 * 4 + 2], x[j4 * 4 + 3]);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = i * 4 + (lane >> 3), cc = lane & 7;
                const float4 v4 = *reinterpret_cast<const float4*>(stg + rr * 32 + ((cc ^ (rr & 7)) << 2));
                long long orow;
                if (out_row(acc_row0 + q * 32 + rr, orow)) *reinterpret_cast<float4*>(dst + orow * g.ldc + n + cc * 4) = v4;
            }
        };
        // global -> registers (my row, 32 cols), same access pattern in reverse
        auto load_t = [&](const float* __restrict__ src, float (&x)[32], int n, int row_mod = 0) {
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = i * 4 + (lane >> 3), cc = lane & 7;
                long long orow;
                float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (out_row(acc_row0 + q * 32 + rr, orow)) v4 = __ldg(reinterpret_cast<const float4*>(src + (row_mod > 0 ? orow % row_mod : orow) * g.ldc + n + cc * 4));
                *reinterpret_cast<float4*>(stg + rr * 32 + ((cc ^ (rr & 7)) << 2)) = v4;
            }
            __syncwarp();
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
                const float4 v4 = *reinterpret_cast<const float4*>(stg + lane * 32 + ((j4 ^ (lane & 7)) << 2));
                x[j4 * 4] = v4.x; x[j4 * 4 + 1] = v4.y; x[j4 * 4 + 2] = v4.z; x[j4 * 4 + 3] = v4.w;
            }
        };
#pragma unroll 1
        for (int cc2 = 0; cc2 < NACC * (BN / 32); ++cc2) {
            const int c = cc2 % (BN / 32);
            acc_row0 = (cc2 / (BN / 32)) * TC_BM;
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cc2 * 32), v);
            const int n = n0 + c * 32;
            float x[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
            if (gridDim.z > 1 && !grouped) {   // split-K: raw partial sums, finished by the LN/reduce kernel
                store_t(g.C + blockIdx.z * g.split_stride, x, n);
                continue;
            }
            if (g.bias) {
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] += __ldg(g.bias + w_row0 + n + j);
            }
            if (g.flags & GEMM_RELU) {
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
            }
            if (g.flags & GEMM_CLAMP5E3) {
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = fminf(x[j], 5e3f);
            }
            if (g.flags & GEMM_GATE) {   // pe = gx * sigmoid(acc) + gs ; kin = pe + gfeat   (pe.py:44-48,166)
                float t[32];
                load_t(g.gx, t, n);
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = t[j] * sigmoid_f(x[j]);
                load_t(g.gs, t, n, g.gs_mod);
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] += t[j];
                if (g.kin) {
                    load_t(g.gfeat, t, n);
#pragma unroll
                    for (int j = 0; j < 32; ++j) t[j] += x[j];
                    store_t(g.kin, t, n);
                }
            }
            if (g.flags & GEMM_ROUND_TF32) {
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = round_tf32(x[j]);
            }
            if (g.flags & GEMM_SPLIT_OUT) {
                float lo[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) { const float hi = round_tf32(x[j]); lo[j] = round_tf32(x[j] - hi); x[j] = hi; }
                store_t(g.C_lo, lo, n);
            }
            store_t(g.C, x, n);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(NACC * BN));
    }
}

// ---------------------------------------------------------------------------- cluster-multicast variant
// The decoder's wide GEMMs (M ~ 300, 3xTF32, 128x64 tiles) are bound by L2->SM traffic: every N-tile CTA
// re-reads the same A tile (hi and lo).  Here the four CTAs of a thread-block cluster that share an M-tile
// each load one quarter of the A rows and TMA-multicast it into all four shared memories, so A crosses the
// L2->SM fabric once per cluster instead of four times.  A smem slot is refilled by the PEERS as well, so
// its "empty" barrier collects one tcgen05.commit arrival from each of the 4 CTAs (multicast commit).
#define TC_MC 4
template <int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_mc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
                  const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmWlo, TcArgs g) {
    constexpr int BN = 64, PASSES = 3;
    constexpr int A_BYTES = TC_BM * TC_BK * 4, W_BYTES = BN * TC_BK * 4;
    constexpr int STAGE_BYTES = 2 * (A_BYTES + W_BYTES);
    constexpr int A_QUARTER = A_BYTES / TC_MC;                 // 32 rows x 128 B
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tile = blockIdx.y, n0 = blockIdx.x * BN;
    const int nkb = g.nkb_per_split, kb0 = blockIdx.z * g.nkb_per_split;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    constexpr uint16_t mask = (1u << TC_MC) - 1;

    if (warp == 0 && lane == 0) {
        tmap_prefetch(&tmA); tmap_prefetch(&tmW); tmap_prefetch(&tmAlo); tmap_prefetch(&tmWlo);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], TC_MC); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                        // every CTA's barriers exist before any peer signals them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();
    pdl_trigger();

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES, ph = (kb / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);                       // all 4 CTAs released slot s
                uint8_t* st = smem + s * STAGE_BYTES;
                mbar_expect_tx(&full_bar[s], STAGE_BYTES);              // own W tiles + 4 multicast quarters of A hi/lo
                const int k = (kb0 + kb) * TC_BK;
                tma_load_2d_mc(&tmA, &full_bar[s], st + rank * A_QUARTER, k, m_tile * TC_BM + rank * 32, mask);
                tma_load_2d_mc(&tmAlo, &full_bar[s], st + A_BYTES + rank * A_QUARTER, k, m_tile * TC_BM + rank * 32, mask);
                tma_load_2d(&tmW, &full_bar[s], st + 2 * A_BYTES, k, n0);
                tma_load_2d(&tmWlo, &full_bar[s], st + 2 * A_BYTES + W_BYTES, k, n0);
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES, ph = (kb / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
                const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + A_BYTES);
                const uint64_t w_hi = make_desc(sa + 2 * A_BYTES), w_lo = make_desc(sa + 2 * A_BYTES + W_BYTES);
#pragma unroll
                for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
                    const uint64_t adv = (uint64_t)((k * TC_UMMA_K * 4) >> 4);
                    umma_tf32(tmem_base, a_hi + adv, w_hi + adv, idesc, (kb | k) != 0);
                    umma_tf32(tmem_base, a_hi + adv, w_lo + adv, idesc, 1);
                    umma_tf32(tmem_base, a_lo + adv, w_hi + adv, idesc, 1);
                }
                umma_commit_mc(&empty_bar[s], mask);      // slot s is free in THIS CTA: tell all 4 producers
            }
            umma_commit(tmem_full_bar);
        }
    } else {
        mbar_wait(tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;
        // the peers may still be multicasting into the pipeline stages: stage through a private area behind them
        float* stg = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256) + q * 1024;
        auto store_t = [&](float* __restrict__ dst, const float (&x)[32], int n) {
            __syncwarp();
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4)
                *reinterpret_cast<float4*>(stg + lane * 32 + ((j4 ^ (lane & 7)) << 2)) =
                    make_float4(x[j4 * 4], x[j4 * 4 + 1], x[j4 * 4 + 2], x[j4 * 4 + 3]);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = i * 4 + (lane >> 3), cc = lane & 7;
                const float4 v4 = *reinterpret_cast<const float4*>(stg + rr * 32 + ((cc ^ (rr & 7)) << 2));
                const long long orow = (long long)m_tile * TC_BM + q * 32 + rr;
                if (orow < g.M) *reinterpret_cast<float4*>(dst + orow * g.ldc + n + cc * 4) = v4;
            }
        };
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
            const int n = n0 + c * 32;
            float x[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
            if (gridDim.z > 1) { store_t(g.C + blockIdx.z * g.split_stride, x, n); continue; }
            if (g.bias) {
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] += __ldg(g.bias + n + j);
            }
            if (g.flags & GEMM_RELU) {
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
            }
            if (g.flags & GEMM_SPLIT_OUT) {
                float lo[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) { const float hi = round_tf32(x[j]); lo[j] = round_tf32(x[j] - hi); x[j] = hi; }
                store_t(g.C_lo, lo, n);
            }
            store_t(g.C, x, n);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                        // nobody leaves while a peer can still write its smem / barriers
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN));
    }
}

// ---------------------------------------------------------------------------- persistent variant
// For the GPU-filling 3xTF32 GEMMs of a batch (M = 2400 rows, N = 2048: the absorbed query projection and FFN layer 1,
// >= 2 tiles per SM).  One CTA per SM walks the (m tile, n tile) units round-robin; the 256 TMEM columns hold TWO
// 128-column accumulators, so the epilogue of unit i (TMEM -> registers -> smem transpose -> coalesced stores, hi / lo
// split) overlaps the main loop of unit i + 1 and the TMA producer runs ahead across unit boundaries.  The one-tile-per-CTA
// kernel above exposes barrier setup, TMEM allocation, pipeline fill and the whole epilogue once per tile (1 CTA / SM at
// 192 KB of shared memory): 33-40 us for these two GEMMs against ~9 us of tensor time.
// Pre-split operands only; epilogue: bias, ReLU, TF32 rounding, hi / lo split.
struct PgArgs {
    TcArgs t;
    int m_tiles, n_tiles;
};

template <int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
                       const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmWlo, PgArgs pg) {
    constexpr int BN = 128;
    constexpr int A_BYTES = TC_BM * TC_BK * 4, W_BYTES = BN * TC_BK * 4;
    constexpr int STAGE_BYTES = 2 * (A_BYTES + W_BYTES);       // 64 KB
    const TcArgs& g = pg.t;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* stg_all = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);        // 4 x 4 KB epilogue transpose tiles
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + 4 * 4096);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;       // [2]
    uint64_t* tmem_empty = tmem_full + 2;           // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = g.nkb_per_split, total = pg.m_tiles * pg.n_tiles;

    if (warp == 0 && lane == 0) {
        tmap_prefetch(&tmA); tmap_prefetch(&tmW); tmap_prefetch(&tmAlo); tmap_prefetch(&tmWlo);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();
    pdl_trigger();

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int u = blockIdx.x; u < total; u += gridDim.x) {
                const int m_tile = u / pg.n_tiles, n0 = (u % pg.n_tiles) * BN;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* st = smem + s * STAGE_BYTES;
                    mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                    tma_load_2d(&tmA, &full_bar[s], st, kb * TC_BK, m_tile * TC_BM);
                    tma_load_2d(&tmAlo, &full_bar[s], st + A_BYTES, kb * TC_BK, m_tile * TC_BM);
                    tma_load_2d(&tmW, &full_bar[s], st + 2 * A_BYTES, kb * TC_BK, n0);
                    tma_load_2d(&tmWlo, &full_bar[s], st + 2 * A_BYTES + W_BYTES, kb * TC_BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        if (lane == 0) {
            int it = 0, ui = 0;
            for (int u = blockIdx.x; u < total; u += gridDim.x, ++ui) {
                const int buf = ui & 1;
                mbar_wait(&tmem_empty[buf], ((ui >> 1) & 1) ^ 1);      // the epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = tmem_base + (uint32_t)(buf * BN);
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(&full_bar[s], ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
                    const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + A_BYTES);
                    const uint64_t w_hi = make_desc(sa + 2 * A_BYTES), w_lo = make_desc(sa + 2 * A_BYTES + W_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
                        const uint64_t adv = (uint64_t)((k * TC_UMMA_K * 4) >> 4);
                        umma_tf32(acc, a_hi + adv, w_hi + adv, idesc, (kb | k) != 0);
                        umma_tf32(acc, a_hi + adv, w_lo + adv, idesc, 1);
                        umma_tf32(acc, a_lo + adv, w_hi + adv, idesc, 1);
                    }
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&tmem_full[buf]);
            }
        }
    } else {
        const int q = warp & 3;
        float* stg = stg_all + q * 1024;
        int ui = 0;
        for (int u = blockIdx.x; u < total; u += gridDim.x, ++ui) {
            const int m_tile = u / pg.n_tiles, n0 = (u % pg.n_tiles) * BN, buf = ui & 1;
            auto store_t = [&](float* __restrict__ dst, const float (&x)[32], int n) {
                __syncwarp();
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4)
                    *reinterpret_cast<float4*>(stg + lane * 32 + ((j4 ^ (lane & 7)) << 2)) =
                        make_float4(x[j4 * 4], x[j4 * 4 + 1], x[j4 * 4 + 2], x[j4 * 4 + 3]);
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rr = i * 4 + (lane >> 3), cc = lane & 7;
                    const float4 v4 = *reinterpret_cast<const float4*>(stg + rr * 32 + ((cc ^ (rr & 7)) << 2));
                    const long long orow = (long long)m_tile * TC_BM + q * 32 + rr;
                    if (orow < g.M) *reinterpret_cast<float4*>(dst + orow * g.ldc + n + cc * 4) = v4;
                }
            };
            mbar_wait(&tmem_full[buf], (ui >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + c * 32), v);
                const int n = n0 + c * 32;
                float x[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
                if (g.bias) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] += __ldg(g.bias + n + j);
                }
                if (g.flags & GEMM_RELU) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
                }
                if (g.flags & GEMM_ROUND_TF32) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] = round_tf32(x[j]);
                }
                if (g.flags & GEMM_SPLIT_OUT) {
                    float lo[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) { const float hi = round_tf32(x[j]); lo[j] = round_tf32(x[j] - hi); x[j] = hi; }
                    store_t(g.C_lo, lo, n);
                }
                store_t(g.C, x, n);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[buf])) : "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN));
    }
}

// ---------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// Encoding a tensor map is a pure host computation (~1 us); the decoder issues ~100 GEMMs per sample on
// a handful of stable buffers, so the encoded maps are cached by (base, rows, K, ld, box_rows).
struct MapKey {
    const void* base; int rows, K, ld, box;
    bool operator==(const MapKey& o) const { return base == o.base && rows == o.rows && K == o.K && ld == o.ld && box == o.box; }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = (size_t)k.base;
        h = h * 1000003u ^ (size_t)k.rows; h = h * 1000003u ^ (size_t)k.K;
        h = h * 1000003u ^ (size_t)k.ld; h = h * 1000003u ^ (size_t)k.box;
        return h;
    }
};
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;
static std::mutex g_map_mutex;

static int make_map_2d_uncached(CUtensorMap* m, const float* base, int rows, int K, int ld, int box_rows);
static int make_map_2d(CUtensorMap* m, const float* base, int rows, int K, int ld, int box_rows) {
    MapKey key{base, rows, K, ld, box_rows};
    std::lock_guard<std::mutex> lock(g_map_mutex);
    auto it = g_map_cache.find(key);
    if (it != g_map_cache.end()) { *m = it->second; return 0; }
    int rc = make_map_2d_uncached(m, base, rows, K, ld, box_rows);
    if (rc == 0) {
        if (g_map_cache.size() > 4096) g_map_cache.clear();
        g_map_cache.emplace(key, *m);
    }
    return rc;
}

// 2-D K-contiguous operand [rows, K] with leading dimension ld (floats); box = [32 floats, box_rows]
static int make_map_2d_uncached(CUtensorMap* m, const float* base, int rows, int K, int ld, int box_rows) {
    EncodeTiledFn enc = get_encode();
    MV2D_CHECK_ARG(enc != nullptr, "gemm_tc: cuTensorMapEncodeTiled not available");
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MV2D_CHECK_ARG(r == CUDA_SUCCESS, "gemm_tc: cuTensorMapEncodeTiled(2d) failed with %d", (int)r);
    return 0;
}

int tc_make_map_2d(void* map, const float* base, int rows, int K, int ld, int box_rows) {
    return make_map_2d(reinterpret_cast<CUtensorMap*>(map), base, rows, K, ld, box_rows);
}

// 4-D RoI tokens [n_rois, 7, 7, 256]; box = [32 ch, 7, 7, 1]
static int make_map_tokens(CUtensorMap* m, const float* base, int n_rois) {
    EncodeTiledFn enc = get_encode();
    MV2D_CHECK_ARG(enc != nullptr, "gemm_tc: cuTensorMapEncodeTiled not available");
    cuuint64_t dims[4] = {MV2D_C, MV2D_ROI, MV2D_ROI, (cuuint64_t)n_rois};
    cuuint64_t strides[3] = {MV2D_C * 4, MV2D_ROI * MV2D_C * 4, MV2D_TOK * MV2D_C * 4};
    cuuint32_t box[4] = {TC_BK, MV2D_ROI, MV2D_ROI, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)base, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MV2D_CHECK_ARG(r == CUDA_SUCCESS, "gemm_tc: cuTensorMapEncodeTiled(4d) failed with %d", (int)r);
    return 0;
}

// 4-D feature map [V, h, w, 256]; box = [32 ch, 8 x, 16 y, 1]
static int make_map_fmap(CUtensorMap* m, const float* base, int V, int h, int w) {
    EncodeTiledFn enc = get_encode();
    MV2D_CHECK_ARG(enc != nullptr, "gemm_tc: cuTensorMapEncodeTiled not available");
    cuuint64_t dims[4] = {MV2D_C, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)V};
    cuuint64_t strides[3] = {MV2D_C * 4, (cuuint64_t)w * MV2D_C * 4, (cuuint64_t)h * w * MV2D_C * 4};
    cuuint32_t box[4] = {TC_BK, 8, 16, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)base, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MV2D_CHECK_ARG(r == CUDA_SUCCESS, "gemm_tc: cuTensorMapEncodeTiled(fmap) failed with %d", (int)r);
    return 0;
}

static bool pack5_enabled() {
    static const bool on = []() { const char* e = getenv("MV2D_CONV_PACK5"); return !(e && e[0] == '0'); }();
    return on;
}

template <int BN, int PASSES, int IM2COL, int STAGES, bool RAW = false>
static int launch_tc(const CUtensorMap& a, const CUtensorMap& alo, const CUtensorMap& w, const CUtensorMap& wlo,
                     const TcArgs& g, int m_tiles, int nsplit, cudaStream_t st, const CUtensorMap* a2 = nullptr,
                     const CUtensorMap* a2lo = nullptr) {
    constexpr int NOP = PASSES == 3 ? 2 : 1;
    constexpr size_t smem = (size_t)STAGES * NOP * ((IM2COL == 3 ? 2 : 1) * TC_BM * TC_BK * 4 + BN * TC_BK * 4) + 1024 + 256;
    auto kern = gemm_tc_kernel<BN, PASSES, IM2COL, STAGES, RAW>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("gemm_tc: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr_set = true;
    }
    dim3 grid(g.N / BN, m_tiles, g.groups > 1 ? g.groups : nsplit);
    launch_k(kern, grid, dim3(TC_THREADS), smem, st, a, alo, w, wlo, a2 ? *a2 : a, a2lo ? *a2lo : alo, g);
    MV2D_CHECK_LAUNCH("gemm_tc");
    return 0;
}

// Cluster-multicast variant of the BN=64 3xTF32 GEMM.  Measured on B200 (profiles/README.md): no gain over the
// plain kernel at the decoder's sizes (the A tile is L2-resident either way), so it is opt-in: MV2D_TC_MULTICAST=1.
static bool mc_enabled() {
    static const bool on = []() { const char* e = getenv("MV2D_TC_MULTICAST"); return e && e[0] == '1'; }();
    return on;
}

static int launch_tc_mc(const CUtensorMap& a, const CUtensorMap& alo, const CUtensorMap& w, const CUtensorMap& wlo,
                        const TcArgs& g, int m_tiles, int nsplit, cudaStream_t st) {
    constexpr int STAGES = 4;
    constexpr size_t smem = (size_t)STAGES * 2 * (TC_BM * TC_BK * 4 + 64 * TC_BK * 4) + 256 + 4 * 4096 + 1024;
    auto kern = gemm_tc_mc_kernel<STAGES>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("gemm_tc_mc: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr_set = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(g.N / 64, m_tiles, nsplit); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = TC_MC; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    cfg.attrs = attr; cfg.numAttrs = (pdl_enabled() && cap == cudaStreamCaptureStatusNone) ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a, alo, w, wlo, g);
    note_launch();
    if (e != cudaSuccess) { set_error("gemm_tc_mc: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

static bool persist_enabled() {
    static const bool on = []() { const char* e = getenv("MV2D_TC_PERSIST"); return !(e && e[0] == '0'); }();
    return on;
}

static int launch_tc_persist(const CUtensorMap& a, const CUtensorMap& alo, const CUtensorMap& w, const CUtensorMap& wlo, const TcArgs& g,
                             int m_tiles, int num_sms, cudaStream_t st) {
    constexpr int STAGES = 3;
    constexpr size_t smem = (size_t)STAGES * 2 * (TC_BM * TC_BK * 4 + 128 * TC_BK * 4) + 4 * 4096 + 256 + 1024;
    auto kern = gemm_tc_persist_kernel<STAGES>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("gemm_tc_persist: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr_set = true;
    }
    PgArgs pg{};
    pg.t = g; pg.m_tiles = m_tiles; pg.n_tiles = g.N / 128;
    const int total = pg.m_tiles * pg.n_tiles;
    launch_k(kern, dim3(total < num_sms ? total : num_sms), dim3(TC_THREADS), smem, st, a, alo, w, wlo, pg);
    MV2D_CHECK_LAUNCH("gemm_tc_persist");
    return 0;
}

int launch_gemm_tc(const TcGemm& t, cudaStream_t st) {
    // small-M problems (the decoder, M ~ 300) are latency bound: 64-wide N tiles double the CTA count
    // (a 256-wide, 2-stage im2col variant measured slower than 128-wide / 3 stages; 64-wide tiles for the N = 256 outputs of
    // a batch -- 38 CTAs at 128 wide -- measured slower too: 28.5 vs 25 us per launch, the A tile is re-read by every N tile)
    const int bn = (t.M <= 512 && !t.im2col && t.passes == 3) ? 64 : 128;
    MV2D_CHECK_ARG(t.M > 0 && t.N % bn == 0 && t.K % TC_BK == 0, "gemm_tc: need N%%%d==0 and K%%32==0 (N=%d K=%d)", bn, t.N, t.K);
    MV2D_CHECK_ARG((t.ldc & 3) == 0 && ((uintptr_t)t.C & 15) == 0, "gemm_tc: C must be 16-byte aligned");
    // 3xTF32 with both lo operands null: plain fp32 operands, split inside the kernel (RAW)
    const bool raw = t.passes == 3 && !t.A_lo && !t.W_lo && !t.im2col;
    MV2D_CHECK_ARG(t.passes == 1 || raw || (t.A_lo && t.W_lo), "gemm_tc: 3xTF32 needs both lo operands (or neither: in-kernel split)");
    CUtensorMap a, alo, w, wlo;
    int rc;
    TcArgs g{};
    const int nsplit = t.nsplit > 1 ? t.nsplit : 1;
    const int ngroups = t.groups > 1 ? t.groups : 1;
    MV2D_CHECK_ARG(ngroups == 1 || (nsplit == 1 && !t.im2col && !t.A2 && t.group_rows >= t.M && !(t.flags & GEMM_GATE)),
                   "gemm_tc: a grouped launch excludes split-K, im2col, a second A operand and the gate epilogue");
    g.groups = ngroups; g.group_rows = (int)t.group_rows;
    MV2D_CHECK_ARG((t.K / TC_BK) % nsplit == 0, "gemm_tc: K=%d does not split %d ways into 32-wide blocks", t.K, nsplit);
    MV2D_CHECK_ARG(!(t.flags & GEMM_SPLIT_OUT) || t.C_lo, "gemm_tc: split output needs C_lo");
    g.C = t.C; g.C_lo = t.C_lo; g.ldc = t.ldc; g.bias = t.bias; g.M = t.M; g.N = t.N; g.K = t.K; g.flags = t.flags;
    g.nkb_per_split = t.K / TC_BK / nsplit; g.split_stride = t.split_stride;
    g.gx = t.gx; g.gs = t.gs; g.gfeat = t.gfeat; g.kin = t.kin; g.gs_mod = t.gs_mod;
    g.m_tile_live = t.m_tile_live;
    int m_tiles;
    if (t.im2col == 2) {
        MV2D_CHECK_ARG(t.K == 9 * MV2D_C && t.passes == 3 && t.fm_v > 0 && t.fm_h > 0 && t.fm_w > 0 && t.M == t.fm_v * t.fm_h * t.fm_w,
                       "gemm_tc: feature-map im2col expects K=2304, 3 passes, M = V*h*w");
        g.fm_h = t.fm_h; g.fm_w = t.fm_w;
        m_tiles = t.fm_v * cdiv(t.fm_h, 16) * cdiv(t.fm_w, 8);
        if ((rc = make_map_fmap(&a, t.A, t.fm_v, t.fm_h, t.fm_w))) return rc;
        if ((rc = make_map_fmap(&alo, t.A_lo, t.fm_v, t.fm_h, t.fm_w))) return rc;
    } else if (t.im2col) {
        MV2D_CHECK_ARG(t.K == 9 * MV2D_C && t.passes == 3, "gemm_tc: im2col expects K=2304, 3 passes");
        g.n_rois = t.M / MV2D_TOK;
        m_tiles = pack5_enabled() ? cdiv(g.n_rois, 5) : cdiv(g.n_rois, 2);
        if ((rc = make_map_tokens(&a, t.A, g.n_rois))) return rc;
        if ((rc = make_map_tokens(&alo, t.A_lo, g.n_rois))) return rc;
    } else {
        m_tiles = cdiv(t.M, TC_BM);
        const int a_rows = ngroups > 1 ? (int)((ngroups - 1) * t.group_rows + t.M) : t.M;
        if ((rc = make_map_2d(&a, t.A, a_rows, t.K, t.lda, TC_BM))) return rc;
        if ((rc = make_map_2d(&alo, (t.passes == 3 && !raw) ? t.A_lo : t.A, a_rows, t.K, t.lda, TC_BM))) return rc;
    }
    if ((rc = make_map_2d(&w, t.W, ngroups * t.N, t.K, t.ldw, bn))) return rc;
    if ((rc = make_map_2d(&wlo, (t.passes == 3 && !raw) ? t.W_lo : t.W, ngroups * t.N, t.K, t.ldw, bn))) return rc;
    CUtensorMap a2, a2lo;
    const bool two_a = t.A2 != nullptr && t.n_switch > 0;
    if (two_a) {
        MV2D_CHECK_ARG(t.passes == 3 && !raw && !t.im2col && t.A2_lo && t.n_switch % bn == 0,
                       "gemm_tc: the second A operand needs pre-split 3xTF32 operands and n_switch %% %d == 0", bn);
        g.n_switch = t.n_switch;
        if ((rc = make_map_2d(&a2, t.A2, t.M, t.K, t.lda, TC_BM))) return rc;
        if ((rc = make_map_2d(&a2lo, t.A2_lo, t.M, t.K, t.lda, TC_BM))) return rc;
    }
    if (persist_enabled() && t.passes == 3 && !raw && !t.im2col && ngroups == 1 && !two_a && nsplit == 1 && !t.m_tile_live &&
        !(t.flags & ~(GEMM_RELU | GEMM_ROUND_TF32 | GEMM_SPLIT_OUT)) && t.N % 128 == 0) {
        static int num_sms = 0;
        if (num_sms == 0) {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        }
        // at least two tiles per SM: below that the one-tile-per-CTA kernel spreads the work just as well
        if (m_tiles * (t.N / 128) >= 2 * num_sms) {
            if (bn != 128) {
                if ((rc = make_map_2d(&w, t.W, t.N, t.K, t.ldw, 128))) return rc;
                if ((rc = make_map_2d(&wlo, t.W_lo, t.N, t.K, t.ldw, 128))) return rc;
            }
            return launch_tc_persist(a, alo, w, wlo, g, m_tiles, num_sms, st);
        }
    }
    if (raw && bn == 64) return launch_tc<64, 3, false, 4, true>(a, alo, w, wlo, g, m_tiles, nsplit, st);
    if (raw) return launch_tc<128, 3, false, 3, true>(a, alo, w, wlo, g, m_tiles, nsplit, st);
    if (t.im2col == 2) return launch_tc<128, 3, 2, 3>(a, alo, w, wlo, g, m_tiles, nsplit, st);
    if (t.im2col && pack5_enabled()) return launch_tc<128, 3, 3, 2>(a, alo, w, wlo, g, m_tiles, nsplit, st);
    if (t.im2col && bn == 256) return launch_tc<256, 3, 1, 2>(a, alo, w, wlo, g, m_tiles, nsplit, st);
    if (t.im2col) return launch_tc<128, 3, 1, 3>(a, alo, w, wlo, g, m_tiles, nsplit, st);
    if (t.passes == 3 && bn == 64 && mc_enabled() && (t.N / 64) % TC_MC == 0 && !(t.A2 != nullptr && t.n_switch > 0) && ngroups == 1) {
        // A is loaded in 32-row quarters and multicast across the 4-CTA cluster that shares the M-tile
        if ((rc = make_map_2d(&a, t.A, t.M, t.K, t.lda, 32))) return rc;
        if ((rc = make_map_2d(&alo, t.A_lo, t.M, t.K, t.lda, 32))) return rc;
        return launch_tc_mc(a, alo, w, wlo, g, m_tiles, nsplit, st);
    }
    if (t.passes == 3 && bn == 64) return launch_tc<64, 3, false, 4>(a, alo, w, wlo, g, m_tiles, nsplit, st, two_a ? &a2 : nullptr, two_a ? &a2lo : nullptr);
    if (t.passes == 3) return launch_tc<128, 3, false, 3>(a, alo, w, wlo, g, m_tiles, nsplit, st, two_a ? &a2 : nullptr, two_a ? &a2lo : nullptr);
    return launch_tc<128, 1, false, 3>(a, alo, w, wlo, g, m_tiles, nsplit, st);
}

__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, long long n) {
    pdl_wait();
    pdl_trigger();
    // grid-stride over float4s (16-byte loads / stores), scalar tail
    const long long n4 = n >> 2, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        float4 h, l;
        h.x = round_tf32(v.x); h.y = round_tf32(v.y); h.z = round_tf32(v.z); h.w = round_tf32(v.w);
        l.x = round_tf32(v.x - h.x); l.y = round_tf32(v.y - h.y); l.z = round_tf32(v.z - h.z); l.w = round_tf32(v.w - h.w);
        reinterpret_cast<float4*>(hi)[i] = h;
        reinterpret_cast<float4*>(lo)[i] = l;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const long long i = (n4 << 2) + threadIdx.x;
        const float v = x[i], h = round_tf32(v);
        hi[i] = h;
        lo[i] = round_tf32(v - h);
    }
}

int launch_split_tf32(const float* x, float* hi, float* lo, long long n, cudaStream_t st) {
    if (n <= 0) return 0;
    MV2D_CHECK_ARG((((uintptr_t)x | (uintptr_t)hi | (uintptr_t)lo) & 15) == 0, "split_tf32: pointers must be 16-byte aligned");
    const long long blocks = ((n >> 2) + 255) / 256;
    launch_k(split_tf32_kernel, dim3((unsigned)(blocks < 1 ? 1 : (blocks > 148 * 16 ? 148 * 16 : blocks))), dim3(256), 0, st, x, hi, lo, n);
    MV2D_CHECK_LAUNCH("split_tf32");
    return 0;
}

// Routing used by the stage code: single-pass tcgen05 for big TF32-tolerant problems, FFMA otherwise.
int launch_gemm_tc_or_simt(const GemmArgs& g, cudaStream_t stream) {
    const bool shape_ok = g.N % 128 == 0 && g.K % TC_BK == 0 && g.batch == 1 && g.nsplit == 1 && !(g.flags & GEMM_CLAMP5E3);
    const bool want = (g.flags & GEMM_FORCE_TC) || ((g.flags & GEMM_TF32_OK) && g.M >= 512);
    if (shape_ok && want) {
        TcGemm t{};
        t.A = g.A; t.lda = g.lda; t.W = g.W; t.ldw = g.ldw; t.C = g.C; t.ldc = g.ldc; t.bias = g.bias;
        t.M = g.M; t.N = g.N; t.K = g.K; t.passes = 1; t.im2col = 0;
        t.flags = g.flags & (GEMM_RELU | GEMM_GATE | GEMM_ROUND_TF32);
        t.nsplit = 1;
        t.gx = g.gx; t.gs = g.gs; t.gfeat = g.gfeat; t.kin = g.kin; t.gs_mod = g.gs_mod;
        return launch_gemm_tc(t, stream);
    }
    GemmArgs h = g;
    h.flags &= ~(GEMM_TF32_OK | GEMM_FORCE_TC | GEMM_ROUND_TF32);
    return launch_gemm_simt(h, A_PLAIN, stream);
}

}  // namespace mv2d
