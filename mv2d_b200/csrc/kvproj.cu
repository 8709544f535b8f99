// Two-frame head, key-stationary cross-attention: K_l = (memory + pos) Wk_l^T and V_l = memory Wv_l^T of every feature
// cell for the decoder layers [layer_begin, layer_end) in ONE persistent launch
// (utils/petr_transformer.py:503-508 -> torch.nn.MultiheadAttention in_proj on key = memory + pos, value = memory).
//
// One CTA per SM walks the work units (128-row tile of cells, side K | V, layer) round-robin; a unit is a [128 x 256]
// output tile = one [256,256] weight matrix applied to 128 rows, as error-compensated 3xTF32 tcgen05 MMAs (M = 128,
// N = 256: one instruction stream per unit, the A rows cross the L2 -> SM fabric once per matrix instead of once per
// 128-column half).  Warp roles: warp 0 TMA producer (2 stages of A hi/lo [128 x 32] + W hi/lo [256 x 32], 96 KB each),
// warp 1 MMA issue, warps 2..5 epilogue.  The 512 TMEM columns hold TWO accumulators, so the epilogue of unit i (TMEM ->
// registers -> smem transpose -> coalesced stores) overlaps the main loop of unit i + 1, and the producer runs ahead
// across unit boundaries: neither the prologue nor the epilogue of a tile is exposed, unlike the one-tile-per-CTA
// kernel (gemm_tc.cu) this replaces for the projections.  The six CTAs that work on the same rows at the same time
// (consecutive units = the layers of one tile) share the A tile through L2.
// Tiles no query has a key in (row_tile_live == 0) are skipped by all three roles.
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"
#include "gemm_tc.cuh"
#include "mv2d_internal.h"
#include "tc_ptx.cuh"

namespace mv2d {

namespace {

constexpr int KV_THREADS = 192;
constexpr int KV_BN = 256;
constexpr int KV_STAGES = 2;
constexpr int KV_A_BYTES = TC_BM * TC_BK * 4;          // 16 KB
constexpr int KV_W_BYTES = KV_BN * TC_BK * 4;          // 32 KB
constexpr int KV_STAGE_BYTES = 2 * (KV_A_BYTES + KV_W_BYTES);      // hi + lo: 96 KB
constexpr int KV_NKB = MV2D_C / TC_BK;                 // 8 K blocks
constexpr int KV_STG_BYTES = 4 * 4096;                 // epilogue transpose tiles, one per warp
constexpr int KV_SMEM = KV_STAGES * KV_STAGE_BYTES + KV_STG_BYTES + 256 + 1024;

struct KvArgs {
    int m_tiles, nl, lb, num_rows;       // nl = layers in this launch, lb = first layer
    const uint8_t* live;                 // nullable [m_tiles]
    float* kp; float* vp;                // [L, num_rows, 256]
    long long RC;                        // num_rows * 256
};

__global__ void __launch_bounds__(KV_THREADS, 1)
kv_proj_kernel(const __grid_constant__ CUtensorMap tmKinHi, const __grid_constant__ CUtensorMap tmKinLo,
               const __grid_constant__ CUtensorMap tmMemHi, const __grid_constant__ CUtensorMap tmMemLo,
               const __grid_constant__ CUtensorMap tmWhi, const __grid_constant__ CUtensorMap tmWlo, KvArgs g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* stg_all = reinterpret_cast<float*>(smem + KV_STAGES * KV_STAGE_BYTES);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + KV_STAGES * KV_STAGE_BYTES + KV_STG_BYTES);
    uint64_t* empty_bar = full_bar + KV_STAGES;
    uint64_t* tmem_full = empty_bar + KV_STAGES;       // [2]
    uint64_t* tmem_empty = tmem_full + 2;              // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per_tile = 2 * g.nl, total = g.m_tiles * per_tile;

    if (warp == 0 && lane == 0) {
        tmap_prefetch(&tmKinHi); tmap_prefetch(&tmKinLo); tmap_prefetch(&tmMemHi); tmap_prefetch(&tmMemLo);
        tmap_prefetch(&tmWhi); tmap_prefetch(&tmWlo);
        for (int s = 0; s < KV_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * KV_BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();
    pdl_trigger();

    if (warp == 0) {
        // ================= TMA producer: runs ahead across unit boundaries =================
        if (lane == 0) {
            int it = 0;
            for (int u = blockIdx.x; u < total; u += gridDim.x) {
                const int m = u / per_tile, j = u % per_tile, side = j / g.nl, l = g.lb + j % g.nl;
                if (g.live && g.live[m] == 0) continue;
                const CUtensorMap* ah = side ? &tmMemHi : &tmKinHi;
                const CUtensorMap* al = side ? &tmMemLo : &tmKinLo;
                const int wrow = (l * 2 + side) * MV2D_C;
                for (int kb = 0; kb < KV_NKB; ++kb, ++it) {
                    const int s = it % KV_STAGES, ph = (it / KV_STAGES) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* st = smem + s * KV_STAGE_BYTES;
                    mbar_expect_tx(&full_bar[s], KV_STAGE_BYTES);
                    tma_load_2d(ah, &full_bar[s], st, kb * TC_BK, m * TC_BM);
                    tma_load_2d(al, &full_bar[s], st + KV_A_BYTES, kb * TC_BK, m * TC_BM);
                    tma_load_2d(&tmWhi, &full_bar[s], st + 2 * KV_A_BYTES, kb * TC_BK, wrow);
                    tma_load_2d(&tmWlo, &full_bar[s], st + 2 * KV_A_BYTES + KV_W_BYTES, kb * TC_BK, wrow);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer: accumulator ui & 1 =================
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(KV_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        if (lane == 0) {
            int it = 0, ui = 0;
            for (int u = blockIdx.x; u < total; u += gridDim.x) {
                if (g.live && g.live[u / per_tile] == 0) continue;
                const int buf = ui & 1;
                mbar_wait(&tmem_empty[buf], ((ui >> 1) & 1) ^ 1);     // the epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = tmem_base + (uint32_t)(buf * KV_BN);
                for (int kb = 0; kb < KV_NKB; ++kb, ++it) {
                    const int s = it % KV_STAGES, ph = (it / KV_STAGES) & 1;
                    mbar_wait(&full_bar[s], ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sa = smem_u32(smem + s * KV_STAGE_BYTES);
                    const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + KV_A_BYTES);
                    const uint64_t w_hi = make_desc(sa + 2 * KV_A_BYTES), w_lo = make_desc(sa + 2 * KV_A_BYTES + KV_W_BYTES);
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
                ++ui;
            }
        }
    } else {
        // ================= epilogue: TMEM -> registers -> smem transpose -> 4 rows x 128 B per store instruction =================
        const int q = warp & 3;
        float* stg = stg_all + q * 1024;
        int ui = 0;
        for (int u = blockIdx.x; u < total; u += gridDim.x) {
            const int m = u / per_tile, j = u % per_tile, side = j / g.nl, l = g.lb + j % g.nl;
            if (g.live && g.live[m] == 0) continue;
            const int buf = ui & 1;
            float* __restrict__ dst = (side ? g.vp : g.kp) + (long long)l * g.RC;
            mbar_wait(&tmem_full[buf], (ui >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int c = 0; c < KV_BN / 32; ++c) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * KV_BN + c * 32), v);
                __syncwarp();
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4)
                    *reinterpret_cast<float4*>(stg + lane * 32 + ((j4 ^ (lane & 7)) << 2)) =
                        make_float4(__uint_as_float(v[j4 * 4]), __uint_as_float(v[j4 * 4 + 1]), __uint_as_float(v[j4 * 4 + 2]),
                                    __uint_as_float(v[j4 * 4 + 3]));
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rr = i * 4 + (lane >> 3), cc = lane & 7;
                    const float4 v4 = *reinterpret_cast<const float4*>(stg + rr * 32 + ((cc ^ (rr & 7)) << 2));
                    const long long row = (long long)m * TC_BM + q * 32 + rr;
                    if (row < g.num_rows) *reinterpret_cast<float4*>(dst + row * MV2D_C + c * 32 + cc * 4) = v4;
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[buf])) : "memory");
            ++ui;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * KV_BN));
    }
}

}  // namespace

// The persistent kernel needs the 2 L weight matrices as ONE [L*2*256, 256] operand (rows (l*2 + side)*256 ..), hi and
// lo: mv2d_pack_weights lays xa_k_w / xa_v_w out that way.  Returns 1 when the layers' pointers are stacked like that.
static bool kv_weights_stacked(const Mv2dLayerWeights* layers, int L) {
    const float *hi = layers[0].xa_k_w, *lo = layers[0].xa_k_w_lo;
    if (!hi || !lo) return false;
    const long long CC = (long long)MV2D_C * MV2D_C;
    for (int l = 0; l < L; ++l)
        if (layers[l].xa_k_w != hi + (2 * l) * CC || layers[l].xa_v_w != hi + (2 * l + 1) * CC ||
            layers[l].xa_k_w_lo != lo + (2 * l) * CC || layers[l].xa_v_w_lo != lo + (2 * l + 1) * CC)
            return false;
    return true;
}

bool kv_persistent_usable(const Mv2dKvParams& p) {
    static const bool on = []() { const char* e = getenv("MV2D_KV_PERSISTENT"); return !(e && e[0] == '0'); }();
    return on && p.kin_lo && p.mem_lo && kv_weights_stacked(p.layers, p.L);
}

int run_kv_project_persistent(const Mv2dKvParams& p, cudaStream_t st) {
    const int le = p.layer_end > 0 ? p.layer_end : p.L;
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    CUtensorMap kh, kl, mh, ml, wh, wl;
    int rc;
    if ((rc = tc_make_map_2d(&kh, p.kin_hi, p.num_rows, MV2D_C, MV2D_C, TC_BM))) return rc;
    if ((rc = tc_make_map_2d(&kl, p.kin_lo, p.num_rows, MV2D_C, MV2D_C, TC_BM))) return rc;
    if ((rc = tc_make_map_2d(&mh, p.mem_hi, p.num_rows, MV2D_C, MV2D_C, TC_BM))) return rc;
    if ((rc = tc_make_map_2d(&ml, p.mem_lo, p.num_rows, MV2D_C, MV2D_C, TC_BM))) return rc;
    if ((rc = tc_make_map_2d(&wh, p.layers[0].xa_k_w, p.L * 2 * MV2D_C, MV2D_C, MV2D_C, KV_BN))) return rc;
    if ((rc = tc_make_map_2d(&wl, p.layers[0].xa_k_w_lo, p.L * 2 * MV2D_C, MV2D_C, MV2D_C, KV_BN))) return rc;
    KvArgs g{};
    g.m_tiles = cdiv(p.num_rows, TC_BM); g.nl = le - p.layer_begin; g.lb = p.layer_begin; g.num_rows = p.num_rows;
    g.live = p.row_tile_live; g.kp = p.kp; g.vp = p.vp; g.RC = (long long)p.num_rows * MV2D_C;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kv_proj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KV_SMEM);
        if (e != cudaSuccess) { set_error("kv_project: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr_set = true;
    }
    const int total = g.m_tiles * 2 * g.nl;
    launch_k(kv_proj_kernel, dim3(total < num_sms ? total : num_sms), dim3(KV_THREADS), (size_t)KV_SMEM, st, kh, kl, mh, ml, wh, wl, g);
    MV2D_CHECK_LAUNCH("kv_project");
    return 0;
}

}  // namespace mv2d
