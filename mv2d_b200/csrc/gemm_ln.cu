// Fused  out = LayerNorm( A[M,K] . W[256,K]^T + bias + residual )  for the three "projection -> add -> norm" steps of a
// decoder layer (self-attention out_proj, cross-attention out_proj, FFN layer 2; utils/petr_transformer.py:240-311 over
// mmcv BaseTransformerLayer's `norm` steps).
//
// One thread-block CLUSTER per 128-row tile.  The CL CTAs of a cluster split K: each streams its K / CL slice of A and of
// the whole [256, K] weight (3xTF32: hi and lo copies, TMA, 128B swizzle) and accumulates a full-width [128 x 256]
// partial tile in TMEM -- a 256-column tile holds whole rows, which is what LayerNorm needs.  The partial tiles are then
// parked in each CTA's own shared memory (the pipeline stages are free by then), the cluster synchronises, and CTA r
// finishes rows [r * 128 / CL, (r + 1) * 128 / CL): it sums the CL partials of those rows over distributed shared memory
// in rank order (deterministic), adds bias + residual, normalises, and writes the row in every form the next kernels
// read (plain, + query_pos, second LayerNorm, TF32 hi / lo splits) -- see ln_tail (ln.cuh).
// Replaces a split-K GEMM launch that wrote CL partial tiles to global memory plus the ln_kernel launch that re-read
// them.
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"
#include "gemm_tc.cuh"
#include "ln.cuh"
#include "tc_ptx.cuh"

namespace mv2d {

namespace {

constexpr int GL_THREADS = 192;
constexpr int GL_BN = 256;
constexpr int GL_STAGES = 2;
constexpr int GL_A_BYTES = TC_BM * TC_BK * 4;          // 16 KB
constexpr int GL_W_BYTES = GL_BN * TC_BK * 4;          // 32 KB
constexpr int GL_STAGE_BYTES = 2 * (GL_A_BYTES + GL_W_BYTES);   // hi + lo of both operands: 96 KB
constexpr int GL_SMEM = GL_STAGES * GL_STAGE_BYTES + 256 + 1024;

struct GlArgs {
    int M, nkb;            // rows; 32-wide K blocks per CTA
    LnArgs ln;             // bias / residual / gammas / outputs (partial, nsplit unused)
};

__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t local_addr, uint32_t rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(remote) : "memory");
    return v;
}

template <int CL>
__global__ void __launch_bounds__(GL_THREADS, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
               const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmWlo, GlArgs g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + GL_STAGES * GL_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + GL_STAGES;
    uint64_t* tmem_full_bar = empty_bar + GL_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tile = blockIdx.y;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int nkb = g.nkb, kb0 = (int)rank * g.nkb;

    if (warp == 0 && lane == 0) {
        tmap_prefetch(&tmA); tmap_prefetch(&tmW); tmap_prefetch(&tmAlo); tmap_prefetch(&tmWlo);
        for (int s = 0; s < GL_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(GL_BN));
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
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % GL_STAGES, ph = (kb / GL_STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t* st = smem + s * GL_STAGE_BYTES;
                mbar_expect_tx(&full_bar[s], GL_STAGE_BYTES);
                const int k = (kb0 + kb) * TC_BK;
                tma_load_2d(&tmA, &full_bar[s], st, k, m_tile * TC_BM);
                tma_load_2d(&tmAlo, &full_bar[s], st + GL_A_BYTES, k, m_tile * TC_BM);
                tma_load_2d(&tmW, &full_bar[s], st + 2 * GL_A_BYTES, k, 0);
                tma_load_2d(&tmWlo, &full_bar[s], st + 2 * GL_A_BYTES + GL_W_BYTES, k, 0);
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(GL_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % GL_STAGES, ph = (kb / GL_STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem_u32(smem + s * GL_STAGE_BYTES);
                const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + GL_A_BYTES);
                const uint64_t w_hi = make_desc(sa + 2 * GL_A_BYTES), w_lo = make_desc(sa + 2 * GL_A_BYTES + GL_W_BYTES);
#pragma unroll
                for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
                    const uint64_t adv = (uint64_t)((k * TC_UMMA_K * 4) >> 4);
                    umma_tf32(tmem_base, a_hi + adv, w_hi + adv, idesc, (kb | k) != 0);
                    umma_tf32(tmem_base, a_hi + adv, w_lo + adv, idesc, 1);
                    umma_tf32(tmem_base, a_lo + adv, w_hi + adv, idesc, 1);
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(tmem_full_bar);
        }
    } else {
        // ---- park the partial tile in shared memory: row r, 32-column chunk c, float4 j4 at ((r*8 + c)*8 + (j4 ^ (r & 7)))
        mbar_wait(tmem_full_bar, 0);                    // every MMA has retired: the pipeline stages are free
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3, r = q * 32 + lane;
        float4* tile = reinterpret_cast<float4*>(smem);
#pragma unroll 1
        for (int c = 0; c < GL_BN / 32; ++c) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4)
                tile[(r * 8 + c) * 8 + (j4 ^ (r & 7))] = make_float4(__uint_as_float(v[j4 * 4]), __uint_as_float(v[j4 * 4 + 1]),
                                                                     __uint_as_float(v[j4 * 4 + 2]), __uint_as_float(v[j4 * 4 + 3]));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                                  // all CL partial tiles are in place (release / acquire)
    if (warp >= 2) {
        // ---- CTA `rank` finishes rows [rank * RPC, (rank + 1) * RPC) of the tile: one warp per row
        constexpr int RPC = TC_BM / CL;
        const uint32_t tile_addr = smem_u32(smem);
        for (int rr = warp - 2; rr < RPC; rr += 4) {
            const int r = (int)rank * RPC + rr, row = m_tile * TC_BM + r;
            if (row >= g.M) break;
            float4 pt[2][CL];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int col = i * 128 + lane * 4, c = col >> 5, j4 = (col & 31) >> 2;
                const uint32_t addr = tile_addr + (uint32_t)(((r * 8 + c) * 8 + (j4 ^ (r & 7))) * 16);
#pragma unroll
                for (int k = 0; k < CL; ++k) pt[i][k] = ld_dsmem_f4(addr, (uint32_t)k);
            }
            float v[8];
            const long long o = (long long)row * MV2D_C;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int col = i * 128 + lane * 4;
                float4 s = pt[i][0];
#pragma unroll
                for (int k = 1; k < CL; ++k) { s.x += pt[i][k].x; s.y += pt[i][k].y; s.z += pt[i][k].z; s.w += pt[i][k].w; }
                if (g.ln.bias) {
                    const float4 t = __ldg(reinterpret_cast<const float4*>(g.ln.bias + col));
                    s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
                }
                if (g.ln.residual) {
                    const float4 t = *reinterpret_cast<const float4*>(g.ln.residual + o + col);
                    s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
                }
                v[i * 4 + 0] = s.x; v[i * 4 + 1] = s.y; v[i * 4 + 2] = s.z; v[i * 4 + 3] = s.w;
            }
            ln_tail(g.ln, row, 0, lane, v);
        }
    }
    cluster_sync_all();                                  // nobody leaves while a peer still reads its tile
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(GL_BN));
    }
}

template <int CL>
int launch_cl(const CUtensorMap& a, const CUtensorMap& alo, const CUtensorMap& w, const CUtensorMap& wlo, const GlArgs& g, int m_tiles,
              cudaStream_t st) {
    auto kern = gemm_ln_kernel<CL>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GL_SMEM);
        if (e != cudaSuccess) { set_error("gemm_ln: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr_set = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CL, m_tiles, 1); cfg.blockDim = dim3(GL_THREADS); cfg.dynamicSmemBytes = GL_SMEM; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    cfg.attrs = attr; cfg.numAttrs = (pdl_enabled() && cap == cudaStreamCaptureStatusNone) ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a, alo, w, wlo, g);
    note_launch();
    if (e != cudaSuccess) { set_error("gemm_ln: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

}  // namespace

bool gemm_ln_enabled() {
    // opt-in (MV2D_GEMM_LN=1): measured slower than the split-K GEMM + ln_kernel pair at both M ~ 300 and M = 2400 --
    // a full-row tile puts the MMAs and the operand stream of a 128-row tile on at most 8 SMs
    static const bool on = []() { const char* e = getenv("MV2D_GEMM_LN"); return e && e[0] == '1'; }();
    return on;
}

// A [M,K] and W [256,K] as TF32 hi / lo pairs (K-contiguous, leading dimensions lda / ldw); `ln` carries bias, residual,
// gammas and the outputs (its partial / nsplit / group fields are ignored).  cluster = 0 picks the split from K and M.
int launch_gemm_ln(const float* A_hi, const float* A_lo, int lda, const float* W_hi, const float* W_lo, int ldw, int M, int K,
                   const LnArgs& ln, int cluster, cudaStream_t st) {
    MV2D_CHECK_ARG(M > 0 && K % TC_BK == 0 && A_hi && A_lo && W_hi && W_lo && ln.out && ln.gamma && ln.beta, "gemm_ln: bad arguments");
    MV2D_CHECK_ARG(ln.rows_per_group == 0 && !ln.relu && !ln.bcast_in, "gemm_ln: grouped / ReLU epilogues stay with ln_kernel");
    const int nkb_total = K / TC_BK, m_tiles = cdiv(M, TC_BM);
    static const int env_cl = []() { const char* e = getenv("MV2D_GEMM_LN_CL"); return e ? atoi(e) : 0; }();
    int cl = cluster > 0 ? cluster : env_cl;
    if (cl <= 0) {
        // enough CTAs to cover the SMs once (clusters of 8 land two per GPC: ~16 of them are resident at a time)
        cl = 8;
        while (cl > 2 && (m_tiles * cl > 128 || nkb_total / cl < 4)) cl >>= 1;
    }
    while (cl > 1 && nkb_total % cl) cl >>= 1;
    MV2D_CHECK_ARG(cl == 2 || cl == 4 || cl == 8, "gemm_ln: K=%d does not split over a cluster of 2, 4 or 8", K);
    CUtensorMap a, alo, w, wlo;
    int rc;
    if ((rc = tc_make_map_2d(&a, A_hi, M, K, lda, TC_BM))) return rc;
    if ((rc = tc_make_map_2d(&alo, A_lo, M, K, lda, TC_BM))) return rc;
    if ((rc = tc_make_map_2d(&w, W_hi, GL_BN, K, ldw, GL_BN))) return rc;
    if ((rc = tc_make_map_2d(&wlo, W_lo, GL_BN, K, ldw, GL_BN))) return rc;
    GlArgs g{};
    g.M = M; g.nkb = nkb_total / cl; g.ln = ln; g.ln.rows = M;
    if (cl == 8) return launch_cl<8>(a, alo, w, wlo, g, m_tiles, st);
    if (cl == 4) return launch_cl<4>(a, alo, w, wlo, g, m_tiles, st);
    return launch_cl<2>(a, alo, w, wlo, g, m_tiles, st);
}

}  // namespace mv2d
