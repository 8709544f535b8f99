// Fused two-layer MLP on the 5th-gen tensor cores:  out[M,256] = epi( act(A[M,K0] W0[H,K0]^T + b0) W2[256,H]^T + b2 )
// with the H-wide hidden activation never leaving the SM (kernel K1).
//   reference: the three 1x1-conv MLPs of PE.forward, utils/pe.py:64-77 (position_encoder 192 -> 1024 -> 256),
//   :78-82 (adapt_pos3d 384 -> 1024 -> 256), :44-48 (SELayer 256 -> 256 -> 256 gate) and the combine :158-166.
//
// One CTA per 128-row tile of A.  The hidden dimension is walked in chunks of 128:
//   G1(c)  acc_h[c & 1] (TMEM, 128 columns) = A . W0[128c .. 128c+127, :]^T        (K0 / 8 tcgen05.mma, N = 128)
//   E(c)   eight epilogue warps: tcgen05.ld -> + b0 -> ReLU -> round to TF32 -> shared memory, written directly in
//          the K-major 128-byte-swizzled layout the next MMA reads its A operand in (Hc, 4 k-blocks = 64 KB, double buffered)
//   G2(c)  acc_out (TMEM, 256 columns) += Hc . W2[:, 128c .. 128c+127]^T             (16 tcgen05.mma, N = 256)
// issued in the order G1(0) G1(1) G2(0) G1(2) G2(1) ... so the tensor pipe works on chunk c+1 while the epilogue
// warps convert chunk c.  All operands stream through one ring of 32 KB slots (a G1 item = A k-block + W0 k-block,
// a G2 item = a [256 x 32] k-block of W2); TMEM = 2 x 128 + 256 = all 512 columns, one CTA per SM.
// Single-pass TF32 (operands pre-rounded by their producers), as the PE MLPs of the unfused path.
// Final epilogue: + b2, then either a plain store or the SE gate / combine  pe = gx * sigmoid(acc) + gs ; kin = pe + gfeat.
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "tc_ptx.cuh"
#include "mlp2.cuh"

namespace mv2d {

static constexpr int M2_THREADS = 320;                   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quadrant)
static constexpr int M2_SLOT_BYTES = 32 * 1024;
static constexpr int M2_NSLOT = 3;
static constexpr int M2_HC_BYTES = 4 * 16 * 1024;        // hidden chunk as an A operand: 4 k-blocks of [128 x 32]; double buffered
static constexpr int M2_SMEM_BYTES = M2_NSLOT * M2_SLOT_BYTES + 2 * M2_HC_BYTES + 1024 /*alignment*/ + 512 /*barriers*/;

struct Mlp2Args {
    int M, K0, H;
    const float* b0; const float* b2;
    float* out; int ldo;
    int gate;                                   // 1: SE gate / combine epilogue
    const float* gx; const float* gs; int gs_mod; const float* gfeat; float* kin;
    int round_out;                              // 1: round the stored result to TF32
};

// CL = CTAs per cluster.  The CL CTAs of a cluster work on adjacent row tiles and need the same W0 / W2 blocks at the same
// step, so each loads 1/CL of every weight block and TMA-multicasts it into all CL shared memories: the weights cross
// the L2 -> SM fabric once per cluster instead of once per CTA (the kernel is bound by exactly that traffic: 2.5 MB per
// tile against 117 MFLOP).  A ring slot is refilled by the peers as well, so its "empty" barrier collects one
// tcgen05.commit arrival from each CTA of the cluster.
template <int CL>
__global__ void __launch_bounds__(M2_THREADS, 1)
mlp2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW0,
            const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmOut,
            const __grid_constant__ CUtensorMap tmKin, const __grid_constant__ CUtensorMap tmGx,
            const __grid_constant__ CUtensorMap tmGs, const __grid_constant__ CUtensorMap tmGf, Mlp2Args g) {
    extern __shared__ __align__(1024) uint8_t m2_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(m2_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* ring = smem;                                   // [NSLOT][32 KB]
    uint8_t* hc = smem + M2_NSLOT * M2_SLOT_BYTES;          // [2][4][16 KB]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(hc + 2 * M2_HC_BYTES);
    uint64_t* empty_bar = full_bar + M2_NSLOT;
    uint64_t* hacc_full = empty_bar + M2_NSLOT;             // [2] G1 of a chunk retired: acc_h[b] may be read
    uint64_t* hacc_empty = hacc_full + 2;                   // [2] the epilogue is done with acc_h[b]
    uint64_t* hsm_full = hacc_empty + 2;                    // [2] Hc[b] written (256 arrivals)
    uint64_t* hsm_empty = hsm_full + 2;                     // [2] G2 of the chunk retired: Hc[b] may be overwritten
    uint64_t* out_full = hsm_empty + 2;
    uint64_t* gate_full = out_full + 1;                     // [2] gate operands of a column group landed
    uint64_t* gate_empty = gate_full + 2;                   // [2] ... and were read by the 128 epilogue threads
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gate_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tile = blockIdx.x;
    const int nkb0 = g.K0 / TC_BK, nchunk = g.H / 128;
    uint32_t rank = 0;
    if (CL > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    constexpr uint16_t cmask = (uint16_t)((1u << CL) - 1);

    if (warp == 0 && lane == 0) {
        tmap_prefetch(&tmA); tmap_prefetch(&tmW0); tmap_prefetch(&tmW2);
        for (int s = 0; s < M2_NSLOT; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CL); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&hacc_full[b], 1); mbar_init(&hacc_empty[b], 256);
            mbar_init(&hsm_full[b], 256); mbar_init(&hsm_empty[b], 1);
        }
        mbar_init(out_full, 1);
        for (int b = 0; b < 2; ++b) { mbar_init(&gate_full[b], 1); mbar_init(&gate_empty[b], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CL > 1) cluster_sync_all();            // every CTA's barriers exist before any peer signals them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();
    pdl_trigger();

    // The schedule both the producer and the MMA warp walk: step = (kind, chunk); kind 0 = G1, 1 = G2.
    // G1(0), then for c = 1 .. nchunk-1: G1(c), G2(c-1); finally G2(nchunk-1).
    const int nsteps = 2 * nchunk;
    auto step_of = [&](int i, int& kind, int& c) {
        if (i == 0) { kind = 0; c = 0; }
        else if (i == nsteps - 1) { kind = 1; c = nchunk - 1; }
        else if (i & 1) { kind = 0; c = (i + 1) >> 1; }
        else { kind = 1; c = (i >> 1) - 1; }
    };

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int it = 0;
            for (int i = 0; i < nsteps; ++i) {
                int kind, c;
                step_of(i, kind, c);
                const int nitems = kind == 0 ? nkb0 : 4;
                for (int kb = 0; kb < nitems; ++kb, ++it) {
                    const int s = it % M2_NSLOT, ph = (it / M2_NSLOT) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* slot = ring + s * M2_SLOT_BYTES;
                    mbar_expect_tx(&full_bar[s], M2_SLOT_BYTES);
                    if (kind == 0) {
                        tma_load_2d(&tmA, &full_bar[s], slot, kb * TC_BK, m_tile * TC_BM);
                        if (CL == 1) tma_load_2d(&tmW0, &full_bar[s], slot + 16 * 1024, kb * TC_BK, c * 128);
                        else tma_load_2d_mc(&tmW0, &full_bar[s], slot + 16 * 1024 + rank * (16 * 1024 / CL), kb * TC_BK,
                                            c * 128 + rank * (128 / CL), cmask);
                    } else {
                        if (CL == 1) tma_load_2d(&tmW2, &full_bar[s], slot, c * 128 + kb * TC_BK, 0);        // [256 rows x 32] of W2
                        else tma_load_2d_mc(&tmW2, &full_bar[s], slot + rank * (32 * 1024 / CL), c * 128 + kb * TC_BK, rank * (256 / CL), cmask);
                    }
                }
            }
            if (g.gate) {
                // gate / combine operands of the output tile, one 32-column group at a time, double buffered in the (by
                // then idle) ring: gx, gs[, gfeat] boxes of [128 rows x 32 columns]
                mbar_wait(out_full, 0);                       // every MMA has retired: nobody reads the ring any more
                const int nbox = g.kin ? 3 : 2;
                const int gs_row = g.gs_mod > 0 ? (m_tile * TC_BM) % g.gs_mod : m_tile * TC_BM;
                for (int cgp = 0; cgp < 8; ++cgp) {
                    const int b = cgp & 1;
                    if (cgp >= 2) mbar_wait(&gate_empty[b], ((cgp >> 1) - 1) & 1);
                    uint8_t* dst = ring + b * 3 * 16 * 1024;
                    mbar_expect_tx(&gate_full[b], nbox * 16 * 1024);
                    tma_load_2d(&tmGx, &gate_full[b], dst, cgp * 32, m_tile * TC_BM);
                    tma_load_2d(&tmGs, &gate_full[b], dst + 16 * 1024, cgp * 32, gs_row);
                    if (g.kin) tma_load_2d(&tmGf, &gate_full[b], dst + 32 * 1024, cgp * 32, m_tile * TC_BM);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        constexpr uint32_t idesc128 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        constexpr uint32_t idesc256 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        if (lane == 0) {
            int it = 0;
            for (int i = 0; i < nsteps; ++i) {
                int kind, c;
                step_of(i, kind, c);
                if (kind == 0) {
                    const int b = c & 1;
                    // acc_h[b] was last read by the epilogue of chunk c - 2
                    if (c >= 2) { mbar_wait(&hacc_empty[b], ((c >> 1) - 1) & 1); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
                    for (int kb = 0; kb < nkb0; ++kb, ++it) {
                        const int s = it % M2_NSLOT, ph = (it / M2_NSLOT) & 1;
                        mbar_wait(&full_bar[s], ph);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t sa = smem_u32(ring + s * M2_SLOT_BYTES);
                        const uint64_t da = make_desc(sa), dw = make_desc(sa + 16 * 1024);
#pragma unroll
                        for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
                            const uint64_t adv = (uint64_t)((k * TC_UMMA_K * 4) >> 4);
                            umma_tf32(tmem_base + b * 128, da + adv, dw + adv, idesc128, (kb | k) != 0);
                        }
                        if (CL == 1) umma_commit(&empty_bar[s]); else umma_commit_mc(&empty_bar[s], cmask);
                    }
                    umma_commit(&hacc_full[b]);
                } else {
                    mbar_wait(&hsm_full[c & 1], (c >> 1) & 1);       // Hc of chunk c is in shared memory
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    for (int kb = 0; kb < 4; ++kb, ++it) {
                        const int s = it % M2_NSLOT, ph = (it / M2_NSLOT) & 1;
                        mbar_wait(&full_bar[s], ph);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint64_t da = make_desc(smem_u32(hc + (c & 1) * M2_HC_BYTES + kb * 16 * 1024));
                        const uint64_t dw = make_desc(smem_u32(ring + s * M2_SLOT_BYTES));
#pragma unroll
                        for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
                            const uint64_t adv = (uint64_t)((k * TC_UMMA_K * 4) >> 4);
                            umma_tf32(tmem_base + 256, da + adv, dw + adv, idesc256, (c | kb | k) != 0);
                        }
                        if (CL == 1) umma_commit(&empty_bar[s]); else umma_commit_mc(&empty_bar[s], cmask);
                    }
                    umma_commit(&hsm_empty[c & 1]);
                }
            }
            umma_commit(out_full);
        }
    } else {
        // ================= epilogue warps: hidden chunks, then the output tile =================
        const int q = warp & 3;                           // TMEM lane quadrant of this warp
        const int hf = (warp - 2) >> 2;                   // which half of the columns this warp of the quadrant's pair takes
        const int r = q * 32 + lane;                      // tile row of this thread
        for (int c = 0; c < nchunk; ++c) {
            const int b = c & 1;
            mbar_wait(&hacc_full[b], (c >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (c >= 2) mbar_wait(&hsm_empty[b], ((c >> 1) - 1) & 1);     // G2 of chunk c - 2 no longer reads Hc[b]
#pragma unroll 1
            for (int kk = 0; kk < 2; ++kk) {
                const int kb = hf * 2 + kk;
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * 128 + kb * 32), v);
                const float4* bias = reinterpret_cast<const float4*>(g.b0 + c * 128 + kb * 32);
                uint8_t* row = hc + b * M2_HC_BYTES + kb * 16 * 1024 + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 bb = __ldg(bias + j4);
                    float4 x;
                    x.x = round_tf32(fmaxf(__uint_as_float(v[j4 * 4 + 0]) + bb.x, 0.f));
                    x.y = round_tf32(fmaxf(__uint_as_float(v[j4 * 4 + 1]) + bb.y, 0.f));
                    x.z = round_tf32(fmaxf(__uint_as_float(v[j4 * 4 + 2]) + bb.z, 0.f));
                    x.w = round_tf32(fmaxf(__uint_as_float(v[j4 * 4 + 3]) + bb.w, 0.f));
                    *reinterpret_cast<float4*>(row + ((j4 ^ (r & 7)) << 4)) = x;
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&hacc_empty[b])) : "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy stores -> visible to the MMA
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&hsm_full[b])) : "memory");
        }
        // ---- output tile
        mbar_wait(out_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // staging for the TMA stores: per warp two [32 rows x 32 cols] boxes (out, kin), 128-byte swizzled, behind the
        // two gate-operand buffers of the ring
        uint8_t* stg_out = hc + (warp - 2) * 8192;           // the Hc buffers are idle as well
        uint8_t* stg_kin = stg_out + 4096;
        const int row0 = m_tile * TC_BM + q * 32;             // first output row of this warp
        auto row_read = [&](const uint8_t* box, float (&x)[32]) {     // this thread's row of a [128 x 32] swizzled box
            const uint8_t* rp = box + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
                const float4 v4 = *reinterpret_cast<const float4*>(rp + ((j4 ^ (r & 7)) << 4));
                x[j4 * 4] = v4.x; x[j4 * 4 + 1] = v4.y; x[j4 * 4 + 2] = v4.z; x[j4 * 4 + 3] = v4.w;
            }
        };
        auto row_write = [&](uint8_t* box32, const float (&x)[32]) {  // row `lane` of a [32 x 32] swizzled box
            uint8_t* rp = box32 + (lane >> 3) * 1024 + (lane & 7) * 128;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4)
                *reinterpret_cast<float4*>(rp + ((j4 ^ (lane & 7)) << 4)) = make_float4(x[j4 * 4], x[j4 * 4 + 1], x[j4 * 4 + 2], x[j4 * 4 + 3]);
        };
        // the two warps of a quadrant take alternate 32-column groups (and so alternate gate-operand buffers)
#pragma unroll 1
        for (int cgp = hf; cgp < 8; cgp += 2) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(256 + cgp * 32), v);
            const int n = cgp * 32;
            float x[32], t[32];
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(g.b2 + n) + j4);
                x[j4 * 4 + 0] = __uint_as_float(v[j4 * 4 + 0]) + bb.x; x[j4 * 4 + 1] = __uint_as_float(v[j4 * 4 + 1]) + bb.y;
                x[j4 * 4 + 2] = __uint_as_float(v[j4 * 4 + 2]) + bb.z; x[j4 * 4 + 3] = __uint_as_float(v[j4 * 4 + 3]) + bb.w;
            }
            if (g.gate) {
                const int b = cgp & 1;
                const uint8_t* src = ring + b * 3 * 16 * 1024;
                mbar_wait(&gate_full[b], (cgp >> 1) & 1);
                row_read(src, t);
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = t[j] * __fdividef(1.f, 1.f + __expf(-x[j]));     // fast sigmoid: ~1e-6 relative, the stage is judged at TF32 accuracy
                row_read(src + 16 * 1024, t);
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] += t[j];
                if (g.kin) {
                    row_read(src + 32 * 1024, t);
#pragma unroll
                    for (int j = 0; j < 32; ++j) t[j] += x[j];
                }
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&gate_empty[b])) : "memory");
            }
            if (g.round_out) {
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = round_tf32(x[j]);
            }
            // the previous group's bulk stores must have read the staging boxes before they are overwritten
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
            row_write(stg_out, x);
            if (g.gate && g.kin) row_write(stg_kin, t);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                tma_store_2d(&tmOut, stg_out, n, row0);
                if (g.gate && g.kin) tma_store_2d(&tmKin, stg_kin, n, row0);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CL > 1) cluster_sync_all();            // nobody leaves while a peer can still write its shared memory / barriers
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    }
}

int launch_mlp2(const Mlp2& m, cudaStream_t st) {
    MV2D_CHECK_ARG(m.M > 0 && m.K0 % TC_BK == 0 && m.K0 >= TC_BK && m.H % 128 == 0 && m.H >= 128,
                   "mlp2: need K0 %% 32 == 0 and H %% 128 == 0 (K0=%d H=%d)", m.K0, m.H);
    MV2D_CHECK_ARG(m.A && m.W0 && m.W2 && m.b0 && m.b2 && m.out, "mlp2: null pointer");
    MV2D_CHECK_ARG(!m.gate || (m.gx && m.gs), "mlp2: the gate epilogue needs gx and gs");
    MV2D_CHECK_ARG(!m.gate || m.gs_mod == 0 || m.gs_mod % TC_BM == 0, "mlp2: gs_mod=%d must be a multiple of 128", m.gs_mod);
    CUtensorMap a, w0, w2, mo, mk, mgx, mgs, mgf;
    int rc;
    static const int cl_env = []() { const char* v = getenv("MV2D_MLP2_CLUSTER"); return v ? atoi(v) : 1; }();   // measured on B200: 2 / 4 give no gain (the ring depth, not L2 -> SM bytes, bounds the kernel)
    const int tiles = cdiv(m.M, TC_BM);
    const int cl = (cl_env == 4 && tiles >= 4) ? 4 : ((cl_env >= 2 && tiles >= 2) ? 2 : 1);
    if ((rc = tc_make_map_2d(&a, m.A, m.M, m.K0, m.lda, TC_BM))) return rc;
    if ((rc = tc_make_map_2d(&w0, m.W0, m.H, m.K0, m.K0, 128 / cl))) return rc;
    if ((rc = tc_make_map_2d(&w2, m.W2, MV2D_C, m.H, m.H, 256 / cl))) return rc;
    if ((rc = tc_make_map_2d(&mo, m.out, m.M, MV2D_C, MV2D_C, 32))) return rc;
    mk = mo; mgx = mo; mgs = mo; mgf = mo;
    if (m.gate) {
        if ((rc = tc_make_map_2d(&mgx, m.gx, m.M, MV2D_C, MV2D_C, TC_BM))) return rc;
        if ((rc = tc_make_map_2d(&mgs, m.gs, m.gs_mod > 0 ? m.gs_mod : m.M, MV2D_C, MV2D_C, TC_BM))) return rc;
        if (m.kin) {
            MV2D_CHECK_ARG(m.gfeat != nullptr, "mlp2: kin needs gfeat");
            if ((rc = tc_make_map_2d(&mk, m.kin, m.M, MV2D_C, MV2D_C, 32))) return rc;
            if ((rc = tc_make_map_2d(&mgf, m.gfeat, m.M, MV2D_C, MV2D_C, TC_BM))) return rc;
        }
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(mlp2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, M2_SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, M2_SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, M2_SMEM_BYTES);
        if (e != cudaSuccess) { set_error("mlp2: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr_set = true;
    }
    Mlp2Args g{};
    g.M = m.M; g.K0 = m.K0; g.H = m.H; g.b0 = m.b0; g.b2 = m.b2; g.out = m.out; g.ldo = MV2D_C;
    g.gate = m.gate; g.gx = m.gx; g.gs = m.gs; g.gs_mod = m.gs_mod; g.gfeat = m.gfeat; g.kin = m.kin; g.round_out = m.round_out;
    if (cl == 1) {
        launch_k(mlp2_kernel<1>, dim3(tiles), dim3(M2_THREADS), (size_t)M2_SMEM_BYTES, st, a, w0, w2, mo, mk, mgx, mgs, mgf, g);
        MV2D_CHECK_LAUNCH("mlp2");
        return 0;
    }
    // cluster launch: the grid is padded to a multiple of the cluster size; a padding CTA runs the whole pipeline on
    // out-of-range rows (TMA zero-fills its loads and clips its stores) because its peers wait for its share of the weights
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(cdiv(tiles, cl) * cl); cfg.blockDim = dim3(M2_THREADS); cfg.dynamicSmemBytes = M2_SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    cfg.attrs = attr; cfg.numAttrs = (pdl_enabled() && cap == cudaStreamCaptureStatusNone) ? 2 : 1;
    cudaError_t e = cl == 4 ? cudaLaunchKernelEx(&cfg, mlp2_kernel<4>, a, w0, w2, mo, mk, mgx, mgs, mgf, g)
                            : cudaLaunchKernelEx(&cfg, mlp2_kernel<2>, a, w0, w2, mo, mk, mgx, mgs, mgf, g);
    note_launch();
    if (e != cudaSuccess) { set_error("mlp2: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

}  // namespace mv2d
