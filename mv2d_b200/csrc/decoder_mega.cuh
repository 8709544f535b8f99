// Persistent decoder: ALL layers of the sparse cross-attention decoder (and the branches) in ONE
// cooperative launch.  At ~300 queries every stage of a decoder layer is a few microseconds of work, so
// the launch-per-stage version (run_decoder, ~75 dependent launches) spends most of its time on kernel
// boundaries.  Here one CTA per SM stays resident and walks through the same stages ("phases"),
// separated by a device-wide barrier (one atomic + acquire spin, ~1 us) instead of a kernel boundary:
//   per layer:  in-proj | self-attn | out-proj | LN | q~ GEMM | sparse cross-attn | output GEMM | LN |
//               FFN1 | FFN2 | LN(+post-norm)         then the cls/reg branches, batched over layers.
// The stage bodies are the same device functions the stand-alone kernels use (decoder.cu); the four wide
// GEMMs of a layer run on tcgen05 (3xTF32, TMA-fed, accumulator in TMEM allocated once per CTA), their
// tensor maps travel in the (large) kernel parameter block.  Included by decoder.cu.
#pragma once
#include <cuda.h>
#include <cooperative_groups.h>

namespace mv2d {

// ------------------------------------------------------------------ device-wide barrier
// Monotonic counter (zeroed by the host before the launch): generation g is complete when the counter
// reaches g * gridDim.x.  Requires co-resident CTAs => cooperative launch.
__device__ __forceinline__ void grid_sync(unsigned* bar, unsigned& gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) {   // phase timestamps (ns) for tools/mega_phases.py: bar[16 + 2*gen ..]
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            reinterpret_cast<unsigned long long*>(bar + 16)[gen] = t;
        }
        gen += 1;
        const unsigned target = gen * gridDim.x;
        __threadfence();
        atomicAdd(bar, 1u);
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
            if (v < target) __nanosleep(32);
        } while (v < target);
        __threadfence();
    }
    __syncthreads();
}

// ------------------------------------------------------------------ 32x32 FFMA GEMM tile, K = 256, in-CTA split-K
// 256 threads = 4 k-quarters x 64 threads (4x4 outputs each, same swizzled smem layout as gemm_small_kernel);
// the four partial tiles are folded through shared memory and every thread finishes one float4 of the tile.
struct G32 {
    const float* A; const float* A2; int n_switch; int lda;
    const float* W; int ldw; const float* bias; float* C; int ldc; int M, N; int relu;
};

__device__ __forceinline__ void g32_cp16(void* smem, const void* gmem, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}

__device__ __forceinline__ void gemm32_tile(const G32& g, int tm, int tn, float* smem) {
    const int tid = threadIdx.x, grp = tid >> 6, t = tid & 63, tx = t & 7, ty = t >> 3;
    float* As = smem + grp * 4096;              // [2][32][32]
    float* Ws = smem + grp * 4096 + 2048;       // [2][32][32]
    const int m0 = tm * 32, n0 = tn * 32;
    const float* __restrict__ A = (g.A2 && n0 >= g.n_switch) ? g.A2 : g.A;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int f = t + 64 * i, kt = f >> 8, row = (f >> 3) & 31, kc = f & 7, sw = kc ^ ((row >> 2) & 7);
        const int k = grp * 64 + kt * 32 + kc * 4;
        const int m = m0 + row, n = n0 + row;
        g32_cp16(As + kt * 1024 + row * 32 + sw * 4, A + (long long)min(m, g.M - 1) * g.lda + k, m < g.M);
        g32_cp16(Ws + kt * 1024 + row * 32 + sw * 4, g.W + (long long)min(n, g.N - 1) * g.ldw + k, n < g.N);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll
    for (int kt = 0; kt < 2; ++kt)
#pragma unroll
        for (int kc = 0; kc < 8; ++kc) {
            float4 av[4], wv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int ra = ty * 4 + i, rw = tx * 4 + i;
                av[i] = *reinterpret_cast<const float4*>(As + kt * 1024 + ra * 32 + ((kc ^ ((ra >> 2) & 7)) << 2));
                wv[i] = *reinterpret_cast<const float4*>(Ws + kt * 1024 + rw * 32 + ((kc ^ ((rw >> 2) & 7)) << 2));
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[i][j] = fmaf(av[i].x, wv[j].x, acc[i][j]);
                    acc[i][j] = fmaf(av[i].y, wv[j].y, acc[i][j]);
                    acc[i][j] = fmaf(av[i].z, wv[j].z, acc[i][j]);
                    acc[i][j] = fmaf(av[i].w, wv[j].w, acc[i][j]);
                }
        }
    __syncthreads();                             // operands consumed; reuse the buffer for the fold
    float* red = smem;                           // [4][32][32]
#pragma unroll
    for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(red + grp * 1024 + (ty * 4 + i) * 32 + tx * 4) =
            make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    __syncthreads();
    {
        const int row = tid >> 3, c4 = (tid & 7) * 4, m = m0 + row, n = n0 + c4;
        float4 s = *reinterpret_cast<const float4*>(red + row * 32 + c4);
#pragma unroll
        for (int gq = 1; gq < 4; ++gq) {
            const float4 v = *reinterpret_cast<const float4*>(red + gq * 1024 + row * 32 + c4);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        if (m < g.M && n < g.N) {
            if (g.bias) { s.x += __ldg(g.bias + n); s.y += __ldg(g.bias + n + 1); s.z += __ldg(g.bias + n + 2); s.w += __ldg(g.bias + n + 3); }
            if (g.relu) { s.x = fmaxf(s.x, 0.f); s.y = fmaxf(s.y, 0.f); s.z = fmaxf(s.z, 0.f); s.w = fmaxf(s.w, 0.f); }
            *reinterpret_cast<float4*>(g.C + (long long)m * g.ldc + n) = s;
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------ tcgen05 3xTF32 tile (BM 128 x BN 64), one per CTA per phase
namespace mega_tc {
constexpr int BN = 64, BK = 32, STAGES = 4;
constexpr int A_BYTES = 128 * BK * 4, W_BYTES = BN * BK * 4, STAGE_BYTES = 2 * (A_BYTES + W_BYTES);   // 48 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tMW_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra MW_DONE;\n\tbra MW_LOOP;\n\tMW_DONE:\n\t}"
        ::"r"(s32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(s32(dst)), "l"((uint64_t)map), "r"(s32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* b) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct Tile {
    const CUtensorMap *a_hi, *a_lo, *w_hi, *w_lo;
    float* C; float* C_lo; int ldc; const float* bias; int M; int relu, split_out;
    int nkb; int kb0; int raw;         // raw: split-K partial (no bias/activation)
    int m_tile, n0;
    int rot;                           // K-block rotation: CTAs sharing an operand tile start at different k
};

// One 128x64 output tile.  Must be called by all 256 threads of the CTA.
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

__device__ __forceinline__ void tile(const Tile& tl, uint8_t* smem, uint32_t tmem_base, unsigned long long* dbg = nullptr) {
    if (dbg && threadIdx.x == 0) dbg[0] = gtime();
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* done_bar = empty_bar + STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < tl.nkb; ++kb) {
                const int s = kb % STAGES, ph = (kb / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t* st = smem + s * STAGE_BYTES;
                mbar_expect(&full_bar[s], STAGE_BYTES);
                const int k = (tl.kb0 + (kb + tl.rot) % tl.nkb) * BK;
                tma2d(tl.a_hi, &full_bar[s], st, k, tl.m_tile * 128);
                tma2d(tl.a_lo, &full_bar[s], st + A_BYTES, k, tl.m_tile * 128);
                tma2d(tl.w_hi, &full_bar[s], st + 2 * A_BYTES, k, tl.n0);
                tma2d(tl.w_lo, &full_bar[s], st + 2 * A_BYTES + W_BYTES, k, tl.n0);
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        if (lane == 0) {
            for (int kb = 0; kb < tl.nkb; ++kb) {
                const int s = kb % STAGES, ph = (kb / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                if (dbg && kb == 0) dbg[1] = gtime();
                if (dbg && kb == tl.nkb - 1) dbg[2] = gtime();
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = s32(smem + s * STAGE_BYTES);
                const uint64_t a_hi = desc(sa), a_lo = desc(sa + A_BYTES), w_hi = desc(sa + 2 * A_BYTES), w_lo = desc(sa + 2 * A_BYTES + W_BYTES);
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) {
                    const uint64_t adv = (uint64_t)((k * 32) >> 4);
                    umma(tmem_base, a_hi + adv, w_hi + adv, idesc, (kb | k) != 0);
                    umma(tmem_base, a_hi + adv, w_lo + adv, idesc, 1);
                    umma(tmem_base, a_lo + adv, w_hi + adv, idesc, 1);
                }
                commit(&empty_bar[s]);
            }
            commit(done_bar);
        }
    } else if (warp >= 4) {
        mbar_wait(done_bar, 0);
        if (dbg && threadIdx.x == 128) dbg[3] = gtime();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;
        float* stg = reinterpret_cast<float*>(smem) + q * 1024;      // pipeline stages are free now
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
                const long long orow = (long long)tl.m_tile * 128 + q * 32 + rr;
                if (orow < tl.M) *reinterpret_cast<float4*>(dst + orow * tl.ldc + n + cc * 4) = v4;
            }
        };
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t v[32];
            ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
            const int n = tl.n0 + c * 32;
            float x[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
            if (!tl.raw) {
                if (tl.bias) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] += __ldg(tl.bias + n + j);
                }
                if (tl.relu) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
                }
                if (tl.split_out) {
                    float lo[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) { const float hi = round_tf32(x[j]); lo[j] = round_tf32(x[j] - hi); x[j] = hi; }
                    store_t(tl.C_lo, lo, n);
                }
            }
            store_t(tl.C, x, n);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (dbg && threadIdx.x == 0) dbg[4] = gtime();
}
}  // namespace mega_tc

// ------------------------------------------------------------------ parameters
struct MegaLayer {
    const float *sa_in_w, *sa_in_b, *sa_out_w, *sa_out_b;
    const float *ca_q_b, *ca_o_b, *ffn_b1, *ffn_b2;
    const float *ln_g[3], *ln_b[3];
    // tensor maps of the four tcgen05 GEMMs: [gemm][a_hi, a_lo, w_hi, w_lo]
    CUtensorMap maps[4][4];
};

struct MegaParams {
    int N, L, mode, max_match, mask_words, klist_cap;
    float pc_range[6]; float vel_dt; int vel_row_start;
    const float *query_pos, *ref, *kin_rows, *mem_rows;
    const int *match, *match_cnt; const uint32_t* keymask; const uint8_t* self_attn_mask;
    const uint16_t* key_list; const int* key_cnt;
    float *x, *xq, *x1, *x1q, *x2, *x1q_hi, *x1q_lo, *x2_hi, *x2_lo, *qkv, *sa, *qt, *ctx, *ctx_lo, *hdn, *hdn_lo, *part;
    float *b0, *b1, *b2, *b3;
    float *cls, *box, *outs_dec;
    Mv2dBranchWeights br;
    unsigned* barrier;
    MegaLayer layer[MV2D_MAX_LAYERS];
};

#define MEGA_SPLIT 8
#define MEGA_THREADS 256

__global__ void __launch_bounds__(MEGA_THREADS, 1)
decoder_mega_kernel(const __grid_constant__ MegaParams p) {
    extern __shared__ __align__(1024) uint8_t mega_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(mega_smem_raw) + 1023) & ~(uintptr_t)1023);
    float* smem_f = reinterpret_cast<float*>(smem);
    __shared__ uint32_t tmem_slot;
    const int cta = blockIdx.x, ncta = gridDim.x, warp = threadIdx.x >> 5;
    const int N = p.N, C = MV2D_C;
    const long long NC = (long long)N * C;
    unsigned gen = 0;

    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(mega_tc::s32(&tmem_slot)), "n"(mega_tc::BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_slot;

    const int mt = cdiv(N, 32);            // 32-row tiles of the FFMA GEMMs
    const int m128 = cdiv(N, 128);         // 128-row tiles of the tcgen05 GEMMs

    auto ln_phase = [&](LnArgs a) {
        for (int vb = cta; vb < cdiv(a.rows, 8); vb += ncta) ln_body(a, vb);
    };
    int l_dbg = 0;
    auto tc_phase = [&](const MegaLayer& ly, int gi, float* Cc, float* C_lo, int ldc, const float* bias, int Nn, int K,
                        int relu, int split_out, int nsplit, long long split_stride) {
        const int ntile_n = Nn / mega_tc::BN, tiles = m128 * ntile_n * nsplit;
        for (int t = cta; t < tiles; t += ncta) {
            const int z = t / (m128 * ntile_n), r = t % (m128 * ntile_n);
            mega_tc::Tile tl;
            tl.a_hi = &ly.maps[gi][0]; tl.a_lo = &ly.maps[gi][1]; tl.w_hi = &ly.maps[gi][2]; tl.w_lo = &ly.maps[gi][3];
            tl.C = Cc + z * split_stride; tl.C_lo = C_lo; tl.ldc = ldc; tl.bias = bias; tl.M = N; tl.relu = relu;
            tl.split_out = split_out; tl.nkb = K / mega_tc::BK / nsplit; tl.kb0 = z * tl.nkb; tl.raw = nsplit > 1;
            tl.m_tile = r / ntile_n; tl.n0 = (r % ntile_n) * mega_tc::BN;
            tl.rot = ((r % ntile_n) + 3 * (r / ntile_n)) % tl.nkb;
            mega_tc::tile(tl, smem, tmem_base, (cta == 0 && l_dbg == 0) ? reinterpret_cast<unsigned long long*>(p.barrier + 16) + 400 + gi * 8 : nullptr);
        }
    };

    for (int l = 0; l < p.L; ++l) {
        l_dbg = l;
        const MegaLayer& ly = p.layer[l];
        // ---- 1. self-attention in-projection: q,k from (x + qpos), v from x
        {
            G32 g{p.xq, p.x, 512, C, ly.sa_in_w, C, ly.sa_in_b, p.qkv, 768, N, 768, 0};
            for (int t = cta; t < mt * 24; t += ncta) gemm32_tile(g, t / 24, t % 24, smem_f);
        }
        grid_sync(p.barrier, gen);
        // ---- 2. self-attention core
        {
            const int qb = cdiv(N, 8 * SA1_QPW);
            for (int t = cta; t < qb * MV2D_HEADS; t += ncta)
                self_attn_body_v1(p.qkv, p.self_attn_mask, N, p.sa, t % qb, t / qb, smem_f);
        }
        grid_sync(p.barrier, gen);
        // ---- 3. out-projection (raw; bias + residual + LN in the next phase)
        {
            G32 g{p.sa, nullptr, 0, C, ly.sa_out_w, C, nullptr, p.part, C, N, C, 0};
            for (int t = cta; t < mt * 8; t += ncta) gemm32_tile(g, t / 8, t % 8, smem_f);
        }
        grid_sync(p.barrier, gen);
        // ---- 4. LN1
        {
            LnArgs a{}; a.partial = p.part; a.nsplit = 1; a.bias = ly.sa_out_b; a.residual = p.x;
            a.gamma = ly.ln_g[0]; a.beta = ly.ln_b[0]; a.qpos = p.query_pos; a.out = p.x1; a.out_q = p.x1q; a.rows = N;
            a.outq_hi = p.x1q_hi; a.outq_lo = p.x1q_lo;
            ln_phase(a);
        }
        grid_sync(p.barrier, gen);
        // ---- 5. absorbed query projection (tcgen05 3xTF32)
        tc_phase(ly, 0, p.qt, nullptr, 2048, ly.ca_q_b, 2048, C, 0, 0, 1, 0);
        grid_sync(p.barrier, gen);
        // ---- 6. sparse cross-attention
        {
            XaArgs a{}; a.qt = p.qt; a.kin_rows = p.kin_rows; a.mem_rows = p.mem_rows; a.match = p.match;
            a.match_cnt = p.match_cnt; a.max_match = p.max_match; a.keymask = p.keymask; a.mask_words = p.mask_words;
            a.key_list = p.key_list; a.key_cnt = p.key_cnt;
            a.mode = p.mode; a.N = N; a.klist_cap = p.klist_cap; a.ctx = p.ctx; a.ctx_lo = p.ctx_lo;
            for (int n = cta; n < N; n += ncta) {
                cross_attn_body<32, 256>(a, n, smem);
                __syncthreads();
            }
        }
        grid_sync(p.barrier, gen);
        // ---- 7. absorbed output projection, split-K
        tc_phase(ly, 1, p.part, nullptr, C, nullptr, C, 2048, 0, 0, MEGA_SPLIT, NC);
        grid_sync(p.barrier, gen);
        // ---- 8. LN2
        {
            LnArgs a{}; a.partial = p.part; a.nsplit = MEGA_SPLIT; a.split_stride = NC; a.bias = ly.ca_o_b; a.residual = p.x1;
            a.gamma = ly.ln_g[1]; a.beta = ly.ln_b[1]; a.out = p.x2; a.rows = N; a.out_hi = p.x2_hi; a.out_lo = p.x2_lo;
            ln_phase(a);
        }
        grid_sync(p.barrier, gen);
        // ---- 9. FFN1 (ReLU, TF32 hi/lo split of the hidden)
        tc_phase(ly, 2, p.hdn, p.hdn_lo, 2048, ly.ffn_b1, 2048, C, 1, 1, 1, 0);
        grid_sync(p.barrier, gen);
        // ---- 10. FFN2, split-K
        tc_phase(ly, 3, p.part, nullptr, C, nullptr, C, 2048, 0, 0, MEGA_SPLIT, NC);
        grid_sync(p.barrier, gen);
        // ---- 11. LN3 (+ post_norm -> intermediate l)
        {
            LnArgs a{}; a.partial = p.part; a.nsplit = MEGA_SPLIT; a.split_stride = NC; a.bias = ly.ffn_b2; a.residual = p.x2;
            a.gamma = ly.ln_g[2]; a.beta = ly.ln_b[2]; a.qpos = p.query_pos; a.out = p.x; a.out_q = p.xq;
            a.gamma2 = p.br.post_g; a.beta2 = p.br.post_b; a.out2 = p.outs_dec + (long long)l * NC; a.rows = N;
            ln_phase(a);
        }
        grid_sync(p.barrier, gen);
    }
    // ---- branches, batched over layers (cross_attention_head.py:216-231)
    const long long CC = (long long)C * C;
    auto branch_gemm = [&](const float* A, const float* W, const float* bias, float* out, int relu, int t) {
        const int l = t / (mt * 8), r = t % (mt * 8);
        G32 g{A + l * NC, nullptr, 0, C, W + l * CC, C, bias ? bias + l * C : nullptr, out + l * NC, C, N, C, relu};
        gemm32_tile(g, r / 8, r % 8, smem_f);
    };
    const int bt = p.L * mt * 8;
    for (int t = cta; t < 2 * bt; t += ncta) {
        if (t < bt) branch_gemm(p.outs_dec, p.br.cls_w0, nullptr, p.b0, 0, t);
        else        branch_gemm(p.outs_dec, p.br.reg_w0, p.br.reg_b0, p.b2, 1, t - bt);
    }
    grid_sync(p.barrier, gen);
    {
        LnArgs a{}; a.partial = p.b0; a.nsplit = 1; a.bias = p.br.cls_b0; a.gamma = p.br.cls_g0; a.beta = p.br.cls_be0;
        a.rows_per_group = N; a.group_stride = C; a.relu = 1; a.out = p.b1; a.rows = p.L * N;
        ln_phase(a);
    }
    grid_sync(p.barrier, gen);
    for (int t = cta; t < 2 * bt; t += ncta) {
        if (t < bt) branch_gemm(p.b1, p.br.cls_w1, nullptr, p.b0, 0, t);
        else        branch_gemm(p.b2, p.br.reg_w1, p.br.reg_b1, p.b3, 1, t - bt);
    }
    grid_sync(p.barrier, gen);
    {
        LnArgs a{}; a.partial = p.b0; a.nsplit = 1; a.bias = p.br.cls_b1; a.gamma = p.br.cls_g1; a.beta = p.br.cls_be1;
        a.rows_per_group = N; a.group_stride = C; a.relu = 1; a.out = p.b1; a.rows = p.L * N;
        ln_phase(a);
    }
    grid_sync(p.barrier, gen);
    for (int vb = cta; vb < cdiv(p.L * N, 8); vb += ncta)
        head10_body(p.b1, p.b3, p.br.cls_w2, p.br.cls_b2, p.br.reg_w2, p.br.reg_b2, p.ref, p.L, N, p.pc_range[0], p.pc_range[1],
                    p.pc_range[2], p.pc_range[3], p.pc_range[4], p.pc_range[5], p.vel_dt, p.vel_row_start, p.cls, p.box, vb);

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(mega_tc::BN));
    }
}

}  // namespace mv2d
