// K4/K5: the sparse multi-view cross-attention decoder (rows a13-a19 of SURVEY.md section 8a).
//   reference: roi_heads/bbox_heads/cross_attention_head.py:22-49,202-242;
//              utils/petr_transformer.py:194-593 (over mmcv BaseTransformerLayer / FFN and
//              torch.nn.MultiheadAttention)
//
// Cross-attention is evaluated in "absorbed" form (exact in real arithmetic): with
//   q~_h = scale * Wk_h^T (Wq_h x + bq_h)           (256 numbers per head, one small GEMM)
// the logits are  q~_h . (memory_k + pos_k)  (+ a per-(query,head) constant that softmax
// cancels), and the head output is  Wv_h (sum_k p_hk memory_k) + bv_h.  So the kernel streams
// the RAW key rows once per query -- no K/V projection of 49*N (or ~25k) keys per layer --
// and the work per layer is bound by the bytes of the key rows it reads (HBM/L2), not flops.
#include <stdlib.h>
#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "mv2d_internal.h"

#include "ln.cuh"

namespace mv2d {

__global__ void __launch_bounds__(256, 3) ln_kernel(LnArgs a) {      // 3 CTAs per SM: the 300 CTAs of a batch of 8 are one wave (296 slots at 2)
    pdl_wait();
    pdl_trigger();
    ln_body(a, blockIdx.x);
}

// ------------------------------------------------------------------------------------------
// FlattenMHSelfAttention core (petr_transformer.py:314-370): all N queries form ONE sequence.
// qkv [N,768] (q | k | v, head h = channels 32h..32h+31), q already includes the bias; the
// 1/sqrt(32) scale is applied here.  grid (ceil(N/8), 8 heads), 256 threads, one query per warp.
// K/V tiles of up to 320 keys (one tile at N = 300) are copied with cp.async into shared memory with a 36-float
// row stride (16-byte aligned, conflict-free LDS.128).  QK^T: lane = key (up to 10 keys per lane), q in registers.
// PV: probabilities go through shared memory; lane = (key phase, channel quad) so each step is
// one LDS + one LDS.128 + 4 FMA; the 4 key phases are folded with two shuffles at the end.
#define SA_KT 320
#define SA_LD 36
#define SA_SMEM_BYTES ((2 * SA_KT * SA_LD + 8 * SA_KT) * 4)
#define SA_QPW 1     // queries per warp
__device__ __forceinline__ void sa_cp16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}

// A sample of a batch owns rows [seg0, seg0 + nq); its keys are the first N of them (seg0 = 0, nq = N: one sample).
__device__ __forceinline__ void
self_attn_body(const float* __restrict__ qkv, const uint8_t* __restrict__ mask, int N, float* __restrict__ out,
               int vbx, int hd, float* smem_f, int seg0 = 0, int nq = -1, float* __restrict__ out_lo = nullptr) {
    float (*Ks)[SA_LD] = reinterpret_cast<float (*)[SA_LD]>(smem_f);
    float (*Vs)[SA_LD] = reinterpret_cast<float (*)[SA_LD]>(smem_f + SA_KT * SA_LD);
    float (*Ps)[SA_KT] = reinterpret_cast<float (*)[SA_KT]>(smem_f + 2 * SA_KT * SA_LD);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kq = lane >> 3, cq = lane & 7;
    if (nq < 0) nq = N;
    const int ql = vbx * 8 + warp;                // query row inside the sample
    const int qi = ql;                            // row of the (single-sample) attention mask
    qkv += (long long)seg0 * 768;
    out += (long long)seg0 * MV2D_C;
    float q[32], m = -INFINITY, l = 0.f;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k0 = 0; k0 < N; k0 += SA_KT) {
        const int ng = min(SA_KT, N - k0);
        __syncthreads();
        // the head's K and V rows of this tile (128 B each) go straight to shared memory, all copies in flight at once
        for (int i = threadIdx.x; i < ng * 8; i += 256) {
            const int r = i >> 3, c4 = i & 7;
            const float* src = qkv + (long long)(k0 + r) * 768 + 256 + hd * 32 + c4 * 4;
            sa_cp16(&Ks[r][c4 * 4], src);
            sa_cp16(&Vs[r][c4 * 4], src + 256);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (k0 == 0) {
#pragma unroll
            for (int d4 = 0; d4 < 8; ++d4) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                // ld.global.cg: inside the persistent kernel qkv was written by other CTAs a phase ago (no .nc path)
                if (ql < nq) v = __ldcg(reinterpret_cast<const float4*>(qkv + (long long)ql * 768 + hd * 32) + d4);
                q[d4 * 4 + 0] = v.x * 0.17677669529663687f; q[d4 * 4 + 1] = v.y * 0.17677669529663687f;
                q[d4 * 4 + 2] = v.z * 0.17677669529663687f; q[d4 * 4 + 3] = v.w * 0.17677669529663687f;
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (ql >= nq) continue;                   // warp-uniform
        const int ngrp = (ng + 31) >> 5;
        float s[SA_KT / 32];
        float mx = -INFINITY;
#pragma unroll
        for (int g = 0; g < SA_KT / 32; ++g) {
            s[g] = -INFINITY;
            if (g < ngrp) {
                const int kk = g * 32 + lane;
                float x = 0.f;
                if (kk < ng) {
#pragma unroll
                    for (int d4 = 0; d4 < 8; ++d4) {
                        const float4 kv = *reinterpret_cast<const float4*>(&Ks[kk][d4 * 4]);
                        x = fmaf(q[d4 * 4 + 0], kv.x, x); x = fmaf(q[d4 * 4 + 1], kv.y, x);
                        x = fmaf(q[d4 * 4 + 2], kv.z, x); x = fmaf(q[d4 * 4 + 3], kv.w, x);
                    }
                }
                if (kk >= ng || (mask && mask[(long long)qi * N + k0 + kk])) x = -INFINITY;
                s[g] = x;
                mx = fmaxf(mx, x);
            }
        }
        const float mn = fmaxf(m, warp_max(mx));
        if (mn == -INFINITY) continue;
        const float alpha = __expf(m - mn);       // exp(-inf) = 0 on the first tile
        float psum = 0.f;
        __syncwarp();
#pragma unroll
        for (int g = 0; g < SA_KT / 32; ++g) {
            if (g < ngrp) {
                const float pv = __expf(s[g] - mn);
                Ps[warp][g * 32 + lane] = pv;
                psum += pv;
            }
        }
        l = l * alpha + warp_sum(psum);
        acc.x *= alpha; acc.y *= alpha; acc.z *= alpha; acc.w *= alpha;
        __syncwarp();
        const int nt = (ng + 3) >> 2;             // keys beyond ng inside the last group carry p = 0 and finite (stale) V
#pragma unroll 5
        for (int t = 0; t < nt; ++t) {
            const int j = t * 4 + kq;
            const float pj = j < ng ? Ps[warp][j] : 0.f;
            const float4 vv = *reinterpret_cast<const float4*>(&Vs[j < ng ? j : 0][cq * 4]);
            acc.x = fmaf(pj, vv.x, acc.x); acc.y = fmaf(pj, vv.y, acc.y);
            acc.z = fmaf(pj, vv.z, acc.z); acc.w = fmaf(pj, vv.w, acc.w);
        }
        m = mn;
    }
    // fold the 4 key phases (lanes differing in bits 3,4)
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
        acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
    }
    if (ql < nq && kq == 0) {
        const float inv = l > 0.f ? 1.f / l : 0.f;
        float4 r = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
        if (out_lo) {     // TF32 hi / lo split: A operand of the 3xTF32 output projection
            const float4 hi = make_float4(round_tf32(r.x), round_tf32(r.y), round_tf32(r.z), round_tf32(r.w));
            *reinterpret_cast<float4*>(out_lo + (long long)(seg0 + ql) * MV2D_C + hd * 32 + cq * 4) =
                make_float4(round_tf32(r.x - hi.x), round_tf32(r.y - hi.y), round_tf32(r.z - hi.z), round_tf32(r.w - hi.w));
            r = hi;
        }
        *reinterpret_cast<float4*>(out + (long long)ql * MV2D_C + hd * 32 + cq * 4) = r;
    }
}

// Previous formulation (128-key tiles, two queries per warp, register-staged loads): kept for the persistent
// decoder (decoder_mega.cuh).  Measured on B200: with the 100 KB single-tile variant above inlined, the persistent
// kernel's later tcgen05 split-K phase returned wrong sums (self-attention output itself was right; the TMA
// stage buffers alias the region the variant dirties -- suspected generic/async proxy ordering, not root-caused).
// The staged decoder (one kernel per stage, no aliasing) is unaffected and uses the variant above.
#define SA1_KT 128
#define SA1_LD 36
#define SA1_SMEM_BYTES ((2 * SA1_KT * SA1_LD + 8 * SA1_KT) * 4)
#define SA1_QPW 2     // queries per warp: every K/V tile staged in shared memory serves 16 queries
__device__ __forceinline__ void
self_attn_body_v1(const float* __restrict__ qkv, const uint8_t* __restrict__ mask, int N, float* __restrict__ out,
               int vbx, int hd, float* smem_f) {
    float (*Ks)[SA1_LD] = reinterpret_cast<float (*)[SA1_LD]>(smem_f);
    float (*Vs)[SA1_LD] = reinterpret_cast<float (*)[SA1_LD]>(smem_f + SA1_KT * SA1_LD);
    float (*Ps)[SA1_KT] = reinterpret_cast<float (*)[SA1_KT]>(smem_f + 2 * SA1_KT * SA1_LD);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kq = lane >> 3, cq = lane & 7;
    int qi[SA1_QPW];
    float q[SA1_QPW][32], m[SA1_QPW], l[SA1_QPW];
    float4 acc[SA1_QPW];
#pragma unroll
    for (int u = 0; u < SA1_QPW; ++u) {
        qi[u] = (vbx * 8 + warp) * SA1_QPW + u;
        m[u] = -INFINITY; l[u] = 0.f; acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int d4 = 0; d4 < 8; ++d4) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (qi[u] < N) v = __ldg(reinterpret_cast<const float4*>(qkv + (long long)qi[u] * 768 + hd * 32) + d4);
            q[u][d4 * 4 + 0] = v.x * 0.17677669529663687f; q[u][d4 * 4 + 1] = v.y * 0.17677669529663687f;
            q[u][d4 * 4 + 2] = v.z * 0.17677669529663687f; q[u][d4 * 4 + 3] = v.w * 0.17677669529663687f;
        }
    }
    for (int k0 = 0; k0 < N; k0 += SA1_KT) {
        __syncthreads();
        {   // all 8 loads of a thread are issued before the first store (one memory latency per tile)
            float4 kv[4], vv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = threadIdx.x + u * 256, r = i >> 3, c4 = i & 7, k = k0 + r;
                kv[u] = make_float4(0.f, 0.f, 0.f, 0.f); vv[u] = kv[u];
                if (k < N) {
                    kv[u] = __ldg(reinterpret_cast<const float4*>(qkv + (long long)k * 768 + 256 + hd * 32) + c4);
                    vv[u] = __ldg(reinterpret_cast<const float4*>(qkv + (long long)k * 768 + 512 + hd * 32) + c4);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = threadIdx.x + u * 256, r = i >> 3, c4 = i & 7;
                *reinterpret_cast<float4*>(&Ks[r][c4 * 4]) = kv[u];
                *reinterpret_cast<float4*>(&Vs[r][c4 * 4]) = vv[u];
            }
        }
        __syncthreads();
        const int ng = min(SA1_KT, N - k0);
#pragma unroll
        for (int u = 0; u < SA1_QPW; ++u) {
            if (qi[u] >= N) continue;                    // warp-uniform
            float s[SA1_KT / 32];
            float mx = -INFINITY;
#pragma unroll
            for (int g = 0; g < SA1_KT / 32; ++g) {
                const int kk = g * 32 + lane;
                float x = 0.f;
#pragma unroll
                for (int d4 = 0; d4 < 8; ++d4) {
                    const float4 kv = *reinterpret_cast<const float4*>(&Ks[kk][d4 * 4]);
                    x = fmaf(q[u][d4 * 4 + 0], kv.x, x); x = fmaf(q[u][d4 * 4 + 1], kv.y, x);
                    x = fmaf(q[u][d4 * 4 + 2], kv.z, x); x = fmaf(q[u][d4 * 4 + 3], kv.w, x);
                }
                if (kk >= ng || (mask && mask[(long long)qi[u] * N + k0 + kk])) x = -INFINITY;
                s[g] = x;
                mx = fmaxf(mx, x);
            }
            const float mn = fmaxf(m[u], warp_max(mx));
            if (mn == -INFINITY) continue;
            const float alpha = __expf(m[u] - mn);       // exp(-inf) = 0 on the first tile
            float psum = 0.f;
            __syncwarp();
#pragma unroll
            for (int g = 0; g < SA1_KT / 32; ++g) {
                const float pv = __expf(s[g] - mn);
                Ps[warp][g * 32 + lane] = pv;
                psum += pv;
            }
            l[u] = l[u] * alpha + warp_sum(psum);
            acc[u].x *= alpha; acc[u].y *= alpha; acc[u].z *= alpha; acc[u].w *= alpha;
            __syncwarp();
#pragma unroll 8
            for (int t = 0; t < SA1_KT / 4; ++t) {
                const int j = t * 4 + kq;
                const float pj = Ps[warp][j];
                const float4 vv = *reinterpret_cast<const float4*>(&Vs[j][cq * 4]);
                acc[u].x = fmaf(pj, vv.x, acc[u].x); acc[u].y = fmaf(pj, vv.y, acc[u].y);
                acc[u].z = fmaf(pj, vv.z, acc[u].z); acc[u].w = fmaf(pj, vv.w, acc[u].w);
            }
            m[u] = mn;
        }
    }
#pragma unroll
    for (int u = 0; u < SA1_QPW; ++u) {
        // fold the 4 key phases (lanes differing in bits 3,4)
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {
            acc[u].x += __shfl_xor_sync(0xffffffffu, acc[u].x, o); acc[u].y += __shfl_xor_sync(0xffffffffu, acc[u].y, o);
            acc[u].z += __shfl_xor_sync(0xffffffffu, acc[u].z, o); acc[u].w += __shfl_xor_sync(0xffffffffu, acc[u].w, o);
        }
        if (qi[u] < N && kq == 0) {
            const float inv = l[u] > 0.f ? 1.f / l[u] : 0.f;
            *reinterpret_cast<float4*>(out + (long long)qi[u] * MV2D_C + hd * 32 + cq * 4) =
                make_float4(acc[u].x * inv, acc[u].y * inv, acc[u].z * inv, acc[u].w * inv);
        }
    }
}

// grid (ceil(rows_per_sample / 8), heads, batch).  n_real (nullable, device [batch]): keys of sample b = its first n_real[b] rows.
__global__ void __launch_bounds__(256)
self_attn_kernel(const float* __restrict__ qkv, const uint8_t* __restrict__ mask, int N, float* __restrict__ out,
                 int rows_per_sample, const int* __restrict__ n_real, float* __restrict__ out_lo) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ __align__(16) float sa_smem[];
    if (rows_per_sample > 0) {
        const int b = blockIdx.z;
        const int nk = n_real ? min(n_real[b], rows_per_sample) : rows_per_sample;
        self_attn_body(qkv, nullptr, nk, out, blockIdx.x, blockIdx.y, sa_smem, b * rows_per_sample, rows_per_sample, out_lo);
    } else {
        self_attn_body(qkv, mask, N, out, blockIdx.x, blockIdx.y, sa_smem, 0, -1, out_lo);
    }
}

// ------------------------------------------------------------------------------------------
// FlattenMHSelfAttention core, blocked form (petr_transformer.py:314-370).  One CTA = (64 queries, head, sample); the
// head's keys stream through shared memory in tiles of 160 (K transposed, V row-major) and every staged tile serves all
// 64 queries (the one-query-per-warp kernel above re-stages the whole K/V slice for every 8 queries).
//   phase A  S = Q K^T: thread = 4 queries x 10 keys (register tile, 21 LDS per 80 FMA), online softmax statistics by
//            half-warp shuffles, P^T -> shared memory;
//   phase B  O += P V: warp = 8 queries, lane = (key phase 0..3, 4 channels), 3 LDS.128 per 32 FMA; the four key phases
//            are folded by two shuffles at the end.
// fp32 FFMA throughout (the parity gate rules out single-pass TF32 logits).
#define SB_KT 160
#define SB_KP 33           // K row stride: lane = key reads of one channel are conflict-free (consecutive keys, odd stride)
template <int QT> struct SbCfg {
    static constexpr int PP = QT + 4;                  // P^T row stride (conflict-free float4 stores of consecutive keys)
    static constexpr int NQG = QT / 4;                 // phase A: groups of 4 queries
    static constexpr int KL = 256 / NQG;               // lanes (keys) per group: 16 (QT = 64) or 32 (QT = 32)
    static constexpr int JN = SB_KT / KL;              // keys per thread: 10 or 5
    static constexpr int QW = QT / 8;                  // phase B: queries per warp: 8 or 4
    static constexpr int SMEM_FLOATS = 32 * QT + SB_KT * SB_KP + SB_KT * 32 + SB_KT * PP + 2 * QT;
    static constexpr int SMEM_BYTES = SMEM_FLOATS * 4;
};

// QT = 64 queries per CTA for batches; QT = 32 when the 64-query grid would leave most SMs idle (one sample)
template <int QT>
__global__ void __launch_bounds__(256, QT == 32 ? 3 : 2)
self_attn_blk_kernel(const float* __restrict__ qkv, const uint8_t* __restrict__ mask, int N, float* __restrict__ out,
                     int rows_per_sample, const int* __restrict__ n_real, float* __restrict__ out_lo) {
    using Cf = SbCfg<QT>;
    constexpr int PP = Cf::PP, KL = Cf::KL, JN = Cf::JN, QW = Cf::QW;
    pdl_wait();
    pdl_trigger();
    extern __shared__ __align__(16) float sb_smem[];
    float* Qt = sb_smem;                       // [32][QT]   q * scale, transposed
    float* Kt = Qt + 32 * QT;                  // [160][33]  keys of the tile (row stride 33)
    float* Vs = Kt + SB_KT * SB_KP;            // [160][32]
    float* Pt = Vs + SB_KT * 32;               // [160][PP]  probabilities, transposed
    float* alpha_s = Pt + SB_KT * PP;          // [QT] rescale of the running output for this tile
    float* l_s = alpha_s + QT;                 // [QT] softmax denominators
    const int t = threadIdx.x, hd = blockIdx.y;
    int seg0 = 0, nq = N, nk = N;
    if (rows_per_sample > 0) {
        const int b = blockIdx.z;
        seg0 = b * rows_per_sample; nq = rows_per_sample;
        nk = n_real ? min(n_real[b], rows_per_sample) : rows_per_sample;
        mask = nullptr;
    }
    const int q0 = blockIdx.x * QT;
    const float* base = qkv + (long long)seg0 * 768 + hd * 32;
    // ---- Q tile, transposed: consecutive lanes take consecutive queries (conflict-free stores)
    for (int i = t; i < QT * 8; i += 256) {
        const int q = i % QT, d4 = i / QT;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q0 + q < nq) v = __ldcg(reinterpret_cast<const float4*>(base + (long long)(q0 + q) * 768) + d4);
        Qt[(d4 * 4 + 0) * QT + q] = v.x * 0.17677669529663687f; Qt[(d4 * 4 + 1) * QT + q] = v.y * 0.17677669529663687f;
        Qt[(d4 * 4 + 2) * QT + q] = v.z * 0.17677669529663687f; Qt[(d4 * 4 + 3) * QT + q] = v.w * 0.17677669529663687f;
    }
    const int qg = t / KL, kl = t % KL;                 // phase A: queries 4qg..4qg+3, keys kl + KL j
    const int warp = t >> 5, lane = t & 31;
    const int ks = lane >> 3, dl = lane & 7;            // phase B: queries QW warp .. +QW-1, keys ks + 4 i, channels 4dl..4dl+3
    float m_run[4], l_run[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { m_run[i] = -INFINITY; l_run[i] = 0.f; }
    float o[QW][4];
#pragma unroll
    for (int i = 0; i < QW; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) o[i][c] = 0.f;
    for (int k0 = 0; k0 < nk; k0 += SB_KT) {
        const int ng = min(SB_KT, nk - k0);
        __syncthreads();                                // previous tile fully consumed (and Qt written, first time)
        // the whole tile goes to shared memory asynchronously (one memory latency per tile, not one per loop trip):
        // K with 4-byte copies, a warp per 128-byte key row (the odd row stride rules out 16-byte copies), V with
        // 16-byte copies, a quarter warp per row
        for (int i = t; i < SB_KT * 32; i += 256) {
            const int key = i >> 5, d = i & 31;
            if (key < ng) {
                const unsigned sa = (unsigned)__cvta_generic_to_shared(Kt + key * SB_KP + d);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(base + (long long)(k0 + key) * 768 + 256 + d) : "memory");
            } else {
                Kt[key * SB_KP + d] = 0.f;
            }
        }
        for (int i = t; i < SB_KT * 8; i += 256) {
            const int key = i >> 3, c4 = i & 7;
            if (key < ng) sa_cp16(Vs + key * 32 + c4 * 4, base + (long long)(k0 + key) * 768 + 512 + c4 * 4);
            else *reinterpret_cast<float4*>(Vs + key * 32 + c4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        // ---- phase A
        float acc[4][JN];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < JN; ++j) acc[i][j] = 0.f;
#pragma unroll 4
        for (int d = 0; d < 32; ++d) {
            const float4 q4 = *reinterpret_cast<const float4*>(Qt + d * QT + qg * 4);
            const float* kr = Kt + kl * SB_KP + d;
#pragma unroll
            for (int j = 0; j < JN; ++j) {
                const float kv = kr[KL * j * SB_KP];
                acc[0][j] = fmaf(q4.x, kv, acc[0][j]); acc[1][j] = fmaf(q4.y, kv, acc[1][j]);
                acc[2][j] = fmaf(q4.z, kv, acc[2][j]); acc[3][j] = fmaf(q4.w, kv, acc[3][j]);
            }
        }
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < JN; ++j) {
            const int key = kl + KL * j;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int q = q0 + qg * 4 + i;
                bool off = key >= ng;
                if (mask && !off && q < nq) off = mask[(long long)q * N + k0 + key] != 0;
                if (off) acc[i][j] = -INFINITY;
                mx[i] = fmaxf(mx[i], acc[i][j]);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int ofs = KL / 2; ofs > 0; ofs >>= 1) mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], ofs));
            const float m_new = fmaxf(m_run[i], mx[i]);
            const float al = (m_run[i] == -INFINITY) ? 0.f : __expf(m_run[i] - m_new);
            float sm = 0.f;
#pragma unroll
            for (int j = 0; j < JN; ++j) {
                const float pv = (m_new == -INFINITY) ? 0.f : __expf(acc[i][j] - m_new);
                acc[i][j] = pv;
                sm += pv;
            }
#pragma unroll
            for (int ofs = KL / 2; ofs > 0; ofs >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, ofs);
            l_run[i] = l_run[i] * al + sm;
            m_run[i] = m_new;
            if (kl == 0) alpha_s[qg * 4 + i] = al;
        }
#pragma unroll
        for (int j = 0; j < JN; ++j)
            *reinterpret_cast<float4*>(Pt + (kl + KL * j) * PP + qg * 4) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
        __syncthreads();
        // ---- phase B
#pragma unroll
        for (int i = 0; i < QW; ++i) {
            const float al = alpha_s[warp * QW + i];
#pragma unroll
            for (int c = 0; c < 4; ++c) o[i][c] *= al;
        }
#pragma unroll 2
        for (int key = ks; key < ng; key += 4) {
            const float4 v = *reinterpret_cast<const float4*>(Vs + key * 32 + dl * 4);
            float pp[QW];
#pragma unroll
            for (int i4 = 0; i4 < QW / 4; ++i4) {
                const float4 p4 = *reinterpret_cast<const float4*>(Pt + key * PP + warp * QW + i4 * 4);
                pp[i4 * 4] = p4.x; pp[i4 * 4 + 1] = p4.y; pp[i4 * 4 + 2] = p4.z; pp[i4 * 4 + 3] = p4.w;
            }
#pragma unroll
            for (int i = 0; i < QW; ++i) {
                o[i][0] = fmaf(pp[i], v.x, o[i][0]); o[i][1] = fmaf(pp[i], v.y, o[i][1]);
                o[i][2] = fmaf(pp[i], v.z, o[i][2]); o[i][3] = fmaf(pp[i], v.w, o[i][3]);
            }
        }
    }
    if (kl == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) l_s[qg * 4 + i] = l_run[i];
    }
    __syncthreads();
    // fold the 4 key phases (lanes differing in bits 3, 4), normalise, store
#pragma unroll
    for (int i = 0; i < QW; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            o[i][c] += __shfl_xor_sync(0xffffffffu, o[i][c], 8);
            o[i][c] += __shfl_xor_sync(0xffffffffu, o[i][c], 16);
        }
    if (ks == 0) {
#pragma unroll
        for (int i = 0; i < QW; ++i) {
            const int q = q0 + warp * QW + i;
            if (q >= nq) continue;
            const float l = l_s[warp * QW + i], inv = l > 0.f ? 1.f / l : 0.f;
            float4 r = make_float4(o[i][0] * inv, o[i][1] * inv, o[i][2] * inv, o[i][3] * inv);
            const long long off = (long long)(seg0 + q) * MV2D_C + hd * 32 + dl * 4;
            if (out_lo) {
                const float4 hi = make_float4(round_tf32(r.x), round_tf32(r.y), round_tf32(r.z), round_tf32(r.w));
                *reinterpret_cast<float4*>(out_lo + off) =
                    make_float4(round_tf32(r.x - hi.x), round_tf32(r.y - hi.y), round_tf32(r.z - hi.z), round_tf32(r.w - hi.w));
                r = hi;
            }
            *reinterpret_cast<float4*>(out + off) = r;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Sparse multi-view cross-attention core, absorbed form.  One CTA (8 warps) per query.
//   qt   [N, 8*256]   q~ per head (scale folded in)
//   keys: mode 0 -> RoI tokens of the matched RoIs (match list); mode 1 -> set bits of keymask
//   kin_rows / mem_rows [num_rows, 256]   key input (memory + pos) / value input (memory)
//   ctx  [N, 8*256]   sum_k softmax_h(k) * mem_row_k
// The key rows of a query are streamed through shared memory in chunks of XA_CH keys with
// cp.async (16-byte copies, all 64 KB of a chunk in flight at once, double buffered so the next
// chunk loads while this one is consumed).  Per chunk: logits (warp per key, lanes along the
// 256 channels, warp-shuffle transpose-reduce over 8 heads), online softmax (warp per head,
// lane per key), probability-weighted sum of the memory rows.  Per-lane state: 8 heads x 8 ch.
// S head (~64 keys / query): 4 warps, 16-key chunks, 3 CTAs / SM.  T head (~2000 keys / query): 8 warps, 32-key chunks.
__device__ __forceinline__ void xa_cp16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}

__device__ inline void reduce8(float (&v)[8], int lane) {
    // transpose-reduce 8 per-lane partials over the warp: 9 shuffles instead of 40.
    // result: every lane holds the full sum for head (lane >> 2) & 7 in v[0].
    const bool up16 = lane & 16, up8 = lane & 8, up4 = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float send = up16 ? v[i] : v[i + 4], keep = up16 ? v[i + 4] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        float send = up8 ? v[i] : v[i + 2], keep = up8 ? v[i + 2] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    {
        float send = up4 ? v[0] : v[1], keep = up4 ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

struct XaArgs {
    const float* qt; const float* kin_rows; const float* mem_rows;
    const int* match; const int* match_cnt; int max_match;
    const uint32_t* keymask; int mask_words;
    const uint16_t* key_list; const int* key_cnt;    // mode 1: compacted key ids from box_corr (nullptr: compact here)
    int mode; int N; int klist_cap;
    float* ctx; float* ctx_lo;    // ctx_lo != nullptr: write the TF32 hi/lo split (operands of the 3xTF32 output GEMM)
};

template <int XA_CH, int XA_THREADS>
__device__ __forceinline__ void cross_attn_body(const XaArgs& a, int n, unsigned char* smem_raw) {
    static_assert(XA_CH == XA_THREADS / 8, "one softmax lane per key: CH = 4 * warps");
    float* kbuf = reinterpret_cast<float*>(smem_raw);               // [2][XA_CH][256] key-input rows
    float* vbuf = kbuf + 2 * XA_CH * MV2D_C;                        // [2][XA_CH][256] memory rows
    float* sc = vbuf + 2 * XA_CH * MV2D_C;                          // [XA_CH][8] logits -> probs
    float* stat = sc + XA_CH * 8;                                   // m[8], l[8], alpha[8]
    const uint16_t* klist = reinterpret_cast<uint16_t*>(stat + 24); // [klist_cap] (or the global list of box_corr)
    uint16_t* klist_s = reinterpret_cast<uint16_t*>(stat + 24);
    __shared__ int nkeys_s;
    __shared__ int grp_cnt[128];
    constexpr int NW = XA_THREADS / 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = threadIdx.x;

    // ---- key list
    if (t == 0) nkeys_s = 0;
    if (t < 8) { stat[t] = -INFINITY; stat[8 + t] = 0.f; stat[16 + t] = 1.f; }
    __syncthreads();
    if (a.mode == 0) {
        const int cnt = a.match_cnt[n];
        for (int i = t; i < cnt * MV2D_TOK; i += XA_THREADS)
            klist_s[i] = (uint16_t)(a.match[(long long)n * a.max_match + i / MV2D_TOK] * MV2D_TOK + i % MV2D_TOK);
        if (t == 0) nkeys_s = cnt * MV2D_TOK;
    } else if (a.key_list) {
        klist = a.key_list + (long long)n * a.mask_words * 32;
        if (t == 0) nkeys_s = a.key_cnt[n];
    } else {
        // deterministic compaction of the set bits: per-32-word group counts, then a prefix
        const uint32_t* km = a.keymask + (long long)n * a.mask_words;
        const int ngroups = (a.mask_words + 31) / 32;           // host guarantees <= 128
        for (int g = warp; g < ngroups; g += NW) {
            const int w = g * 32 + lane;
            const int c = __popc((w < a.mask_words) ? km[w] : 0u);
            const int tot = __reduce_add_sync(0xffffffffu, c);
            if (lane == 0) grp_cnt[g] = tot;
        }
        __syncthreads();
        for (int g = warp; g < ngroups; g += NW) {
            int base = 0;
            for (int i = 0; i < g; ++i) base += grp_cnt[i];
            const int w = g * 32 + lane;
            const uint32_t bits = (w < a.mask_words) ? km[w] : 0u;
            const int c = __popc(bits);
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
            int pos = base + incl - c;
            uint32_t b = bits;
            while (b) { const int bit = __ffs(b) - 1; b &= b - 1; klist_s[pos++] = (uint16_t)(w * 32 + bit); }
            if (g == ngroups - 1 && lane == 31) nkeys_s = base + incl;
        }
    }
    __syncthreads();
    const int nkeys = nkeys_s;
    const int nchunks = (nkeys + XA_CH - 1) / XA_CH;

    // stage chunk c into buffer c&1: 16 keys x (1 KB + 1 KB) = 2048 16-byte copies, 16 per thread
    auto stage = [&](int c) {
        const int base = c * XA_CH, cn = min(XA_CH, nkeys - base), buf = c & 1;
        for (int i = t; i < cn * 64; i += XA_THREADS) {
            const int j = i >> 6, q4 = i & 63;
            const long long row = (long long)klist[base + j] * MV2D_C + q4 * 4;
            xa_cp16(kbuf + (buf * XA_CH + j) * MV2D_C + q4 * 4, a.kin_rows + row);
            xa_cp16(vbuf + (buf * XA_CH + j) * MV2D_C + q4 * 4, a.mem_rows + row);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (nchunks > 0) stage(0);

    // ---- q~ slice of this lane: 8 heads x channels {lane*4..+3, 128+lane*4..+3}
    float qv[8][8];
#pragma unroll
    for (int h = 0; h < 8; ++h) {
        const float4 x0 = *reinterpret_cast<const float4*>(a.qt + (long long)n * 2048 + h * 256 + lane * 4);
        const float4 x1 = *reinterpret_cast<const float4*>(a.qt + (long long)n * 2048 + h * 256 + 128 + lane * 4);
        qv[h][0] = x0.x; qv[h][1] = x0.y; qv[h][2] = x0.z; qv[h][3] = x0.w;
        qv[h][4] = x1.x; qv[h][5] = x1.y; qv[h][6] = x1.z; qv[h][7] = x1.w;
    }
    float acc[8][8];
#pragma unroll
    for (int h = 0; h < 8; ++h)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[h][c] = 0.f;

    for (int c = 0; c < nchunks; ++c) {
        const int cn = min(XA_CH, nkeys - c * XA_CH), buf = c & 1;
        if (c + 1 < nchunks) { stage(c + 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        // pass 1: logits, 4 keys per warp
        for (int j = warp; j < cn; j += NW) {
            const float* row = kbuf + (buf * XA_CH + j) * MV2D_C;
            const float4 k0 = *reinterpret_cast<const float4*>(row + lane * 4);
            const float4 k1 = *reinterpret_cast<const float4*>(row + 128 + lane * 4);
            float sv[8];
#pragma unroll
            for (int h = 0; h < 8; ++h) {
                float x = qv[h][0] * k0.x;
                x = fmaf(qv[h][1], k0.y, x); x = fmaf(qv[h][2], k0.z, x); x = fmaf(qv[h][3], k0.w, x);
                x = fmaf(qv[h][4], k1.x, x); x = fmaf(qv[h][5], k1.y, x); x = fmaf(qv[h][6], k1.z, x);
                x = fmaf(qv[h][7], k1.w, x);
                sv[h] = x;
            }
            reduce8(sv, lane);
            if ((lane & 3) == 0) sc[j * 8 + (lane >> 2)] = sv[0];
        }
        __syncthreads();
        // online softmax: a warp owns 8/NW heads, XA_CH lanes each (lane % XA_CH = key)
        {
            const int h = warp + (lane / XA_CH) * NW, key = lane % XA_CH;
            const float sv = key < cn ? sc[key * 8 + h] : -INFINITY;
            float mx = sv;
#pragma unroll
            for (int o = XA_CH / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            const float m_old = stat[h], m_new = fmaxf(m_old, mx);
            const float p = key < cn ? __expf(sv - m_new) : 0.f;
            float sum = p;
#pragma unroll
            for (int o = XA_CH / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            __syncwarp();
            if (key < cn) sc[key * 8 + h] = p;
            if (key == 0) {
                const float alpha = (m_old == -INFINITY) ? 0.f : __expf(m_old - m_new);
                stat[16 + h] = alpha;
                stat[8 + h] = stat[8 + h] * alpha + sum;
                stat[h] = m_new;
            }
        }
        __syncthreads();
        // pass 2: acc = acc*alpha + sum_k p_k * mem_k
#pragma unroll
        for (int h = 0; h < 8; ++h) {
            const float al = stat[16 + h];
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) acc[h][cc] *= al;
        }
        for (int j = warp; j < cn; j += NW) {
            const float* row = vbuf + (buf * XA_CH + j) * MV2D_C;
            const float4 v0 = *reinterpret_cast<const float4*>(row + lane * 4);
            const float4 v1 = *reinterpret_cast<const float4*>(row + 128 + lane * 4);
            const float4 p0 = *reinterpret_cast<const float4*>(sc + j * 8);
            const float4 p1 = *reinterpret_cast<const float4*>(sc + j * 8 + 4);
            const float pp[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
            for (int h = 0; h < 8; ++h) {
                acc[h][0] = fmaf(pp[h], v0.x, acc[h][0]); acc[h][1] = fmaf(pp[h], v0.y, acc[h][1]);
                acc[h][2] = fmaf(pp[h], v0.z, acc[h][2]); acc[h][3] = fmaf(pp[h], v0.w, acc[h][3]);
                acc[h][4] = fmaf(pp[h], v1.x, acc[h][4]); acc[h][5] = fmaf(pp[h], v1.y, acc[h][5]);
                acc[h][6] = fmaf(pp[h], v1.z, acc[h][6]); acc[h][7] = fmaf(pp[h], v1.w, acc[h][7]);
            }
        }
        __syncthreads();   // everyone done with buffer `buf` and `sc` before they are refilled
    }
    // ---- cross-warp tree sum through shared memory (fixed order => bitwise reproducible):
    // upper half of the warps -> lower half, repeatedly; warp 0 normalises and stores.
    float4* T = reinterpret_cast<float4*>(kbuf);   // [NW/2][512] float4, the staging buffers are free now
    auto put = [&](int slot) {
#pragma unroll
        for (int h = 0; h < 8; ++h) {
            T[slot * 512 + h * 64 + lane] = make_float4(acc[h][0], acc[h][1], acc[h][2], acc[h][3]);
            T[slot * 512 + h * 64 + 32 + lane] = make_float4(acc[h][4], acc[h][5], acc[h][6], acc[h][7]);
        }
    };
    auto add = [&](int slot) {
#pragma unroll
        for (int h = 0; h < 8; ++h) {
            const float4 x0 = T[slot * 512 + h * 64 + lane], x1 = T[slot * 512 + h * 64 + 32 + lane];
            acc[h][0] += x0.x; acc[h][1] += x0.y; acc[h][2] += x0.z; acc[h][3] += x0.w;
            acc[h][4] += x1.x; acc[h][5] += x1.y; acc[h][6] += x1.z; acc[h][7] += x1.w;
        }
    };
#pragma unroll
    for (int stride = NW / 2; stride >= 1; stride >>= 1) {
        if (warp >= stride && warp < 2 * stride) put(warp - stride);
        __syncthreads();
        if (warp < stride) add(warp);
        __syncthreads();
    }
    if (warp == 0) {
#pragma unroll
        for (int h = 0; h < 8; ++h) {
            const float l = stat[8 + h], inv = l > 0.f ? 1.f / l : 0.f;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) v[k] = acc[h][half * 4 + k] * inv;
                const long long o = (long long)n * 2048 + h * 256 + half * 128 + lane * 4;
                if (a.ctx_lo) {
                    float hi[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) { hi[k] = round_tf32(v[k]); v[k] = round_tf32(v[k] - hi[k]); }
                    *reinterpret_cast<float4*>(a.ctx + o) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4*>(a.ctx_lo + o) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
                    *reinterpret_cast<float4*>(a.ctx + o) = make_float4(v[0], v[1], v[2], v[3]);
                }
            }
        }
    }
}

template <int XA_CH, int XA_THREADS>
__global__ void __launch_bounds__(XA_THREADS)
cross_attn_kernel(XaArgs a) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ __align__(16) unsigned char xa_smem_raw[];
    cross_attn_body<XA_CH, XA_THREADS>(a, blockIdx.x, xa_smem_raw);
}

// ------------------------------------------------------------------------------------------
// S head cross-attention, one CTA per (query, matched RoI).  The 49 tokens of a RoI are CONTIGUOUS in
// tok_kin / tok_feat, so one elected thread fetches each 50 176-byte block with a single bulk copy
// (cp.async.bulk + mbarrier); every unit of work is the same size, which removes the tail the per-query kernel
// has (a query with 6 matches streams 6x the keys of the median query and set the kernel time).
// A query with one match (~3 of 4) is finished by its CTA.  Otherwise each CTA leaves its un-normalised
// (acc, m, l) record in global memory and the LAST one to arrive (atomic ticket) merges the records in slot
// order -- fixed order => bitwise reproducible whichever CTA does the merge -- and re-arms the ticket.
#define XR_THREADS 256
#define XR_BLK_BYTES (MV2D_TOK * MV2D_C * 4)
#define XR_REC (2048 + 16)        // floats per partial record: acc[8][256], m[8], l[8]
#define XR_MAXM 8                 // match slots the partial-record scratch is sized for
#define XR_SMEM_BYTES (2 * XR_BLK_BYTES + 64 * 8 * 4 + 32 * 4 + 16)

struct XrArgs {
    const float* qt; const float* kin_rows; const float* mem_rows;
    const int* match; const int* match_cnt; int max_match;
    float* ctx; float* ctx_lo; float* part; int* ticket;
    const int* units;            // [total] (query << 3 | slot), query-major, from xr_units_kernel; units[-1..]: see XR_UNITS_HDR
    const int* total;            // device: number of units
};

// Work list of xa_roi_kernel: one unit per (query, matched RoI), query-major.  Built once per decoder call (the match
// lists do not change between layers).  One block.
#define XR_UNITS_THREADS 1024
__global__ void __launch_bounds__(XR_UNITS_THREADS) xr_units_kernel(const int* __restrict__ match_cnt, int N, int* __restrict__ units,
                                                                    int* __restrict__ total) {
    pdl_wait();
    pdl_trigger();
    __shared__ int wsum[XR_UNITS_THREADS / 32];
    __shared__ int base_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) base_s = 0;
    __syncthreads();
    for (int n0 = 0; n0 < N; n0 += XR_UNITS_THREADS) {
        const int n = n0 + tid;
        const int c = n < N ? min(match_cnt[n], 8) : 0;
        int incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = wsum[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, d); if (lane >= d) w += y; }
            wsum[lane] = w;
        }
        __syncthreads();
        const int off = base_s + (warp ? wsum[warp - 1] : 0) + incl - c;
        for (int s2 = 0; s2 < c; ++s2) units[off + s2] = (n << 3) | s2;
        __syncthreads();
        if (tid == 0) base_s += wsum[XR_UNITS_THREADS / 32 - 1];
        __syncthreads();
    }
    if (tid == 0) *total = base_s;
}

__device__ __forceinline__ uint32_t xr_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// The two inner loops of xa_roi_kernel are bound by instruction issue, not by the FMA pipe: pairing the channels (f32x2,
// common.cuh) halves their FMA count.

// one (query n, match slot) unit; `phase` = parity of this use of the CTA's mbarrier
__device__ __forceinline__ void xr_unit(const XrArgs& a, int n, int slot, int cnt, uint32_t phase, unsigned char* xr_smem) {
    float* Ks = reinterpret_cast<float*>(xr_smem);
    float* Vs = Ks + MV2D_TOK * MV2D_C;
    float* sc = Vs + MV2D_TOK * MV2D_C;          // [64][8] logits -> probabilities
    float* stat = sc + 64 * 8;                   // m[8], l[8]
    uint64_t* bar = reinterpret_cast<uint64_t*>(stat + 32);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    constexpr int NW = XR_THREADS / 32;
    const long long blk = (long long)a.match[(long long)n * a.max_match + slot] * MV2D_TOK * MV2D_C;
    if (t == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the previous unit's generic-proxy use of Ks / Vs is ordered before these copies
        // bar[0]: the key block has landed, bar[1]: the value block (the logits do not wait for the values)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(xr_smem_u32(bar)), "r"(XR_BLK_BYTES) : "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(xr_smem_u32(bar + 1)), "r"(XR_BLK_BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(xr_smem_u32(Ks)), "l"(a.kin_rows + blk), "r"(XR_BLK_BYTES), "r"(xr_smem_u32(bar)) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(xr_smem_u32(Vs)), "l"(a.mem_rows + blk), "r"(XR_BLK_BYTES), "r"(xr_smem_u32(bar + 1)) : "memory");
    }
    {
        // ---- logits.  q~ slice of this lane: 8 heads x channels {lane*4..+3, 128+lane*4..+3}
        f32x2 qv[8][4];             // channel pairs (4 lane, 4 lane + 1), (+2, +3), (128 + 4 lane, ..), (.. + 2, + 3)
#pragma unroll
        for (int h = 0; h < 8; ++h) {
            const ulonglong2 x0 = *reinterpret_cast<const ulonglong2*>(a.qt + (long long)n * 2048 + h * 256 + lane * 4);
            const ulonglong2 x1 = *reinterpret_cast<const ulonglong2*>(a.qt + (long long)n * 2048 + h * 256 + 128 + lane * 4);
            qv[h][0] = x0.x; qv[h][1] = x0.y; qv[h][2] = x1.x; qv[h][3] = x1.y;
        }
        {
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(xr_smem_u32(bar)), "r"(phase) : "memory");
        }
        // two rows per trip (j and j + NW): the five dependent shuffle stages of the transpose-reduce of one row run under
        // the FMAs / shuffles of the other (at 16 warps per SM the kernel is bound by exactly these dependent chains)
        for (int j = warp; j < MV2D_TOK; j += 2 * NW) {
            const bool two = j + NW < MV2D_TOK;
            const float* row = Ks + j * MV2D_C;
            const float* rowb = Ks + (two ? j + NW : j) * MV2D_C;
            const ulonglong2 k0 = *reinterpret_cast<const ulonglong2*>(row + lane * 4);
            const ulonglong2 k1 = *reinterpret_cast<const ulonglong2*>(row + 128 + lane * 4);
            const ulonglong2 m0 = *reinterpret_cast<const ulonglong2*>(rowb + lane * 4);
            const ulonglong2 m1 = *reinterpret_cast<const ulonglong2*>(rowb + 128 + lane * 4);
            float sv[8], sw[8];
#pragma unroll
            for (int h = 0; h < 8; ++h) {
                f32x2 x = mul2(qv[h][0], k0.x), y = mul2(qv[h][0], m0.x);
                x = fma2(qv[h][1], k0.y, x); y = fma2(qv[h][1], m0.y, y);
                x = fma2(qv[h][2], k1.x, x); y = fma2(qv[h][2], m1.x, y);
                x = fma2(qv[h][3], k1.y, x); y = fma2(qv[h][3], m1.y, y);
                float lo, hi;
                unpack2(x, lo, hi);
                sv[h] = lo + hi;
                unpack2(y, lo, hi);
                sw[h] = lo + hi;
            }
            reduce8(sv, lane);
            reduce8(sw, lane);
            if ((lane & 3) == 0) {
                sc[j * 8 + (lane >> 2)] = sv[0];
                if (two) sc[(j + NW) * 8 + (lane >> 2)] = sw[0];
            }
        }
    }
    __syncthreads();
    {   // ---- softmax statistics of this block: warp = head, lanes = keys {lane, lane + 32}
        const int h = warp;
        const float s0 = sc[lane * 8 + h];
        const float s1 = lane + 32 < MV2D_TOK ? sc[(lane + 32) * 8 + h] : -INFINITY;
        const float mx = warp_max(fmaxf(s0, s1));
        const float p0 = __expf(s0 - mx), p1 = lane + 32 < MV2D_TOK ? __expf(s1 - mx) : 0.f;
        const float sum = warp_sum(p0 + p1);
        sc[lane * 8 + h] = p0;
        if (lane + 32 < MV2D_TOK) sc[(lane + 32) * 8 + h] = p1;
        if (lane == 0) { stat[h] = mx; stat[8 + h] = sum; }
    }
    __syncthreads();
    // ---- acc = sum_k p_k * mem_k
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(xr_smem_u32(bar + 1)), "r"(phase) : "memory");
    }
    float acc[8][8];
    {
        f32x2 acc2[8][4];
#pragma unroll
        for (int h = 0; h < 8; ++h)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc2[h][c] = 0ull;
        for (int j = warp; j < MV2D_TOK; j += NW) {
            const float* row = Vs + j * MV2D_C;
            const ulonglong2 v0 = *reinterpret_cast<const ulonglong2*>(row + lane * 4);
            const ulonglong2 v1 = *reinterpret_cast<const ulonglong2*>(row + 128 + lane * 4);
            const float4 p0 = *reinterpret_cast<const float4*>(sc + j * 8);
            const float4 p1 = *reinterpret_cast<const float4*>(sc + j * 8 + 4);
            const float pp[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
            for (int h = 0; h < 8; ++h) {
                const f32x2 p2 = pack2(pp[h], pp[h]);
                acc2[h][0] = fma2(p2, v0.x, acc2[h][0]); acc2[h][1] = fma2(p2, v0.y, acc2[h][1]);
                acc2[h][2] = fma2(p2, v1.x, acc2[h][2]); acc2[h][3] = fma2(p2, v1.y, acc2[h][3]);
            }
        }
#pragma unroll
        for (int h = 0; h < 8; ++h)
#pragma unroll
            for (int c = 0; c < 4; ++c) unpack2(acc2[h][c], acc[h][2 * c], acc[h][2 * c + 1]);
    }
    // ---- cross-warp sum, fixed order: every warp parks its partial (8 heads x 256) in the now free K / V blocks, then
    // warp h adds up head h's eight partials and finishes that head (normalise, store / record, merge)
    __syncthreads();                             // all warps are done reading Vs and sc
    float4* T = reinterpret_cast<float4*>(Ks);   // [NW][8 heads][64] float4 = 64 KB of the 98 KB
#pragma unroll
    for (int h = 0; h < 8; ++h) {
        T[(warp * 8 + h) * 64 + lane] = make_float4(acc[h][0], acc[h][1], acc[h][2], acc[h][3]);
        T[(warp * 8 + h) * 64 + 32 + lane] = make_float4(acc[h][4], acc[h][5], acc[h][6], acc[h][7]);
    }
    __syncthreads();
    const int h = warp;                          // NW == 8 == heads
    float o[8];
    {
        float4 x0 = T[h * 64 + lane], x1 = T[h * 64 + 32 + lane];
#pragma unroll
        for (int w = 1; w < NW; ++w) {
            const float4 y0 = T[(w * 8 + h) * 64 + lane], y1 = T[(w * 8 + h) * 64 + 32 + lane];
            x0.x += y0.x; x0.y += y0.y; x0.z += y0.z; x0.w += y0.w;
            x1.x += y1.x; x1.y += y1.y; x1.z += y1.z; x1.w += y1.w;
        }
        o[0] = x0.x; o[1] = x0.y; o[2] = x0.z; o[3] = x0.w; o[4] = x1.x; o[5] = x1.y; o[6] = x1.z; o[7] = x1.w;
    }
    float inv;
    if (cnt == 1) {
        inv = 1.f / stat[8 + h];
    } else {
        // leave the un-normalised record; the last CTA of this query merges all of them in slot order
        float* rec = a.part + ((long long)n * XR_MAXM + slot) * XR_REC;
        *reinterpret_cast<float4*>(rec + h * 256 + lane * 4) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(rec + h * 256 + 128 + lane * 4) = make_float4(o[4], o[5], o[6], o[7]);
        if (lane == 0) { rec[2048 + h] = stat[h]; rec[2056 + h] = stat[8 + h]; }
        __threadfence();
        __syncthreads();                         // the whole record is written (and fenced) before the ticket is taken
        int* last = reinterpret_cast<int*>(stat + 16);
        if (t == 0) *last = atomicAdd(a.ticket + n, 1) == cnt - 1;
        __syncthreads();
        if (!*last) return;                      // CTA-uniform
        __threadfence();
        const float* base = a.part + (long long)n * XR_MAXM * XR_REC;
        float M = -INFINITY, Lsum = 0.f;
        for (int s2 = 0; s2 < cnt; ++s2) M = fmaxf(M, __ldcg(base + s2 * XR_REC + 2048 + h));
#pragma unroll
        for (int c = 0; c < 8; ++c) o[c] = 0.f;
        for (int s2 = 0; s2 < cnt; ++s2) {
            const float* r2 = base + s2 * XR_REC;
            const float w = __expf(__ldcg(r2 + 2048 + h) - M);
            Lsum = fmaf(__ldcg(r2 + 2056 + h), w, Lsum);
            const float4 x0 = __ldcg(reinterpret_cast<const float4*>(r2 + h * 256 + lane * 4));
            const float4 x1 = __ldcg(reinterpret_cast<const float4*>(r2 + h * 256 + 128 + lane * 4));
            o[0] = fmaf(x0.x, w, o[0]); o[1] = fmaf(x0.y, w, o[1]); o[2] = fmaf(x0.z, w, o[2]); o[3] = fmaf(x0.w, w, o[3]);
            o[4] = fmaf(x1.x, w, o[4]); o[5] = fmaf(x1.y, w, o[5]); o[6] = fmaf(x1.z, w, o[6]); o[7] = fmaf(x1.w, w, o[7]);
        }
        inv = 1.f / Lsum;
        if (t == 0) a.ticket[n] = 0;             // re-armed for the next layer (stream order)
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = o[half * 4 + k] * inv;
        const long long oo = (long long)n * 2048 + h * 256 + half * 128 + lane * 4;
        if (a.ctx_lo) {
            float hi[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { hi[k] = round_tf32(v[k]); v[k] = round_tf32(v[k] - hi[k]); }
            *reinterpret_cast<float4*>(a.ctx + oo) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(a.ctx_lo + oo) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
            *reinterpret_cast<float4*>(a.ctx + oo) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

// Persistent: 2 CTAs per SM walk the unit list (no CTA is launched for the ~80 % of (query, slot) pairs that do not
// exist, and while one CTA of an SM computes the other one's copies are in flight).
__global__ void __launch_bounds__(XR_THREADS, 2)
xa_roi_kernel(XrArgs a) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ __align__(128) unsigned char xr_smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(reinterpret_cast<float*>(xr_smem) + 2 * MV2D_TOK * MV2D_C + 64 * 8 + 32);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(xr_smem_u32(bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(xr_smem_u32(bar + 1)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();            // barriers initialised before anyone arms or polls them
    const int total = *a.total;
    uint32_t phase = 0;
    for (int u = blockIdx.x; u < total; u += gridDim.x, phase ^= 1u) {
        const int code = a.units[u], n = code >> 3, slot = code & 7;
        xr_unit(a, n, slot, min(a.match_cnt[n], XR_MAXM), phase, xr_smem);
        __syncthreads();        // every warp is done with this unit's shared memory
    }
}

// ------------------------------------------------------------------------------------------
// Final 256 -> 10 heads of both branches + reference-point refinement
// (cross_attention_head.py:221-238).  One warp per (layer, query).
__device__ __forceinline__ void
head10_body(const float* __restrict__ xc, const float* __restrict__ xr, const float* __restrict__ wc,
            const float* __restrict__ bc, const float* __restrict__ wr, const float* __restrict__ br,
            const float* __restrict__ ref, int L, int N, float pc0, float pc1, float pc2, float pc3,
            float pc4, float pc5, float vel_dt, int vel_row_start, float* __restrict__ cls, float* __restrict__ box, int vb,
            const float* __restrict__ vel_dt_batch = nullptr, int rows_per_sample = 0) {
    const int row = vb * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= L * N) return;
    const int l = row / N, n = row % N;
    if (vel_dt_batch) vel_dt = __ldg(vel_dt_batch + n / rows_per_sample);
    float a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = xc[(long long)row * MV2D_C + i * 32 + lane];
        b[i] = xr[(long long)row * MV2D_C + i * 32 + lane];
    }
    for (int o = 0; o < 10; ++o) {
        const float* w1 = wc + ((long long)l * 10 + o) * MV2D_C;
        const float* w2 = wr + ((long long)l * 10 + o) * MV2D_C;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s1 = fmaf(a[i], __ldg(w1 + i * 32 + lane), s1);
            s2 = fmaf(b[i], __ldg(w2 + i * 32 + lane), s2);
        }
        s1 = warp_sum(s1) + __ldg(bc + l * 10 + o);
        s2 = warp_sum(s2) + __ldg(br + l * 10 + o);
        if (lane == 0) {
            cls[(long long)row * 10 + o] = s1;
            if (o == 0) s2 = sigmoid_f(s2 + inverse_sigmoid_f(ref[n * 3 + 0])) * (pc3 - pc0) + pc0;
            else if (o == 1) s2 = sigmoid_f(s2 + inverse_sigmoid_f(ref[n * 3 + 1])) * (pc4 - pc1) + pc1;
            else if (o == 4) s2 = sigmoid_f(s2 + inverse_sigmoid_f(ref[n * 3 + 2])) * (pc5 - pc2) + pc2;
            else if (o >= 8 && vel_dt != 0.f && n >= vel_row_start) s2 = s2 / vel_dt;
            box[(long long)row * 10 + o] = s2;
        }
    }
}

__global__ void __launch_bounds__(256)
head10_kernel(const float* __restrict__ xc, const float* __restrict__ xr, const float* __restrict__ wc,
              const float* __restrict__ bc, const float* __restrict__ wr, const float* __restrict__ br,
              const float* __restrict__ ref, int L, int N, float pc0, float pc1, float pc2, float pc3,
              float pc4, float pc5, float vel_dt, int vel_row_start, float* __restrict__ cls, float* __restrict__ box,
              const float* __restrict__ vel_dt_batch, int rows_per_sample) {
    pdl_wait();
    pdl_trigger();
    head10_body(xc, xr, wc, bc, wr, br, ref, L, N, pc0, pc1, pc2, pc3, pc4, pc5, vel_dt, vel_row_start, cls, box, blockIdx.x,
                vel_dt_batch, rows_per_sample);
}

}  // namespace mv2d
#include "decoder_mega.cuh"
#include "xa_tile.cuh"
#include "sa_mma.cuh"
namespace mv2d {

// ------------------------------------------------------------------------------------------
static int gemm(const float* A, int lda, const float* W, int ldw, const float* bias, float* C, int ldc, int M, int N,
                int K, int flags, cudaStream_t st, int nsplit = 1, long long split_stride = 0, int batch = 1,
                long long sA = 0, long long sW = 0, long long sC = 0, long long sB = 0) {
    GemmArgs g{};
    g.A = A; g.lda = lda; g.W = W; g.ldw = ldw; g.C = C; g.ldc = ldc; g.bias = bias;
    g.M = M; g.N = N; g.K = K; g.batch = batch; g.nsplit = nsplit; g.splitStride = split_stride; g.flags = flags;
    g.strideA = sA; g.strideW = sW; g.strideC = sC; g.strideBias = sB;
    return launch_gemm_simt(g, A_PLAIN, st);
}

// 3xTF32 tcgen05 GEMM on pre-split operands (the four wide products of a decoder layer)
static int tc3(const float* A_hi, const float* A_lo, int lda, const float* W_hi, const float* W_lo, int ldw,
               const float* bias, float* Cc, float* C_lo, int ldc, int M, int N, int K, int flags, int nsplit,
               long long split_stride, cudaStream_t st) {
    TcGemm t{};
    t.A = A_hi; t.A_lo = A_lo; t.lda = lda; t.W = W_hi; t.W_lo = W_lo; t.ldw = ldw; t.bias = bias;
    t.C = Cc; t.C_lo = C_lo; t.ldc = ldc; t.M = M; t.N = N; t.K = K; t.passes = 3; t.im2col = 0; t.flags = flags;
    t.nsplit = nsplit; t.split_stride = split_stride;
    return launch_gemm_tc(t, st);
}

static int ln(const LnArgs& a, cudaStream_t st) {
    if (a.rows == 0) return 0;
    MV2D_CHECK_ARG(a.nsplit >= 1 && a.nsplit <= 8, "ln: nsplit=%d must be in [1,8]", a.nsplit);
    launch_k(ln_kernel, dim3(cdiv(a.rows, 8)), dim3(256), 0, st, a);
    MV2D_CHECK_LAUNCH("ln");
    return 0;
}

#define DEC_SPLIT 8
static size_t xa_smem_mega(int mode, int klist_cap) {
    (void)mode;
    return (size_t)(4 * 32 * MV2D_C + 32 * 8 + 24) * sizeof(float) + (size_t)klist_cap * sizeof(uint16_t);
}

size_t decoder_workspace_bytes(int N, int L) {
    size_t n = (size_t)(N > 0 ? N : 1), l = (size_t)(L > 0 ? L : 1);
    // x, xq, x1, x1q, x2 + 4 hi/lo copies (9*256) + qkv 768 + sa 256 + qt 2048 + ctx hi/lo + hdn hi/lo + partials
    // + (3xTF32 everywhere, > 512 rows) x / xq hi/lo, sa lo, xt ctx lo: 6*256
    size_t per = 9 * MV2D_C + 768 + MV2D_C + 2048 + 2 * 2048 + 2 * 2048 + DEC_SPLIT * MV2D_C + XR_MAXM * XR_REC + 1 + XR_MAXM + 6 * MV2D_C;   // + ticket + unit list
    // branches: 4 x [L,N,256] + hi/lo splits of the post-normed states and of three branch activations
    return (n * per + (4 + 7) * l * n * MV2D_C) * sizeof(float) + 4096 + 512;   // + the device-wide barrier word and phase timestamps of the persistent kernel
}

// ---- key-stationary cross-attention of the two-frame head (xa_tile.cuh): caller-owned scratch
struct XtWs {
    int* tile_cnt; int* tile_work; int* order; uint16_t* tile_q; unsigned long long* tile_mask; short* slot_of;
    int* qlist; int* qcnt; float* qp; float* ctx; float* rec;
    size_t bytes;
};
// Np query rows and tiles_ps tiles per sample, batch samples (every (tile, query) pair lives inside one sample)
static XtWs xt_carve(void* base, int Np, int tiles_ps, int batch = 1) {
    XtWs w{};
    size_t off = 0;
    auto take = [&](size_t bytes) { void* r = base ? (char*)base + off : nullptr; off += (bytes + 255) & ~(size_t)255; return r; };
    const size_t bb = (size_t)(batch > 0 ? batch : 1);
    const size_t np = (size_t)(Np > 0 ? Np : 1), n = bb * np, tp = (size_t)(tiles_ps > 0 ? tiles_ps : 1), t = bb * tp;
    w.tile_cnt = (int*)take(t * sizeof(int));
    w.tile_work = (int*)take(t * sizeof(int));
    w.order = (int*)take(t * sizeof(int));
    w.qlist = (int*)take(tp * n * sizeof(int));
    w.qcnt = (int*)take(n * sizeof(int));
    w.tile_q = (uint16_t*)take(t * np * sizeof(uint16_t));
    w.tile_mask = (unsigned long long*)take(t * np * sizeof(unsigned long long));
    w.slot_of = (short*)take(tp * n * sizeof(short));
    w.qp = (float*)take(n * MV2D_C * sizeof(float));
    w.ctx = (float*)take(n * MV2D_C * sizeof(float));
    w.rec = (float*)take(t * np * XT_REC * sizeof(float));
    w.bytes = off;
    return w;
}
static bool xt_mma_enabled() {
    static const bool on = []() { const char* v = getenv("MV2D_XT_MMA"); return !(v && v[0] == '0'); }();
    return on;
}
// per-tile attention of the key-stationary form: TF32 tensor-core kernel (16 queries per MMA row block), or the FFMA
// kernel (a warp per query) with MV2D_XT_MMA=0
static int launch_xt_attn(const XtAttnArgs& a, int ntiles, cudaStream_t st) {
    if (xt_mma_enabled()) {
        static bool attr_set = false;
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(xt_attn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XTM_SMEM_BYTES);
            if (e != cudaSuccess) { set_error("xt_attn: smem attr %s", cudaGetErrorString(e)); return (int)e; }
            attr_set = true;
        }
        launch_k(xt_attn_mma_kernel, dim3(ntiles, a.qsplit), dim3(XTM_THREADS), (size_t)XTM_SMEM_BYTES, st, a);
    } else {
        launch_k(xt_attn_kernel, dim3(ntiles, a.qsplit), dim3(XT_THREADS), (size_t)XT_SMEM_BYTES, st, a);
    }
    MV2D_CHECK_LAUNCH("xt_attn");
    return 0;
}

static int xt_ntiles(int V, int h, int w) { return V * cdiv(h, XT_TS) * cdiv(w, XT_TS); }

size_t xa_tile_workspace_bytes(int N, int V, int h, int w, int batch) {
    return xt_carve(nullptr, N, xt_ntiles(V, h, w), batch > 0 ? batch : 1).bytes;
}

int run_kv_project(const Mv2dKvParams& p, cudaStream_t st) {
    const int le = p.layer_end > 0 ? p.layer_end : p.L;
    MV2D_CHECK_ARG(p.num_rows > 0 && p.L >= 1 && p.L <= MV2D_MAX_LAYERS && p.layer_begin >= 0 && p.layer_begin < le && le <= p.L,
                   "kv_project: bad num_rows=%d / layers [%d,%d) of %d", p.num_rows, p.layer_begin, le, p.L);
    const long long RC = (long long)p.num_rows * MV2D_C;
    // pre-split rows + weights stacked by mv2d_pack_weights: one persistent launch for all the layers asked for (kvproj.cu)
    if (kv_persistent_usable(p)) return run_kv_project_persistent(p, st);
    for (int l = p.layer_begin; l < le; ++l) {
        const Mv2dLayerWeights& w = p.layers[l];
        const bool raw = p.kin_lo == nullptr;      // plain fp32 rows: the GEMM splits rows and weights in shared memory
        MV2D_CHECK_ARG(w.xa_k_w && w.xa_v_w && (raw ? (w.xa_k_raw && w.xa_v_raw) : (w.xa_k_w_lo && w.xa_v_w_lo)),
                       "kv_project: layer %d has no xa_k / xa_v weights", l);
        int rc;
        TcGemm t{};
        t.lda = MV2D_C; t.ldw = MV2D_C; t.ldc = MV2D_C; t.M = p.num_rows; t.N = MV2D_C; t.K = MV2D_C; t.passes = 3; t.nsplit = 1;
        t.m_tile_live = p.row_tile_live;
        t.A = p.kin_hi; t.A_lo = p.kin_lo; t.W = raw ? w.xa_k_raw : w.xa_k_w; t.W_lo = raw ? nullptr : w.xa_k_w_lo; t.C = p.kp + l * RC;
        if ((rc = launch_gemm_tc(t, st))) return rc;
        t.A = p.mem_hi; t.A_lo = p.mem_lo; t.W = raw ? w.xa_v_raw : w.xa_v_w; t.W_lo = raw ? nullptr : w.xa_v_w_lo; t.C = p.vp + l * RC;
        if ((rc = launch_gemm_tc(t, st))) return rc;
    }
    return 0;
}

static int xt_setup(const Mv2dDecoderParams& p, XtGeom& xg, XtWs& xw) {
    const int N = p.N;
    MV2D_CHECK_ARG(p.grid_h > 0 && p.grid_w > 0 && p.num_rows > 0 && p.num_rows % (p.grid_h * p.grid_w) == 0,
                   "decoder: xa_form 1 needs the feature grid (num_rows=%d, grid %dx%d)", p.num_rows, p.grid_h, p.grid_w);
    const int B = p.batch > 0 ? p.batch : 1;
    MV2D_CHECK_ARG(p.num_rows % B == 0 && (p.batch == 0 || N == B * p.rows_per_sample), "decoder: batch=%d does not divide num_rows=%d / N=%d", B, p.num_rows, N);
    const int rows_ps = p.num_rows / B;          // feature cells of one sample
    MV2D_CHECK_ARG(rows_ps % (p.grid_h * p.grid_w) == 0, "decoder: rows per sample %d is not a multiple of the %dx%d grid", rows_ps, p.grid_h, p.grid_w);
    xg.N = N; xg.h = p.grid_h; xg.w = p.grid_w; xg.V = rows_ps / (p.grid_h * p.grid_w);
    xg.B = B; xg.Np = N / B;
    xg.tiles_x = cdiv(xg.w, XT_TS); xg.tiles_y = cdiv(xg.h, XT_TS); xg.tiles_ps = xg.V * xg.tiles_x * xg.tiles_y;
    xg.ntiles = B * xg.tiles_ps;
    MV2D_CHECK_ARG(xg.tiles_ps <= XT_MERGE_MAXT && xg.ntiles <= XT_ORDER_MAXT && N <= 32767,
                   "decoder: xa_form 1 supports <= %d tiles per sample, <= %d per batch and <= 32767 queries", XT_MERGE_MAXT, XT_ORDER_MAXT);
    MV2D_CHECK_ARG(p.keymask && p.xa_workspace, "decoder: xa_form 1 needs the key masks and xa_workspace");
    MV2D_CHECK_ARG(p.mask_words * 32 >= rows_ps, "decoder: keymask has %d words for %d cells", p.mask_words, rows_ps);
    xw = xt_carve(p.xa_workspace, xg.Np, xg.tiles_ps, B);
    MV2D_CHECK_ARG(xw.bytes <= p.xa_workspace_bytes, "decoder: xa_workspace too small (%zu < %zu)", p.xa_workspace_bytes, xw.bytes);
    return 0;
}

static int xt_prepare(const Mv2dDecoderParams& p, const XtGeom& xg, const XtWs& xw, cudaStream_t st) {
    XtPrepArgs a{}; a.g = xg; a.keymask = p.keymask; a.mask_words = p.mask_words;
    a.tile_cnt = xw.tile_cnt; a.tile_q = xw.tile_q; a.tile_mask = xw.tile_mask; a.slot_of = xw.slot_of;
    a.tile_work = xw.tile_work;
    launch_k(xt_prep_kernel, dim3(xg.ntiles), dim3(256), 0, st, a);
    MV2D_CHECK_LAUNCH("xt_prep");
    XtListArgs b{}; b.g = xg; b.slot_of = xw.slot_of; b.tile_work = xw.tile_work; b.qlist = xw.qlist; b.qcnt = xw.qcnt;
    b.order = xw.order;
    launch_k(xt_list_kernel, dim3(xg.N + cdiv(xg.ntiles, 256)), dim3(256), 0, st, b);
    MV2D_CHECK_LAUNCH("xt_list");
    if (p.row_tile_live) {
        const int rows_ps = p.num_rows / xg.B;
        MV2D_CHECK_ARG(xg.B == 1 || rows_ps % 128 == 0, "xa_tile_prepare: row_tile_live of a batch needs V*h*w %% 128 == 0 (got %d)", rows_ps);
        launch_k(xt_rowlive_kernel, dim3(cdiv(p.num_rows, 128)), dim3(128), 0, st, p.keymask, p.mask_words, xg.N, p.num_rows, p.row_tile_live,
                 xg.B > 1 ? rows_ps : 0, xg.Np);
        MV2D_CHECK_LAUNCH("xt_rowlive");
    }
    return 0;
}

int run_xa_tile_prepare(const Mv2dDecoderParams& p, cudaStream_t st) {
    if (p.N == 0) return 0;
    XtGeom xg{};
    XtWs xw{};
    int rc;
    if ((rc = xt_setup(p, xg, xw))) return rc;
    return xt_prepare(p, xg, xw, st);
}

// first float of the xa_roi partial-record scratch inside the decoder workspace (must match the carve of run_decoder)
// persistent xa_roi_kernel: two CTAs per SM (or one per possible unit when that is fewer)
static int xr_grid(int N, int max_match) {
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const long long cap = (long long)N * max_match;
    return (int)(cap < 2LL * num_sms ? cap : 2LL * num_sms);
}

static size_t xr_part_offset(int N, int L) {
    return (size_t)N * (9 * MV2D_C + 768 + MV2D_C + 5 * 2048 + DEC_SPLIT * MV2D_C) + (size_t)4 * L * N * MV2D_C;
}

// The sparse cross-attention core of ONE decoder layer on its own (measurement / stage-level binding; the attention-only
// microbenchmark of BASELINE configs[4]): exactly the launches run_decoder issues between the query projection and the
// output projection.
int run_cross_attention_core(const Mv2dDecoderParams& p, int layer, const float* q, float* ctx, float* ctx_lo, cudaStream_t st) {
    const int N = p.N;
    MV2D_CHECK_ARG(N > 0 && layer >= 0 && layer < p.L && q && ctx, "cross_attention_core: bad N / layer / pointers");
    cudaError_t e;
    if (p.mode == 0) {
        MV2D_CHECK_ARG(p.match && p.match_cnt && p.max_match > 0 && p.max_match <= XR_MAXM && p.workspace,
                       "cross_attention_core: S head needs match lists (max_match <= %d) and the decoder workspace", XR_MAXM);
        MV2D_CHECK_ARG((xr_part_offset(N, p.L) + (size_t)N * XR_MAXM * XR_REC + N + (size_t)N * XR_MAXM + 1) * sizeof(float) <= p.workspace_bytes,
                       "cross_attention_core: workspace too small");
        float* part = p.workspace + xr_part_offset(N, p.L);
        int* ticket = reinterpret_cast<int*>(part + (size_t)N * XR_MAXM * XR_REC);
        int* units = ticket + N;
        int* total = units + (size_t)N * XR_MAXM;
        launch_k(xr_units_kernel, dim3(1), dim3(XR_UNITS_THREADS), 0, st, p.match_cnt, N, units, total);
        MV2D_CHECK_LAUNCH("xr_units");
        if ((e = cudaFuncSetAttribute(xa_roi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XR_SMEM_BYTES)) != cudaSuccess ||
            (e = cudaMemsetAsync(ticket, 0, (size_t)N * sizeof(int), st)) != cudaSuccess) {
            set_error("cross_attention_core: setup %s", cudaGetErrorString(e));
            return (int)e;
        }
        XrArgs a{}; a.qt = q; a.kin_rows = p.kin_rows; a.mem_rows = p.mem_rows; a.match = p.match;
        a.match_cnt = p.match_cnt; a.max_match = p.max_match; a.ctx = ctx; a.ctx_lo = ctx_lo; a.part = part; a.ticket = ticket;
        a.units = units; a.total = total;
        launch_k(xa_roi_kernel, dim3(xr_grid(N, p.max_match)), dim3(XR_THREADS), (size_t)XR_SMEM_BYTES, st, a);
        MV2D_CHECK_LAUNCH("xa_roi");
        return 0;
    }
    MV2D_CHECK_ARG(p.xa_form == 1 && p.kp && p.vp && p.xa_prepared, "cross_attention_core: T head needs xa_form 1, kp / vp and prepared tile lists");
    XtGeom xg{};
    XtWs xw{};
    int rc;
    if ((rc = xt_setup(p, xg, xw))) return rc;
    if ((e = cudaFuncSetAttribute(xt_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XT_SMEM_BYTES)) != cudaSuccess ||
        (e = cudaFuncSetAttribute(xt_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, XT_MERGE_MAXT * 36)) != cudaSuccess) {
        set_error("cross_attention_core: xt smem attr %s", cudaGetErrorString(e));
        return (int)e;
    }
    XtAttnArgs a{}; a.g = xg; a.q = q; a.kp = p.kp + (long long)layer * p.num_rows * MV2D_C; a.vp = p.vp + (long long)layer * p.num_rows * MV2D_C;
    a.tile_cnt = xw.tile_cnt; a.tile_q = xw.tile_q; a.tile_mask = xw.tile_mask; a.rec = xw.rec; a.order = xw.order; a.qsplit = 1;
    a.row_live = p.row_tile_live;
    { int rc; if ((rc = launch_xt_attn(a, xg.ntiles, st))) return rc; }
    XtMergeArgs m{}; m.g = xg; m.qlist = xw.qlist; m.qcnt = xw.qcnt; m.rec = xw.rec; m.ctx = ctx; m.ctx_lo = ctx_lo;
    launch_k(xt_merge_kernel, dim3(N), dim3(XT_MERGE_THREADS), (size_t)xg.tiles_ps * 36, st, m);
    MV2D_CHECK_LAUNCH("xt_merge");
    return 0;
}

int run_decoder(const Mv2dDecoderParams& p, cudaStream_t st) {
    const int N = p.N, L = p.L, C = MV2D_C;
    MV2D_CHECK_ARG(N >= 0 && L >= 1 && L <= MV2D_MAX_LAYERS, "decoder: bad N=%d / L=%d", N, L);
    if (N == 0) return 0;
    MV2D_CHECK_ARG(p.num_rows > 0, "decoder: num_rows=%d must be positive", p.num_rows);
    MV2D_CHECK_ARG(p.layers && p.branches, "decoder: missing weights");
    MV2D_CHECK_ARG(p.mode == 0 ? (p.match && p.match_cnt && p.max_match > 0) : (p.keymask && p.mask_words > 0),
                   "decoder: key description missing for mode %d", p.mode);
    MV2D_CHECK_ARG(p.batch == 0 || (p.batch > 0 && p.rows_per_sample > 0 && N == p.batch * p.rows_per_sample && !p.persistent &&
                                    !p.self_attn_mask && (p.mode == 0 || p.xa_form == 1)),
                   "decoder: batch=%d x rows_per_sample=%d must equal N=%d (staged decoder, no attention mask, T head: xa_form 1)",
                   p.batch, p.rows_per_sample, N);
    float* ws = p.workspace;
    float* x = ws;    ws += (size_t)N * C;
    float* xq = ws;   ws += (size_t)N * C;
    float* x1 = ws;   ws += (size_t)N * C;
    float* x1q = ws;  ws += (size_t)N * C;
    float* x2 = ws;   ws += (size_t)N * C;
    float* x1q_hi = ws; ws += (size_t)N * C;
    float* x1q_lo = ws; ws += (size_t)N * C;
    float* x2_hi = ws;  ws += (size_t)N * C;
    float* x2_lo = ws;  ws += (size_t)N * C;
    float* qkv = ws;  ws += (size_t)N * 768;
    float* sa = ws;   ws += (size_t)N * C;
    float* qt = ws;   ws += (size_t)N * 2048;
    float* ctx = ws;  ws += (size_t)N * 2048;
    float* ctx_lo = ws; ws += (size_t)N * 2048;
    float* hdn = ws;  ws += (size_t)N * 2048;
    float* hdn_lo = ws; ws += (size_t)N * 2048;
    float* part = ws; ws += (size_t)DEC_SPLIT * N * C;
    float* b0 = ws;   ws += (size_t)L * N * C;
    float* b1 = ws;   ws += (size_t)L * N * C;
    float* b2 = ws;   ws += (size_t)L * N * C;
    float* b3 = ws;   ws += (size_t)L * N * C;
    MV2D_CHECK_ARG((size_t)(ws - p.workspace) == xr_part_offset(N, L), "decoder: internal workspace layout drifted");
    float* xr_part = ws; ws += (size_t)N * XR_MAXM * XR_REC;
    int* xr_ticket = reinterpret_cast<int*>(ws); ws += N;
    int* xr_units = reinterpret_cast<int*>(ws); ws += (size_t)N * XR_MAXM;      // (query, slot) work list of xa_roi_kernel
    int* xr_total = reinterpret_cast<int*>(ws); ws += 1;
    ws = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);     // TMA operands follow
    // 3xTF32 operand splits of the GEMMs that run as FFMA below 512 rows
    float* x_hi = ws;  ws += (size_t)N * C;
    float* x_lo = ws;  ws += (size_t)N * C;
    float* xq_hi = ws; ws += (size_t)N * C;
    float* xq_lo = ws; ws += (size_t)N * C;
    float* sa_lo = ws; ws += (size_t)N * C;
    float* xctx_lo = ws; ws += (size_t)N * C;
    float* in_hi = ws; ws += (size_t)L * N * C;      // post-normed states (outs_dec) split
    float* in_lo = ws; ws += (size_t)L * N * C;
    float* b1_lo = ws; ws += (size_t)L * N * C;
    float* b2_lo = ws; ws += (size_t)L * N * C;
    float* b4 = ws;    ws += (size_t)L * N * C;
    ws += (size_t)2 * L * N * C;                     // (two more [L,N,C] slots of the layout mv2d_decoder_workspace_bytes reports; unused)
    unsigned* barrier = reinterpret_cast<unsigned*>(reinterpret_cast<uintptr_t>(ws + 63) & ~(uintptr_t)63); ws += 1024;
    MV2D_CHECK_ARG((size_t)(ws - p.workspace) * sizeof(float) <= p.workspace_bytes, "decoder: workspace too small");
    const long long NC = (long long)N * C;
    cudaError_t e;
    const int lb = p.layer_begin, le = p.layer_end > 0 ? p.layer_end : L;
    MV2D_CHECK_ARG(lb >= 0 && lb < le && le <= L, "decoder: bad layer range [%d,%d) of %d", lb, le, L);
    MV2D_CHECK_ARG(!p.persistent || (lb == 0 && le == L), "decoder: the persistent kernel runs all layers");
    const bool first = lb == 0, last = le == L;
    // two-frame head, key-stationary form over 8x8 tiles of projected keys / values (xa_tile.cuh)
    const bool xt = p.mode == 1 && p.xa_form == 1 && !p.persistent;
    XtGeom xg{};
    XtWs xw{};
    if (xt) {
        int rc0;
        if ((rc0 = xt_setup(p, xg, xw))) return rc0;
        MV2D_CHECK_ARG(p.kp && p.vp, "decoder: xa_form 1 needs kp / vp (mv2d_kv_project)");
        if ((e = cudaFuncSetAttribute(xt_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XT_SMEM_BYTES)) != cudaSuccess ||
            (e = cudaFuncSetAttribute(xt_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, XT_MERGE_MAXT * 36)) != cudaSuccess) {
            set_error("decoder: xt smem attr %s", cudaGetErrorString(e));
            return (int)e;
        }
        if (first && !p.xa_prepared && (rc0 = xt_prepare(p, xg, xw, st))) return rc0;
    }
    // target = 0 ; query + query_pos = query_pos   (cross_attention_head.py:32)
    // The first layer's self-attention sees value = target = 0, so every value row is the bias bv and the attention
    // output is out_proj(bv) + bo for EVERY query whatever the weights/mask: a [256] constant packed once
    // (Mv2dLayerWeights.sa_const).  With it, layer 0 starts at LayerNorm 1 and x / xq are first written by its LN 3.
    const bool fold0 = !p.persistent && p.layers[0].sa_const != nullptr;
    if (first && !fold0 &&
        ((e = cudaMemsetAsync(x, 0, NC * sizeof(float), st)) != cudaSuccess ||
         (e = cudaMemcpyAsync(xq, p.query_pos, NC * sizeof(float), cudaMemcpyDeviceToDevice, st)) != cudaSuccess)) {
        set_error("decoder: init %s", cudaGetErrorString(e));
        return (int)e;
    }
    const bool global_list = p.mode == 1 && p.key_list != nullptr;
    const int klist_cap = p.mode == 0 ? p.max_match * MV2D_TOK : (global_list ? 0 : p.mask_words * 32);
    // S head (~64 keys / query): 4 warps, 16-key chunks, 3 CTAs per SM.  T head (~2000 keys / query): 8 warps and
    // 32-key chunks measured faster than the small variant even when that runs 3 CTAs per SM (fewer barriers per key)
    const int xa_ch = p.mode == 0 ? 16 : 32;
    const size_t xa_smem = (size_t)(4 * xa_ch * MV2D_C + xa_ch * 8 + 24) * sizeof(float) + (size_t)klist_cap * sizeof(uint16_t);
    MV2D_CHECK_ARG(xa_smem <= 227 * 1024, "decoder: key list does not fit shared memory");
    auto xa_kernel = xa_ch == 16 ? cross_attn_kernel<16, 128> : cross_attn_kernel<32, 256>;
    MV2D_CHECK_ARG(p.mode == 0 || p.mask_words <= 4096, "decoder: mask_words=%d > 4096", p.mask_words);
    if ((e = cudaFuncSetAttribute(xa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xa_smem)) != cudaSuccess) {
        set_error("decoder: smem attr %s", cudaGetErrorString(e));
        return (int)e;
    }
    // S head: one CTA per (query, matched RoI) with bulk copies (xa_roi_kernel); MV2D_XA_ROI=0 keeps the per-query kernel
    static const bool xr_on = []() { const char* v = getenv("MV2D_XA_ROI"); return !(v && v[0] == '0'); }();
    const bool use_xr = xr_on && p.mode == 0 && !p.persistent && p.max_match <= XR_MAXM;
    // the per-query kernel (and the persistent decoder) address key rows with 16-bit ids
    MV2D_CHECK_ARG(use_xr || xt || p.num_rows <= 65536, "decoder: num_rows=%d must be <= 65536 (16-bit key ids)", p.num_rows);
    if ((e = cudaFuncSetAttribute(self_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SA_SMEM_BYTES)) != cudaSuccess ||
        (e = cudaFuncSetAttribute(self_attn_blk_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SbCfg<64>::SMEM_BYTES)) != cudaSuccess ||
        (e = cudaFuncSetAttribute(self_attn_blk_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SbCfg<32>::SMEM_BYTES)) != cudaSuccess) {
        set_error("decoder: self_attn smem attr %s", cudaGetErrorString(e));
        return (int)e;
    }
    if (use_xr) {
        if ((e = cudaFuncSetAttribute(xa_roi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XR_SMEM_BYTES)) != cudaSuccess ||
            (first && (e = cudaMemsetAsync(xr_ticket, 0, (size_t)N * sizeof(int), st)) != cudaSuccess)) {
            set_error("decoder: xa_roi setup %s", cudaGetErrorString(e));
            return (int)e;
        }
        if (first) {
            launch_k(xr_units_kernel, dim3(1), dim3(XR_UNITS_THREADS), 0, st, p.match_cnt, N, xr_units, xr_total);
            MV2D_CHECK_LAUNCH("xr_units");
        }
    }
    int rc;
    const Mv2dBranchWeights& B = *p.branches;
    if (p.persistent) {
        // ---- one cooperative launch for all layers + branches (decoder_mega.cuh)
        static int num_sms = 0;
        if (num_sms == 0) {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        }
        MegaParams* mp = new MegaParams();           // ~20 KB: keep it off the stack
        MegaParams& m = *mp;
        m.N = N; m.L = L; m.mode = p.mode; m.max_match = p.max_match; m.mask_words = p.mask_words; m.klist_cap = klist_cap;
        for (int i = 0; i < 6; ++i) m.pc_range[i] = p.pc_range[i];
        m.vel_dt = p.vel_dt; m.vel_row_start = p.vel_row_start;
        m.query_pos = p.query_pos; m.ref = p.ref; m.kin_rows = p.kin_rows; m.mem_rows = p.mem_rows;
        m.match = p.match; m.match_cnt = p.match_cnt; m.keymask = p.keymask; m.self_attn_mask = p.self_attn_mask;
        m.key_list = p.key_list; m.key_cnt = p.key_cnt;
        m.x = x; m.xq = xq; m.x1 = x1; m.x1q = x1q; m.x2 = x2; m.x1q_hi = x1q_hi; m.x1q_lo = x1q_lo; m.x2_hi = x2_hi; m.x2_lo = x2_lo;
        m.qkv = qkv; m.sa = sa; m.qt = qt; m.ctx = ctx; m.ctx_lo = ctx_lo; m.hdn = hdn; m.hdn_lo = hdn_lo; m.part = part;
        m.b0 = b0; m.b1 = b1; m.b2 = b2; m.b3 = b3; m.cls = p.cls_scores; m.box = p.bbox_preds; m.outs_dec = p.outs_dec;
        m.br = B; m.barrier = barrier;
        rc = 0;
        for (int l = 0; l < L && rc == 0; ++l) {
            const Mv2dLayerWeights& w = p.layers[l];
            MegaLayer& y = m.layer[l];
            y.sa_in_w = w.sa_in_w; y.sa_in_b = w.sa_in_b; y.sa_out_w = w.sa_out_w; y.sa_out_b = w.sa_out_b;
            y.ca_q_b = w.ca_q_b; y.ca_o_b = w.ca_o_b; y.ffn_b1 = w.ffn_b1; y.ffn_b2 = w.ffn_b2;
            for (int i = 0; i < 3; ++i) { y.ln_g[i] = w.ln_g[i]; y.ln_b[i] = w.ln_b[i]; }
            struct { const float *ah, *al; int lda; const float *wh, *wl; int ldw, n, k; } gm[4] = {
                {x1q_hi, x1q_lo, C, w.ca_q_w, w.ca_q_w_lo, C, 2048, C},
                {ctx, ctx_lo, 2048, w.ca_o_w, w.ca_o_w_lo, 2048, C, 2048},
                {x2_hi, x2_lo, C, w.ffn_w1, w.ffn_w1_lo, C, 2048, C},
                {hdn, hdn_lo, 2048, w.ffn_w2, w.ffn_w2_lo, 2048, C, 2048}};
            for (int g = 0; g < 4 && rc == 0; ++g) {
                if ((rc = tc_make_map_2d(&y.maps[g][0], gm[g].ah, N, gm[g].k, gm[g].lda, 128))) break;
                if ((rc = tc_make_map_2d(&y.maps[g][1], gm[g].al, N, gm[g].k, gm[g].lda, 128))) break;
                if ((rc = tc_make_map_2d(&y.maps[g][2], gm[g].wh, gm[g].n, gm[g].k, gm[g].ldw, mega_tc::BN))) break;
                if ((rc = tc_make_map_2d(&y.maps[g][3], gm[g].wl, gm[g].n, gm[g].k, gm[g].ldw, mega_tc::BN))) break;
            }
        }
        if (rc) { delete mp; return rc; }
        size_t smem = (size_t)mega_tc::SMEM_BYTES;
        if (xa_smem_mega(p.mode, klist_cap) > smem) smem = xa_smem_mega(p.mode, klist_cap);
        if ((size_t)SA_SMEM_BYTES > smem) smem = SA_SMEM_BYTES;
        smem += 1024;
        if ((e = cudaFuncSetAttribute(decoder_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess ||
            (e = cudaMemsetAsync(barrier, 0, sizeof(unsigned), st)) != cudaSuccess) {
            set_error("decoder(persistent): setup %s", cudaGetErrorString(e));
            delete mp;
            return (int)e;
        }
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(num_sms); cfg.blockDim = dim3(MEGA_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, decoder_mega_kernel, m);
        delete mp;
        note_launch();
        if (e != cudaSuccess) { set_error("decoder(persistent): launch %s", cudaGetErrorString(e)); return (int)e; }
        return 0;
    }
    // More than 512 query rows (batches): every GEMM of the layer is a GPU-filling problem, so the ones that run as
    // FFMA kernels at M ~ 300 (in_proj, out_proj, the T head's q / out projections, the branch MLPs) go to the 3xTF32
    // tcgen05 kernel as well; their A operands are split by the producing kernel's epilogue.
    static const bool tc_all_env = []() { const char* v = getenv("MV2D_DEC_TC_ALL"); return !(v && v[0] == '0'); }();
    const bool big = tc_all_env && N > 512 && p.layers[0].sa_in_w_hi && p.layers[0].sa_out_w_hi && B.cls_w0_hi &&
                     (!xt || (p.layers[0].xa_q_w_hi && p.layers[0].xa_o_w_hi));
    if (big && first && !fold0) {      // layer 0 runs its self-attention: x = 0 and x + pos = pos need their splits too
        if ((e = cudaMemsetAsync(x_hi, 0, 2 * NC * sizeof(float), st)) != cudaSuccess) { set_error("decoder: init %s", cudaGetErrorString(e)); return (int)e; }
        if ((rc = launch_split_tf32(p.query_pos, xq_hi, xq_lo, NC, st))) return rc;
    }
    // projection + residual + LayerNorm as one cluster launch (gemm_ln.cu) wherever the operands exist as TF32 hi / lo pairs
    const bool fuse_ln = gemm_ln_enabled();
    const bool fuse_sa = fuse_ln && p.layers[0].sa_out_w_hi && p.layers[0].sa_out_w_lo;
    const bool fuse_xo = fuse_ln && xt && p.layers[0].xa_o_w_hi && p.layers[0].xa_o_w_lo;
    const bool sa_split = big || fuse_sa;
    for (int l = lb; l < le; ++l) {
        const Mv2dLayerWeights& w = p.layers[l];
        float* inter = p.outs_dec + (long long)l * NC;
        if (l == 0 && fold0) {
            LnArgs a{}; a.partial = w.sa_const; a.bcast_in = 1; a.nsplit = 1;
            a.gamma = w.ln_g[0]; a.beta = w.ln_b[0]; a.qpos = p.query_pos; a.out = x1; a.out_q = x1q; a.rows = N;
            a.outq_hi = x1q_hi; a.outq_lo = x1q_lo;
            if ((rc = ln(a, st))) return rc;
        } else {
            // --- self attention: q,k from (x + qpos), v from x
            if (big) {
                TcGemm t{};
                t.A = xq_hi; t.A_lo = xq_lo; t.A2 = x_hi; t.A2_lo = x_lo; t.n_switch = 512; t.lda = C;
                t.W = w.sa_in_w_hi; t.W_lo = w.sa_in_w_lo; t.ldw = C; t.bias = w.sa_in_b; t.C = qkv; t.ldc = 768;
                t.M = N; t.N = 768; t.K = C; t.passes = 3; t.nsplit = 1;
                if ((rc = launch_gemm_tc(t, st))) return rc;
            } else {
                GemmArgs g{};
                g.A = xq; g.lda = C; g.W = w.sa_in_w; g.ldw = C; g.C = qkv; g.ldc = 768; g.bias = w.sa_in_b;
                g.M = N; g.N = 768; g.K = C; g.batch = 1; g.nsplit = 1;
                if ((rc = launch_gemm_small(g, x, 512, st))) return rc;
            }
            static const bool sa_blk = []() { const char* v = getenv("MV2D_SA_BLOCKED"); return !(v && v[0] == '0'); }();
            // MV2D_SA_MMA=0: FFMA kernels only.  Default: the tensor-core kernel whenever the sample's K / V slices of one head fit
            // shared memory (measured per layer: 31 vs 47 us at 8 x 300 queries, 16 vs 22 us at 2 x 300, 11 vs 15 us at 1 x 300)
            static const int sa_mma = []() { const char* v = getenv("MV2D_SA_MMA"); return v ? atoi(v) : 1; }();
            const int sa_rows = p.batch > 0 ? p.rows_per_sample : N, sa_nb = p.batch > 0 ? p.batch : 1;
            if (sa_mma > 0 && sam_smem_bytes(sa_rows) <= 200 * 1024) {
                static int num_sms = 0;
                if (num_sms == 0) {
                    int dev = 0;
                    cudaGetDevice(&dev);
                    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
                }
                // W warps (16-query row blocks) per CTA: the smallest W >= 2 that makes the grid one wave at 2 CTAs per SM
                const int nrb = cdiv(sa_rows, 16), slots = 2 * num_sms;
                int W = 8;
                for (int w = 2; w <= 8; ++w)
                    if (sa_nb * MV2D_HEADS * cdiv(nrb, w) <= slots) { W = w; break; }
                const size_t smem = sam_smem_bytes(sa_rows, W);
                // two warps per row block (alternate key blocks, folded at the end) while the CTA stays within 320 threads
                static const bool ks2_env = []() { const char* v = getenv("MV2D_SA_KS2"); return !(v && v[0] == '0'); }();
                const bool ks2 = ks2_env && W <= 5;
                auto kern = ks2 ? self_attn_mma_kernel<2> : self_attn_mma_kernel<1>;
                if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) {
                    set_error("decoder: self_attn_mma smem attr %s", cudaGetErrorString(e));
                    return (int)e;
                }
                launch_k(kern, dim3(cdiv(nrb, W), MV2D_HEADS, sa_nb), dim3(32 * W * (ks2 ? 2 : 1)), smem, st, (const float*)qkv,
                         p.batch > 0 ? (const uint8_t*)nullptr : p.self_attn_mask, N, sa, p.batch > 0 ? p.rows_per_sample : 0, p.n_real,
                         sa_split ? sa_lo : (float*)nullptr);
            } else if (sa_blk) {
                const int rows = p.batch > 0 ? p.rows_per_sample : N, nb = p.batch > 0 ? p.batch : 1;
                const int rps = p.batch > 0 ? p.rows_per_sample : 0;
                const uint8_t* am = p.batch > 0 ? nullptr : p.self_attn_mask;
                // 64 queries per CTA halve the K / V staging per query, 32 give twice the CTAs: take the variant with the
                // smaller (waves x work per CTA); two 64-query CTAs fit an SM (93 KB of shared memory, 127 registers), three
                // 32-query ones (69 KB, 79 registers)
                const int c64 = cdiv(rows, 64) * MV2D_HEADS * nb, c32 = cdiv(rows, 32) * MV2D_HEADS * nb;
                if (cdiv(c64, 296) * 64 <= cdiv(c32, 444) * 32)
                    launch_k(self_attn_blk_kernel<64>, dim3(cdiv(rows, 64), MV2D_HEADS, nb), dim3(256), (size_t)SbCfg<64>::SMEM_BYTES, st,
                             (const float*)qkv, am, N, sa, rps, p.n_real, sa_split ? sa_lo : (float*)nullptr);
                else
                    launch_k(self_attn_blk_kernel<32>, dim3(cdiv(rows, 32), MV2D_HEADS, nb), dim3(256), (size_t)SbCfg<32>::SMEM_BYTES, st,
                             (const float*)qkv, am, N, sa, rps, p.n_real, sa_split ? sa_lo : (float*)nullptr);
            } else if (p.batch > 0)
                launch_k(self_attn_kernel, dim3(cdiv(p.rows_per_sample, 8 * SA_QPW), MV2D_HEADS, p.batch), dim3(256), SA_SMEM_BYTES, st,
                         (const float*)qkv, (const uint8_t*)nullptr, N, sa, p.rows_per_sample, p.n_real, sa_split ? sa_lo : (float*)nullptr);
            else
                launch_k(self_attn_kernel, dim3(cdiv(N, 8 * SA_QPW), MV2D_HEADS), dim3(256), SA_SMEM_BYTES, st, (const float*)qkv, p.self_attn_mask, N, sa,
                         0, (const int*)nullptr, sa_split ? sa_lo : (float*)nullptr);
            MV2D_CHECK_LAUNCH("self_attn");
            {
                LnArgs a{}; a.partial = part; a.nsplit = 1; a.bias = w.sa_out_b; a.residual = x;
                a.gamma = w.ln_g[0]; a.beta = w.ln_b[0]; a.qpos = p.query_pos; a.out = x1; a.out_q = x1q; a.rows = N;
                a.outq_hi = x1q_hi; a.outq_lo = x1q_lo;
                if (fuse_sa) {          // out_proj + residual + LayerNorm in one cluster launch (gemm_ln.cu)
                    if ((rc = launch_gemm_ln(sa, sa_lo, C, w.sa_out_w_hi, w.sa_out_w_lo, C, N, C, a, 0, st))) return rc;
                } else {
                    if (big) {
                        if ((rc = tc3(sa, sa_lo, C, w.sa_out_w_hi, w.sa_out_w_lo, C, nullptr, part, nullptr, C, N, C, C, 0, 1, 0, st))) return rc;
                    } else {
                        if ((rc = gemm(sa, C, w.sa_out_w, C, nullptr, part, C, N, C, C, 0, st))) return rc;
                    }
                    if ((rc = ln(a, st))) return rc;
                }
            }
        }
        if (xt) {
            // --- sparse cross attention, key-stationary: q projection, per-tile partial attention, merge, output projection
            MV2D_CHECK_ARG(w.xa_q_w && w.xa_q_b && w.xa_o_w && w.xa_o_b, "decoder: layer %d has no xa_q / xa_o weights", l);
            if (big) {
                if ((rc = tc3(x1q_hi, x1q_lo, C, w.xa_q_w_hi, w.xa_q_w_lo, C, w.xa_q_b, xw.qp, nullptr, C, N, C, C, 0, 1, 0, st))) return rc;
            } else {
                if ((rc = gemm(x1q, C, w.xa_q_w, C, w.xa_q_b, xw.qp, C, N, C, C, 0, st))) return rc;
            }
            {
                XtAttnArgs a{}; a.g = xg; a.q = xw.qp; a.kp = p.kp + (long long)l * p.num_rows * C; a.vp = p.vp + (long long)l * p.num_rows * C;
                a.tile_cnt = xw.tile_cnt; a.tile_q = xw.tile_q; a.tile_mask = xw.tile_mask; a.rec = xw.rec; a.order = xw.order;
                static const int qs_env = []() { const char* v = getenv("MV2D_XT_QSPLIT"); return v ? atoi(v) : 0; }();
                a.qsplit = qs_env > 0 ? qs_env : 1;
                a.row_live = p.row_tile_live;
                if ((rc = launch_xt_attn(a, xg.ntiles, st))) return rc;
            }
            {
                XtMergeArgs a{}; a.g = xg; a.qlist = xw.qlist; a.qcnt = xw.qcnt; a.rec = xw.rec; a.ctx = xw.ctx;
                a.ctx_lo = (big || fuse_xo) ? xctx_lo : nullptr;
                launch_k(xt_merge_kernel, dim3(N), dim3(XT_MERGE_THREADS), (size_t)xg.tiles_ps * 36, st, a);
                MV2D_CHECK_LAUNCH("xt_merge");
            }
            LnArgs a{}; a.partial = part; a.nsplit = 1; a.bias = w.xa_o_b; a.residual = x1;
            a.gamma = w.ln_g[1]; a.beta = w.ln_b[1]; a.out = x2; a.rows = N; a.out_hi = x2_hi; a.out_lo = x2_lo;
            if (fuse_xo) {
                if ((rc = launch_gemm_ln(xw.ctx, xctx_lo, C, w.xa_o_w_hi, w.xa_o_w_lo, C, N, C, a, 0, st))) return rc;
            } else {
                if (big) {
                    if ((rc = tc3(xw.ctx, xctx_lo, C, w.xa_o_w_hi, w.xa_o_w_lo, C, nullptr, part, nullptr, C, N, C, C, 0, 1, 0, st))) return rc;
                } else {
                    if ((rc = gemm(xw.ctx, C, w.xa_o_w, C, nullptr, part, C, N, C, C, 0, st))) return rc;
                }
                if ((rc = ln(a, st))) return rc;
            }
        } else {
        // --- sparse cross attention (absorbed)
        if ((rc = tc3(x1q_hi, x1q_lo, C, w.ca_q_w, w.ca_q_w_lo, C, w.ca_q_b, qt, nullptr, 2048, N, 2048, C, 0, 1, 0, st))) return rc;
        if (use_xr) {
            XrArgs a{}; a.qt = qt; a.kin_rows = p.kin_rows; a.mem_rows = p.mem_rows; a.match = p.match;
            a.match_cnt = p.match_cnt; a.max_match = p.max_match; a.ctx = ctx; a.ctx_lo = ctx_lo;
            a.part = xr_part; a.ticket = xr_ticket; a.units = xr_units; a.total = xr_total;
            launch_k(xa_roi_kernel, dim3(xr_grid(N, p.max_match)), dim3(XR_THREADS), (size_t)XR_SMEM_BYTES, st, a);
            MV2D_CHECK_LAUNCH("xa_roi");
        } else {
            XaArgs a{}; a.qt = qt; a.kin_rows = p.kin_rows; a.mem_rows = p.mem_rows; a.match = p.match;
            a.match_cnt = p.match_cnt; a.max_match = p.max_match; a.keymask = p.keymask; a.mask_words = p.mask_words;
            a.key_list = p.key_list; a.key_cnt = p.key_cnt;
            a.mode = p.mode; a.N = N; a.klist_cap = klist_cap; a.ctx = ctx; a.ctx_lo = ctx_lo;
            launch_k(xa_kernel, dim3(N), dim3(xa_ch == 16 ? 128 : 256), xa_smem, st, a);
            MV2D_CHECK_LAUNCH("cross_attn");
        }
        // split-K only while the M tiles alone cannot fill the GPU
        const int ksplit = big ? (N > 2048 ? 2 : 4) : DEC_SPLIT;
        {
            LnArgs a{}; a.partial = part; a.nsplit = ksplit; a.split_stride = NC; a.bias = w.ca_o_b; a.residual = x1;
            a.gamma = w.ln_g[1]; a.beta = w.ln_b[1]; a.out = x2; a.rows = N; a.out_hi = x2_hi; a.out_lo = x2_lo;
            if (fuse_ln) {
                if ((rc = launch_gemm_ln(ctx, ctx_lo, 2048, w.ca_o_w, w.ca_o_w_lo, 2048, N, 2048, a, 0, st))) return rc;
            } else {
                if ((rc = tc3(ctx, ctx_lo, 2048, w.ca_o_w, w.ca_o_w_lo, 2048, nullptr, part, nullptr, C, N, C, 2048, 0, ksplit, NC, st))) return rc;
                if ((rc = ln(a, st))) return rc;
            }
        }
        }
        // --- FFN
        const int ksplit2 = big ? (N > 2048 ? 2 : 4) : DEC_SPLIT;
        if ((rc = tc3(x2_hi, x2_lo, C, w.ffn_w1, w.ffn_w1_lo, C, w.ffn_b1, hdn, hdn_lo, 2048, N, 2048, C,
                      GEMM_RELU | GEMM_SPLIT_OUT, 1, 0, st))) return rc;
        if (!fuse_ln && (rc = tc3(hdn, hdn_lo, 2048, w.ffn_w2, w.ffn_w2_lo, 2048, nullptr, part, nullptr, C, N, C, 2048, 0, ksplit2, NC, st))) return rc;
        {
            LnArgs a{}; a.partial = part; a.nsplit = ksplit2; a.split_stride = NC; a.bias = w.ffn_b2; a.residual = x2;
            a.gamma = w.ln_g[2]; a.beta = w.ln_b[2]; a.qpos = p.query_pos; a.out = x; a.out_q = xq;
            a.gamma2 = B.post_g; a.beta2 = B.post_b; a.out2 = inter; a.rows = N;
            if (big) {
                a.out_hi = x_hi; a.out_lo = x_lo; a.outq_hi = xq_hi; a.outq_lo = xq_lo;
                a.out2_hi = in_hi + (long long)l * NC; a.out2_lo = in_lo + (long long)l * NC;
            }
            if (fuse_ln) {
                if ((rc = launch_gemm_ln(hdn, hdn_lo, 2048, w.ffn_w2, w.ffn_w2_lo, 2048, N, 2048, a, 0, st))) return rc;
            } else if ((rc = ln(a, st))) return rc;
        }
    }
    if (!last) return 0;
    // --- branches, batched over layers (cross_attention_head.py:216-231)
    const long long CC = (long long)C * C;
    if (big) {
        // one grouped 3xTF32 launch per Linear: group = decoder layer (its own [256,256] matrix), M = N rows each
        auto grouped = [&](const float* a_hi, const float* a_lo, const float* w_hi, const float* w_lo, const float* bias, float* out,
                           float* out_lo, int flags) {
            TcGemm t{};
            t.A = a_hi; t.A_lo = a_lo; t.lda = C; t.W = w_hi; t.W_lo = w_lo; t.ldw = C; t.bias = bias; t.C = out; t.C_lo = out_lo; t.ldc = C;
            t.M = N; t.N = C; t.K = C; t.passes = 3; t.nsplit = 1; t.flags = flags; t.groups = L; t.group_rows = N;
            return launch_gemm_tc(t, st);
        };
        if ((rc = grouped(in_hi, in_lo, B.cls_w0_hi, B.cls_w0_lo, nullptr, b0, nullptr, 0))) return rc;
        {
            LnArgs a{}; a.partial = b0; a.nsplit = 1; a.bias = B.cls_b0; a.gamma = B.cls_g0; a.beta = B.cls_be0;
            a.rows_per_group = N; a.group_stride = C; a.relu = 1; a.out = b1; a.out_hi = b4; a.out_lo = b1_lo; a.rows = L * N;
            if ((rc = ln(a, st))) return rc;
        }
        if ((rc = grouped(b4, b1_lo, B.cls_w1_hi, B.cls_w1_lo, nullptr, b0, nullptr, 0))) return rc;
        {
            LnArgs a{}; a.partial = b0; a.nsplit = 1; a.bias = B.cls_b1; a.gamma = B.cls_g1; a.beta = B.cls_be1;
            a.rows_per_group = N; a.group_stride = C; a.relu = 1; a.out = b1; a.rows = L * N;
            if ((rc = ln(a, st))) return rc;
        }
        // reg branch: Linear-ReLU twice; the first epilogue emits the hi / lo split the second one reads (b2 = hi)
        if ((rc = grouped(in_hi, in_lo, B.reg_w0_hi, B.reg_w0_lo, B.reg_b0, b2, b2_lo, GEMM_RELU | GEMM_SPLIT_OUT))) return rc;
        if ((rc = grouped(b2, b2_lo, B.reg_w1_hi, B.reg_w1_lo, B.reg_b1, b3, nullptr, GEMM_RELU))) return rc;
    } else {
    if ((rc = gemm(p.outs_dec, C, B.cls_w0, C, nullptr, b0, C, N, C, C, 0, st, 1, 0, L, NC, CC, NC, 0))) return rc;
    {
        LnArgs a{}; a.partial = b0; a.nsplit = 1; a.bias = B.cls_b0; a.gamma = B.cls_g0; a.beta = B.cls_be0;
        a.rows_per_group = N; a.group_stride = C; a.relu = 1; a.out = b1; a.rows = L * N;
        if ((rc = ln(a, st))) return rc;
    }
    if ((rc = gemm(b1, C, B.cls_w1, C, nullptr, b0, C, N, C, C, 0, st, 1, 0, L, NC, CC, NC, 0))) return rc;
    {
        LnArgs a{}; a.partial = b0; a.nsplit = 1; a.bias = B.cls_b1; a.gamma = B.cls_g1; a.beta = B.cls_be1;
        a.rows_per_group = N; a.group_stride = C; a.relu = 1; a.out = b1; a.rows = L * N;
        if ((rc = ln(a, st))) return rc;
    }
    if ((rc = gemm(p.outs_dec, C, B.reg_w0, C, B.reg_b0, b2, C, N, C, C, GEMM_RELU, st, 1, 0, L, NC, CC, NC, C))) return rc;
    if ((rc = gemm(b2, C, B.reg_w1, C, B.reg_b1, b3, C, N, C, C, GEMM_RELU, st, 1, 0, L, NC, CC, NC, C))) return rc;
    }
    launch_k(head10_kernel, dim3(cdiv(L * N, 8)), dim3(256), 0, st, (const float*)b1, (const float*)b3, B.cls_w2, B.cls_b2, B.reg_w2, B.reg_b2, p.ref, L, N,
                                                 p.pc_range[0], p.pc_range[1], p.pc_range[2], p.pc_range[3],
                                                 p.pc_range[4], p.pc_range[5], p.vel_dt, p.vel_row_start, p.cls_scores, p.bbox_preds,
                                                 p.batch > 0 ? p.vel_dt_batch : (const float*)nullptr, p.rows_per_sample);
    MV2D_CHECK_LAUNCH("head10");
    return 0;
}

}  // namespace mv2d
