// Tensor-core form of the flattened self-attention for batches (petr_transformer.py:314-370 -> nn.MultiheadAttention over
// the N queries of ONE sample as one sequence; 8 heads of 32 channels).
// The FFMA kernel (self_attn_blk_kernel) is bound by its shared-memory reads (6 LDS wavefronts per 20 FMAs).  Here a warp
// takes 16 queries of one (sample, head) as the M rows of warp-level TF32 MMAs (mma.sync.m16n8k8) and walks the sample's
// keys in blocks of 64 with an online softmax:
//     S [16 x 64] = Q_h [16 x 32] . K_h^T        8 key blocks x 4 channel steps
//     O [16 x 32] += P [16 x 64] . V_h           8 key steps x 4 channel blocks
// every product error-compensated 3xTF32 (hi*hi + lo*hi + hi*lo, splits made in registers), so the result is fp32-grade like
// the FFMA kernel's (SURVEY App. E rules out single-pass TF32 logits).  The C fragment of S is reused as the A fragment of
// P.V by reading V's rows in the matching order (see xt_attn_mma_kernel, xa_tile.cuh).
// A CTA = W warps = W consecutive 16-query row blocks of one (sample, head); the head's K / V slices of ALL the sample's keys
// are staged once per CTA (cp.async, 16-byte copies) at a pitch of 36 floats, which makes both B-fragment access patterns
// bank-conflict free.  grid = (C, heads, samples) with C * W >= row blocks; the host picks W so that the whole grid is one
// wave at 2 CTAs per SM (B = 8 x 300 queries: W = 5, 256 CTAs).
#pragma once
#include "common.cuh"
#include "xa_tile.cuh"

namespace mv2d {

#define SAM_PITCH 36
#define SAM_KB 64               // keys per block

static inline size_t sam_smem_bytes(int nk_max, int W = 8) {      // K and V slices + the fold scratch of the two-phase form
    return (size_t)2 * ((nk_max + SAM_KB - 1) / SAM_KB * SAM_KB) * SAM_PITCH * 4 + (size_t)W * 32 * 20 * 4;
}

template <int KS>
__global__ void __launch_bounds__(KS == 2 ? 320 : 256, 2)
self_attn_mma_kernel(const float* __restrict__ qkv, const uint8_t* __restrict__ mask, int N, float* __restrict__ out,
                     int rows_per_sample, const int* __restrict__ n_real, float* __restrict__ out_lo) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ __align__(16) float sam_smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
    const int hd = blockIdx.y, W = (blockDim.x >> 5) / KS;      // KS = 2: two warps share a row block, each takes every other key block
    int seg0 = 0, nq = N, nk = N;
    if (rows_per_sample > 0) {
        const int b = blockIdx.z;
        seg0 = b * rows_per_sample; nq = rows_per_sample;
        nk = n_real ? min(n_real[b], rows_per_sample) : rows_per_sample;
        mask = nullptr;
    }
    const int nkb = (nk + SAM_KB - 1) / SAM_KB, nk_pad = nkb * SAM_KB;
    float* Ks = sam_smem;
    float* Vs = Ks + (size_t)nk_pad * SAM_PITCH;
    const float* base = qkv + (long long)seg0 * 768 + hd * 32;
    // ---- stage the head's K and V slices of every key of the sample (rows past nk are zero: p = 0 times stale data)
    for (int i = tid; i < nk_pad * 8; i += blockDim.x) {
        const int key = i >> 3, c4 = i & 7;
        float* kd = Ks + key * SAM_PITCH + c4 * 4;
        float* vd = Vs + key * SAM_PITCH + c4 * 4;
        if (key < nk) {
            const float* src = base + (long long)key * 768 + c4 * 4;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(kd)), "l"(src + 256) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(vd)), "l"(src + 512) : "memory");
        } else {
            *reinterpret_cast<float4*>(kd) = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(vd) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    const int nrb = (nq + 15) >> 4;
    const int rb = blockIdx.x * W + warp % W, ks = warp / W;      // this warp's row block and key-block phase
    // ---- the block's 16 queries: rows g and g + 8 of this lane, scaled by 1 / sqrt(32)
    const int i0 = rb * 16 + g, i1 = i0 + 8;
    const bool act = rb < nrb, ok0 = act && i0 < nq, ok1 = act && i1 < nq;
    uint32_t qh[4][4], ql[4][4];
    {
        const float* q0 = base + (long long)(ok0 ? i0 : 0) * 768 + tg;
        const float* q1 = base + (long long)(ok1 ? i1 : 0) * 768 + tg;
        constexpr float sc = 0.17677669529663687f;
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
            split_tf32_reg(__ldcg(q0 + kc * 8) * sc, qh[kc][0], ql[kc][0]);
            split_tf32_reg(__ldcg(q1 + kc * 8) * sc, qh[kc][1], ql[kc][1]);
            split_tf32_reg(__ldcg(q0 + kc * 8 + 4) * sc, qh[kc][2], ql[kc][2]);
            split_tf32_reg(__ldcg(q1 + kc * 8 + 4) * sc, qh[kc][3], ql[kc][3]);
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float o[4][4];
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) o[nb][0] = o[nb][1] = o[nb][2] = o[nb][3] = 0.f;
    for (int kb = act ? ks : nkb; kb < nkb; kb += KS) {
        const float* Kb = Ks + (size_t)kb * SAM_KB * SAM_PITCH;
        const float* Vb = Vs + (size_t)kb * SAM_KB * SAM_PITCH;
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
                    const float* krow = Kb + ((jh * 4 + jj) * 8 + g) * SAM_PITCH + tg;
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
        // ---- keys past the sample's count / masked pairs; block maxima of rows g (values 0, 1) and g + 8 (values 2, 3)
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k0 = kb * SAM_KB + j * 8 + 2 * tg;
            bool f00 = k0 >= nk, f01 = k0 + 1 >= nk, f10 = f00, f11 = f01;
            if (mask) {
                if (!f00 && ok0) f00 = mask[(long long)i0 * N + k0] != 0;
                if (!f01 && ok0) f01 = mask[(long long)i0 * N + k0 + 1] != 0;
                if (!f10 && ok1) f10 = mask[(long long)i1 * N + k0] != 0;
                if (!f11 && ok1) f11 = mask[(long long)i1 * N + k0 + 1] != 0;
            }
            if (f00) s[j][0] = -INFINITY;
            if (f01) s[j][1] = -INFINITY;
            if (f10) s[j][2] = -INFINITY;
            if (f11) s[j][3] = -INFINITY;
            mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
            mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float n0 = fmaxf(m0, mx0), n1 = fmaxf(m1, mx1);
        const float al0 = (m0 == -INFINITY) ? 0.f : __expf(m0 - n0), al1 = (m1 == -INFINITY) ? 0.f : __expf(m1 - n1);
        const float sub0 = (n0 == -INFINITY) ? 0.f : n0, sub1 = (n1 == -INFINITY) ? 0.f : n1;     // all masked so far: exp(-inf - 0) = 0
        float b0 = 0.f, b1 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j][0] = __expf(s[j][0] - sub0); s[j][1] = __expf(s[j][1] - sub0);
            s[j][2] = __expf(s[j][2] - sub1); s[j][3] = __expf(s[j][3] - sub1);
            b0 += s[j][0] + s[j][1];
            b1 += s[j][2] + s[j][3];
        }
        b0 += __shfl_xor_sync(0xffffffffu, b0, 1); b0 += __shfl_xor_sync(0xffffffffu, b0, 2);
        b1 += __shfl_xor_sync(0xffffffffu, b1, 1); b1 += __shfl_xor_sync(0xffffffffu, b1, 2);
        l0 = l0 * al0 + b0; l1 = l1 * al1 + b1;
        m0 = n0; m1 = n1;
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) { o[nb][0] *= al0; o[nb][1] *= al0; o[nb][2] *= al1; o[nb][3] *= al1; }
        // ---- O += P V : MMA k index tg <-> key 8 j + 2 tg, tg + 4 <-> key 8 j + 2 tg + 1
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint32_t ph[4], pl[4];
            split_tf32_reg(s[j][0], ph[0], pl[0]);      // (row g,     k = tg)
            split_tf32_reg(s[j][2], ph[1], pl[1]);      // (row g + 8, k = tg)
            split_tf32_reg(s[j][1], ph[2], pl[2]);      // (row g,     k = tg + 4)
            split_tf32_reg(s[j][3], ph[3], pl[3]);      // (row g + 8, k = tg + 4)
            const float* vrow = Vb + (j * 8 + 2 * tg) * SAM_PITCH + g;
            uint32_t bh0[4], bl0[4], bh1[4], bl1[4];
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                split_tf32_reg(vrow[nb * 8], bh0[nb], bl0[nb]);
                split_tf32_reg(vrow[SAM_PITCH + nb * 8], bh1[nb], bl1[nb]);
            }
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) mma_tf32_16x8x8(o[nb], ph, bh0[nb], bh1[nb]);
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) mma_tf32_16x8x8(o[nb], pl, bh0[nb], bh1[nb]);
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) mma_tf32_16x8x8(o[nb], ph, bl0[nb], bl1[nb]);
        }
    }
    if (KS == 2) {
        // ---- fold the two key phases of a row block: the odd phase parks (m, l, o) behind the K / V slices
        float* part = Vs + (size_t)nk_pad * SAM_PITCH + ((warp % W) * 32 + lane) * 20;
        if (ks == 1) {
            part[0] = m0; part[1] = m1; part[2] = l0; part[3] = l1;
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) *reinterpret_cast<float4*>(part + 4 + nb * 4) = make_float4(o[nb][0], o[nb][1], o[nb][2], o[nb][3]);
        }
        __syncthreads();
        if (ks == 1) return;
        const float pm0 = part[0], pm1 = part[1];
        const float M0 = fmaxf(m0, pm0), M1 = fmaxf(m1, pm1);
        const float a0 = (m0 == -INFINITY) ? 0.f : __expf(m0 - M0), a1 = (m1 == -INFINITY) ? 0.f : __expf(m1 - M1);
        const float c0 = (pm0 == -INFINITY) ? 0.f : __expf(pm0 - M0), c1 = (pm1 == -INFINITY) ? 0.f : __expf(pm1 - M1);
        l0 = l0 * a0 + part[2] * c0; l1 = l1 * a1 + part[3] * c1;
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
            const float4 q = *reinterpret_cast<const float4*>(part + 4 + nb * 4);
            o[nb][0] = o[nb][0] * a0 + q.x * c0; o[nb][1] = o[nb][1] * a0 + q.y * c0;
            o[nb][2] = o[nb][2] * a1 + q.z * c1; o[nb][3] = o[nb][3] * a1 + q.w * c1;
        }
    }
    if (!act) return;
    // ---- normalise, store (and the TF32 hi / lo split for a 3xTF32 out_proj)
    const float inv0 = l0 > 0.f ? 1.f / l0 : 0.f, inv1 = l1 > 0.f ? 1.f / l1 : 0.f;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        if (!(half ? ok1 : ok0)) continue;
        const float inv = half ? inv1 : inv0;
        const long long off = (long long)(seg0 + (half ? i1 : i0)) * MV2D_C + hd * 32 + 2 * tg;
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
            float2 r = make_float2(o[nb][half * 2] * inv, o[nb][half * 2 + 1] * inv);
            if (out_lo) {
                const float2 hi = make_float2(round_tf32(r.x), round_tf32(r.y));
                *reinterpret_cast<float2*>(out_lo + off + nb * 8) = make_float2(round_tf32(r.x - hi.x), round_tf32(r.y - hi.y));
                r = hi;
            }
            *reinterpret_cast<float2*>(out + off + nb * 8) = r;
        }
    }
}

}  // namespace mv2d
