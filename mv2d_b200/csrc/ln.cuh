// Row-wise LayerNorm epilogue shared by ln_kernel (decoder.cu), the persistent decoder and the fused GEMM + LN kernel
// (gemm_ln.cu).  sm_100a only.
#pragma once
#include "common.cuh"
#include "gemm_tc.cuh"

namespace mv2d {

// ------------------------------------------------------------------------------------------
// Row-wise: x = LN( sum_s partial[s] + bias + residual ) ; optional second LN (post_norm) and
// optional "+ query_pos" copy.  One warp per row of 256.  Also used (relu=1, grouped gammas)
// for the Linear-LN-ReLU blocks of the classification branch.
struct LnArgs {
    const float* partial; int nsplit; long long split_stride;
    const float* bias; const float* residual;
    const float* gamma; const float* beta; int rows_per_group; int group_stride;  // per-layer params
    int relu;
    const float* qpos;   // nullable
    const float* gamma2; const float* beta2;  // nullable: post_norm
    float* out; float* out_q; float* out2;
    float* out_hi; float* out_lo;      // nullable: TF32 split of `out`   (A operand of a 3xTF32 GEMM)
    float* outq_hi; float* outq_lo;    // nullable: TF32 split of `out_q`
    float* out2_hi; float* out2_lo;    // nullable: TF32 split of `out2`
    int rows;
    int bcast_in;     // 1: `partial` is ONE [256] row shared by every output row
};

__device__ __forceinline__ void ln_tail(const LnArgs& a, int row, int grp, int lane, float (&v)[8]);

// gemm_ln.cu: LayerNorm( A . W^T + bias + residual ) in one cluster launch (split-K over the cluster, partial tiles summed
// over distributed shared memory); MV2D_GEMM_LN=0 keeps the split-K GEMM + ln_kernel pair
bool gemm_ln_enabled();
int launch_gemm_ln(const float* A_hi, const float* A_lo, int lda, const float* W_hi, const float* W_lo, int ldw, int M, int K,
                   const LnArgs& ln, int cluster, cudaStream_t st);

__device__ __forceinline__ void ln_body(const LnArgs& a, int vb) {
    const int row = vb * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= a.rows) return;
    const int grp = a.rows_per_group > 0 ? row / a.rows_per_group : 0;
    const long long o = (long long)row * MV2D_C;
    const long long oi = a.bcast_in ? 0 : o;
    float v[8];
    // all loads of the row (up to 8 split-K partials x 2 halves) are issued before the first add: one memory
    // latency per row instead of one per partial
    float4 pt[2][8];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k < a.nsplit) pt[i][k] = *reinterpret_cast<const float4*>(a.partial + k * a.split_stride + oi + i * 128 + lane * 4);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int c = i * 128 + lane * 4;
        float4 s = pt[i][0];
#pragma unroll
        for (int k = 1; k < 8; ++k) {
            if (k < a.nsplit) { s.x += pt[i][k].x; s.y += pt[i][k].y; s.z += pt[i][k].z; s.w += pt[i][k].w; }
        }
        if (a.bias) {
            float4 t = __ldg(reinterpret_cast<const float4*>(a.bias + grp * a.group_stride + c));
            s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
        }
        if (a.residual) {
            float4 t = *reinterpret_cast<const float4*>(a.residual + o + c);
            s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
        }
        v[i * 4 + 0] = s.x; v[i * 4 + 1] = s.y; v[i * 4 + 2] = s.z; v[i * 4 + 3] = s.w;
    }
    ln_tail(a, row, grp, lane, v);
}

// v = the 8 values of this lane (columns i*128 + lane*4 + k, i = 0..1, k = 0..3) of row `row`, already summed over the
// split-K partials and with bias + residual added: LayerNorm (+ ReLU, + query_pos copy, + second LayerNorm, + TF32 splits)
__device__ __forceinline__ void ln_tail(const LnArgs& a, int row, int grp, int lane, float (&v)[8]) {
    const long long o = (long long)row * MV2D_C;
    auto norm = [&](const float* g, const float* b, float* y) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += v[i];
        const float mean = warp_sum(s) * (1.f / MV2D_C);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { float d = v[i] - mean; q += d * d; }
        const float rstd = rsqrtf(warp_sum(q) * (1.f / MV2D_C) + 1e-5f);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int c = i * 128 + lane * 4;
            float4 gg = __ldg(reinterpret_cast<const float4*>(g + c));
            float4 bb = __ldg(reinterpret_cast<const float4*>(b + c));
            y[i * 4 + 0] = (v[i * 4 + 0] - mean) * rstd * gg.x + bb.x;
            y[i * 4 + 1] = (v[i * 4 + 1] - mean) * rstd * gg.y + bb.y;
            y[i * 4 + 2] = (v[i * 4 + 2] - mean) * rstd * gg.z + bb.z;
            y[i * 4 + 3] = (v[i * 4 + 3] - mean) * rstd * gg.w + bb.w;
        }
    };
    float y[8];
    norm(a.gamma + grp * a.group_stride, a.beta + grp * a.group_stride, y);
    if (a.relu) {
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = fmaxf(y[i], 0.f);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int c = i * 128 + lane * 4;
        *reinterpret_cast<float4*>(a.out + o + c) = make_float4(y[i * 4], y[i * 4 + 1], y[i * 4 + 2], y[i * 4 + 3]);
        if (a.out_hi) {
            float hi[4], lo[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { hi[k] = round_tf32(y[i * 4 + k]); lo[k] = round_tf32(y[i * 4 + k] - hi[k]); }
            *reinterpret_cast<float4*>(a.out_hi + o + c) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(a.out_lo + o + c) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
        if (a.out_q) {
            float4 qp = *reinterpret_cast<const float4*>(a.qpos + o + c);
            const float z[4] = {y[i * 4] + qp.x, y[i * 4 + 1] + qp.y, y[i * 4 + 2] + qp.z, y[i * 4 + 3] + qp.w};
            *reinterpret_cast<float4*>(a.out_q + o + c) = make_float4(z[0], z[1], z[2], z[3]);
            if (a.outq_hi) {
                float hi[4], lo[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) { hi[k] = round_tf32(z[k]); lo[k] = round_tf32(z[k] - hi[k]); }
                *reinterpret_cast<float4*>(a.outq_hi + o + c) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4*>(a.outq_lo + o + c) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
    }
    if (a.out2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = y[i];
        norm(a.gamma2, a.beta2, y);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int c = i * 128 + lane * 4;
            *reinterpret_cast<float4*>(a.out2 + o + c) = make_float4(y[i * 4], y[i * 4 + 1], y[i * 4 + 2], y[i * 4 + 3]);
            if (a.out2_hi) {
                float hi[4], lo[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) { hi[k] = round_tf32(y[i * 4 + k]); lo[k] = round_tf32(y[i * 4 + k] - hi[k]); }
                *reinterpret_cast<float4*>(a.out2_hi + o + c) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4*>(a.out2_lo + o + c) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
    }
}

}  // namespace mv2d
