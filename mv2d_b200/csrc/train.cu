// Training step of the MV2D-S hot path (SURVEY.md 8e / BASELINE configs[3]): forward WITH saved activations, Hungarian
// targets + losses, and the BACKWARD of rows a1-a18 + f3 -- position encoding, RoIAlign, query generator, query
// embedding, six decoder layers (flattened self-attention, sparse per-RoI cross-attention, FFN, LayerNorms), post-norm,
// cls / reg branches with the reference-point refinement, focal + L1 losses -- down to the gradients of every
// hot-path parameter and of the FPN feature map.
//   reference: roi_heads/mv2d_s_head.py:236-307 (forward_train: sum over layers of stage_loss_weight * loss);
//              roi_heads/bbox_heads/cross_attention_head.py:199-242 (forward), :379-434 (loss_single);
//              utils/petr_transformer.py:194-370,373-513,569-593; utils/pe.py:21-33,137-169;
//              roi_heads/utils/query_generator.py:343-405; the backward itself is torch autograd in the reference.
// The reference trains MV2D-S without denoising queries (configs/mv2d/exp/*single_frame*:44 use_denoise=False), so
// every query attends to the 49 tokens of each RoI in its own match list and self-attention is unmasked.
//
// Layout: parameters and gradients are ONE flat fp32 buffer each (mv2d_train_param_info), so the data-parallel
// gradient exchange is a single NCCL all-reduce and the optimizer a single fused pass.  The inference path keeps the
// absorbed cross-attention weights; training uses the plain in_proj / out_proj form because those are the leaves
// the optimizer updates.
//
// Files (one translation unit): train_gemm.cuh -- the contractions (strided fp32 FFMA GEMM; tcgen05 3xTF32 route for
// the GPU-filling ones; linear forward / backward-data / weight-gradient helpers), train_attn.cuh -- self- and
// cross-attention forward / backward, train_front.cuh -- RoIAlign, im2col, pooling, center2lidar, SE gate; this file:
// parameter and workspace layouts, LayerNorm, query embedding, branch tail, loss gradient, the four entry points
// that enqueue a step, fused AdamW.  DESIGN.md section 8 has the numbers.
#include <algorithm>
#include <cstdlib>
#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "mv2d_internal.h"

namespace mv2d {

namespace {

#define TRY(expr)                \
    do {                         \
        int rc__ = (expr);       \
        if (rc__ != 0) return rc__; \
    } while (0)

constexpr int TC_ = MV2D_C;        // 256
constexpr int TH = MV2D_HEADS;     // 8
constexpr int THD = MV2D_HD;       // 32
constexpr int TTOK = MV2D_TOK;     // 49
constexpr int TFF = 2048;
constexpr int TPE = 384;           // 3 x 128 sin/cos features of the query embedding
constexpr int TCODE = 10;

// ------------------------------------------------------------------------------------------------ parameter layout
enum TrainGlobal { TG_QE0_W, TG_QE0_B, TG_QE2_W, TG_QE2_B, TG_POST_G, TG_POST_B, TG_COUNT };
enum TrainLayer {
    TL_SA_IN_W, TL_SA_IN_B, TL_SA_OUT_W, TL_SA_OUT_B, TL_CA_IN_W, TL_CA_IN_B, TL_CA_OUT_W, TL_CA_OUT_B,
    TL_FFN_W1, TL_FFN_B1, TL_FFN_W2, TL_FFN_B2, TL_LN0_G, TL_LN0_B, TL_LN1_G, TL_LN1_B, TL_LN2_G, TL_LN2_B,
    TL_CLS_W0, TL_CLS_B0, TL_CLS_G0, TL_CLS_BE0, TL_CLS_W1, TL_CLS_B1, TL_CLS_G1, TL_CLS_BE1, TL_CLS_W2, TL_CLS_B2,
    TL_REG_W0, TL_REG_B0, TL_REG_W1, TL_REG_B1, TL_REG_W2, TL_REG_B2, TL_COUNT
};
const long long kGlobalNumel[TG_COUNT] = {256 * 384, 256, 256 * 256, 256, 256, 256};
const long long kLayerNumel[TL_COUNT] = {
    768 * 256, 768, 256 * 256, 256, 768 * 256, 768, 256 * 256, 256,
    2048 * 256, 2048, 256 * 2048, 256, 256, 256, 256, 256, 256, 256,
    256 * 256, 256, 256, 256, 256 * 256, 256, 256, 256, 10 * 256, 10,
    256 * 256, 256, 256 * 256, 256, 10 * 256, 10};

// front end (position encoding + query generator), after the layer blocks; the 3x3 conv weight is kept as
// [c_out, (ky, kx, c_in)] (K order of the im2col GEMM), the host permutes it from / to the state_dict layout
enum TrainFront {
    TF_POS0_W, TF_POS0_B, TF_POS2_W, TF_POS2_B, TF_ADAPT0_W, TF_ADAPT0_B, TF_ADAPT2_W, TF_ADAPT2_B,
    TF_SE_R_W, TF_SE_R_B, TF_SE_E_W, TF_SE_E_B, TF_CONV_W, TF_CONV_B, TF_FC_W, TF_FC_B, TF_ENC0_W, TF_ENC0_B,
    TF_ENC2_W, TF_ENC2_B, TF_CENTER_W, TF_CENTER_B, TF_COUNT
};
const long long kFrontNumel[TF_COUNT] = {
    1024 * 192, 1024, 256 * 1024, 256, 1024 * 384, 1024, 256 * 1024, 256, 256 * 256, 256, 256 * 256, 256,
    256 * 2304, 256, 1024 * 256, 1024, 512 * 1040, 512, 256 * 512, 256, 3 * 256, 3};

inline long long pad16(long long n) { return (n + 15) / 16 * 16; }

long long layer_block_floats() {
    long long s = 0;
    for (int t = 0; t < TL_COUNT; ++t) s += pad16(kLayerNumel[t]);
    return s;
}
long long global_block_floats() {
    long long s = 0;
    for (int t = 0; t < TG_COUNT; ++t) s += pad16(kGlobalNumel[t]);
    return s;
}
long long global_off(int t) {
    long long s = 0;
    for (int i = 0; i < t; ++i) s += pad16(kGlobalNumel[i]);
    return s;
}
long long layer_off(int l, int t) {
    long long s = global_block_floats() + (long long)l * layer_block_floats();
    for (int i = 0; i < t; ++i) s += pad16(kLayerNumel[i]);
    return s;
}
long long front_block_floats() {
    long long s = 0;
    for (int t = 0; t < TF_COUNT; ++t) s += pad16(kFrontNumel[t]);
    return s;
}
long long front_off(int L, int t) {
    long long s = global_block_floats() + (long long)L * layer_block_floats();
    for (int i = 0; i < t; ++i) s += pad16(kFrontNumel[i]);
    return s;
}

#include "train_gemm.cuh"

// ------------------------------------------------------------------------------------------------ LayerNorm
// y = a (+ b); xhat = (y - mean) * rstd; out = [relu](xhat * g + beta).  One warp per row of 256.
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                     const float* __restrict__ g, const float* __restrict__ beta,
                                                     float* __restrict__ out, float* __restrict__ xhat, float* __restrict__ rstd,
                                                     int N, int relu) {
    pdl_wait();
    pdl_trigger();
    const int lane = threadIdx.x & 31, r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= N) return;
    float v[8], s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const long long o = (long long)r * TC_ + lane + 32 * k;
        v[k] = a[o] + (b ? b[o] : 0.f);
        s += v[k];
    }
    const float mean = warp_sum(s) * (1.f / TC_);
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { const float d = v[k] - mean; q += d * d; }
    const float var = warp_sum(q) * (1.f / TC_);
    const float rs = 1.f / sqrtf(var + 1e-5f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = lane + 32 * k;
        const long long o = (long long)r * TC_ + c;
        const float xh = (v[k] - mean) * rs;
        xhat[o] = xh;
        float y = xh * g[c] + beta[c];
        if (relu) y = fmaxf(y, 0.f);
        out[o] = y;
    }
    if (lane == 0) rstd[r] = rs;
}
int ln_fwd(const float* a, const float* b, const float* g, const float* beta, float* out, float* xhat, float* rstd, int N,
           bool relu, cudaStream_t st) {
    if (N <= 0) return 0;
    launch_k(ln_fwd_kernel, dim3(cdiv(N, 8)), dim3(256), 0, st, a, b, g, beta, out, xhat, rstd, N, relu ? 1 : 0);
    MV2D_CHECK_LAUNCH("train ln_fwd");
    return 0;
}

// dy (masked by act > 0 when act != NULL: the ReLU that followed the norm) -> dx (written, or added when accumulate),
// dg += sum_rows dy * xhat,  db += sum_rows dy
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ act,
                                                     const float* __restrict__ xhat, const float* __restrict__ rstd,
                                                     const float* __restrict__ g, float* __restrict__ dx, float* __restrict__ dg,
                                                     float* __restrict__ db, int N, int accumulate) {
    pdl_wait();
    pdl_trigger();
    __shared__ float sg[8][TC_], sb[8][TC_];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float pg[8], pb[8], gg[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { pg[k] = 0.f; pb[k] = 0.f; gg[k] = g[lane + 32 * k]; }
    for (int r = blockIdx.x * 8 + warp; r < N; r += gridDim.x * 8) {
        float d[8], xh[8], s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const long long o = (long long)r * TC_ + lane + 32 * k;
            d[k] = dy[o];
            if (act && !(act[o] > 0.f)) d[k] = 0.f;
            xh[k] = xhat[o];
            pb[k] += d[k];
            pg[k] += d[k] * xh[k];
            const float t = d[k] * gg[k];
            s1 += t;
            s2 += t * xh[k];
        }
        s1 = warp_sum(s1) * (1.f / TC_);
        s2 = warp_sum(s2) * (1.f / TC_);
        const float rs = rstd[r];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const long long o = (long long)r * TC_ + lane + 32 * k;
            const float v = rs * (d[k] * gg[k] - s1 - xh[k] * s2);
            dx[o] = accumulate ? dx[o] + v : v;
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) { sg[warp][lane + 32 * k] = pg[k]; sb[warp][lane + 32 * k] = pb[k]; }
    __syncthreads();
    const int c = threadIdx.x;
    float tg = 0.f, tb = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { tg += sg[i][c]; tb += sb[i][c]; }
    atomicAdd(dg + c, tg);
    atomicAdd(db + c, tb);
}
int ln_bwd(const float* dy, const float* act, const float* xhat, const float* rstd, const float* g, float* dx, float* dg,
           float* db, int N, bool accumulate, cudaStream_t st) {
    if (N <= 0) return 0;
    launch_k(ln_bwd_kernel, dim3(std::min(cdiv(N, 8), 148)), dim3(256), 0, st, dy, act, xhat, rstd, g, dx, dg, db, N, accumulate ? 1 : 0);
    MV2D_CHECK_LAUNCH("train ln_bwd");
    return 0;
}

// ------------------------------------------------------------------------------------------------ query embedding
// pos2posemb3d (utils/pe.py:21-33): [N,3] -> [N,384], blocks (y, x, z), interleaved sin / cos of ref * 2 pi / dim_t
__global__ void __launch_bounds__(128) posemb_fwd_kernel(const float* __restrict__ ref, const float* __restrict__ dim_t,
                                                         float* __restrict__ out, int N) {
    pdl_wait();
    pdl_trigger();
    const int n = blockIdx.x, i = threadIdx.x;
    if (n >= N) return;
#pragma unroll
    for (int blk = 0; blk < 3; ++blk) {
        const int comp = blk == 0 ? 1 : (blk == 1 ? 0 : 2);
        const float p = ref[n * 3 + comp] * 6.283185307179586f;
        const float arg = p / dim_t[i];
        out[(long long)n * TPE + blk * 128 + i] = (i & 1) ? cosf(arg) : sinf(arg);
    }
}
// d_ref[n][comp] += sum_i dpe[n][blk*128+i] * d/dref
__global__ void __launch_bounds__(96) posemb_bwd_kernel(const float* __restrict__ dpe, const float* __restrict__ ref,
                                                        const float* __restrict__ dim_t, float* __restrict__ d_ref, int N) {
    pdl_wait();
    pdl_trigger();
    const int n = blockIdx.x, blk = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (n >= N) return;
    const int comp = blk == 0 ? 1 : (blk == 1 ? 0 : 2);
    const float p = ref[n * 3 + comp] * 6.283185307179586f;
    float a = 0.f;
    for (int i = lane; i < 128; i += 32) {
        const float w = 6.283185307179586f / dim_t[i];
        const float arg = p / dim_t[i];
        const float d = dpe[(long long)n * TPE + blk * 128 + i];
        a += (i & 1) ? -d * sinf(arg) * w : d * cosf(arg) * w;
    }
    a = warp_sum(a);
    if (lane == 0) d_ref[n * 3 + comp] += a;
}

#include "train_attn.cuh"

// ------------------------------------------------------------------------------------------------ reg branch tail
struct Range6 { float v[6]; };

// cross_attention_head.py:224-237: codes 0,1 (+ inverse_sigmoid(ref) 0,1) and 4 (+ ref 2) go through a sigmoid and
// are scaled to pc_range; sig [N,4] keeps the three sigmoid values for the backward
// vel_dt != 0: rows >= vel_row_start (the matching queries of the two-frame head) get their velocity outputs divided by
// the frame interval (mv2d_t_head.py:130-142); the denoising rows in front are left alone (:112-118 come first)
__global__ void __launch_bounds__(128) reg_tail_fwd_kernel(float* __restrict__ box, const float* __restrict__ ref,
                                                           float* __restrict__ sig, Range6 pc, int N, float vel_dt, int vel_row_start) {
    pdl_wait();
    pdl_trigger();
    const int n = blockIdx.x * 128 + threadIdx.x;
    if (n >= N) return;
    float* b = box + (long long)n * TCODE;
    if (vel_dt != 0.f && n >= vel_row_start) { b[8] = b[8] / vel_dt; b[9] = b[9] / vel_dt; }
    const float s0 = sigmoid_f(b[0] + inverse_sigmoid_f(ref[n * 3 + 0]));
    const float s1 = sigmoid_f(b[1] + inverse_sigmoid_f(ref[n * 3 + 1]));
    const float s4 = sigmoid_f(b[4] + inverse_sigmoid_f(ref[n * 3 + 2]));
    b[0] = s0 * (pc.v[3] - pc.v[0]) + pc.v[0];
    b[1] = s1 * (pc.v[4] - pc.v[1]) + pc.v[1];
    b[4] = s4 * (pc.v[5] - pc.v[2]) + pc.v[2];
    sig[n * 4 + 0] = s0; sig[n * 4 + 1] = s1; sig[n * 4 + 2] = s4; sig[n * 4 + 3] = 0.f;
}

// d inverse_sigmoid(x) / dx with mmdet's clamps (x in [0,1], numerator / denominator floored at 1e-5)
__device__ __forceinline__ float inverse_sigmoid_grad(float x) {
    if (x < 0.f || x > 1.f) return 0.f;
    float g = 0.f;
    if (x >= 1e-5f) g += 1.f / x;
    if (1.f - x >= 1e-5f) g += 1.f / (1.f - x);
    return g;
}

// dbox [N,10] -> gradient of the raw reg output in place; d_ref += through inverse_sigmoid(ref)
__global__ void __launch_bounds__(128) reg_tail_bwd_kernel(float* __restrict__ dbox, const float* __restrict__ ref,
                                                           const float* __restrict__ sig, float* __restrict__ d_ref, Range6 pc, int N,
                                                           float vel_dt, int vel_row_start) {
    pdl_wait();
    pdl_trigger();
    const int n = blockIdx.x * 128 + threadIdx.x;
    if (n >= N) return;
    float* d = dbox + (long long)n * TCODE;
    if (vel_dt != 0.f && n >= vel_row_start) { d[8] = d[8] / vel_dt; d[9] = d[9] / vel_dt; }
    const float s0 = sig[n * 4 + 0], s1 = sig[n * 4 + 1], s4 = sig[n * 4 + 2];
    const float g0 = d[0] * (pc.v[3] - pc.v[0]) * s0 * (1.f - s0);
    const float g1 = d[1] * (pc.v[4] - pc.v[1]) * s1 * (1.f - s1);
    const float g4 = d[4] * (pc.v[5] - pc.v[2]) * s4 * (1.f - s4);
    d[0] = g0; d[1] = g1; d[4] = g4;
    d_ref[n * 3 + 0] += g0 * inverse_sigmoid_grad(ref[n * 3 + 0]);
    d_ref[n * 3 + 1] += g1 * inverse_sigmoid_grad(ref[n * 3 + 1]);
    d_ref[n * 3 + 2] += g4 * inverse_sigmoid_grad(ref[n * 3 + 2]);
}

// ------------------------------------------------------------------------------------------------ loss gradient
// d (sum_l w_l (loss_cls_l + loss_bbox_l)) / d cls_scores, bbox_preds given the assignment of mv2d_loss
// (cross_attention_head.py:379-434; mmdet py_sigmoid_focal_loss, L1Loss; avg_factor = max(num_pos, 1) + eps).
struct LossGradArgs {
    const float* cls; const float* box; const int* assigned; const float* gt_boxes; const int* gt_labels;
    float* dcls; float* dbox;
    int N, G, num_classes;
    float alpha, gamma, cls_lw, box_lw;
    float code_w[TCODE];
    float stage_w[MV2D_MAX_LAYERS];
    const float* bbox_avg_factor;     // nullable [L]: cross-rank avg factor of loss_bbox (cross_attention_head.py:419-420)
    float* losses;                    // [L,4]: with bbox_avg_factor, losses[l][1] is rescaled from the local factor to it
    // denoising queries (dn_loss_single, cross_attention_head.py:475-538): the first `pad` of the NT = pad + N rows of
    // every layer; N, assigned refer to the matching rows behind them
    int pad; const int* dn_labels; float dn_split, dn_weight; int neg_bbox_loss;
};

__global__ void __launch_bounds__(256) loss_grad_kernel(LossGradArgs a) {
    pdl_wait();
    pdl_trigger();
    const int l = blockIdx.x, tid = threadIdx.x;
    const int* asg = a.assigned + (long long)l * a.N;
    __shared__ int sp[256];
    int np = 0;
    for (int n = tid; n < a.N; n += 256) np += asg[n] >= 0;
    sp[tid] = np;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) sp[tid] += sp[tid + o];
        __syncthreads();
    }
    const double avg = fmax((double)sp[0], 1.0) + 1.1920928955078125e-07;
    const double avg_box = a.bbox_avg_factor ? (double)a.bbox_avg_factor[l] + 1.1920928955078125e-07 : avg;
    const float kc_m = (float)((double)a.cls_lw * a.stage_w[l] / avg);
    const float kb_m = (float)((double)a.box_lw * a.stage_w[l] / avg_box);
    if (a.bbox_avg_factor && a.losses && tid == 0) a.losses[l * 4 + 1] = (float)((double)a.losses[l * 4 + 1] * avg / avg_box);
    // denoising rows: avg factors pad * pi / 6 * split^3 (classification) and pad (boxes), both floored at 1
    const double dn_cls_avg = fmax((double)a.pad * 3.14159 / 6.0 * a.dn_split * a.dn_split * a.dn_split, 1.0) + 1.1920928955078125e-07;
    const double dn_box_avg = fmax((double)a.pad, 1.0) + 1.1920928955078125e-07;
    const float kc_d = (float)((double)a.cls_lw * a.stage_w[l] * a.dn_weight / dn_cls_avg);
    const float kb_d = (float)((double)a.box_lw * a.stage_w[l] * a.dn_weight / dn_box_avg);
    const int NT = a.pad + a.N;
    for (int row = tid; row < NT; row += 256) {
        const bool dn = row < a.pad;
        const int n = row - a.pad;
        int g, label;
        if (dn) {
            label = a.dn_labels[row];
            g = (a.G > 0 && (a.neg_bbox_loss || label != a.num_classes)) ? row % a.G : -1;
        } else {
            g = asg[n];
            label = g >= 0 ? a.gt_labels[g] : a.num_classes;
        }
        const float kc = dn ? kc_d : kc_m, kb = dn ? kb_d : kb_m;
        const long long oc = ((long long)l * NT + row) * a.num_classes;
        for (int c = 0; c < a.num_classes; ++c) {
            const float x = a.cls[oc + c];
            const float p = 1.f / (1.f + expf(-x));
            // log p = -softplus(-x), log(1 - p) = -softplus(x), both stable
            const float l1p = log1pf(expf(-fabsf(x)));
            const float logp = fminf(x, 0.f) - l1p, log1mp = fminf(-x, 0.f) - l1p;
            float d;
            if (c == label) {
                // a (1-p)^g (-log p):  d/dx = a [ g (1-p)^(g-1) (-p (1-p)) (-log p) - (1-p)^g (1-p) ]
                const float om = 1.f - p;
                d = a.alpha * (a.gamma * powf(om, a.gamma - 1.f) * p * om * logp - powf(om, a.gamma) * om);
            } else {
                // (1-a) p^g (-log(1-p)):  d/dx = (1-a) [ g p^(g-1) p (1-p) (-log(1-p)) + p^g p ]
                d = (1.f - a.alpha) * (-a.gamma * powf(p, a.gamma - 1.f) * p * (1.f - p) * log1mp + powf(p, a.gamma) * p);
            }
            a.dcls[oc + c] = d * kc;
        }
        const long long ob = ((long long)l * NT + row) * TCODE;
        float gn[TCODE];
        bool ok = false;
        if (g >= 0) {
            const float* b = a.gt_boxes + g * 9;
            gn[0] = b[0]; gn[1] = b[1]; gn[2] = logf(b[3]); gn[3] = logf(b[4]); gn[4] = b[2]; gn[5] = logf(b[5]);
            gn[6] = sinf(b[6]); gn[7] = cosf(b[6]); gn[8] = b[7]; gn[9] = b[8];
            ok = true;
#pragma unroll
            for (int j = 0; j < TCODE; ++j) ok = ok && isfinite(gn[j]);
        }
#pragma unroll
        for (int j = 0; j < TCODE; ++j) {
            float d = 0.f;
            if (ok) {
                const float e = a.box[ob + j] - gn[j];
                const float cw = (dn && (j == 6 || j == 7)) ? 0.f : a.code_w[j];     // the denoising loss leaves sin / cos out (:527)
                d = (e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f)) * cw * kb;
            }
            a.dbox[ob + j] = d;
        }
    }
}

// ------------------------------------------------------------------------------------------------ workspace layout
struct LayerAct {
    float *x_in, *xq, *qkv, *P_sa, *attn_o, *xhat0, *rstd0, *x1, *xq1, *cq, *Kp, *Vp, *P_xa, *ctx, *xhat1, *rstd1, *x2,
        *hdn, *xhat2, *rstd2, *x3, *xhatp, *rstdp, *inter, *xhat_c0, *rstd_c0, *c0n, *xhat_c1, *rstd_c1, *c1n, *r0, *r1, *rsig, *xa_stats;
};
struct TrainWs {
    float *posemb, *h0, *qpos, *zero;
    LayerAct layer[MV2D_MAX_LAYERS];
    // backward scratch
    float *dx, *dqpos, *t1, *t2, *t3, *dinter, *dcq, *dhdn, *dqkv, *dS_sa, *dS_xa, *dKp, *dVp, *dcls, *dbox, *dposemb;
    int *inv_cnt, *inv_list;
    float* loss_ws;
    size_t loss_ws_bytes;
    TcScratch tc;
    size_t total_bytes;
};

// N = query rows of the decoder (pad + matching queries in mode 1), Nm = matching queries (Hungarian loss);
// mode 1 (two-frame head): keys = R feature cells -> Kp / Vp [R,256] per layer, P / dS dense [N,8,R]
TrainWs train_layout(float* base, int N, int L, int max_match, int G, int mode = 0, int R = 0, int Nm = -1) {
    TrainWs w{};
    size_t off = 0;   // in floats
    auto take = [&](size_t n) -> float* {
        float* p = base ? base + off : nullptr;
        off += (n + 63) / 64 * 64;
        return p;
    };
    if (Nm < 0) Nm = N;
    const size_t n = (size_t)(N > 0 ? N : 1), NC = n * TC_;
    const size_t krows = mode == 1 ? (size_t)(R > 0 ? R : 1) : n * TTOK;          // key rows of the cross-attention
    const size_t PM = mode == 1 ? krows : (size_t)max_match * TTOK, NK = krows * TC_;
    w.posemb = take(n * TPE); w.h0 = take(NC); w.qpos = take(NC); w.zero = take(NC);
    for (int l = 0; l < L; ++l) {
        LayerAct& a = w.layer[l];
        a.xq = take(NC); a.qkv = take(n * 768); a.P_sa = take(TH * n * n); a.attn_o = take(NC);
        a.xhat0 = take(NC); a.rstd0 = take(n); a.x1 = take(NC); a.xq1 = take(NC); a.cq = take(NC);
        a.Kp = take(NK); a.Vp = take(NK); a.P_xa = take(n * TH * PM); a.ctx = take(NC);
        a.xhat1 = take(NC); a.rstd1 = take(n); a.x2 = take(NC); a.hdn = take(n * TFF);
        a.xhat2 = take(NC); a.rstd2 = take(n); a.x3 = take(NC);
        a.xhatp = take(NC); a.rstdp = take(n); a.inter = take(NC);
        a.xhat_c0 = take(NC); a.rstd_c0 = take(n); a.c0n = take(NC);
        a.xhat_c1 = take(NC); a.rstd_c1 = take(n); a.c1n = take(NC);
        a.r0 = take(NC); a.r1 = take(NC); a.rsig = take(n * 4); a.xa_stats = take(n * TH * 2);
        a.x_in = l == 0 ? w.zero : w.layer[l - 1].x3;
    }
    w.dx = take(NC); w.dqpos = take(NC); w.t1 = take(NC); w.t2 = take(NC); w.t3 = take(NC); w.dinter = take(NC);
    w.dcq = take(NC); w.dhdn = take(n * TFF); w.dqkv = take(n * 768); w.dS_sa = take(TH * n * n); w.dS_xa = take(n * TH * PM);
    w.dKp = take(NK * L); w.dVp = take(NK * L); w.dcls = take((size_t)L * n * TCODE); w.dbox = take((size_t)L * n * TCODE);
    w.dposemb = take(n * TPE);
    w.inv_cnt = reinterpret_cast<int*>(take(n));
    w.inv_list = reinterpret_cast<int*>(take(mode == 1 ? 64 : n * n));
    w.loss_ws_bytes = loss_workspace_bytes(Nm, G, L);
    w.loss_ws = take(w.loss_ws_bytes / 4 + 1);
    {   // tensor-core route: transposed operands of the K/V weight gradients, W^T, split-K partials, accumulate temp
        const size_t Mp = (size_t)round32((int)krows);
        // at: [L*256, Mp] (the K / V gradients of all layers transposed at once); part: split-K partials of a [L*256, 256] product
        w.tc.at_cap = (size_t)L * TC_ * Mp; w.tc.bt_cap = TC_ * Mp; w.tc.wt_cap = (size_t)(L * TC_ > TFF ? L * TC_ : TFF) * TC_;
        w.tc.part_cap = (size_t)48 * L * TC_ * TC_;
        w.tc.tmp_cap = NK;
        w.tc.at = take(w.tc.at_cap); w.tc.bt = take(w.tc.bt_cap); w.tc.wt = take(w.tc.wt_cap); w.tc.part = take(w.tc.part_cap);
        w.tc.tmp = take(w.tc.tmp_cap);
    }
    w.total_bytes = off * sizeof(float);
    return w;
}

struct LayerPtr {   // views into a flat parameter (or gradient) buffer
    float* t[TL_COUNT];
};
LayerPtr layer_ptrs(float* flat, int l) {
    LayerPtr p{};
    for (int t = 0; t < TL_COUNT; ++t) p.t[t] = flat + layer_off(l, t);
    return p;
}


int check_params(const Mv2dTrainParams& p) {
    MV2D_CHECK_ARG(p.N >= 1 && p.L >= 1 && p.L <= MV2D_MAX_LAYERS && p.max_match >= 1 && p.G >= 0, "train: bad N=%d L=%d max_match=%d G=%d",
                   p.N, p.L, p.max_match, p.G);
    MV2D_CHECK_ARG(p.num_classes == TCODE, "train: num_classes must be 10 (got %d)", p.num_classes);
    MV2D_CHECK_ARG(p.mode == 0 || p.mode == 1, "train: mode must be 0 (single-frame) or 1 (two-frame)");
    MV2D_CHECK_ARG(p.params && p.ref, "train: null input");
    if (p.mode == 0) {
        MV2D_CHECK_ARG(p.tok_kin && p.tok_mem && p.match && p.match_cnt, "train: null input");
        MV2D_CHECK_ARG(p.pad == 0 && !p.self_attn_mask, "train: denoising queries need mode 1 (the reference trains MV2D-S without them)");
    } else {
        MV2D_CHECK_ARG(p.pad >= 0 && p.num_rows >= 1 && p.num_rows <= 65536 && p.mask_words * 32 >= p.num_rows,
                       "train: mode 1 needs 1 <= num_rows <= 65536 and mask_words covering them (num_rows=%d mask_words=%d)", p.num_rows, p.mask_words);
        MV2D_CHECK_ARG(p.kin_map && p.mem_map && p.keymask && p.key_list && p.key_cnt, "train: mode 1 needs kin_map / mem_map / keymask / key_list / key_cnt");
        MV2D_CHECK_ARG(p.pad == 0 || (p.dn_labels && p.G > 0), "train: denoising queries need dn_labels and ground truth");
        MV2D_CHECK_ARG(((uintptr_t)p.kin_map & 15) == 0 && ((uintptr_t)p.mem_map & 15) == 0, "train: kin_map / mem_map must be 16-byte aligned");
    }
    MV2D_CHECK_ARG(p.dim_t, "train: null dim_t");
    MV2D_CHECK_ARG(p.cls_scores && p.bbox_preds && p.assigned && p.losses, "train: null output");
    MV2D_CHECK_ARG(p.G == 0 || (p.gt_boxes && p.gt_labels), "train: null ground truth");
    MV2D_CHECK_ARG(((uintptr_t)p.params & 15) == 0 && ((uintptr_t)p.tok_kin & 15) == 0 && ((uintptr_t)p.tok_mem & 15) == 0,
                   "train: params / tokens must be 16-byte aligned");
    MV2D_CHECK_ARG(p.workspace && ((uintptr_t)p.workspace & 255) == 0, "train: workspace must be 256-byte aligned");
    const TrainWs w = train_layout(nullptr, p.pad + p.N, p.L, p.max_match, p.G, p.mode, p.num_rows, p.N);
    MV2D_CHECK_ARG(p.workspace_bytes >= w.total_bytes, "train: workspace too small (%zu < %zu)", p.workspace_bytes, w.total_bytes);
    return 0;
}


#include "train_front.cuh"

struct FrontWs {
    // position encoding ([P, .])
    float *coords, *sine, *hp, *x, *hs, *sb, *g1, *gate, *pe;
    // query generator
    float *tok_pe, *col, *y, *pool, *cat, *e0, *e2, *c, *m_roi;
    // backward scratch
    float *dc, *de2, *de0, *df1, *dpool, *dy, *dcol, *dtok, *dpe_map, *dx, *dg2, *dh;
    TcScratch tc;
    size_t total_bytes;
};
FrontWs front_layout(float* base, int N, int P) {
    FrontWs w{};
    size_t off = 0;
    auto take = [&](size_t n) -> float* {
        float* p = base ? base + off : nullptr;
        off += (n + 63) / 64 * 64;
        return p;
    };
    const size_t n = (size_t)(N > 0 ? N : 1), p = (size_t)P, NK = n * TTOK * TC_;
    w.coords = take(p * 192); w.sine = take(p * 387); w.hp = take(p * 1024); w.x = take(p * TC_); w.hs = take(p * 1024);
    w.sb = take(p * TC_); w.g1 = take(p * TC_); w.gate = take(p * TC_); w.pe = take(p * TC_);
    w.tok_pe = take(NK); w.col = take(NK * 9); w.y = take(NK); w.pool = take(n * TC_); w.cat = take(n * 1040);
    w.e0 = take(n * 512); w.e2 = take(n * TC_); w.c = take(n * 4); w.m_roi = take(n * 16);
    w.dc = take(n * 4); w.de2 = take(n * TC_); w.de0 = take(n * 512); w.df1 = take(n * 1024); w.dpool = take(n * TC_);
    w.dy = take(NK); w.dcol = take(NK * 9); w.dtok = take(NK); w.dpe_map = take(p * TC_); w.dx = take(p * TC_);
    w.dg2 = take(p * TC_); w.dh = take(p * 1024);
    {   // tensor-core route (widest operands: the im2col matrix [*, 2304] and the 1024-wide MLP hidden layers)
        const size_t rows = p > n * TTOK ? p : n * TTOK, Mp = (size_t)round32((int)rows);
        w.tc.at_cap = 1024 * Mp; w.tc.bt_cap = 2304 * Mp; w.tc.wt_cap = (size_t)2304 * 256; w.tc.part_cap = (size_t)48 * 2304 * 256;
        w.tc.tmp_cap = rows * TC_;
        w.tc.at = take(w.tc.at_cap); w.tc.bt = take(w.tc.bt_cap); w.tc.wt = take(w.tc.wt_cap); w.tc.part = take(w.tc.part_cap);
        w.tc.tmp = take(w.tc.tmp_cap);
    }
    w.total_bytes = off * sizeof(float);
    return w;
}

int check_front(const Mv2dFrontTrainParams& p) {
    MV2D_CHECK_ARG(p.N >= 1 && p.V >= 1 && p.V <= MV2D_MAXV && p.h >= 1 && p.w >= 1 && p.L >= 1 && p.L <= MV2D_MAX_LAYERS,
                   "front train: bad N=%d V=%d h=%d w=%d L=%d", p.N, p.V, p.h, p.w, p.L);
    MV2D_CHECK_ARG(p.depth_num == 64, "front train: depth_num must be 64 (position_encoder.0 has 192 inputs)");
    MV2D_CHECK_ARG(p.params && p.rois && p.roi_intrinsics && p.extrinsics && p.feat && p.img2lidar && p.not_mask && p.dim_t,
                   "front train: null input");
    MV2D_CHECK_ARG(p.tok_mem && p.tok_kin && p.ref, "front train: null output");
    MV2D_CHECK_ARG(((uintptr_t)p.params & 15) == 0 && ((uintptr_t)p.feat & 15) == 0 && ((uintptr_t)p.tok_mem & 15) == 0 &&
                   ((uintptr_t)p.tok_kin & 15) == 0, "front train: buffers must be 16-byte aligned");
    MV2D_CHECK_ARG(p.workspace && ((uintptr_t)p.workspace & 255) == 0, "front train: workspace must be 256-byte aligned");
    const FrontWs w = front_layout(nullptr, p.N, p.V * p.h * p.w);
    MV2D_CHECK_ARG(p.workspace_bytes >= w.total_bytes, "front train: workspace too small (%zu < %zu)", p.workspace_bytes, w.total_bytes);
    return 0;
}

}  // namespace

// ================================================================================================ host entry points
int train_set_tensor_cores(int on) {
    const int prev = tc_enabled() ? 1 : 0;
    g_tc_mode = on ? 1 : 0;
    return prev;
}

long long train_param_total(int L) { return global_block_floats() + (long long)L * layer_block_floats() + front_block_floats(); }

int train_param_info(int L, int tensor_id, long long* offset, long long* numel) {
    MV2D_CHECK_ARG(L >= 1 && L <= MV2D_MAX_LAYERS, "train_param_info: bad L=%d", L);
    MV2D_CHECK_ARG(tensor_id >= 0 && tensor_id < TG_COUNT + L * TL_COUNT + TF_COUNT, "train_param_info: bad tensor id %d", tensor_id);
    if (tensor_id < TG_COUNT) {
        if (offset) *offset = global_off(tensor_id);
        if (numel) *numel = kGlobalNumel[tensor_id];
    } else if (tensor_id >= TG_COUNT + L * TL_COUNT) {
        const int t = tensor_id - TG_COUNT - L * TL_COUNT;
        if (offset) *offset = front_off(L, t);
        if (numel) *numel = kFrontNumel[t];
    } else {
        const int l = (tensor_id - TG_COUNT) / TL_COUNT, t = (tensor_id - TG_COUNT) % TL_COUNT;
        if (offset) *offset = layer_off(l, t);
        if (numel) *numel = kLayerNumel[t];
    }
    return 0;
}

size_t train_workspace_bytes(int N, int L, int max_match, int G) {
    return train_layout(nullptr, N, L, max_match, G).total_bytes;
}
size_t train_workspace_bytes_p(const Mv2dTrainParams& p) {
    return train_layout(nullptr, p.pad + p.N, p.L, p.max_match > 0 ? p.max_match : 1, p.G, p.mode, p.num_rows, p.N).total_bytes;
}

// float offset of a saved activation inside the workspace (tests compare them with the oracle's intermediates):
// which = 0 qpos [N,256] (layer ignored); 1 x1 (after norms.0), 2 x2 (after norms.1), 3 x3 (layer output),
// 4 inter (post-normed), 5 ctx (cross-attention context), 6 attn_o (self-attention context).  -1 = unknown.
long long train_debug_offset(int N, int L, int max_match, int G, int layer, int which) {
    if (N < 1 || L < 1 || L > MV2D_MAX_LAYERS || layer < 0 || layer >= L) return -1;
    float* base = reinterpret_cast<float*>(uintptr_t(1) << 20);
    const TrainWs w = train_layout(base, N, L, max_match, G);
    const LayerAct& a = w.layer[layer];
    const float* t = nullptr;
    switch (which) {
        case 0: t = w.qpos; break;
        case 1: t = a.x1; break;
        case 2: t = a.x2; break;
        case 3: t = a.x3; break;
        case 4: t = a.inter; break;
        case 5: t = a.ctx; break;
        case 6: t = a.attn_o; break;
        default: return -1;
    }
    return (long long)(t - base);
}

int run_train_forward(const Mv2dTrainParams& p, cudaStream_t st) {
    TRY(check_params(p));
    // N = query rows of the decoder: the denoising rows (mode 1) come first, then the p.N matching queries
    const int N = p.pad + p.N, L = p.L, NK = p.mode == 1 ? p.num_rows : N * TTOK;
    const bool two_frame = p.mode == 1;
    const TrainWs w = train_layout(p.workspace, N, L, p.max_match, p.G, p.mode, p.num_rows, p.N);
    g_tc = w.tc;
    float* P = const_cast<float*>(p.params);
    Range6 pc;
    for (int i = 0; i < 6; ++i) pc.v[i] = p.pc_range[i];

    // query embedding of the reference points (cross_attention_head.py:199-206)
    launch_k(posemb_fwd_kernel, dim3(N), dim3(128), 0, st, p.ref, p.dim_t, w.posemb, N);
    MV2D_CHECK_LAUNCH("train posemb");
    TRY(linear_fwd(w.posemb, TPE, P + global_off(TG_QE0_W), TPE, P + global_off(TG_QE0_B), w.h0, TC_, N, TC_, TPE, true, st));
    TRY(linear_fwd(w.h0, TC_, P + global_off(TG_QE2_W), TC_, P + global_off(TG_QE2_B), w.qpos, TC_, N, TC_, TC_, false, st));
    cudaError_t e = cudaMemsetAsync(w.zero, 0, (size_t)N * TC_ * sizeof(float), st);   // target = 0 (cross_attention_head.py:32)
    if (e != cudaSuccess) { set_error("train: memset %s", cudaGetErrorString(e)); return (int)e; }

    const float* post_g = P + global_off(TG_POST_G);
    const float* post_b = P + global_off(TG_POST_B);
    for (int l = 0; l < L; ++l) {
        const LayerAct& a = w.layer[l];
        const LayerPtr W = layer_ptrs(P, l);
        // --- flattened self-attention over all N queries (petr_transformer.py:314-370)
        TRY(add(a.xq, a.x_in, w.qpos, (long long)N * TC_, st));
        TRY(linear_fwd(a.xq, TC_, W.t[TL_SA_IN_W], TC_, W.t[TL_SA_IN_B], a.qkv, 768, N, 512, TC_, false, st));
        TRY(linear_fwd(a.x_in, TC_, W.t[TL_SA_IN_W] + 512 * TC_, TC_, W.t[TL_SA_IN_B] + 512, a.qkv + 512, 768, N, TC_, TC_, false, st));
        if (sa_use_smem(N)) {
            TRY(sa_set_attr());
            launch_k(sa_fwd_smem_kernel, dim3(cdiv(N, SA_QB), TH), dim3(256), sa_smem_bytes(N), st, (const float*)a.qkv, a.P_sa, a.attn_o, N,
                     p.self_attn_mask);
        } else {
            launch_k(sa_fwd_kernel, dim3(N), dim3(256), 0, st, (const float*)a.qkv, a.P_sa, a.attn_o, N, p.self_attn_mask);
        }
        MV2D_CHECK_LAUNCH("train sa_fwd");
        TRY(linear_fwd(a.attn_o, TC_, W.t[TL_SA_OUT_W], TC_, W.t[TL_SA_OUT_B], w.t1, TC_, N, TC_, TC_, false, st));
        TRY(ln_fwd(a.x_in, w.t1, W.t[TL_LN0_G], W.t[TL_LN0_B], a.x1, a.xhat0, a.rstd0, N, false, st));
        // --- sparse cross-attention over the matched RoIs' tokens (petr_transformer.py:373-513)
        TRY(add(a.xq1, a.x1, w.qpos, (long long)N * TC_, st));
        TRY(linear_fwd(a.xq1, TC_, W.t[TL_CA_IN_W], TC_, W.t[TL_CA_IN_B], a.cq, TC_, N, TC_, TC_, false, st));
        const float* kin = two_frame ? p.kin_map : p.tok_kin;
        const float* mem = two_frame ? p.mem_map : p.tok_mem;
        TRY(linear_fwd(kin, TC_, W.t[TL_CA_IN_W] + 256 * TC_, TC_, W.t[TL_CA_IN_B] + 256, a.Kp, TC_, NK, TC_, TC_, false, st));
        TRY(linear_fwd(mem, TC_, W.t[TL_CA_IN_W] + 512 * TC_, TC_, W.t[TL_CA_IN_B] + 512, a.Vp, TC_, NK, TC_, TC_, false, st));
        if (two_frame && xt2_enabled()) {
            launch_k(xt2_fwd_kernel, dim3(cdiv(N, XT2_QB), TH), dim3(256), 0, st, (const float*)a.cq, (const float*)a.Kp, (const float*)a.Vp,
                     p.keymask, p.mask_words, N, p.num_rows, a.P_xa, a.xa_stats, a.ctx);
        } else if (two_frame) {
            launch_k(xt_train_fwd_kernel, dim3(N), dim3(256), 0, st, (const float*)a.cq, (const float*)a.Kp, (const float*)a.Vp, p.key_list,
                     p.key_cnt, p.mask_words * 32, p.num_rows, a.P_xa, a.ctx);
        } else {
            launch_k(xa_fwd_kernel, dim3(N), dim3(256), 0, st, (const float*)a.cq, (const float*)a.Kp, (const float*)a.Vp, p.match,
                     p.match_cnt, p.max_match, a.P_xa, a.ctx, N);
        }
        MV2D_CHECK_LAUNCH("train xa_fwd");
        TRY(linear_fwd(a.ctx, TC_, W.t[TL_CA_OUT_W], TC_, W.t[TL_CA_OUT_B], w.t1, TC_, N, TC_, TC_, false, st));
        TRY(ln_fwd(a.x1, w.t1, W.t[TL_LN1_G], W.t[TL_LN1_B], a.x2, a.xhat1, a.rstd1, N, false, st));
        // --- FFN (mmcv FFN: x + W2 relu(W1 x))
        TRY(linear_fwd(a.x2, TC_, W.t[TL_FFN_W1], TC_, W.t[TL_FFN_B1], a.hdn, TFF, N, TFF, TC_, true, st));
        TRY(linear_fwd(a.hdn, TFF, W.t[TL_FFN_W2], TFF, W.t[TL_FFN_B2], w.t1, TC_, N, TC_, TFF, false, st));
        TRY(ln_fwd(a.x2, w.t1, W.t[TL_LN2_G], W.t[TL_LN2_B], a.x3, a.xhat2, a.rstd2, N, false, st));
        // --- post_norm on the intermediate, then the branches (cross_attention_head.py:216-242)
        TRY(ln_fwd(a.x3, nullptr, post_g, post_b, a.inter, a.xhatp, a.rstdp, N, false, st));
        TRY(linear_fwd(a.inter, TC_, W.t[TL_CLS_W0], TC_, W.t[TL_CLS_B0], w.t1, TC_, N, TC_, TC_, false, st));
        TRY(ln_fwd(w.t1, nullptr, W.t[TL_CLS_G0], W.t[TL_CLS_BE0], a.c0n, a.xhat_c0, a.rstd_c0, N, true, st));
        TRY(linear_fwd(a.c0n, TC_, W.t[TL_CLS_W1], TC_, W.t[TL_CLS_B1], w.t1, TC_, N, TC_, TC_, false, st));
        TRY(ln_fwd(w.t1, nullptr, W.t[TL_CLS_G1], W.t[TL_CLS_BE1], a.c1n, a.xhat_c1, a.rstd_c1, N, true, st));
        float* cls_l = p.cls_scores + (long long)l * N * TCODE;
        float* box_l = p.bbox_preds + (long long)l * N * TCODE;
        TRY(linear_fwd(a.c1n, TC_, W.t[TL_CLS_W2], TC_, W.t[TL_CLS_B2], cls_l, TCODE, N, TCODE, TC_, false, st));
        TRY(linear_fwd(a.inter, TC_, W.t[TL_REG_W0], TC_, W.t[TL_REG_B0], a.r0, TC_, N, TC_, TC_, true, st));
        TRY(linear_fwd(a.r0, TC_, W.t[TL_REG_W1], TC_, W.t[TL_REG_B1], a.r1, TC_, N, TC_, TC_, true, st));
        TRY(linear_fwd(a.r1, TC_, W.t[TL_REG_W2], TC_, W.t[TL_REG_B2], box_l, TCODE, N, TCODE, TC_, false, st));
        launch_k(reg_tail_fwd_kernel, dim3(cdiv(N, 128)), dim3(128), 0, st, box_l, p.ref, a.rsig, pc, N, two_frame ? p.vel_dt : 0.f, p.pad);
        MV2D_CHECK_LAUNCH("train reg_tail");
    }
    // Hungarian targets + loss values of every layer (row f3)
    Mv2dLossParams lp{};
    // the matching queries are rows pad .. pad + p.N of every layer's [N,10] block; the denoising rows come first
    lp.N = p.N; lp.G = p.G; lp.L = L; lp.num_classes = p.num_classes; lp.pad = p.pad; lp.neg_bbox_loss = p.neg_bbox_loss;
    lp.layer_stride = (long long)N * TCODE; lp.dn_layer_stride = (long long)N * TCODE;
    lp.cls_cost_weight = p.cls_cost_weight; lp.reg_cost_weight = p.reg_cost_weight; lp.cls_loss_weight = p.cls_loss_weight;
    lp.bbox_loss_weight = p.bbox_loss_weight; lp.focal_alpha = p.focal_alpha; lp.focal_gamma = p.focal_gamma; lp.dn_split = p.dn_split;
    for (int j = 0; j < TCODE; ++j) lp.code_weights[j] = p.code_weights[j];
    lp.cls_scores = p.cls_scores + (long long)p.pad * TCODE; lp.bbox_preds = p.bbox_preds + (long long)p.pad * TCODE;
    if (p.pad > 0) { lp.dn_cls = p.cls_scores; lp.dn_box = p.bbox_preds; lp.dn_labels = p.dn_labels; }
    lp.gt_boxes = p.gt_boxes; lp.gt_labels = p.gt_labels;
    lp.assigned = p.assigned; lp.losses = p.losses; lp.workspace = w.loss_ws; lp.workspace_bytes = w.loss_ws_bytes;
    lp.num_pos = p.num_pos;
    return run_loss(lp, st);
}

int run_train_backward(const Mv2dTrainParams& p, cudaStream_t st) {
    TRY(check_params(p));
    const bool two_frame = p.mode == 1;
    MV2D_CHECK_ARG(p.grads && p.d_ref && (two_frame ? (p.d_kin_map && p.d_mem_map) : (p.d_tok_kin && p.d_tok_mem)),
                   "train backward: null gradient output");
    const int N = p.pad + p.N, L = p.L, NK = two_frame ? p.num_rows : N * TTOK;
    const long long NC = (long long)N * TC_;
    const TrainWs w = train_layout(p.workspace, N, L, p.max_match, p.G, p.mode, p.num_rows, p.N);
    float* d_kin = two_frame ? p.d_kin_map : p.d_tok_kin;       // gradient of the key input rows / of the value input rows
    float* d_mem = two_frame ? p.d_mem_map : p.d_tok_mem;
    const float* kin = two_frame ? p.kin_map : p.tok_kin;
    const float* mem = two_frame ? p.mem_map : p.tok_mem;
    g_tc = w.tc;
    float* P = const_cast<float*>(p.params);
    float* G = p.grads;
    Range6 pc;
    for (int i = 0; i < 6; ++i) pc.v[i] = p.pc_range[i];
    cudaError_t e;
#define ZERO(ptr, count)                                                                     \
    if ((e = cudaMemsetAsync(ptr, 0, (size_t)(count) * sizeof(float), st)) != cudaSuccess) { \
        set_error("train backward: memset %s", cudaGetErrorString(e));                       \
        return (int)e;                                                                       \
    }
    ZERO(p.d_ref, N * 3);
    ZERO(d_kin, (long long)NK * TC_);
    ZERO(d_mem, (long long)NK * TC_);
    ZERO(w.dqpos, NC);
    ZERO(w.dx, NC);
#undef ZERO

    LossGradArgs lg{};
    lg.cls = p.cls_scores; lg.box = p.bbox_preds; lg.assigned = p.assigned; lg.gt_boxes = p.gt_boxes; lg.gt_labels = p.gt_labels;
    lg.dcls = w.dcls; lg.dbox = w.dbox; lg.N = p.N; lg.G = p.G; lg.num_classes = p.num_classes;
    lg.pad = p.pad; lg.dn_labels = p.dn_labels; lg.dn_split = p.dn_split; lg.dn_weight = p.denoise_weight; lg.neg_bbox_loss = p.neg_bbox_loss;
    lg.alpha = p.focal_alpha; lg.gamma = p.focal_gamma; lg.cls_lw = p.cls_loss_weight; lg.box_lw = p.bbox_loss_weight;
    for (int j = 0; j < TCODE; ++j) lg.code_w[j] = p.code_weights[j];
    for (int l = 0; l < MV2D_MAX_LAYERS; ++l) lg.stage_w[l] = p.stage_loss_weights[l];
    lg.bbox_avg_factor = p.bbox_avg_factor; lg.losses = p.losses;
    launch_k(loss_grad_kernel, dim3(L), dim3(256), 0, st, lg);
    MV2D_CHECK_LAUNCH("train loss_grad");
    if (!two_frame) {
        launch_k(xa_inverse_kernel, dim3(N), dim3(32), 0, st, p.match, p.match_cnt, p.max_match, N, w.inv_cnt, w.inv_list);
        MV2D_CHECK_LAUNCH("train xa_inverse");
    }

    float* g_post_g = G + global_off(TG_POST_G);
    float* g_post_b = G + global_off(TG_POST_B);
    const int LC = L * TC_;           // row stride of the K / V gradients of all layers side by side
    const int Mp = round32(NK);
    // tensor-core mode with enough RoI tokens: the K / V projection gradients of ALL layers as four GEMMs after the loop
    // (dW_k of every layer = [dKp_0 | .. | dKp_L-1]^T tok_kin, d tok_kin = [dKp_0 | .. ] [Wk_0 ; .. ]: K = L*256 instead
    // of L GEMMs + L accumulation passes), 24 launches instead of 16 per layer
    const bool kv_batched = tc_enabled() && NK >= 1024 && tc_shape_ok(LC, TC_, Mp) && tc_shape_ok(NK, TC_, LC) &&
                            (size_t)LC * Mp <= w.tc.at_cap && (size_t)TC_ * Mp <= w.tc.bt_cap && (size_t)LC * TC_ <= w.tc.wt_cap;
    for (int l = L - 1; l >= 0; --l) {
        const LayerAct& a = w.layer[l];
        const LayerPtr W = layer_ptrs(P, l);
        const LayerPtr D = layer_ptrs(G, l);
        float* dcls = w.dcls + (long long)l * N * TCODE;
        float* dbox = w.dbox + (long long)l * N * TCODE;
        // --- reg branch
        launch_k(reg_tail_bwd_kernel, dim3(cdiv(N, 128)), dim3(128), 0, st, dbox, p.ref, (const float*)a.rsig, p.d_ref, pc, N,
                 two_frame ? p.vel_dt : 0.f, p.pad);
        MV2D_CHECK_LAUNCH("train reg_tail_bwd");
        TRY(linear_wgrad(dbox, TCODE, a.r1, TC_, D.t[TL_REG_W2], TC_, N, TCODE, TC_, st, D.t[TL_REG_B2]));
        TRY(linear_dgrad(dbox, TCODE, W.t[TL_REG_W2], TC_, w.t1, TC_, N, TCODE, TC_, a.r1, TC_, false, st));
        TRY(linear_wgrad(w.t1, TC_, a.r0, TC_, D.t[TL_REG_W1], TC_, N, TC_, TC_, st, D.t[TL_REG_B1]));
        TRY(linear_dgrad(w.t1, TC_, W.t[TL_REG_W1], TC_, w.t2, TC_, N, TC_, TC_, a.r0, TC_, false, st));
        TRY(linear_wgrad(w.t2, TC_, a.inter, TC_, D.t[TL_REG_W0], TC_, N, TC_, TC_, st, D.t[TL_REG_B0]));
        TRY(linear_dgrad(w.t2, TC_, W.t[TL_REG_W0], TC_, w.dinter, TC_, N, TC_, TC_, nullptr, 0, false, st));
        // --- cls branch
        TRY(linear_wgrad(dcls, TCODE, a.c1n, TC_, D.t[TL_CLS_W2], TC_, N, TCODE, TC_, st, D.t[TL_CLS_B2]));
        TRY(linear_dgrad(dcls, TCODE, W.t[TL_CLS_W2], TC_, w.t1, TC_, N, TCODE, TC_, nullptr, 0, false, st));
        TRY(ln_bwd(w.t1, a.c1n, a.xhat_c1, a.rstd_c1, W.t[TL_CLS_G1], w.t2, D.t[TL_CLS_G1], D.t[TL_CLS_BE1], N, false, st));
        TRY(linear_wgrad(w.t2, TC_, a.c0n, TC_, D.t[TL_CLS_W1], TC_, N, TC_, TC_, st, D.t[TL_CLS_B1]));
        TRY(linear_dgrad(w.t2, TC_, W.t[TL_CLS_W1], TC_, w.t1, TC_, N, TC_, TC_, nullptr, 0, false, st));
        TRY(ln_bwd(w.t1, a.c0n, a.xhat_c0, a.rstd_c0, W.t[TL_CLS_G0], w.t2, D.t[TL_CLS_G0], D.t[TL_CLS_BE0], N, false, st));
        TRY(linear_wgrad(w.t2, TC_, a.inter, TC_, D.t[TL_CLS_W0], TC_, N, TC_, TC_, st, D.t[TL_CLS_B0]));
        TRY(linear_dgrad(w.t2, TC_, W.t[TL_CLS_W0], TC_, w.dinter, TC_, N, TC_, TC_, nullptr, 0, true, st));
        // --- post_norm: dx (gradient of this layer's output) += LN'(dinter)
        TRY(ln_bwd(w.dinter, nullptr, a.xhatp, a.rstdp, P + global_off(TG_POST_G), w.dx, g_post_g, g_post_b, N, true, st));
        // --- norms.2 and the FFN
        TRY(ln_bwd(w.dx, nullptr, a.xhat2, a.rstd2, W.t[TL_LN2_G], w.t1, D.t[TL_LN2_G], D.t[TL_LN2_B], N, false, st));
        TRY(linear_wgrad(w.t1, TC_, a.hdn, TFF, D.t[TL_FFN_W2], TFF, N, TC_, TFF, st, D.t[TL_FFN_B2]));
        TRY(linear_dgrad(w.t1, TC_, W.t[TL_FFN_W2], TFF, w.dhdn, TFF, N, TC_, TFF, a.hdn, TFF, false, st));
        TRY(linear_wgrad(w.dhdn, TFF, a.x2, TC_, D.t[TL_FFN_W1], TC_, N, TFF, TC_, st, D.t[TL_FFN_B1]));
        TRY(linear_dgrad(w.dhdn, TFF, W.t[TL_FFN_W1], TC_, w.t1, TC_, N, TFF, TC_, nullptr, 0, true, st));   // t1 = d x2
        // --- norms.1 and the cross-attention
        TRY(ln_bwd(w.t1, nullptr, a.xhat1, a.rstd1, W.t[TL_LN1_G], w.t2, D.t[TL_LN1_G], D.t[TL_LN1_B], N, false, st));
        TRY(linear_wgrad(w.t2, TC_, a.ctx, TC_, D.t[TL_CA_OUT_W], TC_, N, TC_, TC_, st, D.t[TL_CA_OUT_B]));
        TRY(linear_dgrad(w.t2, TC_, W.t[TL_CA_OUT_W], TC_, w.t3, TC_, N, TC_, TC_, nullptr, 0, false, st));   // t3 = d ctx
        if (two_frame && xt2_enabled()) {
            launch_k(xt2_bwd_dq_kernel, dim3(cdiv(N, XT2_QB), TH), dim3(256), 0, st, (const float*)a.Kp, (const float*)a.Vp, (const float*)a.P_xa,
                     (const float*)a.xa_stats, (const float*)w.t3, (const float*)a.ctx, p.keymask, p.mask_words, N, p.num_rows, w.dS_xa, w.dcq);
            MV2D_CHECK_LAUNCH("train xt2_bwd_dq");
            launch_k(xt2_bwd_dkv_kernel, dim3(cdiv(NK, XT2_KT), TH), dim3(128), 0, st, (const float*)a.cq, (const float*)w.t3, (const float*)a.P_xa,
                     (const float*)a.xa_stats, (const float*)w.dS_xa, p.keymask, p.mask_words, N, p.num_rows, w.dKp + l * TC_, w.dVp + l * TC_, LC);
            MV2D_CHECK_LAUNCH("train xt2_bwd_dkv");
        } else if (two_frame) {
            launch_k(xt_train_bwd_dq_kernel, dim3(N), dim3(256), 0, st, (const float*)a.Kp, (const float*)a.Vp, (const float*)a.P_xa,
                     (const float*)w.t3, p.key_list, p.key_cnt, p.mask_words * 32, p.num_rows, w.dS_xa, w.dcq);
            MV2D_CHECK_LAUNCH("train xt_bwd_dq");
            launch_k(xt_train_bwd_dkv_kernel, dim3(NK), dim3(256), 0, st, (const float*)a.cq, (const float*)w.t3, (const float*)a.P_xa,
                     (const float*)w.dS_xa, p.keymask, p.mask_words, N, p.num_rows, w.dKp + l * TC_, w.dVp + l * TC_, LC);
            MV2D_CHECK_LAUNCH("train xt_bwd_dkv");
        } else {
        launch_k(xa_bwd_dq_kernel, dim3(N), dim3(256), 0, st, (const float*)a.Kp, (const float*)a.Vp, (const float*)a.P_xa,
                 (const float*)w.t3, p.match, p.match_cnt, p.max_match, w.dS_xa, w.dcq, N);
        MV2D_CHECK_LAUNCH("train xa_bwd_dq");
        launch_k(xa_bwd_dkv_kernel, dim3(NK), dim3(256), 0, st, (const float*)a.cq, (const float*)w.t3, (const float*)a.P_xa,
                 (const float*)w.dS_xa, (const int*)w.inv_cnt, (const int*)w.inv_list, p.max_match, w.dKp + l * TC_, w.dVp + l * TC_, LC, N);
        MV2D_CHECK_LAUNCH("train xa_bwd_dkv");
        }
        TRY(linear_wgrad(w.dcq, TC_, a.xq1, TC_, D.t[TL_CA_IN_W], TC_, N, TC_, TC_, st, D.t[TL_CA_IN_B]));
        TRY(linear_dgrad(w.dcq, TC_, W.t[TL_CA_IN_W], TC_, w.t3, TC_, N, TC_, TC_, nullptr, 0, false, st));   // t3 = d (x1 + qpos)
        TRY(add(w.t2, w.t2, w.t3, NC, st));          // t2 = d x1
        TRY(add(w.dqpos, w.dqpos, w.t3, NC, st));
        if (!kv_batched) {     // per layer; otherwise all layers' K / V projection gradients in four GEMMs after the loop
            TRY(linear_wgrad(w.dKp + l * TC_, LC, kin, TC_, D.t[TL_CA_IN_W] + 256 * TC_, TC_, NK, TC_, TC_, st, D.t[TL_CA_IN_B] + 256));
            TRY(linear_dgrad(w.dKp + l * TC_, LC, W.t[TL_CA_IN_W] + 256 * TC_, TC_, d_kin, TC_, NK, TC_, TC_, nullptr, 0, true, st));
            TRY(linear_wgrad(w.dVp + l * TC_, LC, mem, TC_, D.t[TL_CA_IN_W] + 512 * TC_, TC_, NK, TC_, TC_, st, D.t[TL_CA_IN_B] + 512));
            TRY(linear_dgrad(w.dVp + l * TC_, LC, W.t[TL_CA_IN_W] + 512 * TC_, TC_, d_mem, TC_, NK, TC_, TC_, nullptr, 0, true, st));
        }
        // --- norms.0 and the self-attention
        TRY(ln_bwd(w.t2, nullptr, a.xhat0, a.rstd0, W.t[TL_LN0_G], w.t1, D.t[TL_LN0_G], D.t[TL_LN0_B], N, false, st));   // t1 = d (x_in + sa)
        TRY(linear_wgrad(w.t1, TC_, a.attn_o, TC_, D.t[TL_SA_OUT_W], TC_, N, TC_, TC_, st, D.t[TL_SA_OUT_B]));
        TRY(linear_dgrad(w.t1, TC_, W.t[TL_SA_OUT_W], TC_, w.t3, TC_, N, TC_, TC_, nullptr, 0, false, st));   // t3 = d attn_o
        if (sa_use_smem(N)) {
            TRY(sa_set_attr());
            launch_k(sa_bwd_dq_smem_kernel, dim3(cdiv(N, SA_QB), TH), dim3(256), sa_smem_bytes(N), st, (const float*)a.qkv,
                     (const float*)a.P_sa, (const float*)w.t3, w.dS_sa, w.dqkv, N);
            MV2D_CHECK_LAUNCH("train sa_bwd_dq");
            launch_k(sa_bwd_dkv_smem_kernel, dim3(cdiv(N, SA_QB), TH), dim3(256), sa_smem_bytes(N), st, (const float*)a.qkv,
                     (const float*)a.P_sa, (const float*)w.dS_sa, (const float*)w.t3, w.dqkv, N);
            MV2D_CHECK_LAUNCH("train sa_bwd_dkv");
        } else {
            launch_k(sa_bwd_dq_kernel, dim3(N), dim3(256), 0, st, (const float*)a.qkv, (const float*)a.P_sa, (const float*)w.t3, w.dS_sa, w.dqkv, N);
            MV2D_CHECK_LAUNCH("train sa_bwd_dq");
            launch_k(sa_bwd_dkv_kernel, dim3(N), dim3(256), 0, st, (const float*)a.qkv, (const float*)a.P_sa, (const float*)w.dS_sa,
                     (const float*)w.t3, w.dqkv, N);
            MV2D_CHECK_LAUNCH("train sa_bwd_dkv");
        }
        TRY(linear_wgrad(w.dqkv, 768, a.xq, TC_, D.t[TL_SA_IN_W], TC_, N, 512, TC_, st, D.t[TL_SA_IN_B]));
        TRY(linear_wgrad(w.dqkv + 512, 768, a.x_in, TC_, D.t[TL_SA_IN_W] + 512 * TC_, TC_, N, TC_, TC_, st, D.t[TL_SA_IN_B] + 512));
        TRY(linear_dgrad(w.dqkv, 768, W.t[TL_SA_IN_W], TC_, w.t3, TC_, N, 512, TC_, nullptr, 0, false, st));   // t3 = d (x_in + qpos)
        TRY(add(w.dqpos, w.dqpos, w.t3, NC, st));
        TRY(add(w.dx, w.t1, w.t3, NC, st));           // dx = gradient of the previous layer's output
        TRY(linear_dgrad(w.dqkv + 512, 768, W.t[TL_SA_IN_W] + 512 * TC_, TC_, w.dx, TC_, N, TC_, TC_, nullptr, 0, true, st));
    }
    if (kv_batched) {
        const long long lstride = layer_block_floats();
        const int tiles = cdiv(LC, 128) * (TC_ / (LC <= 512 ? 64 : 128)), nkb = Mp / 32;
        int nsplit = 1;
        for (int d = 1; d <= 48 && d <= nkb; ++d)
            if (nkb % d == 0 && nkb / d >= 4 && (size_t)d * LC * TC_ <= w.tc.part_cap) { nsplit = d; if (tiles * d >= 148) break; }
        for (int kv = 0; kv < 2; ++kv) {
            const float* dYall = kv == 0 ? w.dKp : w.dVp;             // [NK, L*256]
            const float* X = kv == 0 ? kin : mem;                     // [NK, 256]
            float* dX = kv == 0 ? d_kin : d_mem;
            const int wrow = kv == 0 ? 256 : 512;                     // rows of in_proj: q | k | v
            // weight + bias gradients of every layer
            TRY(transpose_pad(dYall, LC, NK, LC, w.tc.at, Mp, st));
            TRY(transpose_pad(X, TC_, NK, TC_, w.tc.bt, Mp, st));
            TRY(tc_gemm(w.tc.at, Mp, w.tc.bt, Mp, nullptr, w.tc.part, TC_, LC, TC_, Mp, false, nsplit, (long long)LC * TC_, st));
            launch_k(wgrad_fold_kernel, dim3(ew_grid_n((long long)LC * TC_)), dim3(256), 0, st, (const float*)w.tc.part, nsplit,
                     (long long)LC * TC_, LC, TC_, 0, G + layer_off(0, TL_CA_IN_W) + (long long)wrow * TC_, TC_, TC_, lstride);
            MV2D_CHECK_LAUNCH("train kv wgrad_fold");
            launch_k(rowsum_kernel, dim3(LC), dim3(256), 0, st, (const float*)w.tc.at, Mp, LC, G + layer_off(0, TL_CA_IN_B) + wrow, TC_, lstride);
            MV2D_CHECK_LAUNCH("train kv rowsum");
            // input gradient: d tok = [dY_0 | .. | dY_L-1] [W_0 ; .. ; W_L-1]  (d_tok_* was zeroed above; written once here)
            for (int l = 0; l < L; ++l)
                TRY(transpose_pad(P + layer_off(l, TL_CA_IN_W) + (long long)wrow * TC_, TC_, TC_, TC_, w.tc.wt + l * TC_, TC_, st, LC));
            TRY(tc_gemm(dYall, LC, w.tc.wt, LC, nullptr, dX, TC_, NK, TC_, LC, false, 1, 0, st));
        }
    }
    // --- query embedding MLP and the sin / cos features
    TRY(linear_wgrad(w.dqpos, TC_, w.h0, TC_, G + global_off(TG_QE2_W), TC_, N, TC_, TC_, st, G + global_off(TG_QE2_B)));
    TRY(linear_dgrad(w.dqpos, TC_, P + global_off(TG_QE2_W), TC_, w.t1, TC_, N, TC_, TC_, w.h0, TC_, false, st));
    TRY(linear_wgrad(w.t1, TC_, w.posemb, TPE, G + global_off(TG_QE0_W), TPE, N, TC_, TPE, st, G + global_off(TG_QE0_B)));
    TRY(linear_dgrad(w.t1, TC_, P + global_off(TG_QE0_W), TPE, w.dposemb, TPE, N, TC_, TPE, nullptr, 0, false, st));
    launch_k(posemb_bwd_kernel, dim3(N), dim3(96), 0, st, (const float*)w.dposemb, p.ref, p.dim_t, p.d_ref, N);
    MV2D_CHECK_LAUNCH("train posemb_bwd");
    return 0;
}


size_t front_train_workspace_bytes(int N, int V, int h, int w) { return front_layout(nullptr, N, V * h * w).total_bytes; }

int run_front_train_forward(const Mv2dFrontTrainParams& p, cudaStream_t st) {
    TRY(check_front(p));
    const int N = p.N, P = p.V * p.h * p.w, NK = N * TTOK;
    const FrontWs w = front_layout(p.workspace, N, P);
    g_tc = w.tc;
    float* Wt = const_cast<float*>(p.params);
    auto W = [&](int t) { return Wt + front_off(p.L, t); };
    Range6 pc;
    for (int i = 0; i < 6; ++i) pc.v[i] = p.pc_range[i];
    const float scale = 1.0f / (float)p.stride;
    // --- PE.forward: frustum coordinates and sine features are parameter-free inputs (un-rounded fp32 here)
    TRY(run_pe_train_inputs(p.V, p.h, p.w, p.depth_num, p.pad_h, p.pad_w, p.stride, p.depth_start, p.position_range, p.img2lidar,
                            p.not_mask, p.dim_t, w.coords, w.sine, st));
    TRY(linear_fwd(w.coords, 192, W(TF_POS0_W), 192, W(TF_POS0_B), w.hp, 1024, P, 1024, 192, true, st));
    TRY(linear_fwd(w.hp, 1024, W(TF_POS2_W), 1024, W(TF_POS2_B), w.x, TC_, P, TC_, 1024, false, st));
    TRY(linear_fwd(w.sine, 384, W(TF_ADAPT0_W), 384, W(TF_ADAPT0_B), w.hs, 1024, P, 1024, 384, true, st));
    TRY(linear_fwd(w.hs, 1024, W(TF_ADAPT2_W), 1024, W(TF_ADAPT2_B), w.sb, TC_, P, TC_, 1024, false, st));
    TRY(linear_fwd(p.feat, TC_, W(TF_SE_R_W), TC_, W(TF_SE_R_B), w.g1, TC_, P, TC_, TC_, true, st));
    TRY(linear_fwd(w.g1, TC_, W(TF_SE_E_W), TC_, W(TF_SE_E_B), w.gate, TC_, P, TC_, TC_, false, st));
    launch_k(pe_gate_fwd_kernel, dim3(ew_grid_n((long long)P * TC_)), dim3(256), 0, st, (const float*)w.x, w.gate, (const float*)w.sb, w.pe,
             (long long)P * TC_);
    MV2D_CHECK_LAUNCH("front pe_gate");
    if (p.pe_out) {
        cudaError_t e = cudaMemcpyAsync(p.pe_out, w.pe, (size_t)P * TC_ * sizeof(float), cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) { set_error("front train: memcpy %s", cudaGetErrorString(e)); return (int)e; }
    }
    if (p.kin_out) TRY(add(p.kin_out, p.feat, w.pe, (long long)P * TC_, st));       // key input of the two-frame head
    // --- RoIAlign of the feature and of the position embedding (torch.cat + SingleRoIExtractor + split)
    launch_k(roi_align_fwd_kernel, dim3(TTOK, N), dim3(64), 0, st, p.rois, p.feat, p.h, p.w, scale, (const float*)nullptr, p.tok_mem);
    MV2D_CHECK_LAUNCH("front roi_align(feat)");
    launch_k(roi_align_fwd_kernel, dim3(TTOK, N), dim3(64), 0, st, p.rois, (const float*)w.pe, p.h, p.w, scale, (const float*)p.tok_mem, p.tok_kin);
    MV2D_CHECK_LAUNCH("front roi_align(pe)");
    // --- query generator
    launch_k(front_params_kernel, dim3(cdiv(N, 64)), dim3(64), 0, st, p.rois, p.roi_intrinsics, p.extrinsics, N, p.intrins_feat_scale,
             w.cat, w.m_roi);
    MV2D_CHECK_LAUNCH("front params");
    launch_k(im2col_kernel, dim3(ew_grid_n((long long)NK * 9 * 64)), dim3(256), 0, st, (const float*)p.tok_mem, w.col, N);
    MV2D_CHECK_LAUNCH("front im2col");
    TRY(linear_fwd(w.col, 9 * TC_, W(TF_CONV_W), 9 * TC_, W(TF_CONV_B), w.y, TC_, NK, TC_, 9 * TC_, true, st));
    launch_k(pool49_fwd_kernel, dim3(cdiv(N * TC_, 256)), dim3(256), 0, st, (const float*)w.y, w.pool, N);
    MV2D_CHECK_LAUNCH("front pool49");
    TRY(linear_fwd(w.pool, TC_, W(TF_FC_W), TC_, W(TF_FC_B), w.cat, 1040, N, 1024, TC_, true, st, SG_CLAMP5E3));
    TRY(linear_fwd(w.cat, 1040, W(TF_ENC0_W), 1040, W(TF_ENC0_B), w.e0, 512, N, 512, 1040, true, st));
    TRY(linear_fwd(w.e0, 512, W(TF_ENC2_W), 512, W(TF_ENC2_B), w.e2, TC_, N, TC_, 512, true, st));
    TRY(linear_fwd(w.e2, TC_, W(TF_CENTER_W), TC_, W(TF_CENTER_B), w.c, 3, N, 3, TC_, false, st));
    launch_k(center_fwd_kernel, dim3(cdiv(N, 128)), dim3(128), 0, st, (const float*)w.c, (const float*)w.m_roi, pc, p.ref, N);
    MV2D_CHECK_LAUNCH("front center");
    return 0;
}

int run_front_train_backward(const Mv2dFrontTrainParams& p, cudaStream_t st) {
    TRY(check_front(p));
    MV2D_CHECK_ARG(p.grads && p.d_ref && p.d_tok_kin && p.d_tok_mem && p.d_feat, "front train backward: null gradient pointer");
    const int N = p.N, P = p.V * p.h * p.w, NK = N * TTOK;
    const long long PC = (long long)P * TC_;
    const FrontWs w = front_layout(p.workspace, N, P);
    g_tc = w.tc;
    float* Wt = const_cast<float*>(p.params);
    auto W = [&](int t) { return Wt + front_off(p.L, t); };
    auto D = [&](int t) { return p.grads + front_off(p.L, t); };
    Range6 pc;
    for (int i = 0; i < 6; ++i) pc.v[i] = p.pc_range[i];
    const float scale = 1.0f / (float)p.stride;
    cudaError_t e;
    if ((e = cudaMemsetAsync(p.d_feat, 0, (size_t)PC * sizeof(float), st)) != cudaSuccess ||
        (e = cudaMemsetAsync(w.dpe_map, 0, (size_t)PC * sizeof(float), st)) != cudaSuccess) {
        set_error("front train backward: memset %s", cudaGetErrorString(e));
        return (int)e;
    }
    // --- reference points -> fc_center -> FC chain -> avg-pool -> conv
    launch_k(center_bwd_kernel, dim3(cdiv(N, 128)), dim3(128), 0, st, (const float*)w.c, (const float*)w.m_roi, pc, p.d_ref, w.dc, N);
    MV2D_CHECK_LAUNCH("front center_bwd");
    TRY(linear_wgrad(w.dc, 3, w.e2, TC_, D(TF_CENTER_W), TC_, N, 3, TC_, st, D(TF_CENTER_B)));
    TRY(linear_dgrad(w.dc, 3, W(TF_CENTER_W), TC_, w.de2, TC_, N, 3, TC_, w.e2, TC_, false, st));
    TRY(linear_wgrad(w.de2, TC_, w.e0, 512, D(TF_ENC2_W), 512, N, TC_, 512, st, D(TF_ENC2_B)));
    TRY(linear_dgrad(w.de2, TC_, W(TF_ENC2_W), 512, w.de0, 512, N, TC_, 512, w.e0, 512, false, st));
    TRY(linear_wgrad(w.de0, 512, w.cat, 1040, D(TF_ENC0_W), 1040, N, 512, 1040, st, D(TF_ENC0_B)));
    // only the 1024 FC columns of the concatenation carry a gradient (the intrinsics feature is an input)
    TRY(linear_dgrad(w.de0, 512, W(TF_ENC0_W), 1040, w.df1, 1024, N, 512, 1024, w.cat, 1040, false, st, SG_MASK_LT5E3));
    TRY(linear_wgrad(w.df1, 1024, w.pool, TC_, D(TF_FC_W), TC_, N, 1024, TC_, st, D(TF_FC_B)));
    TRY(linear_dgrad(w.df1, 1024, W(TF_FC_W), TC_, w.dpool, TC_, N, 1024, TC_, nullptr, 0, false, st));
    launch_k(pool49_bwd_kernel, dim3(ew_grid_n((long long)NK * TC_)), dim3(256), 0, st, (const float*)w.y, (const float*)w.dpool, w.dy, N);
    MV2D_CHECK_LAUNCH("front pool49_bwd");
    TRY(linear_wgrad(w.dy, TC_, w.col, 9 * TC_, D(TF_CONV_W), 9 * TC_, NK, TC_, 9 * TC_, st, D(TF_CONV_B)));
    TRY(linear_dgrad(w.dy, TC_, W(TF_CONV_W), 9 * TC_, w.dcol, 9 * TC_, NK, TC_, 9 * TC_, nullptr, 0, false, st));
    // d tok_mem = conv path + value path + key path (tok_kin = tok_mem + RoIAlign(pe)); d RoIAlign(pe) = d tok_kin
    launch_k(col2im_kernel, dim3(ew_grid_n((long long)NK * 64)), dim3(256), 0, st, (const float*)w.dcol, p.d_tok_mem, p.d_tok_kin, w.dtok, N);
    MV2D_CHECK_LAUNCH("front col2im");
    launch_k(roi_align_bwd_kernel, dim3(TTOK, N), dim3(64), 0, st, p.rois, (const float*)w.dtok, p.h, p.w, scale, p.d_feat);
    MV2D_CHECK_LAUNCH("front roi_align_bwd(feat)");
    launch_k(roi_align_bwd_kernel, dim3(TTOK, N), dim3(64), 0, st, p.rois, p.d_tok_kin, p.h, p.w, scale, w.dpe_map);
    MV2D_CHECK_LAUNCH("front roi_align_bwd(pe)");
    if (p.d_pe_extra) TRY(add(w.dpe_map, w.dpe_map, p.d_pe_extra, PC, st));        // two-frame head: keys = the whole feat + pe map
    if (p.d_feat_extra) TRY(add(p.d_feat, p.d_feat, p.d_feat_extra, PC, st));
    if (p.d_feat_extra2) TRY(add(p.d_feat, p.d_feat, p.d_feat_extra2, PC, st));
    // --- PE: pe = x * gate + sb
    launch_k(pe_gate_bwd_kernel, dim3(ew_grid_n(PC)), dim3(256), 0, st, (const float*)w.dpe_map, (const float*)w.x, (const float*)w.gate, w.dx, w.dg2, PC);
    MV2D_CHECK_LAUNCH("front pe_gate_bwd");
    // sine branch (adapt_pos3d): d sb = dpe
    TRY(linear_wgrad(w.dpe_map, TC_, w.hs, 1024, D(TF_ADAPT2_W), 1024, P, TC_, 1024, st, D(TF_ADAPT2_B)));
    TRY(linear_dgrad(w.dpe_map, TC_, W(TF_ADAPT2_W), 1024, w.dh, 1024, P, TC_, 1024, w.hs, 1024, false, st));
    TRY(linear_wgrad(w.dh, 1024, w.sine, 384, D(TF_ADAPT0_W), 384, P, 1024, 384, st, D(TF_ADAPT0_B)));
    // position MLP (position_encoder)
    TRY(linear_wgrad(w.dx, TC_, w.hp, 1024, D(TF_POS2_W), 1024, P, TC_, 1024, st, D(TF_POS2_B)));
    TRY(linear_dgrad(w.dx, TC_, W(TF_POS2_W), 1024, w.dh, 1024, P, TC_, 1024, w.hp, 1024, false, st));
    TRY(linear_wgrad(w.dh, 1024, w.coords, 192, D(TF_POS0_W), 192, P, 1024, 192, st, D(TF_POS0_B)));
    // SE gate (fpe): g2 = We relu(Wr feat + br) + be
    TRY(linear_wgrad(w.dg2, TC_, w.g1, TC_, D(TF_SE_E_W), TC_, P, TC_, TC_, st, D(TF_SE_E_B)));
    TRY(linear_dgrad(w.dg2, TC_, W(TF_SE_E_W), TC_, w.dx, TC_, P, TC_, TC_, w.g1, TC_, false, st));      // dx reused: d g1
    TRY(linear_wgrad(w.dx, TC_, p.feat, TC_, D(TF_SE_R_W), TC_, P, TC_, TC_, st, D(TF_SE_R_B)));
    TRY(linear_dgrad(w.dx, TC_, W(TF_SE_R_W), TC_, p.d_feat, TC_, P, TC_, TC_, nullptr, 0, true, st));
    return 0;
}

// ---- fused AdamW over the flat buffers (torch.optim.AdamW semantics, the exp configs' optimizer):
// p -= lr * wd * p;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr / (1 - b1^t) * m / (sqrt(v / (1 - b2^t)) + eps)
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                                                    float wd, float bc1, float bc2, float grad_scale) {
    pdl_wait();
    pdl_trigger();
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float gi = g[i] * grad_scale;
        float pi = p[i];
        pi -= lr * wd * pi;
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
        p[i] = pi - (lr / bc1) * (mi / denom);
    }
}

int run_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps, float wd,
              int step, float grad_scale, cudaStream_t st) {
    MV2D_CHECK_ARG(p && g && m && v && n >= 0 && step >= 1, "adamw: bad arguments");
    if (n == 0) return 0;
    const float bc1 = 1.f - powf(b1, (float)step), bc2 = 1.f - powf(b2, (float)step);
    const long long want = (n + 255) / 256;
    const int grid = (int)(want < 148 * 16 ? want : 148 * 16);
    launch_k(adamw_kernel, dim3(grid), dim3(256), 0, st, p, g, m, v, n, lr, b1, b2, eps, wd, bc1, bc2, grad_scale);
    MV2D_CHECK_LAUNCH("adamw");
    return 0;
}

}  // namespace mv2d
