// Row a20: denoising queries of the training-mode forward (include/mv2d_b200.h, Mv2dDnParams).
// Reference: MV2DSHead.prepare_for_dn (roi_heads/mv2d_s_head.py:39-120), the cross-mask extension of both heads
// (mv2d_s_head.py:159-172, mv2d_t_head.py:79-98) and query_embedding(pos2posemb3d(.)) over the padded
// reference points (cross_attention_head.py:199-206).  Compiled with -fmad=false: the noised centres are
// plain fp32 mul/add/div in the reference.
#include "common.cuh"
#include "keylist.cuh"
#include "gemm_simt.cuh"
#include "mv2d_internal.h"

namespace mv2d {

// One CTA per output query row i: reference point, label (denoising rows), one row of the self-attention
// mask, and the sine embedding of the reference point (input of the query-embedding MLP).
__global__ void __launch_bounds__(128)
dn_rows_kernel(Mv2dDnParams p, int pad, int T, float* __restrict__ posemb) {
    pdl_wait();
    pdl_trigger();
    const int i = blockIdx.x, t = threadIdx.x;
    __shared__ float r[3];
    if (t == 0) {
        if (i < pad) {
            const int g = i % p.G;
            const float* b = p.gt_boxes + g * 9;
            float sq = 0.f;
            for (int k = 0; k < 3; ++k) {
                const float rp = p.rand[i * 3 + k] * 2.0f - 1.0f;
                const float diff = b[3 + k] / 2.0f + p.noise_trans;
                float c = b[k];
                if (p.noise_scale > 0.f) c = c + (rp * diff) * p.noise_scale;
                c = (c - p.pc_range[k]) / (p.pc_range[k + 3] - p.pc_range[k]);
                r[k] = fminf(fmaxf(c, p.eps), 1.0f - p.eps);
                sq += rp * rp;
            }
            p.dn_labels[i] = (p.noise_scale > 0.f && sqrtf(sq) > p.split) ? p.num_classes : p.gt_labels[g];
        } else {
            for (int k = 0; k < 3; ++k) r[k] = p.ref[(i - pad) * 3 + k];
        }
        for (int k = 0; k < 3; ++k) p.ref_all[i * 3 + k] = r[k];
    }
    __syncthreads();
    // self-attention mask row (mv2d_s_head.py:93-104)
    uint8_t* mrow = p.attn_mask + (long long)i * T;
    const int grp = i < pad ? i / p.G : -1;
    for (int j = t; j < T; j += blockDim.x) mrow[j] = (j < pad && (grp < 0 || j / p.G != grp)) ? 1 : 0;
    // pos2posemb3d: cat(emb(y), emb(x), emb(z)), interleaved sin/cos (utils/pe.py:21-33)
    for (int idx = t; idx < 384; idx += blockDim.x) {
        const int part = idx >> 7, k = idx & 127;
        const float pos = (part == 0 ? r[1] : (part == 1 ? r[0] : r[2])) * 6.283185307179586f;
        const float a = pos / __ldg(p.dim_t + k);
        posemb[(long long)i * 384 + idx] = (k & 1) ? cosf(a) : sinf(a);
    }
}

// S head: match lists of all T rows.
__global__ void dn_match_kernel(Mv2dDnParams p, int pad) {
    pdl_wait();
    pdl_trigger();
    const int i = blockIdx.x;
    int* out = p.match_all + (long long)i * p.max_match_all;
    if (i < pad) {
        for (int j = threadIdx.x; j < p.N; j += blockDim.x) out[j] = j;
        if (threadIdx.x == 0) p.match_cnt_all[i] = p.N;
    } else {
        const int n = i - pad, cnt = p.match_cnt[n];
        for (int j = threadIdx.x; j < cnt; j += blockDim.x) out[j] = p.match[(long long)n * p.max_match + j];
        if (threadIdx.x == 0) p.match_cnt_all[i] = cnt;
    }
}

// T head, step 1: the union of all per-query key masks (what the denoising rows attend to).
__global__ void dn_union_kernel(Mv2dDnParams p, uint32_t* __restrict__ uni) {
    pdl_wait();
    pdl_trigger();
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= p.mask_words) return;
    uint32_t b = 0u;
    for (int n = 0; n < p.N; ++n) b |= p.keymask[(long long)n * p.mask_words + w];
    if (w == 0 && p.train_unmask)
        for (int n = 0; n < p.N; ++n) if (p.key_cnt[n] == 0) b |= 1u;
    uni[w] = b;
}

// T head, step 2: one CTA per output row -- mask row, key count and the ordered key list.
__global__ void __launch_bounds__(256)
dn_keys_kernel(Mv2dDnParams p, int pad, const uint32_t* __restrict__ uni) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ uint32_t bits[];
    __shared__ int total;
    __shared__ int grp_cnt[128];
    const int i = blockIdx.x, t = threadIdx.x, words = p.mask_words;
    const uint32_t* src = i < pad ? uni : p.keymask + (long long)(i - pad) * words;
    const bool unmask = i >= pad && p.train_unmask && p.key_cnt[i - pad] == 0;
    if (t == 0) total = 0;
    __syncthreads();
    int local = 0;
    for (int w = t; w < words; w += blockDim.x) {
        uint32_t b = src[w];
        if (unmask && w == 0) b |= 1u;
        bits[w] = b;
        p.keymask_all[(long long)i * words + w] = b;
        local += __popc(b);
    }
    atomicAdd(&total, local);
    __syncthreads();
    if (t == 0) p.key_cnt_all[i] = total;
    compact_key_bits(bits, words, grp_cnt, p.key_list_all + (long long)i * words * 32);
}

// pos2posemb3d (utils/pe.py:21-33) of N reference points: [N,384] = (emb(y) | emb(x) | emb(z)), interleaved sin / cos.
// Same arithmetic as qg_tail_kernel (roi.cu) and dn_rows_kernel above.
__global__ void __launch_bounds__(128) posemb3d_kernel(const float* __restrict__ ref, const float* __restrict__ dim_t, int N,
                                                       float* __restrict__ posemb) {
    pdl_wait();
    pdl_trigger();
    const int n = blockIdx.x;
    if (n >= N) return;
    const float p[3] = {ref[n * 3 + 0], ref[n * 3 + 1], ref[n * 3 + 2]};
    for (int idx = threadIdx.x; idx < 384; idx += blockDim.x) {
        const int part = idx >> 7, i = idx & 127;
        const float pos = (part == 0 ? p[1] : (part == 1 ? p[0] : p[2])) * 6.283185307179586f;
        const float a = pos / __ldg(dim_t + i);
        posemb[(long long)n * 384 + idx] = (i & 1) ? cosf(a) : sinf(a);
    }
}

// CrossAttentionBoxHead.position_embedding (cross_attention_head.py:199-200): query_embedding(pos2posemb3d(ref)).
// workspace: N * (384 + 256) floats.
int run_query_embedding(const float* ref, int N, const float* w_qe0, const float* b_qe0, const float* w_qe2, const float* b_qe2,
                        const float* dim_t, float* query_pos, float* workspace, cudaStream_t st) {
    if (N == 0) return 0;
    float* pemb = workspace;
    float* qh = workspace + (size_t)N * 384;
    launch_k(posemb3d_kernel, dim3(N), dim3(128), 0, st, ref, dim_t, N, pemb);
    MV2D_CHECK_LAUNCH("posemb3d");
    GemmArgs g{};
    g.A = pemb; g.lda = 384; g.W = w_qe0; g.ldw = 384; g.bias = b_qe0; g.C = qh; g.ldc = MV2D_C;
    g.M = N; g.N = MV2D_C; g.K = 384; g.batch = 1; g.nsplit = 1; g.flags = GEMM_RELU;
    int rc;
    if ((rc = launch_gemm_simt(g, A_PLAIN, st))) return rc;
    g.A = qh; g.lda = MV2D_C; g.W = w_qe2; g.ldw = MV2D_C; g.bias = b_qe2; g.C = query_pos; g.K = MV2D_C; g.flags = 0;
    return launch_gemm_simt(g, A_PLAIN, st);
}

size_t dn_workspace_bytes(int T, int mask_words) {
    size_t t = (size_t)(T > 0 ? T : 1);
    return t * (384 + MV2D_C) * sizeof(float) + (size_t)(mask_words > 0 ? mask_words : 0) * sizeof(uint32_t) + 256;
}

int run_dn_prepare(const Mv2dDnParams& p, cudaStream_t st) {
    MV2D_CHECK_ARG(p.N >= 1 && p.G >= 0 && p.scalar >= 1, "dn_prepare: bad N=%d / G=%d / scalar=%d", p.N, p.G, p.scalar);
    const int pad = p.G * p.scalar, T = pad + p.N;
    MV2D_CHECK_ARG(p.ref && p.ref_all && p.attn_mask && p.query_pos_all && (pad == 0 || (p.gt_boxes && p.gt_labels && p.rand && p.dn_labels)),
                   "dn_prepare: missing buffer");
    MV2D_CHECK_ARG(dn_workspace_bytes(T, p.mode == 1 ? p.mask_words : 0) <= p.workspace_bytes, "dn_prepare: workspace too small");
    float* pemb = p.workspace;
    float* qh = pemb + (size_t)T * 384;
    uint32_t* uni = reinterpret_cast<uint32_t*>(qh + (size_t)T * MV2D_C);
    launch_k(dn_rows_kernel, dim3(T), dim3(128), 0, st, p, pad, T, pemb);
    MV2D_CHECK_LAUNCH("dn_rows");
    if (p.mode == 0) {
        MV2D_CHECK_ARG(p.match && p.match_cnt && p.match_all && p.match_cnt_all, "dn_prepare: missing match buffers");
        MV2D_CHECK_ARG(p.max_match_all >= p.N && p.max_match_all >= p.max_match, "dn_prepare: max_match_all=%d too small", p.max_match_all);
        launch_k(dn_match_kernel, dim3(T), dim3(128), 0, st, p, pad);
        MV2D_CHECK_LAUNCH("dn_match");
    } else {
        MV2D_CHECK_ARG(p.keymask && p.key_cnt && p.keymask_all && p.key_list_all && p.key_cnt_all, "dn_prepare: missing key-mask buffers");
        MV2D_CHECK_ARG(p.mask_words > 0 && p.mask_words <= 4096, "dn_prepare: mask_words=%d out of range", p.mask_words);
        launch_k(dn_union_kernel, dim3(cdiv(p.mask_words, 128)), dim3(128), 0, st, p, uni);
        MV2D_CHECK_LAUNCH("dn_union");
        launch_k(dn_keys_kernel, dim3(T), dim3(256), p.mask_words * sizeof(uint32_t), st, p, pad, (const uint32_t*)uni);
        MV2D_CHECK_LAUNCH("dn_keys");
    }
    // query_embedding: 384 -> 256 (ReLU) -> 256 over all T rows
    GemmArgs g{};
    g.A = pemb; g.lda = 384; g.W = p.w_qe0; g.ldw = 384; g.bias = p.b_qe0; g.C = qh; g.ldc = MV2D_C;
    g.M = T; g.N = MV2D_C; g.K = 384; g.batch = 1; g.nsplit = 1; g.flags = GEMM_RELU;
    int rc;
    if ((rc = launch_gemm_simt(g, A_PLAIN, st))) return rc;
    g.A = qh; g.lda = MV2D_C; g.W = p.w_qe2; g.ldw = MV2D_C; g.bias = p.b_qe2; g.C = p.query_pos_all; g.K = MV2D_C; g.flags = 0;
    if ((rc = launch_gemm_simt(g, A_PLAIN, st))) return rc;
    return 0;
}

}  // namespace mv2d
