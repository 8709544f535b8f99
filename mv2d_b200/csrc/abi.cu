// extern "C" surface of libmv2d_b200 (include/mv2d_b200.h).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "mv2d_internal.h"

namespace mv2d {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
bool pdl_enabled() {
    static const bool on = []() { const char* e = getenv("MV2D_NO_PDL"); return !(e && e[0] == '1'); }();
    return on;
}
static std::atomic<unsigned long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace mv2d

using namespace mv2d;

extern "C" {

int mv2d_abi_version(void) { return MV2D_ABI_VERSION; }
const char* mv2d_last_error(void) { return g_err; }
unsigned long long mv2d_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

size_t mv2d_sizeof(int which) {
    switch (which) {
        case 0: return sizeof(Mv2dPeParams);
        case 1: return sizeof(Mv2dQgParams);
        case 2: return sizeof(Mv2dCorrParams);
        case 3: return sizeof(Mv2dDecoderParams);
        case 4: return sizeof(Mv2dLayerWeights);
        case 5: return sizeof(Mv2dBranchWeights);
        case 6: return sizeof(Mv2dDnParams);
        case 7: return sizeof(Mv2dKvParams);
        case 8: return sizeof(Mv2dLossParams);
        case 9: return sizeof(Mv2dNeckParams);
        case 10: return sizeof(Mv2dTrainParams);
        case 11: return sizeof(Mv2dFrontTrainParams);
        default: return 0;
    }
}

#define NONNULL(p, what) MV2D_CHECK_ARG((p) != nullptr, what ": null params")

int mv2d_geom_prep(const double* lidar2img, int V, double* img2lidar, double* trans, void* stream) {
    MV2D_CHECK_ARG(lidar2img && img2lidar && trans, "geom_prep: null pointer");
    return run_geom_prep(lidar2img, V, img2lidar, trans, (cudaStream_t)stream, 1);
}

int mv2d_geom_prep_batch(const double* lidar2img, int batch, int V, double* img2lidar, double* trans, void* stream) {
    MV2D_CHECK_ARG(lidar2img && img2lidar && trans, "geom_prep_batch: null pointer");
    return run_geom_prep(lidar2img, V, img2lidar, trans, (cudaStream_t)stream, batch);
}

int mv2d_nchw_to_nhwc(const float* in, float* out, float* out_tf32, int V, int C, int HW, void* stream) {
    MV2D_CHECK_ARG(in && out && V > 0 && C > 0 && HW > 0, "nchw_to_nhwc: bad arguments");
    return run_nchw_to_nhwc(in, out, out_tf32, V, C, HW, (cudaStream_t)stream);
}

int mv2d_nchw_to_nhwc_split(const float* in, float* out, float* out_tf32, float* out_lo, int V, int C, int HW, void* stream) {
    MV2D_CHECK_ARG(in && out && out_tf32 && out_lo && V > 0 && C == MV2D_C && HW > 0, "nchw_to_nhwc_split: bad arguments (C must be 256)");
    return run_nchw_to_nhwc(in, out, out_tf32, V, C, HW, (cudaStream_t)stream, nullptr, out_lo);
}

int mv2d_nchw_add_to_nhwc(const float* in, const float* in2, float* out, int V, int C, int HW, void* stream) {
    MV2D_CHECK_ARG(in && out && V > 0 && C > 0 && HW > 0, "nchw_add_to_nhwc: bad arguments");
    return run_nchw_to_nhwc(in, out, nullptr, V, C, HW, (cudaStream_t)stream, in2);
}

int mv2d_query_embedding(const float* ref, int N, const float* w_qe0, const float* b_qe0, const float* w_qe2, const float* b_qe2,
                         const float* dim_t, float* query_pos, float* workspace, void* stream) {
    MV2D_CHECK_ARG(N >= 0 && (N == 0 || (ref && w_qe0 && b_qe0 && w_qe2 && b_qe2 && dim_t && query_pos && workspace)), "query_embedding: null pointer");
    return run_query_embedding(ref, N, w_qe0, b_qe0, w_qe2, b_qe2, dim_t, query_pos, workspace, (cudaStream_t)stream);
}

int mv2d_split_tf32(const float* x, float* hi, float* lo, long long n, void* stream) {
    MV2D_CHECK_ARG(x && hi && lo && n >= 0, "split_tf32: bad arguments");
    return launch_split_tf32(x, hi, lo, n, (cudaStream_t)stream);
}

int mv2d_gemm_3xtf32(const float* A_hi, const float* A_lo, int lda, const float* W_hi, const float* W_lo, int ldw,
                     const float* bias, float* C, int ldc, int M, int N, int K, int flags, void* stream) {
    MV2D_CHECK_ARG(A_hi && W_hi && C && ((A_lo && W_lo) || (!A_lo && !W_lo)), "gemm_3xtf32: null pointer");
    TcGemm t{};
    t.A = A_hi; t.A_lo = A_lo; t.lda = lda; t.W = W_hi; t.W_lo = W_lo; t.ldw = ldw; t.bias = bias;
    t.C = C; t.ldc = ldc; t.M = M; t.N = N; t.K = K; t.passes = 3; t.im2col = (flags & 128) ? 1 : 0;
    t.flags = flags & GEMM_RELU;
    return launch_gemm_tc(t, (cudaStream_t)stream);
}

size_t mv2d_pe3d_workspace_bytes(int V, int h, int w, int depth_num) { return pe3d_workspace_bytes(V, h, w, depth_num); }
int mv2d_pe3d(const Mv2dPeParams* p, void* stream) {
    NONNULL(p, "pe3d");
    MV2D_CHECK_ARG(p->feat && p->img2lidar && p->pe && p->workspace && p->dim_t && (p->not_mask || p->sine_branch_cached),
                   "pe3d: null pointer");
    return run_pe3d(*p, (cudaStream_t)stream);
}

size_t mv2d_roi_align_qg_workspace_bytes(int N) { return roi_align_qg_workspace_bytes(N); }
int mv2d_roi_align_qg(const Mv2dQgParams* p, void* stream) {
    NONNULL(p, "roi_align_qg");
    MV2D_CHECK_ARG(p->N == 0 || ((p->phase == 3 || (p->rois && p->intrinsics && p->extrinsics && p->feat)) && p->tok_feat && p->ref &&
                                 p->query_pos && p->workspace && p->dim_t),
                   "roi_align_qg: null pointer");
    return run_roi_align_qg(*p, (cudaStream_t)stream);
}

int mv2d_box_corr(const Mv2dCorrParams* p, void* stream) {
    NONNULL(p, "box_corr");
    MV2D_CHECK_ARG(p->N == 0 || (p->rois && p->roi_start && p->trans && p->lin && p->depths && p->match && p->match_cnt),
                   "box_corr: null pointer");
    return run_box_corr(*p, (cudaStream_t)stream);
}

int mv2d_handoff_2d(const float* det, const int* det_start, const float* gt, const int* gt_start, int V, float min_bbox_size,
                    float complement_thr, float* out, int* out_count, void* stream) {
    return run_handoff_2d(det, det_start, gt, gt_start, V, min_bbox_size, complement_thr, out, out_count, (cudaStream_t)stream);
}

size_t mv2d_dn_workspace_bytes(int T, int mask_words) { return dn_workspace_bytes(T, mask_words); }
int mv2d_dn_prepare(const Mv2dDnParams* p, void* stream) {
    NONNULL(p, "dn_prepare");
    MV2D_CHECK_ARG(p->workspace && p->w_qe0 && p->b_qe0 && p->w_qe2 && p->b_qe2 && p->dim_t, "dn_prepare: null pointer");
    return run_dn_prepare(*p, (cudaStream_t)stream);
}

size_t mv2d_decoder_workspace_bytes(int N, int L) { return decoder_workspace_bytes(N, L); }
size_t mv2d_xa_tile_workspace_bytes(int N, int V, int grid_h, int grid_w) { return xa_tile_workspace_bytes(N, V, grid_h, grid_w); }
size_t mv2d_xa_tile_workspace_bytes_batch(int batch, int rows_per_sample, int V, int grid_h, int grid_w) {
    return xa_tile_workspace_bytes(rows_per_sample, V, grid_h, grid_w, batch);
}
int mv2d_xa_tile_prepare(const Mv2dDecoderParams* p, void* stream) {
    NONNULL(p, "xa_tile_prepare");
    MV2D_CHECK_ARG(p->N == 0 || (p->keymask && p->xa_workspace), "xa_tile_prepare: null pointer");
    return run_xa_tile_prepare(*p, (cudaStream_t)stream);
}
int mv2d_kv_project(const Mv2dKvParams* p, void* stream) {
    NONNULL(p, "kv_project");
    MV2D_CHECK_ARG(p->kin_hi && p->mem_hi && p->layers && p->kp && p->vp && ((p->kin_lo && p->mem_lo) || (!p->kin_lo && !p->mem_lo)),
                   "kv_project: null pointer");
    return run_kv_project(*p, (cudaStream_t)stream);
}
int mv2d_cross_attention_core(const Mv2dDecoderParams* p, int layer, const float* q, float* ctx, float* ctx_lo, void* stream) {
    NONNULL(p, "cross_attention_core");
    MV2D_CHECK_ARG(p->kin_rows && p->mem_rows, "cross_attention_core: null pointer");
    return run_cross_attention_core(*p, layer, q, ctx, ctx_lo, (cudaStream_t)stream);
}
int mv2d_decoder(const Mv2dDecoderParams* p, void* stream) {
    NONNULL(p, "decoder");
    MV2D_CHECK_ARG(p->N == 0 || (p->query_pos && p->ref && p->kin_rows && p->mem_rows && p->cls_scores &&
                                 p->bbox_preds && p->outs_dec && p->workspace),
                   "decoder: null pointer");
    return run_decoder(*p, (cudaStream_t)stream);
}

size_t mv2d_loss_workspace_bytes(int N, int G, int L) { return loss_workspace_bytes(N, G, L); }
int mv2d_loss(const Mv2dLossParams* p, void* stream) {
    NONNULL(p, "loss");
    return run_loss(*p, (cudaStream_t)stream);
}

size_t mv2d_fpn_neck_workspace_bytes(int V, int h, int w) { return fpn_neck_workspace_bytes(V, h, w); }
int mv2d_fpn_neck(const Mv2dNeckParams* p, void* stream) {
    NONNULL(p, "fpn_neck");
    MV2D_CHECK_ARG(p->x && p->lat_w && p->lat_w_lo && p->lat_b && p->fpn_w && p->fpn_w_lo && p->fpn_b && p->feat, "fpn_neck: null pointer");
    return run_fpn_neck(*p, (cudaStream_t)stream);
}

long long mv2d_train_param_total(int L) { return train_param_total(L); }
int mv2d_train_set_tensor_cores(int on) { return train_set_tensor_cores(on); }
int mv2d_train_param_info(int L, int tensor_id, long long* offset, long long* numel) {
    return train_param_info(L, tensor_id, offset, numel);
}
size_t mv2d_decoder_train_workspace_bytes(int N, int L, int max_match, int G) { return train_workspace_bytes(N, L, max_match, G); }
size_t mv2d_decoder_train_workspace_bytes_p(const Mv2dTrainParams* p) { return p ? train_workspace_bytes_p(*p) : 0; }
long long mv2d_train_debug_offset(int N, int L, int max_match, int G, int layer, int which) {
    return train_debug_offset(N, L, max_match, G, layer, which);
}
int mv2d_decoder_train_forward(const Mv2dTrainParams* p, void* stream) {
    NONNULL(p, "decoder_train_forward");
    return run_train_forward(*p, (cudaStream_t)stream);
}
int mv2d_decoder_train_backward(const Mv2dTrainParams* p, void* stream) {
    NONNULL(p, "decoder_train_backward");
    return run_train_backward(*p, (cudaStream_t)stream);
}
size_t mv2d_front_train_workspace_bytes(int N, int V, int h, int w) { return front_train_workspace_bytes(N, V, h, w); }
int mv2d_front_train_forward(const Mv2dFrontTrainParams* p, void* stream) {
    NONNULL(p, "front_train_forward");
    return run_front_train_forward(*p, (cudaStream_t)stream);
}
int mv2d_front_train_backward(const Mv2dFrontTrainParams* p, void* stream) {
    NONNULL(p, "front_train_backward");
    return run_front_train_backward(*p, (cudaStream_t)stream);
}
int mv2d_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                    float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream) {
    return run_adamw(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale,
                     (cudaStream_t)stream);
}

int mv2d_gemm(const float* A, int lda, const float* W, int ldw, const float* bias, float* C, int ldc, int M, int N,
              int K, int flags, void* stream) {
    MV2D_CHECK_ARG(A && W && C, "gemm: null pointer");
    GemmArgs g{};
    g.A = A; g.lda = lda; g.W = W; g.ldw = ldw; g.C = C; g.ldc = ldc; g.bias = bias;
    g.M = M; g.N = N; g.K = K; g.batch = 1; g.nsplit = 1; g.flags = flags;
    return launch_gemm_tc_or_simt(g, (cudaStream_t)stream);
}

int mv2d_nms_free_decode(const float* cls, const float* box, int N, int max_num, const float* post_range,
                         float* out_boxes, float* out_scores, int* out_labels, uint8_t* out_valid, void* stream) {
    MV2D_CHECK_ARG(cls && box && post_range && out_boxes && out_scores && out_labels && out_valid,
                   "nms_free_decode: null pointer");
    return run_nms_free_decode(cls, box, N, max_num, post_range, out_boxes, out_scores, out_labels, out_valid,
                               (cudaStream_t)stream);
}

int mv2d_scene_nms(const float* boxes, const float* scores, const int* labels, const uint8_t* valid, int n, float score_thr,
                   float nms_thr, int max_num, float* out_boxes, float* out_scores, int* out_labels, int* out_count, void* stream) {
    MV2D_CHECK_ARG(out_boxes && out_scores && out_labels && out_count && (n == 0 || (boxes && scores && labels)), "scene_nms: null pointer");
    return run_scene_nms(boxes, scores, labels, valid, n, score_thr, nms_thr, max_num, out_boxes, out_scores, out_labels, out_count,
                         (cudaStream_t)stream);
}

int mv2d_debug_clock_probe(long long cycles, long long* out, void* stream) {
    MV2D_CHECK_ARG(out != nullptr && cycles > 0, "clock_probe: bad arguments");
    return run_clock_probe(cycles, out, (cudaStream_t)stream);
}

}  // extern "C"
