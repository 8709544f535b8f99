// Internal declarations shared by the translation units of libmv2d_b200.
#pragma once
#include <cuda_runtime.h>
#include "../../include/mv2d_b200.h"

namespace mv2d {

int run_geom_prep(const double* lidar2img, int V, double* img2lidar, double* trans, cudaStream_t st, int batch = 1);
int run_nchw_to_nhwc(const float* in, float* out, float* out_tf32, int V, int C, int HW, cudaStream_t st, const float* in2 = nullptr,
                     float* out_lo = nullptr);
int run_query_embedding(const float* ref, int N, const float* w_qe0, const float* b_qe0, const float* w_qe2, const float* b_qe2,
                        const float* dim_t, float* query_pos, float* workspace, cudaStream_t st);
int run_pe3d(const Mv2dPeParams& p, cudaStream_t st);
size_t pe3d_workspace_bytes(int V, int h, int w, int depth_num);
int run_roi_align_qg(const Mv2dQgParams& p, cudaStream_t st);
size_t roi_align_qg_workspace_bytes(int N);
int run_box_corr(const Mv2dCorrParams& p, cudaStream_t st);
int run_handoff_2d(const float* det, const int* det_start, const float* gt, const int* gt_start, int V, float min_size, float thr,
                   float* out, int* out_count, cudaStream_t st);
int run_decoder(const Mv2dDecoderParams& p, cudaStream_t st);
int run_cross_attention_core(const Mv2dDecoderParams& p, int layer, const float* q, float* ctx, float* ctx_lo, cudaStream_t st);
size_t decoder_workspace_bytes(int N, int L);
size_t xa_tile_workspace_bytes(int N, int V, int h, int w, int batch = 1);   // N = query rows per sample
int run_kv_project(const Mv2dKvParams& p, cudaStream_t st);
bool kv_persistent_usable(const Mv2dKvParams& p);                       // kvproj.cu
int run_kv_project_persistent(const Mv2dKvParams& p, cudaStream_t st);
int run_xa_tile_prepare(const Mv2dDecoderParams& p, cudaStream_t st);
int run_dn_prepare(const Mv2dDnParams& p, cudaStream_t st);
size_t dn_workspace_bytes(int T, int mask_words);
int run_nms_free_decode(const float* cls, const float* box, int N, int max_num, const float* post_range,
                        float* out_boxes, float* out_scores, int* out_labels, uint8_t* out_valid,
                        cudaStream_t st);

int run_loss(const Mv2dLossParams& p, cudaStream_t st);
size_t loss_workspace_bytes(int N, int G, int L);

int run_fpn_neck(const Mv2dNeckParams& p, cudaStream_t st);
size_t fpn_neck_workspace_bytes(int V, int h, int w);

int run_scene_nms(const float* boxes, const float* scores, const int* labels, const uint8_t* valid, int n, float score_thr,
                  float nms_thr, int max_num, float* out_boxes, float* out_scores, int* out_labels, int* out_count,
                  cudaStream_t st);

int run_clock_probe(long long cycles, long long* out, cudaStream_t st);

// pe.cu, for train.cu
int run_pe_train_inputs(int V, int h, int w, int D, int pad_h, int pad_w, int stride, double depth_start, const double* pr,
                        const double* img2lidar, const uint8_t* not_mask, const float* dim_t, float* coords, float* sine,
                        cudaStream_t st);

// train.cu
size_t front_train_workspace_bytes(int N, int V, int h, int w);
int run_front_train_forward(const Mv2dFrontTrainParams& p, cudaStream_t st);
int run_front_train_backward(const Mv2dFrontTrainParams& p, cudaStream_t st);
long long train_param_total(int L);
int train_set_tensor_cores(int on);
int train_param_info(int L, int tensor_id, long long* offset, long long* numel);
size_t train_workspace_bytes(int N, int L, int max_match, int G);
size_t train_workspace_bytes_p(const Mv2dTrainParams& p);
long long train_debug_offset(int N, int L, int max_match, int G, int layer, int which);
int run_train_forward(const Mv2dTrainParams& p, cudaStream_t st);
int run_train_backward(const Mv2dTrainParams& p, cudaStream_t st);
int run_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps, float wd,
              int step, float grad_scale, cudaStream_t st);

}  // namespace mv2d
