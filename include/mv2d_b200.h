/* libmv2d_b200 -- C ABI of the B200-native MV2D decoder hot path.
 *
 * Drop-in boundary: these entry points are what the reference's Python
 * modules would bind (ctypes) in place of the torch / mmcv composites they run today.  Each
 * entry cites the reference interface it replaces (paths relative to
 * /root/reference/mmdet3d_plugin/models/).
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the field comment says "host";
 *   - caller owns every buffer (inputs, outputs, workspace); the library never allocates,
 *     frees or retains a pointer past the call;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*), no host
 *     synchronisation, no host read of device data => CUDA-graph capturable;
 *   - return 0 = OK; <0 = argument/shape/alignment error (nothing launched);
 *     >0 = cudaError_t of a failed launch.  Text via mv2d_last_error() (thread-local);
 *   - float tensors are fp32, geometry is fp64 where the reference uses .double();
 *   - feature maps are channels-last: [V, h, w, 256]; RoI tokens are [N, 49, 256].
 *
 * Batches (ABI 5).  The reference asserts one sample per call (detectors/mv2d.py:143, roi_heads/mv2d_head.py:210,251).
 * Here B samples travel through ONE kernel chain as a segment dimension:
 *   - the views of all samples are stacked: feature map [B*V, h, w, 256], camera matrices [B*V, 16]; a RoI's view
 *     index (rois[n][0]) is the GLOBAL view b*V + v;
 *   - every sample owns `rows_per_sample` consecutive query rows (row b*rows_per_sample + i); only the first
 *     n_real[b] of them are detections, the rest are padding rows (any valid box of the sample, e.g. the reference's
 *     dummy box) whose results the caller ignores.  Padding rows never enter another row's result: they are outside
 *     every view's roi_start range (box correlation candidates) and outside the self-attention key range;
 *   - self-attention, box correlation and the T head's key masks stay inside a sample; GEMMs / LayerNorms / branches
 *     simply see B*rows_per_sample rows.  Results are those of B separate calls (same arithmetic per row).
 * batch = 0 in a parameter struct means the single-sample layout of ABI 4.
 */
#ifndef MV2D_B200_H_
#define MV2D_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MV2D_API __attribute__((visibility("default")))
#else
#define MV2D_API
#endif

#define MV2D_ABI_VERSION 5
#define MV2D_MAX_LAYERS 8

MV2D_API int mv2d_abi_version(void);
MV2D_API const char* mv2d_last_error(void);
/* number of kernel launches this library has enqueued so far in this process */
MV2D_API unsigned long long mv2d_launch_count(void);
/* sizeof() of the parameter structs, so a binding can verify its mirror of this header */
MV2D_API size_t mv2d_sizeof(int which); /* 0 Pe, 1 Qg, 2 Corr, 3 Decoder, 4 LayerWeights, 5 BranchWeights, 6 Dn, 7 Kv, 8 Loss, 9 Neck, 10 Train, 11 FrontTrain */

/* ---- geometry (utils/pe.py:111; roi_heads/utils/box_correlation.py:118-122,174-178)
 * img2lidar[v] = inv(lidar2img[v]);  trans[src][dst] = lidar2img[dst] @ img2lidar[src]  (fp64) */
MV2D_API int mv2d_geom_prep(const double* lidar2img /*[V,16]*/, int V, double* img2lidar /*[V,16]*/,
                   double* trans /*[V,V,16]*/, void* stream);
/* the same for a batch: lidar2img / img2lidar [batch*V,16], trans [batch,V,V,16] (views of one sample only) */
MV2D_API int mv2d_geom_prep_batch(const double* lidar2img, int batch, int V, double* img2lidar, double* trans, void* stream);

/* NCHW [V,C,HW] -> NHWC [V,HW,C] (the FPN output layout -> this library's layout).
 * out_tf32 (nullable) receives a copy rounded to TF32 (operand of the single-pass tensor-core SE gate GEMM). */
MV2D_API int mv2d_nchw_to_nhwc(const float* in, float* out, float* out_tf32, int V, int C, int HW, void* stream);
/* the same, additionally writing out_lo = tf32(x - tf32(x)): with out_tf32 the hi / lo operand pair of a 3xTF32 GEMM over the map
 * (two-frame head: the value rows of mv2d_kv_project), so no separate mv2d_split_tf32 pass over it is needed */
MV2D_API int mv2d_nchw_to_nhwc_split(const float* in, float* out, float* out_tf32, float* out_lo, int V, int C, int HW, void* stream);
/* the same with a second map added on the way: out = nhwc(in + in2)  (key = memory + key_pos of the dense interfaces) */
MV2D_API int mv2d_nchw_add_to_nhwc(const float* in, const float* in2, float* out, int V, int C, int HW, void* stream);
/* CrossAttentionBoxHead.position_embedding (bbox_heads/cross_attention_head.py:199-200):
 * query_pos = query_embedding(pos2posemb3d(ref)); workspace: N * 640 floats */
MV2D_API int mv2d_query_embedding(const float* ref /*[N,3]*/, int N, const float* w_qe0, const float* b_qe0, const float* w_qe2,
                                  const float* b_qe2, const float* dim_t, float* query_pos /*[N,256]*/, float* workspace,
                                  void* stream);

/* ---- K1  PE.forward  (utils/pe.py:137-169 incl. position_encoding :84-135, SELayer :44-48,
 * SinePositionalEncoding3D positional_encoding.py:58-96 + adapt_pos3d) */
typedef struct Mv2dPeParams {
    /* phase: 0 = everything; 1 = the part that does not read `feat` (frustum coordinates, position MLP, sine
     *        branch: can overlap the host-to-device copy of the feature map); 2 = SE gate + combine (after 1) */
    int V, h, w, depth_num, pad_h, pad_w, stride, phase;
    double depth_start;
    double position_range[6];
    const float* feat;         /* [V,h,w,256] */
    const float* feat_tf32;    /* nullable: feat rounded to TF32 (from mv2d_nchw_to_nhwc); NULL = use feat */
    const double* img2lidar;   /* [V,16] from mv2d_geom_prep */
    const uint8_t* not_mask;   /* [V,h,w] 1 = inside the un-padded image */
    const float* dim_t;        /* [128] temperature ** (2*(i//2)/128) */
    /* the six weight matrices below must be pre-rounded to TF32 (mv2d_b200/pack.py does it):
     * the PE MLPs run as single-pass tcgen05 kind::tf32 GEMMs (inside the gate) */
    const float *w_pos0, *b_pos0;       /* position_encoder.0  [1024,192] */
    const float *w_pos2, *b_pos2;       /* position_encoder.2  [256,1024] */
    const float *w_adapt0, *b_adapt0;   /* adapt_pos3d.0       [1024,384] */
    const float *w_adapt2, *b_adapt2;   /* adapt_pos3d.2       [256,1024] */
    const float *w_se_reduce, *b_se_reduce; /* fpe.conv_reduce [256,256] */
    const float *w_se_expand, *b_se_expand; /* fpe.conv_expand [256,256] */
    const float* sine_branch_cached;    /* nullable [V*h*w,256]: adapt_pos3d(sine) computed earlier */
    float* sine_branch_out;             /* nullable: receives adapt_pos3d(sine) */
    float* pe;                 /* out [V,h,w,256] */
    float* kin;                /* out, nullable: feat + pe (keys of the two-frame head) */
    float* workspace;
    size_t workspace_bytes;
    int sine_separable;        /* 1 = the caller guarantees not_mask is all ones (no padded cells): the first layer of
                                * the sine branch is evaluated as W1.s(v,y,x) = Tv[v] + Ty[y] + Tx[x] from three
                                * (V + h + w)-row tables instead of a GEMM over all V*h*w cells; 0 = general masks */
    int views_per_sample;      /* batch: V counts the views of all samples, this many belong to one sample (0 = V).
                                * The view axis of SinePositionalEncoding3D (cumsum over views) is per sample. */
    int sine_shared;           /* batch: 1 = the caller guarantees all samples have identical not_mask, so
                                * adapt_pos3d(sine) is the same tensor for every sample: it is evaluated once, for the
                                * first views_per_sample views, and shared (a common subexpression, not a cache: it is
                                * recomputed on every call) */
    int unfused_mlp;           /* 1 = one GEMM per MLP layer (the 1024-wide hidden activations round-trip through HBM);
                                * 0 = the fused two-layer kernel (csrc/mlp2.cu): hidden activations stay in TMEM / shared
                                * memory.  Both are single-pass TF32; kept selectable for A/B tests. */
} Mv2dPeParams;
MV2D_API size_t mv2d_pe3d_workspace_bytes(int V, int h, int w, int depth_num);
MV2D_API int mv2d_pe3d(const Mv2dPeParams* p, void* stream);

/* ---- K2  RoIAlign + dynamic query generator + query embedding
 * (roi_heads/mv2d_head.py:51-72,95-101,114-134; utils/query_generator.py:333-405;
 *  bbox_heads/cross_attention_head.py:199-200; utils/pe.py:21-33; mmcv RoIAlign avg/aligned) */
typedef struct Mv2dQgParams {
    /* phase: 0 = everything; 1 = everything that does not need `pe` (can run concurrently with mv2d_pe3d);
     *        2 = only tok_kin = tok_feat + RoIAlign(pe) (after phases 1 and mv2d_pe3d);
     *        4 = only roi_intrinsics (get_box_params, mv2d_head.py:51-72: reads rois / intrinsics / extrinsics, no weights) */
    int N, V, h, w, stride, phase;
    float pc_range[6];
    float intrins_feat_scale;
    float reserved1;
    const float* rois;          /* [N,5] (view, x1, y1, x2, y2) px */
    const double* intrinsics;   /* [V,16] */
    const double* extrinsics;   /* [V,16] (= lidar2cam transposed) */
    const float* feat;          /* [V,h,w,256] */
    const float* pe;            /* [V,h,w,256], nullable when tok_kin == NULL */
    const float* dim_t;         /* [128] */
    const float *w_conv, *b_conv;     /* shared_convs.0.conv repacked [256, 9*256] (tap-major K), TF32 "hi" part */
    const float *w_conv_lo;           /* TF32 "lo" part: w = hi + lo (3xTF32 error-compensated tcgen05 GEMM) */
    const float *w_fc, *b_fc;         /* shared_fcs.0 [1024,256] */
    const float *w_enc0, *b_enc0;     /* extra_enc.0 [512,1040] zero-padded along K to [512,1056] */
    const float *w_enc2, *b_enc2;     /* extra_enc.2 [256,512] */
    const float *w_center, *b_center; /* fc_center [3,256] */
    const float *w_qe0, *b_qe0;       /* query_embedding.0 [256,384] */
    const float *w_qe2, *b_qe2;       /* query_embedding.2 [256,256] */
    float* tok_feat;       /* out [N,49,256] RoI-pooled image feature (bbox_feats, channels-last) */
    float* tok_kin;        /* out, nullable [N,49,256] RoI-pooled (feat + pe) */
    double* roi_intrinsics;/* out, nullable [N,16] K' */
    float* center_lidar;   /* out, nullable [N,3] */
    float* ref;            /* out [N,3] normalised reference points */
    float* query_pos;      /* out [N,256] */
    float* workspace;
    size_t workspace_bytes;
    /* ---- ABI 5, all nullable: TF32 hi / lo splits of the five FC matrices.  With them, and more than 512 RoIs
     * (batches), the FC chain runs as 3xTF32 tcgen05 GEMMs instead of the small-M FFMA kernels. */
    const float *w_fc_hi, *w_fc_lo, *w_enc0_hi, *w_enc0_lo, *w_enc2_hi, *w_enc2_lo, *w_qe0_hi, *w_qe0_lo, *w_qe2_hi, *w_qe2_lo;
    /* ---- ABI 5, phase 3 = QueryGenerator.forward on its own (utils/query_generator.py:343-350): the RoI features and
     * the per-RoI camera parameters are INPUTS -- tok_feat [N,49,256] (channels-last RoI features), roi_intrinsics
     * [N,16] fp64 (K' of get_box_params), roi_extrinsics [N,16] fp64, intrins_feat [N,16] (extra_feats['intrinsic']);
     * rois / feat / intrinsics / extrinsics are not read.  Outputs as in phase 0 (center_lidar, ref, query_pos). */
    const double* roi_extrinsics;
    const float* intrins_feat;
    float* enc_out;            /* out, nullable (any phase but 2) [N,256]: the encoded RoI feature (return_feats['enc']) */
} Mv2dQgParams;
MV2D_API size_t mv2d_roi_align_qg_workspace_bytes(int N);
MV2D_API int mv2d_roi_align_qg(const Mv2dQgParams* p, void* stream);

/* ---- K3  BoxCorrelation  (roi_heads/utils/box_correlation.py:95-398, topk_matched mode)
 * S head: match list per RoI (self first).  T head: additionally the bit-packed per-query key
 * mask over the [V,h,w] feature cells (bit c of word c/32, c = (v*h+y)*w+x). */
typedef struct Mv2dCorrParams {
    int N, V, img_h, img_w;
    int topk, sample_size, num_depth, max_match;
    float ratio, iou_thr, depth_start, reserved0;
    const float* rois;        /* [N,5] */
    const int* roi_start;     /* [V+1] first RoI of each view (RoIs are sorted by view) */
    const double* trans;      /* [V,V,16] from mv2d_geom_prep */
    const float* lin;         /* [sample_size] torch.linspace(0,1,S) */
    const float* depths;      /* [num_depth] LID depth bins, fp32 */
    int* match;               /* out [N,max_match] */
    int* match_cnt;           /* out [N] */
    /* two-frame head only (keymask != NULL) */
    int h, w, stride, expand_stride;
    const uint8_t* pad_mask;  /* nullable [V,h,w] 1 = padded-out cell (key_padding_mask) */
    uint32_t* keymask;        /* out, nullable [N, ceil(V*h*w/32)] */
    int* key_cnt;             /* out, nullable [N] */
    uint16_t* key_list;       /* out, nullable [N, V*h*w]: the set bits of keymask in ascending order (compacted once
                               * here instead of once per decoder layer); needs key_cnt */
    /* ---- ABI 5: batch > 0: N = batch * rows_per_sample query rows, V = views of ONE sample, rois[n][0] = global view,
     * roi_start [batch, V+1] global row ranges of the real detections (roi_start[b][V] = b*rows_per_sample + n_real[b]),
     * trans [batch,V,V,16], pad_mask [batch*V,h,w]; keymask bits index the cells of the row's own sample
     * ((v_local*h+y)*w+x) */
    int batch, rows_per_sample;
} Mv2dCorrParams;
MV2D_API int mv2d_box_corr(const Mv2dCorrParams* p, void* stream);

/* ---- K4/K5  decoder: MV2DTransformer + PETRTransformerDecoder + branches
 * (bbox_heads/cross_attention_head.py:22-49,202-242; utils/petr_transformer.py:194-593) */
typedef struct Mv2dLayerWeights {
    const float *sa_in_w, *sa_in_b;     /* attentions.0.attn.in_proj [768,256] */
    const float *sa_out_w, *sa_out_b;   /* attentions.0.attn.out_proj [256,256] */
    /* the four wide matrices run as 3xTF32 tcgen05 GEMMs: w = hi + lo, both TF32-representable */
    const float *ca_q_w, *ca_q_w_lo, *ca_q_b;   /* absorbed  scale*Wk_h^T Wq_h : [2048,256], [2048] */
    const float *ca_o_w, *ca_o_w_lo, *ca_o_b;   /* absorbed  Wo[:,h] Wv_h      : [256,2048], [256]  */
    const float *ffn_w1, *ffn_w1_lo, *ffn_b1;   /* ffns.0.layers.0.0 [2048,256] */
    const float *ffn_w2, *ffn_w2_lo, *ffn_b2;   /* ffns.0.layers.1   [256,2048] */
    const float *ln_g[3], *ln_b[3];     /* norms.{0,1,2} */
    const float *sa_const;              /* nullable, read for layers[0] only: [256] = out_proj(bv) + bo.  The decoder's
                                         * target starts at zero (cross_attention_head.py:32), so in the first layer
                                         * every self-attention VALUE row is the bias bv and the attention output is
                                         * this constant for every query, whatever q, k and the mask are.  With it the
                                         * staged decoder skips layer 0's in_proj / attention / out_proj launches. */
    /* plain (not absorbed) cross-attention projections, used by the key-stationary form of the two-frame head
     * (xa_form 1): attentions.1.attn.in_proj / out_proj split per role */
    const float *xa_q_w, *xa_q_b;       /* Wq / sqrt(32), bq / sqrt(32) : [256,256], [256] */
    const float *xa_k_w, *xa_k_w_lo;    /* Wk as TF32 hi + lo : [256,256] (bk only shifts the logits of a query by a
                                         * constant, softmax cancels it) */
    const float *xa_v_w, *xa_v_w_lo;    /* Wv as TF32 hi + lo : [256,256] */
    const float *xa_o_w, *xa_o_b;       /* Wo [256,256], Wo bv + bo [256] (probabilities sum to one) */
    const float *xa_k_raw, *xa_v_raw;   /* Wk, Wv as plain fp32 (= hi + lo): operands of the in-kernel-split GEMM */
    /* ---- ABI 5: with more than 512 query rows (batches) every decoder GEMM runs as 3xTF32 on the tensor cores.
     * sa_in_w / sa_out_w / xa_q_w / xa_o_w above stay plain fp32 (the small-M FFMA kernels read them); these are their
     * TF32 hi / lo splits (w = hi + lo).  All nullable: NULL keeps the FFMA kernels at every size. */
    const float *sa_in_w_hi, *sa_in_w_lo;     /* [768,256] */
    const float *sa_out_w_hi, *sa_out_w_lo;   /* [256,256] */
    const float *xa_q_w_hi, *xa_q_w_lo;       /* [256,256] */
    const float *xa_o_w_hi, *xa_o_w_lo;       /* [256,256] */
} Mv2dLayerWeights;

typedef struct Mv2dBranchWeights {      /* stacked over layers: leading dim L */
    const float *cls_w0, *cls_b0, *cls_g0, *cls_be0;   /* Linear [L,256,256], LN */
    const float *cls_w1, *cls_b1, *cls_g1, *cls_be1;
    const float *cls_w2, *cls_b2;                      /* [L,10,256] */
    const float *reg_w0, *reg_b0, *reg_w1, *reg_b1;    /* [L,256,256] */
    const float *reg_w2, *reg_b2;                      /* [L,10,256] */
    const float *post_g, *post_b;                      /* decoder.post_norm */
    /* ABI 5, nullable: TF32 hi / lo splits of the four [L,256,256] branch matrices (3xTF32 tensor-core GEMMs for
     * batches of more than 512 query rows; NULL keeps the FFMA kernels) */
    const float *cls_w0_hi, *cls_w0_lo, *cls_w1_hi, *cls_w1_lo, *reg_w0_hi, *reg_w0_lo, *reg_w1_hi, *reg_w1_lo;
} Mv2dBranchWeights;

/* ---- weight packing (host only, no CUDA calls): the one-time re-layout of the reference state_dict into the buffers
 * the kernels read -- absorbed cross-attention matrices (fp64, rounded once), K-major 3x3 convolutions, TF32 hi / lo
 * splits, stacked branches, layer 0's constant self-attention output.  Replaces what a reference user gets from
 * `load_state_dict` (mmdet3d checkpoint keys; the `roi_head.` prefix is optional).
 *   1. bytes = mv2d_pack_weights_bytes(L, fold, &n_entries); allocate a device arena (256-byte aligned) and a host image
 *   2. mv2d_pack_weights(...): fills the host image, the directory (name -> byte offset) and the two weight structs with
 *      DEVICE pointers (device_base + offset)
 *   3. copy the host image to the device arena (one cudaMemcpy)
 * Directory names are the ones the other parameter structs use (w_pos0, b_pos0, ..., w_conv, w_conv_lo, w_fc, w_fc_hi,
 * ..., l<k>.<field>, br.<field>, post_g, post_b, dim_t). */
typedef struct Mv2dNamedTensor { const char* name; const float* data; int64_t numel; } Mv2dNamedTensor;   /* HOST fp32 */
typedef struct Mv2dPackedEntry { char name[48]; int64_t offset /*bytes into the arena*/; int64_t numel; } Mv2dPackedEntry;
MV2D_API int64_t mv2d_pack_weights_bytes(int num_layers, int fold_first_self_attn, int* n_entries);
MV2D_API int mv2d_pack_weights(const Mv2dNamedTensor* state_dict, int n_tensors, int num_layers, int fold_first_self_attn,
                               void* host_arena, int64_t arena_bytes, const void* device_base, Mv2dPackedEntry* dir,
                               int dir_cap, int* n_dir, struct Mv2dLayerWeights* layers /*out [num_layers]*/,
                               struct Mv2dBranchWeights* branches /*out*/);
/* `neck.*` (one-level FPN, configs/mv2d/exp/*.py:32-39) as the operands of mv2d_fpn_neck: lat_w / lat_w_lo / lat_b,
 * fpn_w / fpn_w_lo (K ordered (ky, kx, c_in)) / fpn_b.  Keys with or without the `neck.` prefix. */
MV2D_API int64_t mv2d_pack_neck_bytes(int* n_entries);
MV2D_API int mv2d_pack_neck(const Mv2dNamedTensor* state_dict, int n_tensors, void* host_arena, int64_t arena_bytes,
                            const void* device_base, Mv2dPackedEntry* dir, int dir_cap, int* n_dir);

typedef struct Mv2dDecoderParams {
    int N, L, mode /*0 = RoI-token keys (S), 1 = feature-map keys (T)*/, num_rows;
    int max_match, mask_words;
    int persistent;             /* 1 = all layers + branches in ONE cooperative launch (one CTA per SM, device-wide
                                 *     barriers between stages); 0 = one launch per stage (~75 launches) */
    int vel_row_start;          /* vel_dt applies to query rows >= this (denoising rows are prepended and are not
                                 *     rescaled, mv2d_t_head.py:112-118 vs :136-140); 0 = all rows */
    float pc_range[6];
    float vel_dt;               /* T head: bbox_preds[..., 8:10] /= vel_dt ; 0 = off (mv2d_t_head.py:130-142) */
    float reserved2;
    const float* query_pos;     /* [N,256] */
    const float* ref;           /* [N,3] */
    const float* kin_rows;      /* [num_rows,256] key input  (memory + pos) */
    const float* mem_rows;      /* [num_rows,256] value input (memory) */
    const int* match;           /* mode 0: [N,max_match] RoI ids */
    const int* match_cnt;       /* mode 0: [N] */
    const uint32_t* keymask;    /* mode 1: [N,mask_words] */
    const uint16_t* key_list;   /* mode 1, nullable: [N, mask_words*32] compacted key ids from mv2d_box_corr */
    const int* key_cnt;         /* mode 1, with key_list: [N] */
    const uint8_t* self_attn_mask; /* nullable [N,N] 1 = masked (DN training) */
    const Mv2dLayerWeights* layers;   /* HOST array [L] */
    const Mv2dBranchWeights* branches;/* HOST pointer */
    float* cls_scores;          /* out [L,N,10] */
    float* bbox_preds;          /* out [L,N,10] */
    float* outs_dec;            /* out [L,N,256] post-normed intermediates */
    float* workspace;
    size_t workspace_bytes;
    /* ---- ABI 3 */
    int layer_begin, layer_end; /* run decoder layers [layer_begin, layer_end) (layer_end 0 = L); the branches run when
                                 * layer_end reaches L.  The state between calls lives in `workspace`, so a caller can
                                 * interleave the layers with mv2d_kv_project on another stream. */
    int xa_form;                /* mode 1 only: 0 = query-stationary absorbed form (streams each query's raw key rows),
                                 *              1 = key-stationary form over 8x8 cell tiles of projected K / V */
    int grid_h, grid_w;         /* xa_form 1: the feature grid; num_rows = V * grid_h * grid_w */
    int xa_prepared;            /* xa_form 1: 1 = mv2d_xa_tile_prepare already ran on xa_workspace for this sample */
    const float* kp;            /* xa_form 1: [L,num_rows,256] projected keys   from mv2d_kv_project */
    const float* vp;            /* xa_form 1: [L,num_rows,256] projected values from mv2d_kv_project */
    void* xa_workspace;         /* xa_form 1: mv2d_xa_tile_workspace_bytes(N, V, grid_h, grid_w) */
    size_t xa_workspace_bytes;
    uint8_t* row_tile_live;     /* nullable [ceil(num_rows/128)]: 1 = some query has a key among rows 128 t .. 128 t + 127.
                                 * mv2d_xa_tile_prepare WRITES it (it is then the input of mv2d_kv_project);
                                 * mv2d_decoder / mv2d_cross_attention_core READ it when it is passed: set it exactly when
                                 * the projection that made kp / vp was given the same flags -- the tensor-core tile
                                 * attention loads whole 8x8 cell tiles and zero-fills the cells of skipped row tiles
                                 * instead of reading their stale kp / vp rows.  NULL = every row was projected. */
    /* ---- ABI 5: batch > 0: N = batch * rows_per_sample; self-attention keys of a row are the first n_real[b] rows of
     * its sample; mode 0: match ids are global rows; mode 1 (xa_form 1 only): num_rows = batch * V*h*w rows of
     * kin_rows / mem_rows / kp / vp, keymask bits and mask_words refer to ONE sample's V*h*w cells */
    int batch, rows_per_sample;
    const int* n_real;          /* device [batch], nullable = every row is real */
    const float* vel_dt_batch;  /* device [batch], nullable: per-sample vel_dt (replaces the scalar; 0 = off for that sample) */
} Mv2dDecoderParams;
MV2D_API size_t mv2d_decoder_workspace_bytes(int N, int L);
MV2D_API size_t mv2d_xa_tile_workspace_bytes(int N, int V, int grid_h, int grid_w);
/* batch > 0: `batch` samples of rows_per_sample query rows and V views each */
MV2D_API size_t mv2d_xa_tile_workspace_bytes_batch(int batch, int rows_per_sample, int V, int grid_h, int grid_w);
/* xa_form 1: per-tile query lists, per-query record lists and the tile order, built from the key masks alone
 * (so it can run right after mv2d_box_corr, beside the position embedding).  Reads N, num_rows, grid_h, grid_w,
 * keymask, mask_words, xa_workspace(_bytes) of the decoder parameters.  A decoder call with xa_prepared = 1 then
 * skips this step. */
MV2D_API int mv2d_xa_tile_prepare(const Mv2dDecoderParams* p, void* stream);
MV2D_API int mv2d_decoder(const Mv2dDecoderParams* p, void* stream);
/* The sparse cross-attention core of decoder layer `layer` on its own -- PETRMultiheadAttention between its query and
 * output projections (utils/petr_transformer.py:426-513) -- for stage-level bindings and the attention-only
 * microbenchmark (BASELINE configs[4]).  Reads the key description of `p` (match lists / prepared key tiles).
 *   mode 0: q = absorbed queries [N,2048] (scale * Wk_h^T (Wq_h x + bq_h)), ctx [N,2048] = per-head softmax-weighted
 *           sums of the raw value rows; uses the partial-record scratch inside p->workspace;
 *   mode 1 (xa_form 1, xa_prepared): q = projected queries [N,256], ctx [N,256] (heads concatenated).
 * ctx_lo (nullable): ctx then receives the TF32 hi part and ctx_lo the lo part. */
MV2D_API int mv2d_cross_attention_core(const Mv2dDecoderParams* p, int layer, const float* q, float* ctx, float* ctx_lo,
                                       void* stream);

/* ---- K/V projection of the two-frame head's keys (SURVEY.md 8b "kv_proj"; utils/petr_transformer.py:503-508 ->
 * torch.nn.MultiheadAttention in_proj on key = memory + pos and value = memory).  For every layer l in
 * [layer_begin, layer_end):  kp[l] = (kin_hi + kin_lo) (Wk_l)^T,  vp[l] = (mem_hi + mem_lo) (Wv_l)^T  as
 * error-compensated 3xTF32 tcgen05 GEMMs (fp32-grade).  The operands are the TF32 splits mv2d_split_tf32 makes. */
typedef struct Mv2dKvParams {
    int num_rows, L, layer_begin, layer_end;
    const float *kin_hi, *kin_lo;      /* [num_rows,256]; kin_lo == mem_lo == NULL: kin_hi / mem_hi are the plain fp32 */
    const float *mem_hi, *mem_lo;      /* [num_rows,256]  rows and the GEMM splits operands and weights in-kernel    */
    const Mv2dLayerWeights* layers;    /* HOST array [L] */
    float* kp;                         /* out [L,num_rows,256] */
    float* vp;                         /* out [L,num_rows,256] */
    const uint8_t* row_tile_live;      /* nullable, device [ceil(num_rows/128)] from mv2d_xa_tile_prepare: 128-row tiles no
                                        * query has a key in are not projected (their kp / vp rows stay untouched and
                                        * are never read by the attention) */
} Mv2dKvParams;
MV2D_API int mv2d_kv_project(const Mv2dKvParams* p, void* stream);

/* ---- f2 (next row): the 2D-detections hand-off of MV2D.forward_train, boxes staying on the device
 * (detectors/mv2d.py:60-86 process_2d_detections' min-size filter; :88-117 box_iou + complement_2d_gt).
 * det [n,6] = (x1, y1, x2, y2, score, label) of all views, view v in rows det_start[v] .. det_start[v+1]; gt [m,6] the
 * 2D ground truth in the same format (score 1, process_2d_gt) with gt_start.  For every view: the detections whose
 * sides reach min_bbox_size, in input order, then -- complement_thr > 0 -- the ground-truth boxes whose best IoU with
 * those is below complement_thr and whose sides reach min_bbox_size (all of them, unfiltered, when no detection is
 * left: the reference's early return).  View v's result starts at row det_start[v] + gt_start[v] of `out`
 * ([n + m, 6]) and has out_count[v] rows.  gt / gt_start may be NULL when complement_thr <= 0. */
MV2D_API int mv2d_handoff_2d(const float* det, const int* det_start /*device [V+1]*/, const float* gt, const int* gt_start /*device [V+1]*/,
                             int V, float min_bbox_size, float complement_thr, float* out, int* out_count /*device [V]*/, void* stream);

/* ---- row a20: denoising queries of the training-mode forward -----------------------------------------------
 * Replaces MV2DSHead.prepare_for_dn (mv2d_s_head.py:39-120, training branch, batch_size 1) plus the way both
 * heads extend the cross-attention mask for the prepended queries (mv2d_s_head.py:159-172,
 * mv2d_t_head.py:79-98) and the query embedding of the padded reference points
 * (cross_attention_head.py:199-206).  T = scalar*G + N query rows come out, denoising rows first:
 *   ref_all[i]   i < pad: clamp(normalise(centre_g + (2*rand-1) * (size_g/2 + noise_trans) * noise_scale), eps, 1-eps)
 *                         with g = i % G ; i >= pad: ref[i-pad]
 *   dn_labels[i] = |2*rand-1|_2 > split ? num_classes : gt_labels[g]
 *   attn_mask    [T,T] u8, 1 = masked: matching rows do not see denoising rows, denoising groups do not see
 *                each other
 *   keys         S: match_all[i] = every RoI for i < pad, the RoI's own match list otherwise
 *                T: keymask_all[i] = OR of all rows for i < pad, the own row otherwise; with train_unmask a
 *                   matching row without any key gets key 0 (mv2d_t_head.py:80-82)
 *   query_pos_all = query_embedding(pos2posemb3d(ref_all))
 * The uniform noise is an INPUT (the caller draws it, as torch.rand_like does in the reference). */
typedef struct Mv2dDnParams {
    int N, G, scalar, num_classes;
    int mode;                   /* 0 = S (match lists), 1 = T (key masks) */
    int max_match;              /* S: row stride of `match` */
    int max_match_all;          /* S: row stride of `match_all`, >= max(N, max_match) */
    int mask_words;             /* T */
    int train_unmask;           /* T */
    int reserved0;
    float noise_scale, noise_trans, split, eps;
    float pc_range[6];
    const float* gt_boxes;      /* [G,9] gravity centre xyz, w, l, h, yaw, vx, vy */
    const int* gt_labels;       /* [G] */
    const float* rand;          /* [scalar*G,3] uniform [0,1) */
    const float* ref;           /* [N,3] from mv2d_roi_align_qg */
    const int* match;           /* S: [N,max_match] from mv2d_box_corr */
    const int* match_cnt;       /* S: [N] */
    const uint32_t* keymask;    /* T: [N,mask_words] from mv2d_box_corr */
    const int* key_cnt;         /* T: [N] */
    const float *w_qe0, *b_qe0, *w_qe2, *b_qe2, *dim_t;   /* query_embedding weights, as in Mv2dQgParams */
    float* ref_all;             /* out [T,3] */
    int* dn_labels;             /* out [scalar*G] */
    uint8_t* attn_mask;         /* out [T,T] */
    float* query_pos_all;       /* out [T,256] */
    int* match_all;             /* S out [T,max_match_all] */
    int* match_cnt_all;         /* S out [T] */
    uint32_t* keymask_all;      /* T out [T,mask_words] */
    uint16_t* key_list_all;     /* T out [T,mask_words*32] */
    int* key_cnt_all;           /* T out [T] */
    float* workspace;
    size_t workspace_bytes;
} Mv2dDnParams;
MV2D_API size_t mv2d_dn_workspace_bytes(int T, int mask_words);
MV2D_API int mv2d_dn_prepare(const Mv2dDnParams* p, void* stream);

/* ---- f3 (next row): training targets and losses of one sample, all decoder layers in one call
 * (core/bbox/assigners/hungarian_assigner_3d.py:66-150; core/bbox/match_costs/match_cost.py:6-26;
 *  core/bbox/util.py:38-58; bbox_heads/cross_attention_head.py:244-343,379-434 loss_single, :475-538 dn_loss_single,
 *  as roi_heads/mv2d_s_head.py:278-299 calls them: per layer, one sample).  For every layer l:
 *   cost[n,g]   = cls_cost_weight * FocalLossCost(cls[l,n], label_g) + reg_cost_weight * |box[l,n,:8] - norm(gt_g)[:8]|_1
 *   assigned[l] = linear sum assignment of cost (the reference: scipy on the host after a .cpu() sync), -1 = background
 *   losses[l]   = { cls_loss_weight * focal(cls, targets) / max(num_pos, 1),
 *                   bbox_loss_weight * sum_pos |box - norm(gt)| * code_weights / max(num_pos, 1),
 *                   the same two for the denoising queries (avg factors pad*pi/6*split^3 and pad; codes 6:8 unweighted) }
 * Ties between equal costs are broken in favour of unassigned columns, then the lower index (scipy's scan order can
 * differ there; exact ties do not occur with real-valued predictions). */
typedef struct Mv2dLossParams {
    int N, G, L, num_classes;
    int pad;                    /* number of denoising queries, 0 = none */
    int neg_bbox_loss;          /* 1: negative denoising queries keep their box target (exp two_frames config :45,
                                 *    cross_attention_head.py:521-523); 0: only positives enter the denoising box loss */
    long long layer_stride;     /* elements between layers of cls_scores / bbox_preds (rows are [*,10] contiguous) */
    long long dn_layer_stride;  /* same for dn_cls / dn_box */
    float cls_cost_weight, reg_cost_weight, cls_loss_weight, bbox_loss_weight;
    float focal_alpha, focal_gamma, dn_split, reserved1;
    float code_weights[10];
    const float* cls_scores;    /* [L,N,num_classes] */
    const float* bbox_preds;    /* [L,N,10] */
    const float* gt_boxes;      /* [G,9] gravity centre xyz, w, l, h, yaw, vx, vy */
    const int* gt_labels;       /* [G] */
    const float* dn_cls;        /* nullable [L,pad,num_classes] */
    const float* dn_box;        /* nullable [L,pad,10] */
    const int* dn_labels;       /* nullable [pad]: label, num_classes = negative; query i was noised from box i % G */
    int* assigned;              /* out [L,N] */
    float* losses;              /* out [L,4]: loss_cls, loss_bbox, dn_loss_cls, dn_loss_bbox */
    float* workspace;
    size_t workspace_bytes;
    /* ---- ABI 5: the reference divides loss_bbox by clamp(reduce_mean(num_total_pos), min=1) taken ACROSS RANKS
     * (cross_attention_head.py:419-420); loss_cls keeps the local count (sync_cls_avg_factor = False). */
    float* num_pos;             /* out, nullable [L]: positive matching queries of every layer (local count, as float) */
    const float* bbox_avg_factor; /* in, nullable [L]: replaces max(num_pos, 1) as the avg factor of loss_bbox (the caller
                                 * all-reduces num_pos, divides by the number of samples and clamps at 1) */
} Mv2dLossParams;
MV2D_API size_t mv2d_loss_workspace_bytes(int N, int G, int L);
MV2D_API int mv2d_loss(const Mv2dLossParams* p, void* stream);

/* ---- f4 (next row): the MV2D neck = one-level mmdet FPN on the 2D detector's P4 (configs/mv2d/exp/*.py:32-39,
 * detectors/mv2d.py:122-127):  feat = conv3x3(conv1x1(x) + lat_b, padding 1) + fpn_b,  emitted channels-last
 * (the layout the rest of this library consumes), both convolutions as 3xTF32 tcgen05 GEMMs (fp32-grade). */
typedef struct Mv2dNeckParams {
    int V, h, w, in_is_nhwc;
    const float* x;                          /* detector P4: [V,256,h,w], or [V,h,w,256] when in_is_nhwc */
    const float *lat_w, *lat_w_lo, *lat_b;   /* neck.lateral_convs.0.conv: [256,256] as TF32 hi + lo, bias [256] */
    const float *fpn_w, *fpn_w_lo, *fpn_b;   /* neck.fpn_convs.0.conv: [256, 9*256], K ordered (ky, kx, c_in), hi + lo, bias */
    float* feat;                             /* out [V,h,w,256] */
    float* feat_tf32;                        /* out, nullable: feat rounded to TF32 (as mv2d_nchw_to_nhwc emits it) */
    float* workspace;
    size_t workspace_bytes;
} Mv2dNeckParams;
MV2D_API size_t mv2d_fpn_neck_workspace_bytes(int V, int h, int w);
MV2D_API int mv2d_fpn_neck(const Mv2dNeckParams* p, void* stream);

/* ---- row e / BASELINE configs[3]: the training step of the MV2D-S decoder slice (ABI 4) ---------------------------
 * Forward with saved activations + Hungarian targets + losses, then the BACKWARD (torch autograd in the reference)
 * of: CrossAttentionBoxHead.forward (bbox_heads/cross_attention_head.py:199-242: query embedding, MV2DTransformer /
 * PETRTransformerDecoder utils/petr_transformer.py:194-593, post_norm, cls / reg branches, reference-point
 * refinement) and loss_single (:379-434), summed over layers with stage_loss_weights as MV2DSHead.forward_train does
 * (roi_heads/mv2d_s_head.py:262-307).  The single-frame configs train WITHOUT denoising queries
 * (configs/mv2d/exp/mv2d_r50_frcnn_single_frame_roi_1408x512_ep*.py:44), so query i attends to the 49 tokens of
 * every RoI in match[i][0 .. match_cnt[i]) and self-attention is unmasked.
 *
 * Parameters and gradients live in ONE flat fp32 buffer each (so the data-parallel exchange is a single NCCL
 * all-reduce and the optimizer a single fused pass).  Tensor ids, in buffer order (every tensor starts at a
 * multiple of 16 floats; names relative to `bbox_head.`):
 *   0 query_embedding.0.weight [256,384]   1 query_embedding.0.bias   2 query_embedding.2.weight [256,256]
 *   3 query_embedding.2.bias   4 transformer.decoder.post_norm.weight   5 .bias
 *   then for every layer l, 6 + 34 l + k with k =
 *    0 attentions.0.attn.in_proj_weight [768,256]  1 in_proj_bias  2 out_proj.weight [256,256]  3 out_proj.bias
 *    4 attentions.1.attn.in_proj_weight [768,256]  5 in_proj_bias  6 out_proj.weight            7 out_proj.bias
 *    8 ffns.0.layers.0.0.weight [2048,256]  9 .bias  10 ffns.0.layers.1.weight [256,2048]  11 .bias
 *    12..17 norms.{0,1,2}.{weight,bias}
 *    18 cls_branches.l.0.weight  19 .0.bias  20 .1.weight (LN)  21 .1.bias  22 .3.weight  23 .3.bias  24 .4.weight (LN)
 *    25 .4.bias  26 .6.weight [10,256]  27 .6.bias [10]
 *    28 reg_branches.l.0.weight  29 .0.bias  30 .2.weight  31 .2.bias  32 .4.weight [10,256]  33 .4.bias [10]      */
#define MV2D_TRAIN_GLOBAL_TENSORS 6
#define MV2D_TRAIN_LAYER_TENSORS 34
MV2D_API long long mv2d_train_param_total(int L);                      /* floats in the flat buffer */
/* Arithmetic of the training step's GPU-filling contractions (M >= 1024 rows: K/V projections, the 3x3 conv, the
 * position-encoding MLPs): 1 (default; env MV2D_TRAIN_TC=0 turns it off) = error-compensated 3xTF32 on the tcgen05
 * tensor cores, forward and backward; 0 = fp32 FFMA everywhere.  Both are fp32-grade, but 3xTF32 is ~16x coarser
 * (2^-20 vs 2^-24 of sum |a||b|), so a ReLU whose input sits within ~1e-5 of zero can land on the other side than in
 * the reference: the gradient is then exact for a forward that differs by that rounding.  Returns the previous mode. */
MV2D_API int mv2d_train_set_tensor_cores(int on);
MV2D_API int mv2d_train_param_info(int L, int tensor_id, long long* offset, long long* numel);

typedef struct Mv2dTrainParams {
    int N, L, max_match, G;
    int num_classes, reserved0;
    float pc_range[6];
    float cls_cost_weight, reg_cost_weight, cls_loss_weight, bbox_loss_weight, focal_alpha, focal_gamma;
    float code_weights[10];
    float stage_loss_weights[MV2D_MAX_LAYERS];   /* roi_head.stage_loss_weights (exp configs: 0.1 per layer ... 1.0 last) */
    const float* params;        /* flat parameters */
    float* grads;               /* flat gradients; the backward ACCUMULATES into it (the caller zeroes it per step) */
    const float* dim_t;         /* [128] temperature ** (2*(i//2)/128) */
    const float* ref;           /* [N,3] normalised reference points (mv2d_roi_align_qg) */
    const float* tok_kin;       /* [N,49,256] RoI-pooled (feat + pe): key input */
    const float* tok_mem;       /* [N,49,256] RoI-pooled feat: value input */
    const int* match;           /* [N,max_match] from mv2d_box_corr */
    const int* match_cnt;       /* [N] */
    const float* gt_boxes;      /* [G,9] gravity centre xyz, w, l, h, yaw, vx, vy */
    const int* gt_labels;       /* [G] */
    float* cls_scores;          /* out of the forward, in of the backward [L,N,10] */
    float* bbox_preds;          /* out / in [L,N,10] */
    int* assigned;              /* out / in [L,N] */
    float* losses;              /* out [L,4] as mv2d_loss (unweighted by stage_loss_weights) */
    float* d_ref;               /* backward out [N,3] */
    float* d_tok_kin;           /* backward out [N,49,256] */
    float* d_tok_mem;           /* backward out [N,49,256] */
    float* workspace;           /* saved activations of the forward + scratch of the backward; must survive between the
                                 * two calls */
    size_t workspace_bytes;
    /* ---- ABI 5: cross-rank loss normaliser (see Mv2dLossParams) */
    float* num_pos;             /* forward out, nullable [L] */
    const float* bbox_avg_factor; /* backward in, nullable [L]: the gradient of loss_bbox uses it instead of the local
                                 * max(num_pos, 1), and losses[l][1] is rescaled to it in place */
    /* ---- ABI 5: the two-frame head and denoising queries (roi_heads/mv2d_t_head.py:26-142, mv2d_s_head.py:39-120,
     * 278-299, bbox_heads/cross_attention_head.py:475-538) -- the configuration the reference trains MV2D-T with.
     * mode 1: the decoder runs over pad + N query rows (denoising rows first); `ref`, `d_ref`, cls_scores and
     * bbox_preds have pad + N rows ([L, pad+N, 10]); the cross-attention keys are the num_rows feature cells of
     * kin_map / mem_map, query i attends to key_list[i][0 .. key_cnt[i]) (= the set bits of keymask[i]); tok_kin /
     * tok_mem / match* / d_tok_* are not used.  The matching rows' velocity outputs are divided by vel_dt before
     * the loss (mv2d_t_head.py:130-142); the denoising rows add dn_loss_cls / dn_loss_bbox (losses[l][2..3]) times
     * denoise_weight.  `assigned`, `num_pos` and the Hungarian losses refer to the N matching rows. */
    int mode, pad, num_rows, mask_words;
    int neg_bbox_loss, reserved5;
    float vel_dt, dn_split, denoise_weight, reserved6;
    const float* kin_map;       /* [num_rows,256] key input (feat + pe), channels-last */
    const float* mem_map;       /* [num_rows,256] value input (feat) */
    const uint32_t* keymask;    /* [pad+N, mask_words] */
    const uint16_t* key_list;   /* [pad+N, mask_words*32] */
    const int* key_cnt;         /* [pad+N] */
    const uint8_t* self_attn_mask; /* nullable [pad+N, pad+N], 1 = masked */
    const int* dn_labels;       /* [pad]: label of denoising query i (num_classes = negative); its box is gt_boxes[i % G] */
    float* d_kin_map;           /* backward out [num_rows,256] */
    float* d_mem_map;           /* backward out [num_rows,256] */
} Mv2dTrainParams;
MV2D_API size_t mv2d_decoder_train_workspace_bytes(int N, int L, int max_match, int G);
/* the same from a filled parameter block (needed for mode 1: N, pad, L, num_rows, G are read) */
MV2D_API size_t mv2d_decoder_train_workspace_bytes_p(const Mv2dTrainParams* p);
MV2D_API int mv2d_decoder_train_forward(const Mv2dTrainParams* p, void* stream);
/* tests: float offset of a saved activation in `workspace` after the forward.  which = 0 query_pos [N,256];
 * per layer: 1 x after norms.0, 2 after norms.1, 3 layer output, 4 post-normed intermediate, 5 cross-attention
 * context, 6 self-attention context (all [N,256]).  -1 = unknown */
MV2D_API long long mv2d_train_debug_offset(int N, int L, int max_match, int G, int layer, int which);
MV2D_API int mv2d_decoder_train_backward(const Mv2dTrainParams* p, void* stream);

/* ---- row e, front end: training forward / backward of rows a1-a8 -- PE.forward (utils/pe.py:137-169), RoIAlign of the
 * feature and the position embedding (roi_heads/mv2d_s_head.py:133-138; mmcv RoIAlign), QueryGenerator.forward
 * (roi_heads/utils/query_generator.py:343-405) and the reference-point normalisation (mv2d_s_head.py:147-152).
 * The forward recomputes the stage in fp32 with saved activations and emits the inputs of mv2d_decoder_train_forward;
 * the backward consumes that call's input gradients and adds the parameter gradients to the same flat buffer
 * (tensor ids 6 + 34 L + k, names relative to roi_head: k = 0 position_encoding.position_encoder.0.weight [1024,192]
 * 1 .0.bias  2 .2.weight [256,1024]  3 .2.bias  4 position_encoding.adapt_pos3d.0.weight [1024,384]  5 .0.bias
 * 6 .2.weight [256,1024]  7 .2.bias  8 position_encoding.fpe.conv_reduce.weight [256,256]  9 .bias
 * 10 fpe.conv_expand.weight  11 .bias  12 query_generator.shared_convs.0.conv.weight stored [256, (ky,kx,c_in)]
 * 13 .bias  14 query_generator.shared_fcs.0.weight [1024,256]  15 .bias  16 extra_enc.0.weight [512,1040]  17 .bias
 * 18 extra_enc.2.weight [256,512]  19 .bias  20 fc_center.weight [3,256]  21 .bias [3]) and returns d loss / d feat. */
#define MV2D_TRAIN_FRONT_TENSORS 22
typedef struct Mv2dFrontTrainParams {
    int N, V, h, w, L, stride;
    int depth_num, pad_h, pad_w, reserved0;
    double depth_start;
    double position_range[6];
    float pc_range[6];
    float intrins_feat_scale, reserved1;
    const float* params;            /* flat parameters (layout of L decoder layers) */
    float* grads;                   /* flat gradients, accumulated */
    const float* rois;              /* [N,5] */
    const double* roi_intrinsics;   /* [N,16] K' from mv2d_roi_align_qg */
    const double* extrinsics;       /* [V,16] */
    const double* img2lidar;        /* [V,16] from mv2d_geom_prep */
    const uint8_t* not_mask;        /* [V,h,w] */
    const float* dim_t;             /* [128] */
    const float* feat;              /* [V,h,w,256] channels-last */
    float* tok_mem;                 /* forward out [N,49,256] RoI-pooled feat */
    float* tok_kin;                 /* forward out [N,49,256] RoI-pooled feat + pe */
    float* ref;                     /* forward out [N,3] */
    float* pe_out;                  /* forward out, nullable [V,h,w,256] */
    const float* d_ref;             /* backward in  (from mv2d_decoder_train_backward) */
    const float* d_tok_kin;         /* backward in */
    const float* d_tok_mem;         /* backward in */
    float* d_feat;                  /* backward out [V,h,w,256] */
    float* workspace;               /* saved activations + scratch; must survive between the two calls */
    size_t workspace_bytes;
    /* ---- ABI 5 (two-frame head: the decoder's keys are the whole maps feat + pe and feat) */
    const float* d_pe_extra;        /* backward in, nullable [V,h,w,256]: added to d loss / d pe before the PE backward */
    const float* d_feat_extra;      /* backward in, nullable [V,h,w,256]: added to d_feat */
    const float* d_feat_extra2;     /* backward in, nullable: a second map added to d_feat (value-input gradient) */
    float* kin_out;                 /* forward out, nullable [V,h,w,256]: feat + pe, the key input of the two-frame head */
} Mv2dFrontTrainParams;
MV2D_API size_t mv2d_front_train_workspace_bytes(int N, int V, int h, int w);
MV2D_API int mv2d_front_train_forward(const Mv2dFrontTrainParams* p, void* stream);
MV2D_API int mv2d_front_train_backward(const Mv2dFrontTrainParams* p, void* stream);

/* fused AdamW step over flat buffers (torch.optim.AdamW semantics; configs/mv2d/exp/*.py optimizer):
 * g is multiplied by grad_scale first (1 / world_size after a sum all-reduce); step counts from 1 */
MV2D_API int mv2d_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                             float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream);

/* ---- low-level GEMM, exposed for tests and microbenchmarks:
 * C[M,N] = act(A[M,K] . W[N,K]^T + bias); flags: 1 relu, 8 allow TF32 tensor cores,
 * 16 force the single-pass tcgen05 kernel (operands should be TF32-representable), 64 round C to TF32 */
MV2D_API int mv2d_gemm(const float* A, int lda, const float* W, int ldw, const float* bias, float* C, int ldc,
              int M, int N, int K, int flags, void* stream);

/* x -> hi = tf32(x), lo = tf32(x - hi): the operand split of the 3xTF32 GEMM */
MV2D_API int mv2d_split_tf32(const float* x, float* hi, float* lo, long long n, void* stream);
/* error-compensated 3xTF32 tcgen05 GEMM: C = act((A_hi+A_lo) . (W_hi+W_lo)^T + bias), fp32-grade.
 * With A_lo == W_lo == NULL the operands are plain fp32 and the kernel splits them in shared memory itself (one
 * copy of each operand crosses L2 instead of two); the result is bit-identical to the pre-split call.
 * flags: 1 relu, 128 A is the implicit 3x3 im2col of [M/49,7,7,256] RoI tokens (K = 2304; pre-split operands only) */
MV2D_API int mv2d_gemm_3xtf32(const float* A_hi, const float* A_lo, int lda, const float* W_hi, const float* W_lo,
                              int ldw, const float* bias, float* C, int ldc, int M, int N, int K, int flags,
                              void* stream);

/* ---- f1 (next row): NMSFreeCoder.decode_single + get_bboxes z-shift
 * (core/bbox/coders/nms_free_coder.py:49-102; bbox_heads/cross_attention_head.py:372).
 * Device top-k over N*10 sigmoid scores; writes max_num rows, valid[i] = inside post range. */
MV2D_API int mv2d_nms_free_decode(const float* cls /*[N,10]*/, const float* box /*[N,10]*/, int N, int max_num,
                         const float* post_range /*host [6]*/, float* out_boxes /*[max_num,9]*/,
                         float* out_scores /*[max_num]*/, int* out_labels /*[max_num]*/,
                         uint8_t* out_valid /*[max_num]*/, void* stream);

/* ---- f1, scene level: the mmdet3d box3d_multiclass_nms call of MV2D.simple_test (detectors/mv2d.py:266-282) on the
 * decoded boxes: keep valid[i] (nullable) && score > score_thr, regroup by class then descending score, cap at
 * max_num (by score, over all classes).  nms_thr must be >= 1.0 (the configs' value: rotated BEV NMS at IoU 1.0
 * suppresses nothing); lower thresholds return an error. */
MV2D_API int mv2d_scene_nms(const float* boxes /*[n,9]*/, const float* scores /*[n]*/, const int* labels /*[n]*/,
                            const uint8_t* valid /*[n] nullable*/, int n, float score_thr, float nms_thr, int max_num,
                            float* out_boxes /*[max_num,9]*/, float* out_scores /*[max_num]*/, int* out_labels /*[max_num]*/,
                            int* out_count /*device [1]*/, void* stream);

/* debug: spin `cycles` SM clocks in a 1-warp kernel; out[0] = elapsed ns (globaltimer), out[1] = cycles */
MV2D_API int mv2d_debug_clock_probe(long long cycles, long long* out /*device [2]*/, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MV2D_B200_H_ */
