// Row f4 of SURVEY.md 8f: the MV2D neck.  The reference builds an mmdet FPN with one level
// (configs/mv2d/exp/*.py:32-39: in_channels [256]*5, start_level = end_level = 2, num_outs = 1) and calls it in
// detectors/mv2d.py:122-127 (process_detector_feat) on the 2D detector's FPN outputs: with a single level there is
// no top-down path, so the neck is  P4' = conv3x3(conv1x1(P4) + b_lat) + b_fpn  (mmdet FPN.forward, no norm / act).
// Here: layout change to channels-last, the 1x1 as a 3xTF32 tcgen05 GEMM whose epilogue writes the TF32 hi / lo
// operands of the next GEMM, the 3x3 as the same GEMM with the A operand gathered by 4-D TMA boxes
// (32 ch, 8 x, 16 y, 1 view) with out-of-bounds zero fill = implicit im2col with padding 1.  The output is the
// channels-last feature map the rest of the path consumes (no separate transpose), plus its TF32-rounded copy.
#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "mv2d_internal.h"

namespace mv2d {

size_t fpn_neck_workspace_bytes(int V, int h, int w) {
    const size_t P = (size_t)(V > 0 ? V : 1) * (size_t)(h > 0 ? h : 1) * (size_t)(w > 0 ? w : 1);
    return 5 * P * MV2D_C * sizeof(float) + 256;      // x (nhwc), x_hi, x_lo, lat_hi, lat_lo
}

int run_fpn_neck(const Mv2dNeckParams& p, cudaStream_t st) {
    MV2D_CHECK_ARG(p.V >= 1 && p.V <= MV2D_MAXV && p.h > 0 && p.w > 0, "fpn_neck: bad V/h/w");
    const long long P = (long long)p.V * p.h * p.w;
    MV2D_CHECK_ARG(p.workspace && p.workspace_bytes >= fpn_neck_workspace_bytes(p.V, p.h, p.w), "fpn_neck: workspace too small");
    float* ws = p.workspace;
    float* x = ws;      ws += P * MV2D_C;
    float* x_hi = ws;   ws += P * MV2D_C;
    float* x_lo = ws;   ws += P * MV2D_C;
    float* l_hi = ws;   ws += P * MV2D_C;
    float* l_lo = ws;   ws += P * MV2D_C;
    int rc;
    const float* xin = p.x;
    if (!p.in_is_nhwc) {
        if ((rc = run_nchw_to_nhwc(p.x, x, nullptr, p.V, MV2D_C, p.h * p.w, st))) return rc;
        xin = x;
    }
    if ((rc = launch_split_tf32(xin, x_hi, x_lo, P * MV2D_C, st))) return rc;
    {   // lateral 1x1: [P,256] . [256,256]^T + b, result straight into the hi / lo operands of the 3x3
        TcGemm t{};
        t.A = x_hi; t.A_lo = x_lo; t.lda = MV2D_C; t.W = p.lat_w; t.W_lo = p.lat_w_lo; t.ldw = MV2D_C; t.bias = p.lat_b;
        t.C = l_hi; t.C_lo = l_lo; t.ldc = MV2D_C; t.M = (int)P; t.N = MV2D_C; t.K = MV2D_C; t.passes = 3; t.nsplit = 1;
        t.flags = GEMM_SPLIT_OUT;
        if ((rc = launch_gemm_tc(t, st))) return rc;
    }
    {   // fpn 3x3, padding 1: implicit im2col over the feature map
        TcGemm t{};
        t.A = l_hi; t.A_lo = l_lo; t.lda = MV2D_C; t.W = p.fpn_w; t.W_lo = p.fpn_w_lo; t.ldw = 9 * MV2D_C; t.bias = p.fpn_b;
        t.C = p.feat; t.ldc = MV2D_C; t.M = (int)P; t.N = MV2D_C; t.K = 9 * MV2D_C; t.passes = 3; t.nsplit = 1;
        t.im2col = 2; t.fm_v = p.V; t.fm_h = p.h; t.fm_w = p.w;
        if ((rc = launch_gemm_tc(t, st))) return rc;
    }
    if (p.feat_tf32) {      // operand of the single-pass SE-gate GEMM of the position embedding (x_lo is free: scratch)
        if ((rc = launch_split_tf32(p.feat, p.feat_tf32, x_lo, P * MV2D_C, st))) return rc;
    }
    return 0;
}

}  // namespace mv2d
