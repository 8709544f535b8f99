// K2 (RoIAlign + query generator) and K3 (box correlation).  This translation unit is compiled
// with -fmad=false: the fp32/fp64 geometry below feeds discontinuous decisions (hit tests,
// IoU>0, top-k) and must round exactly like the reference's un-fused torch ops.
//   reference: roi_heads/mv2d_head.py:51-72,95-101; roi_heads/utils/query_generator.py:333-405;
//              roi_heads/utils/box_correlation.py:95-398; mmcv RoIAlign
#include "common.cuh"
#include "keylist.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "mv2d_internal.h"

namespace mv2d {

// ------------------------------------------------------------------------------------------
// RoIAlign (avg, aligned=True, adaptive sampling grid) on channels-last maps.
// grid (N, 1 or 4), 256 threads: one CTA per RoI (four for a single sample's few hundred RoIs), warp = bins {w, w + 8, ..}, lane = eight channels {4 lane .. 4 lane + 3,
// 128 + 4 lane .. + 3} => a warp reads two 512-byte runs per bilinear corner and the per-point geometry (the larger part of
// the instruction stream: the kernel is bound by instruction issue) is shared by eight channels instead of four.
// Writes tok_feat and (optionally) tok_kin = feat + pe tokens (RoIAlign is linear, so pooling feat+pe equals pooling them
// separately) and the TF32 hi / lo split of tok_feat.
// Channel pairs as packed f32x2; roi.cu is compiled with -fmad=false for the box geometry, so these FMAs are spelled out.
__global__ void __launch_bounds__(256)
roi_align_tokens_kernel(const float* __restrict__ rois, const float* __restrict__ feat,
                        const float* __restrict__ pe, int h, int w, float spatial_scale,
                        float* __restrict__ tok_feat, float* __restrict__ tok_kin,
                        float* __restrict__ tok_hi, float* __restrict__ tok_lo) {
    // feat == nullptr: second phase -- pool only pe and add the already pooled tok_feat (tok_kin = tok_feat + pool(pe))
    pdl_wait();
    pdl_trigger();
    const int n = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* r = rois + n * 5;
    const int v = (int)r[0];
    const float x1 = r[1] * spatial_scale - 0.5f, y1 = r[2] * spatial_scale - 0.5f;
    const float x2 = r[3] * spatial_scale - 0.5f, y2 = r[4] * spatial_scale - 0.5f;
    const float rw = x2 - x1, rh = y2 - y1;
    const float bw = rw / (float)MV2D_ROI, bh = rh / (float)MV2D_ROI;
    const int gh = (int)ceilf(rh / (float)MV2D_ROI), gw = (int)ceilf(rw / (float)MV2D_ROI);
    const float count = (float)max(gh * gw, 1);
    const ulonglong2* f4 = feat ? reinterpret_cast<const ulonglong2*>(feat) + (long long)v * h * w * 64 + lane : nullptr;
    const ulonglong2* p4 = pe ? reinterpret_cast<const ulonglong2*>(pe) + (long long)v * h * w * 64 + lane : nullptr;
    for (int bin = blockIdx.y * 8 + warp; bin < MV2D_TOK; bin += 8 * gridDim.y) {     // gridDim.y > 1: few RoIs (one sample)
        const int ph = bin / MV2D_ROI, pw = bin % MV2D_ROI;
        f32x2 af0 = 0ull, af1 = 0ull, af2 = 0ull, af3 = 0ull, ap0 = 0ull, ap1 = 0ull, ap2 = 0ull, ap3 = 0ull;
        for (int iy = 0; iy < gh; ++iy) {
            float y = y1 + (float)ph * bh + ((float)iy + 0.5f) * bh / (float)gh;
            for (int ix = 0; ix < gw; ++ix) {
                float x = x1 + (float)pw * bw + ((float)ix + 0.5f) * bw / (float)gw;
                if (y < -1.f || y > (float)h || x < -1.f || x > (float)w) continue;
                float yy = fmaxf(y, 0.f), xx = fmaxf(x, 0.f);
                int yl = (int)yy, xl = (int)xx, yh, xh;
                if (yl >= h - 1) { yh = yl = h - 1; yy = (float)yl; } else yh = yl + 1;
                if (xl >= w - 1) { xh = xl = w - 1; xx = (float)xl; } else xh = xl + 1;
                const float ly = yy - (float)yl, lx = xx - (float)xl, hy = 1.f - ly, hx = 1.f - lx;
                const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
                const f32x2 q1 = pack2(w1, w1), q2 = pack2(w2, w2), q3 = pack2(w3, w3), q4 = pack2(w4, w4);
                const int o1 = (yl * w + xl) * 64, o2 = (yl * w + xh) * 64, o3 = (yh * w + xl) * 64,
                          o4 = (yh * w + xh) * 64;
                if (f4) {
                    const ulonglong2 a = __ldg(f4 + o1), b = __ldg(f4 + o2), c = __ldg(f4 + o3), d = __ldg(f4 + o4);
                    const ulonglong2 a2 = __ldg(f4 + o1 + 32), b2 = __ldg(f4 + o2 + 32), c2 = __ldg(f4 + o3 + 32), d2 = __ldg(f4 + o4 + 32);
                    af0 = fma2(q1, a.x, af0); af1 = fma2(q1, a.y, af1); af2 = fma2(q1, a2.x, af2); af3 = fma2(q1, a2.y, af3);
                    af0 = fma2(q2, b.x, af0); af1 = fma2(q2, b.y, af1); af2 = fma2(q2, b2.x, af2); af3 = fma2(q2, b2.y, af3);
                    af0 = fma2(q3, c.x, af0); af1 = fma2(q3, c.y, af1); af2 = fma2(q3, c2.x, af2); af3 = fma2(q3, c2.y, af3);
                    af0 = fma2(q4, d.x, af0); af1 = fma2(q4, d.y, af1); af2 = fma2(q4, d2.x, af2); af3 = fma2(q4, d2.y, af3);
                }
                if (p4) {
                    const ulonglong2 a = __ldg(p4 + o1), b = __ldg(p4 + o2), c = __ldg(p4 + o3), d = __ldg(p4 + o4);
                    const ulonglong2 a2 = __ldg(p4 + o1 + 32), b2 = __ldg(p4 + o2 + 32), c2 = __ldg(p4 + o3 + 32), d2 = __ldg(p4 + o4 + 32);
                    ap0 = fma2(q1, a.x, ap0); ap1 = fma2(q1, a.y, ap1); ap2 = fma2(q1, a2.x, ap2); ap3 = fma2(q1, a2.y, ap3);
                    ap0 = fma2(q2, b.x, ap0); ap1 = fma2(q2, b.y, ap1); ap2 = fma2(q2, b2.x, ap2); ap3 = fma2(q2, b2.y, ap3);
                    ap0 = fma2(q3, c.x, ap0); ap1 = fma2(q3, c.y, ap1); ap2 = fma2(q3, c2.x, ap2); ap3 = fma2(q3, c2.y, ap3);
                    ap0 = fma2(q4, d.x, ap0); ap1 = fma2(q4, d.y, ap1); ap2 = fma2(q4, d2.x, ap2); ap3 = fma2(q4, d2.y, ap3);
                }
            }
        }
        const f32x2 accf[2][2] = {{af0, af1}, {af2, af3}}, accp[2][2] = {{ap0, ap1}, {ap2, ap3}};
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float4 af, ap;
            unpack2(accf[half][0], af.x, af.y); unpack2(accf[half][1], af.z, af.w);
            unpack2(accp[half][0], ap.x, ap.y); unpack2(accp[half][1], ap.z, ap.w);
            const long long o = ((long long)n * MV2D_TOK + bin) * 64 + half * 32 + lane;
            if (!f4) af = reinterpret_cast<const float4*>(tok_feat)[o];   // phase 2: pooled feature from phase 1
            else { af.x /= count; af.y /= count; af.z /= count; af.w /= count; }
            if (f4) reinterpret_cast<float4*>(tok_feat)[o] = af;
            if (f4) {   // TF32 hi/lo split of the pooled feature: operands of the 3xTF32 conv GEMM
                float4 hi = make_float4(round_tf32(af.x), round_tf32(af.y), round_tf32(af.z), round_tf32(af.w));
                reinterpret_cast<float4*>(tok_hi)[o] = hi;
                reinterpret_cast<float4*>(tok_lo)[o] = make_float4(round_tf32(af.x - hi.x), round_tf32(af.y - hi.y),
                                                                  round_tf32(af.z - hi.z), round_tf32(af.w - hi.w));
            }
            if (tok_kin) {
                ap.x /= count; ap.y /= count; ap.z /= count; ap.w /= count;
                reinterpret_cast<float4*>(tok_kin)[o] = make_float4(af.x + ap.x, af.y + ap.y, af.z + ap.z, af.w + ap.w);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Per-RoI camera parameters (mv2d_head.py:51-72,95-101 + query_generator.py:338-339):
// K' = RoI-frame intrinsics (fp64), feature = clamp(float(K')*scale), M = float(inv(K' E^T)).
__global__ void box_params_kernel(const float* __restrict__ rois, const double* __restrict__ intrinsics,
                                  const double* __restrict__ extrinsics, int N, float feat_scale,
                                  float* __restrict__ cat /*[N,MV2D_CAT_LD]*/, float* __restrict__ m_roi /*[N,16]*/,
                                  double* __restrict__ k_out /*nullable [N,16]*/, float* __restrict__ cat_lo /*nullable: cat = TF32 hi part*/) {
    pdl_wait();
    pdl_trigger();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float* r = rois + n * 5;
    const int v = (int)r[0];
    double K[16], E[16], L[16], Li[16];
    for (int i = 0; i < 16; ++i) { K[i] = intrinsics[v * 16 + i]; E[i] = extrinsics[v * 16 + i]; }
    const float wx = r[3] - r[1], wy = r[4] - r[2];
    const float sx = 7.0f / wx, sy = 7.0f / wy;
    K[2] = K[2] - (double)r[1] - (double)(0.5f / sx);
    K[6] = K[6] - (double)r[2] - (double)(0.5f / sy);
    for (int j = 0; j < 4; ++j) { K[j] = K[j] * (double)sx; K[4 + j] = K[4 + j] * (double)sy; }
    const bool invalid = (wx < 4.f) || (wy < 4.f);
    for (int i = 0; i < 16; ++i) {
        float f = invalid ? 0.f : (float)K[i] * feat_scale;
        f = fminf(fmaxf(f, -5e3f), 5e3f);
        if (cat_lo) {
            const float hi = round_tf32(f);
            cat_lo[(long long)n * MV2D_CAT_LD + 1024 + i] = round_tf32(f - hi);
            cat_lo[(long long)n * MV2D_CAT_LD + 1040 + i] = 0.f;
            f = hi;
        }
        cat[(long long)n * MV2D_CAT_LD + 1024 + i] = f;
        cat[(long long)n * MV2D_CAT_LD + 1040 + i] = 0.f;   // zero K-padding (weights are zero-padded too)
        if (k_out) k_out[n * 16 + i] = K[i];
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s += K[i * 4 + k] * E[j * 4 + k];
            L[i * 4 + j] = s;
        }
    inv4x4(L, Li);
    for (int i = 0; i < 16; ++i) m_roi[n * 16 + i] = (float)Li[i];
}

// Stage-level QueryGenerator.forward (phase 3): the per-RoI camera inputs are given instead of derived from the boxes:
// K' [N,16] fp64 (get_box_params), per-RoI extrinsics [N,16] fp64 and the intrinsics feature [N,16] (extra_feats).
__global__ void qg_inputs_kernel(const double* __restrict__ k_roi, const double* __restrict__ e_roi, const float* __restrict__ ifeat,
                                 int N, float* __restrict__ cat, float* __restrict__ m_roi, float* __restrict__ cat_lo) {
    pdl_wait();
    pdl_trigger();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    double K[16], E[16], L[16], Li[16];
    for (int i = 0; i < 16; ++i) { K[i] = k_roi[n * 16 + i]; E[i] = e_roi[n * 16 + i]; }
    for (int i = 0; i < 16; ++i) {
        float f = fminf(fmaxf(ifeat[n * 16 + i], -5e3f), 5e3f);
        if (cat_lo) {
            const float hi = round_tf32(f);
            cat_lo[(long long)n * MV2D_CAT_LD + 1024 + i] = round_tf32(f - hi);
            cat_lo[(long long)n * MV2D_CAT_LD + 1040 + i] = 0.f;
            f = hi;
        }
        cat[(long long)n * MV2D_CAT_LD + 1024 + i] = f;
        cat[(long long)n * MV2D_CAT_LD + 1040 + i] = 0.f;
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s += K[i * 4 + k] * E[j * 4 + k];
            L[i * 4 + j] = s;
        }
    inv4x4(L, Li);
    for (int i = 0; i < 16; ++i) m_roi[n * 16 + i] = (float)Li[i];
}

// AvgPool2d(7) over the ReLU'd conv output: [N,49,256] -> [N,256]
__global__ void avgpool49_kernel(const float* __restrict__ x, float* __restrict__ out, int N, float* __restrict__ out_lo) {
    pdl_wait();
    pdl_trigger();
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= N * MV2D_C) return;
    const int n = gid / MV2D_C, c = gid % MV2D_C;
    float s = 0.f;
    for (int t = 0; t < MV2D_TOK; ++t) s += x[((long long)n * MV2D_TOK + t) * MV2D_C + c];
    s = s / 49.0f;
    if (out_lo) { const float hi = round_tf32(s); out_lo[gid] = round_tf32(s - hi); s = hi; }   // 3xTF32 operand split
    out[gid] = s;
}

// fc_center + center2lidar + normalisation + pos2posemb3d  (query_generator.py:333-341,400-403;
// mv2d_s_head.py:147-152; utils/pe.py:21-33).  One warp per RoI.
__global__ void qg_tail_kernel(const float* __restrict__ enc, const float* __restrict__ w_center,
                               const float* __restrict__ b_center, const float* __restrict__ m_roi,
                               const float* __restrict__ dim_t, int N, float pc0, float pc1, float pc2,
                               float pc3, float pc4, float pc5, float* __restrict__ ref,
                               float* __restrict__ center_lidar, float* __restrict__ posemb, float* __restrict__ posemb_lo) {
    pdl_wait();
    pdl_trigger();
    const int n = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x & 31;
    if (n >= N) return;
    float c[3];
#pragma unroll
    for (int o = 0; o < 3; ++o) {
        float s = 0.f;
        for (int k = lane; k < MV2D_C; k += 32) s = fmaf(enc[(long long)n * MV2D_C + k], __ldg(w_center + o * MV2D_C + k), s);
        c[o] = warp_sum(s) + __ldg(b_center + o);
    }
    const float hom[4] = {c[0] * c[2], c[1] * c[2], c[2], 1.0f};
    const float* m = m_roi + n * 16;
    float xyz[3], p[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
        xyz[i] = m[i * 4 + 0] * hom[0] + m[i * 4 + 1] * hom[1] + m[i * 4 + 2] * hom[2] + m[i * 4 + 3] * hom[3];
    p[0] = (xyz[0] - pc0) / (pc3 - pc0);
    p[1] = (xyz[1] - pc1) / (pc4 - pc1);
    p[2] = (xyz[2] - pc2) / (pc5 - pc2);
    if (lane < 3) {
        ref[n * 3 + lane] = p[lane];
        if (center_lidar) center_lidar[n * 3 + lane] = xyz[lane];
    }
    // posemb = cat(emb(y), emb(x), emb(z)), interleaved sin/cos
    for (int idx = lane; idx < 384; idx += 32) {
        const int part = idx >> 7, i = idx & 127;
        const float pos = (part == 0 ? p[1] : (part == 1 ? p[0] : p[2])) * 6.283185307179586f;
        const float a = pos / __ldg(dim_t + i);
        float e = (i & 1) ? cosf(a) : sinf(a);
        if (posemb_lo) { const float hi = round_tf32(e); posemb_lo[(long long)n * 384 + idx] = round_tf32(e - hi); e = hi; }
        posemb[(long long)n * 384 + idx] = e;
    }
}

static int gemm(const float* A, int lda, const float* W, int ldw, const float* bias, float* C, int ldc,
                int M, int N, int K, int flags, int amode, cudaStream_t st) {
    GemmArgs g{};
    g.A = A; g.lda = lda; g.W = W; g.ldw = ldw; g.C = C; g.ldc = ldc; g.bias = bias;
    g.M = M; g.N = N; g.K = K; g.batch = 1; g.nsplit = 1; g.flags = flags;
    return launch_gemm_simt(g, amode, st);
}

size_t roi_align_qg_workspace_bytes(int N) {
    size_t n = (size_t)(N > 0 ? N : 1);
    // + the lo halves of pool, cat, e0, posemb, qh (3xTF32 FC chain of a batch)
    return n * ((size_t)3 * MV2D_TOK * MV2D_C + MV2D_C + MV2D_CAT_LD + 512 + MV2D_C + 16 + 384 + MV2D_C +
                MV2D_C + MV2D_CAT_LD + 512 + 384 + MV2D_C) * sizeof(float);
}

int run_roi_align_qg(const Mv2dQgParams& p, cudaStream_t st) {
    const int N = p.N, C = MV2D_C;
    MV2D_CHECK_ARG(N >= 0 && p.V >= 1 && p.V <= MV2D_MAXVB, "roi_align_qg: bad N/V");
    if (N == 0) return 0;
    MV2D_CHECK_ARG(p.phase >= 0 && p.phase <= 4, "roi_align_qg: phase must be 0 .. 4");
    MV2D_CHECK_ARG(p.phase == 1 || p.phase >= 3 || p.tok_kin == nullptr || p.pe != nullptr, "roi_align_qg: tok_kin needs pe");
    MV2D_CHECK_ARG(p.phase != 3 || (p.roi_intrinsics && p.roi_extrinsics && p.intrins_feat), "roi_align_qg: phase 3 needs K', E and the intrinsics feature per RoI");
    if (p.phase == 2) {   // only the position-embedding tokens: tok_kin = tok_feat + RoIAlign(pe)
        if (p.tok_kin == nullptr) return 0;
        launch_k(roi_align_tokens_kernel, dim3(N, N < 1024 ? 4 : 1), dim3(256), 0, st, p.rois, (const float*)nullptr, p.pe, p.h, p.w,
                 1.0f / (float)p.stride, p.tok_feat, p.tok_kin, (float*)nullptr, (float*)nullptr);
        MV2D_CHECK_LAUNCH("roi_align_tokens(pe)");
        return 0;
    }
    const bool with_pe = p.phase == 0 && p.tok_kin != nullptr;
    float* ws = p.workspace;
    float* conv = ws;   ws += (size_t)N * MV2D_TOK * C;
    float* thi = ws;    ws += (size_t)N * MV2D_TOK * C;
    float* tlo = ws;    ws += (size_t)N * MV2D_TOK * C;
    float* pool = ws;   ws += (size_t)N * C;
    float* cat = ws;    ws += (size_t)N * MV2D_CAT_LD;
    float* e0 = ws;     ws += (size_t)N * 512;
    float* enc = ws;    ws += (size_t)N * C;
    float* mroi = ws;   ws += (size_t)N * 16;
    float* pemb = ws;   ws += (size_t)N * 384;
    float* qh = ws;     ws += (size_t)N * C;
    float* pool_lo = ws; ws += (size_t)N * C;
    float* cat_lo = ws;  ws += (size_t)N * MV2D_CAT_LD;
    float* e0_lo = ws;   ws += (size_t)N * 512;
    float* pemb_lo = ws; ws += (size_t)N * 384;
    float* qh_lo = ws;   ws += (size_t)N * C;
    MV2D_CHECK_ARG((size_t)(ws - p.workspace) * sizeof(float) <= p.workspace_bytes, "roi_align_qg: workspace too small");
    // more than 512 RoIs (a batch): the FC chain is GPU-filling, it runs as 3xTF32 tcgen05 GEMMs on split operands
    static const bool tc_fc_env = []() { const char* v = getenv("MV2D_QG_TC_FC"); return !(v && v[0] == '0'); }();
    const bool big = tc_fc_env && N > 512 && p.w_fc_hi && p.w_fc_lo && p.w_enc0_hi && p.w_enc0_lo && p.w_enc2_hi && p.w_enc2_lo &&
                     p.w_qe0_hi && p.w_qe0_lo && p.w_qe2_hi && p.w_qe2_lo;
    auto tc3 = [&](const float* a_hi, const float* a_lo, int lda, const float* w_hi, const float* w_lo, int ldw, const float* bias,
                   float* c, float* c_lo, int ldc, int n_out, int k, int flags) {
        TcGemm t{};
        t.A = a_hi; t.A_lo = a_lo; t.lda = lda; t.W = w_hi; t.W_lo = w_lo; t.ldw = ldw; t.bias = bias; t.C = c; t.C_lo = c_lo; t.ldc = ldc;
        t.M = N; t.N = n_out; t.K = k; t.passes = 3; t.nsplit = 1; t.flags = flags;
        return launch_gemm_tc(t, st);
    };
    int rc;
    if (p.phase == 4) {     // per-RoI intrinsics K' only (weight independent: what the training forward needs from this entry)
        MV2D_CHECK_ARG(p.roi_intrinsics != nullptr, "roi_align_qg: phase 4 writes roi_intrinsics");
        launch_k(box_params_kernel, dim3(cdiv(N, 64)), dim3(64), 0, st, p.rois, p.intrinsics, p.extrinsics, N, p.intrins_feat_scale, cat,
                                                      mroi, p.roi_intrinsics, (float*)nullptr);
        MV2D_CHECK_LAUNCH("box_params");
        return 0;
    }
    if (p.phase == 3) {     // tokens and per-RoI camera parameters are inputs (QueryGenerator.forward on its own)
        if ((rc = launch_split_tf32(p.tok_feat, thi, tlo, (long long)N * MV2D_TOK * C, st))) return rc;
        launch_k(qg_inputs_kernel, dim3(cdiv(N, 64)), dim3(64), 0, st, (const double*)p.roi_intrinsics, p.roi_extrinsics, p.intrins_feat, N, cat,
                 mroi, big ? cat_lo : (float*)nullptr);
        MV2D_CHECK_LAUNCH("qg_inputs");
    } else {
    launch_k(roi_align_tokens_kernel, dim3(N, N < 1024 ? 4 : 1), dim3(256), 0, st, p.rois, p.feat, with_pe ? p.pe : (const float*)nullptr, p.h, p.w,
             1.0f / (float)p.stride, p.tok_feat, with_pe ? p.tok_kin : (float*)nullptr, thi, tlo);
    MV2D_CHECK_LAUNCH("roi_align_tokens");
    launch_k(box_params_kernel, dim3(cdiv(N, 64)), dim3(64), 0, st, p.rois, p.intrinsics, p.extrinsics, N, p.intrins_feat_scale, cat,
                                                  mroi, p.roi_intrinsics, big ? cat_lo : (float*)nullptr);
    MV2D_CHECK_LAUNCH("box_params");
    }
    // shared conv 3x3 (+ReLU) as implicit GEMM over the tokens, avg-pool, FC chain
    {   // tcgen05 3xTF32, A staged by 4-D TMA boxes straight from the token tensor
        TcGemm t{};
        t.A = thi; t.A_lo = tlo; t.lda = C; t.W = p.w_conv; t.W_lo = p.w_conv_lo; t.ldw = 9 * C; t.bias = p.b_conv;
        t.C = conv; t.ldc = C; t.M = N * MV2D_TOK; t.N = C; t.K = 9 * C; t.passes = 3; t.im2col = 1; t.flags = GEMM_RELU;
        if ((rc = launch_gemm_tc(t, st))) return rc;
    }
    launch_k(avgpool49_kernel, dim3(cdiv(N * C, 256)), dim3(256), 0, st, (const float*)conv, pool, N, big ? pool_lo : (float*)nullptr);
    MV2D_CHECK_LAUNCH("avgpool49");
    if (big) {
        if ((rc = tc3(pool, pool_lo, C, p.w_fc_hi, p.w_fc_lo, C, p.b_fc, cat, cat_lo, MV2D_CAT_LD, 1024, C,
                      GEMM_RELU | GEMM_CLAMP5E3 | GEMM_SPLIT_OUT))) return rc;
        if ((rc = tc3(cat, cat_lo, MV2D_CAT_LD, p.w_enc0_hi, p.w_enc0_lo, MV2D_CAT_LD, p.b_enc0, e0, e0_lo, 512, 512, MV2D_CAT_LD,
                      GEMM_RELU | GEMM_SPLIT_OUT))) return rc;
        if ((rc = tc3(e0, e0_lo, 512, p.w_enc2_hi, p.w_enc2_lo, 512, p.b_enc2, enc, nullptr, C, C, 512, GEMM_RELU))) return rc;
    } else {
        if ((rc = gemm(pool, C, p.w_fc, C, p.b_fc, cat, MV2D_CAT_LD, N, 1024, C, GEMM_RELU | GEMM_CLAMP5E3, A_PLAIN, st))) return rc;
        if ((rc = gemm(cat, MV2D_CAT_LD, p.w_enc0, MV2D_CAT_LD, p.b_enc0, e0, 512, N, 512, MV2D_CAT_LD, GEMM_RELU, A_PLAIN, st))) return rc;
        if ((rc = gemm(e0, 512, p.w_enc2, 512, p.b_enc2, enc, C, N, C, 512, GEMM_RELU, A_PLAIN, st))) return rc;
    }
    if (p.enc_out) {
        cudaError_t e = cudaMemcpyAsync(p.enc_out, enc, (size_t)N * C * sizeof(float), cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) { set_error("roi_align_qg: enc copy %s", cudaGetErrorString(e)); return (int)e; }
    }
    launch_k(qg_tail_kernel, dim3(cdiv(N, 4)), dim3(128), 0, st, (const float*)enc, p.w_center, p.b_center, mroi, p.dim_t, N, p.pc_range[0],
                                               p.pc_range[1], p.pc_range[2], p.pc_range[3], p.pc_range[4],
                                               p.pc_range[5], p.ref, p.center_lidar, pemb, big ? pemb_lo : (float*)nullptr);
    MV2D_CHECK_LAUNCH("qg_tail");
    // query_embedding: 384 -> 256 (ReLU) -> 256
    if (big) {
        if ((rc = tc3(pemb, pemb_lo, 384, p.w_qe0_hi, p.w_qe0_lo, 384, p.b_qe0, qh, qh_lo, C, C, 384, GEMM_RELU | GEMM_SPLIT_OUT))) return rc;
        if ((rc = tc3(qh, qh_lo, C, p.w_qe2_hi, p.w_qe2_lo, C, p.b_qe2, p.query_pos, nullptr, C, C, C, 0))) return rc;
    } else {
        if ((rc = gemm(pemb, 384, p.w_qe0, 384, p.b_qe0, qh, C, N, C, 384, GEMM_RELU, A_PLAIN, st))) return rc;
        if ((rc = gemm(qh, C, p.w_qe2, C, p.b_qe2, p.query_pos, C, N, C, C, 0, A_PLAIN, st))) return rc;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// K3 box correlation.  One CTA (128 threads = 16 sample points x 8 depths) per RoI.
// For every other view: project the samples (fp64), test them against that view's RoIs,
// and when any valid sample lands in any RoI, IoU-match the samples' bounding box against the
// view's RoIs and keep the top-k with IoU > 0  (box_correlation.py:260-382).
__device__ inline float block_min128(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = fminf(fminf(red[0], red[1]), fminf(red[2], red[3]));
    __syncthreads();
    return r;
}

#define CORR_MAXR 256  // max RoIs per view handled by the smem IoU table

__global__ void __launch_bounds__(128)
box_corr_kernel(Mv2dCorrParams p) {
    pdl_wait();
    pdl_trigger();
    __shared__ float red[4];
    __shared__ float iou_s[CORR_MAXR];
    __shared__ int cnt_s, kept_s;
    const int n = blockIdx.x, t = threadIdx.x;
    const float* r = p.rois + n * 5;
    const int src = (int)r[0];
    const int S = p.sample_size, Dn = p.num_depth;
    const int pt = t / Dn, d = t % Dn;          // host guarantees S*S*Dn == 128
    const int iy = pt / S, ix = pt % S;
    const float wx = r[3] - r[1], wy = r[4] - r[2];
    const float px = r[1] + wx * __ldg(p.lin + ix);
    const float py = r[2] + wy * __ldg(p.lin + iy);
    const double dep = (double)__ldg(p.depths + d);
    const double hx = (double)px * dep, hy = (double)py * dep;
    int* match = p.match + (long long)n * p.max_match;
    if (t == 0) { match[0] = n; cnt_s = 1; }
    __syncthreads();
    // batch: the RoI's sample b owns views [b*V, (b+1)*V); roi_start is [batch, V+1], trans [batch, V, V, 16]
    const int b = p.batch > 0 ? src / p.V : 0;
    const int src_l = src - b * p.V;
    const int* roi_start = p.roi_start + b * (p.V + 1);
    const double* trans = p.trans + (long long)b * p.V * p.V * 16;
    for (int v = 0; v < p.V; ++v) {
        const int rs = roi_start[v], nv = roi_start[v + 1] - rs;
        if (v == src_l || nv == 0) continue;      // block-uniform
        const double* T = trans + ((long long)src_l * p.V + v) * 16;
        const double cx = T[0] * hx + T[1] * hy + T[2] * dep + T[3];
        const double cy = T[4] * hx + T[5] * hy + T[6] * dep + T[7];
        const double cz = T[8] * hx + T[9] * hy + T[10] * dep + T[11];
        const double zc = fmax(cz, 1e-2);
        const double tx = cx / zc, ty = cy / zc;
        bool valid = !(cz < (double)p.depth_start);
        valid = valid && (0.0 <= tx) && (tx <= (double)(p.img_w - 1)) && (0.0 <= ty) && (ty <= (double)(p.img_h - 1));
        const float fx = (float)tx, fy = (float)ty;
        bool hit = false;
        if (valid) {
            for (int j = 0; j < nv; ++j) {
                const float* q = p.rois + (rs + j) * 5;
                hit = hit || (q[1] <= fx && fx <= q[3] && q[2] <= fy && fy <= q[4]);
            }
        }
        if (!__syncthreads_or(hit ? 1 : 0)) continue;
        // bounding box of the valid samples in view v (1e4 / -1e4 sentinels as the reference)
        const float xmin = block_min128(valid ? fx : 1e4f, red);
        const float ymin = block_min128(valid ? fy : 1e4f, red);
        const float xmax = -block_min128(valid ? -fx : 1e4f, red);
        const float ymax = -block_min128(valid ? -fy : 1e4f, red);
        const float area_a = (xmax - xmin) * (ymax - ymin);
        if (t == 0) kept_s = 0;
        for (int j = t; j < nv && j < CORR_MAXR; j += 128) {
            const float* q = p.rois + (rs + j) * 5;
            const float iw = fmaxf(fminf(xmax, q[3]) - fmaxf(xmin, q[1]), 0.f);
            const float ih = fmaxf(fminf(ymax, q[4]) - fmaxf(ymin, q[2]), 0.f);
            const float inter = iw * ih;
            const float area_b = (q[3] - q[1]) * (q[4] - q[2]);
            const float uni = area_a + area_b - inter;
            iou_s[j] = inter / (uni + 1e-4f);
        }
        __syncthreads();
        const int nvc = min(nv, CORR_MAXR);
        float mx = 0.f;
        for (int j = 0; j < nvc; ++j) mx = fmaxf(mx, iou_s[j]);
        const int base = cnt_s;
        int kept_local = 0;
        for (int j = t; j < nvc; j += 128) {
            const float me = iou_s[j];
            if (!(me > 0.f)) continue;
            int rank = 0;   // position in the descending-IoU order (ties: lower index first)
            for (int k = 0; k < nvc; ++k) {
                const float o = iou_s[k];
                rank += (o > me) || (o == me && k < j);
            }
            // kept entries form a prefix of the ranking (the condition is monotone in IoU)
            const bool keep = rank < p.topk && ((me > p.ratio * mx) || (me > p.iou_thr));
            if (keep && base + rank < p.max_match) { match[base + rank] = rs + j; kept_local++; }
        }
        if (kept_local) atomicAdd(&kept_s, kept_local);
        __syncthreads();
        if (t == 0) cnt_s = min(base + kept_s, p.max_match);
        __syncthreads();
    }
    if (t == 0) p.match_cnt[n] = cnt_s;
}

// T head: per-query key mask = OR of the (expanded) own-view cell rectangles of all matched
// RoIs, minus padded-out cells (box_correlation.py:102-115,147-157; mv2d_t_head.py:67-88).
__global__ void __launch_bounds__(256)
key_mask_kernel(Mv2dCorrParams p, int words) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ uint32_t bits[];
    __shared__ int total;
    const int n = blockIdx.x, t = threadIdx.x;
    for (int i = t; i < words; i += blockDim.x) bits[i] = 0u;
    if (t == 0) total = 0;
    __syncthreads();
    const int cnt = p.match_cnt[n];
    const int* match = p.match + (long long)n * p.max_match;
    const float st = (float)p.stride;
    const float margin1 = 0.5f * st, margin2 = (float)p.expand_stride * st;
    for (int i = 0; i < cnt; ++i) {
        const float* q = p.rois + match[i] * 5;
        const int vg = (int)q[0];                               // global view (pad_mask index)
        const int v = p.batch > 0 ? vg % p.V : vg;              // view inside the sample (key bit index)
        // candidate cell window (generous), exact fp32 test per cell as the reference does
        int cx0 = max((int)floorf((q[1] - margin1 - margin2) / st) - 1, 0);
        int cx1 = min((int)ceilf((q[3] + margin1 + margin2) / st) + 1, p.w - 1);
        int cy0 = max((int)floorf((q[2] - margin1 - margin2) / st) - 1, 0);
        int cy1 = min((int)ceilf((q[4] + margin1 + margin2) / st) + 1, p.h - 1);
        const int ww = cx1 - cx0 + 1, hh = cy1 - cy0 + 1;
        for (int c = t; c < ww * hh; c += blockDim.x) {
            const int x = cx0 + c % ww, y = cy0 + c / ww;
            const float xs = ((float)x + 0.5f) * st - 0.5f, ys = ((float)y + 0.5f) * st - 0.5f;
            const bool in = (xs + margin1 + margin2 >= q[1]) && (xs - margin1 - margin2 <= q[3]) &&
                            (ys + margin1 + margin2 >= q[2]) && (ys - margin1 - margin2 <= q[4]);
            if (in) {
                const int cell = (v * p.h + y) * p.w + x;
                if (!(p.pad_mask && p.pad_mask[(vg * p.h + y) * p.w + x])) atomicOr(&bits[cell >> 5], 1u << (cell & 31));
            }
        }
    }
    __syncthreads();
    int local = 0;
    for (int i = t; i < words; i += blockDim.x) {
        const uint32_t b = bits[i];
        p.keymask[(long long)n * words + i] = b;
        local += __popc(b);
    }
    if (p.key_cnt) {
        atomicAdd(&total, local);
        __syncthreads();
        if (t == 0) p.key_cnt[n] = total;
    }
    if (p.key_list) {
        // ordered compaction of the set bits (done once here; every decoder layer streams this list)
        __shared__ int grp_cnt[128];
        compact_key_bits(bits, words, grp_cnt, p.key_list + (long long)n * words * 32);
    }
}

// ------------------------------------------------------------------------------------------
// Next row f2: the 2D-detections hand-off of MV2D.forward_train, on the device (detectors/mv2d.py:60-117):
// process_2d_detections' min-size filter, then complement_2d_gt -- the 2D ground-truth boxes whose best IoU with the
// kept detections is below `thr` (and whose sides reach the minimum) are appended.  One CTA per view; both steps are
// ORDERED compactions (the reference's boolean indexing keeps the input order).  fp32 operation order of box_iou as in
// the reference (this unit is compiled with -fmad=false), so the `max_iou < thr` decisions are bit-identical.
__device__ __forceinline__ int ordered_slot(bool flag, int* warp_cnt, int& running) {
    // position of this thread's element among the flagged ones of the current 256-wide chunk (+ running), or -1
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const int c = warp_cnt[i]; if (i < warp) before += c; total += c; }
    const int pos = flag ? running + before + __popc(bal & ((1u << lane) - 1u)) : -1;
    running += total;
    __syncthreads();
    return pos;
}

__global__ void __launch_bounds__(256)
handoff_2d_kernel(const float* __restrict__ det, const int* __restrict__ det_start, const float* __restrict__ gt,
                  const int* __restrict__ gt_start, float min_size, float thr, float* __restrict__ out, int* __restrict__ out_count) {
    pdl_wait();
    pdl_trigger();
    __shared__ int warp_cnt[8];
    const int v = blockIdx.x, t = threadIdx.x;
    const int d0 = det_start[v], n = det_start[v + 1] - d0;
    const int g0 = gt ? gt_start[v] : 0, m = gt ? gt_start[v + 1] - g0 : 0;
    float* o = out + (long long)(d0 + g0) * 6;
    int kept = 0;
    for (int base = 0; base < n; base += 256) {
        const int i = base + t;
        bool keep = false;
        float r[6];
        if (i < n) {
#pragma unroll
            for (int k = 0; k < 6; ++k) r[k] = det[(long long)(d0 + i) * 6 + k];
            keep = !(min_size > 0.f) || ((r[2] - r[0]) >= min_size && (r[3] - r[1]) >= min_size);
        }
        const int pos = ordered_slot(keep, warp_cnt, kept);
        if (pos >= 0) {
#pragma unroll
            for (int k = 0; k < 6; ++k) o[(long long)pos * 6 + k] = r[k];
        }
    }
    __syncthreads();        // the kept detections are read back below (same CTA)
    int total = kept;
    if (thr > 0.f && m > 0) {
        const bool no_det = kept == 0;       // the reference returns ALL ground-truth boxes then, unfiltered
        for (int base = 0; base < m; base += 256) {
            const int j = base + t;
            bool keep = false;
            float r[6];
            if (j < m) {
#pragma unroll
                for (int k = 0; k < 6; ++k) r[k] = gt[(long long)(g0 + j) * 6 + k];
                if (no_det) {
                    keep = true;
                } else {
                    const float area_a = (r[2] - r[0]) * (r[3] - r[1]);
                    float best = -INFINITY;
                    for (int i = 0; i < kept; ++i) {
                        const float* b = o + (long long)i * 6;
                        const float w = fmaxf(fminf(r[2], b[2]) - fmaxf(r[0], b[0]), 0.f);
                        const float h = fmaxf(fminf(r[3], b[3]) - fmaxf(r[1], b[1]), 0.f);
                        const float inter = w * h;
                        const float area_b = (b[2] - b[0]) * (b[3] - b[1]);
                        const float uni = area_a + area_b - inter;
                        best = fmaxf(best, inter / (uni + 1e-4f));
                    }
                    keep = best < thr && (r[2] - r[0]) >= min_size && (r[3] - r[1]) >= min_size;
                }
            }
            const int pos = ordered_slot(keep, warp_cnt, total);
            if (pos >= 0) {
#pragma unroll
                for (int k = 0; k < 6; ++k) o[(long long)pos * 6 + k] = r[k];
            }
        }
    }
    if (t == 0) out_count[v] = total;
}

int run_handoff_2d(const float* det, const int* det_start, const float* gt, const int* gt_start, int V, float min_size, float thr,
                   float* out, int* out_count, cudaStream_t st) {
    MV2D_CHECK_ARG(V >= 1 && det_start && out && out_count && (thr <= 0.f || !gt || gt_start), "handoff_2d: bad arguments");
    launch_k(handoff_2d_kernel, dim3(V), dim3(256), 0, st, det, det_start, thr > 0.f ? gt : (const float*)nullptr, gt_start, min_size, thr, out,
             out_count);
    MV2D_CHECK_LAUNCH("handoff_2d");
    return 0;
}

int run_box_corr(const Mv2dCorrParams& p, cudaStream_t st) {
    MV2D_CHECK_ARG(p.N >= 0 && p.V >= 1 && p.V <= MV2D_MAXV, "box_corr: bad N/V");
    MV2D_CHECK_ARG(p.batch == 0 || (p.batch > 0 && p.rows_per_sample > 0 && p.N == p.batch * p.rows_per_sample),
                   "box_corr: batch=%d x rows_per_sample=%d != N=%d", p.batch, p.rows_per_sample, p.N);
    MV2D_CHECK_ARG(p.sample_size * p.sample_size * p.num_depth == 128,
                   "box_corr: sample_size^2*num_depth must be 128 (got %d)", p.sample_size * p.sample_size * p.num_depth);
    MV2D_CHECK_ARG(p.max_match >= 1 && p.topk >= 1, "box_corr: bad max_match/topk");
    if (p.N == 0) return 0;
    launch_k(box_corr_kernel, dim3(p.N), dim3(128), 0, st, p);
    MV2D_CHECK_LAUNCH("box_corr");
    if (p.keymask) {
        const int words = cdiv(p.V * p.h * p.w, 32);
        MV2D_CHECK_ARG(words <= 4096, "box_corr: V*h*w too large for the key list (words=%d)", words);
        MV2D_CHECK_ARG(!p.key_list || p.key_cnt, "box_corr: key_list needs key_cnt");
        launch_k(key_mask_kernel, dim3(p.N), dim3(256), words * sizeof(uint32_t), st, p, words);
        MV2D_CHECK_LAUNCH("key_mask");
    }
    return 0;
}

}  // namespace mv2d
