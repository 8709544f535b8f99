// Row f3 of SURVEY.md 8f: training targets and losses of one sample, all decoder layers at once, on the device.
//   reference: core/bbox/assigners/hungarian_assigner_3d.py:66-150 (cost = FocalLossCost + BBox3DL1Cost on the first
//              8 normalised codes, nan_to_num, scipy linear_sum_assignment on the HOST after a .cpu() sync),
//              core/bbox/match_costs/match_cost.py:6-26, core/bbox/util.py:38-58 (normalize_bbox),
//              roi_heads/bbox_heads/cross_attention_head.py:244-343 (targets), :379-434 (loss_single),
//              :475-538 (dn_loss_single); mmdet 2.25.1 FocalLoss / L1Loss / FocalLossCost.
// Three launches, no host round trip (the reference syncs once per decoder layer for the assignment):
//   cost_kernel  [L, N, G] fp32 cost matrices
//   lsa_kernel   one CTA per layer: rectangular linear sum assignment by shortest augmenting paths (the algorithm
//                behind scipy.optimize.linear_sum_assignment, Crouse 2016), duals in fp64, the column scan of
//                every Dijkstra step spread over 1024 threads
//   loss_kernel  one CTA per (layer, {matching, denoising}): focal + weighted L1 sums in fp64, fixed-order reduction
#include "common.cuh"
#include "mv2d_internal.h"

namespace mv2d {

#define LOSS_CODE 10
#define LSA_THREADS 1024

__device__ __forceinline__ void normalize_gt(const float* __restrict__ b, float (&o)[LOSS_CODE]) {
    o[0] = b[0]; o[1] = b[1]; o[2] = logf(b[3]); o[3] = logf(b[4]); o[4] = b[2]; o[5] = logf(b[5]);
    o[6] = sinf(b[6]); o[7] = cosf(b[6]); o[8] = b[7]; o[9] = b[8];
}

struct CostArgs {
    const float* cls; const float* box; long long layer_stride;
    const float* gt_boxes; const int* gt_labels;
    int N, G, L, num_classes;
    float alpha, gamma, cls_w, reg_w;
    float* cost;     // [L,N,G]
};

__global__ void __launch_bounds__(256) cost_kernel(CostArgs a) {
    pdl_wait();
    pdl_trigger();
    const int l = blockIdx.y;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (long long)a.N * a.G) return;
    const int n = (int)(idx / a.G), g = (int)(idx % a.G);
    const float* cls = a.cls + l * a.layer_stride + (long long)n * a.num_classes;
    const float* box = a.box + l * a.layer_stride + (long long)n * LOSS_CODE;
    float gn[LOSS_CODE];
    normalize_gt(a.gt_boxes + g * 9, gn);
    const float x = cls[a.gt_labels[g]];
    const float p = 1.f / (1.f + expf(-x));
    const float neg = -logf(1.f - p + 1e-12f) * (1.f - a.alpha) * powf(p, a.gamma);
    const float pos = -logf(p + 1e-12f) * a.alpha * powf(1.f - p, a.gamma);
    float reg = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) reg += fabsf(box[j] - gn[j]);
    float c = (pos - neg) * a.cls_w + reg * a.reg_w;
    if (isnan(c)) c = 100.f;                       // torch.nan_to_num(cost, nan=100, posinf=100, neginf=-100)
    else if (isinf(c)) c = c > 0.f ? 100.f : -100.f;
    a.cost[((long long)l * a.N + n) * a.G + g] = c;
}

struct LsaArgs {
    const float* cost;   // [L,N,G]
    int N, G;
    int* assigned;       // [L,N] gt index or -1
};

// rows = the smaller side (every row gets a column), cols = the larger side.  elem(i, j) reads the [N,G] matrix.
__global__ void __launch_bounds__(LSA_THREADS) lsa_kernel(LsaArgs a) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ __align__(16) unsigned char lsa_smem[];
    const int l = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool tr = a.G <= a.N;                  // rows are the GT boxes, columns the queries
    const int R = tr ? a.G : a.N, C = tr ? a.N : a.G;
    const float* cost = a.cost + (long long)l * a.N * a.G;
    auto elem = [&](int i, int j) -> double { return (double)(tr ? cost[(long long)j * a.G + i] : cost[(long long)i * a.G + j]); };
    double* u = reinterpret_cast<double*>(lsa_smem);      // [R]
    double* v = u + R;                                     // [C]
    double* spc = v + C;                                   // [C] shortest path costs
    int* path = reinterpret_cast<int*>(spc + C);           // [C]
    int* col4row = path + C;                               // [R]
    int* row4col = col4row + R;                            // [C]
    unsigned char* SR = reinterpret_cast<unsigned char*>(row4col + C);   // [R]
    unsigned char* SC = SR + R;                            // [C]
    __shared__ double red_v[32];
    __shared__ int red_j[32], red_f[32];
    __shared__ int s_i, s_sink, s_j;
    __shared__ double s_min;
    for (int k = tid; k < R; k += LSA_THREADS) { u[k] = 0.0; col4row[k] = -1; }
    for (int k = tid; k < C; k += LSA_THREADS) { v[k] = 0.0; row4col[k] = -1; }
    int* out = a.assigned + (long long)l * a.N;
    for (int k = tid; k < a.N; k += LSA_THREADS) out[k] = -1;
    __syncthreads();
    for (int cur = 0; cur < R; ++cur) {
        for (int k = tid; k < C; k += LSA_THREADS) { spc[k] = INFINITY; SC[k] = 0; path[k] = -1; }
        for (int k = tid; k < R; k += LSA_THREADS) SR[k] = 0;
        if (tid == 0) { s_i = cur; s_sink = -1; s_min = 0.0; }
        __syncthreads();
        while (true) {
            const int i = s_i;
            const double minVal = s_min, ui = u[i];
            if (tid == 0) SR[i] = 1;
            // relax the row's edges and find the closest unscanned column (ties: an unassigned column first)
            double bv = INFINITY; int bj = -1, bf = 0;
            for (int j = tid; j < C; j += LSA_THREADS) {
                if (SC[j]) continue;
                const double r = minVal + elem(i, j) - ui - v[j];
                double s = spc[j];
                if (r < s) { path[j] = i; spc[j] = r; s = r; }
                const int f = row4col[j] == -1;
                if (s < bv || (s == bv && f > bf)) { bv = s; bj = j; bf = f; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oj = __shfl_xor_sync(0xffffffffu, bj, o), of = __shfl_xor_sync(0xffffffffu, bf, o);
                const bool take = oj >= 0 && (bj < 0 || ov < bv || (ov == bv && (of > bf || (of == bf && oj < bj))));
                if (take) { bv = ov; bj = oj; bf = of; }
            }
            if (lane == 0) { red_v[warp] = bv; red_j[warp] = bj; red_f[warp] = bf; }
            __syncthreads();
            if (warp == 0) {
                bv = red_v[lane]; bj = red_j[lane]; bf = red_f[lane];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    const int oj = __shfl_xor_sync(0xffffffffu, bj, o), of = __shfl_xor_sync(0xffffffffu, bf, o);
                    const bool take = oj >= 0 && (bj < 0 || ov < bv || (ov == bv && (of > bf || (of == bf && oj < bj))));
                    if (take) { bv = ov; bj = oj; bf = of; }
                }
                if (lane == 0) {
                    s_j = bj; s_min = bv;
                    if (bj < 0 || bv == INFINITY) s_sink = -2;            // infeasible: cannot happen after nan_to_num
                    else {
                        SC[bj] = 1;
                        if (row4col[bj] == -1) s_sink = bj; else s_i = row4col[bj];
                    }
                }
            }
            __syncthreads();
            if (s_sink != -1) break;
        }
        if (s_sink < 0) break;
        const double minVal = s_min;
        // dual updates
        for (int k = tid; k < R; k += LSA_THREADS) {
            if (k == cur) u[k] += minVal;
            else if (SR[k]) u[k] += minVal - spc[col4row[k]];
        }
        for (int k = tid; k < C; k += LSA_THREADS)
            if (SC[k]) v[k] -= minVal - spc[k];
        __syncthreads();
        if (tid == 0) {      // augment along the path (a handful of steps)
            int j = s_sink;
            while (true) {
                const int i = path[j];
                row4col[j] = i;
                const int t = col4row[i]; col4row[i] = j; j = t;
                if (i == cur) break;
            }
        }
        __syncthreads();
    }
    // rows -> output: assigned[query] = gt
    for (int k = tid; k < R; k += LSA_THREADS) {
        const int j = col4row[k];
        if (j >= 0) { if (tr) out[j] = k; else out[k] = j; }
    }
}

struct LossArgs {
    const float* cls; const float* box; long long layer_stride;      // matching queries [L,N,*]
    const float* dn_cls; const float* dn_box; long long dn_layer_stride; const int* dn_labels; int pad; int neg_bbox_loss;
    const float* gt_boxes; const int* gt_labels; const int* assigned;
    int N, G, num_classes;
    float alpha, gamma, cls_lw, box_lw, dn_split;
    float code_w[LOSS_CODE];
    float* losses;       // [L,4]
    float* num_pos;                    // nullable out [L]
    const float* bbox_avg_factor;      // nullable in [L]
};

__device__ __forceinline__ double focal_elem(float x, bool t, float alpha, float gamma) {
    const float p = 1.f / (1.f + expf(-x));
    const float pt = t ? 1.f - p : p;
    const float fw = (t ? alpha : 1.f - alpha) * powf(pt, gamma);
    // binary_cross_entropy_with_logits(x, t) = max(x, 0) - x t + log1p(exp(-|x|))
    const float bce = fmaxf(x, 0.f) - (t ? x : 0.f) + log1pf(expf(-fabsf(x)));
    return (double)(bce * fw);
}

// grid (L, 2): y = 0 matching queries, y = 1 denoising queries.  256 threads.
__global__ void __launch_bounds__(256) loss_kernel(LossArgs a) {
    pdl_wait();
    pdl_trigger();
    const int l = blockIdx.x, dn = blockIdx.y, tid = threadIdx.x;
    if (dn && (a.pad == 0 || a.dn_cls == nullptr)) {
        if (tid == 0) { a.losses[l * 4 + 2] = 0.f; a.losses[l * 4 + 3] = 0.f; }
        return;
    }
    const int M = dn ? a.pad : a.N;
    const float* cls = dn ? a.dn_cls + l * a.dn_layer_stride : a.cls + l * a.layer_stride;
    const float* box = dn ? a.dn_box + l * a.dn_layer_stride : a.box + l * a.layer_stride;
    const int* asg = a.assigned + (long long)l * a.N;
    double fsum = 0.0, bsum = 0.0;
    int npos = 0;
    for (int n = tid; n < M; n += 256) {
        int label, g = -1;
        if (dn) { label = a.dn_labels[n]; g = (label != a.num_classes || a.neg_bbox_loss) ? n % a.G : -1; }
        else { g = asg[n]; label = g >= 0 ? a.gt_labels[g] : a.num_classes; }
        for (int c = 0; c < a.num_classes; ++c) fsum += focal_elem(cls[(long long)n * a.num_classes + c], c == label, a.alpha, a.gamma);
        if (g >= 0) {
            npos += 1;
            float gn[LOSS_CODE];
            normalize_gt(a.gt_boxes + g * 9, gn);
            bool ok = true;
#pragma unroll
            for (int j = 0; j < LOSS_CODE; ++j) ok = ok && isfinite(gn[j]);
            if (ok) {
#pragma unroll
                for (int j = 0; j < LOSS_CODE; ++j) {
                    const float w = (dn && (j == 6 || j == 7)) ? 0.f : a.code_w[j];
                    bsum += (double)(fabsf(box[(long long)n * LOSS_CODE + j] - gn[j]) * w);
                }
            }
        }
    }
    __shared__ double sf[256], sb[256];
    __shared__ int sp[256];
    sf[tid] = fsum; sb[tid] = bsum; sp[tid] = npos;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) { sf[tid] += sf[tid + o]; sb[tid] += sb[tid + o]; sp[tid] += sp[tid + o]; }
        __syncthreads();
    }
    if (tid == 0) {
        const double eps = 1.1920928955078125e-07;       // torch.finfo(float32).eps (mmdet weight_reduce_loss)
        double cls_avg, box_avg;
        if (dn) {
            cls_avg = fmax((double)a.pad * 3.14159 / 6.0 * a.dn_split * a.dn_split * a.dn_split, 1.0);
            box_avg = fmax((double)a.pad, 1.0);
        } else {
            cls_avg = fmax((double)sp[0], 1.0);          // bg_cls_weight = 0 with a sigmoid focal loss (cross_attention_head.py:153)
            // clamp(reduce_mean(num_total_pos), min=1) over ranks when the caller supplies it (cross_attention_head.py:419-420)
            box_avg = a.bbox_avg_factor ? (double)a.bbox_avg_factor[l] : fmax((double)sp[0], 1.0);
            if (a.num_pos) a.num_pos[l] = (float)sp[0];
        }
        float lc = (float)(a.cls_lw * sf[0] / (cls_avg + eps));
        float lb = (float)(a.box_lw * sb[0] / (box_avg + eps));
        if (isnan(lc)) lc = 0.f;                          // torch.nan_to_num
        if (isnan(lb)) lb = 0.f;
        a.losses[l * 4 + dn * 2 + 0] = lc;
        a.losses[l * 4 + dn * 2 + 1] = lb;
    }
}

static size_t lsa_smem_bytes(int N, int G) {
    const size_t R = (size_t)(G <= N ? G : N), C = (size_t)(G <= N ? N : G);
    return R * 8 + C * 16 + C * 4 + R * 4 + C * 4 + R + C + 16;
}

size_t loss_workspace_bytes(int N, int G, int L) {
    return ((size_t)(L > 0 ? L : 1) * (size_t)(N > 0 ? N : 1) * (size_t)(G > 0 ? G : 1)) * sizeof(float) + 256;
}

int run_loss(const Mv2dLossParams& p, cudaStream_t st) {
    MV2D_CHECK_ARG(p.N >= 0 && p.G >= 0 && p.L >= 1 && p.L <= MV2D_MAX_LAYERS && p.num_classes >= 1 && p.num_classes <= 64,
                   "loss: bad N=%d G=%d L=%d classes=%d", p.N, p.G, p.L, p.num_classes);
    MV2D_CHECK_ARG(p.losses && (p.N == 0 || p.assigned), "loss: null output");
    MV2D_CHECK_ARG(p.N == 0 || (p.cls_scores && p.bbox_preds), "loss: null predictions");
    MV2D_CHECK_ARG(p.G == 0 || (p.gt_boxes && p.gt_labels), "loss: null ground truth");
    MV2D_CHECK_ARG(p.pad == 0 || p.G > 0, "loss: denoising queries need ground truth");
    cudaError_t e;
    if (p.N > 0 && (e = cudaMemsetAsync(p.assigned, 0xff, (size_t)p.L * p.N * sizeof(int), st)) != cudaSuccess) {
        set_error("loss: memset %s", cudaGetErrorString(e));
        return (int)e;
    }
    if (p.N > 0 && p.G > 0) {
        MV2D_CHECK_ARG(p.workspace && p.workspace_bytes >= loss_workspace_bytes(p.N, p.G, p.L), "loss: workspace too small");
        const size_t smem = lsa_smem_bytes(p.N, p.G);
        MV2D_CHECK_ARG(smem <= 200 * 1024, "loss: N=%d x G=%d does not fit the assignment kernel's shared memory", p.N, p.G);
        CostArgs c{};
        c.cls = p.cls_scores; c.box = p.bbox_preds; c.layer_stride = p.layer_stride; c.gt_boxes = p.gt_boxes; c.gt_labels = p.gt_labels;
        c.N = p.N; c.G = p.G; c.L = p.L; c.num_classes = p.num_classes; c.alpha = p.focal_alpha; c.gamma = p.focal_gamma;
        c.cls_w = p.cls_cost_weight; c.reg_w = p.reg_cost_weight; c.cost = p.workspace;
        launch_k(cost_kernel, dim3((unsigned)(((long long)p.N * p.G + 255) / 256), p.L), dim3(256), 0, st, c);
        MV2D_CHECK_LAUNCH("loss cost");
        if ((e = cudaFuncSetAttribute(lsa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) {
            set_error("loss: smem attr %s", cudaGetErrorString(e));
            return (int)e;
        }
        LsaArgs s{}; s.cost = p.workspace; s.N = p.N; s.G = p.G; s.assigned = p.assigned;
        launch_k(lsa_kernel, dim3(p.L), dim3(LSA_THREADS), smem, st, s);
        MV2D_CHECK_LAUNCH("loss lsa");
    }
    LossArgs a{};
    a.cls = p.cls_scores; a.box = p.bbox_preds; a.layer_stride = p.layer_stride;
    a.dn_cls = p.dn_cls; a.dn_box = p.dn_box; a.dn_layer_stride = p.dn_layer_stride; a.dn_labels = p.dn_labels; a.pad = p.pad; a.neg_bbox_loss = p.neg_bbox_loss;
    a.gt_boxes = p.gt_boxes; a.gt_labels = p.gt_labels; a.assigned = p.assigned;
    a.N = p.N; a.G = p.G; a.num_classes = p.num_classes; a.alpha = p.focal_alpha; a.gamma = p.focal_gamma;
    a.cls_lw = p.cls_loss_weight; a.box_lw = p.bbox_loss_weight; a.dn_split = p.dn_split;
    for (int j = 0; j < LOSS_CODE; ++j) a.code_w[j] = p.code_weights[j];
    a.losses = p.losses; a.num_pos = p.num_pos; a.bbox_avg_factor = p.bbox_avg_factor;
    launch_k(loss_kernel, dim3(p.L, 2), dim3(256), 0, st, a);
    MV2D_CHECK_LAUNCH("loss");
    return 0;
}

}  // namespace mv2d
