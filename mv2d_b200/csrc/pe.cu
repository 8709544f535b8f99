// K1: 3D position embedding (rows a1-a3 of SURVEY.md section 8a).
//   reference: mmdet3d_plugin/models/utils/pe.py:84-169, positional_encoding.py:58-96
// Layout: everything channels-last, pixel p = (v*h + y)*w + x, row p of a [P, C] matrix.
#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "mlp2.cuh"
#include "mv2d_internal.h"

namespace mv2d {

// ---- geometry prep: img2lidar[v] = inv(lidar2img[v]); trans[src][dst] = lidar2img[dst] @ img2lidar[src]
// (pe.py:111; box_correlation.py:118-122).  One block, V*V threads.
__global__ void geom_prep_kernel(const double* __restrict__ lidar2img, int V,
                                 double* __restrict__ img2lidar, double* __restrict__ trans) {
    pdl_wait();
    pdl_trigger();
    __shared__ double inv_s[MV2D_MAXV][16];
    int t = threadIdx.x;
    // one block per sample of a batch: the V views of sample blockIdx.x (trans couples views of ONE sample only)
    lidar2img += (long long)blockIdx.x * V * 16;
    img2lidar += (long long)blockIdx.x * V * 16;
    trans += (long long)blockIdx.x * V * V * 16;
    if (t < V) {
        double out[16];
        inv4x4(lidar2img + t * 16, out);
        for (int i = 0; i < 16; ++i) { inv_s[t][i] = out[i]; img2lidar[t * 16 + i] = out[i]; }
    }
    __syncthreads();
    if (t < V * V) {
        int src = t / V, dst = t % V;
        double out[16];
        mat4_mul(lidar2img + dst * 16, inv_s[src], out);
        for (int i = 0; i < 16; ++i) trans[(src * V + dst) * 16 + i] = out[i];
    }
}

// ---- NCHW -> NHWC (the FPN hands us NCHW; every kernel below wants C contiguous)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, float* __restrict__ out_tf32, int C, int HW,
                                    const float* __restrict__ in2) {
    pdl_wait();
    pdl_trigger();
    __shared__ float tile[32][33];
    const int v = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float* src = in + (long long)v * C * HW;
    const float* src2 = in2 ? in2 + (long long)v * C * HW : nullptr;      // optional second map, added (key = memory + pos)
    float* dst = out + (long long)v * C * HW;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i, p = p0 + threadIdx.x;
        float val = (c < C && p < HW) ? src[(long long)c * HW + p] : 0.f;
        if (src2 && c < C && p < HW) val += src2[(long long)c * HW + p];
        tile[i][threadIdx.x] = val;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int p = p0 + i, c = c0 + threadIdx.x;
        if (c < C && p < HW) {
            const float val = tile[threadIdx.x][i];
            dst[(long long)p * C + c] = val;
            if (out_tf32) out_tf32[(long long)v * C * HW + (long long)p * C + c] = round_tf32(val);
        }
    }
}

// C = 256 (the hot path's feature maps): one CTA moves 64 pixels x all 256 channels, so it writes 64 KB of consecutive
// channels-last rows (the 32 x 32 tiles above write 128-byte pieces of 1 KB rows from four different CTAs) and reads
// 256-byte runs of every channel plane.
constexpr int NHWC_PT = 64;
constexpr int NHWC_SMEM = 256 * (NHWC_PT + 1) * 4;
__global__ void __launch_bounds__(256)
nchw_to_nhwc_c256_kernel(const float* __restrict__ in, float* __restrict__ out, float* __restrict__ out_tf32, int HW,
                         const float* __restrict__ in2, float* __restrict__ out_lo) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ float nh_tile[];          // [256][65]
    const int v = blockIdx.y, p0 = blockIdx.x * NHWC_PT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* src = in + (long long)v * 256 * HW + p0;
    const float* src2 = in2 ? in2 + (long long)v * 256 * HW + p0 : nullptr;
    const bool ok0 = p0 + lane < HW, ok1 = p0 + 32 + lane < HW;
#pragma unroll 8
    for (int c = warp; c < 256; c += 8) {
        float a = ok0 ? __ldg(src + (long long)c * HW + lane) : 0.f;
        float b = ok1 ? __ldg(src + (long long)c * HW + 32 + lane) : 0.f;
        if (src2) {
            if (ok0) a += __ldg(src2 + (long long)c * HW + lane);
            if (ok1) b += __ldg(src2 + (long long)c * HW + 32 + lane);
        }
        nh_tile[c * (NHWC_PT + 1) + lane] = a;
        nh_tile[c * (NHWC_PT + 1) + 32 + lane] = b;
    }
    __syncthreads();
    for (int p = warp; p < NHWC_PT; p += 8) {
        if (p0 + p >= HW) break;
        const long long o = ((long long)v * HW + p0 + p) * 256;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float val = nh_tile[(k * 32 + lane) * (NHWC_PT + 1) + p];
            out[o + k * 32 + lane] = val;
            if (out_tf32) {
                const float hi = round_tf32(val);
                out_tf32[o + k * 32 + lane] = hi;
                if (out_lo) out_lo[o + k * 32 + lane] = round_tf32(val - hi);
            }
        }
    }
}

// ---- frustum coordinates -> [P, 3*D] (pe.py:93-130).  fp64 geometry as in the reference.
// thread = (pixel, depth); 3 consecutive floats per thread => a warp writes 384 contiguous bytes.
__global__ void pe_coords_kernel(const double* __restrict__ img2lidar, float* __restrict__ out,
                                 int V, int h, int w, int D, double pad_h, double pad_w,
                                 double depth_start, double pr0, double pr1, double pr2,
                                 double bin, double ir0, double ir1, double ir2, int tf32) {
    // bin = (pr3 - depth_start) / (D (1 + D)) and ir = 1 / (hi - lo) come from the host: fp64 divisions are ~40-instruction
    // sequences and were most of this kernel
    pdl_wait();
    pdl_trigger();
    // 32-bit index arithmetic (the launchers check V h w D < 2^31): 64-bit divisions are long instruction sequences too
    const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned total = (unsigned)V * h * w * D;
    if (gid >= total) return;
    const unsigned d = gid % (unsigned)D;
    const unsigned p = gid / (unsigned)D;
    const unsigned x = p % (unsigned)w, y = (p / (unsigned)w) % (unsigned)h, v = p / ((unsigned)w * h);
    const double cw = ((double)x + 0.5) * pad_w - 0.5;        // pad_w, pad_h: already divided by w, h on the host
    const double ch = ((double)y + 0.5) * pad_h - 0.5;
    const double cd = depth_start + bin * (double)d * ((double)d + 1.0);
    const double s = fmax(cd, 1e-3);
    const double c0 = cw * s, c1 = ch * s;
    const double* m = img2lidar + v * 16;
    const double lo[3] = {pr0, pr1, pr2}, ir[3] = {ir0, ir1, ir2};
    float r[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        // the frustum point and its normalisation in fp64 as the reference; the logit itself in fp32: the value is
        // rounded to TF32 (2^-11) right below because it feeds a TF32 tensor-core GEMM, fp32 log error (2^-23) is noise
        double c = m[i * 4 + 0] * c0 + m[i * 4 + 1] * c1 + m[i * 4 + 2] * cd + m[i * 4 + 3];
        c = (c - lo[i]) * ir[i];
        const float cf = (float)fmin(fmax(c, 0.0), 1.0);
        const float num = fmaxf(cf, 1e-5f), den = fmaxf((float)(1.0 - fmin(fmax(c, 0.0), 1.0)), 1e-5f);
        const float lg = __logf(__fdividef(num, den));     // inverse_sigmoid (fast intrinsics: ~1e-7 absolute, the kernel is instruction bound)
        r[i] = tf32 ? round_tf32(lg) : lg;                 // inference: operand of a TF32 GEMM; training keeps fp32
    }
    float* o = out + (size_t)p * (3 * D) + d * 3;
    o[0] = r[0]; o[1] = r[1]; o[2] = r[2];
}

// ---- SinePositionalEncoding3D, normalize=True (positional_encoding.py:58-96).
// Step 1: per pixel the three normalised embeds (view, y, x) from the not-mask cumsums.
__global__ void sine_prep_kernel(const uint8_t* __restrict__ not_mask, float* __restrict__ emb,
                                 int V, int h, int w, float stride, float scale, float eps, int vps) {
    pdl_wait();
    pdl_trigger();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= V * h * w) return;
    const int x = p % w, y = (p / w) % h, v = p / (w * h);
    float n = 0.f, nl = 0.f, ye = 0.f, yl = 0.f, xe = 0.f, xl = 0.f;
    // the view cumsum runs over the views of the cell's own sample (vps views per sample; one sample: vps = V)
    const int v0 = (v / vps) * vps;
    for (int i = v0; i < v0 + vps; ++i) {
        float m = (float)not_mask[(i * h + y) * w + x];
        nl += m;
        if (i <= v) n += m;
    }
    for (int i = 0; i < h; ++i) {
        float m = (float)not_mask[(v * h + i) * w + x];
        yl += m;
        if (i <= y) ye += m;
    }
    for (int i = 0; i < w; ++i) {
        float m = (float)not_mask[(v * h + y) * w + i];
        xl += m;
        if (i <= x) xe += m;
    }
    if (stride > 0.f) {
        ye = (ye - 0.5f) * stride; yl = (yl - 0.5f) * stride;
        xe = (xe - 0.5f) * stride; xl = (xl - 0.5f) * stride;
    }
    emb[p * 3 + 0] = n / (nl + eps) * scale;
    emb[p * 3 + 1] = ye / (yl + eps) * scale;
    emb[p * 3 + 2] = xe / (xl + eps) * scale;
}

// Step 2: [P, 384] = per embed (n, y, x): 64 sines of even dim_t then 64 cosines of odd dim_t.
__global__ void sine_embed_kernel(const float* __restrict__ emb, const float* __restrict__ dim_t,
                                  float* __restrict__ out, int P, int tf32) {
    pdl_wait();
    pdl_trigger();
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)P * 384) return;
    const int c = (int)(gid % 384);
    const long long p = gid / 384;
    const int e = c >> 7, i = c & 127;
    const float val = emb[p * 3 + e];
    // arguments lie in [0, 2*pi]; the SFU sine/cosine (abs. error ~1e-6 there) is far inside the TF32 rounding below
    float r;
    if (!tf32) {                // training: fp32 operands, libm sine / cosine
        out[gid] = i < 64 ? sinf(val / __ldg(dim_t + 2 * i)) : cosf(val / __ldg(dim_t + 2 * (i - 64) + 1));
        return;
    }
    if (i < 64) r = __sinf(val / __ldg(dim_t + 2 * i));
    else        r = __cosf(val / __ldg(dim_t + 2 * (i - 64) + 1));
    out[gid] = round_tf32(r);   // operand of a TF32 GEMM
}

// ---- separable form of the sine branch's first layer (no padded cells: not_mask is all ones).
// The 384 sine features of a cell are [f(e_view) | f(e_y) | f(e_x)], and each normalised embed depends on ONE
// coordinate only, so  W1 . s(v,y,x) = Tv[v] + Ty[y] + Tx[x]  with three tiny tables (V + h + w rows instead of
// V*h*w): the 384 -> 1024 GEMM over all cells (13 GFLOP at V = 6) becomes a 126-row GEMM plus an element-wise
// gather-add.  Exact in real arithmetic; the tables are accumulated in fp32 FFMA (the full-size GEMM was TF32).
// Step A: rows r < V: view v = r; r < V + h: y = r - V; else x = r - V - h.  F[r] holds the row's 128 features in its
// own 128-column slot of a zero-padded [rows, 384] matrix, so ONE GEMM against W1 [1024, 384] yields all three tables.
__global__ void __launch_bounds__(128) sine_axis_kernel(const float* __restrict__ dim_t, float* __restrict__ F,
                                                        int V, int h, int w, float stride, float scale, float eps) {
    pdl_wait();
    pdl_trigger();
    const int r = blockIdx.x, i = threadIdx.x;
    int axis, idx, len;
    if (r < V) { axis = 0; idx = r; len = V; }
    else if (r < V + h) { axis = 1; idx = r - V; len = h; }
    else { axis = 2; idx = r - V - h; len = w; }
    // same fp32 operations as sine_prep_kernel with an all-ones mask: cumsum = idx + 1, last = len
    float e = (float)(idx + 1), l = (float)len;
    if (axis > 0 && stride > 0.f) { e = (e - 0.5f) * stride; l = (l - 0.5f) * stride; }
    const float val = e / (l + eps) * scale;
    float f;
    if (i < 64) f = __sinf(val / __ldg(dim_t + 2 * i));
    else        f = __cosf(val / __ldg(dim_t + 2 * (i - 64) + 1));
    float* row = F + (long long)r * 384;
    row[i] = axis == 0 ? f : 0.f;
    row[128 + i] = axis == 1 ? f : 0.f;
    row[256 + i] = axis == 2 ? f : 0.f;
}

// Step C: hidden[p, :] = tf32(relu(Tv[v] + Ty[y] + Tx[x] + b)), the A operand of the 1024 -> 256 TF32 GEMM.
__global__ void __launch_bounds__(256) sine_hidden_kernel(const float* __restrict__ T, const float* __restrict__ bias,
                                                          float* __restrict__ Hd, int V, int h, int w, int vps) {
    pdl_wait();
    pdl_trigger();
    const long long gid = (long long)blockIdx.x * 256 + threadIdx.x;      // one float4 of a 1024-wide row
    const long long total = (long long)V * h * w * 256;
    if (gid >= total) return;
    const int j4 = (int)(gid & 255);
    const long long p = gid >> 8;
    // tables: vps view rows (the view index inside the cell's sample), then h row rows, then w column rows
    const int x = (int)(p % w), y = (int)((p / w) % h), v = (int)(p / ((long long)w * h)) % vps;
    const float4 a = __ldg(reinterpret_cast<const float4*>(T + (long long)v * 1024) + j4);
    const float4 b = __ldg(reinterpret_cast<const float4*>(T + (long long)(vps + y) * 1024) + j4);
    const float4 c = __ldg(reinterpret_cast<const float4*>(T + (long long)(vps + h + x) * 1024) + j4);
    const float4 d = __ldg(reinterpret_cast<const float4*>(bias) + j4);
    float4 o;
    o.x = round_tf32(fmaxf(((a.x + b.x) + c.x) + d.x, 0.f)); o.y = round_tf32(fmaxf(((a.y + b.y) + c.y) + d.y, 0.f));
    o.z = round_tf32(fmaxf(((a.z + b.z) + c.z) + d.z, 0.f)); o.w = round_tf32(fmaxf(((a.w + b.w) + c.w) + d.w, 0.f));
    reinterpret_cast<float4*>(Hd)[gid] = o;
}

// ---- separable sine branch on the fused MLP kernel (mlp2.cu): the gather-add of the three table rows becomes the FIRST
// GEMM of the fused kernel.  A = [onehot | onehot] (a one in column v, vps + y and vps + h + x of each half), W0 = the
// transposed tables split into TF32 hi / lo halves: A . W0^T = Tv[v] + Ty[y] + Tx[x] with every product exact and fp32
// accumulation, i.e. fp32-grade like the gather kernel.  KH = table rows rounded up to 32, K0 = 2 KH.
__global__ void __launch_bounds__(256) sine_onehot_kernel(float* __restrict__ A, int Ps, int h, int w, int vps, int KH) {
    pdl_wait();
    pdl_trigger();
    const int k4n = KH >> 1;                                    // float4s per row (K0 / 4)
    const long long gid = (long long)blockIdx.x * 256 + threadIdx.x;
    if (gid >= (long long)Ps * k4n) return;
    const int k0 = (int)(gid % k4n) * 4;
    const long long p = gid / k4n;
    const int x = (int)(p % w), y = (int)((p / w) % h), v = (int)(p / ((long long)w * h)) % vps;
    const int kk = k0 >= KH ? k0 - KH : k0;                     // column inside the half
    float o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = kk + i;
        o[i] = (c == v || c == vps + y || c == vps + h + x) ? 1.f : 0.f;
    }
    reinterpret_cast<float4*>(A)[gid] = make_float4(o[0], o[1], o[2], o[3]);
}

// W0s[j][r] = tf32_hi(T[r][j]), W0s[j][KH + r] = tf32_lo(T[r][j]), zero for r >= rows.  grid (1024 / 32, KH / 32), block (32, 8)
__global__ void sine_w0_kernel(const float* __restrict__ T, float* __restrict__ W0s, int rows, int KH) {
    pdl_wait();
    pdl_trigger();
    __shared__ float tile[32][33];
    const int j0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int r = r0 + i;
        tile[i][threadIdx.x] = r < rows ? T[(long long)r * 1024 + j0 + threadIdx.x] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const float val = tile[threadIdx.x][i], hi = round_tf32(val);
        float* o = W0s + (long long)(j0 + i) * (2 * KH) + r0 + threadIdx.x;
        o[0] = hi;
        o[KH] = round_tf32(val - hi);
    }
}

static int gemm(const float* A, int lda, const float* W, int ldw, const float* bias, float* C, int ldc,
                int M, int N, int K, int flags, cudaStream_t st, const float* gx = nullptr,
                const float* gs = nullptr, const float* gfeat = nullptr, float* kin = nullptr, int gs_mod = 0) {
    GemmArgs g{};
    g.A = A; g.lda = lda; g.W = W; g.ldw = ldw; g.C = C; g.ldc = ldc; g.bias = bias;
    g.M = M; g.N = N; g.K = K; g.batch = 1; g.nsplit = 1; g.flags = flags;
    g.gx = gx; g.gs = gs; g.gfeat = gfeat; g.kin = kin; g.gs_mod = gs_mod;
    return launch_gemm_tc_or_simt(g, st);
}

int run_geom_prep(const double* lidar2img, int V, double* img2lidar, double* trans, cudaStream_t st, int batch) {
    MV2D_CHECK_ARG(V >= 1 && V <= MV2D_MAXV && batch >= 1, "geom_prep: V=%d / batch=%d out of range", V, batch);
    launch_k(geom_prep_kernel, dim3(batch), dim3(V * V), 0, st, lidar2img, V, img2lidar, trans);
    MV2D_CHECK_LAUNCH("geom_prep");
    return 0;
}

int run_nchw_to_nhwc(const float* in, float* out, float* out_tf32, int V, int C, int HW, cudaStream_t st, const float* in2, float* out_lo) {
    MV2D_CHECK_ARG(out_lo == nullptr || (C == 256 && out_tf32 != nullptr), "nchw_to_nhwc: the lo half needs C = 256 and the TF32 copy");
    MV2D_CHECK_ARG(V <= 65535, "nchw_to_nhwc: at most 65535 maps per call (got %d)", V);
    static const bool wide = []() { const char* e = getenv("MV2D_NHWC_WIDE"); return !(e && e[0] == '0'); }();
    if (C == 256 && (wide || out_lo)) {
        static bool attr_set = false;
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(nchw_to_nhwc_c256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NHWC_SMEM);
            if (e != cudaSuccess) { set_error("nchw_to_nhwc: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
            attr_set = true;
        }
        launch_k(nchw_to_nhwc_c256_kernel, dim3(cdiv(HW, NHWC_PT), V), dim3(256), (size_t)NHWC_SMEM, st, in, out, out_tf32, HW, in2, out_lo);
        MV2D_CHECK_LAUNCH("nchw_to_nhwc");
        return 0;
    }
    dim3 grid(cdiv(HW, 32), cdiv(C, 32), V), block(32, 8);
    launch_k(nchw_to_nhwc_kernel, grid, block, 0, st, in, out, out_tf32, C, HW, in2);
    MV2D_CHECK_LAUNCH("nchw_to_nhwc");
    return 0;
}

// PE.forward.  feat is NHWC [P,256].  Outputs pe [P,256] and (optional) kin = feat + pe.
int run_pe3d(const Mv2dPeParams& p, cudaStream_t st) {
    const int P = p.V * p.h * p.w, D = p.depth_num, C = MV2D_C;
    MV2D_CHECK_ARG(p.V >= 1 && p.V <= MV2D_MAXVB && P > 0, "pe3d: bad V/h/w");
    // batch: V counts the views of ALL samples, vps = views of one sample (the sine embedding's view axis is per sample)
    const int vps = p.views_per_sample > 0 ? p.views_per_sample : p.V;
    MV2D_CHECK_ARG(vps <= MV2D_MAXV && p.V % vps == 0, "pe3d: V=%d is not a multiple of views_per_sample=%d", p.V, vps);
    // sine_shared: every sample of the batch has the same padding masks, so adapt_pos3d(sine) is identical for all of
    // them: it is evaluated for the first sample's cells only and read with the row index modulo vps*h*w
    const int Vs = p.sine_shared ? vps : p.V;          // views the sine branch is evaluated for
    const int Ps = Vs * p.h * p.w;
    MV2D_CHECK_ARG(!p.sine_shared || (!p.sine_branch_cached && !p.sine_branch_out), "pe3d: sine_shared excludes the cached sine branch");
    MV2D_CHECK_ARG((3 * D) % 16 == 0, "pe3d: 3*depth_num must be a multiple of 16");
    float* ws = p.workspace;
    float* A1 = ws;                      ws += (size_t)P * 3 * D;
    float* Hd = ws;                      ws += (size_t)P * 4 * C;
    float* X  = ws;                      ws += (size_t)P * C;
    float* G1 = ws;                      ws += (size_t)P * C;
    float* S  = ws;                      ws += (size_t)P * 384;
    float* SB = ws;                      ws += (size_t)P * C;
    float* EM = ws;                      ws += (size_t)P * 3;
    MV2D_CHECK_ARG((size_t)(ws - p.workspace) * sizeof(float) <= p.workspace_bytes,
                   "pe3d: workspace too small (%zu needed)", (size_t)(ws - p.workspace) * sizeof(float));
    int rc;
    MV2D_CHECK_ARG(p.phase >= 0 && p.phase <= 2, "pe3d: phase must be 0, 1 or 2");
    MV2D_CHECK_ARG(p.phase == 0 || !p.sine_branch_cached, "pe3d: phases cannot be combined with a cached sine branch");
    if (p.phase != 2) {
        long long total = (long long)P * D;
        MV2D_CHECK_ARG(total < (1LL << 31), "pe3d: V h w D = %lld must be below 2^31", total);
        launch_k(pe_coords_kernel, dim3((unsigned)cdiv((int)total, 256)), dim3(256), 0, st, 
            p.img2lidar, A1, p.V, p.h, p.w, D, (double)p.pad_h / (double)p.h, (double)p.pad_w / (double)p.w, (double)p.depth_start,
            (double)p.position_range[0], (double)p.position_range[1], (double)p.position_range[2],
            ((double)p.position_range[3] - (double)p.depth_start) / ((double)D * (1.0 + (double)D)),
            1.0 / ((double)p.position_range[3] - (double)p.position_range[0]),
            1.0 / ((double)p.position_range[4] - (double)p.position_range[1]),
            1.0 / ((double)p.position_range[5] - (double)p.position_range[2]), 1);
        MV2D_CHECK_LAUNCH("pe_coords");
    }
    // Fused MLPs (mlp2.cu): the 1024-wide hidden activations of the two branches and the SE gate's hidden stay in
    // TMEM / shared memory.  MV2D_PE_FUSED=0 keeps one tcgen05 GEMM per layer (hidden round trip through HBM).
    static const bool fused_env = []() { const char* v = getenv("MV2D_PE_FUSED"); return !(v && v[0] == '0'); }();
    const bool fused = fused_env && !p.unfused_mlp && (3 * D) % 32 == 0;
    // position_encoder: 192 -> 1024 -> 256
    if (p.phase != 2) {
        if (fused) {
            Mlp2 m{};
            m.A = A1; m.lda = 3 * D; m.W0 = p.w_pos0; m.b0 = p.b_pos0; m.W2 = p.w_pos2; m.b2 = p.b_pos2; m.out = X;
            m.M = P; m.K0 = 3 * D; m.H = 4 * C;
            if ((rc = launch_mlp2(m, st))) return rc;
        } else {
    if ((rc = gemm(A1, 3 * D, p.w_pos0, 3 * D, p.b_pos0, Hd, 4 * C, P, 4 * C, 3 * D, GEMM_RELU | GEMM_TF32_OK | GEMM_ROUND_TF32, st))) return rc;
    if ((rc = gemm(Hd, 4 * C, p.w_pos2, 4 * C, p.b_pos2, X, C, P, C, 4 * C, GEMM_TF32_OK, st))) return rc;
        }
    }
    // sine branch: 384 -> 1024 -> 256  (input-independent given masks + weights; recomputed here)
    if (p.phase == 2) {
        // X and SB were left in the workspace by phase 1
    } else if (!p.sine_branch_cached && p.sine_separable) {
        // no padded cells: per-axis tables instead of the 384 -> 1024 GEMM over every cell
        const int rows = vps + p.h + p.w;
        float* Fm = S;                              // [rows, 384]
        float* Tm = S + (size_t)rows * 384;         // [rows, 1024]
        MV2D_CHECK_ARG((size_t)rows * (384 + 1024) <= (size_t)P * 384, "pe3d: separable sine tables do not fit the workspace");
        launch_k(sine_axis_kernel, dim3(rows), dim3(128), 0, st, p.dim_t, Fm, vps, p.h, p.w, (float)p.stride, 6.283185307179586f, 1e-6f);
        MV2D_CHECK_LAUNCH("sine_axis");
        if ((rc = gemm(Fm, 384, p.w_adapt0, 384, nullptr, Tm, 4 * C, rows, 4 * C, 384, 0, st))) return rc;     // fp32 FFMA
        const int KH = (rows + 31) / 32 * 32;
        // measured on B200 (profiles/r02_pe_fused.md): as a one-hot first GEMM the gather costs as much L2 -> SM traffic as a
        // real layer (47 us for V = 6 against 25 + 25 for gather kernel + GEMM), so it is opt-in: MV2D_PE_FUSED_SINE=1
        static const bool fused_sine = []() { const char* v = getenv("MV2D_PE_FUSED_SINE"); return v && v[0] == '1'; }();
        if (fused && fused_sine && (size_t)Ps * 2 * KH + (size_t)1024 * 2 * KH <= (size_t)P * 4 * C) {
            float* Aoh = Hd;                                 // [Ps, 2 KH] one-hot rows
            float* W0s = Hd + (size_t)Ps * 2 * KH;           // [1024, 2 KH] transposed tables, hi | lo
            launch_k(sine_w0_kernel, dim3(1024 / 32, KH / 32), dim3(32, 8), 0, st, (const float*)Tm, W0s, rows, KH);
            MV2D_CHECK_LAUNCH("sine_w0");
            launch_k(sine_onehot_kernel, dim3((unsigned)(((long long)Ps * (KH >> 1) + 255) / 256)), dim3(256), 0, st, Aoh, Ps, p.h, p.w, vps, KH);
            MV2D_CHECK_LAUNCH("sine_onehot");
            Mlp2 m{};
            m.A = Aoh; m.lda = 2 * KH; m.W0 = W0s; m.b0 = p.b_adapt0; m.W2 = p.w_adapt2; m.b2 = p.b_adapt2; m.out = SB;
            m.M = Ps; m.K0 = 2 * KH; m.H = 4 * C;
            if ((rc = launch_mlp2(m, st))) return rc;
        } else {
        launch_k(sine_hidden_kernel, dim3((unsigned)(((long long)Ps * 256 + 255) / 256)), dim3(256), 0, st, (const float*)Tm, p.b_adapt0, Hd, Vs, p.h, p.w, vps);
        MV2D_CHECK_LAUNCH("sine_hidden");
        if ((rc = gemm(Hd, 4 * C, p.w_adapt2, 4 * C, p.b_adapt2, SB, C, Ps, C, 4 * C, GEMM_TF32_OK, st))) return rc;
        }
    } else if (!p.sine_branch_cached) {
        launch_k(sine_prep_kernel, dim3(cdiv(Ps, 128)), dim3(128), 0, st, p.not_mask, EM, Vs, p.h, p.w, (float)p.stride,
                                                      6.283185307179586f, 1e-6f, vps);
        MV2D_CHECK_LAUNCH("sine_prep");
        launch_k(sine_embed_kernel, dim3((unsigned)(((long long)Ps * 384 + 255) / 256)), dim3(256), 0, st, (const float*)EM, p.dim_t, S, Ps, 1);
        MV2D_CHECK_LAUNCH("sine_embed");
        if ((rc = gemm(S, 384, p.w_adapt0, 384, p.b_adapt0, Hd, 4 * C, Ps, 4 * C, 384, GEMM_RELU | GEMM_TF32_OK | GEMM_ROUND_TF32, st))) return rc;
        if ((rc = gemm(Hd, 4 * C, p.w_adapt2, 4 * C, p.b_adapt2, SB, C, Ps, C, 4 * C, GEMM_TF32_OK, st))) return rc;
    } else {
        SB = const_cast<float*>(p.sine_branch_cached);
    }
    if (p.phase == 1) return 0;   // everything that does not read the image feature is done
    // SE gate on the image feature, fused combine: pe = X * sigmoid(gate) + SB ; kin = pe + feat
    if (fused && (!p.sine_shared || Ps % 128 == 0)) {
        Mlp2 m{};
        m.A = p.feat_tf32 ? p.feat_tf32 : p.feat; m.lda = C; m.W0 = p.w_se_reduce; m.b0 = p.b_se_reduce;
        m.W2 = p.w_se_expand; m.b2 = p.b_se_expand; m.out = p.pe; m.M = P; m.K0 = C; m.H = C;
        m.gate = 1; m.gx = X; m.gs = SB; m.gs_mod = p.sine_shared ? Ps : 0; m.gfeat = p.feat; m.kin = p.kin;
        if ((rc = launch_mlp2(m, st))) return rc;
    } else {
    if ((rc = gemm(p.feat_tf32 ? p.feat_tf32 : p.feat, C, p.w_se_reduce, C, p.b_se_reduce, G1, C, P, C, C,
                   GEMM_RELU | GEMM_TF32_OK | GEMM_ROUND_TF32, st))) return rc;
    if ((rc = gemm(G1, C, p.w_se_expand, C, p.b_se_expand, p.pe, C, P, C, C, GEMM_GATE | GEMM_TF32_OK, st, X, SB,
                   p.feat, p.kin, p.sine_shared ? Ps : 0))) return rc;
    }
    if (p.sine_branch_out && !p.sine_branch_cached) {
        cudaError_t e = cudaMemcpyAsync(p.sine_branch_out, SB, (size_t)P * C * sizeof(float),
                                        cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) { set_error("pe3d: memcpy %s", cudaGetErrorString(e)); return (int)e; }
    }
    return 0;
}

// The parameter-free inputs of PE.forward for the training path (train.cu): frustum coordinates [P,3*D] after
// inverse_sigmoid and the 384 sine features [P,384], both as plain fp32 (the inference path rounds them to TF32).
// `sine` must have room for P*384 + 3*P floats: the three normalised embeds per cell are staged behind the features.
int run_pe_train_inputs(int V, int h, int w, int D, int pad_h, int pad_w, int stride, double depth_start, const double* pr,
                        const double* img2lidar, const uint8_t* not_mask, const float* dim_t, float* coords, float* sine,
                        cudaStream_t st) {
    const int P = V * h * w;
    MV2D_CHECK_ARG(V >= 1 && V <= MV2D_MAXV && P > 0 && D > 0, "pe_train_inputs: bad V/h/w/D");
    const long long total = (long long)P * D;
    MV2D_CHECK_ARG(total < (1LL << 31), "pe_train_inputs: V h w D = %lld must be below 2^31", total);
    launch_k(pe_coords_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, img2lidar, coords, V, h, w, D, (double)pad_h / (double)h,
             (double)pad_w / (double)w, (double)depth_start, (double)pr[0], (double)pr[1], (double)pr[2],
             ((double)pr[3] - (double)depth_start) / ((double)D * (1.0 + (double)D)), 1.0 / ((double)pr[3] - (double)pr[0]),
             1.0 / ((double)pr[4] - (double)pr[1]), 1.0 / ((double)pr[5] - (double)pr[2]), 0);
    MV2D_CHECK_LAUNCH("pe_coords(train)");
    float* emb = sine + (size_t)P * 384;     // the caller's `sine` buffer holds [P,384] + 3 P floats for the embeds
    launch_k(sine_prep_kernel, dim3(cdiv(P, 128)), dim3(128), 0, st, not_mask, emb, V, h, w, (float)stride, 6.283185307179586f, 1e-6f, V);
    MV2D_CHECK_LAUNCH("sine_prep(train)");
    launch_k(sine_embed_kernel, dim3((unsigned)(((long long)P * 384 + 255) / 256)), dim3(256), 0, st, (const float*)emb, dim_t, sine, P, 0);
    MV2D_CHECK_LAUNCH("sine_embed(train)");
    return 0;
}

size_t pe3d_workspace_bytes(int V, int h, int w, int depth_num) {
    size_t P = (size_t)V * h * w;
    return P * (3 * depth_num + 4 * MV2D_C + MV2D_C + MV2D_C + 384 + MV2D_C + 3) * sizeof(float);
}

}  // namespace mv2d
