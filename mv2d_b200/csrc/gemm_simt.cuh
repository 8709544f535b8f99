// Generic fp32 FFMA GEMM:  C[M,N] = epi( A[M,K] . W[N,K]^T )   (both operands K-contiguous,
// i.e. nn.Linear / 1x1-conv weight layout), register-blocked, double-buffered through shared
// memory.  Used for the precision-critical and the small (M ~ 300) contractions of the path,
// where fp32 accuracy is required by the parity gate and tensor-core tiles
// would be mostly padding.  The big PE / conv contractions go through gemm_tc.cu instead.
#pragma once
#include "common.cuh"

namespace mv2d {

enum GemmFlags : int {
    GEMM_RELU = 1,        // max(x, 0)
    GEMM_CLAMP5E3 = 2,    // min(x, 5e3) after relu  (query_generator.py:369 clamp on the cat)
    GEMM_GATE = 4,        // out = gx * sigmoid(acc+bias) + gs ; kin = out + gfeat   (pe.py:44-48,166)
    GEMM_TF32_OK = 8,     // caller allows single-pass TF32 tensor cores (PE MLPs only)
};

enum GemmAMode : int { A_PLAIN = 0, A_IM2COL3X3 = 1 };

struct GemmArgs {
    const float* A; int lda; long long strideA;
    const float* W; int ldw; long long strideW;
    float* C; int ldc; long long strideC;
    const float* bias; long long strideBias;
    int M, N, K;
    int batch;            // blockIdx.z = b * nsplit + s
    int nsplit;           // split-K: split s writes RAW partial sums to C + s * splitStride
    long long splitStride;
    int flags;
    const float* gx; const float* gs; const float* gfeat; float* kin;  // GEMM_GATE extras (ld = ldc)
    int gs_mod;           // GEMM_GATE: > 0 = gs has gs_mod rows and is read at row (m % gs_mod) (batch-shared sine branch)
};

template <int BM, int BN, int BK, int RM, int RN, int AMODE>
__global__ void __launch_bounds__((BM / (4 * RM)) * (BN / (4 * RN)))
gemm_simt_kernel(GemmArgs g) {
    pdl_wait();
    pdl_trigger();
    constexpr int TY = BM / (4 * RM), TX = BN / (4 * RN), NT = TY * TX;
    constexpr int KQ = BK / 4;
    constexpr int A_F4 = BM * KQ, W_F4 = BN * KQ;
    constexpr int A_PER = (A_F4 + NT - 1) / NT, W_PER = (W_F4 + NT - 1) / NT;
    constexpr int PAD = 4;
    __shared__ float As[2][BK][BM + PAD];
    __shared__ float Ws[2][BK][BN + PAD];

    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;
    const int b = blockIdx.z / g.nsplit, split = blockIdx.z % g.nsplit;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int Ks = g.K / g.nsplit;           // host guarantees Ks % BK == 0
    const int kbeg = split * Ks;
    const int nk = Ks / BK;

    const float* __restrict__ A = g.A + b * g.strideA;
    const float* __restrict__ W = g.W + b * g.strideW;

    float4 ra[A_PER], rw[W_PER];

    auto load_tiles = [&](int kt) {
        const int k0 = kbeg + kt * BK;
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            int f = tid + i * NT;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (A_F4 % NT == 0 || f < A_F4) {
                int row = f / KQ, kq = f % KQ;
                int m = m0 + row;
                if (m < g.M) {
                    if (AMODE == A_PLAIN) {
                        v = __ldg(reinterpret_cast<const float4*>(A + (long long)m * g.lda + k0 + kq * 4));
                    } else {
                        // implicit im2col over [roi, 7, 7, 256] tokens; K ordered (tap, ci)
                        int roi = m / MV2D_TOK, pos = m - roi * MV2D_TOK;
                        int y = pos / MV2D_ROI, x = pos - y * MV2D_ROI;
                        int tap = k0 / MV2D_C, ci = k0 - tap * MV2D_C + kq * 4;
                        int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
                        if (yy >= 0 && yy < MV2D_ROI && xx >= 0 && xx < MV2D_ROI)
                            v = __ldg(reinterpret_cast<const float4*>(
                                A + ((long long)roi * MV2D_TOK + yy * MV2D_ROI + xx) * MV2D_C + ci));
                    }
                }
            }
            ra[i] = v;
        }
#pragma unroll
        for (int i = 0; i < W_PER; ++i) {
            int f = tid + i * NT;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (W_F4 % NT == 0 || f < W_F4) {
                int row = f / KQ, kq = f % KQ;
                int n = n0 + row;
                if (n < g.N)
                    v = __ldg(reinterpret_cast<const float4*>(W + (long long)n * g.ldw + k0 + kq * 4));
            }
            rw[i] = v;
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            int f = tid + i * NT;
            if (A_F4 % NT == 0 || f < A_F4) {
                int row = f / KQ, kq = f % KQ;
                As[buf][kq * 4 + 0][row] = ra[i].x;
                As[buf][kq * 4 + 1][row] = ra[i].y;
                As[buf][kq * 4 + 2][row] = ra[i].z;
                As[buf][kq * 4 + 3][row] = ra[i].w;
            }
        }
#pragma unroll
        for (int i = 0; i < W_PER; ++i) {
            int f = tid + i * NT;
            if (W_F4 % NT == 0 || f < W_F4) {
                int row = f / KQ, kq = f % KQ;
                Ws[buf][kq * 4 + 0][row] = rw[i].x;
                Ws[buf][kq * 4 + 1][row] = rw[i].y;
                Ws[buf][kq * 4 + 2][row] = rw[i].z;
                Ws[buf][kq * 4 + 3][row] = rw[i].w;
            }
        }
    };

    float acc[4 * RM][4 * RN];
#pragma unroll
    for (int i = 0; i < 4 * RM; ++i)
#pragma unroll
        for (int j = 0; j < 4 * RN; ++j) acc[i][j] = 0.f;

    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tiles(kt + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[4 * RM], w[4 * RN];
#pragma unroll
            for (int r = 0; r < RM; ++r) {
                float4 v = *reinterpret_cast<const float4*>(&As[buf][k][r * (BM / RM) + ty * 4]);
                a[r * 4 + 0] = v.x; a[r * 4 + 1] = v.y; a[r * 4 + 2] = v.z; a[r * 4 + 3] = v.w;
            }
#pragma unroll
            for (int r = 0; r < RN; ++r) {
                float4 v = *reinterpret_cast<const float4*>(&Ws[buf][k][r * (BN / RN) + tx * 4]);
                w[r * 4 + 0] = v.x; w[r * 4 + 1] = v.y; w[r * 4 + 2] = v.z; w[r * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < 4 * RM; ++i)
#pragma unroll
                for (int j = 0; j < 4 * RN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        if (kt + 1 < nk) store_tiles(buf ^ 1);
        __syncthreads();
    }

    // ---- epilogue
    const bool raw = g.nsplit > 1;
    float* __restrict__ C = g.C + (raw ? split * g.splitStride : 0) + b * g.strideC;
    const float* __restrict__ bias = (g.bias && !raw) ? g.bias + b * g.strideBias : nullptr;
    const bool vec = ((g.N & 3) == 0) && ((g.ldc & 3) == 0);
#pragma unroll
    for (int rm = 0; rm < RM; ++rm)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + rm * (BM / RM) + ty * 4 + i;
            if (m >= g.M) continue;
#pragma unroll
            for (int rn = 0; rn < RN; ++rn) {
                const int n = n0 + rn * (BN / RN) + tx * 4;
                if (n >= g.N) continue;
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float x = acc[rm * 4 + i][rn * 4 + j];
                    if (bias && n + j < g.N) x += __ldg(bias + n + j);
                    if (!raw) {
                        if (g.flags & GEMM_RELU) x = fmaxf(x, 0.f);
                        if (g.flags & GEMM_CLAMP5E3) x = fminf(x, 5e3f);
                    }
                    v[j] = x;
                }
                const long long o = (long long)m * g.ldc + n;
                if (!raw && (g.flags & GEMM_GATE)) {
                    // host guarantees vec for GATE
                    float4 xx = __ldg(reinterpret_cast<const float4*>(g.gx + o));
                    float4 ss = __ldg(reinterpret_cast<const float4*>(g.gs + (g.gs_mod > 0 ? (long long)(m % g.gs_mod) * g.ldc + n : o)));
                    v[0] = xx.x * sigmoid_f(v[0]) + ss.x;
                    v[1] = xx.y * sigmoid_f(v[1]) + ss.y;
                    v[2] = xx.z * sigmoid_f(v[2]) + ss.z;
                    v[3] = xx.w * sigmoid_f(v[3]) + ss.w;
                    if (g.kin) {
                        float4 ff = __ldg(reinterpret_cast<const float4*>(g.gfeat + o));
                        *reinterpret_cast<float4*>(g.kin + o) =
                            make_float4(v[0] + ff.x, v[1] + ff.y, v[2] + ff.z, v[3] + ff.w);
                    }
                }
                if (vec) {
                    *reinterpret_cast<float4*>(C + o) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (n + j < g.N) C[o + j] = v[j];
                }
            }
        }
}

// ------------------------------------------------------------------------------------------
// Small-M variant (M ~ 300: every decoder / query-generator GEMM).  These problems are latency
// bound, not FLOP bound: one 32x32 tile per CTA (64 threads, 4x4 outputs each) so that even a
// [300,256]x[256,256] product spreads over 80 CTAs, and a 4-stage cp.async pipeline with
// BK = 32 keeps three 8 KB k-tiles in flight per CTA.  Shared tiles are stored [row][k] with
// the 16-byte chunks XOR-swizzled by (row >> 2) & 7, so both the cp.async stores and the
// LDS.128 reads (4 consecutive k per thread) are bank-conflict free.
// A may switch to a second matrix (A2) for output columns >= n_switch: the self-attention
// in-projection takes q,k from (x + query_pos) and v from x in ONE launch.
struct GemmSmallArgs {
    GemmArgs g;
    const float* A2; int n_switch;    // A2 == nullptr: unused
};

// host-side launchers (gemm_simt.cu / gemm_tc.cu)
int launch_gemm_simt(const GemmArgs& g, int amode, cudaStream_t stream);
// small-M kernel with the optional second A operand
int launch_gemm_small(const GemmArgs& g, const float* A2, int n_switch, cudaStream_t stream);
// routes big TF32-tolerant problems to the tcgen05 kernel, everything else to the FFMA kernel
int launch_gemm_tc_or_simt(const GemmArgs& g, cudaStream_t stream);

}  // namespace mv2d
