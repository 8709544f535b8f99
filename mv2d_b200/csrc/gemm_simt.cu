#include <stdlib.h>
#include "gemm_simt.cuh"

namespace mv2d {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;   // src-size 0 => zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

#define GS_BM 32
#define GS_BN 32
#define GS_BK 32
#define GS_STAGES 4

__global__ void __launch_bounds__(64)
gemm_small_kernel(GemmSmallArgs a) {
    pdl_wait();
    pdl_trigger();
    const GemmArgs& g = a.g;
    __shared__ __align__(16) float As[GS_STAGES][GS_BM][GS_BK];
    __shared__ __align__(16) float Ws[GS_STAGES][GS_BN][GS_BK];
    const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
    const int b = blockIdx.z / g.nsplit, split = blockIdx.z % g.nsplit;
    const int m0 = blockIdx.y * GS_BM, n0 = blockIdx.x * GS_BN;
    const int Ks = g.K / g.nsplit, kbeg = split * Ks, nk = Ks / GS_BK;
    const float* __restrict__ A = ((a.A2 && n0 >= a.n_switch) ? a.A2 : g.A) + b * g.strideA;
    const float* __restrict__ W = g.W + b * g.strideW;

    // each thread copies 4 chunks of A and 4 of W per k-tile: chunk id f = tid + 64*i -> (row = f/8, kc = f%8)
    auto issue = [&](int kt, int stage) {
        const int k0 = kbeg + kt * GS_BK;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int f = tid + 64 * i, row = f >> 3, kc = f & 7, sw = kc ^ ((row >> 2) & 7);
            const int m = m0 + row, n = n0 + row;
            cp_async16(&As[stage][row][sw * 4], A + (long long)min(m, g.M - 1) * g.lda + k0 + kc * 4, m < g.M);
            cp_async16(&Ws[stage][row][sw * 4], W + (long long)min(n, g.N - 1) * g.ldw + k0 + kc * 4, n < g.N);
        }
    };
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

#pragma unroll
    for (int s = 0; s < GS_STAGES - 1; ++s) {
        if (s < nk) issue(s, s);
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<GS_STAGES - 2>();
        __syncthreads();
        if (kt + GS_STAGES - 1 < nk) issue(kt + GS_STAGES - 1, (kt + GS_STAGES - 1) % GS_STAGES);
        cp_async_commit();
        const int st = kt % GS_STAGES;
#pragma unroll
        for (int kc = 0; kc < 8; ++kc) {
            float4 av[4], wv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int ra = ty * 4 + i, rw = tx * 4 + i;
                av[i] = *reinterpret_cast<const float4*>(&As[st][ra][(kc ^ ((ra >> 2) & 7)) * 4]);
                wv[i] = *reinterpret_cast<const float4*>(&Ws[st][rw][(kc ^ ((rw >> 2) & 7)) * 4]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[i][j] = fmaf(av[i].x, wv[j].x, acc[i][j]);
                    acc[i][j] = fmaf(av[i].y, wv[j].y, acc[i][j]);
                    acc[i][j] = fmaf(av[i].z, wv[j].z, acc[i][j]);
                    acc[i][j] = fmaf(av[i].w, wv[j].w, acc[i][j]);
                }
        }
    }
    cp_async_wait<0>();

    const bool raw = g.nsplit > 1;
    float* __restrict__ C = g.C + (raw ? split * g.splitStride : 0) + b * g.strideC;
    const float* __restrict__ bias = (g.bias && !raw) ? g.bias + b * g.strideBias : nullptr;
    const int n = n0 + tx * 4;
    if (n >= g.N) return;
    const bool vec = ((g.N & 3) == 0) && ((g.ldc & 3) == 0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= g.M) continue;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float x = acc[i][j];
            if (bias && n + j < g.N) x += __ldg(bias + n + j);
            if (!raw) {
                if (g.flags & GEMM_RELU) x = fmaxf(x, 0.f);
                if (g.flags & GEMM_CLAMP5E3) x = fminf(x, 5e3f);
            }
            v[j] = x;
        }
        const long long o = (long long)m * g.ldc + n;
        if (vec) {
            *reinterpret_cast<float4*>(C + o) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n + j < g.N) C[o + j] = v[j];
        }
    }
}



// Same 32x32 tile, but the K range of the CTA is dealt to four 64-thread groups (in-CTA split-K): at M ~ 300 these
// GEMMs are a few hundred CTAs of pure latency -- a 64-thread CTA walks K = 256 in 8 dependent k-tiles; here every
// group walks a quarter of them (for K = 256 both of its k-tiles are in flight from the start), the four partial
// tiles are folded through shared memory and each of the 256 threads finishes one float4 of the tile.
#define SK_GROUPS 4
#define SK_STAGES 3
#define SK_SMEM_BYTES (SK_GROUPS * SK_STAGES * 2 * GS_BM * GS_BK * 4)

__global__ void __launch_bounds__(256)
gemm_sk4_kernel(GemmSmallArgs a) {
    pdl_wait();
    pdl_trigger();
    const GemmArgs& g = a.g;
    extern __shared__ __align__(16) float sk_smem[];
    const int tid = threadIdx.x, grp = tid >> 6, t = tid & 63, tx = t & 7, ty = t >> 3;
    float (*As)[GS_BM][GS_BK] = reinterpret_cast<float (*)[GS_BM][GS_BK]>(sk_smem + grp * SK_STAGES * 2 * GS_BM * GS_BK);
    float (*Ws)[GS_BN][GS_BK] = As + SK_STAGES;
    const int b = blockIdx.z / g.nsplit, split = blockIdx.z % g.nsplit;
    const int m0 = blockIdx.y * GS_BM, n0 = blockIdx.x * GS_BN;
    const int Ks = g.K / g.nsplit, nk = Ks / GS_BK / SK_GROUPS, kbeg = split * Ks + grp * nk * GS_BK;
    const float* __restrict__ A = ((a.A2 && n0 >= a.n_switch) ? a.A2 : g.A) + b * g.strideA;
    const float* __restrict__ W = g.W + b * g.strideW;

    auto issue = [&](int kt, int stage) {
        const int k0 = kbeg + kt * GS_BK;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int f = t + 64 * i, row = f >> 3, kc = f & 7, sw = kc ^ ((row >> 2) & 7);
            const int m = m0 + row, n = n0 + row;
            cp_async16(&As[stage][row][sw * 4], A + (long long)min(m, g.M - 1) * g.lda + k0 + kc * 4, m < g.M);
            cp_async16(&Ws[stage][row][sw * 4], W + (long long)min(n, g.N - 1) * g.ldw + k0 + kc * 4, n < g.N);
        }
    };
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll
    for (int s = 0; s < SK_STAGES - 1; ++s) {
        if (s < nk) issue(s, s);
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<SK_STAGES - 2>();
        __syncthreads();
        if (kt + SK_STAGES - 1 < nk) issue(kt + SK_STAGES - 1, (kt + SK_STAGES - 1) % SK_STAGES);
        cp_async_commit();
        const int st = kt % SK_STAGES;
#pragma unroll
        for (int kc = 0; kc < 8; ++kc) {
            float4 av[4], wv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int ra = ty * 4 + i, rw = tx * 4 + i;
                av[i] = *reinterpret_cast<const float4*>(&As[st][ra][(kc ^ ((ra >> 2) & 7)) * 4]);
                wv[i] = *reinterpret_cast<const float4*>(&Ws[st][rw][(kc ^ ((rw >> 2) & 7)) * 4]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[i][j] = fmaf(av[i].x, wv[j].x, acc[i][j]);
                    acc[i][j] = fmaf(av[i].y, wv[j].y, acc[i][j]);
                    acc[i][j] = fmaf(av[i].z, wv[j].z, acc[i][j]);
                    acc[i][j] = fmaf(av[i].w, wv[j].w, acc[i][j]);
                }
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    // fold the four partial tiles (fixed order => bitwise reproducible): red[grp][row][col], row stride 36 floats
    float* red = sk_smem;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(red + (grp * GS_BM + ty * 4 + i) * 36 + tx * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    __syncthreads();
    const int row = tid >> 3, c4 = tid & 7;
    float4 v = *reinterpret_cast<const float4*>(red + row * 36 + c4 * 4);
#pragma unroll
    for (int q = 1; q < SK_GROUPS; ++q) {
        const float4 y = *reinterpret_cast<const float4*>(red + (q * GS_BM + row) * 36 + c4 * 4);
        v.x += y.x; v.y += y.y; v.z += y.z; v.w += y.w;
    }
    const bool raw = g.nsplit > 1;
    float* __restrict__ C = g.C + (raw ? split * g.splitStride : 0) + b * g.strideC;
    const float* __restrict__ bias = (g.bias && !raw) ? g.bias + b * g.strideBias : nullptr;
    const int m = m0 + row, n = n0 + c4 * 4;
    if (m >= g.M || n >= g.N) return;
    float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (bias && n + j < g.N) o[j] += __ldg(bias + n + j);
        if (!raw) {
            if (g.flags & GEMM_RELU) o[j] = fmaxf(o[j], 0.f);
            if (g.flags & GEMM_CLAMP5E3) o[j] = fminf(o[j], 5e3f);
        }
    }
    const long long off = (long long)m * g.ldc + n;
    if (((g.N & 3) == 0) && ((g.ldc & 3) == 0)) {
        *reinterpret_cast<float4*>(C + off) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (n + j < g.N) C[off + j] = o[j];
    }
}

template <int BM, int BN, int BK, int RM, int RN>
static int launch_cfg(const GemmArgs& g, int amode, cudaStream_t stream) {
    constexpr int NT = (BM / (4 * RM)) * (BN / (4 * RN));
    dim3 grid(cdiv(g.N, BN), cdiv(g.M, BM), g.batch * g.nsplit);
    if (amode == A_PLAIN)
        launch_k(gemm_simt_kernel<BM, BN, BK, RM, RN, A_PLAIN>, grid, dim3(NT), 0, stream, g);
    else
        launch_k(gemm_simt_kernel<BM, BN, BK, RM, RN, A_IM2COL3X3>, grid, dim3(NT), 0, stream, g);
    MV2D_CHECK_LAUNCH("gemm_simt");
    return 0;
}

int launch_gemm_simt(const GemmArgs& g, int amode, cudaStream_t stream) {
    MV2D_CHECK_ARG(g.M >= 0 && g.N > 0 && g.K > 0, "gemm: bad dims M=%d N=%d K=%d", g.M, g.N, g.K);
    if (g.M == 0) return 0;
    MV2D_CHECK_ARG(g.nsplit >= 1 && g.batch >= 1, "gemm: bad nsplit/batch");
    MV2D_CHECK_ARG(g.K % g.nsplit == 0 && (g.K / g.nsplit) % 16 == 0,
                   "gemm: K=%d must split into multiples of 16 (nsplit=%d)", g.K, g.nsplit);
    MV2D_CHECK_ARG((g.lda & 3) == 0 && (g.ldw & 3) == 0, "gemm: lda/ldw must be multiples of 4");
    MV2D_CHECK_ARG(((uintptr_t)g.A & 15) == 0 && ((uintptr_t)g.W & 15) == 0 && ((uintptr_t)g.C & 15) == 0,
                   "gemm: operands must be 16-byte aligned");
    if (g.flags & GEMM_GATE)
        MV2D_CHECK_ARG((g.N & 3) == 0 && (g.ldc & 3) == 0 && g.gx && g.gs, "gemm: gate epilogue needs N%%4==0");
    if (amode == A_IM2COL3X3)
        MV2D_CHECK_ARG(g.K == 9 * MV2D_C && g.M % MV2D_TOK == 0, "gemm: im2col expects K=2304, M=49*n");
    const long long tiles_big = (long long)cdiv(g.M, 128) * cdiv(g.N, 128) * g.batch * g.nsplit;
    if (g.M >= 1024 && g.N >= 128 && tiles_big >= 96) return launch_cfg<128, 128, 8, 2, 2>(g, amode, stream);
    if (g.M >= 1024 || amode != A_PLAIN || (g.flags & GEMM_GATE) || (g.K / g.nsplit) % GS_BK != 0)
        return launch_cfg<64, 64, 16, 1, 1>(g, amode, stream);
    return launch_gemm_small(g, nullptr, 0, stream);
}

int launch_gemm_small(const GemmArgs& g, const float* A2, int n_switch, cudaStream_t stream) {
    MV2D_CHECK_ARG(g.M >= 0 && g.N > 0 && g.K > 0 && g.nsplit >= 1 && g.batch >= 1, "gemm_small: bad dims");
    if (g.M == 0) return 0;
    MV2D_CHECK_ARG(g.K % g.nsplit == 0 && (g.K / g.nsplit) % GS_BK == 0, "gemm_small: K=%d / nsplit=%d must be a multiple of 32", g.K, g.nsplit);
    MV2D_CHECK_ARG((g.lda & 3) == 0 && (g.ldw & 3) == 0, "gemm_small: lda/ldw must be multiples of 4");
    MV2D_CHECK_ARG(!(g.flags & GEMM_GATE), "gemm_small: gate epilogue not supported");
    MV2D_CHECK_ARG(A2 == nullptr || n_switch % GS_BN == 0, "gemm_small: n_switch must be a multiple of 32");
    GemmSmallArgs a{};
    a.g = g; a.A2 = A2; a.n_switch = n_switch;
    dim3 grid(cdiv(g.N, GS_BN), cdiv(g.M, GS_BM), g.batch * g.nsplit);
    static const bool sk_on = []() { const char* v = getenv("MV2D_GEMM_SK4"); return !(v && v[0] == '0'); }();
    if (sk_on && (g.K / g.nsplit) % (GS_BK * SK_GROUPS) == 0) {
        static bool attr_set = false;
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(gemm_sk4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM_BYTES);
            if (e != cudaSuccess) { set_error("gemm_sk4: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
            attr_set = true;
        }
        launch_k(gemm_sk4_kernel, grid, dim3(256), (size_t)SK_SMEM_BYTES, stream, a);
        MV2D_CHECK_LAUNCH("gemm_sk4");
        return 0;
    }
    launch_k(gemm_small_kernel, grid, dim3(64), 0, stream, a);
    MV2D_CHECK_LAUNCH("gemm_small");
    return 0;
}

}  // namespace mv2d
