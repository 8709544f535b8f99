// Shared device/host helpers for libmv2d_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define MV2D_C 256          // embed dims
#define MV2D_HEADS 8
#define MV2D_HD 32          // head dim
#define MV2D_ROI 7
#define MV2D_TOK 49         // 7x7 RoI tokens
#define MV2D_MAXV 16        // max views per sample
#define MV2D_MAXVB 1024      // max views of a whole batch (samples x views)
#define MV2D_CAT_LD 1056    // 1024 FC features + 16 intrinsics, K padded to a multiple of 32

namespace mv2d {

// thread-local last-error text, set by the host wrappers (abi.cu)
void set_error(const char* fmt, ...);
void note_launch();   // counts kernel launches enqueued by this library (mv2d_launch_count)

#define MV2D_CHECK_ARG(cond, ...)                         \
    do {                                                  \
        if (!(cond)) {                                    \
            mv2d::set_error(__VA_ARGS__);                 \
            return -1;                                    \
        }                                                 \
    } while (0)

#define MV2D_CHECK_LAUNCH(what)                                              \
    do {                                                                     \
        cudaError_t e__ = cudaGetLastError();                                \
        mv2d::note_launch();                                                 \
        if (e__ != cudaSuccess) {                                            \
            mv2d::set_error("%s: %s", what, cudaGetErrorString(e__));        \
            return (int)e__;                                                 \
        }                                                                    \
    } while (0)

__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---- Programmatic dependent launch (PDL).  Every kernel of the library starts with pdl_wait()
// (griddepcontrol.wait: blocks until the preceding kernel in the stream has completed and its writes are
// visible) and then lets ITS successor begin launching (griddepcontrol.launch_dependents), so launch
// latency, CTA scheduling and kernel prologues overlap with the predecessor's tail.  The path is a chain
// of ~90 short dependent kernels per sample, so this matters more than any single kernel's speed.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();   // abi.cu: false when MV2D_NO_PDL=1

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    // measured on B200: PDL shortens the eager chain by ~8 %, but programmatic edges inside a captured
    // CUDA graph replay ~5 % slower than plain kernel-node edges, so the attribute is dropped under capture
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    cfg.attrs = attr; cfg.numAttrs = (pdl_enabled() && cap == cudaStreamCaptureStatusNone) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- fp64 4x4 inverse, Gauss-Jordan with partial pivoting (what LAPACK getrf/getri amount to
// for a 4x4; replaces np.linalg.inv pe.py:111 and torch.inverse box_correlation.py:120,176,
// query_generator.py:339).  m and out are row-major.
__device__ inline void inv4x4(const double* m, double* out) {
    double a[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            a[i][j] = m[i * 4 + j];
            a[i][4 + j] = (i == j) ? 1.0 : 0.0;
        }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        int p = c;
        double best = fabs(a[c][c]);
#pragma unroll
        for (int r = c + 1; r < 4; ++r) {
            double v = fabs(a[r][c]);
            if (v > best) { best = v; p = r; }
        }
        if (p != c) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { double t = a[c][j]; a[c][j] = a[p][j]; a[p][j] = t; }
        }
        double inv = 1.0 / a[c][c];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[c][j] *= inv;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (r == c) continue;
            double f = a[r][c];
#pragma unroll
            for (int j = 0; j < 8; ++j) a[r][j] -= f * a[c][j];
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) out[i * 4 + j] = a[i][4 + j];
}

__device__ inline void mat4_mul(const double* a, const double* b, double* c) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j];
            c[i * 4 + j] = s;
        }
}

__device__ inline float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ inline float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Packed fp32 pairs (sm_100a: fma.rn.f32x2 issues two IEEE fp32 FMAs per instruction): for loops bound by instruction issue.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// mmdet inverse_sigmoid(x, eps=1e-5)
__device__ inline float inverse_sigmoid_f(float x) {
    x = fminf(fmaxf(x, 0.f), 1.f);
    float a = fmaxf(x, 1e-5f), b = fmaxf(1.f - x, 1e-5f);
    return logf(a / b);
}
__device__ inline float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

}  // namespace mv2d
