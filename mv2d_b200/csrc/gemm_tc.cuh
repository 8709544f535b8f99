// Host interface of the tcgen05 GEMM (gemm_tc.cu).
#pragma once
#include "common.cuh"

namespace mv2d {

enum GemmTcFlags : int {
    GEMM_FORCE_TC = 16,       // mv2d_gemm: run the single-pass tcgen05 kernel whatever the shape heuristics say
    GEMM_ROUND_TF32 = 64,     // epilogue rounds the result to TF32 (it feeds a single-pass TF32 GEMM)
    GEMM_SPLIT_OUT = 256,     // epilogue writes the TF32 hi part to C and the lo part to C_lo (feeds a 3xTF32 GEMM)
};

// round-to-nearest (ties away) fp32 -> tf32, kept in an fp32 container
// (half an ulp added to the bit pattern, low 13 bits masked: bit-identical to cvt.rna.tf32.f32 for every finite input and on
// overflow to infinity -- the same expression the host packer uses --, 2 instructions instead of the ~7 of the emulated cvt on
// sm_100a; NaNs do not reach these epilogues)
__device__ __forceinline__ float round_tf32(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

struct TcGemm {
    const float* A; const float* A_lo; int lda;     // A_lo only for passes == 3
    const float* W; const float* W_lo; int ldw;
    float* C; float* C_lo; int ldc;
    const float* bias;
    int M, N, K;
    int nsplit; long long split_stride;   // split-K: raw partial sums of split z go to C + z * split_stride
    int passes;        // 1 = single-pass TF32 (operands already TF32-representable), 3 = 3xTF32
    int im2col;        // 1: A = [n_rois,7,7,256] tokens, M = 49*n_rois, K = 2304 ordered (tap, c_in)
                       // 2: A = [fm_v,fm_h,fm_w,256] feature map, M = fm_v*fm_h*fm_w, same K order (3x3, padding 1)
    int fm_v, fm_h, fm_w;
    const uint8_t* m_tile_live;   // nullable, device [ceil(M/128)]: row tiles with 0 are skipped (their C rows stay untouched)
    int flags;         // GEMM_RELU | GEMM_GATE | GEMM_ROUND_TF32
    const float* gx; const float* gs; const float* gfeat; float* kin;
    int gs_mod;        // GEMM_GATE: > 0 = gs has gs_mod rows, read at (row % gs_mod)
    int groups;        // > 1 (plain 2-D operands, nsplit == 1): `groups` independent problems of M rows each in one launch
                       // (blockIdx.z): A / C rows of group z start at z * group_rows, W rows at z * N, bias at z * N
                       // (the per-layer branch MLPs: one weight matrix per decoder layer)
    long long group_rows;
    const float* A2; const float* A2_lo; int n_switch;   // nullable: output columns >= n_switch take their A rows from A2
                                                          // (the self-attention in_proj: q, k from x + pos, v from x)
};

int launch_gemm_tc(const TcGemm& t, cudaStream_t st);

// encode a 2-D, K-contiguous, 128B-swizzled tensor map [rows, K] (box = 32 floats x box_rows); `map` points at a CUtensorMap
int tc_make_map_2d(void* map, const float* base, int rows, int K, int ld, int box_rows);

// x -> hi = rna_tf32(x), lo = rna_tf32(x - hi)   (both exactly TF32-representable)
int launch_split_tf32(const float* x, float* hi, float* lo, long long n, cudaStream_t st);

}  // namespace mv2d
