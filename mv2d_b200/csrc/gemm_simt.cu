#include "gemm_simt.cuh"

namespace mv2d {

template <int BM, int BN, int BK, int RM, int RN>
static int launch_cfg(const GemmArgs& g, int amode, cudaStream_t stream) {
    constexpr int NT = (BM / (4 * RM)) * (BN / (4 * RN));
    dim3 grid(cdiv(g.N, BN), cdiv(g.M, BM), g.batch * g.nsplit);
    if (amode == A_PLAIN)
        gemm_simt_kernel<BM, BN, BK, RM, RN, A_PLAIN><<<grid, NT, 0, stream>>>(g);
    else
        gemm_simt_kernel<BM, BN, BK, RM, RN, A_IM2COL3X3><<<grid, NT, 0, stream>>>(g);
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
    if (g.M >= 512) return launch_cfg<64, 64, 16, 1, 1>(g, amode, stream);
    return launch_cfg<32, 64, 16, 1, 1>(g, amode, stream);
}

}  // namespace mv2d
