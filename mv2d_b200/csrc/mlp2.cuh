// Host interface of the fused two-layer MLP kernel (mlp2.cu).
#pragma once
#include "common.cuh"

namespace mv2d {

struct Mlp2 {
    const float* A; int lda;          // [M, K0], TF32-representable values
    const float* W0; const float* b0; // [H, K0] (TF32-representable), [H]
    const float* W2; const float* b2; // [256, H] (TF32-representable), [256]
    float* out;                       // [M, 256]
    int M, K0, H;
    int round_out;                    // round the result to TF32
    int gate;                         // SE gate / combine epilogue: out = gx * sigmoid(acc + b2) + gs ; kin = out + gfeat
    const float* gx; const float* gs; int gs_mod; const float* gfeat; float* kin;
};

int launch_mlp2(const Mlp2& m, cudaStream_t st);

}  // namespace mv2d
