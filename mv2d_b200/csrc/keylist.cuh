// Ordered compaction of the set bits of a per-query key mask into a list of uint16 key ids.
// Shared by the box-correlation mask kernel (roi.cu) and the denoising-query preparation (dn.cu).
#pragma once
#include "common.cuh"

namespace mv2d {

// bits: [words] in shared memory; grp_cnt: [128] ints of shared memory; out: [words*32] u16 in global memory.
// Must be called by every thread of the block (contains __syncthreads).  words <= 4096.
__device__ __forceinline__ void compact_key_bits(const uint32_t* bits, int words, int* grp_cnt, uint16_t* out) {
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31, nw = blockDim.x >> 5;
    const int ngroups = (words + 31) / 32;
    for (int g = warp; g < ngroups; g += nw) {
        const int w = g * 32 + lane;
        const int c = __popc((w < words) ? bits[w] : 0u);
        const int tot = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0) grp_cnt[g] = tot;
    }
    __syncthreads();
    for (int g = warp; g < ngroups; g += nw) {
        int base = 0;
        for (int i = 0; i < g; ++i) base += grp_cnt[i];
        const int w = g * 32 + lane;
        uint32_t b = (w < words) ? bits[w] : 0u;
        const int c = __popc(b);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
        int pos = base + incl - c;
        while (b) { const int bit = __ffs(b) - 1; b &= b - 1; out[pos++] = (uint16_t)(w * 32 + bit); }
    }
}

}  // namespace mv2d
