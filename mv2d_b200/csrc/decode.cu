// f1 ("next" row of SURVEY.md section 8f): NMSFreeCoder.decode_single + get_bboxes z-shift.
//   reference: core/bbox/coders/nms_free_coder.py:49-102, core/bbox/util.py:60-87,
//              roi_heads/bbox_heads/cross_attention_head.py:372
// One CTA: sigmoid scores of the N*10 logits are bitonic-sorted (descending, ties by lower
// flat index) in shared memory; the first max_num entries are decoded.
#include "common.cuh"
#include "mv2d_internal.h"

namespace mv2d {

__global__ void __launch_bounds__(1024)
nms_free_decode_kernel(const float* __restrict__ cls, const float* __restrict__ box, int total, int npow2,
                       int max_num, float r0, float r1, float r2, float r3, float r4, float r5,
                       float* __restrict__ out_boxes, float* __restrict__ out_scores, int* __restrict__ out_labels,
                       uint8_t* __restrict__ out_valid) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ unsigned char raw[];
    float* key = reinterpret_cast<float*>(raw);
    int* idx = reinterpret_cast<int*>(key + npow2);
    for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
        key[i] = i < total ? sigmoid_f(cls[i]) : -1.f;
        idx[i] = i;
    }
    __syncthreads();
    for (int k = 2; k <= npow2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
                const int p = i ^ j;
                if (p > i) {
                    const bool desc = (i & k) == 0;
                    const float a = key[i], b = key[p];
                    const int ia = idx[i], ib = idx[p];
                    const bool a_first = (a > b) || (a == b && ia < ib);   // a should precede b in descending order
                    if (desc ? !a_first : a_first) { key[i] = b; key[p] = a; idx[i] = ib; idx[p] = ia; }
                }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < max_num; i += blockDim.x) {
        if (i >= total) { out_valid[i] = 0; out_scores[i] = 0.f; out_labels[i] = 0; continue; }
        const int f = idx[i], q = f / 10;
        const float* b = box + (long long)q * 10;
        float o[9];
        o[0] = b[0]; o[1] = b[1]; o[2] = b[4];
        o[3] = expf(b[2]); o[4] = expf(b[3]); o[5] = expf(b[5]);
        o[6] = atan2f(b[6], b[7]); o[7] = b[8]; o[8] = b[9];
        const bool ok = o[0] >= r0 && o[1] >= r1 && o[2] >= r2 && o[0] <= r3 && o[1] <= r4 && o[2] <= r5;
        o[2] = o[2] - o[5] * 0.5f;
        for (int c = 0; c < 9; ++c) out_boxes[i * 9 + c] = o[c];
        out_scores[i] = key[i];
        out_labels[i] = f % 10;
        out_valid[i] = ok ? 1 : 0;
    }
}

int run_nms_free_decode(const float* cls, const float* box, int N, int max_num, const float* post_range,
                        float* out_boxes, float* out_scores, int* out_labels, uint8_t* out_valid, cudaStream_t st) {
    MV2D_CHECK_ARG(N >= 1 && max_num >= 1, "nms_free_decode: bad N/max_num");
    const int total = N * 10;
    int npow2 = 1;
    while (npow2 < total) npow2 <<= 1;
    const size_t smem = (size_t)npow2 * 8;
    MV2D_CHECK_ARG(smem <= 200 * 1024, "nms_free_decode: N=%d too large for the single-CTA sort", N);
    cudaError_t e = cudaFuncSetAttribute(nms_free_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("nms_free_decode: %s", cudaGetErrorString(e)); return (int)e; }
    launch_k(nms_free_decode_kernel, dim3(1), dim3(1024), smem, st, cls, box, total, npow2, max_num, post_range[0], post_range[1],
                                                 post_range[2], post_range[3], post_range[4], post_range[5],
                                                 out_boxes, out_scores, out_labels, out_valid);
    MV2D_CHECK_LAUNCH("nms_free_decode");
    return 0;
}

// Scene-level tail of MV2D.simple_test (detectors/mv2d.py:266-282): mmdet3d box3d_multiclass_nms on the decoded
// boxes.  With the configs' nms_thr = 1.0 the rotated BEV NMS suppresses nothing (an IoU never exceeds 1), so what
// the call does is: keep score > score_thr, regroup by class (ascending), each class by descending score; and if
// more than max_num boxes remain, keep the max_num best of all classes in descending score order.
// One CTA, rank by counting (n is at most a few hundred).
__global__ void __launch_bounds__(1024)
scene_nms_kernel(const float* __restrict__ boxes, const float* __restrict__ scores, const int* __restrict__ labels,
                 const uint8_t* __restrict__ valid, int n, float score_thr, int max_num,
                 float* __restrict__ out_boxes, float* __restrict__ out_scores, int* __restrict__ out_labels,
                 int* __restrict__ out_count) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ unsigned char raw[];
    float* sc = reinterpret_cast<float*>(raw);
    int* lb = reinterpret_cast<int*>(sc + n);          // label, or -1 = dropped
    __shared__ int kept_s;
    if (threadIdx.x == 0) kept_s = 0;
    __syncthreads();
    int kept_local = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const bool k = (valid == nullptr || valid[i]) && scores[i] > score_thr;
        sc[i] = scores[i];
        lb[i] = k ? labels[i] : -1;
        kept_local += k;
    }
    atomicAdd(&kept_s, kept_local);
    __syncthreads();
    const int kept = kept_s;
    const bool by_class = kept <= max_num;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (lb[i] < 0) continue;
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            if (lb[j] < 0) continue;
            const bool score_first = sc[j] > sc[i] || (sc[j] == sc[i] && j < i);
            rank += by_class ? (lb[j] < lb[i] || (lb[j] == lb[i] && score_first)) : score_first;
        }
        if (rank < max_num) {
            for (int c = 0; c < 9; ++c) out_boxes[rank * 9 + c] = boxes[i * 9 + c];
            out_scores[rank] = sc[i];
            out_labels[rank] = lb[i];
        }
    }
    if (threadIdx.x == 0) *out_count = kept < max_num ? kept : max_num;
}

int run_scene_nms(const float* boxes, const float* scores, const int* labels, const uint8_t* valid, int n, float score_thr,
                  float nms_thr, int max_num, float* out_boxes, float* out_scores, int* out_labels, int* out_count,
                  cudaStream_t st) {
    MV2D_CHECK_ARG(n >= 0 && max_num >= 1, "scene_nms: bad n/max_num");
    MV2D_CHECK_ARG(nms_thr >= 1.f, "scene_nms: rotated BEV NMS below an IoU threshold of 1.0 is not implemented "
                                   "(the MV2D configs use nms_thr = 1.0, which suppresses nothing)");
    MV2D_CHECK_ARG((size_t)n * 8 <= 200 * 1024, "scene_nms: n=%d too large", n);
    if (n == 0) {
        cudaError_t e = cudaMemsetAsync(out_count, 0, sizeof(int), st);
        if (e != cudaSuccess) { set_error("scene_nms: %s", cudaGetErrorString(e)); return (int)e; }
        return 0;
    }
    const size_t smem = (size_t)n * 8;
    cudaError_t e = cudaFuncSetAttribute(scene_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 48 * 1024 ? smem : 48 * 1024));
    if (e != cudaSuccess) { set_error("scene_nms: %s", cudaGetErrorString(e)); return (int)e; }
    launch_k(scene_nms_kernel, dim3(1), dim3(1024), smem, st, boxes, scores, labels, valid, n, score_thr, max_num,
             out_boxes, out_scores, out_labels, out_count);
    MV2D_CHECK_LAUNCH("scene_nms");
    return 0;
}

// debug: SM clock rate as seen by a running kernel: spin `cycles` SM clocks, report elapsed globaltimer ns
__global__ void clock_probe_kernel(long long cycles, long long* out) {
    pdl_wait();
    pdl_trigger();
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    const long long c0 = clock64();
    while (clock64() - c0 < cycles) {}
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = (long long)(t1 - t0); out[1] = cycles; }
}
int run_clock_probe(long long cycles, long long* out, cudaStream_t st) {
    launch_k(clock_probe_kernel, dim3(1), dim3(32), 0, st, cycles, out);
    MV2D_CHECK_LAUNCH("clock_probe");
    return 0;
}

}  // namespace mv2d
