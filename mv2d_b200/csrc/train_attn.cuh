// Attention kernels of the training step: flattened self-attention (L2 and shared-memory variants) and the sparse per-RoI cross-attention, forward and backward.
// Included by train.cu only (inside namespace mv2d { namespace { ... } }): one translation unit, several files.
#pragma once

// ------------------------------------------------------------------------------------------------ self-attention
// One CTA per query, one warp per head; lane = key.  qkv [N,768] = (q | k | v), q and k from x + query_pos.
// mask (nullable, [N,N] u8, 1 = masked): the denoising groups' attention mask; a masked logit is -inf, its probability 0,
// so the backward kernels (which work from P) need no mask of their own.
__global__ void __launch_bounds__(256) sa_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ P,
                                                     float* __restrict__ attn_o, int N, const uint8_t* __restrict__ mask) {
    pdl_wait();
    pdl_trigger();
    const int i = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float scale = 0.17677669529663687f;   // 1 / sqrt(32)
    float q[THD];
#pragma unroll
    for (int c = 0; c < THD; ++c) q[c] = qkv[(long long)i * 768 + h * THD + c] * scale;
    float* Prow = P + ((long long)h * N + i) * N;
    float mx = -INFINITY;
    for (int j = lane; j < N; j += 32) {
        const float4* kr = reinterpret_cast<const float4*>(qkv + (long long)j * 768 + 256 + h * THD);
        float s = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < THD / 4; ++c4) {
            const float4 k = kr[c4];
            s += q[c4 * 4] * k.x + q[c4 * 4 + 1] * k.y + q[c4 * 4 + 2] * k.z + q[c4 * 4 + 3] * k.w;
        }
        Prow[j] = s;
        mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < N; j += 32) {
        const float e = expf(Prow[j] - mx);
        Prow[j] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    float o[THD];
#pragma unroll
    for (int c = 0; c < THD; ++c) o[c] = 0.f;
    for (int j = lane; j < N; j += 32) {
        const float p = Prow[j] * inv;
        Prow[j] = p;
        const float4* vr = reinterpret_cast<const float4*>(qkv + (long long)j * 768 + 512 + h * THD);
#pragma unroll
        for (int c4 = 0; c4 < THD / 4; ++c4) {
            const float4 v = vr[c4];
            o[c4 * 4] += p * v.x; o[c4 * 4 + 1] += p * v.y; o[c4 * 4 + 2] += p * v.z; o[c4 * 4 + 3] += p * v.w;
        }
    }
    float mine = 0.f;
#pragma unroll
    for (int c = 0; c < THD; ++c) {
        const float r = warp_sum(o[c]);
        if (lane == c) mine = r;
    }
    attn_o[(long long)i * TC_ + h * THD + lane] = mine;
}

// dO [N,256] -> dS (probability-space gradient folded to logits) and dq (rows 0:256 of dqkv)
__global__ void __launch_bounds__(256) sa_bwd_dq_kernel(const float* __restrict__ qkv, const float* __restrict__ P,
                                                        const float* __restrict__ dO, float* __restrict__ dS,
                                                        float* __restrict__ dqkv, int N) {
    pdl_wait();
    pdl_trigger();
    const int i = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float scale = 0.17677669529663687f;
    float go[THD];
#pragma unroll
    for (int c = 0; c < THD; ++c) go[c] = dO[(long long)i * TC_ + h * THD + c];
    const float* Prow = P + ((long long)h * N + i) * N;
    float* Srow = dS + ((long long)h * N + i) * N;
    float D = 0.f;
    for (int j = lane; j < N; j += 32) {
        const float4* vr = reinterpret_cast<const float4*>(qkv + (long long)j * 768 + 512 + h * THD);
        float dp = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < THD / 4; ++c4) {
            const float4 v = vr[c4];
            dp += go[c4 * 4] * v.x + go[c4 * 4 + 1] * v.y + go[c4 * 4 + 2] * v.z + go[c4 * 4 + 3] * v.w;
        }
        Srow[j] = dp;
        D += Prow[j] * dp;
    }
    D = warp_sum(D);
    float dq[THD];
#pragma unroll
    for (int c = 0; c < THD; ++c) dq[c] = 0.f;
    for (int j = lane; j < N; j += 32) {
        const float ds = Prow[j] * (Srow[j] - D);
        Srow[j] = ds;
        const float4* kr = reinterpret_cast<const float4*>(qkv + (long long)j * 768 + 256 + h * THD);
#pragma unroll
        for (int c4 = 0; c4 < THD / 4; ++c4) {
            const float4 k = kr[c4];
            dq[c4 * 4] += ds * k.x; dq[c4 * 4 + 1] += ds * k.y; dq[c4 * 4 + 2] += ds * k.z; dq[c4 * 4 + 3] += ds * k.w;
        }
    }
    float mine = 0.f;
#pragma unroll
    for (int c = 0; c < THD; ++c) {
        const float r = warp_sum(dq[c]);
        if (lane == c) mine = r;
    }
    dqkv[(long long)i * 768 + h * THD + lane] = mine * scale;
}

// one CTA per key j, one warp per head; lane = query.  dk -> dqkv[:, 256:512], dv -> dqkv[:, 512:768]
__global__ void __launch_bounds__(256) sa_bwd_dkv_kernel(const float* __restrict__ qkv, const float* __restrict__ P,
                                                         const float* __restrict__ dS, const float* __restrict__ dO,
                                                         float* __restrict__ dqkv, int N) {
    pdl_wait();
    pdl_trigger();
    const int j = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float scale = 0.17677669529663687f;
    float dk[THD], dv[THD];
#pragma unroll
    for (int c = 0; c < THD; ++c) { dk[c] = 0.f; dv[c] = 0.f; }
    for (int i = lane; i < N; i += 32) {
        const long long o = ((long long)h * N + i) * N + j;
        const float p = P[o], ds = dS[o];
        const float4* qr = reinterpret_cast<const float4*>(qkv + (long long)i * 768 + h * THD);
        const float4* gr = reinterpret_cast<const float4*>(dO + (long long)i * TC_ + h * THD);
#pragma unroll
        for (int c4 = 0; c4 < THD / 4; ++c4) {
            const float4 q = qr[c4], g = gr[c4];
            dk[c4 * 4] += ds * q.x; dk[c4 * 4 + 1] += ds * q.y; dk[c4 * 4 + 2] += ds * q.z; dk[c4 * 4 + 3] += ds * q.w;
            dv[c4 * 4] += p * g.x; dv[c4 * 4 + 1] += p * g.y; dv[c4 * 4 + 2] += p * g.z; dv[c4 * 4 + 3] += p * g.w;
        }
    }
    float mk = 0.f, mv = 0.f;
#pragma unroll
    for (int c = 0; c < THD; ++c) {
        const float rk = warp_sum(dk[c]), rv = warp_sum(dv[c]);
        if (lane == c) { mk = rk; mv = rv; }
    }
    dqkv[(long long)j * 768 + 256 + h * THD + lane] = mk * scale;
    dqkv[(long long)j * 768 + 512 + h * THD + lane] = mv;
}

// ---- shared-memory variants (the default whenever one head's K and V of all N queries fit: N <= 775).  The kernels
// above read every key row from L2 once per (query, head) -- 184 MB of L2 traffic per launch at N = 300, which is what
// bounded them (66 us).  Here a CTA owns one head and 16 queries (or 16 keys), stages the head's two [N,32] operand
// slices in shared memory once (row stride 33 floats: lane = row reads are conflict-free) and the warps walk them.
constexpr int SA_QB = 16;
__device__ __forceinline__ void sa_stage(const float* __restrict__ src, int ld, int col, int N, float* __restrict__ dst) {
    for (int idx = threadIdx.x; idx < N * 8; idx += 256) {
        const int r = idx >> 3, c4 = idx & 7;
        const float4 v = *reinterpret_cast<const float4*>(src + (long long)r * ld + col + c4 * 4);
        float* d = dst + r * 33 + c4 * 4;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
}

__global__ void __launch_bounds__(256) sa_fwd_smem_kernel(const float* __restrict__ qkv, float* __restrict__ P,
                                                          float* __restrict__ attn_o, int N, const uint8_t* __restrict__ mask) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ float sa_sm[];
    float* Ks = sa_sm;
    float* Vs = sa_sm + (size_t)N * 33;
    const int h = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    sa_stage(qkv, 768, 256 + h * THD, N, Ks);
    sa_stage(qkv, 768, 512 + h * THD, N, Vs);
    __syncthreads();
    const float scale = 0.17677669529663687f;
    for (int qi = warp; qi < SA_QB; qi += 8) {
        const int i = blockIdx.x * SA_QB + qi;
        if (i >= N) break;
        float q[THD];
#pragma unroll
        for (int c = 0; c < THD; ++c) q[c] = qkv[(long long)i * 768 + h * THD + c] * scale;
        float* Prow = P + ((long long)h * N + i) * N;
        float mx = -INFINITY;
        for (int j = lane; j < N; j += 32) {
            const float* k = Ks + j * 33;
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < THD; ++c) s = fmaf(q[c], k[c], s);
            if (mask && mask[(long long)i * N + j]) s = -INFINITY;
            Prow[j] = s;
            mx = fmaxf(mx, s);
        }
        mx = warp_max(mx);
        float sum = 0.f;
        for (int j = lane; j < N; j += 32) {
            const float e = expf(Prow[j] - mx);
            Prow[j] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = 1.f / sum;
        float o[THD];
#pragma unroll
        for (int c = 0; c < THD; ++c) o[c] = 0.f;
        for (int j = lane; j < N; j += 32) {
            const float p = Prow[j] * inv;
            Prow[j] = p;
            const float* v = Vs + j * 33;
#pragma unroll
            for (int c = 0; c < THD; ++c) o[c] = fmaf(p, v[c], o[c]);
        }
        float mine = 0.f;
#pragma unroll
        for (int c = 0; c < THD; ++c) {
            const float r = warp_sum(o[c]);
            if (lane == c) mine = r;
        }
        attn_o[(long long)i * TC_ + h * THD + lane] = mine;
    }
}

__global__ void __launch_bounds__(256) sa_bwd_dq_smem_kernel(const float* __restrict__ qkv, const float* __restrict__ P,
                                                             const float* __restrict__ dO, float* __restrict__ dS,
                                                             float* __restrict__ dqkv, int N) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ float sa_sm[];
    float* Ks = sa_sm;
    float* Vs = sa_sm + (size_t)N * 33;
    const int h = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    sa_stage(qkv, 768, 256 + h * THD, N, Ks);
    sa_stage(qkv, 768, 512 + h * THD, N, Vs);
    __syncthreads();
    const float scale = 0.17677669529663687f;
    for (int qi = warp; qi < SA_QB; qi += 8) {
        const int i = blockIdx.x * SA_QB + qi;
        if (i >= N) break;
        float go[THD];
#pragma unroll
        for (int c = 0; c < THD; ++c) go[c] = dO[(long long)i * TC_ + h * THD + c];
        const float* Prow = P + ((long long)h * N + i) * N;
        float* Srow = dS + ((long long)h * N + i) * N;
        float D = 0.f;
        for (int j = lane; j < N; j += 32) {
            const float* v = Vs + j * 33;
            float dp = 0.f;
#pragma unroll
            for (int c = 0; c < THD; ++c) dp = fmaf(go[c], v[c], dp);
            Srow[j] = dp;
            D += Prow[j] * dp;
        }
        D = warp_sum(D);
        float dq[THD];
#pragma unroll
        for (int c = 0; c < THD; ++c) dq[c] = 0.f;
        for (int j = lane; j < N; j += 32) {
            const float ds = Prow[j] * (Srow[j] - D);
            Srow[j] = ds;
            const float* k = Ks + j * 33;
#pragma unroll
            for (int c = 0; c < THD; ++c) dq[c] = fmaf(ds, k[c], dq[c]);
        }
        float mine = 0.f;
#pragma unroll
        for (int c = 0; c < THD; ++c) {
            const float r = warp_sum(dq[c]);
            if (lane == c) mine = r;
        }
        dqkv[(long long)i * 768 + h * THD + lane] = mine * scale;
    }
}

// CTA = (16 keys, head): the head's Q and dO slices of all queries are staged; one warp per key, lane = query
__global__ void __launch_bounds__(256) sa_bwd_dkv_smem_kernel(const float* __restrict__ qkv, const float* __restrict__ P,
                                                              const float* __restrict__ dS, const float* __restrict__ dO,
                                                              float* __restrict__ dqkv, int N) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ float sa_sm[];
    float* Qs = sa_sm;
    float* Gs = sa_sm + (size_t)N * 33;
    const int h = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    sa_stage(qkv, 768, h * THD, N, Qs);
    sa_stage(dO, TC_, h * THD, N, Gs);
    __syncthreads();
    const float scale = 0.17677669529663687f;
    for (int ki = warp; ki < SA_QB; ki += 8) {
        const int j = blockIdx.x * SA_QB + ki;
        if (j >= N) break;
        float dk[THD], dv[THD];
#pragma unroll
        for (int c = 0; c < THD; ++c) { dk[c] = 0.f; dv[c] = 0.f; }
        for (int i = lane; i < N; i += 32) {
            const long long o = ((long long)h * N + i) * N + j;
            const float p = P[o], ds = dS[o];
            const float* q = Qs + i * 33;
            const float* g = Gs + i * 33;
#pragma unroll
            for (int c = 0; c < THD; ++c) { dk[c] = fmaf(ds, q[c], dk[c]); dv[c] = fmaf(p, g[c], dv[c]); }
        }
        float mk = 0.f, mv = 0.f;
#pragma unroll
        for (int c = 0; c < THD; ++c) {
            const float rk = warp_sum(dk[c]), rv = warp_sum(dv[c]);
            if (lane == c) { mk = rk; mv = rv; }
        }
        dqkv[(long long)j * 768 + 256 + h * THD + lane] = mk * scale;
        dqkv[(long long)j * 768 + 512 + h * THD + lane] = mv;
    }
}

inline size_t sa_smem_bytes(int N) { return (size_t)2 * N * 33 * sizeof(float); }
inline bool sa_use_smem(int N) {
    static const bool on = []() { const char* e = getenv("MV2D_TRAIN_SA_SMEM"); return !(e && e[0] == '0'); }();
    return on && sa_smem_bytes(N) <= 200 * 1024;
}
int sa_set_attr() {
    static bool done = false;
    if (done) return 0;
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(sa_fwd_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)) != cudaSuccess ||
        (e = cudaFuncSetAttribute(sa_bwd_dq_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)) != cudaSuccess ||
        (e = cudaFuncSetAttribute(sa_bwd_dkv_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)) != cudaSuccess) {
        set_error("train: self-attention smem attribute: %s", cudaGetErrorString(e));
        return (int)e;
    }
    done = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------ cross-attention
// query i attends to the 49 tokens of every RoI in match[i][0 .. cnt_i); Kp / Vp [N*49,256] projected tokens.
// P [N, 8, PM] with PM = max_match * 49; slot = m * 49 + t.
__global__ void __launch_bounds__(256) xa_fwd_kernel(const float* __restrict__ cq, const float* __restrict__ Kp,
                                                     const float* __restrict__ Vp, const int* __restrict__ match,
                                                     const int* __restrict__ match_cnt, int max_match,
                                                     float* __restrict__ P, float* __restrict__ ctx, int N) {
    pdl_wait();
    pdl_trigger();
    const int i = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int PM = max_match * TTOK;
    const int cnt = min(match_cnt[i], max_match);
    const float scale = 0.17677669529663687f;
    float q[THD];
#pragma unroll
    for (int c = 0; c < THD; ++c) q[c] = cq[(long long)i * TC_ + h * THD + c] * scale;
    float* Prow = P + ((long long)i * TH + h) * PM;
    float mx = -INFINITY;
    for (int m = 0; m < cnt; ++m) {
        const int r = match[i * max_match + m];
        for (int t = lane; t < TTOK; t += 32) {
            const float4* kr = reinterpret_cast<const float4*>(Kp + ((long long)r * TTOK + t) * TC_ + h * THD);
            float s = 0.f;
#pragma unroll
            for (int c4 = 0; c4 < THD / 4; ++c4) {
                const float4 k = kr[c4];
                s += q[c4 * 4] * k.x + q[c4 * 4 + 1] * k.y + q[c4 * 4 + 2] * k.z + q[c4 * 4 + 3] * k.w;
            }
            Prow[m * TTOK + t] = s;
            mx = fmaxf(mx, s);
        }
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int m = 0; m < cnt; ++m)
        for (int t = lane; t < TTOK; t += 32) {
            const float e = expf(Prow[m * TTOK + t] - mx);
            Prow[m * TTOK + t] = e;
            sum += e;
        }
    sum = warp_sum(sum);
    const float inv = cnt > 0 ? 1.f / sum : 0.f;
    float o[THD];
#pragma unroll
    for (int c = 0; c < THD; ++c) o[c] = 0.f;
    for (int m = 0; m < cnt; ++m) {
        const int r = match[i * max_match + m];
        for (int t = lane; t < TTOK; t += 32) {
            const float p = Prow[m * TTOK + t] * inv;
            Prow[m * TTOK + t] = p;
            const float4* vr = reinterpret_cast<const float4*>(Vp + ((long long)r * TTOK + t) * TC_ + h * THD);
#pragma unroll
            for (int c4 = 0; c4 < THD / 4; ++c4) {
                const float4 v = vr[c4];
                o[c4 * 4] += p * v.x; o[c4 * 4 + 1] += p * v.y; o[c4 * 4 + 2] += p * v.z; o[c4 * 4 + 3] += p * v.w;
            }
        }
    }
    float mine = 0.f;
#pragma unroll
    for (int c = 0; c < THD; ++c) {
        const float r = warp_sum(o[c]);
        if (lane == c) mine = r;
    }
    ctx[(long long)i * TC_ + h * THD + lane] = mine;
}

__global__ void __launch_bounds__(256) xa_bwd_dq_kernel(const float* __restrict__ Kp, const float* __restrict__ Vp,
                                                        const float* __restrict__ P, const float* __restrict__ dctx,
                                                        const int* __restrict__ match, const int* __restrict__ match_cnt,
                                                        int max_match, float* __restrict__ dS, float* __restrict__ dcq, int N) {
    pdl_wait();
    pdl_trigger();
    const int i = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int PM = max_match * TTOK;
    const int cnt = min(match_cnt[i], max_match);
    const float scale = 0.17677669529663687f;
    float go[THD];
#pragma unroll
    for (int c = 0; c < THD; ++c) go[c] = dctx[(long long)i * TC_ + h * THD + c];
    const float* Prow = P + ((long long)i * TH + h) * PM;
    float* Srow = dS + ((long long)i * TH + h) * PM;
    float D = 0.f;
    for (int m = 0; m < cnt; ++m) {
        const int r = match[i * max_match + m];
        for (int t = lane; t < TTOK; t += 32) {
            const float4* vr = reinterpret_cast<const float4*>(Vp + ((long long)r * TTOK + t) * TC_ + h * THD);
            float dp = 0.f;
#pragma unroll
            for (int c4 = 0; c4 < THD / 4; ++c4) {
                const float4 v = vr[c4];
                dp += go[c4 * 4] * v.x + go[c4 * 4 + 1] * v.y + go[c4 * 4 + 2] * v.z + go[c4 * 4 + 3] * v.w;
            }
            Srow[m * TTOK + t] = dp;
            D += Prow[m * TTOK + t] * dp;
        }
    }
    D = warp_sum(D);
    float dq[THD];
#pragma unroll
    for (int c = 0; c < THD; ++c) dq[c] = 0.f;
    for (int m = 0; m < cnt; ++m) {
        const int r = match[i * max_match + m];
        for (int t = lane; t < TTOK; t += 32) {
            const float ds = Prow[m * TTOK + t] * (Srow[m * TTOK + t] - D);
            Srow[m * TTOK + t] = ds;
            const float4* kr = reinterpret_cast<const float4*>(Kp + ((long long)r * TTOK + t) * TC_ + h * THD);
#pragma unroll
            for (int c4 = 0; c4 < THD / 4; ++c4) {
                const float4 k = kr[c4];
                dq[c4 * 4] += ds * k.x; dq[c4 * 4 + 1] += ds * k.y; dq[c4 * 4 + 2] += ds * k.z; dq[c4 * 4 + 3] += ds * k.w;
            }
        }
    }
    float mine = 0.f;
#pragma unroll
    for (int c = 0; c < THD; ++c) {
        const float r = warp_sum(dq[c]);
        if (lane == c) mine = r;
    }
    dcq[(long long)i * TC_ + h * THD + lane] = mine * scale;
}

// inverse of the match lists: for every RoI r the (query, list position) pairs that attend to it, in ascending
// order (deterministic summation order in xa_bwd_dkv).  One warp per RoI.
__global__ void __launch_bounds__(32) xa_inverse_kernel(const int* __restrict__ match, const int* __restrict__ match_cnt,
                                                        int max_match, int N, int* __restrict__ inv_cnt, int* __restrict__ inv_list) {
    pdl_wait();
    pdl_trigger();
    const int r = blockIdx.x, lane = threadIdx.x;
    int pos = 0;
    const int total = N * max_match;
    for (int base = 0; base < total; base += 32) {
        const int e = base + lane;
        bool hit = false;
        if (e < total) {
            const int i = e / max_match, m = e % max_match;
            hit = m < min(match_cnt[i], max_match) && match[e] == r;
        }
        const unsigned b = __ballot_sync(0xffffffffu, hit);
        if (hit) {
            const int at = pos + __popc(b & ((1u << lane) - 1u));
            if (at < N) inv_list[(long long)r * N + at] = e;
        }
        pos += __popc(b);
    }
    if (lane == 0) inv_cnt[r] = min(pos, N);
}

// one CTA per key token row (r, t), one warp per head, lane = channel of the head: the RoI's inverse list is short
// (the RoI's own query plus the few that matched it), so the entries are walked sequentially with broadcast loads of
// the two scalars and coalesced 128-byte loads of the query / output-gradient rows -- no shuffles
__global__ void __launch_bounds__(256) xa_bwd_dkv_kernel(const float* __restrict__ cq, const float* __restrict__ dctx,
                                                         const float* __restrict__ P, const float* __restrict__ dS,
                                                         const int* __restrict__ inv_cnt, const int* __restrict__ inv_list,
                                                         int max_match, float* __restrict__ dKp, float* __restrict__ dVp, int ldo, int N) {
    pdl_wait();
    pdl_trigger();
    const int row = blockIdx.x, r = row / TTOK, t = row % TTOK;
    const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int PM = max_match * TTOK;
    const float scale = 0.17677669529663687f;
    const int n = inv_cnt[r];
    float dk = 0.f, dv = 0.f;
    for (int e = 0; e < n; ++e) {
        const int code = __ldg(inv_list + (long long)r * N + e);
        const int i = code / max_match, m = code % max_match;
        const long long o = ((long long)i * TH + h) * PM + m * TTOK + t;
        const float p = __ldg(P + o), ds = __ldg(dS + o);
        dk = fmaf(ds, __ldg(cq + (long long)i * TC_ + h * THD + lane), dk);
        dv = fmaf(p, __ldg(dctx + (long long)i * TC_ + h * THD + lane), dv);
    }
    dKp[(long long)row * ldo + h * THD + lane] = dk * scale;      // ldo: all layers' gradients side by side, [N*49, L*256]
    dVp[(long long)row * ldo + h * THD + lane] = dv;
}

// ------------------------------------------------------------------------------------------------ cross-attention, two-frame head
// Keys = the R feature cells (projected once per layer: Kp / Vp [R,256]); query i attends to key_list[i][0 .. cnt_i)
// (ascending cell ids = the set bits of its key mask; denoising queries carry the union of all masks).
// P / dS are stored DENSELY indexed, [NT, 8, R] -- only the entries of a query's keys are ever written or read -- so the
// key-side backward finds the entry of (query, key) without ranking the key inside the query's list.
// One CTA per query, one warp per head, lane = key slot (same structure as xa_fwd_kernel).
__global__ void __launch_bounds__(256) xt_train_fwd_kernel(const float* __restrict__ cq, const float* __restrict__ Kp,
                                                           const float* __restrict__ Vp, const uint16_t* __restrict__ key_list,
                                                           const int* __restrict__ key_cnt, int klist_ld, int R,
                                                           float* __restrict__ P, float* __restrict__ ctx) {
    pdl_wait();
    pdl_trigger();
    const int i = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cnt = key_cnt[i];
    const uint16_t* kl = key_list + (long long)i * klist_ld;
    const float scale = 0.17677669529663687f;
    float q[THD];
#pragma unroll
    for (int c = 0; c < THD; ++c) q[c] = cq[(long long)i * TC_ + h * THD + c] * scale;
    float* Prow = P + ((long long)i * TH + h) * R;
    float mx = -INFINITY;
    for (int j = lane; j < cnt; j += 32) {
        const int k = kl[j];
        const float4* kr = reinterpret_cast<const float4*>(Kp + (long long)k * TC_ + h * THD);
        float s = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < THD / 4; ++c4) {
            const float4 kk = kr[c4];
            s += q[c4 * 4] * kk.x + q[c4 * 4 + 1] * kk.y + q[c4 * 4 + 2] * kk.z + q[c4 * 4 + 3] * kk.w;
        }
        Prow[k] = s;
        mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < cnt; j += 32) {
        const int k = kl[j];
        const float e = expf(Prow[k] - mx);
        Prow[k] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    const float inv = cnt > 0 ? 1.f / sum : 0.f;
    float o[THD];
#pragma unroll
    for (int c = 0; c < THD; ++c) o[c] = 0.f;
    for (int j = lane; j < cnt; j += 32) {
        const int k = kl[j];
        const float p = Prow[k] * inv;
        Prow[k] = p;
        const float4* vr = reinterpret_cast<const float4*>(Vp + (long long)k * TC_ + h * THD);
#pragma unroll
        for (int c4 = 0; c4 < THD / 4; ++c4) {
            const float4 v = vr[c4];
            o[c4 * 4] += p * v.x; o[c4 * 4 + 1] += p * v.y; o[c4 * 4 + 2] += p * v.z; o[c4 * 4 + 3] += p * v.w;
        }
    }
    float mine = 0.f;
#pragma unroll
    for (int c = 0; c < THD; ++c) {
        const float r = warp_sum(o[c]);
        if (lane == c) mine = r;
    }
    ctx[(long long)i * TC_ + h * THD + lane] = mine;
}

__global__ void __launch_bounds__(256) xt_train_bwd_dq_kernel(const float* __restrict__ Kp, const float* __restrict__ Vp,
                                                              const float* __restrict__ P, const float* __restrict__ dctx,
                                                              const uint16_t* __restrict__ key_list, const int* __restrict__ key_cnt,
                                                              int klist_ld, int R, float* __restrict__ dS, float* __restrict__ dcq) {
    pdl_wait();
    pdl_trigger();
    const int i = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cnt = key_cnt[i];
    const uint16_t* kl = key_list + (long long)i * klist_ld;
    const float scale = 0.17677669529663687f;
    float go[THD];
#pragma unroll
    for (int c = 0; c < THD; ++c) go[c] = dctx[(long long)i * TC_ + h * THD + c];
    const float* Prow = P + ((long long)i * TH + h) * R;
    float* Srow = dS + ((long long)i * TH + h) * R;
    float D = 0.f;
    for (int j = lane; j < cnt; j += 32) {
        const int k = kl[j];
        const float4* vr = reinterpret_cast<const float4*>(Vp + (long long)k * TC_ + h * THD);
        float dp = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < THD / 4; ++c4) {
            const float4 v = vr[c4];
            dp += go[c4 * 4] * v.x + go[c4 * 4 + 1] * v.y + go[c4 * 4 + 2] * v.z + go[c4 * 4 + 3] * v.w;
        }
        Srow[k] = dp;
        D += Prow[k] * dp;
    }
    D = warp_sum(D);
    float dq[THD];
#pragma unroll
    for (int c = 0; c < THD; ++c) dq[c] = 0.f;
    for (int j = lane; j < cnt; j += 32) {
        const int k = kl[j];
        const float ds = Prow[k] * (Srow[k] - D);
        Srow[k] = ds;
        const float4* kr = reinterpret_cast<const float4*>(Kp + (long long)k * TC_ + h * THD);
#pragma unroll
        for (int c4 = 0; c4 < THD / 4; ++c4) {
            const float4 kk = kr[c4];
            dq[c4 * 4] += ds * kk.x; dq[c4 * 4 + 1] += ds * kk.y; dq[c4 * 4 + 2] += ds * kk.z; dq[c4 * 4 + 3] += ds * kk.w;
        }
    }
    float mine = 0.f;
#pragma unroll
    for (int c = 0; c < THD; ++c) {
        const float r = warp_sum(dq[c]);
        if (lane == c) mine = r;
    }
    dcq[(long long)i * TC_ + h * THD + lane] = mine * scale;
}

// one CTA per key cell, one warp per head, lane = channel: the queries are walked in ascending order (deterministic
// summation) and the key-mask bit of (query, cell) decides -- warp-uniformly -- whether the query attends to the cell
__global__ void __launch_bounds__(256) xt_train_bwd_dkv_kernel(const float* __restrict__ cq, const float* __restrict__ dctx,
                                                               const float* __restrict__ P, const float* __restrict__ dS,
                                                               const uint32_t* __restrict__ keymask, int mask_words, int NT, int R,
                                                               float* __restrict__ dKp, float* __restrict__ dVp, int ldo) {
    pdl_wait();
    pdl_trigger();
    const int k = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int word = k >> 5;
    const uint32_t bit = 1u << (k & 31);
    const float scale = 0.17677669529663687f;
    float dk = 0.f, dv = 0.f;
    for (int i = 0; i < NT; ++i) {
        if (!(__ldg(keymask + (long long)i * mask_words + word) & bit)) continue;
        const long long o = ((long long)i * TH + h) * R + k;
        const float p = __ldg(P + o), ds = __ldg(dS + o);
        dk = fmaf(ds, __ldg(cq + (long long)i * TC_ + h * THD + lane), dk);
        dv = fmaf(p, __ldg(dctx + (long long)i * TC_ + h * THD + lane), dv);
    }
    dKp[(long long)k * ldo + h * THD + lane] = dk * scale;
    dVp[(long long)k * ldo + h * THD + lane] = dv;
}
