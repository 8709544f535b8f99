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

// ---- tiled variants (default).  The per-query kernels above stream every query's K / V rows from L2 on their own: with
// denoising queries (each attends to the UNION of all key masks, ~25 k cells) that is 15 GB of L2 traffic per layer.
// Here a tile of 128 keys of one head is staged in shared memory once per CTA and serves a block of 32 queries
// (forward, query-side backward), or a tile of 128 keys accumulates over all queries (key-side backward); tiles / word
// groups without a set mask bit are skipped warp-uniformly.
#define XT2_QB 8           // queries per CTA (one per warp: the dense denoising rows need the parallelism more than the reuse)
#define XT2_KT 128         // keys per tile
#define XT2_LD 36          // row stride of the staged K / V tiles (16-byte aligned rows, conflict-free LDS.128 with lane = key)

__device__ __forceinline__ void xt2_stage_tile(const float* __restrict__ src, int R, int row0, int h, float* __restrict__ dst) {
    // rows row0 .. row0 + 127 of src [R,256], channels 32 h .. 32 h + 31 -> dst [128][36]
    for (int i = threadIdx.x; i < XT2_KT * 8; i += 256) {
        const int r = i >> 3, c4 = i & 7;
        float* d = dst + r * XT2_LD + c4 * 4;
        if (row0 + r < R) {
            const unsigned sa = (unsigned)__cvta_generic_to_shared(d);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src + (long long)(row0 + r) * TC_ + h * THD + c4 * 4) : "memory");
        } else {
            *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

// grid (ceil(NT / XT2_QB), 8 heads), 256 threads.  P keeps the RAW logits of a query's keys; stats [NT, 8, 2] = (max,
// 1 / sum) of every (query, head): the backward kernels form p = exp(s - max) / sum on the fly (a normalising sweep over
// the 25 k keys of a denoising query would be one long dependent chain per warp).
__global__ void __launch_bounds__(256) xt2_fwd_kernel(const float* __restrict__ cq, const float* __restrict__ Kp, const float* __restrict__ Vp,
                                                      const uint32_t* __restrict__ keymask, int mask_words,
                                                      int NT, int R, float* __restrict__ P, float* __restrict__ stats, float* __restrict__ ctx) {
    pdl_wait();
    pdl_trigger();
    __shared__ __align__(16) float Ks[XT2_KT * XT2_LD];
    __shared__ __align__(16) float Vs[XT2_KT * XT2_LD];
    __shared__ __align__(16) float psm[8][32];
    const int h = blockIdx.y, q0 = blockIdx.x * XT2_QB, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float scale = 0.17677669529663687f;
    constexpr int QPW = XT2_QB / 8;
    float m_i[QPW], l_i[QPW], o_i[QPW];
#pragma unroll
    for (int u = 0; u < QPW; ++u) { m_i[u] = -INFINITY; l_i[u] = 0.f; o_i[u] = 0.f; }
    const int ntiles = (R + XT2_KT - 1) / XT2_KT;
    for (int t = 0; t < ntiles; ++t) {
        // does any query of the block have a key in this tile?
        int any = 0;
        if (threadIdx.x < XT2_QB * 4) {
            const int q = q0 + (threadIdx.x >> 2), w = t * 4 + (threadIdx.x & 3);
            if (q < NT && w < mask_words) any = keymask[(long long)q * mask_words + w] != 0u;
        }
        if (!__syncthreads_or(any)) continue;
        xt2_stage_tile(Kp, R, t * XT2_KT, h, Ks);
        xt2_stage_tile(Vp, R, t * XT2_KT, h, Vs);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
#pragma unroll
        for (int u = 0; u < QPW; ++u) {
            const int i = q0 + warp * QPW + u;
            if (i >= NT) break;
            const uint32_t* km = keymask + (long long)i * mask_words + t * 4;
            uint32_t wd = (lane < 4 && t * 4 + lane < mask_words) ? km[lane] : 0u;
            if (__ballot_sync(0xffffffffu, wd != 0u) == 0u) continue;
            float q[THD];
            const float4* qr = reinterpret_cast<const float4*>(cq + (long long)i * TC_ + h * THD);
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
                const float4 v = __ldg(qr + c4);
                q[c4 * 4] = v.x * scale; q[c4 * 4 + 1] = v.y * scale; q[c4 * 4 + 2] = v.z * scale; q[c4 * 4 + 3] = v.w * scale;
            }
            float* Prow = P + ((long long)i * TH + h) * R + t * XT2_KT;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const uint32_t bits = __shfl_sync(0xffffffffu, wd, g);
                if (bits == 0u) continue;
                const bool act = (bits >> lane) & 1u;
                float sv = -INFINITY;
                if (act) {
                    const float* kr = Ks + (g * 32 + lane) * XT2_LD;
                    float a = 0.f;
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        const float4 kk = *reinterpret_cast<const float4*>(kr + c4 * 4);
                        a = fmaf(q[c4 * 4], kk.x, a); a = fmaf(q[c4 * 4 + 1], kk.y, a); a = fmaf(q[c4 * 4 + 2], kk.z, a); a = fmaf(q[c4 * 4 + 3], kk.w, a);
                    }
                    sv = a;
                    Prow[g * 32 + lane] = a;               // raw logit (see stats)
                }
                const float m_new = fmaxf(m_i[u], warp_max(sv));
                const float alpha = m_i[u] == -INFINITY ? 0.f : expf(m_i[u] - m_new);
                const float pv = act ? expf(sv - m_new) : 0.f;
                l_i[u] = l_i[u] * alpha + warp_sum(pv);
                m_i[u] = m_new;
                __syncwarp();
                psm[warp][lane] = pv;
                __syncwarp();
                // all 32 keys of the group, unrolled (an inactive key carries p = 0): independent LDS, two FMA chains
                float acc0 = o_i[u] * alpha, acc1 = 0.f;
                const float* vcol = Vs + g * 32 * XT2_LD + lane;
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) {
                    const float4 pp = *reinterpret_cast<const float4*>(&psm[warp][k4 * 4]);
                    acc0 = fmaf(pp.x, vcol[(k4 * 4 + 0) * XT2_LD], acc0); acc1 = fmaf(pp.y, vcol[(k4 * 4 + 1) * XT2_LD], acc1);
                    acc0 = fmaf(pp.z, vcol[(k4 * 4 + 2) * XT2_LD], acc0); acc1 = fmaf(pp.w, vcol[(k4 * 4 + 3) * XT2_LD], acc1);
                }
                o_i[u] = acc0 + acc1;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < QPW; ++u) {
        const int i = q0 + warp * QPW + u;
        if (i >= NT) break;
        const float inv = l_i[u] > 0.f ? 1.f / l_i[u] : 0.f;
        ctx[(long long)i * TC_ + h * THD + lane] = o_i[u] * inv;
        if (lane == 0) { stats[((long long)i * TH + h) * 2] = m_i[u]; stats[((long long)i * TH + h) * 2 + 1] = inv; }
    }
}

// query-side backward, same tiling.  D = sum_k P_k dP_k = dO . ctx (head slice): no pass over the keys needed for it.
__global__ void __launch_bounds__(256) xt2_bwd_dq_kernel(const float* __restrict__ Kp, const float* __restrict__ Vp, const float* __restrict__ P,
                                                         const float* __restrict__ stats, const float* __restrict__ dctx, const float* __restrict__ ctx,
                                                         const uint32_t* __restrict__ keymask, int mask_words, int NT, int R,
                                                         float* __restrict__ dS, float* __restrict__ dcq) {
    pdl_wait();
    pdl_trigger();
    __shared__ __align__(16) float Ks[XT2_KT * XT2_LD];
    __shared__ __align__(16) float Vs[XT2_KT * XT2_LD];
    __shared__ __align__(16) float psm[8][32];
    const int h = blockIdx.y, q0 = blockIdx.x * XT2_QB, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float scale = 0.17677669529663687f;
    constexpr int QPW = XT2_QB / 8;
    float D_i[QPW], dq_i[QPW];
#pragma unroll
    for (int u = 0; u < QPW; ++u) {
        const int i = q0 + warp * QPW + u;
        dq_i[u] = 0.f;
        float d = 0.f;
        if (i < NT) d = dctx[(long long)i * TC_ + h * THD + lane] * ctx[(long long)i * TC_ + h * THD + lane];
        D_i[u] = warp_sum(d);
    }
    const int ntiles = (R + XT2_KT - 1) / XT2_KT;
    for (int t = 0; t < ntiles; ++t) {
        int any = 0;
        if (threadIdx.x < XT2_QB * 4) {
            const int q = q0 + (threadIdx.x >> 2), w = t * 4 + (threadIdx.x & 3);
            if (q < NT && w < mask_words) any = keymask[(long long)q * mask_words + w] != 0u;
        }
        if (!__syncthreads_or(any)) continue;
        xt2_stage_tile(Kp, R, t * XT2_KT, h, Ks);
        xt2_stage_tile(Vp, R, t * XT2_KT, h, Vs);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
#pragma unroll
        for (int u = 0; u < QPW; ++u) {
            const int i = q0 + warp * QPW + u;
            if (i >= NT) break;
            const uint32_t* km = keymask + (long long)i * mask_words + t * 4;
            uint32_t wd = (lane < 4 && t * 4 + lane < mask_words) ? km[lane] : 0u;
            if (__ballot_sync(0xffffffffu, wd != 0u) == 0u) continue;
            float go[THD];
            const float4* gr = reinterpret_cast<const float4*>(dctx + (long long)i * TC_ + h * THD);
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
                const float4 v = __ldg(gr + c4);
                go[c4 * 4] = v.x; go[c4 * 4 + 1] = v.y; go[c4 * 4 + 2] = v.z; go[c4 * 4 + 3] = v.w;
            }
            const long long prow = ((long long)i * TH + h) * R + t * XT2_KT;
            const float smax = __ldg(stats + ((long long)i * TH + h) * 2), sinv = __ldg(stats + ((long long)i * TH + h) * 2 + 1);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const uint32_t bits = __shfl_sync(0xffffffffu, wd, g);
                if (bits == 0u) continue;
                const bool act = (bits >> lane) & 1u;
                float ds = 0.f;
                if (act) {
                    const float* vr = Vs + (g * 32 + lane) * XT2_LD;
                    float dp = 0.f;
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        const float4 vv = *reinterpret_cast<const float4*>(vr + c4 * 4);
                        dp = fmaf(go[c4 * 4], vv.x, dp); dp = fmaf(go[c4 * 4 + 1], vv.y, dp); dp = fmaf(go[c4 * 4 + 2], vv.z, dp); dp = fmaf(go[c4 * 4 + 3], vv.w, dp);
                    }
                    ds = expf(P[prow + g * 32 + lane] - smax) * sinv * (dp - D_i[u]);
                    dS[prow + g * 32 + lane] = ds;
                }
                __syncwarp();
                psm[warp][lane] = ds;
                __syncwarp();
                float acc0 = dq_i[u], acc1 = 0.f;
                const float* kcol = Ks + g * 32 * XT2_LD + lane;
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) {
                    const float4 pp = *reinterpret_cast<const float4*>(&psm[warp][k4 * 4]);
                    acc0 = fmaf(pp.x, kcol[(k4 * 4 + 0) * XT2_LD], acc0); acc1 = fmaf(pp.y, kcol[(k4 * 4 + 1) * XT2_LD], acc1);
                    acc0 = fmaf(pp.z, kcol[(k4 * 4 + 2) * XT2_LD], acc0); acc1 = fmaf(pp.w, kcol[(k4 * 4 + 3) * XT2_LD], acc1);
                }
                dq_i[u] = acc0 + acc1;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < QPW; ++u) {
        const int i = q0 + warp * QPW + u;
        if (i < NT) dcq[(long long)i * TC_ + h * THD + lane] = dq_i[u] * scale;
    }
}

// key-side backward: CTA = (128 keys, head), 128 threads = one key each with its 2 x 32 accumulators in registers; the
// queries are walked in ascending order in chunks of 32 whose q / dO head slices are staged in shared memory
__global__ void __launch_bounds__(128) xt2_bwd_dkv_kernel(const float* __restrict__ cq, const float* __restrict__ dctx, const float* __restrict__ P,
                                                          const float* __restrict__ stats,
                                                          const float* __restrict__ dS, const uint32_t* __restrict__ keymask, int mask_words,
                                                          int NT, int R, float* __restrict__ dKp, float* __restrict__ dVp, int ldo) {
    pdl_wait();
    pdl_trigger();
    __shared__ __align__(16) float Qs[32 * THD];
    __shared__ __align__(16) float Gs[32 * THD];
    __shared__ float Ss[32 * 2];
    const int t = blockIdx.x, h = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int key = t * XT2_KT + threadIdx.x, word = t * 4 + warp;
    const float scale = 0.17677669529663687f;
    float dk[THD], dv[THD];
#pragma unroll
    for (int c = 0; c < THD; ++c) { dk[c] = 0.f; dv[c] = 0.f; }
    for (int i0 = 0; i0 < NT; i0 += 32) {
        __syncthreads();
        for (int e = threadIdx.x; e < 32 * 8; e += 128) {
            const int r = e >> 3, c4 = e & 7;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
            if (i0 + r < NT) {
                a = __ldg(reinterpret_cast<const float4*>(cq + (long long)(i0 + r) * TC_ + h * THD) + c4);
                b = __ldg(reinterpret_cast<const float4*>(dctx + (long long)(i0 + r) * TC_ + h * THD) + c4);
            }
            *reinterpret_cast<float4*>(Qs + r * THD + c4 * 4) = a;
            *reinterpret_cast<float4*>(Gs + r * THD + c4 * 4) = b;
        }
        if (threadIdx.x < 64) Ss[threadIdx.x] = (i0 + (threadIdx.x >> 1) < NT) ? stats[((long long)(i0 + (threadIdx.x >> 1)) * TH + h) * 2 + (threadIdx.x & 1)] : 0.f;
        __syncthreads();
        // the mask word of (query i0 + lane, this warp's 32 keys): one coalesced-ish load per chunk, then shuffles
        uint32_t mine = 0u;
        if (i0 + lane < NT && word < mask_words) mine = __ldg(keymask + (long long)(i0 + lane) * mask_words + word);
        if (__ballot_sync(0xffffffffu, mine != 0u) == 0u) continue;
        const int nq = min(32, NT - i0);
        for (int r = 0; r < nq; ++r) {
            const uint32_t bits = __shfl_sync(0xffffffffu, mine, r);
            if (bits == 0u) continue;
            if (!((bits >> lane) & 1u)) continue;
            const long long o = ((long long)(i0 + r) * TH + h) * R + key;
            const float p = expf(__ldg(P + o) - Ss[r * 2]) * Ss[r * 2 + 1], ds = __ldg(dS + o);
            const float* qr = Qs + r * THD;
            const float* gr = Gs + r * THD;
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
                const float4 qq = *reinterpret_cast<const float4*>(qr + c4 * 4);
                const float4 gg = *reinterpret_cast<const float4*>(gr + c4 * 4);
                dk[c4 * 4] = fmaf(ds, qq.x, dk[c4 * 4]); dk[c4 * 4 + 1] = fmaf(ds, qq.y, dk[c4 * 4 + 1]);
                dk[c4 * 4 + 2] = fmaf(ds, qq.z, dk[c4 * 4 + 2]); dk[c4 * 4 + 3] = fmaf(ds, qq.w, dk[c4 * 4 + 3]);
                dv[c4 * 4] = fmaf(p, gg.x, dv[c4 * 4]); dv[c4 * 4 + 1] = fmaf(p, gg.y, dv[c4 * 4 + 1]);
                dv[c4 * 4 + 2] = fmaf(p, gg.z, dv[c4 * 4 + 2]); dv[c4 * 4 + 3] = fmaf(p, gg.w, dv[c4 * 4 + 3]);
            }
        }
    }
    if (key < R) {
        float4* ok = reinterpret_cast<float4*>(dKp + (long long)key * ldo + h * THD);
        float4* ov = reinterpret_cast<float4*>(dVp + (long long)key * ldo + h * THD);
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
            ok[c4] = make_float4(dk[c4 * 4] * scale, dk[c4 * 4 + 1] * scale, dk[c4 * 4 + 2] * scale, dk[c4 * 4 + 3] * scale);
            ov[c4] = make_float4(dv[c4 * 4], dv[c4 * 4 + 1], dv[c4 * 4 + 2], dv[c4 * 4 + 3]);
        }
    }
}

inline bool xt2_enabled() {
    static const bool on = []() { const char* e = getenv("MV2D_TRAIN_XT_TILED"); return !(e && e[0] == '0'); }();
    return on;
}
