// Front-end kernels of the training step: RoIAlign forward / backward, per-RoI camera parameters, im2col / col2im, average pool, center2lidar, SE gate.
// Included by train.cu only (inside namespace mv2d { namespace { ... } }): one translation unit, several files.
#pragma once

// ================================================================================================ front end
// Training forward / backward of rows a1-a8: position encoding (pe.py:137-169), RoIAlign of feat and pe
// (mv2d_s_head.py:133-138), query generator (query_generator.py:343-405) and the reference-point normalisation
// (mv2d_s_head.py:147-152), all in fp32 FFMA (the inference path runs the same MLPs as single-pass TF32).

// ---- RoIAlign (mmcv: avg, aligned=True, adaptive sampling grid), channels-last maps.  The bin geometry is evaluated
// with explicitly rounded fp32 operations (no FMA contraction) so that forward, backward and csrc/roi.cu make the
// same floor / in-range decisions.
struct RoiBin { int v, gh, gw; float x1, y1, bw, bh, count; };
__device__ __forceinline__ RoiBin roi_bin(const float* __restrict__ r, float spatial_scale) {
    RoiBin b;
    b.v = (int)r[0];
    b.x1 = __fadd_rn(__fmul_rn(r[1], spatial_scale), -0.5f);
    b.y1 = __fadd_rn(__fmul_rn(r[2], spatial_scale), -0.5f);
    const float x2 = __fadd_rn(__fmul_rn(r[3], spatial_scale), -0.5f), y2 = __fadd_rn(__fmul_rn(r[4], spatial_scale), -0.5f);
    const float rw = __fsub_rn(x2, b.x1), rh = __fsub_rn(y2, b.y1);
    b.bw = __fdiv_rn(rw, (float)MV2D_ROI);
    b.bh = __fdiv_rn(rh, (float)MV2D_ROI);
    b.gh = (int)ceilf(__fdiv_rn(rh, (float)MV2D_ROI));
    b.gw = (int)ceilf(__fdiv_rn(rw, (float)MV2D_ROI));
    b.count = (float)max(b.gh * b.gw, 1);
    return b;
}
struct RoiTap { int o1, o2, o3, o4; float w1, w2, w3, w4; bool ok; };
// sample (iy, ix) of bin (ph, pw): the four corner offsets (in pixels) and bilinear weights
__device__ __forceinline__ RoiTap roi_tap(const RoiBin& b, int ph, int pw, int iy, int ix, int h, int w) {
    RoiTap t;
    const float y = __fadd_rn(__fadd_rn(b.y1, __fmul_rn((float)ph, b.bh)), __fdiv_rn(__fmul_rn(__fadd_rn((float)iy, 0.5f), b.bh), (float)b.gh));
    const float x = __fadd_rn(__fadd_rn(b.x1, __fmul_rn((float)pw, b.bw)), __fdiv_rn(__fmul_rn(__fadd_rn((float)ix, 0.5f), b.bw), (float)b.gw));
    t.ok = !(y < -1.f || y > (float)h || x < -1.f || x > (float)w);
    float yy = fmaxf(y, 0.f), xx = fmaxf(x, 0.f);
    int yl = (int)yy, xl = (int)xx, yh, xh;
    if (yl >= h - 1) { yh = yl = h - 1; yy = (float)yl; } else yh = yl + 1;
    if (xl >= w - 1) { xh = xl = w - 1; xx = (float)xl; } else xh = xl + 1;
    const float ly = __fsub_rn(yy, (float)yl), lx = __fsub_rn(xx, (float)xl), hy = __fsub_rn(1.f, ly), hx = __fsub_rn(1.f, lx);
    t.w1 = __fmul_rn(hy, hx); t.w2 = __fmul_rn(hy, lx); t.w3 = __fmul_rn(ly, hx); t.w4 = __fmul_rn(ly, lx);
    t.o1 = yl * w + xl; t.o2 = yl * w + xh; t.o3 = yh * w + xl; t.o4 = yh * w + xh;
    return t;
}

// grid (49, N), 64 threads (one float4 of the 256 channels each): tok[n, bin] = pooled map (+ addend[n, bin])
__global__ void __launch_bounds__(64) roi_align_fwd_kernel(const float* __restrict__ rois, const float* __restrict__ map, int h, int w,
                                                           float spatial_scale, const float* __restrict__ addend, float* __restrict__ tok) {
    pdl_wait();
    pdl_trigger();
    const int n = blockIdx.y, bin = blockIdx.x, ph = bin / MV2D_ROI, pw = bin % MV2D_ROI;
    const RoiBin b = roi_bin(rois + n * 5, spatial_scale);
    const float4* m4 = reinterpret_cast<const float4*>(map) + (long long)b.v * h * w * 64 + threadIdx.x;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int iy = 0; iy < b.gh; ++iy)
        for (int ix = 0; ix < b.gw; ++ix) {
            const RoiTap t = roi_tap(b, ph, pw, iy, ix, h, w);
            if (!t.ok) continue;
            const float4 c1 = __ldg(m4 + (long long)t.o1 * 64), c2 = __ldg(m4 + (long long)t.o2 * 64),
                         c3 = __ldg(m4 + (long long)t.o3 * 64), c4 = __ldg(m4 + (long long)t.o4 * 64);
            a.x += t.w1 * c1.x + t.w2 * c2.x + t.w3 * c3.x + t.w4 * c4.x;
            a.y += t.w1 * c1.y + t.w2 * c2.y + t.w3 * c3.y + t.w4 * c4.y;
            a.z += t.w1 * c1.z + t.w2 * c2.z + t.w3 * c3.z + t.w4 * c4.z;
            a.w += t.w1 * c1.w + t.w2 * c2.w + t.w3 * c3.w + t.w4 * c4.w;
        }
    a.x /= b.count; a.y /= b.count; a.z /= b.count; a.w /= b.count;
    const long long o = ((long long)n * MV2D_TOK + bin) * 64 + threadIdx.x;
    if (addend) {
        const float4 e = reinterpret_cast<const float4*>(addend)[o];
        a.x += e.x; a.y += e.y; a.z += e.z; a.w += e.w;
    }
    reinterpret_cast<float4*>(tok)[o] = a;
}

// dmap[v, corner] += w * dtok[n, bin] / count   (atomic: RoIs and bins overlap on the map)
__global__ void __launch_bounds__(64) roi_align_bwd_kernel(const float* __restrict__ rois, const float* __restrict__ dtok, int h, int w,
                                                           float spatial_scale, float* __restrict__ dmap) {
    pdl_wait();
    pdl_trigger();
    const int n = blockIdx.y, bin = blockIdx.x, ph = bin / MV2D_ROI, pw = bin % MV2D_ROI;
    const RoiBin b = roi_bin(rois + n * 5, spatial_scale);
    float4 g = reinterpret_cast<const float4*>(dtok)[((long long)n * MV2D_TOK + bin) * 64 + threadIdx.x];
    g.x /= b.count; g.y /= b.count; g.z /= b.count; g.w /= b.count;
    float* base = dmap + (long long)b.v * h * w * MV2D_C + threadIdx.x * 4;
    for (int iy = 0; iy < b.gh; ++iy)
        for (int ix = 0; ix < b.gw; ++ix) {
            const RoiTap t = roi_tap(b, ph, pw, iy, ix, h, w);
            if (!t.ok) continue;
            const int off[4] = {t.o1, t.o2, t.o3, t.o4};
            const float wt[4] = {t.w1, t.w2, t.w3, t.w4};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float* d = base + (long long)off[k] * MV2D_C;
                atomicAdd(d + 0, wt[k] * g.x); atomicAdd(d + 1, wt[k] * g.y);
                atomicAdd(d + 2, wt[k] * g.z); atomicAdd(d + 3, wt[k] * g.w);
            }
        }
}

// per RoI: the intrinsics feature of get_roi_feat (mv2d_head.py:95-101: flatten(K') * scale, zero when the box is
// narrower than 4 px, clamped with the concatenation) into cat[:, 1024:1040], and float(inv(K' E^T)) of center2lidar
__global__ void front_params_kernel(const float* __restrict__ rois, const double* __restrict__ k_roi, const double* __restrict__ extrinsics,
                                    int N, float feat_scale, float* __restrict__ cat, float* __restrict__ m_roi) {
    pdl_wait();
    pdl_trigger();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float* r = rois + n * 5;
    const int v = (int)r[0];
    double K[16], E[16], L[16], Li[16];
    for (int i = 0; i < 16; ++i) { K[i] = k_roi[n * 16 + i]; E[i] = extrinsics[v * 16 + i]; }
    const bool invalid = (__fsub_rn(r[3], r[1]) < 4.f) || (__fsub_rn(r[4], r[2]) < 4.f);
    for (int i = 0; i < 16; ++i) {
        const float f = invalid ? 0.f : __fmul_rn((float)K[i], feat_scale);
        cat[(long long)n * 1040 + 1024 + i] = fminf(fmaxf(f, -5e3f), 5e3f);
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s += K[i * 4 + k] * E[j * 4 + k];
            L[i * 4 + j] = s;
        }
    inv4x4(L, Li);
    for (int i = 0; i < 16; ++i) m_roi[n * 16 + i] = (float)Li[i];
}

// [N,7,7,256] tokens -> [N*49, 9*256] patches of the 3x3 / padding 1 convolution, K ordered (ky, kx, c)
__global__ void __launch_bounds__(256) im2col_kernel(const float* __restrict__ tok, float* __restrict__ col, int N) {
    pdl_wait();
    pdl_trigger();
    const long long total = (long long)N * MV2D_TOK * 9 * 64;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int c4 = (int)(i % 64);
        const int tap = (int)((i / 64) % 9);
        const long long row = i / (64 * 9);
        const int t = (int)(row % MV2D_TOK), y = t / MV2D_ROI + tap / 3 - 1, x = t % MV2D_ROI + tap % 3 - 1;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (y >= 0 && y < MV2D_ROI && x >= 0 && x < MV2D_ROI)
            v = reinterpret_cast<const float4*>(tok)[((row / MV2D_TOK) * MV2D_TOK + y * MV2D_ROI + x) * 64 + c4];
        reinterpret_cast<float4*>(col)[(row * 9 + tap) * 64 + c4] = v;
    }
}
// dtok[n,y,x,:] = sum over taps of dcol at the output cell that read (y,x) through that tap, + a1 + a2 (nullable)
__global__ void __launch_bounds__(256) col2im_kernel(const float* __restrict__ dcol, const float* __restrict__ a1,
                                                     const float* __restrict__ a2, float* __restrict__ dtok, int N) {
    pdl_wait();
    pdl_trigger();
    const long long total = (long long)N * MV2D_TOK * 64;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int c4 = (int)(i % 64);
        const long long row = i / 64;
        const long long n = row / MV2D_TOK;
        const int t = (int)(row % MV2D_TOK), y = t / MV2D_ROI, x = t % MV2D_ROI;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a1) { const float4 e = reinterpret_cast<const float4*>(a1)[i]; s.x += e.x; s.y += e.y; s.z += e.z; s.w += e.w; }
        if (a2) { const float4 e = reinterpret_cast<const float4*>(a2)[i]; s.x += e.x; s.y += e.y; s.z += e.z; s.w += e.w; }
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int yo = y - (tap / 3 - 1), xo = x - (tap % 3 - 1);
            if (yo < 0 || yo >= MV2D_ROI || xo < 0 || xo >= MV2D_ROI) continue;
            const float4 e = reinterpret_cast<const float4*>(dcol)[((n * MV2D_TOK + yo * MV2D_ROI + xo) * 9 + tap) * 64 + c4];
            s.x += e.x; s.y += e.y; s.z += e.z; s.w += e.w;
        }
        reinterpret_cast<float4*>(dtok)[i] = s;
    }
}

// AvgPool2d(7) over [N,49,256] and its backward through the ReLU that precedes it
__global__ void __launch_bounds__(256) pool49_fwd_kernel(const float* __restrict__ y, float* __restrict__ out, int N) {
    pdl_wait();
    pdl_trigger();
    const int gid = blockIdx.x * 256 + threadIdx.x;
    if (gid >= N * MV2D_C) return;
    const int n = gid / MV2D_C, c = gid % MV2D_C;
    float s = 0.f;
    for (int t = 0; t < MV2D_TOK; ++t) s += y[((long long)n * MV2D_TOK + t) * MV2D_C + c];
    out[gid] = s / 49.0f;
}
__global__ void __launch_bounds__(256) pool49_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dpool,
                                                         float* __restrict__ dy, int N) {
    pdl_wait();
    pdl_trigger();
    const long long total = (long long)N * MV2D_TOK * MV2D_C;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const long long n = i / (MV2D_TOK * MV2D_C);
        const int c = (int)(i % MV2D_C);
        dy[i] = y[i] > 0.f ? dpool[n * MV2D_C + c] / 49.0f : 0.f;
    }
}

// center2lidar + normalisation (query_generator.py:333-341, mv2d_s_head.py:147-152): c = (u, v, d) -> ref
__global__ void __launch_bounds__(128) center_fwd_kernel(const float* __restrict__ c, const float* __restrict__ m_roi, Range6 pc,
                                                         float* __restrict__ ref, int N) {
    pdl_wait();
    pdl_trigger();
    const int n = blockIdx.x * 128 + threadIdx.x;
    if (n >= N) return;
    const float u = c[n * 3], v = c[n * 3 + 1], d = c[n * 3 + 2];
    const float hom[4] = {u * d, v * d, d, 1.f};
    const float* m = m_roi + n * 16;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float xyz = m[i * 4] * hom[0] + m[i * 4 + 1] * hom[1] + m[i * 4 + 2] * hom[2] + m[i * 4 + 3] * hom[3];
        ref[n * 3 + i] = (xyz - pc.v[i]) / (pc.v[i + 3] - pc.v[i]);
    }
}
__global__ void __launch_bounds__(128) center_bwd_kernel(const float* __restrict__ c, const float* __restrict__ m_roi, Range6 pc,
                                                         const float* __restrict__ d_ref, float* __restrict__ dc, int N) {
    pdl_wait();
    pdl_trigger();
    const int n = blockIdx.x * 128 + threadIdx.x;
    if (n >= N) return;
    const float u = c[n * 3], v = c[n * 3 + 1], d = c[n * 3 + 2];
    const float* m = m_roi + n * 16;
    float dh[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float g = d_ref[n * 3 + i] / (pc.v[i + 3] - pc.v[i]);
        dh[0] += m[i * 4] * g; dh[1] += m[i * 4 + 1] * g; dh[2] += m[i * 4 + 2] * g;
    }
    dc[n * 3] = dh[0] * d;
    dc[n * 3 + 1] = dh[1] * d;
    dc[n * 3 + 2] = dh[0] * u + dh[1] * v + dh[2];
}

// SE gate + combine of PE.forward (pe.py:158-166): pe = x * sigmoid(g2) + sb; gate overwrites g2
__global__ void __launch_bounds__(256) pe_gate_fwd_kernel(const float* __restrict__ x, float* __restrict__ g2, const float* __restrict__ sb,
                                                          float* __restrict__ pe, long long n) {
    pdl_wait();
    pdl_trigger();
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float s = sigmoid_f(g2[i]);
        g2[i] = s;
        pe[i] = x[i] * s + sb[i];
    }
}
// dpe -> dx = dpe * gate, dg2 = dpe * x * gate (1 - gate)   (d sb = dpe itself)
__global__ void __launch_bounds__(256) pe_gate_bwd_kernel(const float* __restrict__ dpe, const float* __restrict__ x,
                                                          const float* __restrict__ gate, float* __restrict__ dx, float* __restrict__ dg2,
                                                          long long n) {
    pdl_wait();
    pdl_trigger();
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float s = gate[i], g = dpe[i];
        dx[i] = g * s;
        dg2[i] = g * x[i] * s * (1.f - s);
    }
}
