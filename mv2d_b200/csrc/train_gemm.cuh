// Contractions of the training step: the strided fp32 FFMA GEMM, the tcgen05 3xTF32 route with re-laid-out operands, and the linear forward / backward-data / weight-gradient helpers built on them.
// Included by train.cu only (inside namespace mv2d { namespace { ... } }): one translation unit, several files.
#pragma once

// ------------------------------------------------------------------------------------------------ generic fp32 GEMM
// C[M,N] (op)= sum_k A(m,k) B(k,n) (+ bias[n]),  A(m,k) = A[m*sam + k*sak],  B(k,n) = B[k*sbk + n*sbn].
enum SgFlags { SG_RELU = 1, SG_ACC = 2, SG_ATOMIC = 4, SG_CLAMP5E3 = 8, SG_MASK_LT5E3 = 16 };
struct Sg {
    const float* A; const float* B; float* C; const float* bias; const float* mask; float* rowsum;
    long long sam, sak, sbk, sbn;
    int ldc, ldmask, M, N, K, klen, flags;
};

// TM x TM outputs per thread, 256 threads: TM = 4 -> 64 x 64 tile with BK = 32, TM = 8 -> 128 x 128 tile with BK = 16
// (eight elements of each operand per thread and k-step, prefetched into registers while the previous k-step is
// multiplied: the M ~ 300 problems of the decoder are short chains of k-steps on a few CTAs, so the loads in flight
// per step set their speed).  AK1 / BN1 say which stride of A / B is 1, i.e. which index runs along a warp when the
// tile is loaded (coalescing only; addressing always goes through the strides).
// rowsum (weight-gradient calls): the CTAs of the first column of tiles also add sum_k A(m,k) -- the bias gradient
// of the same layer -- so no separate column-sum launch is needed.
template <int TM, bool AK1, bool BN1>
__global__ void __launch_bounds__(256) sgemm_kernel(Sg g) {
    pdl_wait();
    pdl_trigger();
    constexpr int BM = 16 * TM, BK = 2048 / BM, LD = BM + 4, E = 8;
    __shared__ __align__(16) float As[BK][LD];
    __shared__ __align__(16) float Bs[BK][LD];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BM;
    const int kbeg = blockIdx.z * g.klen;
    const int kend = min(g.K, kbeg + g.klen);
    const bool do_rowsum = g.rowsum != nullptr && blockIdx.x == 0 && tx == 0;
    float acc[TM][TM], rs[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        rs[i] = 0.f;
#pragma unroll
        for (int j = 0; j < TM; ++j) acc[i][j] = 0.f;
    }
    float ra[E], rb[E];
    // element e of this thread inside a tile: (am, ak) for A, (bn, bk) for B
    auto a_m = [&](int e) { const int idx = tid + e * 256; return AK1 ? idx / BK : idx % BM; };
    auto a_k = [&](int e) { const int idx = tid + e * 256; return AK1 ? idx % BK : idx / BM; };
    auto b_n = [&](int e) { const int idx = tid + e * 256; return BN1 ? idx % BM : idx / BK; };
    auto b_k = [&](int e) { const int idx = tid + e * 256; return BN1 ? idx / BM : idx % BK; };
    auto fetch = [&](int k0) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int am = a_m(e), ak = a_k(e), bn = b_n(e), bk = b_k(e);
            ra[e] = (m0 + am < g.M && k0 + ak < kend) ? __ldg(g.A + (long long)(m0 + am) * g.sam + (long long)(k0 + ak) * g.sak) : 0.f;
            rb[e] = (n0 + bn < g.N && k0 + bk < kend) ? __ldg(g.B + (long long)(k0 + bk) * g.sbk + (long long)(n0 + bn) * g.sbn) : 0.f;
        }
    };
    if (kbeg < kend) fetch(kbeg);
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
        for (int e = 0; e < E; ++e) { As[a_k(e)][a_m(e)] = ra[e]; Bs[b_k(e)][b_n(e)] = rb[e]; }
        __syncthreads();
        if (k0 + BK < kend) fetch(k0 + BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float av[TM], bv[TM];
#pragma unroll
            for (int q = 0; q < TM / 4; ++q) {
                const float4 a = *reinterpret_cast<const float4*>(&As[k][q * 64 + ty * 4]);
                const float4 b = *reinterpret_cast<const float4*>(&Bs[k][q * 64 + tx * 4]);
                av[q * 4] = a.x; av[q * 4 + 1] = a.y; av[q * 4 + 2] = a.z; av[q * 4 + 3] = a.w;
                bv[q * 4] = b.x; bv[q * 4 + 1] = b.y; bv[q * 4 + 2] = b.z; bv[q * 4 + 3] = b.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TM; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            if (do_rowsum) {
#pragma unroll
                for (int i = 0; i < TM; ++i) rs[i] += av[i];
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + (i / 4) * 64 + ty * 4 + (i % 4);
        if (m >= g.M) continue;
        if (do_rowsum) atomicAdd(g.rowsum + m, rs[i]);
#pragma unroll
        for (int j = 0; j < TM; ++j) {
            const int n = n0 + (j / 4) * 64 + tx * 4 + (j % 4);
            if (n >= g.N) continue;
            float v = acc[i][j];
            if (g.bias && blockIdx.z == 0) v += __ldg(g.bias + n);
            float* c = g.C + (long long)m * g.ldc + n;
            if (g.flags & SG_ATOMIC) { atomicAdd(c, v); continue; }
            if (g.flags & SG_RELU) v = fmaxf(v, 0.f);
            if (g.flags & SG_CLAMP5E3) v = fminf(v, 5e3f);
            if (g.mask) {
                const float a = g.mask[(long long)m * g.ldmask + n];
                if (!(a > 0.f) || ((g.flags & SG_MASK_LT5E3) && !(a < 5e3f))) v = 0.f;
            }
            if (g.flags & SG_ACC) v += *c;
            *c = v;
        }
    }
}

// 128 x 128 tiles once the problem fills the GPU with them, 64 x 64 otherwise
inline int sg_tile(int M, int N) {
    static const bool big_ok = []() { const char* e = getenv("MV2D_TRAIN_SGEMM128"); return !(e && e[0] == '0'); }();
    return (big_ok && M >= 512 && N >= 128) ? 128 : 64;
}

template <int TM>
int launch_sgemm_t(const Sg& g, dim3 grid, cudaStream_t st) {
    const bool ak1 = g.sak == 1, bn1 = g.sbn == 1;
    if (ak1 && bn1) launch_k(sgemm_kernel<TM, true, true>, grid, dim3(256), 0, st, g);
    else if (ak1) launch_k(sgemm_kernel<TM, true, false>, grid, dim3(256), 0, st, g);
    else if (bn1) launch_k(sgemm_kernel<TM, false, true>, grid, dim3(256), 0, st, g);
    else launch_k(sgemm_kernel<TM, false, false>, grid, dim3(256), 0, st, g);
    MV2D_CHECK_LAUNCH("train sgemm");
    return 0;
}

int launch_sgemm(const Sg& g, int splits, cudaStream_t st) {
    if (g.M <= 0 || g.N <= 0 || g.K <= 0) return 0;
    const int t = sg_tile(g.M, g.N);
    dim3 grid(cdiv(g.N, t), cdiv(g.M, t), splits);
    return t == 128 ? launch_sgemm_t<8>(g, grid, st) : launch_sgemm_t<4>(g, grid, st);
}


// ------------------------------------------------------------------------------------------------ tensor-core route
// The GPU-filling contractions of the step (K/V projections over all RoI tokens, the 3x3 conv as an im2col GEMM, the
// position-encoding MLPs: M = 14 700 .. 16 896 rows) run on the tcgen05 kernel of gemm_tc.cu as error-compensated
// 3xTF32 (fp32-grade, operands split inside the kernel).  That kernel computes C = A W^T with both operands
// K-contiguous, so the backward forms get their operands re-laid-out first:
//   dX = dY W        -> W^T is materialised (weights are small), the ReLU mask / accumulation is a second pass;
//   dW = dY^T X      -> dY^T and X^T are materialised with the row count zero-padded to a multiple of 32 (the GEMM's
//                       K), the reduction is split over CTAs (raw partial sums) and one kernel folds the partials into
//                       the flat gradient buffer and the bias gradient.
// MV2D_TRAIN_TC=0 keeps everything on the FFMA kernel below (the tests run both).
struct TcScratch {
    float *at, *bt, *wt, *part, *tmp;
    size_t at_cap, bt_cap, wt_cap, part_cap, tmp_cap;   // floats
};
int g_tc_mode = -1;      // -1 = not set yet: MV2D_TRAIN_TC from the environment (default on); mv2d_train_set_tensor_cores overrides
bool tc_enabled() {
    if (g_tc_mode < 0) { const char* e = getenv("MV2D_TRAIN_TC"); g_tc_mode = (e && e[0] == '0') ? 0 : 1; }
    return g_tc_mode == 1;
}
inline bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }
inline int round32(int x) { return (x + 31) / 32 * 32; }

int tc_gemm(const float* A, int lda, const float* W, int ldw, const float* bias, float* C, int ldc, int M, int N, int K, bool relu,
            int nsplit, long long split_stride, cudaStream_t st) {
    TcGemm t{};
    t.A = A; t.A_lo = nullptr; t.lda = lda; t.W = W; t.W_lo = nullptr; t.ldw = ldw; t.bias = bias; t.C = C; t.ldc = ldc;
    t.M = M; t.N = N; t.K = K; t.passes = 3; t.im2col = 0; t.flags = relu ? GEMM_RELU : 0; t.nsplit = nsplit; t.split_stride = split_stride;
    return launch_gemm_tc(t, st);
}
// shape rule of launch_gemm_tc: N tiles are 64 wide for M <= 512, 128 wide otherwise
inline bool tc_shape_ok(int M, int N, int K) { return K % 32 == 0 && N % (M <= 512 ? 64 : 128) == 0; }

// out[c * ldo + r] = in[r * ld + c] for r < Rp (zero for R <= r < Rp), c < C
__global__ void __launch_bounds__(256) transpose_pad_kernel(const float* __restrict__ in, int ld, int R, int C, float* __restrict__ out, int Rp,
                                                            int ldo) {
    pdl_wait();
    pdl_trigger();
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + ty + i * 8, c = c0 + tx;
        tile[ty + i * 8][tx] = (r < R && c < C) ? in[(long long)r * ld + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + i * 8, r = r0 + tx;
        if (c < C && r < Rp) out[(long long)c * ldo + r] = tile[tx][ty + i * 8];
    }
}
int transpose_pad(const float* in, int ld, int R, int C, float* out, int Rp, cudaStream_t st, int ldo = 0) {
    launch_k(transpose_pad_kernel, dim3(cdiv(Rp, 32), cdiv(C, 32)), dim3(256), 0, st, in, ld, R, C, out, Rp, ldo > 0 ? ldo : Rp);
    MV2D_CHECK_LAUNCH("train transpose");
    return 0;
}

// dW[n,k] += sum_z part[z][...]; part is [rows, cols] = [Nout, K], or [K, Nout] when `swapped`
// rows_per_blk / blk_stride: output rows n are grouped in blocks of rows_per_blk that sit blk_stride floats apart in dW (the
// same tensor of consecutive decoder layers in the flat gradient buffer); 0 = one contiguous matrix
__global__ void __launch_bounds__(256) wgrad_fold_kernel(const float* __restrict__ part, int nsplit, long long stride, int Nout, int K,
                                                         int swapped, float* __restrict__ dW, int ldw, int rows_per_blk, long long blk_stride) {
    pdl_wait();
    pdl_trigger();
    const long long total = (long long)Nout * K;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        // walk the partial in ITS row-major order (coalesced reads), scatter into dW
        int n, k;
        if (swapped) { k = (int)(i / Nout); n = (int)(i % Nout); } else { n = (int)(i / K); k = (int)(i % K); }
        float a = 0.f;
        for (int z = 0; z < nsplit; ++z) a += part[z * stride + i];
        if (rows_per_blk > 0) dW[(n / rows_per_blk) * blk_stride + (long long)(n % rows_per_blk) * ldw + k] += a;
        else dW[(long long)n * ldw + k] += a;
    }
}
// db[n] += sum_r yt[n][r]  (rows of the transposed, zero-padded output gradient); one CTA per row
__global__ void __launch_bounds__(256) rowsum_kernel(const float* __restrict__ yt, int Rp, int Nout, float* __restrict__ db, int rows_per_blk,
                                                     long long blk_stride) {
    pdl_wait();
    pdl_trigger();
    __shared__ float red[8];
    const int n = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float a = 0.f;
    for (int r = threadIdx.x; r < Rp; r += 256) a += yt[(long long)n * Rp + r];
    a = warp_sum(a);
    if (lane == 0) red[warp] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i];
        if (rows_per_blk > 0) db[(n / rows_per_blk) * blk_stride + n % rows_per_blk] += t;
        else db[n] += t;
    }
}
// dX = (accumulate ? dX : 0) + src . [mask > 0 (and < 5e3)]   (rows of K floats; ld per operand)
__global__ void __launch_bounds__(256) dgrad_finish_kernel(const float* __restrict__ src, int lds, const float* __restrict__ mask, int ldmask,
                                                           int lt5e3, int accumulate, float* __restrict__ dX, int ldx, int M, int K) {
    pdl_wait();
    pdl_trigger();
    const long long total = (long long)M * K;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const long long m = i / K;
        const int k = (int)(i % K);
        float v = src[m * lds + k];
        if (mask) {
            const float a = mask[m * ldmask + k];
            if (!(a > 0.f) || (lt5e3 && !(a < 5e3f))) v = 0.f;
        }
        float* d = dX + m * ldx + k;
        *d = accumulate ? *d + v : v;
    }
}
inline int ew_grid_n(long long n) {
    const long long want = (n + 255) / 256;
    return (int)(want < 148 * 16 ? (want > 0 ? want : 1) : 148 * 16);
}

// the scratch of the current call (set by the run_* entry points; the library is single-threaded per call)
thread_local TcScratch g_tc{};

// Y[M,Nout] = act(X[M,K] W[Nout,K]^T + b)
int linear_fwd(const float* X, int ldx, const float* W, int ldw, const float* b, float* Y, int ldy, int M, int Nout, int K,
               bool relu, cudaStream_t st, int extra_flags = 0) {
    const bool lds_ok = (ldx & 3) == 0 && (ldw & 3) == 0 && (ldy & 3) == 0 && al16(X) && al16(W) && al16(Y);
    if (tc_enabled() && lds_ok && extra_flags == 0 && M >= 1024 && tc_shape_ok(M, Nout, K))
        return tc_gemm(X, ldx, W, ldw, b, Y, ldy, M, Nout, K, relu, 1, 0, st);
    if (tc_enabled() && lds_ok && M <= 512 && K % 32 == 0 && (Nout & 3) == 0 && (extra_flags & ~SG_CLAMP5E3) == 0) {
        // the inference path's small-M kernel (in-CTA split-K): fp32 FFMA, same arithmetic class as the kernel below
        GemmArgs a{};
        a.A = X; a.lda = ldx; a.W = W; a.ldw = ldw; a.C = Y; a.ldc = ldy; a.bias = b; a.M = M; a.N = Nout; a.K = K;
        a.batch = 1; a.nsplit = 1; a.flags = (relu ? GEMM_RELU : 0) | ((extra_flags & SG_CLAMP5E3) ? GEMM_CLAMP5E3 : 0);
        return launch_gemm_small(a, nullptr, 0, st);
    }
    Sg g{};
    g.A = X; g.sam = ldx; g.sak = 1; g.B = W; g.sbk = 1; g.sbn = ldw; g.C = Y; g.ldc = ldy; g.bias = b;
    g.M = M; g.N = Nout; g.K = K; g.klen = K; g.flags = (relu ? SG_RELU : 0) | extra_flags;
    return launch_sgemm(g, 1, st);
}
// dX[M,K] (+)= (dY[M,Nout] W[Nout,K]) . [mask > 0]
int linear_dgrad(const float* dY, int ldy, const float* W, int ldw, float* dX, int ldx, int M, int Nout, int K,
                 const float* mask, int ldmask, bool accumulate, cudaStream_t st, int extra_flags = 0) {
    const TcScratch& sc = g_tc;
    if (tc_enabled() && sc.wt && M >= 1024 && tc_shape_ok(M, K, Nout) && (ldy & 3) == 0 && (ldx & 3) == 0 && al16(dY) && al16(dX) &&
        (size_t)K * Nout <= sc.wt_cap && (!accumulate || (size_t)M * K <= sc.tmp_cap)) {
        TRY(transpose_pad(W, ldw, Nout, K, sc.wt, Nout, st));                       // W^T [K, Nout]
        float* target = accumulate ? sc.tmp : dX;
        const int ldt = accumulate ? K : ldx;
        TRY(tc_gemm(dY, ldy, sc.wt, Nout, nullptr, target, ldt, M, K, Nout, false, 1, 0, st));
        if (accumulate || mask) {
            launch_k(dgrad_finish_kernel, dim3(ew_grid_n((long long)M * K)), dim3(256), 0, st, (const float*)target, ldt, mask, ldmask,
                     (extra_flags & SG_MASK_LT5E3) ? 1 : 0, accumulate ? 1 : 0, dX, ldx, M, K);
            MV2D_CHECK_LAUNCH("train dgrad_finish");
        }
        return 0;
    }
    if (tc_enabled() && sc.wt && M <= 512 && Nout % 32 == 0 && (K & 3) == 0 && (ldy & 3) == 0 && (ldx & 3) == 0 && al16(dY) && al16(dX) &&
        (size_t)K * Nout <= sc.wt_cap && (!accumulate || (size_t)M * K <= sc.tmp_cap)) {
        // M ~ 300 rows: W^T once, then the inference path's small-M kernel (in-CTA split-K) -- a chain of k-steps on
        // the 20 CTAs the strided FFMA kernel would get for these shapes is latency bound
        TRY(transpose_pad(W, ldw, Nout, K, sc.wt, Nout, st));
        float* target = accumulate ? sc.tmp : dX;
        const int ldt = accumulate ? K : ldx;
        GemmArgs a{};
        a.A = dY; a.lda = ldy; a.W = sc.wt; a.ldw = Nout; a.C = target; a.ldc = ldt; a.M = M; a.N = K; a.K = Nout; a.batch = 1; a.nsplit = 1;
        TRY(launch_gemm_small(a, nullptr, 0, st));
        if (accumulate || mask) {
            launch_k(dgrad_finish_kernel, dim3(ew_grid_n((long long)M * K)), dim3(256), 0, st, (const float*)target, ldt, mask, ldmask,
                     (extra_flags & SG_MASK_LT5E3) ? 1 : 0, accumulate ? 1 : 0, dX, ldx, M, K);
            MV2D_CHECK_LAUNCH("train dgrad_finish");
        }
        return 0;
    }
    Sg g{};
    g.A = dY; g.sam = ldy; g.sak = 1; g.B = W; g.sbk = ldw; g.sbn = 1; g.C = dX; g.ldc = ldx; g.mask = mask; g.ldmask = ldmask;
    g.M = M; g.N = K; g.K = Nout; g.klen = Nout; g.flags = (accumulate ? SG_ACC : 0) | extra_flags;
    return launch_sgemm(g, 1, st);
}
// dW[Nout,K] += dY[M,Nout]^T X[M,K]   (split over the M rows, atomic accumulation);  db[Nout] += sum_rows dY (nullable)
int linear_wgrad(const float* dY, int ldy, const float* X, int ldx, float* dW, int ldw, int M, int Nout, int K, cudaStream_t st,
                 float* db = nullptr) {
    const TcScratch& sc = g_tc;
    if (tc_enabled() && sc.at && M >= 1024) {
        const int Mp = round32(M);
        const bool direct = tc_shape_ok(Nout, K, Mp), swapped = !direct && tc_shape_ok(K, Nout, Mp);
        const int gm = direct ? Nout : K, gn = direct ? K : Nout;           // the GEMM's M and N
        const int tiles = cdiv(gm, 128) * (gn / (gm <= 512 ? 64 : 128));
        const int nkb = Mp / 32;
        int nsplit = 1;
        for (int d = 1; d <= 48 && d <= nkb; ++d)
            if (nkb % d == 0 && nkb / d >= 4) { nsplit = d; if (tiles * d >= 148) break; }
        if ((direct || swapped) && (size_t)Nout * Mp <= sc.at_cap && (size_t)K * Mp <= sc.bt_cap &&
            (size_t)nsplit * Nout * K <= sc.part_cap) {
            TRY(transpose_pad(dY, ldy, M, Nout, sc.at, Mp, st));                     // dY^T [Nout, Mp]
            TRY(transpose_pad(X, ldx, M, K, sc.bt, Mp, st));                         // X^T  [K, Mp]
            const float* ga = direct ? sc.at : sc.bt;
            const float* gw = direct ? sc.bt : sc.at;
            TRY(tc_gemm(ga, Mp, gw, Mp, nullptr, sc.part, gn, gm, gn, Mp, false, nsplit, (long long)Nout * K, st));
            launch_k(wgrad_fold_kernel, dim3(ew_grid_n((long long)Nout * K)), dim3(256), 0, st, (const float*)sc.part, nsplit,
                     (long long)Nout * K, Nout, K, swapped ? 1 : 0, dW, ldw, 0, 0LL);
            MV2D_CHECK_LAUNCH("train wgrad_fold");
            if (db) {
                launch_k(rowsum_kernel, dim3(Nout), dim3(256), 0, st, (const float*)sc.at, Mp, Nout, db, 0, 0LL);
                MV2D_CHECK_LAUNCH("train rowsum");
            }
            return 0;
        }
    }
    Sg g{};
    g.A = dY; g.sam = 1; g.sak = ldy; g.B = X; g.sbk = ldx; g.sbn = 1; g.C = dW; g.ldc = ldw; g.rowsum = db;
    g.M = Nout; g.N = K; g.K = M; g.flags = SG_ATOMIC;
    const int t = sg_tile(Nout, K);
    const int tiles = cdiv(Nout, t) * cdiv(K, t);
    int splits = cdiv(296, tiles);
    splits = std::max(1, std::min(splits, cdiv(M, 64)));
    g.klen = cdiv(cdiv(M, splits), 16) * 16;
    splits = cdiv(M, g.klen);
    return launch_sgemm(g, splits, st);
}

// out = a + b (b nullable); in-place allowed
__global__ void __launch_bounds__(256) add_kernel(float* out, const float* a, const float* b, long long n) {
    pdl_wait();
    pdl_trigger();
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256)
        out[i] = a[i] + (b ? b[i] : 0.f);
}
int add(float* out, const float* a, const float* b, long long n, cudaStream_t st) {
    if (n <= 0) return 0;
    const long long want = (n + 255) / 256;
    const int grid = (int)(want < 148 * 8 ? want : 148 * 8);
    launch_k(add_kernel, dim3(grid), dim3(256), 0, st, out, a, b, n);
    MV2D_CHECK_LAUNCH("train add");
    return 0;
}
