// Tensor-times-matrix kernels of the HOOI loop (CoffeeModel.build).
//
//   pb200_ttm        res[i0,:,:] += val * U[i1,:] (x) W[i2,:]   for a tensor grouped by i0
//                    (replaces dttm_seq / dttm_par, polara/lib/sparse.py:203-234, as called by
//                     ttm3d_*, polara/lib/tensor.py:7-34): an SpMM of the mode-unfolded tensor
//                    against a Khatri-Rao panel that is formed on the fly per nnz.
//   pb200_ttm_reduce the same sum when the grouped mode has only a handful of huge segments
//                    (the feedback mode): a gathered cross-Gram  A[ia,:]^T diag(val) B[ib,:]
//                    accumulated in fp64 with a deterministic two-stage reduction.
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int CB = 2048;
constexpr int WARPS = 8;
constexpr int LONG_ROW = 4096;

__device__ __forceinline__ int64_t lower_bound_i64(const int64_t* a, int64_t n, int64_t key) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

template <int J>
__device__ __forceinline__ void ttm_accumulate(float (&acc)[J], const int (&xo)[J], const int (&yo)[J], int64_t beg,
                                               int64_t end, int64_t step, const int32_t* __restrict__ i1,
                                               const int32_t* __restrict__ i2, const float* __restrict__ values,
                                               const float* __restrict__ U, int64_t ldu,
                                               const float* __restrict__ W, int64_t ldw, int lane) {
    for (int64_t p = beg; p < end; p += 32 * step) {
        int64_t q = p + lane;
        int32_t a = 0, b = 0;
        float v = 0.f;
        if (q < end) { a = __ldg(i1 + q); b = __ldg(i2 + q); v = __ldg(values + q); }
        int cnt = (int)min((int64_t)32, end - p);
        for (int t = 0; t < cnt; ++t) {
            int32_t at = __shfl_sync(0xffffffffu, a, t), bt = __shfl_sync(0xffffffffu, b, t);
            float vt = __shfl_sync(0xffffffffu, v, t);
            const float* u = U + (int64_t)at * ldu;
            const float* w = W + (int64_t)bt * ldw;
#pragma unroll
            for (int j = 0; j < J; ++j)
                if (xo[j] >= 0) acc[j] = fmaf(vt * __ldg(u + xo[j]), __ldg(w + yo[j]), acc[j]);
        }
    }
}

template <int J>
__global__ void __launch_bounds__(WARPS * 32)
ttm_kernel(int64_t n0, int64_t nnz, const int64_t* __restrict__ seg_ptr, const int32_t* __restrict__ i1,
           const int32_t* __restrict__ i2, const float* __restrict__ values, const float* __restrict__ U, int ru,
           int64_t ldu, const float* __restrict__ W, int rw, int64_t ldw, float* __restrict__ out, int64_t ldo,
           int64_t n_blocks) {
    __shared__ int64_t s_rows[2];
    __shared__ int s_next, s_nlong;
    __shared__ int64_t s_long[2];
    __shared__ float s_part[WARPS][32 * J];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t b = blockIdx.x;
    const int width = ru * rw;
    int xo[J], yo[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        int col = lane + 32 * j;
        xo[j] = col < width ? col / rw : -1;
        yo[j] = col < width ? col % rw : 0;
    }
    if (threadIdx.x == 0) {
        s_rows[0] = lower_bound_i64(seg_ptr, n0, b * (int64_t)CB);
        s_rows[1] = (b == n_blocks - 1) ? n0 : lower_bound_i64(seg_ptr, n0, (b + 1) * (int64_t)CB);
        s_next = 0; s_nlong = 0;
    }
    __syncthreads();
    const int64_t row_lo = s_rows[0], row_hi = s_rows[1];
    for (;;) {
        int idx = 0;
        if (lane == 0) idx = atomicAdd(&s_next, 1);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        int64_t row = row_lo + idx;
        if (row >= row_hi) break;
        int64_t beg = seg_ptr[row], end = seg_ptr[row + 1];
        if (end - beg > LONG_ROW) {
            if (lane == 0) { int s = atomicAdd(&s_nlong, 1); if (s < 2) s_long[s] = row; }
            continue;
        }
        float acc[J];
#pragma unroll
        for (int j = 0; j < J; ++j) acc[j] = 0.f;
        ttm_accumulate<J>(acc, xo, yo, beg, end, 1, i1, i2, values, U, ldu, W, ldw, lane);
#pragma unroll
        for (int j = 0; j < J; ++j) if (xo[j] >= 0) out[row * ldo + lane + 32 * j] = acc[j];
    }
    __syncthreads();
    const int nlong = min(s_nlong, 2);
    for (int li = 0; li < nlong; ++li) {
        int64_t row = s_long[li];
        int64_t beg = seg_ptr[row], end = seg_ptr[row + 1];
        float acc[J];
#pragma unroll
        for (int j = 0; j < J; ++j) acc[j] = 0.f;
        ttm_accumulate<J>(acc, xo, yo, beg + 32 * (int64_t)warp, end, WARPS, i1, i2, values, U, ldu, W, ldw, lane);
#pragma unroll
        for (int j = 0; j < J; ++j) s_part[warp][lane + 32 * j] = acc[j];
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int j = 0; j < J; ++j) {
                if (xo[j] < 0) continue;
                float s = 0.f;
                for (int w = 0; w < WARPS; ++w) s += s_part[w][lane + 32 * j];
                out[row * ldo + lane + 32 * j] = s;
            }
        }
        __syncthreads();
    }
}

// ---------------- nnz windows per warp (same scheme as spmm_window_kernel, spmm.cu) -----------------------------------
// Grouping by item gives rows of up to ~1e6 nnz (popular items): the row-owned kernel above leaves such a row to ONE
// block.  Here every warp owns a window of TW consecutive nnz and all ru*rw output columns; a row that straddles windows is
// written by the warp where it starts, later pieces go to a carry buffer and are added in window order (deterministic).
// Two nnz are in flight per step (their factor-row gathers overlap).
constexpr int TW = 512;        // nnz per warp window
constexpr int TWARPS = 8;

__device__ __forceinline__ int64_t lower_bound_ptr(const int64_t* __restrict__ a, int64_t n, int64_t key) {
    int64_t lo = 0, hi = n + 1;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

template <int J>
__global__ void __launch_bounds__(TWARPS * 32)
ttm_window_kernel(int64_t n0, int64_t nnz, const int64_t* __restrict__ seg_ptr, const int32_t* __restrict__ i1,
                  const int32_t* __restrict__ i2, const float* __restrict__ values, const float* __restrict__ U, int ru,
                  int64_t ldu, const float* __restrict__ W, int rw, int64_t ldw, float* __restrict__ out, int64_t ldo,
                  int64_t n_windows, float* __restrict__ carry /*[n_windows][32*J]*/, int64_t* __restrict__ carry_row) {
    const int lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * TWARPS + (threadIdx.x >> 5);
    if (b >= n_windows) return;
    const int width = ru * rw;
    int xo[J], yo[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int col = lane + 32 * j;
        xo[j] = col < width ? col / rw : -1;
        yo[j] = col < width ? col % rw : 0;
    }
    const int64_t w0 = b * (int64_t)TW;
    const int64_t w1 = min(nnz, w0 + TW);
    const int rel_w1 = (int)max(w1 - w0, (int64_t)0);
    const int n_groups = (rel_w1 + 31) / 32;
    const int64_t own_end = (b == n_windows - 1) ? nnz + 1 : w1;
    int64_t cur;
    bool piece_is_carry;
    {
        const int64_t lb = b == 0 ? 0 : lower_bound_ptr(seg_ptr, n0, w0);
        if (lb <= n0 && (b == 0 || __ldg(seg_ptr + lb) == w0)) { cur = lb; piece_is_carry = false; }
        else { cur = lb - 1; piece_is_carry = true; }
    }
    if (lane == 0) carry_row[b] = piece_is_carry ? cur : -1;
    int64_t pbase = cur;
    auto load_ptrs = [&](int64_t base) -> int {
        const int64_t r = min(base + lane, n0);
        const int64_t v = __ldg(seg_ptr + r) - w0;
        return (int)max((int64_t)-1, min(v, (int64_t)TW + 2));
    };
    int ptrs = load_ptrs(pbase);
    auto row_end_rel = [&]() -> int {
        if (cur + 1 - pbase >= 32) { pbase = cur; ptrs = load_ptrs(pbase); }
        return __shfl_sync(0xffffffffu, ptrs, (int)(cur + 1 - pbase));
    };
    float acc[J];
#pragma unroll
    for (int j = 0; j < J; ++j) acc[j] = 0.f;
    auto emit = [&]() {
#pragma unroll
        for (int j = 0; j < J; ++j) {
            if (xo[j] >= 0) {
                if (piece_is_carry) carry[b * (int64_t)(32 * J) + lane + 32 * j] = acc[j];
                else out[cur * ldo + lane + 32 * j] = acc[j];
            }
            acc[j] = 0.f;
        }
        piece_is_carry = false;
    };
    int rel_end = row_end_rel();
    const int rel_own = (int)(own_end - w0);
    for (int i = 0; i < n_groups; ++i) {
        const int rel0 = i * 32;
        const int cnt = min(32, rel_w1 - rel0);
        int32_t a = 0, bb = 0;
        float v = 0.f;
        if (lane < cnt) { a = __ldg(i1 + w0 + rel0 + lane); bb = __ldg(i2 + w0 + rel0 + lane); v = __ldg(values + w0 + rel0 + lane); }
        int t = 0;
        while (t < cnt) {
            while (rel0 + t == rel_end) { emit(); ++cur; rel_end = row_end_rel(); }
            const int seg_end = min(cnt, rel_end - rel0);
            for (; t + 2 <= seg_end; t += 2) {
                const int32_t a0 = __shfl_sync(0xffffffffu, a, t), a1 = __shfl_sync(0xffffffffu, a, t + 1);
                const int32_t b0 = __shfl_sync(0xffffffffu, bb, t), b1 = __shfl_sync(0xffffffffu, bb, t + 1);
                const float v0 = __shfl_sync(0xffffffffu, v, t), v1 = __shfl_sync(0xffffffffu, v, t + 1);
                const float* u0 = U + (int64_t)a0 * ldu; const float* u1 = U + (int64_t)a1 * ldu;
                const float* q0 = W + (int64_t)b0 * ldw; const float* q1 = W + (int64_t)b1 * ldw;
                float x0[J], y0[J], x1[J], y1[J];
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const bool on = xo[j] >= 0;
                    x0[j] = on ? __ldg(u0 + xo[j]) : 0.f; y0[j] = on ? __ldg(q0 + yo[j]) : 0.f;
                    x1[j] = on ? __ldg(u1 + xo[j]) : 0.f; y1[j] = on ? __ldg(q1 + yo[j]) : 0.f;
                }
#pragma unroll
                for (int j = 0; j < J; ++j) { acc[j] = fmaf(v0 * x0[j], y0[j], acc[j]); acc[j] = fmaf(v1 * x1[j], y1[j], acc[j]); }
            }
            for (; t < seg_end; ++t) {
                const int32_t a0 = __shfl_sync(0xffffffffu, a, t);
                const int32_t b0 = __shfl_sync(0xffffffffu, bb, t);
                const float v0 = __shfl_sync(0xffffffffu, v, t);
                const float* u0 = U + (int64_t)a0 * ldu;
                const float* q0 = W + (int64_t)b0 * ldw;
#pragma unroll
                for (int j = 0; j < J; ++j)
                    if (xo[j] >= 0) acc[j] = fmaf(v0 * __ldg(u0 + xo[j]), __ldg(q0 + yo[j]), acc[j]);
            }
        }
    }
    emit();
    while (rel_end < rel_own) {                            // empty rows at the end pointer (last window / empty tensor)
        ++cur;
        if (cur >= n0) break;
        rel_end = row_end_rel();
        emit();
    }
}

__global__ void ttm_fixup_kernel(const float* __restrict__ carry, const int64_t* __restrict__ carry_row, int64_t n_windows,
                                 float* __restrict__ out, int64_t ldo, int width, int stride) {
    const int64_t b = blockIdx.x;
    const int64_t r = carry_row[b];
    if (r < 0) return;
    if (b > 0 && carry_row[b - 1] == r) return;
    for (int c = threadIdx.x; c < width; c += blockDim.x) {
        float y = out[r * ldo + c];
        for (int64_t bb = b; bb < n_windows && carry_row[bb] == r; ++bb) y += carry[bb * stride + c];
        out[r * ldo + c] = y;
    }
}

// ---------------- gathered cross-Gram for the few-segment mode ------------------------
constexpr int GT = 64, GR = 32;

__global__ void __launch_bounds__(256)
xgram_partial_kernel(const float* __restrict__ A, int ra, int64_t lda, const float* __restrict__ B, int rb,
                     int64_t ldb, const int32_t* __restrict__ ia, const int32_t* __restrict__ ib,
                     const float* __restrict__ val, int64_t p_lo, int64_t p_hi, int64_t rows_per_block,
                     int tiles_b, double* __restrict__ partial) {
    const int ti = blockIdx.y / tiles_b, tj = blockIdx.y % tiles_b;
    __shared__ float sa[GR][GT + 4];
    __shared__ float sb[GR][GT + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    const int64_t r0 = p_lo + (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = min(p_hi, r0 + rows_per_block);
    for (int64_t base = r0; base < r1; base += GR) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            int e = threadIdx.x + it * 256;
            int rr = e >> 6, cc = e & 63;
            int64_t p = base + rr;
            float va = 0.f, vb = 0.f;
            if (p < r1) {
                int ca = ti * GT + cc, cb = tj * GT + cc;
                if (ca < ra) va = __ldg(val + p) * __ldg(A + (int64_t)__ldg(ia + p) * lda + ca);
                if (cb < rb) vb = __ldg(B + (int64_t)__ldg(ib + p) * ldb + cb);
            }
            sa[rr][cc] = va;
            sb[rr][cc] = vb;
        }
        __syncthreads();
#pragma unroll 4
        for (int rr = 0; rr < GR; ++rr) {
            float4 a4 = *reinterpret_cast<const float4*>(&sa[rr][ty * 4]);
            float4 b4 = *reinterpret_cast<const float4*>(&sb[rr][tx * 4]);
            double a[4] = {(double)a4.x, (double)a4.y, (double)a4.z, (double)a4.w};
            double b[4] = {(double)b4.x, (double)b4.y, (double)b4.z, (double)b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    double* out = partial + ((int64_t)blockIdx.x * gridDim.y + blockIdx.y) * (GT * GT);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) out[(ty * 4 + i) * GT + tx * 4 + j] = acc[i][j];
}

__global__ void xgram_reduce_kernel(const double* __restrict__ partial, int nblk, int ntiles, int tiles_b, int ra,
                                    int rb, float* __restrict__ out /* [ra*rb] */) {
    int tile = blockIdx.x;
    int ti = tile / tiles_b, tj = tile % tiles_b;
    for (int e = threadIdx.x; e < GT * GT; e += blockDim.x) {
        int x = ti * GT + e / GT, y = tj * GT + e % GT;
        if (x >= ra || y >= rb) continue;
        double s = 0.0;
        for (int b = 0; b < nblk; ++b) s += partial[((int64_t)b * ntiles + tile) * (GT * GT) + e];
        out[(int64_t)x * rb + y] = (float)s;
    }
}

}  // namespace

extern "C" int pb200_ttm(pb200_ctx* ctx, int64_t n0, int64_t nnz, const int64_t* seg_ptr, const int32_t* i1,
                         const int32_t* i2, const float* values, const float* U, int ru, int64_t ldu,
                         const float* W, int rw, int64_t ldw, float* out, int64_t ldo) {
    PB_ENTER(ctx);
    int width = ru * rw;
    PB_REQUIRE(ctx, ru > 0 && rw > 0 && width <= 1024, "ttm: need ru*rw <= 1024");
    PB_REQUIRE(ctx, ldo >= width && ldu >= ru && ldw >= rw, "ttm: leading dimension too small");
    if (n0 == 0) return PB200_OK;
    if (width <= 512 && ctx->spmm_kernel >= 3) {
        // nnz windows per warp + carried row pieces (see ttm_window_kernel): balanced for skewed groupings
        Scratch sc(ctx);
        const int64_t n_windows = std::max<int64_t>(1, ceil_div64(nnz, TW));
        const int jj = width <= 128 ? 4 : width <= 256 ? 8 : 16;
        float* carry = nullptr;
        int64_t* carry_row = nullptr;
        PB_TRY(sc.alloc(&carry, (size_t)n_windows * 32 * jj));
        PB_TRY(sc.alloc(&carry_row, (size_t)n_windows));
        const unsigned blocks = (unsigned)ceil_div64(n_windows, TWARPS);
#define PB_TTMW_LAUNCH(JJ) ttm_window_kernel<JJ><<<blocks, TWARPS * 32, 0, ctx->stream>>>(n0, nnz, seg_ptr, i1, i2, values, U, ru, ldu, W, rw, ldw, out, ldo, n_windows, carry, carry_row)
        if (jj == 4) PB_TTMW_LAUNCH(4);
        else if (jj == 8) PB_TTMW_LAUNCH(8);
        else PB_TTMW_LAUNCH(16);
#undef PB_TTMW_LAUNCH
        ttm_fixup_kernel<<<(unsigned)n_windows, 128, 0, ctx->stream>>>(carry, carry_row, n_windows, out, ldo, width, 32 * jj);
        ctx->stats[0] += 2;
        PB_CUDA(ctx, cudaGetLastError());
        return PB200_OK;
    }
    int64_t n_blocks = std::max<int64_t>(1, ceil_div64(nnz, CB));
    dim3 grid((unsigned)n_blocks), block(WARPS * 32);
#define PB_TTM_LAUNCH(JJ) ttm_kernel<JJ><<<grid, block, 0, ctx->stream>>>(n0, nnz, seg_ptr, i1, i2, values, U, ru, ldu, W, rw, ldw, out, ldo, n_blocks)
    if (width <= 128) PB_TTM_LAUNCH(4);
    else if (width <= 256) PB_TTM_LAUNCH(8);
    else if (width <= 512) PB_TTM_LAUNCH(16);
    else PB_TTM_LAUNCH(32);
#undef PB_TTM_LAUNCH
    ctx->stats[0] += 1;
    PB_CUDA(ctx, cudaGetLastError());
    return PB200_OK;
}

extern "C" int pb200_ttm_reduce(pb200_ctx* ctx, int n_seg, int64_t nnz, const int64_t* seg_ptr_host_or_dev,
                                const int32_t* ia, const int32_t* ib, const float* values, const float* A, int ra,
                                int64_t lda, const float* B, int rb, int64_t ldb, float* out, int64_t ldo) {
    PB_ENTER(ctx);
    PB_REQUIRE(ctx, n_seg > 0 && n_seg <= 4096, "ttm_reduce: 1..4096 segments");
    PB_REQUIRE(ctx, ra > 0 && rb > 0 && ldo >= (int64_t)ra * rb, "ttm_reduce: bad shape");
    std::vector<int64_t> seg(n_seg + 1);
    PB_CUDA(ctx, cudaMemcpyAsync(seg.data(), seg_ptr_host_or_dev, sizeof(int64_t) * (n_seg + 1),
                                 cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    PB_REQUIRE(ctx, seg[n_seg] == nnz, "ttm_reduce: seg_ptr does not end at nnz");
    Scratch sc(ctx);
    int tiles_a = (ra + GT - 1) / GT, tiles_b = (rb + GT - 1) / GT, ntiles = tiles_a * tiles_b;
    int max_blk = 2 * ctx->num_sms;
    double* partial = nullptr;
    PB_TRY(sc.alloc(&partial, (size_t)max_blk * ntiles * GT * GT));
    for (int s = 0; s < n_seg; ++s) {
        int64_t lo = seg[s], hi = seg[s + 1], len = hi - lo;
        int nblk = (int)std::min<int64_t>(std::max<int64_t>(1, ceil_div64(len, 1024)), max_blk);
        int64_t rpb = ceil_div64(std::max<int64_t>(len, 1), nblk);
        rpb = ceil_div64(rpb, GR) * GR;
        nblk = (int)std::max<int64_t>(1, ceil_div64(std::max<int64_t>(len, 1), rpb));
        xgram_partial_kernel<<<dim3(nblk, ntiles), 256, 0, ctx->stream>>>(A, ra, lda, B, rb, ldb, ia, ib, values, lo,
                                                                          hi, rpb, tiles_b, partial);
        xgram_reduce_kernel<<<ntiles, 256, 0, ctx->stream>>>(partial, nblk, ntiles, tiles_b, ra, rb, out + s * ldo);
        ctx->stats[0] += 2;
    }
    PB_CUDA(ctx, cudaGetLastError());
    return PB200_OK;
}
