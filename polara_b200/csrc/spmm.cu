// CSR SpMM  Y[n_rows x ell] = A * X  (fp32), rows owned by exactly one warp or one block
// so the result is deterministic (fixed summation order, no atomics).
//
// Replaces csr_matrix.dot(ndarray) (polara/recommender/models.py:860) and the
// A x / A^T x products inside scipy.sparse.linalg.svds (models.py:844).
//
// Work split: block b owns the rows whose first nnz lies in [b*CB, (b+1)*CB).
//   - rows up to LONG_ROW nnz: one warp per row, lanes read (col,val) coalesced in
//     batches of 32, broadcast them with shuffles and gather X rows (each lane owns
//     LPT columns: col = lane + 32*j -> every gather is a coalesced 128 B segment).
//   - longer rows (popular items in A^T): all warps of the block take interleaved
//     batches, partial sums meet in shared memory and are added in warp order.
// HBM-bound on (8 B * nnz + 4*ell*(rows+cols)); the X gather is served by L2.
#include "common.cuh"

namespace {

constexpr int CB = 2048;        // nnz window per block
constexpr int WARPS = 8;
constexpr int LONG_ROW = 4096;  // rows longer than this are processed by the whole block
constexpr int MAX_LONG = CB / LONG_ROW + 2;

__device__ __forceinline__ int64_t lower_bound_i64(const int64_t* a, int64_t n, int64_t key) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

template <int LPT, bool FULL>
__device__ __forceinline__ void accumulate_range(float (&acc)[LPT], int64_t beg, int64_t end,
                                                 int64_t step_batches, const int32_t* __restrict__ indices,
                                                 const float* __restrict__ values,
                                                 const float* __restrict__ X, int64_t ldx, int lane, int live) {
    // columns >= live are padding: their lanes neither load nor accumulate (fewer 32-byte sectors per gathered row)
    bool on[LPT];
#pragma unroll
    for (int j = 0; j < LPT; ++j) on[j] = FULL || lane + 32 * j < live;      // FULL: no predicates in the generated code
    // processes batches [beg + b*32*step_batches ...) ; step_batches = 1 for a warp-owned row
    for (int64_t p = beg; p < end; p += 32 * step_batches) {
        int64_t q = p + lane;
        int32_t c = 0;
        float v = 0.f;
        if (q < end) { c = __ldg(indices + q); v = __ldg(values + q); }
        int cnt = (int)min((int64_t)32, end - p);
        int t = 0;
        for (; t + 4 <= cnt; t += 4) {
            int32_t c0 = __shfl_sync(0xffffffffu, c, t), c1 = __shfl_sync(0xffffffffu, c, t + 1);
            int32_t c2 = __shfl_sync(0xffffffffu, c, t + 2), c3 = __shfl_sync(0xffffffffu, c, t + 3);
            float v0 = __shfl_sync(0xffffffffu, v, t), v1 = __shfl_sync(0xffffffffu, v, t + 1);
            float v2 = __shfl_sync(0xffffffffu, v, t + 2), v3 = __shfl_sync(0xffffffffu, v, t + 3);
            const float* x0 = X + (int64_t)c0 * ldx + lane;
            const float* x1 = X + (int64_t)c1 * ldx + lane;
            const float* x2 = X + (int64_t)c2 * ldx + lane;
            const float* x3 = X + (int64_t)c3 * ldx + lane;
            float a0[LPT], a1[LPT], a2[LPT], a3[LPT];
#pragma unroll
            for (int j = 0; j < LPT; ++j) {
                a0[j] = on[j] ? __ldg(x0 + 32 * j) : 0.f; a1[j] = on[j] ? __ldg(x1 + 32 * j) : 0.f;
                a2[j] = on[j] ? __ldg(x2 + 32 * j) : 0.f; a3[j] = on[j] ? __ldg(x3 + 32 * j) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < LPT; ++j) {
                acc[j] = fmaf(v0, a0[j], acc[j]); acc[j] = fmaf(v1, a1[j], acc[j]);
                acc[j] = fmaf(v2, a2[j], acc[j]); acc[j] = fmaf(v3, a3[j], acc[j]);
            }
        }
        for (; t < cnt; ++t) {
            int32_t c0 = __shfl_sync(0xffffffffu, c, t);
            float v0 = __shfl_sync(0xffffffffu, v, t);
            const float* x0 = X + (int64_t)c0 * ldx + lane;
#pragma unroll
            for (int j = 0; j < LPT; ++j) acc[j] = fmaf(v0, on[j] ? __ldg(x0 + 32 * j) : 0.f, acc[j]);
        }
    }
}

template <int LPT, bool FULL>
__global__ void __launch_bounds__(WARPS * 32)
spmm_csr_kernel(int64_t n_rows, int64_t nnz, const int64_t* __restrict__ indptr,
                const int32_t* __restrict__ indices, const float* __restrict__ values,
                const float* __restrict__ X, int64_t ldx, float* __restrict__ Y, int64_t ldy,
                int64_t n_blocks, int live) {
    __shared__ int64_t s_rows[2];
    __shared__ int s_next;
    __shared__ int s_nlong;
    __shared__ int64_t s_long[MAX_LONG];
    __shared__ float s_part[WARPS][32 * LPT];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t b = blockIdx.x;
    if (threadIdx.x == 0) {
        s_rows[0] = lower_bound_i64(indptr, n_rows, b * (int64_t)CB);
        s_rows[1] = (b == n_blocks - 1) ? n_rows : lower_bound_i64(indptr, n_rows, (b + 1) * (int64_t)CB);
        s_next = 0;
        s_nlong = 0;
    }
    __syncthreads();
    const int64_t row_lo = s_rows[0], row_hi = s_rows[1];
    // ---- warp-owned rows (dynamic assignment: order does not affect results) ----
    for (;;) {
        int idx = 0;
        if (lane == 0) idx = atomicAdd(&s_next, 1);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        int64_t row = row_lo + idx;
        if (row >= row_hi) break;
        int64_t beg = indptr[row], end = indptr[row + 1];
        if (end - beg > LONG_ROW) {
            if (lane == 0) { int s = atomicAdd(&s_nlong, 1); if (s < MAX_LONG) s_long[s] = row; }
            continue;
        }
        float acc[LPT];
#pragma unroll
        for (int j = 0; j < LPT; ++j) acc[j] = 0.f;
        accumulate_range<LPT, FULL>(acc, beg, end, 1, indices, values, X, ldx, lane, live);
        float* y = Y + row * ldy + lane;
#pragma unroll
        for (int j = 0; j < LPT; ++j) y[32 * j] = acc[j];
    }
    __syncthreads();
    // ---- long rows: the whole block, deterministic reduction in warp order ----
    const int nlong = min(s_nlong, MAX_LONG);
    for (int li = 0; li < nlong; ++li) {
        // canonical order of long rows does not matter (each is independent)
        int64_t row = s_long[li];
        int64_t beg = indptr[row], end = indptr[row + 1];
        float acc[LPT];
#pragma unroll
        for (int j = 0; j < LPT; ++j) acc[j] = 0.f;
        accumulate_range<LPT, FULL>(acc, beg + 32 * (int64_t)warp, end, WARPS, indices, values, X, ldx, lane, live);
#pragma unroll
        for (int j = 0; j < LPT; ++j) s_part[warp][lane + 32 * j] = acc[j];
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int j = 0; j < LPT; ++j) {
                float s = 0.f;
                for (int w = 0; w < WARPS; ++w) s += s_part[w][lane + 32 * j];
                Y[row * ldy + lane + 32 * j] = s;
            }
        }
        __syncthreads();
    }
}

}  // namespace

int pb_spmm_impl(pb200_ctx* ctx, int64_t n_rows, int64_t nnz, const int64_t* indptr,
                 const int32_t* indices, const float* values, const float* X, int64_t ldx,
                 float* Y, int64_t ldy, int ell) {
    PB_REQUIRE(ctx, ell > 0, "spmm: ell must be positive");
    PB_REQUIRE(ctx, n_rows >= 0 && nnz >= 0, "spmm: negative size");
    if (n_rows == 0) return PB200_OK;
    int64_t n_blocks = ceil_div64(nnz, CB);
    if (n_blocks == 0) n_blocks = 1;
    PB_REQUIRE(ctx, n_blocks < (int64_t)2147483647, "spmm: nnz too large for one launch");
    // Y is written in whole groups of 32 columns (zeros beyond ell), X is read up to column ell only
    int done = 0;
    while (done < ell) {
        int w = ell - done;
        const float* x = X + done;
        float* y = Y + done;
        dim3 grid((unsigned)n_blocks), block(WARPS * 32);
        if (w > 96) {
            if (w >= 32 * 4) spmm_csr_kernel<4, true><<<grid, block, 0, ctx->stream>>>(n_rows, nnz, indptr, indices, values, x, ldx, y, ldy, n_blocks, w);
            else spmm_csr_kernel<4, false><<<grid, block, 0, ctx->stream>>>(n_rows, nnz, indptr, indices, values, x, ldx, y, ldy, n_blocks, w);
            done += 128;
        } else if (w > 64) {
            if (w >= 32 * 3) spmm_csr_kernel<3, true><<<grid, block, 0, ctx->stream>>>(n_rows, nnz, indptr, indices, values, x, ldx, y, ldy, n_blocks, w);
            else spmm_csr_kernel<3, false><<<grid, block, 0, ctx->stream>>>(n_rows, nnz, indptr, indices, values, x, ldx, y, ldy, n_blocks, w);
            done += 96;
        } else if (w > 32) {
            if (w >= 32 * 2) spmm_csr_kernel<2, true><<<grid, block, 0, ctx->stream>>>(n_rows, nnz, indptr, indices, values, x, ldx, y, ldy, n_blocks, w);
            else spmm_csr_kernel<2, false><<<grid, block, 0, ctx->stream>>>(n_rows, nnz, indptr, indices, values, x, ldx, y, ldy, n_blocks, w);
            done += 64;
        } else {
            if (w >= 32 * 1) spmm_csr_kernel<1, true><<<grid, block, 0, ctx->stream>>>(n_rows, nnz, indptr, indices, values, x, ldx, y, ldy, n_blocks, w);
            else spmm_csr_kernel<1, false><<<grid, block, 0, ctx->stream>>>(n_rows, nnz, indptr, indices, values, x, ldx, y, ldy, n_blocks, w);
            done += 32;
        }
        ctx->stats[0]++;
    }
    PB_CUDA(ctx, cudaGetLastError());
    return PB200_OK;
}

extern "C" int pb200_spmm(pb200_ctx* ctx, int64_t n_rows, int64_t n_cols, int64_t nnz,
                          const int64_t* indptr, const int32_t* indices, const float* values,
                          const float* X, int64_t ldx, float* Y, int64_t ldy, int ell) {
    if (!ctx) return PB200_EINVAL;
    (void)n_cols;
    PB_REQUIRE(ctx, ldx >= ell && ldy >= (ell + 31) / 32 * 32, "spmm: need ldx >= ell and ldy >= ell rounded up to 32");
    return pb_spmm_impl(ctx, n_rows, nnz, indptr, indices, values, X, ldx, Y, ldy, ell);
}
