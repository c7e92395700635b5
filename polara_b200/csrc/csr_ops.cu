// CSR format helpers: transpose (A -> CSR of A^T) and the ScaledSVD row/column scaling.
//
// The transpose is a one-time format conversion per build() (scipy does the analogous
// coo->csr/csc conversion at polara/recommender/models.py:169-174).  It is a stable
// radix sort of nnz positions keyed by column (CUB DeviceRadixSort -- library code, used
// for this format conversion only, not on the scoring hot path), followed by gathers.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace {

__global__ void expand_rows_kernel(const int64_t* __restrict__ indptr, int64_t n_rows, int32_t* __restrict__ rows,
                                   uint32_t* __restrict__ pos, int64_t nnz) {
    // one warp per row
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = w; r < n_rows; r += nw) {
        int64_t b = indptr[r], e = indptr[r + 1];
        for (int64_t p = b + lane; p < e; p += 32) { rows[p] = (int32_t)r; pos[p] = (uint32_t)p; }
    }
}

__global__ void gather_transposed_kernel(const uint32_t* __restrict__ perm, const int32_t* __restrict__ rows,
                                         const float* __restrict__ values, int64_t nnz,
                                         int32_t* __restrict__ t_indices, float* __restrict__ t_values) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < nnz; i += stride) {
        uint32_t p = perm[i];
        t_indices[i] = rows[p];
        t_values[i] = values[p];
    }
}

__global__ void gather_group_kernel(const uint32_t* __restrict__ perm, const int32_t* __restrict__ a,
                                    const int32_t* __restrict__ b, const float* __restrict__ val, int64_t nnz,
                                    int32_t* __restrict__ ao, int32_t* __restrict__ bo, float* __restrict__ vo) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < nnz; i += stride) {
        uint32_t p = perm[i];
        ao[i] = a[p]; bo[i] = b[p]; vo[i] = val[p];
    }
}

__global__ void iota_kernel(uint32_t* __restrict__ pos, int64_t nnz) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < nnz; i += stride) pos[i] = (uint32_t)i;
}

__global__ void indptr_from_sorted_kernel(const int32_t* __restrict__ sorted_keys, int64_t nnz, int64_t n_keys,
                                          int64_t* __restrict__ indptr) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_keys) return;
    // first position whose key >= c
    int64_t lo = 0, hi = nnz;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if ((int64_t)sorted_keys[mid] < c) lo = mid + 1; else hi = mid;
    }
    indptr[c] = lo;
}

__global__ void count_cols_kernel(const int32_t* __restrict__ indices, int64_t nnz, int32_t* __restrict__ counts) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < nnz; i += stride) atomicAdd(counts + indices[i], 1);   // integer counts: order-independent
}

__global__ void rescale_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                               float* __restrict__ values, int64_t n_rows, const int32_t* __restrict__ col_counts,
                               double row_pow, double col_pow, int do_rows, int do_cols) {
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = w; r < n_rows; r += nw) {
        int64_t b = indptr[r], e = indptr[r + 1];
        // matrices.py:78-83: norm = sqrt(count); factor = norm**(scaling-1) where norm != 0
        double rf = 1.0;
        if (do_rows && e > b) rf = pow(sqrt((double)(e - b)), row_pow);
        for (int64_t p = b + lane; p < e; p += 32) {
            double v = (double)values[p] * rf;
            if (do_cols) v *= pow(sqrt((double)col_counts[indices[p]]), col_pow);
            values[p] = (float)v;
        }
    }
}

}  // namespace

extern "C" int pb200_csr_transpose(pb200_ctx* ctx, int64_t n_rows, int64_t n_cols, int64_t nnz,
                                   const int64_t* indptr, const int32_t* indices, const float* values,
                                   int64_t* t_indptr, int32_t* t_indices, float* t_values) {
    if (!ctx) return PB200_EINVAL;
    PB_REQUIRE(ctx, nnz < (int64_t)4294967295ll, "transpose: nnz must be < 2^32");
    PB_REQUIRE(ctx, n_cols < (int64_t)2147483647 && n_rows < (int64_t)2147483647, "transpose: dimension must fit int32");
    Scratch sc(ctx);
    int32_t *rows = nullptr, *keys_out = nullptr;
    uint32_t *pos = nullptr, *perm = nullptr;
    PB_TRY(sc.alloc(&rows, (size_t)nnz));
    PB_TRY(sc.alloc(&keys_out, (size_t)nnz));
    PB_TRY(sc.alloc(&pos, (size_t)nnz));
    PB_TRY(sc.alloc(&perm, (size_t)nnz));
    int blocks = 8 * ctx->num_sms;
    expand_rows_kernel<<<blocks, 256, 0, ctx->stream>>>(indptr, n_rows, rows, pos, nnz);
    int bits = 1;
    while (((int64_t)1 << bits) < n_cols) ++bits;
    size_t temp_bytes = 0;
    PB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, indices, keys_out, pos, perm, nnz, 0, bits, ctx->stream));
    uint8_t* temp = nullptr;
    PB_TRY(sc.alloc(&temp, temp_bytes));
    PB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(temp, temp_bytes, indices, keys_out, pos, perm, nnz, 0, bits, ctx->stream));
    gather_transposed_kernel<<<blocks, 256, 0, ctx->stream>>>(perm, rows, values, nnz, t_indices, t_values);
    indptr_from_sorted_kernel<<<(unsigned)ceil_div64(n_cols + 1, 256), 256, 0, ctx->stream>>>(keys_out, nnz, n_cols, t_indptr);
    ctx->stats[0] += 4;
    PB_CUDA(ctx, cudaGetLastError());
    return PB200_OK;
}

extern "C" int pb200_rescale(pb200_ctx* ctx, int64_t n_rows, int64_t n_cols, int64_t nnz,
                             const int64_t* indptr, const int32_t* indices, float* values,
                             double row_scaling, double col_scaling) {
    if (!ctx) return PB200_EINVAL;
    int do_rows = row_scaling != 1.0, do_cols = col_scaling != 1.0;
    if (!do_rows && !do_cols) return PB200_OK;
    Scratch sc(ctx);
    int32_t* counts = nullptr;
    PB_TRY(sc.alloc(&counts, (size_t)n_cols));
    int blocks = 8 * ctx->num_sms;
    if (do_cols) {
        PB_CUDA(ctx, cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)n_cols, ctx->stream));
        count_cols_kernel<<<blocks, 256, 0, ctx->stream>>>(indices, nnz, counts);
        ctx->stats[0] += 1;
        PB_TRY(pb_reduce(ctx, counts, n_cols, PB200_I32));       // row-sharded matrix: column counts are global
    }
    // NOTE the reference scales rows first and recounts nothing in between: both counts are
    // structural nnz counts of the same pattern (matrices.py:79, binary=True), so one pass suffices.
    rescale_kernel<<<blocks, 256, 0, ctx->stream>>>(indptr, indices, values, n_rows, counts,
                                                    row_scaling - 1.0, col_scaling - 1.0, do_rows, do_cols);
    ctx->stats[0] += 1;
    PB_CUDA(ctx, cudaGetLastError());
    return PB200_OK;
}

// Stable grouping of a 3-way COO tensor by one mode (arrange_indices, polara/lib/sparse.py:239-264,
// does the analogous host-side preparation for dttm_par).
extern "C" int pb200_coo_group(pb200_ctx* ctx, int64_t nnz, int64_t n_keys, const int32_t* key, const int32_t* a,
                               const int32_t* b, const float* val, int64_t* seg_ptr, int32_t* a_out,
                               int32_t* b_out, float* val_out) {
    if (!ctx) return PB200_EINVAL;
    PB_REQUIRE(ctx, nnz < (int64_t)4294967295ll && n_keys > 0, "coo_group: nnz must be < 2^32, n_keys > 0");
    Scratch sc(ctx);
    int32_t* keys_out = nullptr;
    uint32_t *pos = nullptr, *perm = nullptr;
    PB_TRY(sc.alloc(&keys_out, (size_t)nnz));
    PB_TRY(sc.alloc(&pos, (size_t)nnz));
    PB_TRY(sc.alloc(&perm, (size_t)nnz));
    int blocks = 8 * ctx->num_sms;
    iota_kernel<<<blocks, 256, 0, ctx->stream>>>(pos, nnz);
    int bits = 1;
    while (((int64_t)1 << bits) < n_keys) ++bits;
    size_t temp_bytes = 0;
    PB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, key, keys_out, pos, perm, nnz, 0, bits, ctx->stream));
    uint8_t* temp = nullptr;
    PB_TRY(sc.alloc(&temp, temp_bytes));
    PB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(temp, temp_bytes, key, keys_out, pos, perm, nnz, 0, bits, ctx->stream));
    gather_group_kernel<<<blocks, 256, 0, ctx->stream>>>(perm, a, b, val, nnz, a_out, b_out, val_out);
    indptr_from_sorted_kernel<<<(unsigned)ceil_div64(n_keys + 1, 256), 256, 0, ctx->stream>>>(keys_out, nnz, n_keys, seg_ptr);
    ctx->stats[0] += 4;
    PB_CUDA(ctx, cudaGetLastError());
    return PB200_OK;
}
