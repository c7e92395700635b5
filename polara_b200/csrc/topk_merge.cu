// Merge `parts` sorted candidate lists per user into the final top-k (one warp per user)
// and, if fewer than k unseen items exist, append the user's seen items by descending
// score -- the order that downvote_seen_items (polara/recommender/models.py:517-519:
// seen scores are pushed below the minimum but keep their mutual order) followed by
// get_topk_elements (models.py:561-563) produces.
#include <algorithm>

#include "topk_common.cuh"

namespace {

constexpr int MAX_PARTS_PER_LANE = 8;   // up to 256 parts

__global__ void __launch_bounds__(256)
merge_lists_kernel(const pb200_cand* __restrict__ lists, int parts, int64_t part_stride, int64_t m, int k,
                   int64_t item_offset, int64_t* __restrict__ out_ids, float* __restrict__ out_scores,
                   pb200_cand* __restrict__ out_cands, const float* __restrict__ E, int64_t lde,
                   const float* __restrict__ V, int64_t ldv, int r, int64_t n,
                   const int64_t* __restrict__ seen_indptr, const int32_t* __restrict__ seen_indices,
                   const unsigned char* __restrict__ todo) {
    const int lane = threadIdx.x & 31;
    const int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (u >= m) return;
    if (todo && !todo[u]) return;                 // already merged by the per-thread fast path
    // lane owns parts lane, lane+32, ... ; head[] = next unread position in each
    int head[MAX_PARTS_PER_LANE];
#pragma unroll
    for (int i = 0; i < MAX_PARTS_PER_LANE; ++i) head[i] = 0;
    int produced = 0;
    for (; produced < k; ++produced) {
        float bs = -CUDART_INF_F; int bi = -1, bslot = -1;
#pragma unroll
        for (int i = 0; i < MAX_PARTS_PER_LANE; ++i) {
            int p = lane + 32 * i;
            if (p < parts && head[i] < k) {
                pb200_cand c = lists[(int64_t)p * part_stride + u * k + head[i]];
                if (c.id >= 0 && (bi < 0 || cand_before(c.score, c.id, bs, bi))) { bs = c.score; bi = c.id; bslot = i; }
            }
        }
        // warp argmax under the (score desc, id asc) order
        float ws = bs; int wi = bi; int wl = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float os = __shfl_xor_sync(0xffffffffu, ws, o);
            int oi = __shfl_xor_sync(0xffffffffu, wi, o);
            int ol = __shfl_xor_sync(0xffffffffu, wl, o);
            bool take = (oi >= 0) && (wi < 0 || cand_before(os, oi, ws, wi) || (os == ws && oi == wi && ol < wl));
            if (take) { ws = os; wi = oi; wl = ol; }
        }
        if (wi < 0) break;                       // all lists exhausted
        if (lane == wl) {
#pragma unroll
            for (int i = 0; i < MAX_PARTS_PER_LANE; ++i) if (i == bslot) head[i]++;
        }
        if (lane == 0) {
            if (out_ids) out_ids[u * k + produced] = (int64_t)wi + item_offset;
            if (out_scores) out_scores[u * k + produced] = ws;
            if (out_cands) { pb200_cand c; c.score = ws; c.id = (int)(wi + item_offset); out_cands[u * k + produced] = c; }
        }
    }
    if (produced == k) return;
    // ---- fewer than k unseen items: continue with seen ones, best first -------------
    int64_t sb = 0, se = 0;
    if (seen_indptr && E && V) { sb = seen_indptr[u]; se = seen_indptr[u + 1]; }
    float prev_s = CUDART_INF_F; int prev_i = -1;   // last emitted (strictly ordered walk)
    while (produced < k) {
        float bs = -CUDART_INF_F; int bi = -1;
        for (int64_t p = sb + lane; p < se; p += 32) {
            int it = (int)(__ldg(seen_indices + p) - item_offset);      // seen ids are global ids
            if (it < 0 || it >= n) continue;
            float s = exact_score(E + u * lde, V + (int64_t)it * ldv, r);
            // strictly after the previously emitted (prev_s, prev_i) in the total order
            bool after_prev = (prev_i < 0) || cand_before(prev_s, prev_i, s, it);
            if (after_prev && (bi < 0 || cand_before(s, it, bs, bi))) { bs = s; bi = it; }
        }
        float ws = bs; int wi = bi;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float os = __shfl_xor_sync(0xffffffffu, ws, o);
            int oi = __shfl_xor_sync(0xffffffffu, wi, o);
            if (oi >= 0 && (wi < 0 || cand_before(os, oi, ws, wi))) { ws = os; wi = oi; }
        }
        if (wi < 0) break;
        if (lane == 0) {
            if (out_ids) out_ids[u * k + produced] = (int64_t)wi + item_offset;
            if (out_scores) out_scores[u * k + produced] = ws;
            if (out_cands) { pb200_cand c; c.score = ws; c.id = (int)(wi + item_offset); out_cands[u * k + produced] = c; }
        }
        prev_s = ws; prev_i = wi;
        ++produced;
    }
    for (int j = produced + lane; j < k; j += 32) {   // nothing left: pad (reference pads with -1, models.py:73)
        if (out_ids) out_ids[u * k + j] = -1;
        if (out_scores) out_scores[u * k + j] = -CUDART_INF_F;
        if (out_cands) { pb200_cand c; c.score = -CUDART_INF_F; c.id = -1; out_cands[u * k + j] = c; }
    }
}

// Fast path for a handful of lists per user (the common case: two half-lists of the tensor-core sweep plus the
// probe list): one THREAD per user walks the list heads.  Users with fewer than k candidates in total are left
// to the warp kernel above (flagged in `todo`), which also knows how to append seen items.
constexpr int SMALL_PARTS = 8;
__global__ void __launch_bounds__(256)
merge_small_kernel(const pb200_cand* __restrict__ lists, int parts, int64_t part_stride, int64_t m, int k,
                   int64_t item_offset, int64_t* __restrict__ out_ids, float* __restrict__ out_scores,
                   pb200_cand* __restrict__ out_cands, unsigned char* __restrict__ todo) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= m) return;
    int head[SMALL_PARTS];
    pb200_cand cur[SMALL_PARTS];
#pragma unroll
    for (int p = 0; p < SMALL_PARTS; ++p) {
        head[p] = 0;
        cur[p].score = -CUDART_INF_F; cur[p].id = -1;
        if (p < parts) cur[p] = lists[(int64_t)p * part_stride + u * k];
    }
    int produced = 0;
    for (; produced < k; ++produced) {
        int best = -1, bi = -1;
        float bs = -CUDART_INF_F;
#pragma unroll
        for (int p = 0; p < SMALL_PARTS; ++p)
            if (p < parts && cur[p].id >= 0 && (best < 0 || cand_before(cur[p].score, cur[p].id, bs, bi))) {
                best = p; bs = cur[p].score; bi = cur[p].id;
            }
        if (best < 0) break;
        pb200_cand w;
#pragma unroll
        for (int p = 0; p < SMALL_PARTS; ++p) if (p == best) {
            w = cur[p];
            ++head[p];
            if (head[p] < k) cur[p] = lists[(int64_t)p * part_stride + u * k + head[p]]; else { cur[p].score = -CUDART_INF_F; cur[p].id = -1; }
        }
        if (out_ids) out_ids[u * k + produced] = (int64_t)w.id + item_offset;
        if (out_scores) out_scores[u * k + produced] = w.score;
        if (out_cands) { pb200_cand c; c.score = w.score; c.id = (int)(w.id + item_offset); out_cands[u * k + produced] = c; }
    }
    todo[u] = produced < k ? 1 : 0;
}

__global__ void fill_empty_cands_kernel(pb200_cand* __restrict__ c, int64_t count) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < count; i += stride) { pb200_cand e; e.score = -CUDART_INF_F; e.id = -1; c[i] = e; }
}

}  // namespace

extern "C" int pb200_fill_empty_cands(pb200_ctx* ctx, pb200_cand* cands, int64_t count) {
    PB_ENTER(ctx);
    if (count <= 0) return PB200_OK;
    fill_empty_cands_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(count, 256), 8 * (int64_t)ctx->num_sms), 256, 0, ctx->stream>>>(cands, count);
    ctx->stats[0] += 1;
    PB_CUDA(ctx, cudaGetLastError());
    return PB200_OK;
}

int pb_merge_lists(pb200_ctx* ctx, const pb200_cand* lists, int parts, int64_t part_stride, int64_t m, int k,
                   int64_t item_offset, int64_t* out_ids, float* out_scores, pb200_cand* out_cands,
                   const float* E, int64_t lde, const float* V, int64_t ldv, int r, int64_t n,
                   const int64_t* seen_indptr, const int32_t* seen_indices) {
    PB_REQUIRE(ctx, parts >= 1 && parts <= 32 * MAX_PARTS_PER_LANE, "merge: parts must be in 1..256");
    if (m == 0) return PB200_OK;
    unsigned blocks = (unsigned)ceil_div64(m * 32, 256);
    unsigned char* todo = nullptr;
    Scratch sc(ctx);
    if (parts <= SMALL_PARTS) {
        PB_TRY(sc.alloc(&todo, (size_t)m));
        merge_small_kernel<<<(unsigned)ceil_div64(m, 256), 256, 0, ctx->stream>>>(lists, parts, part_stride, m, k, item_offset,
                                                                                  out_ids, out_scores, out_cands, todo);
        ctx->stats[0] += 1;
    }
    merge_lists_kernel<<<blocks, 256, 0, ctx->stream>>>(lists, parts, part_stride, m, k, item_offset, out_ids,
                                                        out_scores, out_cands, E, lde, V, ldv, r, n, seen_indptr,
                                                        seen_indices, todo);
    ctx->stats[0] += 1;
    PB_CUDA(ctx, cudaGetLastError());
    return PB200_OK;
}
