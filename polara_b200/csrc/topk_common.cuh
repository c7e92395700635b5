// Candidate-list conventions shared by the scoring kernels.
//
// A list holds k entries sorted by (score desc, id asc); empty slots are
// {score=-inf, id=-1}.  The canonical score of (user u, item j) is the fp32 value
//     s = fmaf(E[u][r-1], V[j][r-1], ... fmaf(E[u][0], V[j][0], 0.f))
// (ascending k, single accumulator), so every kernel that "rescores exactly"
// produces bit-identical numbers.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "common.cuh"

__device__ __forceinline__ bool cand_before(float sa, int ia, float sb, int ib) {
    // empty slots (id < 0) sort last
    if (ib < 0) return ia >= 0;
    if (ia < 0) return false;
    return sa > sb || (sa == sb && ia < ib);
}

__device__ __forceinline__ float exact_score(const float* __restrict__ e, const float* __restrict__ v, int r) {
    float s = 0.f;
    for (int t = 0; t < r; ++t) s = fmaf(e[t], v[t], s);
    return s;
}

// true if `item` occurs in the sorted range seen[beg, end)
__device__ __forceinline__ bool seen_lookup(const int32_t* __restrict__ seen, int64_t beg, int64_t end, int item) {
    int64_t lo = beg, hi = end;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        int v = __ldg(seen + mid);
        if (v < item) lo = mid + 1; else hi = mid;
    }
    return lo < end && __ldg(seen + lo) == item;
}

// Warp-cooperative insertion of (s, id) into a sorted list of capacity k living in
// global/shared memory.  `cnt` is the current fill (uniform across the warp); returns
// the new fill.  All 32 lanes must call.
__device__ __forceinline__ int warp_list_insert(pb200_cand* list, int k, int cnt, float s, int id, int lane) {
    if (cnt == k) {
        pb200_cand last = list[k - 1];
        if (!cand_before(s, id, last.score, last.id)) return cnt;
    }
    // position = number of entries ranking before the candidate
    int pos = 0;
    for (int base = 0; base < cnt; base += 32) {
        int i = base + lane;
        bool before = false;
        if (i < cnt) { pb200_cand c = list[i]; before = cand_before(c.score, c.id, s, id); }
        pos += __popc(__ballot_sync(0xffffffffu, before));
    }
    int last_dst = min(cnt, k - 1);           // highest destination index after the shift
    // shift [pos, last_dst-1] -> [pos+1, last_dst], walking from the top in chunks of 32
    for (int hi = last_dst; hi > pos; hi -= 32) {
        int dst = hi - lane;
        pb200_cand c;
        bool act = dst > pos;
        if (act) c = list[dst - 1];
        __syncwarp();
        if (act) list[dst] = c;
        __syncwarp();
    }
    if (lane == 0) { pb200_cand c; c.score = s; c.id = id; list[pos] = c; }
    __syncwarp();
    return min(cnt + 1, k);
}
