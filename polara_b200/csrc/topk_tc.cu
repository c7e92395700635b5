// Fused scoring on the 5th-gen tensor cores (tcgen05 / TMEM), sm_100a.
//
//   scores = E V^T  ->  seen-item mask  ->  per-user top-k          (score rows never reach HBM)
//
// replaces the dgemm of SVDModel.slice_recommendations (polara/recommender/models.py:857-861),
// downvote_seen_items (models.py:494-519) and get_topk_elements (models.py:522-564).
//
// Idea: the tensor cores only FILTER.  Operands are packed to bf16 (A = -E, B = V) together with
// three extra K-slots: a per-user threshold t_w (split hi+lo bf16, B holds 1.0 there) and a per-PAIR
// error margin (A: -2^-7 ||e_u||, B: ||v_j||, both rounded up), so the fp32 accumulator in TMEM is
//     d = t_w - s~ - 2^-7 ||e_u|| ||v_j||     with  s~ = bf16 dot product (|s~ - s| <= 2^-8 ||e|| ||v||).
// Items are swept in order of decreasing ||v_j|| (stable radix sort of the norms, CUB), which makes
// the running thresholds tight after the first tile.  The epilogue reads
// TMEM with tcgen05.ld and keeps ONLY THE SIGN BIT of each accumulator (one SHF per pair):
// sign set  <=>  s~ + margin > t_w  <=>  "candidate".  t_w is a lower bound of the user's final k-th
// best exact score, so no true top-k item can be missed.
// Candidates (a few hundred per user out of 1e5 items) are then checked against the user's seen
// list and RESCORED EXACTLY in fp32 (the canonical fmaf chain of topk_common.cuh), which makes the
// result bit-identical to the exact SIMT kernel (topk_simt.cu).  As better candidates arrive the
// owner thread rewrites t_w inside the A operand in shared memory (generic-proxy store +
// fence.proxy.async), so later MMAs filter harder.
//
// Pipeline per CTA (persistent, one CTA per SM, 10 warps):
//   warp 8  producer : cp.async.bulk (UBLKCP) of pre-packed operand tiles, mbarrier complete_tx
//   warp 9  MMA      : one elected thread issues tcgen05.mma (M=128, N=256, K=16 per instr),
//                      tcgen05.commit releases smem stages / publishes TMEM accumulators
//   warps 0-7 epilogue: tcgen05.ld 32x32b.x32, sign-bit masks, staging, flush (rescoring + lists)
// TMEM holds two 128x256 fp32 accumulators (512 columns) so MMA and epilogue overlap.
#include <cuda_bf16.h>
#include <cub/device/device_radix_sort.cuh>

#include "topk_common.cuh"

namespace {

constexpr int BM = 128;          // users per tile (TMEM lanes)
constexpr int BN = 256;          // items per tile (TMEM columns per accumulator)
constexpr int NEPI_WARPS = 8;
constexpr int NTHREADS = 320;
constexpr int CAPS = 16;         // staged (chunk, mask) entries per epilogue thread
constexpr int MAX_STAGES = 6;
constexpr int PROBE_ITEMS = 256; // items scored exactly up front to seed the thresholds
constexpr long long SPIN_LIMIT_CYCLES = 4000000000ll;

struct TcParams {
    const __nv_bfloat16* Ap;     // packed A tiles [user_tiles][BM x KP]
    const __nv_bfloat16* Bp;     // packed B tiles [item_tiles][BN x KP]
    const float* E; int64_t lde;
    const float* V; int64_t ldv;
    const float* enorm;          // [m] ||e_u||
    const int32_t* perm;         // [n] sweep position -> item id (norm-descending order)
    const float* t0;             // [m] seed lower bound of the k-th best score (or -inf)
    int64_t m, n;
    int r, KP, rs, k;
    int64_t user_tiles, item_tiles;
    int parts; int64_t tiles_per_part;
    const int64_t* seen_indptr; const int32_t* seen_indices; int64_t seen_offset;
    pb200_cand* lists;           // [parts*2][m][k]
    int stages;
    uint32_t a_bytes, b_bytes, sbo;
    unsigned long long* stats;   // device counters
};

// ------------------------------------------------------------------ PTX wrappers --
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(count), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, unsigned long long* stats) {
    uint32_t spins = 0;
    long long t_start = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0xFFFu) == 0) {        // never hang the GPU: after ~2 s record and abort the kernel
            long long now = clock64();
            if (t_start == 0) t_start = now;
            else if (now - t_start > SPIN_LIMIT_CYCLES) {
                if (stats) atomicExch(stats + 7, 0xDEAD0000ull | (bar & 0xFFFFu));
                asm volatile("trap;");
            }
        }
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, no swizzle ("interleave"): 8-row x 16-byte core matrices,
// LBO = byte distance between the two 16 B K-chunks of one instruction, SBO = between 8-row groups.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
           (1ull << 46);
}
// instruction descriptor: D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1), K-major both, N>>3 @17, M>>4 @24
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// bf16 bit helpers (round toward -inf so that thresholds stay conservative)
__device__ __forceinline__ uint32_t bf16_floor_bits(float x) {
    uint32_t b = __float_as_uint(x);
    uint32_t hi = b >> 16;
    if ((b & 0xFFFFu) && (b >> 31)) hi += 1;     // negative: truncation rounds up -> step down
    return hi;
}
__device__ __forceinline__ uint32_t bf16_ceil_pos_bits(float x) {      // x >= 0, round up
    uint32_t b = __float_as_uint(x);
    return (b >> 16) + ((b & 0xFFFFu) ? 1u : 0u);
}
// pack threshold t (<= target) into {hi, lo} bf16 pair, hi + lo <= t
__device__ __forceinline__ uint32_t pack_threshold(float t) {
    if (!(t > -3.0e38f)) t = -3.0e38f;
    uint32_t hi = bf16_floor_bits(t);
    float hif = __uint_as_float(hi << 16);
    float rem = t - hif;                          // >= 0, exact
    uint32_t lo = bf16_floor_bits(rem);
    return (lo << 16) | (hi & 0xFFFFu);
}

// --------------------------------------------------------------- packing kernels --
// element (row, k) of a [rows x KP] K-major tile lives at byte
//   (row/8)*SBO + (k/8)*128 + (row%8)*16 + (k%8)*2 ,  SBO = (KP/8)*128
__global__ void pack_items_kernel(const float* __restrict__ V, int64_t ldv, int64_t n, int r, int rs, int KP,
                                  int64_t item_tiles, const int32_t* __restrict__ perm,
                                  const float* __restrict__ vnorm_sorted, __nv_bfloat16* __restrict__ Bp) {
    const int chunks = KP / 8;
    int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // one 16-byte chunk per thread
    int64_t total = item_tiles * BN * chunks;
    if (gid >= total) return;
    int64_t tile = gid / ((int64_t)BN * chunks);
    int rem = (int)(gid % ((int64_t)BN * chunks));
    int row = rem / chunks, ch = rem % chunks;
    int64_t pos = tile * BN + row;
    const int64_t item = pos < n ? (int64_t)__ldg(perm + pos) : -1;
    __align__(16) __nv_bfloat16 out[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        int kk = ch * 8 + j;
        float x = 0.f;
        if (item >= 0) {
            if (kk < r) x = __ldg(V + item * ldv + kk);
            else if (kk == rs || kk == rs + 1) x = 1.0f;
        }
        out[j] = __float2bfloat16_rn(x);
        if (item >= 0 && kk == rs + 2) out[j] = __ushort_as_bfloat16((unsigned short)bf16_ceil_pos_bits(__ldg(vnorm_sorted + pos)));
    }
    size_t byte = (size_t)tile * BN * KP * 2 + (size_t)(row / 8) * (chunks * 128) + (size_t)ch * 128 + (row % 8) * 16;
    *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(Bp) + byte) = *reinterpret_cast<const uint4*>(out);
}

__global__ void pack_users_kernel(const float* __restrict__ E, int64_t lde, int64_t m, int r, int rs, int KP,
                                  int64_t user_tiles, const float* __restrict__ enorm,
                                  const float* __restrict__ t0, __nv_bfloat16* __restrict__ Ap) {
    const int chunks = KP / 8;
    int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = user_tiles * BM * chunks;
    if (gid >= total) return;
    int64_t tile = gid / ((int64_t)BM * chunks);
    int rem = (int)(gid % ((int64_t)BM * chunks));
    int row = rem / chunks, ch = rem % chunks;
    int64_t u = tile * BM + row;
    __align__(16) __nv_bfloat16 out[8];
    uint32_t thr = 0;
    if (ch == rs / 8) {
        if (u < m) {
            thr = pack_threshold(t0[u]);
        } else {
            thr = 0x00007F7Fu;                                            // +3.39e38: padding rows never fire
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        int kk = ch * 8 + j;
        float x = 0.f;
        if (u < m && kk < r) x = -__ldg(E + u * lde + kk);
        out[j] = __float2bfloat16_rn(x);
        if (kk == rs) out[j] = __ushort_as_bfloat16((unsigned short)(thr & 0xFFFFu));
        if (kk == rs + 1) out[j] = __ushort_as_bfloat16((unsigned short)(thr >> 16));
        // per-pair margin slot: -(2^-7 ||e_u||) rounded away from zero (2x the bf16 product bound 2^-8)
        if (kk == rs + 2 && u < m)
            out[j] = __ushort_as_bfloat16((unsigned short)(0x8000u | bf16_ceil_pos_bits(0.0078125f * enorm[u] + 1e-30f)));
    }
    size_t byte = (size_t)tile * BM * KP * 2 + (size_t)(row / 8) * (chunks * 128) + (size_t)ch * 128 + (row % 8) * 16;
    *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(Ap) + byte) = *reinterpret_cast<const uint4*>(out);
}

// row norms (one warp per row); optional max over rows (positive floats order like ints)
__global__ void row_norm_kernel(const float* __restrict__ X, int64_t ld, int64_t rows, int r, float* __restrict__ norms,
                                float* __restrict__ max_out) {
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (w >= rows) return;
    float s = 0.f;
    for (int t = lane; t < r; t += 32) { float x = __ldg(X + w * ld + t); s = fmaf(x, x, s); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    s = sqrtf(s) * 1.0001f;                    // tiny inflation covers the rounding of the norm itself
    if (lane == 0) {
        if (norms) norms[w] = s;
        if (max_out) atomicMax(reinterpret_cast<int*>(max_out), __float_as_int(s));
    }
}

__global__ void iota_i32_kernel(int32_t* __restrict__ x, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = (int32_t)i;
}

__global__ void gather_rows_kernel(const float* __restrict__ V, int64_t ldv, const int32_t* __restrict__ perm, int64_t rows,
                                   int r, float* __restrict__ out, int64_t ldo) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * r) return;
    int64_t row = i / r; int c = (int)(i % r);
    out[row * ldo + c] = __ldg(V + (int64_t)__ldg(perm + row) * ldv + c);
}

__global__ void seed_threshold_kernel(const pb200_cand* __restrict__ probe_lists, int64_t m, int k, float* __restrict__ t0) {
    int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= m) return;
    pb200_cand c = probe_lists[u * k + (k - 1)];
    t0[u] = (c.id >= 0) ? c.score : -CUDART_INF_F;
}

// ------------------------------------------------------------------ main kernel ---
struct ListState {
    pb200_cand* list;   // k slots in global memory, sorted
    int cnt;
    float kth;          // score of slot k-1 once full, else -inf
};

__device__ __forceinline__ void list_insert(ListState& ls, int k, float s, int id) {
    if (ls.cnt == k) {
        pb200_cand last = ls.list[k - 1];
        if (!cand_before(s, id, last.score, last.id)) return;
    }
    int i = ls.cnt < k ? ls.cnt : k - 1;
    while (i > 0) {
        pb200_cand p = ls.list[i - 1];
        if (!cand_before(s, id, p.score, p.id)) break;
        ls.list[i] = p;
        --i;
    }
    pb200_cand c; c.score = s; c.id = id;
    ls.list[i] = c;
    if (ls.cnt < k) ls.cnt++;
    if (ls.cnt == k) ls.kth = ls.list[k - 1].score;
}

__global__ void __launch_bounds__(NTHREADS, 1)
score_topk_tc_kernel(const TcParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    // ---- carve shared memory -------------------------------------------------------
    unsigned char* sA = smem;
    unsigned char* sB = sA + p.a_bytes;
    uint2* sStage = reinterpret_cast<uint2*>(sB + (size_t)p.stages * p.b_bytes);          // [CAPS][256]
    volatile uint2* sThr = reinterpret_cast<volatile uint2*>(sStage + CAPS * 256);          // [2][128] {work tag, k-th score}
    uint64_t* bars = reinterpret_cast<uint64_t*>(const_cast<uint2*>(sThr) + 256);
    // barrier layout: full[S], empty[S], tfull[2], tempty[2], a_full, a_empty
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + MAX_STAGES);
    const uint32_t bar_tfull = smem_u32(bars + 2 * MAX_STAGES), bar_tempty = smem_u32(bars + 2 * MAX_STAGES + 2);
    const uint32_t bar_afull = smem_u32(bars + 2 * MAX_STAGES + 4), bar_aempty = smem_u32(bars + 2 * MAX_STAGES + 5);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 6);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, NEPI_WARPS); }
        mbar_init(bar_afull, 1);
        mbar_init(bar_aempty, 1 + NEPI_WARPS);
        fence_barrier_init();
    }
    if (warp == 9) { tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int64_t n_work = p.user_tiles * p.parts;
    const int kb = p.KP / 16;                          // MMA instructions per tile

    if (warp == 8) {
        // ============================ producer ======================================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, awork = 0;
            for (int64_t w = blockIdx.x; w < n_work; w += gridDim.x, ++awork) {
                const int64_t ut = w / p.parts; const int part = (int)(w % p.parts);
                const int64_t t_lo = min(p.item_tiles, (int64_t)part * p.tiles_per_part);
                const int64_t t_hi = min(p.item_tiles, (int64_t)(part + 1) * p.tiles_per_part);
                mbar_wait(bar_aempty, (awork & 1) ^ 1, p.stats);
                mbar_arrive_expect_tx(bar_afull, p.a_bytes);
                bulk_g2s(smem_u32(sA), reinterpret_cast<const unsigned char*>(p.Ap) + (size_t)ut * p.a_bytes, p.a_bytes, bar_afull);
                for (int64_t t = t_lo; t < t_hi; ++t) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1, p.stats);
                    mbar_arrive_expect_tx(bar_full + 8 * stage, p.b_bytes);
                    bulk_g2s(smem_u32(sB + (size_t)stage * p.b_bytes),
                             reinterpret_cast<const unsigned char*>(p.Bp) + (size_t)t * p.b_bytes, p.b_bytes,
                             bar_full + 8 * stage);
                    if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 9) {
        // ============================ MMA issuer ====================================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(BM, BN);
            uint32_t stage = 0, phase = 0, awork = 0, tcount = 0;
            for (int64_t w = blockIdx.x; w < n_work; w += gridDim.x, ++awork) {
                const int part = (int)(w % p.parts);
                const int64_t t_lo = min(p.item_tiles, (int64_t)part * p.tiles_per_part);
                const int64_t t_hi = min(p.item_tiles, (int64_t)(part + 1) * p.tiles_per_part);
                mbar_wait(bar_afull, awork & 1, p.stats);
                for (int64_t t = t_lo; t < t_hi; ++t, ++tcount) {
                    const uint32_t acc = tcount & 1, aphase = (tcount >> 1) & 1;
                    mbar_wait(bar_tempty + 8 * acc, aphase ^ 1, p.stats);
                    mbar_wait(bar_full + 8 * stage, phase, p.stats);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB + (size_t)stage * p.b_bytes);
                    for (int ks = 0; ks < kb; ++ks) {
                        uint64_t ad = umma_desc(a0 + ks * 256, 128, p.sbo);
                        uint64_t bd = umma_desc(b0 + ks * 256, 128, p.sbo);
                        tc_mma_bf16(tmem_base + acc * BN, ad, bd, idesc, ks > 0 ? 1u : 0u);
                    }
                    tc_commit(bar_empty + 8 * stage);          // smem stage reusable once these MMAs retire
                    tc_commit(bar_tfull + 8 * acc);            // accumulator ready for the epilogue
                    if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
                }
                tc_commit(bar_aempty);                         // A tile no longer read by the tensor cores
            }
        }
    } else {
        // ============================ epilogue ======================================
        const int q = warp & 3, h = warp >> 2;                 // TMEM lane quarter, column half
        const int row = 32 * q + lane;
        const int etid = warp * 32 + lane;                     // 0..255
        // byte offset of this row's threshold pair inside the packed A tile
        const uint32_t thr_off = (uint32_t)(row / 8) * p.sbo + (uint32_t)(p.rs / 8) * 128 + (row % 8) * 16 + (p.rs % 8) * 2;
        const bool vec_ok = ((p.lde | p.ldv) % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.E) | reinterpret_cast<uintptr_t>(p.V)) % 16 == 0);
        const int r4 = p.r / 4;
        uint32_t awork = 0, tcount = 0;
        unsigned long long n_rescored = 0;
        for (int64_t w = blockIdx.x; w < n_work; w += gridDim.x, ++awork) {
            const int64_t ut = w / p.parts; const int part = (int)(w % p.parts);
            const int64_t t_lo = min(p.item_tiles, (int64_t)part * p.tiles_per_part);
            const int64_t t_hi = min(p.item_tiles, (int64_t)(part + 1) * p.tiles_per_part);
            const int64_t u = ut * BM + row;
            const bool live = u < p.m;
            ListState ls;
            ls.list = p.lists + ((int64_t)(part * 2 + h) * p.m + (live ? u : 0)) * p.k;
            ls.cnt = 0; ls.kth = -CUDART_INF_F;
            if (live) for (int j = 0; j < p.k; ++j) { pb200_cand c; c.score = -CUDART_INF_F; c.id = -1; ls.list[j] = c; }
            float t_row = live ? __ldg(p.t0 + u) : CUDART_INF_F;           // best known lower bound of the k-th score
            float t_written = t_row;
            const float* erow = p.E + (live ? u : 0) * p.lde;
            int64_t sb = 0, se = 0;                                        // this user's seen list (sorted item ids)
            if (live && p.seen_indptr) { sb = p.seen_indptr[u]; se = p.seen_indptr[u + 1]; }
            int scount = 0;
            mbar_wait(bar_afull, awork & 1, p.stats);                      // A tile (and its threshold slots) landed

            auto flush = [&]() {
                for (int e = 0; e < scount; ++e) {
                    uint2 ent = sStage[e * 256 + etid];
                    const int64_t base = (int64_t)(t_lo + (ent.x >> 2)) * BN + h * 128 + (ent.x & 3) * 32;
                    uint32_t mask = ent.y;
                    while (mask) {
                        int c = __clz(mask);                   // column c <-> bit 31-c (first column packed first)
                        mask &= ~(0x80000000u >> c);
                        const int64_t pos = base + c;
                        if (pos >= p.n) continue;
                        const int64_t item = __ldg(p.perm + pos);          // sweep position -> item id
                        if (sb < se && seen_lookup(p.seen_indices, sb, se, (int)(item + p.seen_offset))) continue;   // masked
                        const float* vrow = p.V + item * p.ldv;
                        float s = 0.f;
                        if (vec_ok) {
                            const float4* e4 = reinterpret_cast<const float4*>(erow);
                            const float4* v4 = reinterpret_cast<const float4*>(vrow);
                            for (int t = 0; t < r4; ++t) {
                                float4 a = __ldg(e4 + t), b = __ldg(v4 + t);
                                s = fmaf(a.x, b.x, s); s = fmaf(a.y, b.y, s); s = fmaf(a.z, b.z, s); s = fmaf(a.w, b.w, s);
                            }
                            for (int t = r4 * 4; t < p.r; ++t) s = fmaf(__ldg(erow + t), __ldg(vrow + t), s);
                        } else {
                            s = exact_score(erow, vrow, p.r);
                        }
                        ++n_rescored;
                        if (s < t_row) continue;               // cannot be in the final top-k
                        list_insert(ls, p.k, s, (int)item);
                    }
                }
                scount = 0;
                // share the per-half k-th scores of this row; both are lower bounds of the final k-th score
                sThr[h * 128 + row].y = __float_as_uint(ls.kth);
                sThr[h * 128 + row].x = awork + 1;                     // tag: valid for this work item only
                const uint32_t otag = sThr[(1 - h) * 128 + row].x;
                const float oval = __uint_as_float(sThr[(1 - h) * 128 + row].y);
                // the other half may be one update behind or ahead; any value carrying this work's tag
                // is the k-th score of k real unseen items of this user, hence a valid lower bound
                const float other = (otag == awork + 1) ? oval : -CUDART_INF_F;
                t_row = fmaxf(t_row, fmaxf(ls.kth, other));
                if (live && t_row > t_written) {
                    uint32_t packed = pack_threshold(t_row);
                    *reinterpret_cast<volatile uint32_t*>(sA + thr_off) = packed;
                    fence_proxy_async();                       // make the generic-proxy store visible to the MMA reads
                    t_written = t_row;
                }
            };

            for (int64_t t = t_lo; t < t_hi; ++t, ++tcount) {
                const uint32_t acc = tcount & 1, aphase = (tcount >> 1) & 1;
                mbar_wait(bar_tfull + 8 * acc, aphase, p.stats);
                tc_fence_after();
                const uint32_t tbase = tmem_base + ((uint32_t)(32 * q) << 16) + acc * BN + h * 128;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    uint32_t v[32];
                    tmem_ld32(tbase + c * 32, v);
                    tmem_wait_ld();
                    uint32_t mask = 0;
#pragma unroll
                    for (int i = 0; i < 32; ++i) mask = __funnelshift_l(v[i], mask, 1);   // (mask << 1) | sign(v[i])
                    if (mask && live) {
                        sStage[scount * 256 + etid] = make_uint2((uint32_t)((t - t_lo) << 2) | (uint32_t)c, mask);
                        ++scount;
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
                if (__any_sync(0xffffffffu, scount > CAPS - 4)) flush();
            }
            flush();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_aempty);            // this warp no longer touches the A tile
        }
        if (p.stats && n_rescored) atomicAdd(p.stats + 1, n_rescored);
    }
    // ---- teardown ---------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 9) tmem_dealloc(tmem_base, 512);
}

}  // namespace

int pb_score_tc(pb200_ctx* ctx, const float* E, int64_t lde, const float* V, int64_t ldv, int64_t m, int64_t n,
                int r, const int64_t* seen_indptr, const int32_t* seen_indices, int64_t seen_offset, int k,
                int* parts_out, pb200_cand** lists_out, Scratch& sc) {
    const int rs = (r + 1) & ~1;                       // threshold pair, 4-byte aligned
    const int KP = ((rs + 3) + 15) / 16 * 16;         // + threshold hi/lo + margin slot
    const uint32_t a_bytes = BM * KP * 2, b_bytes = BN * KP * 2;
    const size_t fixed = (size_t)a_bytes + CAPS * 256 * sizeof(uint2) + 256 * sizeof(uint2) + (2 * MAX_STAGES + 8) * 8 + 1024;
    int dev_smem = 0;
    PB_CUDA(ctx, cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
    int stages = (int)std::min<int64_t>(MAX_STAGES, ((int64_t)dev_smem - (int64_t)fixed) / b_bytes);
    if (stages < 2) {
        ctx->err = "tcgen05 scoring kernel: rank too large for the shared-memory pipeline (use the simt kernel)";
        return PB200_ENOTIMPL;
    }
    const int64_t user_tiles = ceil_div64(m, BM), item_tiles = ceil_div64(n, BN);
    int parts = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(ceil_div64(2 * (int64_t)ctx->num_sms, user_tiles), 64), item_tiles));
    const int64_t tiles_per_part = ceil_div64(item_tiles, parts);
    parts = (int)ceil_div64(item_tiles, tiles_per_part);

    __nv_bfloat16 *Ap = nullptr, *Bp = nullptr;
    float *enorm = nullptr, *vnorm = nullptr, *vnorm_sorted = nullptr, *t0 = nullptr, *vprobe = nullptr;
    int32_t *iota = nullptr, *perm = nullptr;
    pb200_cand *probe = nullptr, *lists = nullptr;
    PB_TRY(sc.alloc(&Ap, (size_t)user_tiles * BM * KP));
    PB_TRY(sc.alloc(&Bp, (size_t)item_tiles * BN * KP));
    PB_TRY(sc.alloc(&enorm, (size_t)m));
    PB_TRY(sc.alloc(&vnorm, (size_t)n));
    PB_TRY(sc.alloc(&vnorm_sorted, (size_t)n));
    PB_TRY(sc.alloc(&iota, (size_t)n));
    PB_TRY(sc.alloc(&perm, (size_t)n));
    PB_TRY(sc.alloc(&t0, (size_t)m));
    PB_TRY(sc.alloc(&probe, (size_t)m * k));
    PB_TRY(sc.alloc(&lists, (size_t)parts * 2 * m * k));

    // 1) item norms; sweep order = decreasing norm (stable radix sort; CUB is used for this ordering only)
    row_norm_kernel<<<(unsigned)ceil_div64(n * 32, 256), 256, 0, ctx->stream>>>(V, ldv, n, r, vnorm, nullptr);
    iota_i32_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, ctx->stream>>>(iota, n);
    {
        size_t temp_bytes = 0;
        PB_CUDA(ctx, cub::DeviceRadixSort::SortPairsDescending(nullptr, temp_bytes, vnorm, vnorm_sorted, iota, perm, (int64_t)n, 0, 32, ctx->stream));
        uint8_t* temp = nullptr;
        PB_TRY(sc.alloc(&temp, temp_bytes));
        PB_CUDA(ctx, cub::DeviceRadixSort::SortPairsDescending(temp, temp_bytes, vnorm, vnorm_sorted, iota, perm, (int64_t)n, 0, 32, ctx->stream));
    }
    // 2) exact probe pass over the largest-norm items seeds a lower bound of every user's k-th best score
    const int64_t n_probe = std::min<int64_t>(n, PROBE_ITEMS);
    const int64_t ldp = (r + 3) / 4 * 4;
    PB_TRY(sc.alloc(&vprobe, (size_t)n_probe * ldp));
    gather_rows_kernel<<<(unsigned)ceil_div64(n_probe * r, 256), 256, 0, ctx->stream>>>(V, ldv, perm, n_probe, r, vprobe, ldp);
    PB_TRY(pb_score_simt(ctx, E, lde, vprobe, ldp, m, n_probe, r, seen_indptr, seen_indices, seen_offset, k, 1, probe, perm));
    seed_threshold_kernel<<<(unsigned)ceil_div64(m, 256), 256, 0, ctx->stream>>>(probe, m, k, t0);
    // 3) user norms for the per-pair margin, operand packing
    row_norm_kernel<<<(unsigned)ceil_div64(m * 32, 256), 256, 0, ctx->stream>>>(E, lde, m, r, enorm, nullptr);
    {
        int64_t tot_b = item_tiles * BN * (KP / 8), tot_a = user_tiles * BM * (KP / 8);
        pack_items_kernel<<<(unsigned)ceil_div64(tot_b, 256), 256, 0, ctx->stream>>>(V, ldv, n, r, rs, KP, item_tiles, perm, vnorm_sorted, Bp);
        pack_users_kernel<<<(unsigned)ceil_div64(tot_a, 256), 256, 0, ctx->stream>>>(E, lde, m, r, rs, KP, user_tiles, enorm, t0, Ap);
    }
    // 4) the fused tensor-core kernel
    TcParams p;
    p.Ap = Ap; p.Bp = Bp; p.E = E; p.lde = lde; p.V = V; p.ldv = ldv; p.enorm = enorm; p.perm = perm; p.t0 = t0;
    p.m = m; p.n = n; p.r = r; p.KP = KP; p.rs = rs; p.k = k;
    p.user_tiles = user_tiles; p.item_tiles = item_tiles; p.parts = parts; p.tiles_per_part = tiles_per_part;
    p.seen_indptr = seen_indptr; p.seen_indices = seen_indices; p.seen_offset = seen_offset;
    p.lists = lists; p.stages = stages; p.a_bytes = a_bytes; p.b_bytes = b_bytes; p.sbo = (uint32_t)(KP / 8) * 128;
    p.stats = reinterpret_cast<unsigned long long*>(ctx->d_stats);
    const size_t smem_bytes = fixed + (size_t)stages * b_bytes;
    PB_CUDA(ctx, cudaFuncSetAttribute(score_topk_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    const int64_t n_work = user_tiles * parts;
    const unsigned grid = (unsigned)std::min<int64_t>(n_work, ctx->num_sms);
    cudaEventRecord(ctx->ev0, ctx->stream);
    score_topk_tc_kernel<<<grid, NTHREADS, smem_bytes, ctx->stream>>>(p);
    cudaEventRecord(ctx->ev1, ctx->stream);
    ctx->stats[0] += 11;
    ctx->stats[2] = (uint64_t)item_tiles; ctx->stats[3] = (uint64_t)user_tiles;
    PB_CUDA(ctx, cudaGetLastError());
    *parts_out = parts * 2;
    *lists_out = lists;
    return PB200_OK;
}
