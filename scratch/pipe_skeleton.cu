// Round-2 planning microbenchmark (written after the GPU budget of round 1 was spent; compile-checked only).
// Skeleton of the fused scoring kernel's MMA <-> epilogue ring with static operands: which ring shape turns around fastest?
//   BN   = 128 (4 accumulators of 128 columns, today's shape) or 64 (8 accumulators of 64 columns)
//   ALL  = 0: the two epilogue halves take alternate tiles (4 warps read BN columns each, today's scheme)
//          1: all 8 warps read every tile (each BN/2 columns): shorter read-out, twice the hand-shakes per warp
//   TMA  = 1: a producer streams 16 KB bulk copies into shared memory next to the operands (the B-tile traffic)
// Prints cycles per tile and per 128 items.  One tile = 4 x (M128 x BN x K16) bf16 SS MMAs + commit.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
    return (uint64_t)((a & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity) { while (!try_wait(bar, parity)) {} }
__device__ __forceinline__ void arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}
constexpr int NT = 352;
template <int BN, int ALL>
__global__ void __launch_bounds__(NT, 1) k(int n_tiles, int with_tma, const unsigned char* src, long long* out) {
    constexpr int NACC = 512 / BN;
    constexpr int READERS = ALL ? 8 : 4;                 // warps that read one tile
    constexpr int COLS = ALL ? BN / 2 : BN;              // columns per warp and tile
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t slot;
    __shared__ uint64_t bars[2 * 8 + 8];                  // tfull[8], tempty[8], tma[8]
    __shared__ volatile int stop;
    unsigned char* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x3c003c00u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tfull = smem_u32(&bars[0]), tempty = smem_u32(&bars[8]), tbar = smem_u32(&bars[16]);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(1), "r"(tfull + 8 * i) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(READERS), "r"(tempty + 8 * i) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(1), "r"(tbar + 8 * i) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        stop = 0;
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (warp == 9 || warp == 10) {
        // ---- two MMA issuers on alternate tiles, like the real kernel
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)BN >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t ad0 = desc_sw128(smem_u32(base)), bd0 = desc_sw128(smem_u32(base) + 16384);
        const long long t0 = clock64();
        for (int i = warp - 9; i < n_tiles; i += 2) {
            const int acc = i % NACC, use = i / NACC;
            if (use > 0) wait(tempty + 8 * acc, (uint32_t)((use - 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d = tmem + (uint32_t)acc * BN;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(d), "l"(ad0 + 2 * j), "l"(bd0 + 2 * j), "r"(idesc), "r"(j ? 1u : 0u) : "memory");
            asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
                         ::"r"(tfull + 8 * acc) : "memory");
        }
        if (warp == 9) {
            // the last tiles have been read when their tempty phases complete
            for (int i = max(0, n_tiles - NACC); i < n_tiles; ++i) wait(tempty + 8 * (i % NACC), (uint32_t)((i / NACC) & 1));
            const long long t1 = clock64();
            if (lane == 0) { out[blockIdx.x] = t1 - t0; stop = 1; }
        }
    } else if (warp == 8) {
        if (with_tma && lane == 0) {
            uint32_t stage = 0, phase = 0, n = 0;
            while (!stop) {
                const uint32_t bar = tbar + 8 * stage;
                if (n >= 8) wait(bar, phase ^ 1);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(16384u), "r"(bar) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(base) + 32768u + stage * 16384u), "l"(src + (size_t)((n * 16384u) & 0xFFFFFFu)), "r"(16384u), "r"(bar) : "memory");
                ++n;
                if (++stage == 8) { stage = 0; phase ^= 1; }
            }
            for (uint32_t d = 0; d < 8 && d < n; ++d) {
                const uint32_t c = n - 1 - d;
                wait(tbar + 8 * (c & 7), (c >> 3) & 1);
            }
        }
    } else {
        // ---- readers: warp = (quarter q, half h)
        const int q = warp & 3, h = warp >> 2;
        uint32_t sink = 0;
        for (int i = ALL ? 0 : h; i < n_tiles; i += ALL ? 1 : 2) {
            const int acc = i % NACC, use = i / NACC;
            wait(tfull + 8 * acc, (uint32_t)(use & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tb = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)acc * BN + (ALL ? (uint32_t)h * COLS : 0u);
            uint32_t va[32], vb[32];
            if (COLS == 32) {
                ld32(tb, va);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int c = 0; c < 32; ++c) sink = __funnelshift_l(va[c], sink, 1);
            } else {
#pragma unroll
                for (int c0 = 0; c0 < COLS; c0 += 64) {
                    ld32(tb + c0, va);
                    ld32(tb + c0 + 32, vb);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int c = 0; c < 32; ++c) sink = __funnelshift_l(va[c], sink, 1);
#pragma unroll
                    for (int c = 0; c < 32; ++c) sink = __funnelshift_l(vb[c], sink, 1);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) arrive(tempty + 8 * acc);
        }
        if (sink == 0x12345678u) out[0] = 0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
template <int BN, int ALL>
void run(int tma, long long* out, const unsigned char* src) {
    const int smem = 32768 + 8 * 16384 + 2048, n_tiles = 8000;
    cudaFuncSetAttribute(k<BN, ALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaMemset(out, 0, 148 * 8);
    k<BN, ALL><<<148, NT, smem>>>(n_tiles, tma, src, out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("BN %3d ring %d read-out %s tma %d: %.1f cycles/tile = %.1f cycles per 128 items   %s\n", BN, 512 / BN,
           ALL ? "all-8-warps" : "alternate  ", tma, (double)h / n_tiles, (double)h / n_tiles * 128 / BN, cudaGetErrorString(e));
}
int main() {
    long long* out; cudaMalloc(&out, 148 * 8);
    unsigned char* src; cudaMalloc(&src, 32 << 20); cudaMemset(src, 0x3c, 32 << 20);
    for (int tma = 0; tma < 2; ++tma) {
        run<128, 0>(tma, out, src); run<128, 1>(tma, out, src);
        run<64, 0>(tma, out, src);  run<64, 1>(tma, out, src);
    }
    return 0;
}
