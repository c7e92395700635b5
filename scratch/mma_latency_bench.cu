// Round-2 planning microbenchmark (not run yet when written -- GPU budget of round 1 was spent):
// what bounds the accumulator round trip of the fused scoring kernel?
//   depth sweep : issue tile i only after tile i-depth has completed (commit -> mbarrier observed by the issuing warp).
//                 period(depth) = max(E, L / depth)  =>  L (completion latency of an idle pipe) and E (service time)
//   + tma       : a producer warp streams 16 KB tiles global -> shared (cp.async.bulk) into a ring next to the operands
//   + ldtm      : 8 warps keep reading the accumulator columns with tcgen05.ld.32x32b.x32
// One tile = 4 x (M128 N128 K16) bf16 SS MMAs into one of 4 accumulators, operands static in shared memory.
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

constexpr int NT = 352;      // 8 reader warps + producer + 2 spare, like the real kernel
__global__ void __launch_bounds__(NT, 1) k(int n_tiles, int depth, int with_tma, int with_ldtm, const unsigned char* src,
                                           long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t slot;
    __shared__ uint64_t bars[4 + 8];          // 4 accumulator barriers, 8 TMA stage barriers
    __shared__ volatile int stop;
    unsigned char* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x3c003c00u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 12; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(1), "r"(smem_u32(&bars[i])) : "memory");
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
    if (warp == 9) {
        // ---- MMA issuer: tile i goes to accumulator i % 4; before issuing tile i wait for tile i - depth
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t ad0 = desc_sw128(smem_u32(base)), bd0 = desc_sw128(smem_u32(base) + 16384);
        long long t0 = clock64(), first_lat = 0;
        for (int i = 0; i < n_tiles; ++i) {
            if (i >= depth) { const int j = i - depth; wait(smem_u32(&bars[j & 3]), (uint32_t)((j >> 2) & 1)); }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t acc = tmem + (uint32_t)(i & 3) * 128u;
            const long long ti = clock64();
#pragma unroll
            for (int j = 0; j < 4; ++j)
                asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(acc), "l"(ad0 + 2 * j), "l"(bd0 + 2 * j), "r"(idesc), "r"(j ? 1u : 0u) : "memory");
            asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
                         ::"r"(smem_u32(&bars[i & 3])) : "memory");
            if (i == 0 && depth == 1) { wait(smem_u32(&bars[0]), 0); first_lat = clock64() - ti; }
        }
        for (int j = max(0, n_tiles - depth); j < n_tiles; ++j) wait(smem_u32(&bars[j & 3]), (uint32_t)((j >> 2) & 1));
        const long long t1 = clock64();
        if (lane == 0) { out[2 * blockIdx.x] = t1 - t0; out[2 * blockIdx.x + 1] = first_lat; stop = 1; }
    } else if (warp == 8 && with_tma) {
        // ---- producer: 16 KB bulk copies into an 8-stage ring behind the operands until the issuer is done
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, n = 0;
            while (!stop) {
                const uint32_t bar = smem_u32(&bars[4 + stage]);
                if (n >= 8) wait(bar, phase ^ 1);           // the copy issued 8 stages ago has landed
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(16384u), "r"(bar) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(base) + 32768u + stage * 16384u), "l"(src + (size_t)((n * 16384u) & 0xFFFFFFu)), "r"(16384u), "r"(bar) : "memory");
                ++n;
                if (++stage == 8) { stage = 0; phase ^= 1; }
            }
            // drain: every issued copy must land before the CTA exits
            for (uint32_t d = 0; d < 8 && d < n; ++d) {
                const uint32_t s2 = (stage + 8 - 1 - d) & 7;
                const uint32_t uses = (n - 1 - d) / 8;       // completed phases of that stage so far
                wait(smem_u32(&bars[4 + s2]), uses & 1);
            }
        }
    } else if (warp < 8 && with_ldtm) {
        // ---- readers: hammer the accumulator columns (values are ignored)
        const uint32_t tb = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
        uint32_t sink = 0;
        while (!stop) {
            uint32_t v[32];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                         "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                         "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                           "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                           "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(tb + (uint32_t)((warp >> 2) * 64 + (sink & 1) * 32)) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 32; ++i) sink += v[i] >> 31;
        }
        if (sink == 0xFFFFFFFFu) out[0] = 0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
int main() {
    long long* out; cudaMalloc(&out, 148 * 16);
    unsigned char* src; cudaMalloc(&src, 32 << 20); cudaMemset(src, 0x3c, 32 << 20);
    const int smem = 32768 + 8 * 16384 + 2048;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int n_tiles = 4000;
    for (int tma = 0; tma < 2; ++tma)
        for (int ld = 0; ld < 2; ++ld)
            for (int depth = 1; depth <= 4; ++depth) {
                cudaMemset(out, 0, 148 * 16);
                k<<<148, NT, smem>>>(n_tiles, depth, tma, ld, src, out);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
                printf("tma %d ldtm %d depth %d: %.1f cycles/tile", tma, ld, depth, (double)h[0] / n_tiles);
                if (depth == 1) printf("   (first tile issue -> barrier observed: %lld cycles)", h[1]);
                printf("   %s\n", cudaGetErrorString(e));
            }
    return 0;
}
