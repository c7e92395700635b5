// microbenchmark: tcgen05.mma issue throughput vs commit frequency (M128 N256 K16 bf16, SW128 K-major, smem zero)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
    return (uint64_t)((a & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__global__ void __launch_bounds__(128, 1) k(int n_mma, int commit_every, int n_acc, int wait_each, long long* cyc, int N, int alt) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    unsigned char* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0;
    int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(1), "r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem = slot;
    if (warp == 1) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
        uint32_t a0 = smem_u32(base), b0 = a0 + 16384;
        uint32_t parity = 0;
        long long t0 = clock64();
        const uint64_t ad0 = desc_sw128(a0), bd0 = desc_sw128(b0);
        const uint32_t bar_a = smem_u32(&bar);
        for (int i = 0; i < n_mma; i += 4) {
            const uint32_t acc = ((i >> 2) & (n_acc - 1)) * (512 / n_acc);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem + acc), "l"(ad0 + 2 * j), "l"(bd0 + 2 * j), "r"(idesc), "r"(j ? 1u : 0u) : "memory");
            }
            if (commit_every == 4)
                asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar_a) : "memory");
        }
        long long t1 = clock64();
        if ((threadIdx.x & 31) == 0) cyc[blockIdx.x] = t1 - t0;
    }
    __syncthreads();
    // crude: give the pipe time to finish before dealloc
    if (warp == 1) { long long t = clock64(); while (clock64() - t < 2000000) {} }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
int main() {
    long long* cyc; cudaMalloc(&cyc, 148 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    int n = 16384;
    int cfgs[][5] = {{4096, 2, 0, 256, 0}, {4096, 2, 0, 128, 0}, {4096, 2, 0, 64, 0}, {4, 2, 0, 256, 0}, {4, 2, 0, 128, 0}, {4096, 1, 0, 256, 0}, {4096, 2, 0, 32, 0}};
    for (auto& c : cfgs) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k<<<148, 128, 64 * 1024>>>(n, c[0], c[1], c[2], cyc, c[3], c[4]);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long c0; cudaMemcpy(&c0, cyc, 8, cudaMemcpyDeviceToHost);
        printf("N %3d alt %d commit_every %4d n_acc %d wait_each %d: issue-loop %.1f cycles/MMA (kernel %.3f ms incl. 1ms pad) err=%s\n", c[3], c[4], c[0], c[1], c[2], (double)c0 / n, ms, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
