// microbenchmark: tcgen05.mma.cta_group::2 (M256 over a CTA pair) issue/execute rate vs N, next to cta_group::1 M128
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
    return (uint64_t)((a & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void csync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
template <int PAIR>
__global__ void __launch_bounds__(128, 1) k(int n_mma, long long* cyc, int N, int n_acc) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    unsigned char* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0;
    int warp = threadIdx.x >> 5;
    uint32_t crank = 0;
    if (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(1 << 20), "r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (PAIR) csync();
    uint32_t tmem = slot;
    if (warp == 1 && crank == 0) {
        const uint32_t M = PAIR ? 256u : 128u;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)N >> 3) << 17) | ((M >> 4) << 24);
        uint32_t a0 = smem_u32(base), b0 = a0 + 16384;
        const uint64_t ad0 = desc_sw128(a0), bd0 = desc_sw128(b0);
        const uint32_t bar_a = smem_u32(&bar);
        long long t0 = clock64();
        for (int i = 0; i < n_mma; i += 4) {
            const uint32_t acc = ((i >> 2) & (n_acc - 1)) * (512 / n_acc);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (PAIR)
                    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem + acc), "l"(ad0 + 2 * j), "l"(bd0 + 2 * j), "r"(idesc), "r"(j ? 1u : 0u) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem + acc), "l"(ad0 + 2 * j), "l"(bd0 + 2 * j), "r"(idesc), "r"(j ? 1u : 0u) : "memory");
            }
            if (PAIR)
                asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(bar_a), "h"((uint16_t)3) : "memory");
            else
                asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar_a) : "memory");
        }
        long long t1 = clock64();
        if ((threadIdx.x & 31) == 0) cyc[blockIdx.x] = t1 - t0;
    }
    __syncthreads();
    if (warp == 1) { long long t = clock64(); while (clock64() - t < 4000000) {} }   // let the pipe drain
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (PAIR) csync();
    if (warp == 0) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}
template <int PAIR>
void run(int N, int n_acc, long long* cyc) {
    int n = 16384;
    cudaFuncSetAttribute(k<PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 64 * 1024; cfg.stream = 0;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = PAIR ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaMemset(cyc, 0, 148 * 8);
    cudaError_t e = cudaLaunchKernelEx(&cfg, k<PAIR>, n, cyc, N, n_acc);
    cudaError_t e2 = cudaDeviceSynchronize();
    long long c0; cudaMemcpy(&c0, cyc, 8, cudaMemcpyDeviceToHost);
    printf("cta_group::%d M%d N%3d n_acc %d: %.1f cycles/MMA   launch=%s sync=%s\n", PAIR ? 2 : 1, PAIR ? 256 : 128, N, n_acc, (double)c0 / n,
           cudaGetErrorString(e), cudaGetErrorString(e2));
}
int main() {
    long long* cyc; cudaMalloc(&cyc, 148 * 8);
    run<0>(128, 4, cyc); run<0>(256, 2, cyc); run<0>(64, 4, cyc);
    run<1>(64, 4, cyc); run<1>(128, 4, cyc); run<1>(256, 2, cyc); run<1>(32, 4, cyc);
    return 0;
}
