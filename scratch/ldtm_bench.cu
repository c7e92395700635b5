// microbenchmark: TMEM read (tcgen05.ld) throughput per SM for several shapes / warp counts
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int X>
__device__ __forceinline__ uint32_t ld_sink(uint32_t taddr);
template <>
__device__ __forceinline__ uint32_t ld_sink<32>(uint32_t taddr) {
    uint32_t v[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s = __funnelshift_l(v[i], s, 1);
    return s;
}
template <>
__device__ __forceinline__ uint32_t ld_sink<8>(uint32_t taddr) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s = __funnelshift_l(v[i], s, 1);
    return s;
}
template <int X, bool NOWAIT>
__global__ void k(int iters, uint32_t* out, long long* cyc) {
    __shared__ uint32_t slot;
    int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t base = slot + ((uint32_t)(32 * (warp & 3)) << 16);
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
        for (int c = 0; c < 512; c += X) acc += ld_sink<X>(base + ((c + (warp >> 2) * 128) & 511));
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}
int main() {
    uint32_t* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    int iters = 2000;
    for (int warps : {4, 8, 16}) {
        for (int x : {32, 8}) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            if (x == 32) k<32, false><<<148, warps * 32>>>(iters, out, cyc); else k<8, false><<<148, warps * 32>>>(iters, out, cyc);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long c0; cudaMemcpy(&c0, cyc, 8, cudaMemcpyDeviceToHost);
            double bytes_per_sm = (double)iters * (512.0 / x) * warps * 32 * x * 4;
            printf("warps %2d shape x%-2d: %.3f ms, %lld cycles, %.1f B/clk/SM, err=%s\n", warps, x, ms, c0, bytes_per_sm / c0, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
