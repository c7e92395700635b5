// microbenchmark: tcgen05.ld throughput of 8 epilogue warps WHILE one warp keeps the tensor pipe busy
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
    return (uint64_t)((a & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
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
__device__ __forceinline__ uint32_t signs(const uint32_t (&v)[32]) {
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s = __funnelshift_l(v[i], s, 1);
    return s;
}
// mode bit0: MMA warp active ; bit1: epilogue warps active ; nld = loads in flight per wait (1,2,4)
template <int NLD>
__global__ void __launch_bounds__(320, 1) k(int iters, int mode, int N, uint32_t* out, long long* cyc, int pattern) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t slot;
    unsigned char* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0;
    int warp = threadIdx.x >> 5;
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem = slot;
    long long t0 = clock64();
    if (warp == 9 && (mode & 1)) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t ad0 = desc_sw128(smem_u32(base)), bd0 = desc_sw128(smem_u32(base) + 16384);
        for (int i = 0; i < iters; ++i) {
            const uint32_t acc = (i & 1) * 256;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem + acc), "l"(ad0 + 2 * j), "l"(bd0 + 2 * j), "r"(idesc), "r"(j ? 1u : 0u) : "memory");
        }
        long long t1 = clock64();
        if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 2] = t1 - t0;
    } else if (warp < 8 && (mode & 2)) {
        uint32_t tb = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (warp >> 2) * 128;
        uint32_t acc = 0;
        for (int i = 0; i < iters; ++i) {
            uint32_t t = tb + (i & 1) * 256;
            if (pattern == 1) t = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (((i & 3) * 128 + (warp >> 2) * 128) & 511);
            if (pattern == 2 && warp >= 4) continue;
            if (pattern == 3) t = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (warp >> 2) * 128;
            if (pattern == 6) t = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (warp >> 2) * 256 + (i & 1) * 128;   // acc ring of N=256 tiles laid out as [h][acc]
            if (pattern == 7) t = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + ((i & 1) ? 256u : 0u) + (warp >> 2) * 128;
            if (pattern == 8) t = tb + ((i >> 1) & 1) * 256;   // switch accumulator every 2nd iteration
            if (NLD == 1) {
                uint32_t va[32], vb[32];
                ld32(t, va); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                ld32(t + 32, vb); acc += signs(va); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                ld32(t + 64, va); acc += signs(vb); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                ld32(t + 96, vb); acc += signs(va); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc += signs(vb);
            } else if (NLD == 2) {
                uint32_t va[32], vb[32];
                ld32(t, va); ld32(t + 32, vb); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc += signs(va); acc += signs(vb);
                ld32(t + 64, va); ld32(t + 96, vb); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc += signs(va); acc += signs(vb);
            } else {
                uint32_t va[32], vb[32], vc[32], vd[32];
                ld32(t, va); ld32(t + 32, vb); ld32(t + 64, vc); ld32(t + 96, vd); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc += signs(va); acc += signs(vb); acc += signs(vc); acc += signs(vd);
            }
        }
        long long t1 = clock64();
        if (threadIdx.x == 0) cyc[blockIdx.x * 2 + 1] = t1 - t0;
        out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    }
    __syncthreads();
    if (warp == 9) { long long t = clock64(); while (clock64() - t < 1000000) {} }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
template <int NLD> void run(int mode, int N, uint32_t* out, long long* cyc, int pattern = 0) {
    int iters = 4000;
    cudaMemset(cyc, 0, 148 * 16);
    cudaFuncSetAttribute(k<NLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    k<NLD><<<148, 320, 64 * 1024>>>(iters, mode, N, out, cyc, pattern);
    cudaDeviceSynchronize();
    long long c[2]; cudaMemcpy(c, cyc, 16, cudaMemcpyDeviceToHost);
    printf("pattern %d NLD %d mode %d N %3d: MMA %.0f cycles/tile(4 MMAs), epilogue %.0f cycles/tile (128 KB) err=%s\n", pattern, NLD, mode, N,
           (double)c[0] / iters, (double)c[1] / iters, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    uint32_t* out; long long* cyc; cudaMalloc(&out, 148 * 320 * 4); cudaMalloc(&cyc, 148 * 16);
    run<1>(2, 256, out, cyc, 0); run<1>(2, 256, out, cyc, 6); run<1>(2, 256, out, cyc, 7); run<1>(2, 256, out, cyc, 8); run<1>(3, 256, out, cyc, 6); run<4>(3, 256, out, cyc, 6);
    return 0;
}
