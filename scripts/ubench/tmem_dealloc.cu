// Microbenchmark: latency of tcgen05.alloc / tcgen05.dealloc for 32..512 columns, with and without prior tcgen05.ld traffic.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int COLS>
__global__ void k(long long* out, int touch) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    if (warp == 0) {
        t0 = clock64();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "n"(COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        t1 = clock64();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (touch) {
        uint32_t r[16]; uint32_t acc = 0;
        const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
        for (int i = 0; i < 64; ++i) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(base + (i * 16) % (COLS - 15 > 0 ? COLS - 15 : 1)) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 16; ++j) acc ^= r[j];
        }
        if (acc == 0x1234567u) out[100] = acc;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        t2 = clock64();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "n"(COLS) : "memory");
        t3 = clock64();
        if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t3 - t2; }
    }
}
template <int COLS> void run(int touch, int threads) {
    long long* d; cudaMalloc(&d, 8 * 128);
    k<COLS><<<148, threads>>>(d, touch); k<COLS><<<148, threads>>>(d, touch);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("cols %3d touch %d threads %d: alloc %lld clk, dealloc %lld clk (%s)\n", COLS, touch, threads, h[0], h[1], cudaGetErrorString(e));
    cudaFree(d);
}
int main() {
    for (int touch = 0; touch < 2; ++touch) { run<32>(touch, 128); run<128>(touch, 128); run<256>(touch, 128); run<512>(touch, 128); run<512>(touch, 448); }
    return 0;
}
