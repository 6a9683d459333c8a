// Micro-benchmark: tcgen05.ld (TMEM -> registers) throughput on B200, the epilogue rate a
// tensor-core prefilter would need (every accumulator must be read and thresholded).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define LD32(r, addr)                                                                               \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                          \
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"                          \
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"        \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),  \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), \
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), \
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) \
                 : "r"(addr))

template <int WARPS, int BATCH>
__global__ void __launch_bounds__(WARPS * 32, 1) k(uint32_t *out, long long *cycles, int iters) {
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        uint32_t dst = (uint32_t) __cvta_generic_to_shared(&tmem_base);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(dst));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = tmem_base + (((uint32_t) (warp & 3) * 32u) << 16);
    uint32_t acc = 0xffffffffu;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < 512; c += 32 * BATCH) {
            uint32_t r[BATCH][32];
#pragma unroll
            for (int b = 0; b < BATCH; b++) LD32(r[b], base + c + 32 * b);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int b = 0; b < BATCH; b++)
#pragma unroll
                for (int j = 0; j < 32; j += 3) acc &= r[b][j] & r[b][(j + 1) & 31] & r[b][(j + 2) & 31];
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}

template <int WARPS, int BATCH> void run() {
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, 148 * WARPS * 32 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    k<WARPS, BATCH><<<148, WARPS * 32>>>(out, cyc, iters);
    k<WARPS, BATCH><<<148, WARPS * 32>>>(out, cyc, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
    // bytes read per CTA: every warp reads its 32 lanes x 512 columns x 4 B per iteration
    double bytes = (double) iters * WARPS * 32 * 512 * 4;
    printf("warps=%d batch=%d: %s  %.1f B/clk/SM  (%.2f accumulators/clk/SM), %.1f clk per x32 ld per warp\n", WARPS, BATCH,
           cudaGetErrorString(e), bytes / avg, bytes / avg / 4, avg / (iters * 16.0));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<4, 1>(); run<4, 2>(); run<4, 4>();
    run<8, 1>(); run<8, 2>();
    run<16, 1>(); run<16, 2>();
    return 0;
}
