// Micro-benchmark: shared-memory wavefronts per LDS for the address patterns the prefilter can
// use.  Each warp issues ITER x 16 independent loads; reports SM cycles per warp-level LDS at
// full occupancy (16 warps/SM) -- i.e. wavefronts per instruction when LDS-bound.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int BYTES> struct Vec;
template <> struct Vec<4> { using T = uint32_t; };
template <> struct Vec<8> { using T = uint2; };
template <> struct Vec<16> { using T = uint4; };
__device__ __forceinline__ uint32_t fold(uint32_t v) { return v; }
__device__ __forceinline__ uint32_t fold(uint2 v) { return v.x ^ v.y; }
__device__ __forceinline__ uint32_t fold(uint4 v) { return v.x ^ v.y ^ v.z ^ v.w; }

// lane address = base + ((lane_hash % distinct) * stride_bytes); 16 different bases per iteration
template <int BYTES>
__global__ void __launch_bounds__(512, 1) k(uint32_t *out, long long *cycles, int distinct, int stride, int iters, uint32_t seed) {
    extern __shared__ __align__(16) unsigned char sm[];
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) ((uint32_t *) sm)[i] = i * 2654435761u;
    __syncthreads();
    uint32_t off[16];
    uint32_t h = (threadIdx.x * 2654435761u) ^ seed;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        h = h * 1664525u + 1013904223u;
        off[j] = ((h >> 16) % distinct) * stride;
    }
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        const unsigned char *base = sm + (it & 7) * 4096;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            typename Vec<BYTES>::T v = *reinterpret_cast<const typename Vec<BYTES>::T *>(base + j * 256 + off[j]);
            acc += fold(v);
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int BYTES> void run(const char *name, int distinct, int stride) {
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
    int iters = 2000;
    cudaFuncSetAttribute(k<BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    k<BYTES><<<148, 512, 48 * 1024>>>(out, cyc, distinct, stride, iters, 12345u);
    k<BYTES><<<148, 512, 48 * 1024>>>(out, cyc, distinct, stride, iters, 777u);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
    double per = avg / ((double) iters * 16 * 16);  // 16 warps x 16 loads per iteration
    printf("%-46s LDS.%-3d distinct=%3d stride=%3d  cycles/warp-LDS = %.3f\n", name, BYTES * 8, distinct, stride, per);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<4>("32b, 16 words in 64 B (current tables)", 16, 4);
    run<4>("32b, 32 words in 128 B", 32, 4);
    run<4>("32b, 64 words in 256 B (3-mer)", 64, 4);
    run<4>("32b, 256 words in 1 KB (4-mer)", 256, 4);
    run<8>("64b, 16 entries in 128 B (pair layout)", 16, 8);
    run<8>("64b, 32 entries in 256 B", 32, 8);
    run<8>("64b, 8 entries in 64 B", 8, 8);
    run<8>("64b, 64 entries in 512 B (3-mer pairs)", 64, 8);
    run<16>("128b, 8 entries in 128 B", 8, 16);
    run<16>("128b, 16 entries in 256 B (quad layout)", 16, 16);
    run<16>("128b, 32 entries in 512 B", 32, 16);
    run<16>("128b, 4 entries in 64 B", 4, 16);
    return 0;
}
