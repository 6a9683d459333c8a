// fp8_peak.cu -- dense e4m3 x e4m3 -> f16 tensor-core peak of this GPU, the denominator of bench.py's
// roofline for prefilter_tc_kernel (MEASURED_PEAKS.json only holds a bf16 figure).
//
// One CTA per SM.  A (128 x K) and B (256 x K) sit in shared memory in the no-swizzle K-major canonical
// layout the product uses, filled with random non-zero e4m3 bytes; one warp issues
// tcgen05.mma.cta_group::1.kind::f8f6f4, M = 128, N = 256, K = 32, four K steps per accumulator, into two
// alternating TMEM buffers.  Nothing reads the accumulators back: this is the tensor pipe alone, fed from
// shared memory exactly as in the prefilter (12 KB of operands per MMA).
//
//   burst      best of several ~25 ms launches (the GPU at its boost clock)
//   sustained  one ~1.5 s launch (clocks settled under the tensor load)
//
// Prints one JSON line: {"burst_tflops": .., "sustained_tflops": .., ...}.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o bench_micro/fp8_peak bench_micro/fp8_peak.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kKSteps = 4;                       // K = 32 steps per accumulator
constexpr int kABytes = 128 * 32 * kKSteps;      // 16 KB
constexpr int kBBytes = 256 * 32 * kKSteps;      // 32 KB
constexpr int kBatch = 16;                       // accumulators (of kKSteps MMAs) per commit

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_f8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
// K-major, no swizzle (version 1 descriptor): start address, K-chunk byte offset (LBO), 8-row-group byte offset (SBO)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t) ((addr & 0x3FFFFu) >> 4) | ((uint64_t) (lbo >> 4) << 16) | ((uint64_t) (sbo >> 4) << 32) | (1ull << 46);
}

__global__ void __launch_bounds__(128, 1) fp8_peak_kernel(const uint8_t *__restrict__ fill, long long batches) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[2];
    __shared__ uint32_t s_tmem;
    uint8_t *s_a = smem, *s_b = smem + kABytes;
    for (int i = threadIdx.x; i < (kABytes + kBBytes) / 16; i += blockDim.x)
        reinterpret_cast<uint4 *>(smem)[i] = reinterpret_cast<const uint4 *>(fill)[i];
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    if (threadIdx.x == 0) {
        // D = F16, A = B = E4M3, K-major, N = 256, M = 128 (the prefilter's instruction descriptor)
        const uint32_t idesc = ((uint32_t) (256 >> 3) << 17) | ((uint32_t) (128 >> 4) << 24);
        // per K step: A 128 rows x 32 B = two 16-byte chunks of 128 x 16 B (LBO = 2048), B likewise (LBO = 4096)
        for (long long it = 0; it < batches; it++) {
            const int half = (int) (it & 1);
            if (it >= 2) mbar_wait(&bar[half], (uint32_t) (((it >> 1) - 1) & 1));   // the batch two back has drained
            for (int u = 0; u < kBatch; u++) {
                const uint32_t d = tmem + (uint32_t) ((u & 1) * 256);
#pragma unroll
                for (int k = 0; k < kKSteps; k++)
                    umma_f8(d, make_desc(smem_u32(s_a) + k * 4096, 2048, 128), make_desc(smem_u32(s_b) + k * 8192, 4096, 128), idesc, k > 0);
            }
            umma_commit(&bar[half]);
        }
        const long long last = batches - 1;
        if (batches >= 1) mbar_wait(&bar[last & 1], (uint32_t) ((last >> 1) & 1));
        if (batches >= 2) mbar_wait(&bar[(last - 1) & 1], (uint32_t) (((last - 1) >> 1) & 1));
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    if (prop.major != 10) { printf("{\"error\": \"not an sm_100 device\"}\n"); return 0; }
    const int n_sm = prop.multiProcessorCount;
    const int smem = kABytes + kBBytes;
    CK(cudaFuncSetAttribute(fp8_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    std::vector<uint8_t> h(smem);
    uint32_t x = 12345u;
    for (auto &b : h) {               // random e4m3 magnitudes in [2^-3, 2^2): dense, finite, no overflow in f16
        x = x * 1664525u + 1013904223u;
        b = (uint8_t) ((((x >> 24) & 0x80u)) | (0x20u + ((x >> 16) % 0x28u)));
    }
    uint8_t *d_fill;
    CK(cudaMalloc(&d_fill, smem));
    CK(cudaMemcpy(d_fill, h.data(), smem, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const double flop_per_batch = (double) kBatch * kKSteps * 2.0 * 128 * 256 * 32;
    auto run = [&](long long batches) {
        CK(cudaEventRecord(e0));
        fp8_peak_kernel<<<n_sm, 128, smem>>>(d_fill, batches);
        CK(cudaGetLastError());
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        return flop_per_batch * batches * n_sm / (ms * 1e-3) / 1e12;
    };
    run(200);                                            // warm-up
    const double probe = run(2000);                      // TFLOP/s -> size the launches
    const long long per_ms = (long long) (probe * 1e12 * 1e-3 / (flop_per_batch * n_sm)) + 1;
    double burst = 0;
    for (int i = 0; i < 5; i++) { double t = run(25 * per_ms); if (t > burst) burst = t; }
    const double sustained = run(1500 * per_ms);
    printf("{\"burst_tflops\": %.1f, \"sustained_tflops\": %.1f, \"sms\": %d, \"mma\": \"tcgen05.mma.cta_group::1.kind::f8f6f4 "
           "M=128 N=256 K=32, e4m3 x e4m3 -> f16, operands from shared memory, no epilogue\", "
           "\"burst_ms\": 25, \"sustained_ms\": 1500}\n", burst, sustained, n_sm);
    return 0;
}
