// racecheck_repro.cu -- minimal correctly-synchronised cp.async.bulk pattern, for `compute-sanitizer --tool racecheck`.
// One CTA: thread 0 initialises an mbarrier, arms it with the byte count and issues TWO bulk copies into DISJOINT halves of
// a shared buffer; every thread waits on the barrier phase, then reads both halves.  This is exactly the staging pattern
// of the library kernels (mctq_common.cuh: mbar_init / mbar_arrive_expect_tx / bulk_g2s / mbar_wait) reduced to 40 lines.
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O2 -o tools/_build/racecheck_repro tools/racecheck_repro.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void repro(const float* __restrict__ src, float* __restrict__ dst, int copies) {
    __shared__ __align__(16) float buf[512];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(2048u) : "memory");
        const uint32_t part = 2048u / copies;
        for (int c = 0; c < copies; ++c)          // disjoint destination ranges
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(buf) + c * part),
                         "l"(reinterpret_cast<const char*>(src + blockIdx.x * 512) + c * part), "r"(part), "r"(s32(&bar)) : "memory");
    }
    __syncthreads();                              // barrier init visible to everyone
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(s32(&bar)), "r"(0u) : "memory");
    dst[blockIdx.x * 256 + threadIdx.x] = buf[threadIdx.x] + buf[256 + threadIdx.x];
}

int main() {
    float *src, *dst;
    cudaMalloc(&src, 64 * 512 * 4);
    cudaMalloc(&dst, 64 * 256 * 4);
    cudaMemset(src, 0, 64 * 512 * 4);
    for (int copies = 1; copies <= 2; ++copies) {
        repro<<<64, 256>>>(src, dst, copies);
        printf("copies per CTA = %d: %s\n", copies, cudaGetErrorString(cudaDeviceSynchronize()));
    }
    return 0;
}
