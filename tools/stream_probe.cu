// stream_probe.cu -- what is the ceiling for a 1:1 read/write streaming kernel on this B200?
// Standalone experiment (not part of the library): y[i] = f(x[i]) over f32 with different load/store cache hints, tile
// shapes and grid styles; prints achieved GB/s (read + write) for each variant.  Results are summarised in
// profiles/r01_stream_probe.txt and DESIGN.md; the library kernels use the best variant.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/stream_probe tools/stream_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

enum { LD_NC_NOALLOC = 0, LD_CS = 1, LD_NC_L2_256 = 2, LD_PLAIN = 3, LD_EVICT_FIRST = 4 };
enum { ST_NOALLOC = 0, ST_CS = 1, ST_PLAIN = 2, ST_WT = 3, ST_EVICT_FIRST = 4 };

template <int LD> __device__ __forceinline__ uint4 ld(const uint4* p, uint64_t pol) {
    uint4 r;
    if (LD == LD_NC_NOALLOC) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if (LD == LD_CS) asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if (LD == LD_NC_L2_256) asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if (LD == LD_EVICT_FIRST) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
    else r = *p;
    return r;
}
template <int ST> __device__ __forceinline__ void st(uint4* p, uint4 v, uint64_t pol) {
    if (ST == ST_NOALLOC) asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else if (ST == ST_CS) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else if (ST == ST_WT) asm volatile("st.global.wt.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else if (ST == ST_EVICT_FIRST) asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
    else *p = v;
}
__device__ __forceinline__ uint4 work(uint4 v) {      // the fake-quant arithmetic, roughly
    float f[4] = {__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w)};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float t = f[i] * 46.5f;
        t = (t + 12582912.0f) - 12582912.0f;
        t = fminf(fmaxf(t, -116.0f), 139.0f);
        f[i] = t * 0.0215f;
    }
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
}

// one tile per CTA (the library's shape)
template <int LD, int ST, int UNROLL, int THREADS>
__global__ void __launch_bounds__(THREADS) tile_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, size_t nvec) {
    uint64_t pol = 0;
    if (LD == LD_EVICT_FIRST || ST == ST_EVICT_FIRST) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    const size_t base = (size_t)blockIdx.x * (THREADS * UNROLL) + threadIdx.x;
    uint4 v[UNROLL];
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) if (base + (size_t)j * THREADS < nvec) v[j] = ld<LD>(x + base + (size_t)j * THREADS, pol);
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) if (base + (size_t)j * THREADS < nvec) st<ST>(y + base + (size_t)j * THREADS, work(v[j]), pol);
}

// persistent: grid = SMs * k, each CTA walks tiles with a register double buffer (next tile's loads in flight while
// the current one is processed)
template <int LD, int ST, int UNROLL, int THREADS>
__global__ void __launch_bounds__(THREADS) persistent_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, size_t nvec) {
    uint64_t pol = 0;
    if (LD == LD_EVICT_FIRST || ST == ST_EVICT_FIRST) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    const size_t tile = (size_t)THREADS * UNROLL;
    const size_t ntiles = (nvec + tile - 1) / tile;
    uint4 cur[UNROLL], nxt[UNROLL];
    size_t t = blockIdx.x;
    if (t < ntiles) {
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) { size_t i = t * tile + (size_t)j * THREADS + threadIdx.x; if (i < nvec) cur[j] = ld<LD>(x + i, pol); }
    }
    for (; t < ntiles; t += gridDim.x) {
        const size_t tn = t + gridDim.x;
        if (tn < ntiles) {
#pragma unroll
            for (int j = 0; j < UNROLL; ++j) { size_t i = tn * tile + (size_t)j * THREADS + threadIdx.x; if (i < nvec) nxt[j] = ld<LD>(x + i, pol); }
        }
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) { size_t i = t * tile + (size_t)j * THREADS + threadIdx.x; if (i < nvec) st<ST>(y + i, work(cur[j]), pol); }
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) cur[j] = nxt[j];
    }
}

struct Result { const char* name; double gbs1, gbs4; };

template <class F> double time_it(F launch, size_t bytes, int reps) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) launch(i);
    CK(cudaDeviceSynchronize());
    std::vector<float> ts;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(a));
        launch(i);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        ts.push_back(ms);
    }
    std::sort(ts.begin(), ts.end());
    return bytes / (ts[ts.size() / 2] * 1e-3) / 1e9;
}

int main() {
    const size_t n1 = (size_t)256 << 20, n4 = (size_t)1024 << 20;      // f32 elements: 1 GiB and 4 GiB of input
    const int NB = 3;
    float* x[NB]; float* y[NB];
    for (int i = 0; i < NB; ++i) { CK(cudaMalloc(&x[i], n4 * 4)); CK(cudaMalloc(&y[i], n4 * 4)); CK(cudaMemset(x[i], 0x3c, n4 * 4)); }
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    printf("SMs %d\n", sms);
    auto report = [&](const char* name, auto launch1, auto launch4) {
        double g1 = time_it(launch1, 2 * n1 * 4, 20), g4 = time_it(launch4, 2 * n4 * 4, 10);
        printf("%-64s 1 GiB %8.1f GB/s   4 GiB %8.1f GB/s\n", name, g1, g4);
        fflush(stdout);
    };
    {
        auto l1 = [&](int i) { CK(cudaMemcpyAsync(y[i % NB], x[i % NB], n1 * 4, cudaMemcpyDeviceToDevice)); };
        auto l4 = [&](int i) { CK(cudaMemcpyAsync(y[i % NB], x[i % NB], n4 * 4, cudaMemcpyDeviceToDevice)); };
        report("cudaMemcpyAsync D2D", l1, l4);
    }
#define TILE(LDV, STV, U, T, NAME) { \
        auto l1 = [&](int i) { size_t nv = n1 / 4; tile_kernel<LDV, STV, U, T><<<(unsigned)((nv + (size_t)U * T - 1) / ((size_t)U * T)), T>>>((const uint4*)x[i % NB], (uint4*)y[i % NB], nv); }; \
        auto l4 = [&](int i) { size_t nv = n4 / 4; tile_kernel<LDV, STV, U, T><<<(unsigned)((nv + (size_t)U * T - 1) / ((size_t)U * T)), T>>>((const uint4*)x[i % NB], (uint4*)y[i % NB], nv); }; \
        report(NAME, l1, l4); }
#define PERS(LDV, STV, U, T, K, NAME) { \
        auto l1 = [&](int i) { persistent_kernel<LDV, STV, U, T><<<sms * K, T>>>((const uint4*)x[i % NB], (uint4*)y[i % NB], n1 / 4); }; \
        auto l4 = [&](int i) { persistent_kernel<LDV, STV, U, T><<<sms * K, T>>>((const uint4*)x[i % NB], (uint4*)y[i % NB], n4 / 4); }; \
        report(NAME, l1, l4); }
    TILE(LD_NC_NOALLOC, ST_NOALLOC, 4, 256, "tile u4 t256  ld.nc.L1::no_allocate / st.L1::no_allocate (library)")
    TILE(LD_CS, ST_CS, 4, 256, "tile u4 t256  ld.cs / st.cs")
    TILE(LD_NC_NOALLOC, ST_CS, 4, 256, "tile u4 t256  ld.nc.no_allocate / st.cs")
    TILE(LD_NC_L2_256, ST_NOALLOC, 4, 256, "tile u4 t256  ld.nc.no_allocate.L2::256B / st.no_allocate")
    TILE(LD_EVICT_FIRST, ST_NOALLOC, 4, 256, "tile u4 t256  ld L2::evict_first / st.no_allocate")
    TILE(LD_EVICT_FIRST, ST_EVICT_FIRST, 4, 256, "tile u4 t256  ld L2::evict_first / st L2::evict_first")
    TILE(LD_NC_NOALLOC, ST_EVICT_FIRST, 4, 256, "tile u4 t256  ld.nc.no_allocate / st L2::evict_first")
    TILE(LD_NC_NOALLOC, ST_WT, 4, 256, "tile u4 t256  ld.nc.no_allocate / st.wt")
    TILE(LD_PLAIN, ST_PLAIN, 4, 256, "tile u4 t256  plain ld / st")
    TILE(LD_NC_NOALLOC, ST_NOALLOC, 2, 256, "tile u2 t256")
    TILE(LD_NC_NOALLOC, ST_NOALLOC, 8, 256, "tile u8 t256")
    TILE(LD_NC_NOALLOC, ST_NOALLOC, 4, 512, "tile u4 t512")
    TILE(LD_NC_NOALLOC, ST_NOALLOC, 8, 128, "tile u8 t128")
    TILE(LD_NC_NOALLOC, ST_NOALLOC, 4, 1024, "tile u4 t1024")
    PERS(LD_NC_NOALLOC, ST_NOALLOC, 4, 256, 4, "persistent 4 CTA/SM u4 t256 (register double buffer)")
    PERS(LD_NC_NOALLOC, ST_NOALLOC, 4, 256, 8, "persistent 8 CTA/SM u4 t256")
    PERS(LD_NC_NOALLOC, ST_NOALLOC, 8, 256, 2, "persistent 2 CTA/SM u8 t256")
    PERS(LD_NC_NOALLOC, ST_NOALLOC, 4, 512, 2, "persistent 2 CTA/SM u4 t512")
    PERS(LD_NC_NOALLOC, ST_NOALLOC, 2, 1024, 1, "persistent 1 CTA/SM u2 t1024")
    PERS(LD_CS, ST_CS, 4, 256, 4, "persistent 4 CTA/SM u4 t256 ld.cs / st.cs")
    PERS(LD_EVICT_FIRST, ST_EVICT_FIRST, 4, 256, 4, "persistent 4 CTA/SM u4 t256 evict_first both")
    return 0;
}
