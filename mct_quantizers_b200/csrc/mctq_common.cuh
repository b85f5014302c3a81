// mctq_common.cuh -- shared device utilities of libmctq_sm100 (see mctq_affine.cu / mctq_lut.cu / mctq_host.cu).
// Build: mct_quantizers_b200/build.py
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false ... (one object per .cu, linked into one .so)
// (--fmad=false: the reference never contracts a multiply-add; the fused operations these kernels need are
//  written explicitly with __fmaf_rn, which the flag does not touch.)
//
// All kernels are HBM-bound streaming kernels (no tensor cores): one CTA = one tile of kThreads * UNROLL vectors,
// every load of the tile is issued before the first use, per-channel parameters of the rows the tile touches are
// staged in shared memory while the loads are in flight.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <utility>

#include "mctq.h"

namespace mctq {

constexpr int kThreads = 256;

// process-wide state (defined in mctq_host.cu)
extern std::atomic<int64_t> g_launches;
extern int g_unroll;
extern int g_force_rint;
extern int g_force_ieee_div;
extern int g_lut_shfl;      // warp-shuffle search in the generic LUT kernel for tables of <= 32 entries (key 4)
extern int g_wide;          // wide (8-element / 256-bit) vector variants (mctq_set_tuning key 5)
extern int g_multi_span;    // tiles per CTA of the multi-tensor LUT launch (mctq_set_tuning key 6)
extern int g_lut_xy;        // xy-record variant of the prepared LUT kernel where it applies (mctq_set_tuning key 7)
extern int g_pdl;           // programmatic dependent launch for the streaming kernels (mctq_set_tuning key 3)
extern int g_tab_early;     // parameter tables staged before the dependent-launch wait when legal (mctq_set_tuning key 8)
extern int g_chain_max;     // launches that may overlap under the "free" order before one waits again (mctq_set_tuning key 10)
extern int g_nvtx;          // NVTX ranges around the C-ABI compute entry points (mctq_set_tuning key 9 / MCTQ_TUNE=9=1; off by default)

enum ChMode { CH_PT = 0, CH_VEC = 1, CH_ELEM = 2, CH_LAST = 3 };

// ------------------------------------------------------------------------------------------ small utils
struct FastDiv {   // floor(u / d) for 0 <= u < 2^31, 1 <= d < 2^31  (round-up multiplier, Granlund-Montgomery)
    uint32_t mul, shift, d;
};
inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    f.d = d;
    uint32_t s = 0;
    while (s < 32 && (1ull << s) < d) ++s;
    f.shift = s;
    f.mul = (uint32_t)((((1ull << 32) * ((1ull << s) - d)) / d) + 1);
    return f;
}
__device__ __forceinline__ uint32_t fdiv_u32(uint32_t u, const FastDiv& f) {
    return (__umulhi(u, f.mul) + u) >> f.shift;
}

template <int BYTES> struct Vec;
template <> struct Vec<16> { using type = uint4; };
template <> struct Vec<8> { using type = uint2; };
template <> struct Vec<4> { using type = uint32_t; };
template <> struct Vec<2> { using type = uint16_t; };
template <> struct Vec<1> { using type = uint8_t; };

// streaming loads: read-only path, do not allocate in L1 (every byte is touched exactly once)
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ld_stream(const uint2* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(uint4* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream(uint2* p, const uint2& v) {
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void st_stream(uint32_t* p, const uint32_t& v) {
    asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_stream(uint16_t* p, const uint16_t& v) {
    asm volatile("st.global.L1::no_allocate.u16 [%0], %1;" :: "l"(p), "h"(v) : "memory");
}

// 256-bit streaming store (sm_100: STG.256): one instruction covers 32 contiguous bytes per lane, so a warp writes 1 KB of
// contiguous memory -- two 128-bit stores per lane at a 32-byte lane stride would half-fill every sector twice
__device__ __forceinline__ void st_stream256(void* p, const uint32_t* w) {
    asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}
__device__ __forceinline__ void ld_stream256(const void* p, uint32_t* w) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
}

// element <-> f32
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

// unpack a register-resident vector of V elements of T (V * sizeof(T) bytes, as 32-bit words) to floats
template <typename T, int V> struct Pack;
template <int V> struct Pack<float, V> {
    static constexpr int kWords = V;
    __device__ static __forceinline__ void unpack(const uint32_t* w, float* f) {
#pragma unroll
        for (int i = 0; i < V; ++i) f[i] = __uint_as_float(w[i]);
    }
    __device__ static __forceinline__ void pack(const float* f, uint32_t* w) {
#pragma unroll
        for (int i = 0; i < V; ++i) w[i] = __float_as_uint(f[i]);
    }
};
template <int V> struct Pack<__nv_bfloat16, V> {
    static constexpr int kWords = V / 2;
    __device__ static __forceinline__ void unpack(const uint32_t* w, float* f) {
#pragma unroll
        for (int i = 0; i < V / 2; ++i) {
            f[2 * i] = __uint_as_float(w[i] << 16);
            f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    __device__ static __forceinline__ void pack(const float* f, uint32_t* w) {
#pragma unroll
        for (int i = 0; i < V / 2; ++i) {
            __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
    }
};
template <int V> struct Pack<__half, V> {
    static constexpr int kWords = V / 2;
    __device__ static __forceinline__ void unpack(const uint32_t* w, float* f) {
#pragma unroll
        for (int i = 0; i < V / 2; ++i) {
            __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
            float2 v = __half22float2(h);
            f[2 * i] = v.x;
            f[2 * i + 1] = v.y;
        }
    }
    __device__ static __forceinline__ void pack(const float* f, uint32_t* w) {
#pragma unroll
        for (int i = 0; i < V / 2; ++i) {
            __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
    }
};

template <int WORDS> __device__ __forceinline__ void ld_words(const void* p, uint32_t* w);
template <> __device__ __forceinline__ void ld_words<4>(const void* p, uint32_t* w) {
    uint4 v = ld_stream(reinterpret_cast<const uint4*>(p));
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
}
template <> __device__ __forceinline__ void ld_words<2>(const void* p, uint32_t* w) {
    uint2 v = ld_stream(reinterpret_cast<const uint2*>(p));
    w[0] = v.x; w[1] = v.y;
}
template <int WORDS> __device__ __forceinline__ void st_words(void* p, const uint32_t* w);
template <> __device__ __forceinline__ void st_words<4>(void* p, const uint32_t* w) {
    st_stream(reinterpret_cast<uint4*>(p), make_uint4(w[0], w[1], w[2], w[3]));
}
template <> __device__ __forceinline__ void st_words<2>(void* p, const uint32_t* w) {
    st_stream(reinterpret_cast<uint2*>(p), make_uint2(w[0], w[1]));
}

// integer codes of one vector -> global.  INT8: V bytes; INT4: V/2 bytes (element 2j in the low nibble)
template <int V, int CODE> __device__ __forceinline__ void st_codes(void* codes, int64_t elem, const int* c) {
    if (CODE == MCTQ_CODES_INT8) {
        uint32_t w[V / 4];
#pragma unroll
        for (int i = 0; i < V / 4; ++i)
            w[i] = (uint32_t)(c[4 * i] & 0xff) | ((uint32_t)(c[4 * i + 1] & 0xff) << 8) |
                   ((uint32_t)(c[4 * i + 2] & 0xff) << 16) | ((uint32_t)(c[4 * i + 3] & 0xff) << 24);
        uint8_t* p = reinterpret_cast<uint8_t*>(codes) + elem;
        if (V == 4) st_stream(reinterpret_cast<uint32_t*>(p), w[0]);
        else st_stream(reinterpret_cast<uint2*>(p), make_uint2(w[0], w[V / 4 - 1]));
    } else if (CODE == MCTQ_CODES_INT4) {
        uint32_t w = 0;
#pragma unroll
        for (int i = 0; i < V; ++i) w |= (uint32_t)(c[i] & 0xf) << (4 * i);
        uint8_t* p = reinterpret_cast<uint8_t*>(codes) + (elem >> 1);
        if (V == 4) st_stream(reinterpret_cast<uint16_t*>(p), (uint16_t)w);
        else st_stream(reinterpret_cast<uint32_t*>(p), w);
    }
}
// ------------------------------------------------------------------------------------------ channel window
// A tile covers logical elements [g0, g0 + tile) of the [outer][C][inner] view.  It touches rows
// r0 .. r0 + nr - 1 (row = g / inner, channel = row % C).  Slot jj of the window holds channel (r0 + jj) % C,
// for jj < W = min(max rows per tile, C); row j of the tile maps to slot j % W.
struct Window {
    uint32_t off0;     // offset of g0 inside its row (bigrow: unused)
    uint32_t split;    // bigrow: first local element of the second row (or > tile if none)
};

template <class Op>
__device__ __forceinline__ void stage_window(float* sm_par, Window* sm_win, int64_t g0, uint32_t tile_elems,
                                             const typename Op::Args& a) {
    const uint32_t tid = threadIdx.x;
    if (tid < a.W || tid == 0) {
        int64_t r0 = g0 / a.inner;
        int64_t off = g0 - r0 * a.inner;
        if (tid == 0) {
            Window w;
            w.off0 = a.bigrow ? 0u : (uint32_t)off;
            int64_t sp = a.inner - off;
            w.split = (uint32_t)(sp > (int64_t)tile_elems ? (int64_t)tile_elems + 1 : sp);
            *sm_win = w;
        }
        int64_t c0 = r0 % a.C;
        for (uint32_t jj = tid; jj < a.W; jj += kThreads) {
            int64_t c = c0 + jj;
            c = c % a.C;
            Op::stage(sm_par, a.W, jj, c, a);
        }
    }
}

// row (slot) of local element l
template <class Args>
__device__ __forceinline__ void locate(uint32_t l, const Window& w, const Args& a, uint32_t& slot, uint32_t& rem) {
    uint32_t j;
    if (a.bigrow) {
        j = l >= w.split ? 1u : 0u;
        rem = 0;   // unused in bigrow mode
    } else {
        uint32_t u = w.off0 + l;
        j = fdiv_u32(u, a.div_inner);
        rem = u - j * a.div_inner.d;
    }
    slot = j - fdiv_u32(j, a.div_W) * a.div_W.d;
}


// ------------------------------------------------------------------------------------------ dependent launch
// Every streaming kernel is launched with the programmatic-stream-serialization attribute and brackets its work with
// griddepcontrol.wait (returns once the preceding kernel in the stream has completed and its writes are visible; a no-op
// when the launch carries no programmatic dependency) and griddepcontrol.launch_dependents, which lets the NEXT kernel of
// the stream be scheduled while this one is still running: its CTAs take the SM slots that free up during our tail.
//
// Three orders exist, selected per launch by the host (`early` in the kernel arguments, see pdl_plan_launch):
//   0 late  : wait -> trigger -> loads -> math -> stores         the default; legal whatever else runs on the stream
//   1 early : loads -> wait -> trigger -> math -> stores         the dependent CTAs already have their tile in flight while
//             the predecessor's last wave drains.  Needs: the input is not an output of any launch still in flight.
//   2 free  : trigger -> loads -> math -> stores -> wait         the launch does not wait for anything until its CTAs are
//             done: consecutive launches overlap completely, there is no drain / ramp between them.  The wait at the end
//             keeps the stream's completion order (a kernel finishes only after its predecessor has).  Needs: inputs AND
//             outputs are disjoint from the inputs and outputs of every launch still in flight; at most g_chain_max - 1
//             such launches in a row, then a late one.
// Orders 1 and 2 are OPT-IN (mctq_set_tuning key 3 = 2 / 3; Python: `with mct_quantizers_b200.private_stream():`): the library
// cannot see kernels other libraries enqueue between two of its launches -- such a kernel may itself release its dependents
// before its stores are visible (cuDNN / CUTLASS kernels do), and the caching allocator may hand one of its input buffers
// to a later launch of ours as output -- so they are only legal while the stream carries this library's launches alone.
// Measured on the MobileNetV2 step (54 back-to-back launches): late 6656, early 6761 GB/s; an L2-prefetch-before-the-wait
// variant that would be legal unconditionally gained nothing (6647) and was dropped.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter(uint32_t order) {          // first statement of a streaming kernel
    if (order == 0) pdl_wait();
    if (order != 1) pdl_launch_dependents();
}
__device__ __forceinline__ void pdl_loaded(uint32_t order) {         // the tile's loads are issued; nothing has been written yet
    if (order == 1) { pdl_wait(); pdl_launch_dependents(); }
}
__device__ __forceinline__ void pdl_exit(uint32_t order) {           // last statement, every thread
    if (order == 2) pdl_wait();
}

// Host side (mctq_host.cu; consulted when the caller opted in): the library remembers, per (device, stream), the memory
// ranges of its launches that may still be in flight -- back to the last launch that waited before touching memory -- and
// picks the most permissive order the new launch's ranges allow.  Launches whose ranges are not described (multi-tensor
// plans) record "unknown", which forces the late order on their successor.
struct IoSpan { const void* p; size_t bytes; };
// `tables` (optional): the prepared blob the kernel stages into shared memory; *tab_early = 1 when it may do so before its wait
int pdl_plan_launch(cudaStream_t st, const IoSpan* in, int n_in, const IoSpan* out, int n_out,    // 0 late, 1 early, 2 free
                    const IoSpan* tables = nullptr, uint32_t* tab_early = nullptr);
// a prepare kernel (plain launch: waits for everything before it) that writes `blob`
void pdl_note_prepare(cudaStream_t st, const void* blob, size_t bytes);
void pdl_forget_streams();                                           // every stream starts over with a late launch

// ------------------------------------------------------------------------------------------ TMA bulk staging
// Parameter tables that are already laid out in global memory the way a CTA wants them in shared memory (prepared LUT
// records, prepared per-channel affine parameters) are staged with the bulk-copy engine instead of per-thread loops:
// one elected thread arms an mbarrier with the byte count and issues cp.async.bulk (1-D TMA, no tensor map) copies;
// the copies run while all threads of the CTA issue their streaming loads of the data tile, and everybody waits on
// the barrier's phase 0 just before the first use.  Addresses and sizes must be multiples of 16 bytes.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");     // make the init visible to the async proxy
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}

// ------------------------------------------------------------------------------------------ host helpers
// NVTX (header-only v3: no link dependency; a no-op unless a profiler injected itself) -- one range per C-ABI call, named
// after the entry point, so that an `ncu --nvtx` / Nsight Systems timeline shows which call of the reference's API a kernel
// belongs to.  Costs one predictable branch when off.
struct NvtxRange {
    bool on;
    explicit NvtxRange(const char* name);
    ~NvtxRange();
};
#define MCTQ_NVTX(name) mctq::NvtxRange mctq_nvtx_range_(name)
inline int cuda_rc(cudaError_t e) { return e == cudaSuccess ? 0 : (int)e; }

// launch with the programmatic-serialization attribute; the caller has already declared the launch to pdl_plan_launch
template <typename... KArgs, typename... Args>
inline int launch_planned(void (*kernel)(KArgs...), dim3 grid, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cuda_rc(e);
}

// same, for kernels that always use the late order: their outputs are recorded as "unknown"
template <typename... KArgs, typename... Args>
inline int launch_streaming(void (*kernel)(KArgs...), dim3 grid, size_t smem, cudaStream_t st, Args&&... args) {
    pdl_plan_launch(st, nullptr, 0, nullptr, 0);
    return launch_planned(kernel, grid, smem, st, std::forward<Args>(args)...);
}

template <typename K>
int ensure_smem(K kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return 0;
    if (bytes > 200 * 1024) return MCTQ_E_BADARG;
    return cuda_rc(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; }

// geometry of the channel window for a tile of `tile` elements
template <class Args>
void set_window(Args& a, uint32_t tile) {
    a.bigrow = a.inner >= (int64_t)tile ? 1u : 0u;
    int64_t rows = a.bigrow ? 2 : (int64_t)tile / a.inner + 2;
    a.W = (uint32_t)(rows < a.C ? rows : a.C);
    a.div_inner = make_fastdiv(a.bigrow ? 1u : (uint32_t)a.inner);
    a.div_W = make_fastdiv(a.W);
}

}  // namespace mctq
