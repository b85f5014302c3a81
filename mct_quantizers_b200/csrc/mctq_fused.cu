// mctq_fused.cu -- activation fake-quant fused with its elementwise producer.
//
// In an MCT-exported model every PytorchActivationQuantizationHolder follows the op that produced its input
// (mct_quantizers/pytorch/activation_quantization_holder.py:43-53); when that producer is a ReLU / ReLU6 or a residual
// add, running it inside the fake-quant kernel removes one full write + read of the activation:
//     relu -> holder      16 B / element (f32)  ->  8 B
//     add  -> holder      20 B / element        -> 12 B
// The fused result is bit-identical to the eager composition: the producer is evaluated in f32 and rounded to the
// tensor dtype (what the eager kernel would have stored), then quantised with the per-tensor affine recipe of
// mctq_affine.cu.
#include "mctq_common.cuh"

namespace mctq {

constexpr float kMagicF = 12582912.0f;

struct FusedArgs {
    const void* x;
    const void* x2;
    void* y;
    int64_t n;
    float scale;
    int32_t zp, qmin, qmax;
    uint32_t early;          // dependent-launch order: 0 late, 1 early, 2 free (see pdl_plan_launch)
};

template <typename T, int PRE>
__device__ __forceinline__ float producer(float a, float b) {
    float t = a;
    if (PRE == MCTQ_PRE_ADD || PRE == MCTQ_PRE_ADD_RELU) t = to_f32<T>(from_f32<T>(__fadd_rn(a, b)));   // eager add stores T
    if (PRE == MCTQ_PRE_RELU || PRE == MCTQ_PRE_ADD_RELU) t = (t < 0.0f) ? 0.0f : t;                      // NaN stays NaN (-> qmin)
    if (PRE == MCTQ_PRE_RELU6) t = (t < 0.0f) ? 0.0f : ((t > 6.0f) ? 6.0f : t);
    return t;
}

template <typename T, int PRE, int UNROLL>
__global__ void __launch_bounds__(kThreads) fq_affine_pre_kernel(const FusedArgs a) {
    constexpr int V = 16 / sizeof(T);
    constexpr uint32_t TILE = kThreads * UNROLL * V;
    constexpr bool TWO = PRE == MCTQ_PRE_ADD || PRE == MCTQ_PRE_ADD_RELU;
    const uint32_t tid = threadIdx.x;
    const int64_t t0 = (int64_t)blockIdx.x * TILE;
    const int64_t remaining = a.n - t0;
    const bool full = remaining >= (int64_t)TILE;
    const T* xt = reinterpret_cast<const T*>(a.x) + t0;
    const T* bt = reinterpret_cast<const T*>(a.x2) + t0;
    pdl_enter(a.early);

    uint32_t w[UNROLL][4], v[UNROLL][4];
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
        const int64_t l = (int64_t)(j * kThreads + tid) * V;
        if (full || l + V <= remaining) {
            ld_words<4>(xt + l, w[j]);
            if (TWO) ld_words<4>(bt + l, v[j]);
        } else {
            T tx[V], tb[V];
#pragma unroll
            for (int e = 0; e < V; ++e) {
                const bool ok = l + e < remaining;
                tx[e] = ok ? xt[l + e] : from_f32<T>(0.0f);
                if (TWO) tb[e] = ok ? bt[l + e] : from_f32<T>(0.0f);
            }
            memcpy(w[j], tx, 16);
            if (TWO) memcpy(v[j], tb, 16);
        }
    }
    pdl_loaded(a.early);

    const float s = a.scale;
    const float inv = __fdiv_rn(1.0f, s);
    const float lo = (float)(a.qmin - a.zp), hi = (float)(a.qmax - a.zp);
    T* yt = reinterpret_cast<T*>(a.y) + t0;
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
        const int64_t l = (int64_t)(j * kThreads + tid) * V;
        float f[V], g[V];
        Pack<T, V>::unpack(w[j], f);
        if (TWO) Pack<T, V>::unpack(v[j], g);
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const float t = __fmul_rn(producer<T, PRE>(f[e], TWO ? g[e] : 0.0f), inv);
            const float r = __fsub_rn(__fadd_rn(t, kMagicF), kMagicF);
            f[e] = __fmul_rn(fminf(fmaxf(r, lo), hi), s);
        }
        if (full || l + V <= remaining) {
            Pack<T, V>::pack(f, w[j]);
            st_words<4>(yt + l, w[j]);
        } else if (l < remaining) {
            for (int e = 0; e < V && l + e < remaining; ++e) yt[l + e] = from_f32<T>(f[e]);
        }
    }
    pdl_exit(a.early);
}

// element-per-thread variant for misaligned views
template <typename T>
__global__ void __launch_bounds__(kThreads) fq_affine_pre_scalar_kernel(const FusedArgs a, int pre) {
    const float inv = __fdiv_rn(1.0f, a.scale);
    const float lo = (float)(a.qmin - a.zp), hi = (float)(a.qmax - a.zp);
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < a.n; i += stride) {
        const float x = to_f32<T>(reinterpret_cast<const T*>(a.x)[i]);
        const float b = a.x2 ? to_f32<T>(reinterpret_cast<const T*>(a.x2)[i]) : 0.0f;
        float p;
        switch (pre) {
            case MCTQ_PRE_RELU: p = producer<T, MCTQ_PRE_RELU>(x, b); break;
            case MCTQ_PRE_RELU6: p = producer<T, MCTQ_PRE_RELU6>(x, b); break;
            case MCTQ_PRE_ADD: p = producer<T, MCTQ_PRE_ADD>(x, b); break;
            case MCTQ_PRE_ADD_RELU: p = producer<T, MCTQ_PRE_ADD_RELU>(x, b); break;
            default: p = x;
        }
        const float t = __fmul_rn(p, inv);
        const float r = __fsub_rn(__fadd_rn(t, kMagicF), kMagicF);
        reinterpret_cast<T*>(a.y)[i] = from_f32<T>(__fmul_rn(fminf(fmaxf(r, lo), hi), a.scale));
    }
}

}  // namespace mctq

using namespace mctq;

namespace {

template <typename T, int PRE>
int launch_pre(const FusedArgs& a_in, cudaStream_t st) {
    FusedArgs a = a_in;
    constexpr int UNROLL = (PRE == MCTQ_PRE_ADD || PRE == MCTQ_PRE_ADD_RELU) ? 2 : 4;      // two input streams: same bytes in flight
    constexpr uint32_t TILE = kThreads * UNROLL * (16 / sizeof(T));
    const int64_t tiles = (a.n + TILE - 1) / TILE;
    if (tiles > 0x7fffffffLL) return MCTQ_E_BADARG;
    const IoSpan in[2] = {{a.x, (size_t)a.n * sizeof(T)}, {a.x2, (size_t)a.n * sizeof(T)}};
    const IoSpan out[1] = {{a.y, (size_t)a.n * sizeof(T)}};
    a.early = (uint32_t)pdl_plan_launch(st, in, 2, out, 1);
    return launch_planned(fq_affine_pre_kernel<T, PRE, UNROLL>, (unsigned)tiles, 0, st, a);
}

template <typename T>
int launch_pre_typed(const FusedArgs& a, int pre, cudaStream_t st) {
    const bool vec_ok = aligned16(a.x) && aligned16(a.y) && (!a.x2 || aligned16(a.x2));
    if (!vec_ok) {
        int64_t blocks = (a.n + kThreads - 1) / kThreads;
        if (blocks > 148 * 64) blocks = 148 * 64;
        fq_affine_pre_scalar_kernel<T><<<(unsigned)blocks, kThreads, 0, st>>>(a, pre);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        return cuda_rc(cudaGetLastError());
    }
    switch (pre) {
        case MCTQ_PRE_RELU: return launch_pre<T, MCTQ_PRE_RELU>(a, st);
        case MCTQ_PRE_RELU6: return launch_pre<T, MCTQ_PRE_RELU6>(a, st);
        case MCTQ_PRE_ADD: return launch_pre<T, MCTQ_PRE_ADD>(a, st);
        case MCTQ_PRE_ADD_RELU: return launch_pre<T, MCTQ_PRE_ADD_RELU>(a, st);
        default: return MCTQ_E_BADARG;
    }
}

}  // namespace

extern "C" {

int mctq_fq_affine_scalar_pre(const void* x, const void* x2, void* y, int64_t n, int x_dtype, int pre_op, float scale,
                              int32_t zp, int32_t qmin, int32_t qmax, void* stream) {
    MCTQ_NVTX("mctq_fq_affine_scalar_pre");
    const bool two = pre_op == MCTQ_PRE_ADD || pre_op == MCTQ_PRE_ADD_RELU;
    if (!x || !y || n < 0 || (two && !x2) || pre_op < MCTQ_PRE_RELU || pre_op > MCTQ_PRE_ADD_RELU) return MCTQ_E_BADARG;
    if (qmin > qmax || !((int64_t)qmax - qmin < (1 << 21)) || zp < qmin || zp > qmax) return MCTQ_E_RANGE;
    if (n == 0) return 0;
    FusedArgs a;
    a.x = x; a.x2 = two ? x2 : nullptr; a.y = y; a.n = n; a.scale = scale; a.zp = zp; a.qmin = qmin; a.qmax = qmax; a.early = 0;
    cudaStream_t st = (cudaStream_t)stream;
    switch (x_dtype) {
        case MCTQ_F32: return launch_pre_typed<float>(a, pre_op, st);
        case MCTQ_BF16: return launch_pre_typed<__nv_bfloat16>(a, pre_op, st);
        case MCTQ_F16: return launch_pre_typed<__half>(a, pre_op, st);
        default: return MCTQ_E_DTYPE;
    }
}

}  // extern "C"
