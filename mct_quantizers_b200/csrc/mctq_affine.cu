// mctq_affine.cu -- affine (round-clip-dequant) fake-quant kernels and their C ABI entry points.
#include "mctq_common.cuh"

namespace mctq {

constexpr float kMagic = 12582912.0f;      // 1.5 * 2^23: (t + kMagic) - kMagic == rint(t) for |t| < 2^22
constexpr int32_t kFastRangeLimit = 1 << 21;

struct AffineArgs {
    const void* x;
    void* y;
    void* codes;
    int64_t n;
    const float* scale;     // device [C] or NULL (then scale_val / zp_val)
    const int32_t* zp;
    float scale_val;
    int32_t zp_val;
    int64_t C, inner, elem_offset;
    int32_t qmin, qmax;
    // channel-window geometry (host-derived)
    FastDiv div_inner, div_W;
    uint32_t W;             // staged channel slots
    uint32_t bigrow;        // inner >= tile elems: a tile touches at most two rows
    // CH_LAST (channel is the innermost dimension): parameters are staged for one period = lcm(C, V) of the channel
    // pattern, so that every aligned vector of V elements reads V consecutive entries
    uint32_t period;
    FastDiv div_period;
    // prepared parameters (mctq_affine_prepare): records {1/s, s, zp bits, (float)zp} per channel and the same values as
    // three arrays [inv | s | zp] of C4 = roundup(C, 4) entries each; NULL for the raw-parameter entry points
    const float4* prep_rec;
    const float* prep_soa;
    const int32_t* prep_flags;   // [0] != 0: some zero point is non-zero
    int64_t C4;
    uint32_t early;              // dependent-launch order: 0 late, 1 early, 2 free (opt-in, see pdl_plan_launch)
    uint32_t tab_early;          // PREP: the parameter blob may be staged before the dependent-launch wait (nobody is writing it)
};

constexpr size_t kPrepHeaderBytes = 16;
inline size_t prep_bytes(int64_t C) { int64_t C4 = (C + 3) & ~(int64_t)3; return kPrepHeaderBytes + (size_t)C * 16 + (size_t)C4 * 12; }

// (float)k for |k| < 2^22 without I2F (a quarter-rate conversion): the mantissa of 1.5 * 2^23 + k holds k
__device__ __forceinline__ float small_int_to_float(int k) { return __fsub_rn(__int_as_float(0x4B400000 + k), kMagic); }

template <bool RINT> struct AffineOp {
    using Args = AffineArgs;
    struct ChanParams { float inv, s, lo, hi; int zp; };
    static constexpr int kSmemFloatsPerChan = 3;

    __device__ static __forceinline__ ChanParams make(float s, int zp, const Args& a) {
        ChanParams p;
        p.s = s;
        p.inv = __fdiv_rn(1.0f, s);                  // IEEE reciprocal of the f32 scale (ATen: 1.0f / scale)
        p.zp = zp;
        if (RINT) { p.lo = (float)a.qmin; p.hi = (float)a.qmax; }
        else { p.lo = small_int_to_float(a.qmin - zp); p.hi = small_int_to_float(a.qmax - zp); }    // |q - zp| <= 2^21 on this path
        return p;
    }
    __device__ static __forceinline__ ChanParams uniform(const Args& a) {
        float s = a.scale ? __ldg(a.scale) : a.scale_val;
        int zp = a.zp ? __ldg(a.zp) : a.zp_val;
        return make(s, zp, a);
    }
    __device__ static __forceinline__ void stage(float* sm, uint32_t W, uint32_t slot, int64_t c, const Args& a) {
        ChanParams p = make(__ldg(a.scale + c), __ldg(a.zp + c), a);
        sm[slot] = p.inv;
        sm[W + slot] = p.s;
        sm[2 * W + slot] = __int_as_float(p.zp);
    }
    __device__ static __forceinline__ ChanParams fetch(const float* sm, uint32_t W, uint32_t slot, const Args& a) {
        ChanParams p;
        p.inv = sm[slot];
        p.s = sm[W + slot];
        p.zp = __float_as_int(sm[2 * W + slot]);
        if (RINT) { p.lo = (float)a.qmin; p.hi = (float)a.qmax; }
        else { p.lo = small_int_to_float(a.qmin - p.zp); p.hi = small_int_to_float(a.qmax - p.zp); }
        return p;
    }
    // prepared record {1/s, s, zp bits, (float)zp}: one 16-byte shared-memory load, no int -> float conversions
    // ((float)(qmin - zp) == (float)qmin - (float)zp exactly: all three are integers below 2^24)
    __device__ static __forceinline__ ChanParams fetch_rec(const float* sm, uint32_t slot, const Args& a) {
        const float4 r = reinterpret_cast<const float4*>(sm)[slot];
        ChanParams p;
        p.inv = r.x;
        p.s = r.y;
        p.zp = __float_as_int(r.z);
        if (RINT) { p.lo = (float)a.qmin; p.hi = (float)a.qmax; }
        else { p.lo = __fsub_rn((float)a.qmin, r.w); p.hi = __fsub_rn((float)a.qmax, r.w); }
        return p;
    }
    // returns y; code receives the clamped integer q
    template <bool WANT_CODE>
    __device__ static __forceinline__ float apply(float x, const ChanParams& p, int& code) {
        float t = __fmul_rn(x, p.inv);
        if (!RINT) {
            // (t + M) - M rounds half-to-even exactly for |t| < 2^22 and stays out of [lo, hi] beyond it;
            // anything that rounds to zero comes out as +0.0, so y never carries the sign of a tiny negative
            // input (ATen subtracts the integer zero point, which has the same effect).  Clamping
            // r to [qmin - zp, qmax - zp] equals clamp(r + zp, qmin, qmax) - zp on integer-valued floats.
            float r = __fsub_rn(__fadd_rn(t, kMagic), kMagic);
            float c = fminf(fmaxf(r, p.lo), p.hi);           // fmaxf(NaN, lo) = lo: NaN -> qmin
            // integer code without F2I (a quarter-rate conversion that made the codes-only bf16 variant issue bound): c is an
            // integer with |c| < 2^22, so the low mantissa bits of c + 1.5 * 2^23 ARE c in two's complement
            if (WANT_CODE) code = __float_as_int(__fadd_rn(c, kMagic)) + (p.zp - 0x4B400000);
            return __fmul_rn(c, p.s);
        } else {
            float zf = (float)p.zp;
            float q = __fadd_rn(rintf(t), zf);
            q = fminf(fmaxf(q, p.lo), p.hi);
            if (WANT_CODE) code = (int)q;
            return __fmul_rn(__fsub_rn(q, zf), p.s);
        }
    }
};

// ------------------------------------------------------------------------------------------ affine kernel
// PREP: the per-channel parameters come from a prepared blob and are staged with 1-D TMA bulk copies (one elected thread,
// one mbarrier) instead of per-thread load / divide / store loops -- the staging cost of a tile that touches hundreds
// of channels (short rows, channel-innermost layouts) drops to a handful of instructions.
// VB: bytes per vector (16, or 8 for 2-byte types in CH_LAST mode: with 16-byte vectors of bf16 a lane needs 8 consecutive
// parameter entries = two LDS.128 at a lane stride of 32 bytes, which is a 2-way bank conflict and made the kernel
// shared-memory bound; 8-byte vectors need one LDS.128 per array at a lane stride of 16 bytes, conflict free).
template <typename T, int CHMODE, int CODE, int UNROLL, bool RINT, bool PREP, int VB = 16>
__global__ void __launch_bounds__(kThreads) fq_affine_kernel(const AffineArgs a) {
    using Op = AffineOp<RINT>;
    constexpr int V = VB / sizeof(T);
    constexpr int WORDS = VB / 4;
    constexpr uint32_t TILE = kThreads * UNROLL * V;
    extern __shared__ __align__(16) float sm_par[];
    __shared__ Window sm_win;
    __shared__ __align__(8) uint64_t sm_bar;

    const uint32_t tid = threadIdx.x;
    const int64_t t0 = (int64_t)blockIdx.x * TILE;
    const int64_t remaining = a.n - t0;
    const bool full = remaining >= (int64_t)TILE;
    const T* xt = reinterpret_cast<const T*>(a.x) + t0;
    bool has_zp = true;
    // PREP: one elected thread arms the mbarrier and hands the tile's channel window to the bulk-copy engine.  When the
    // host knows that the blob is not being written (tab_early) this happens BEFORE the dependent-launch wait, so a CTA
    // scheduled into the predecessor's tail has its parameters in shared memory when the wait returns.
    if constexpr (PREP && CHMODE != CH_PT) {
        auto stage = [&]() {
            if (tid != 0) return;
            mbar_init(&sm_bar, 1);
            if (CHMODE == CH_LAST) {
                // period / C back-to-back copies of each parameter array: entry i = channel i % C
                const uint32_t Cb = (uint32_t)a.C * 4u, reps = a.period / (uint32_t)a.C;
                mbar_arrive_expect_tx(&sm_bar, 3u * a.period * 4u);
                for (uint32_t arr = 0; arr < 3; ++arr)
                    for (uint32_t r = 0; r < reps; ++r)
                        bulk_g2s(sm_par + arr * a.period + r * (uint32_t)a.C, a.prep_soa + arr * a.C4, Cb, &sm_bar);
            } else {
                const int64_t g0 = a.elem_offset + t0;
                const int64_t r0 = g0 / a.inner;
                const int64_t off = g0 - r0 * a.inner;
                Window wv;
                wv.off0 = a.bigrow ? 0u : (uint32_t)off;
                const int64_t sp = a.inner - off;
                wv.split = (uint32_t)(sp > (int64_t)TILE ? (int64_t)TILE + 1 : sp);
                sm_win = wv;
                const uint32_t c0 = (uint32_t)(r0 % a.C);
                const uint32_t n1 = min(a.W, (uint32_t)a.C - c0);          // records before the channel index wraps
                mbar_arrive_expect_tx(&sm_bar, a.W * 16u);
                bulk_g2s(sm_par, a.prep_rec + c0, n1 * 16u, &sm_bar);
                if (n1 < a.W) bulk_g2s(sm_par + 4u * n1, a.prep_rec, (a.W - n1) * 16u, &sm_bar);
            }
        };
        // staging always precedes the tile loads; what moves is the wait (the host only picks an order other than "late"
        // together with tab_early)
        if (!a.tab_early) pdl_enter(0);
        stage();
        if (a.tab_early) pdl_enter(a.early);
    } else {
        pdl_enter(a.early);
    }

    uint32_t w[UNROLL][WORDS];
    if (full) {
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) ld_words<WORDS>(xt + (size_t)(j * kThreads + tid) * V, w[j]);
    } else {
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) {
            int64_t l = (int64_t)(j * kThreads + tid) * V;
            if (l + V <= remaining) ld_words<WORDS>(xt + l, w[j]);
            else {
                T tmp[V];
#pragma unroll
                for (int e = 0; e < V; ++e) tmp[e] = (l + e < remaining) ? xt[l + e] : from_f32<T>(0.0f);
                memcpy(w[j], tmp, VB);
            }
        }
    }
    pdl_loaded(a.early);                                   // the tile is in flight; nothing is written before this point

    typename Op::ChanParams pu;
    Window win;
    uint32_t base_mod = 0;
    if (CHMODE == CH_PT) {
        pu = Op::uniform(a);
    } else if (PREP) {
        // (read after the tile loads are issued: every thread needs it, and a wait in front of the loads would serialise
        // an L2 round trip with them)
        if (CHMODE == CH_LAST || CHMODE == CH_ELEM) has_zp = __ldg(a.prep_flags) != 0;
        if (CHMODE == CH_LAST) base_mod = (uint32_t)((uint64_t)(a.elem_offset + t0) % a.period);
        __syncthreads();                                               // barrier init + window visible to everyone
        if (CHMODE != CH_LAST) win = sm_win;
        mbar_wait(&sm_bar, 0);
    } else if (CHMODE == CH_LAST) {
        // entry i of the period holds the parameters of channel i % C (struct of arrays: inv | s | zp)
        int any_zp = 0;
        for (uint32_t i = tid; i < a.period; i += kThreads) {
            const uint32_t c = i - fdiv_u32(i, a.div_W) * a.div_W.d;              // div_W divides by C here
            Op::stage(sm_par, a.period, i, (int64_t)c, a);
            any_zp |= __ldg(a.zp + c);
        }
        base_mod = (uint32_t)((uint64_t)(a.elem_offset + t0) % a.period);
        // symmetric quantizers have all-zero zero points: then a third of the shared-memory reads can be skipped
        has_zp = __syncthreads_or(any_zp) != 0;
    } else {
        stage_window<Op>(sm_par, &sm_win, a.elem_offset + t0, TILE, a);
        __syncthreads();
        win = sm_win;
    }

    T* yt = reinterpret_cast<T*>(a.y) + t0;
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
        const uint32_t l = (uint32_t)(j * kThreads + tid) * V;
        float f[V];
        int code[V];
        Pack<T, V>::unpack(w[j], f);
        if (CHMODE == CH_PT) {
#pragma unroll
            for (int e = 0; e < V; ++e) f[e] = Op::template apply<CODE != 0>(f[e], pu, code[e]);
        } else if (CHMODE == CH_LAST) {
            const uint32_t m = base_mod + l;
            const uint32_t i0 = m - fdiv_u32(m, a.div_period) * a.period;            // multiple of V: vector loads below
            float pinv[V], ps[V], pz[V];
#pragma unroll
            for (int e = 0; e < V; e += 4) {
                *reinterpret_cast<float4*>(&pinv[e]) = *reinterpret_cast<const float4*>(&sm_par[i0 + e]);
                *reinterpret_cast<float4*>(&ps[e]) = *reinterpret_cast<const float4*>(&sm_par[a.period + i0 + e]);
                if (has_zp) *reinterpret_cast<float4*>(&pz[e]) = *reinterpret_cast<const float4*>(&sm_par[2 * a.period + i0 + e]);
                else *reinterpret_cast<float4*>(&pz[e]) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (!has_zp) {
                typename Op::ChanParams p;
                p.zp = 0;
                p.lo = (float)a.qmin;
                p.hi = (float)a.qmax;
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    p.inv = pinv[e];
                    p.s = ps[e];
                    f[e] = Op::template apply<CODE != 0>(f[e], p, code[e]);
                }
            } else {
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    typename Op::ChanParams p;
                    p.inv = pinv[e];
                    p.s = ps[e];
                    p.zp = __float_as_int(pz[e]);
                    if (RINT) { p.lo = (float)a.qmin; p.hi = (float)a.qmax; }
                    else { p.lo = small_int_to_float(a.qmin - p.zp); p.hi = small_int_to_float(a.qmax - p.zp); }
                    f[e] = Op::template apply<CODE != 0>(f[e], p, code[e]);
                }
            }
        } else if (CHMODE == CH_VEC) {
            uint32_t slot, rem;
            locate(l, win, a, slot, rem);
            typename Op::ChanParams p = PREP ? Op::fetch_rec(sm_par, slot, a) : Op::fetch(sm_par, a.W, slot, a);
#pragma unroll
            for (int e = 0; e < V; ++e) f[e] = Op::template apply<CODE != 0>(f[e], p, code[e]);
        } else if (a.inner >= V) {
            // CH_ELEM, rows at least one vector long: the vector straddles at most one row boundary, so two
            // parameter sets and a per-element select replace V shared-memory fetches
            uint32_t slot, rem, k;                  // k = elements of this vector that still belong to the first row
            if (a.bigrow) {
                const bool second = l >= win.split;
                slot = second ? 1u : 0u;
                rem = 0;
                k = second ? (uint32_t)V : min((uint32_t)V, win.split - l);
            } else {
                locate(l, win, a, slot, rem);
                k = a.div_inner.d - rem;
            }
            const uint32_t slot1 = (slot + 1 == a.W) ? 0u : slot + 1;
            const typename Op::ChanParams p0 = PREP ? Op::fetch_rec(sm_par, slot, a) : Op::fetch(sm_par, a.W, slot, a);
            const typename Op::ChanParams p1 = PREP ? Op::fetch_rec(sm_par, slot1, a) : Op::fetch(sm_par, a.W, slot1, a);
            if (PREP && !has_zp) {
                // all zero points are zero (symmetric quantizers): the clamp bounds are the same for every channel, only
                // 1/s and s are selected per element
                typename Op::ChanParams p = p0;
                p.zp = 0;
                p.lo = (float)a.qmin;
                p.hi = (float)a.qmax;
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    const bool first = (uint32_t)e < k;
                    p.inv = first ? p0.inv : p1.inv;
                    p.s = first ? p0.s : p1.s;
                    f[e] = Op::template apply<CODE != 0>(f[e], p, code[e]);
                }
            } else {
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    const bool first = (uint32_t)e < k;
                    typename Op::ChanParams p;
                    p.inv = first ? p0.inv : p1.inv;
                    p.s = first ? p0.s : p1.s;
                    p.lo = first ? p0.lo : p1.lo;
                    p.hi = first ? p0.hi : p1.hi;
                    p.zp = first ? p0.zp : p1.zp;
                    f[e] = Op::template apply<CODE != 0>(f[e], p, code[e]);
                }
            }
        } else {
            // rows shorter than a vector (inner < V, not channel-last): per-element parameter fetch
            uint32_t slot, rem;
            locate(l, win, a, slot, rem);
#pragma unroll
            for (int e = 0; e < V; ++e) {
                typename Op::ChanParams p = PREP ? Op::fetch_rec(sm_par, slot, a) : Op::fetch(sm_par, a.W, slot, a);
                f[e] = Op::template apply<CODE != 0>(f[e], p, code[e]);
                if (++rem == a.div_inner.d) { rem = 0; slot = (slot + 1 == a.W) ? 0 : slot + 1; }
            }
        }
        if (full || (int64_t)l + V <= remaining) {
            if (a.y) { Pack<T, V>::pack(f, w[j]); st_words<WORDS>(yt + l, w[j]); }
            if (CODE != 0) st_codes<V, CODE>(a.codes, t0 + l, code);
        } else if ((int64_t)l < remaining) {
            const int cnt = (int)(remaining - l);
            for (int e = 0; e < V; ++e) {
                if (e < cnt) {
                    if (a.y) yt[l + e] = from_f32<T>(f[e]);
                    if (CODE == MCTQ_CODES_INT8) reinterpret_cast<uint8_t*>(a.codes)[t0 + l + e] = (uint8_t)(code[e] & 0xff);
                }
            }
            if (CODE == MCTQ_CODES_INT4) {
                uint8_t* cp = reinterpret_cast<uint8_t*>(a.codes) + ((t0 + l) >> 1);
                for (int e = 0; e < V; e += 2) {
                    if (e < cnt) {
                        int hi = (e + 1 < cnt) ? code[e + 1] : 0;
                        cp[e >> 1] = (uint8_t)((code[e] & 0xf) | ((hi & 0xf) << 4));
                    }
                }
            }
        }
    }
    pdl_exit(a.early);
}

// ------------------------------------------------------------------------------------------ parameter preparation
// blob = [16-byte header: int32 any_nonzero_zp][records float4 x C][inv C4 | s C4 | zp C4]
__global__ void __launch_bounds__(kThreads) affine_prepare_kernel(const float* __restrict__ scale, const int32_t* __restrict__ zp,
                                                                  int64_t C, int64_t C4, uint8_t* blob) {
    int32_t* flags = reinterpret_cast<int32_t*>(blob);
    float4* rec = reinterpret_cast<float4*>(blob + kPrepHeaderBytes);
    float* soa = reinterpret_cast<float*>(blob + kPrepHeaderBytes + (size_t)C * 16);
    for (int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x; c < C4; c += (int64_t)gridDim.x * kThreads) {
        float s = 1.0f;
        int z = 0;
        if (c < C) { s = scale[c]; z = zp[c]; }
        const float inv = __fdiv_rn(1.0f, s);
        if (c < C) rec[c] = make_float4(inv, s, __int_as_float(z), (float)z);
        soa[c] = inv;
        soa[C4 + c] = s;
        soa[2 * C4 + c] = __int_as_float(z);
        if (z != 0) atomicOr(flags, 1);
    }
}

// ------------------------------------------------------------------------------------------ scalar fallbacks
// Misaligned base pointers (views with odd offsets): element-per-thread, still coalesced.
template <typename T, bool RINT>
__global__ void __launch_bounds__(kThreads) fq_affine_scalar_kernel(const AffineArgs a) {
    using Op = AffineOp<RINT>;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < a.n; i += stride) {
        int64_t c = 0;
        if (a.C > 1) c = ((a.elem_offset + i) / a.inner) % a.C;
        typename Op::ChanParams p = a.scale ? Op::make(__ldg(a.scale + c), __ldg(a.zp + c), a) : Op::make(a.scale_val, a.zp_val, a);
        int code;
        float y = Op::template apply<true>(to_f32<T>(reinterpret_cast<const T*>(a.x)[i]), p, code);
        if (a.y) reinterpret_cast<T*>(a.y)[i] = from_f32<T>(y);
        if (a.codes) reinterpret_cast<int8_t*>(a.codes)[i] = (int8_t)code;   // scalar fallback emits INT8 only
    }
}

// ------------------------------------------------------------------------------------------ dequant of codes
template <int CODE>
__global__ void __launch_bounds__(kThreads) dequant_affine_kernel(const void* codes, int is_signed, float* y, int64_t n,
                                                                  const float* scale, const int32_t* zp, int64_t C,
                                                                  int64_t inner, int64_t elem_offset) {
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
        int q;
        if (CODE == MCTQ_CODES_INT8) {
            uint8_t b = reinterpret_cast<const uint8_t*>(codes)[i];
            q = is_signed ? (int)(int8_t)b : (int)b;
        } else {
            uint8_t b = reinterpret_cast<const uint8_t*>(codes)[i >> 1];
            int nib = (i & 1) ? (b >> 4) : (b & 0xf);
            q = is_signed ? ((nib ^ 8) - 8) : nib;
        }
        int64_t c = C > 1 ? ((elem_offset + i) / inner) % C : 0;
        y[i] = __fmul_rn((float)(q - __ldg(zp + c)), __ldg(scale + c));
    }
}

// Streaming consumer of the code wire format: tile = 256 threads x 4 vectors x 4 elements; a vector is 4 codes (one 32-bit
// load for int8, 16 bits for packed int4) and one 16-byte f32 store.  5 (int8) / 4.5 (int4) algorithmic bytes per element.
//   MODE 0  per-tensor: parameters loaded once per thread
//   MODE 1  per-channel, rows a multiple of 4: one channel (one division, two parameter loads) per vector
//   MODE 2  per-channel, any row length below 2^32: the vector's first channel from one 32-bit division, the following
//           elements walk (rem, c) forward
// (q - zp) -> float goes through the magic-number trick (integer add into the mantissa of 1.5 * 2^23, one FADD) instead of
// a quarter-rate I2F: exact for |q - zp| < 2^22.
// V: codes per vector.  4 (one 32-bit / 16-bit load, one 16-byte store) or 8 (64-bit / 32-bit load, ONE 32-byte store:
// STG.256) when the whole vector lies in one channel -- half the loads, stores and channel look-ups per element.
template <int CODE, int MODE, int V>
__global__ void __launch_bounds__(kThreads) dequant_affine_vec_kernel(const void* codes, int is_signed, float* y, int64_t n,
                                                                      const float* __restrict__ scale, const int32_t* __restrict__ zp,
                                                                      uint32_t C, uint32_t inner, int64_t elem_offset,
                                                                      const FastDiv div_inner, const FastDiv div_C) {
    constexpr int UNROLL = 4;
    constexpr int64_t TILE = (int64_t)kThreads * UNROLL * V;
    constexpr int RAW = (CODE == MCTQ_CODES_INT8 && V == 8) ? 2 : 1;           // 32-bit words of codes per vector
    static_assert(V == 4 || (V == 8 && MODE != 2), "vector width");
    const int64_t t0 = (int64_t)blockIdx.x * TILE;
    pdl_wait();
    pdl_launch_dependents();
    uint32_t raw[UNROLL][RAW];
    const uint8_t* cb = reinterpret_cast<const uint8_t*>(codes);
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
        const int64_t e = t0 + (int64_t)(j * kThreads + threadIdx.x) * V;
#pragma unroll
        for (int r = 0; r < RAW; ++r) raw[j][r] = 0;
        if (e + V <= n) {
            if (CODE == MCTQ_CODES_INT8) {
                if (V == 8) { const uint2 v = __ldg(reinterpret_cast<const uint2*>(cb + e)); raw[j][0] = v.x; raw[j][RAW - 1] = v.y; }
                else raw[j][0] = __ldg(reinterpret_cast<const uint32_t*>(cb + e));
            } else {
                if (V == 8) raw[j][0] = __ldg(reinterpret_cast<const uint32_t*>(cb + (e >> 1)));
                else raw[j][0] = __ldg(reinterpret_cast<const uint16_t*>(cb + (e >> 1)));
            }
        } else {
            for (int k = 0; k < V && e + k < n; ++k) {
                const int64_t i = e + k;
                if (CODE == MCTQ_CODES_INT8) raw[j][(k >> 2) % RAW] |= (uint32_t)cb[i] << (8 * (k & 3));
                else raw[j][0] |= (uint32_t)((cb[i >> 1] >> ((i & 1) * 4)) & 0xf) << (4 * k);
            }
        }
    }
    float s0 = 0.0f;
    int z0 = 0;
    if (MODE == 0) { s0 = __ldg(scale); z0 = __ldg(zp); }
    const bool small = (uint64_t)(elem_offset + n) < (1ull << 31);
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
        const int64_t e = t0 + (int64_t)(j * kThreads + threadIdx.x) * V;
        if (e >= n) continue;
        uint32_t c = 0, rem = 0;
        float sv = s0;
        int zv = z0;
        if (MODE != 0) {
            const uint64_t g = (uint64_t)(elem_offset + e);
            if (small) {                                        // the usual case (< 2^31 elements): multiply-shift division
                const uint32_t g32 = (uint32_t)g, r = fdiv_u32(g32, div_inner);
                rem = g32 - r * inner;
                c = r - fdiv_u32(r, div_C) * C;
            } else {
                const uint64_t r = g / inner;
                rem = (uint32_t)(g - r * inner);
                c = (uint32_t)(r % C);
            }
            sv = __ldg(scale + c);
            zv = __ldg(zp + c);
        }
        // MODE 2, rows of at least one vector: the vector crosses at most one row boundary, so the parameters of the next
        // channel are fetched up front and selected per element (no divergent reloads inside the element loop)
        const bool two_rows = MODE == 2 && inner >= (uint32_t)V;
        uint32_t first_cnt = V;
        float sv1 = sv;
        int zv1 = zv;
        if (two_rows) {
            first_cnt = inner - rem;
            const uint32_t c1 = (c + 1 == C) ? 0 : c + 1;
            sv1 = __ldg(scale + c1);
            zv1 = __ldg(zp + c1);
        }
        uint32_t out[V];
#pragma unroll
        for (int k = 0; k < V; ++k) {
            int q;
            if (CODE == MCTQ_CODES_INT8) {
                const uint32_t word = raw[j][(k >> 2) % RAW];
                q = is_signed ? (int)(int8_t)(word >> (8 * (k & 3))) : (int)((word >> (8 * (k & 3))) & 0xffu);
            } else {
                const int nib = (int)((raw[j][0] >> (4 * k)) & 0xfu);
                q = is_signed ? ((nib ^ 8) - 8) : nib;
            }
            float se = sv;
            int ze = zv;
            if (two_rows && (uint32_t)k >= first_cnt) { se = sv1; ze = zv1; }
            const float d = __fsub_rn(__int_as_float(0x4B400000 + (q - ze)), kMagic);        // (float)(q - zp), exact
            out[k] = __float_as_uint(__fmul_rn(d, se));
            if (MODE == 2 && !two_rows) {
                if (++rem == inner) {
                    rem = 0;
                    c = (c + 1 == C) ? 0 : c + 1;
                    sv = __ldg(scale + c);
                    zv = __ldg(zp + c);
                }
            }
        }
        if (e + V <= n) {
            if (V == 8) st_stream256(y + e, out);
            else st_stream(reinterpret_cast<uint4*>(y + e), make_uint4(out[0], out[1], out[2], out[3]));
        } else {
            for (int k = 0; k < V && e + k < n; ++k) y[e + k] = __uint_as_float(out[k]);
        }
    }
}

// ------------------------------------------------------------------------------------------ multi-tensor
constexpr int kMultiTile = 2048;   // elements per tile, any dtype (8192 measured: ResNet-18 29 -> 29 us, MobileNetV2 10 -> 17 us)

// channel of logical element i: 32-bit arithmetic whenever the tensor has fewer than 2^32 elements
__device__ __forceinline__ int64_t multi_channel(int64_t i, const MctqTensorDesc& d, bool small) {
    if (d.C <= 1) return 0;
    if (small) return (int64_t)(((uint32_t)i / (uint32_t)d.inner) % (uint32_t)d.C);
    return (i / d.inner) % d.C;
}

template <typename T, bool RINT>
__device__ __forceinline__ void multi_tile_body(const MctqTensorDesc& d, int64_t e0) {
    using Op = AffineOp<RINT>;
    constexpr int V = 4;                                   // elements per vector: 16 B (f32) or 8 B (bf16 / f16)
    constexpr int WORDS = V * sizeof(T) / 4;
    constexpr int NV = kMultiTile / (kThreads * V);        // vectors per thread (2)
    const T* x = reinterpret_cast<const T*>(d.x);
    T* y = reinterpret_cast<T*>(d.y);
    const int64_t rem = d.n - e0;
    const bool small = d.n < (1LL << 32) && d.inner < (1LL << 32) && d.C < (1LL << 32);
    AffineArgs a;
    a.qmin = d.qmin;
    a.qmax = d.qmax;
    const bool vec = (d.C <= 1 || d.inner % V == 0) && (reinterpret_cast<uintptr_t>(x) % (V * sizeof(T))) == 0 &&
                     (!y || (reinterpret_cast<uintptr_t>(y) % (V * sizeof(T))) == 0) &&
                     (!d.codes || d.code_mode != MCTQ_CODES_INT8 || (reinterpret_cast<uintptr_t>(d.codes) & 3u) == 0);
    if (vec) {
        // one channel per aligned vector of 4: load both vectors and their parameters first, then one reciprocal per vector
        uint32_t w[NV][WORDS];
        float sc[NV];
        int zp[NV];
        bool whole[NV];
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int64_t l = (int64_t)(j * kThreads + threadIdx.x) * V;
            whole[j] = l + V <= rem;
            if (whole[j]) ld_words<WORDS>(x + e0 + l, w[j]);
            else {
                T tmp[V];
#pragma unroll
                for (int e = 0; e < V; ++e) tmp[e] = (l + e < rem) ? x[e0 + l + e] : from_f32<T>(0.0f);
                memcpy(w[j], tmp, sizeof(tmp));
            }
            const int64_t c = l < rem ? multi_channel(e0 + l, d, small) : 0;
            sc[j] = __ldg(d.scale + c);
            zp[j] = __ldg(d.zp + c);
        }
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int64_t l = (int64_t)(j * kThreads + threadIdx.x) * V;
            if (l >= rem) continue;
            const typename Op::ChanParams p = Op::make(sc[j], zp[j], a);
            float f[V];
            int code[V];
            Pack<T, V>::unpack(w[j], f);
#pragma unroll
            for (int e = 0; e < V; ++e) f[e] = Op::template apply<true>(f[e], p, code[e]);
            if (whole[j]) {
                if (y) { Pack<T, V>::pack(f, w[j]); st_words<WORDS>(y + e0 + l, w[j]); }
                if (d.codes && d.code_mode == MCTQ_CODES_INT8) st_codes<V, MCTQ_CODES_INT8>(d.codes, e0 + l, code);
            } else {
                for (int e = 0; e < V && l + e < rem; ++e) {
                    if (y) y[e0 + l + e] = from_f32<T>(f[e]);
                    if (d.codes && d.code_mode == MCTQ_CODES_INT8) reinterpret_cast<uint8_t*>(d.codes)[e0 + l + e] = (uint8_t)(code[e] & 0xff);
                }
            }
        }
        return;
    }
    // rows that are not a multiple of 4 (depthwise 3x3, first conv ...) or misaligned views: 8 elements per thread,
    // strided by the CTA width so that every access is coalesced; one channel look-up per element
    const int64_t e1 = min(e0 + (int64_t)kMultiTile, d.n);
    float xv[kMultiTile / kThreads];
#pragma unroll
    for (int k = 0; k < kMultiTile / kThreads; ++k) {
        int64_t i = e0 + k * kThreads + threadIdx.x;
        xv[k] = i < e1 ? to_f32<T>(x[i]) : 0.0f;
    }
#pragma unroll
    for (int k = 0; k < kMultiTile / kThreads; ++k) {
        int64_t i = e0 + k * kThreads + threadIdx.x;
        if (i < e1) {
            const int64_t c = multi_channel(i, d, small);
            int code;
            const float out = Op::template apply<true>(xv[k], Op::make(__ldg(d.scale + c), __ldg(d.zp + c), a), code);
            if (y) y[i] = from_f32<T>(out);
            if (d.codes && d.code_mode == MCTQ_CODES_INT8) reinterpret_cast<uint8_t*>(d.codes)[i] = (uint8_t)(code & 0xff);
        }
    }
}

__global__ void __launch_bounds__(kThreads) fq_affine_multi_kernel(const MctqTensorDesc* __restrict__ descs,
                                                                   const int32_t* __restrict__ tile_starts, int n_desc) {
    pdl_wait();
    pdl_launch_dependents();
    // which tensor does this tile belong to?  upper-bound binary search over tile_starts
    const int tile = blockIdx.x;
    int lo = 0, hi = n_desc;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (__ldg(tile_starts + mid) <= tile) lo = mid; else hi = mid;
    }
    const MctqTensorDesc d = descs[lo];
    const int64_t e0 = (int64_t)(tile - __ldg(tile_starts + lo)) * kMultiTile;
    const bool fast = (int64_t)d.qmax - d.qmin < kFastRangeLimit;
    if (d.dtype == MCTQ_F32) { if (fast) multi_tile_body<float, false>(d, e0); else multi_tile_body<float, true>(d, e0); }
    else if (d.dtype == MCTQ_BF16) { if (fast) multi_tile_body<__nv_bfloat16, false>(d, e0); else multi_tile_body<__nv_bfloat16, true>(d, e0); }
    else { if (fast) multi_tile_body<__half, false>(d, e0); else multi_tile_body<__half, true>(d, e0); }
}

// ---- many per-tensor sites, one launch (activation holders whose inputs already exist; SURVEY 8f rank 1 applied to the
// activation side).  54 back-to-back launches of a MobileNetV2 step lose ~1.5 us each to the drain of one grid and the
// ramp of the next even with programmatic dependent launch (4 % of the step); one grid over all sites has neither.
// The site table travels as kernel parameters (constant bank: a binary search of a few LDC, no global look-up in front
// of the tile's loads); every CTA runs the per-tensor body of fq_affine_kernel<T, CH_PT> on one 8 KB tile of its site.
constexpr int kSitesCap = 64;
struct alignas(16) SiteEntry {
    const void* x;
    void* y;
    int64_t n;
    float scale;
    int32_t zp, qmin, qmax;
    int32_t dtype, pad;
};
struct SitesParams {
    int32_t n_sites, pad[3];
    int32_t starts[kSitesCap + 4];              // starts[k] = first tile of site k; starts[n_sites] = number of tiles
    SiteEntry e[kSitesCap];
};
static_assert(sizeof(SitesParams) <= 4096, "kernel parameter space");

template <typename T, bool RINT>
__device__ __forceinline__ void site_tile(const SiteEntry& s, const int64_t tile) {
    using Op = AffineOp<RINT>;
    constexpr int UNROLL = 2, V = 16 / sizeof(T), WORDS = 4;
    constexpr uint32_t TILE = kThreads * UNROLL * V;
    const uint32_t tid = threadIdx.x;
    const int64_t t0 = tile * TILE;
    const int64_t remaining = s.n - t0;
    const bool full = remaining >= (int64_t)TILE;
    const T* xt = reinterpret_cast<const T*>(s.x) + t0;
    T* yt = reinterpret_cast<T*>(s.y) + t0;
    uint32_t w[UNROLL][WORDS];
    if (full) {
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) ld_words<WORDS>(xt + (size_t)(j * kThreads + tid) * V, w[j]);
    } else {
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) {
            int64_t l = (int64_t)(j * kThreads + tid) * V;
            if (l + V <= remaining) ld_words<WORDS>(xt + l, w[j]);
            else {
                T tmp[V];
#pragma unroll
                for (int e = 0; e < V; ++e) tmp[e] = (l + e < remaining) ? xt[l + e] : from_f32<T>(0.0f);
                memcpy(w[j], tmp, 16);
            }
        }
    }
    AffineArgs a;                                   // only the fields AffineOp::make reads
    a.qmin = s.qmin;
    a.qmax = s.qmax;
    const typename Op::ChanParams p = Op::make(s.scale, s.zp, a);
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
        const uint32_t l = (uint32_t)(j * kThreads + tid) * V;
        float f[V];
        int code;
        Pack<T, V>::unpack(w[j], f);
#pragma unroll
        for (int e = 0; e < V; ++e) f[e] = Op::template apply<false>(f[e], p, code);
        if (full || (int64_t)l + V <= remaining) {
            Pack<T, V>::pack(f, w[j]);
            st_words<WORDS>(yt + l, w[j]);
        } else if ((int64_t)l < remaining) {
            const int cnt = (int)(remaining - l);
            for (int e = 0; e < V; ++e)
                if (e < cnt) yt[l + e] = from_f32<T>(f[e]);
        }
    }
}

__global__ void __launch_bounds__(kThreads) fq_affine_sites_kernel(const __grid_constant__ SitesParams p) {
    pdl_wait();
    pdl_launch_dependents();
    const int tile = blockIdx.x;
    int lo = 0, hi = p.n_sites;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (p.starts[mid] <= tile) lo = mid; else hi = mid;
    }
    const SiteEntry& s = p.e[lo];
    const int64_t t = (int64_t)(tile - p.starts[lo]);
    const bool fast = (int64_t)s.qmax - s.qmin < kFastRangeLimit;
    if (s.dtype == MCTQ_F32) { if (fast) site_tile<float, false>(s, t); else site_tile<float, true>(s, t); }
    else if (s.dtype == MCTQ_BF16) { if (fast) site_tile<__nv_bfloat16, false>(s, t); else site_tile<__nv_bfloat16, true>(s, t); }
    else { if (fast) site_tile<__half, false>(s, t); else site_tile<__half, true>(s, t); }
}

}  // namespace mctq

using namespace mctq;

namespace {

int check_codes(int32_t qmin, int32_t qmax, int code_mode, const void* codes) {
    if (code_mode == MCTQ_CODES_NONE) return 0;
    if (!codes) return MCTQ_E_BADARG;
    if (code_mode == MCTQ_CODES_INT8) {
        if ((qmin >= -128 && qmax <= 127) || (qmin >= 0 && qmax <= 255)) return 0;
        return MCTQ_E_RANGE;
    }
    if (code_mode == MCTQ_CODES_INT4) {
        if ((qmin >= -8 && qmax <= 7) || (qmin >= 0 && qmax <= 15)) return 0;
        return MCTQ_E_RANGE;
    }
    return MCTQ_E_BADARG;
}

// bytes per vector: 16, except channel-innermost layouts of 2-byte types (8: see fq_affine_kernel)
template <typename T, int CHMODE> constexpr int vec_bytes() { return (CHMODE == CH_LAST && sizeof(T) == 2) ? 8 : 16; }

template <typename T, int CHMODE, int CODE, int UNROLL_IN, bool RINT, bool PREP = false>
int launch_affine_tiles(const AffineArgs& a_in, cudaStream_t st) {
    constexpr int VB = vec_bytes<T, CHMODE>();
    constexpr int UNROLL = UNROLL_IN * 16 / VB;            // same bytes in flight per thread
    constexpr int V = VB / sizeof(T);
    constexpr uint32_t TILE = kThreads * UNROLL * V;
    AffineArgs a = a_in;
    size_t smem = 0;
    if (CHMODE == CH_LAST) {
        a.div_W = make_fastdiv((uint32_t)a.C);
        a.div_period = make_fastdiv(a.period);
        smem = (size_t)a.period * 3 * sizeof(float);
        int rc = ensure_smem(fq_affine_kernel<T, CHMODE, CODE, UNROLL, RINT, PREP, VB>, smem);
        if (rc) return rc;
    } else if (CHMODE != CH_PT) {
        set_window(a, TILE);
        smem = (size_t)a.W * (PREP ? 4 : 3) * sizeof(float);
        int rc = ensure_smem(fq_affine_kernel<T, CHMODE, CODE, UNROLL, RINT, PREP, VB>, smem);
        if (rc) return rc;
    }
    int64_t tiles = (a.n + TILE - 1) / TILE;
    if (tiles > 0x7fffffffLL) return MCTQ_E_BADARG;
    const IoSpan in[1] = {{a.x, (size_t)a.n * sizeof(T)}};
    const IoSpan out[2] = {{a.y, (size_t)a.n * sizeof(T)}, {a.codes, CODE == MCTQ_CODES_INT4 ? (size_t)(a.n + 1) / 2 : (size_t)a.n}};
    if (PREP && CHMODE != CH_PT) {
        const IoSpan tables = {a.prep_flags, prep_bytes(a.C)};            // the blob starts with its flags word
        a.early = (uint32_t)pdl_plan_launch(st, in, 1, out, CODE != 0 ? 2 : 1, &tables, &a.tab_early);
        if (!a.tab_early) a.early = 0;
    } else {
        a.early = (uint32_t)pdl_plan_launch(st, in, 1, out, CODE != 0 ? 2 : 1);
    }
    return launch_planned(fq_affine_kernel<T, CHMODE, CODE, UNROLL, RINT, PREP, VB>, (unsigned)tiles, smem, st, a);
}

// prepared parameters: TMA-staged variants (fast range, unroll 4, every code mode)
template <typename T, int CHMODE>
int launch_affine_prepared(const AffineArgs& a, int code_mode, cudaStream_t st) {
    switch (code_mode) {
        case MCTQ_CODES_INT8: return launch_affine_tiles<T, CHMODE, MCTQ_CODES_INT8, 4, false, true>(a, st);
        case MCTQ_CODES_INT4: return launch_affine_tiles<T, CHMODE, MCTQ_CODES_INT4, 4, false, true>(a, st);
        default: return launch_affine_tiles<T, CHMODE, MCTQ_CODES_NONE, 4, false, true>(a, st);
    }
}

template <typename T, int CHMODE, int CODE, bool RINT>
int launch_affine_unroll(const AffineArgs& a, cudaStream_t st) {
    // the unroll sweep (mctq_set_tuning key 0) exists only for the plain fake-quant variants
    if (CODE == MCTQ_CODES_NONE && !RINT) {
        // automatic: 8 KB tiles for per-tensor launches (nothing to stage per tile; tools/stream_probe.cu measures
        // +2 % over 16 KB tiles), 16 KB tiles when a channel window is staged per tile
        // (4 KB tiles were measured in round 2 and are slower at every size: 74 MB 5.72 -> 5.16 TB/s queued, 295 MB 6.63 ->
        // 5.94, 1 GB 6.91 -> 6.16 for bf16; the ~3 us a mid-size launch loses are drain + ramp, not tile granularity)
        const int u = g_unroll ? g_unroll : (CHMODE == CH_PT ? 2 : 4);
        if (u == 2) return launch_affine_tiles<T, CHMODE, MCTQ_CODES_NONE, 2, false>(a, st);
        if (u == 8) return launch_affine_tiles<T, CHMODE, MCTQ_CODES_NONE, 8, false>(a, st);
    }
    return launch_affine_tiles<T, CHMODE, CODE, 4, RINT>(a, st);
}

template <typename T, int CHMODE, bool RINT>
int launch_affine_code(const AffineArgs& a, int code_mode, cudaStream_t st) {
    switch (code_mode) {
        case MCTQ_CODES_INT8: return launch_affine_unroll<T, CHMODE, MCTQ_CODES_INT8, RINT>(a, st);
        case MCTQ_CODES_INT4: return launch_affine_unroll<T, CHMODE, MCTQ_CODES_INT4, RINT>(a, st);
        default: return launch_affine_unroll<T, CHMODE, MCTQ_CODES_NONE, RINT>(a, st);
    }
}

template <typename T>
int launch_affine_typed(const AffineArgs& a, int code_mode, cudaStream_t st) {
    constexpr int V = 16 / sizeof(T);
    const bool rint_path = g_force_rint || !(a.qmax - (int64_t)a.qmin < kFastRangeLimit);
    // vector path needs 16-byte aligned x / y and suitably aligned codes
    bool vec_ok = aligned16(a.x) && (!a.y || aligned16(a.y));
    if (code_mode != MCTQ_CODES_NONE) vec_ok = vec_ok && (reinterpret_cast<uintptr_t>(a.codes) & 7u) == 0;
    if (!vec_ok) {
        if (code_mode == MCTQ_CODES_INT4) return MCTQ_E_BADARG;
        int64_t blocks = (a.n + kThreads - 1) / kThreads;
        if (blocks > 148 * 64) blocks = 148 * 64;
        if (rint_path) fq_affine_scalar_kernel<T, true><<<(unsigned)blocks, kThreads, 0, st>>>(a);
        else fq_affine_scalar_kernel<T, false><<<(unsigned)blocks, kThreads, 0, st>>>(a);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        return cuda_rc(cudaGetLastError());
    }
    int chmode;
    AffineArgs a2 = a;
    if (a.C == 1) chmode = CH_PT;
    else if (a.inner % V == 0 && a.elem_offset % V == 0) chmode = CH_VEC;
    else {
        chmode = CH_ELEM;
        constexpr int VL = vec_bytes<T, CH_LAST>() / sizeof(T);       // elements per vector in CH_LAST mode
        if (a.inner == 1 && a.elem_offset % VL == 0 && a.C <= 4096) {
            // channel-innermost layout: one period of the channel pattern (lcm(C, VL) entries, <= 48 KB) in shared memory
            uint64_t g = (uint64_t)a.C, h = VL;
            while (h) { uint64_t t = g % h; g = h; h = t; }
            const uint64_t period = (uint64_t)a.C / g * VL;
            if (period <= 4096) { chmode = CH_LAST; a2.period = (uint32_t)period; }
        }
    }
    if (a.prep_rec && !rint_path && (g_unroll == 0 || g_unroll == 4)) {
        // prepared blob: bulk-copy staging needs 16-byte granules (CH_LAST: whole arrays of C floats)
        if (chmode == CH_VEC) return launch_affine_prepared<T, CH_VEC>(a2, code_mode, st);
        if (chmode == CH_ELEM) return launch_affine_prepared<T, CH_ELEM>(a2, code_mode, st);
        if (chmode == CH_LAST && a.C % 4 == 0) return launch_affine_prepared<T, CH_LAST>(a2, code_mode, st);
    }
#define MCTQ_DISPATCH_AFF(CM)                                                                   \
    return rint_path ? launch_affine_code<T, CM, true>(a2, code_mode, st) : launch_affine_code<T, CM, false>(a2, code_mode, st)
    if (chmode == CH_PT) { MCTQ_DISPATCH_AFF(CH_PT); }
    if (chmode == CH_VEC) { MCTQ_DISPATCH_AFF(CH_VEC); }
    if (chmode == CH_LAST) { MCTQ_DISPATCH_AFF(CH_LAST); }
    MCTQ_DISPATCH_AFF(CH_ELEM);
#undef MCTQ_DISPATCH_AFF
}

int launch_affine(const AffineArgs& a, int x_dtype, int code_mode, cudaStream_t st) {
    if (a.n < 0 || a.C < 1 || a.inner < 1 || a.elem_offset < 0 || !a.x || (!a.y && code_mode == MCTQ_CODES_NONE))
        return MCTQ_E_BADARG;
    if (a.qmin > a.qmax) return MCTQ_E_RANGE;
    int rc = check_codes(a.qmin, a.qmax, code_mode, a.codes);
    if (rc) return rc;
    if (a.n == 0) return 0;
    switch (x_dtype) {
        case MCTQ_F32: return launch_affine_typed<float>(a, code_mode, st);
        case MCTQ_BF16: return launch_affine_typed<__nv_bfloat16>(a, code_mode, st);
        case MCTQ_F16: return launch_affine_typed<__half>(a, code_mode, st);
        default: return MCTQ_E_DTYPE;
    }
}

}  // namespace

extern "C" {

int mctq_fq_affine(const void* x, void* y, void* codes, int64_t n, int x_dtype, const float* scale, const int32_t* zp,
                   int64_t C, int64_t inner, int64_t elem_offset, int32_t qmin, int32_t qmax, int code_mode,
                   void* stream) {
    MCTQ_NVTX("mctq_fq_affine");
    if (!scale || !zp) return MCTQ_E_BADARG;
    AffineArgs a;
    memset(&a, 0, sizeof(a));
    a.x = x; a.y = y; a.codes = codes; a.n = n; a.scale = scale; a.zp = zp;
    a.C = C; a.inner = C == 1 ? 1 : inner; a.elem_offset = C == 1 ? 0 : elem_offset; a.qmin = qmin; a.qmax = qmax;
    return launch_affine(a, x_dtype, code_mode, (cudaStream_t)stream);
}

size_t mctq_affine_prepared_bytes(int64_t C) { return C < 1 ? 0 : prep_bytes(C); }

int mctq_affine_prepare(const float* scale, const int32_t* zp, int64_t C, void* prepared_dev, size_t prepared_bytes, void* stream) {
    MCTQ_NVTX("mctq_affine_prepare");
    if (!scale || !zp || !prepared_dev || C < 1 || prepared_bytes < prep_bytes(C) || !aligned16(prepared_dev)) return MCTQ_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(prepared_dev, 0, kPrepHeaderBytes, st);
    if (e != cudaSuccess) return (int)e;
    const int64_t C4 = (C + 3) & ~(int64_t)3;
    int64_t blocks = (C4 + kThreads - 1) / kThreads;
    if (blocks > 148 * 8) blocks = 148 * 8;
    pdl_note_prepare(st, prepared_dev, prep_bytes(C));              // launches that follow stage this blob only after their wait
    affine_prepare_kernel<<<(unsigned)blocks, kThreads, 0, st>>>(scale, zp, C, C4, reinterpret_cast<uint8_t*>(prepared_dev));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cuda_rc(cudaGetLastError());
}

int mctq_fq_affine_prepared(const void* x, void* y, void* codes, int64_t n, int x_dtype, const void* prepared_dev, int64_t C,
                            int64_t inner, int64_t elem_offset, int32_t qmin, int32_t qmax, int code_mode, void* stream) {
    MCTQ_NVTX("mctq_fq_affine_prepared");
    if (!prepared_dev || C < 1 || !aligned16(prepared_dev)) return MCTQ_E_BADARG;
    const uint8_t* blob = reinterpret_cast<const uint8_t*>(prepared_dev);
    AffineArgs a;
    memset(&a, 0, sizeof(a));
    a.C4 = (C + 3) & ~(int64_t)3;
    a.prep_flags = reinterpret_cast<const int32_t*>(blob);
    a.prep_rec = reinterpret_cast<const float4*>(blob + kPrepHeaderBytes);
    a.prep_soa = reinterpret_cast<const float*>(blob + kPrepHeaderBytes + (size_t)C * 16);
    // the arrays inside the blob double as the raw parameters for the variants that do not use bulk staging
    a.scale = a.prep_soa + a.C4;
    a.zp = reinterpret_cast<const int32_t*>(a.prep_soa + 2 * a.C4);
    a.x = x; a.y = y; a.codes = codes; a.n = n;
    a.C = C; a.inner = C == 1 ? 1 : inner; a.elem_offset = C == 1 ? 0 : elem_offset; a.qmin = qmin; a.qmax = qmax;
    return launch_affine(a, x_dtype, code_mode, (cudaStream_t)stream);
}

int mctq_fq_affine_scalar(const void* x, void* y, void* codes, int64_t n, int x_dtype, float scale, int32_t zp,
                          int32_t qmin, int32_t qmax, int code_mode, void* stream) {
    MCTQ_NVTX("mctq_fq_affine_scalar");
    AffineArgs a;
    memset(&a, 0, sizeof(a));
    a.x = x; a.y = y; a.codes = codes; a.n = n; a.scale = nullptr; a.zp = nullptr; a.scale_val = scale; a.zp_val = zp;
    a.C = 1; a.inner = 1; a.elem_offset = 0; a.qmin = qmin; a.qmax = qmax;
    return launch_affine(a, x_dtype, code_mode, (cudaStream_t)stream);
}

int mctq_dequant_affine(const void* codes, int code_mode, int is_signed, float* y, int64_t n, const float* scale,
                        const int32_t* zp, int64_t C, int64_t inner, int64_t elem_offset, void* stream) {
    MCTQ_NVTX("mctq_dequant_affine");
    if (!codes || !y || !scale || !zp || n < 0 || C < 1 || inner < 1) return MCTQ_E_BADARG;
    if (n == 0) return 0;
    if (code_mode != MCTQ_CODES_INT8 && code_mode != MCTQ_CODES_INT4) return MCTQ_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (aligned16(y) && (reinterpret_cast<uintptr_t>(codes) & 3u) == 0 && C < (1LL << 32) && inner < (1LL << 32)) {
        const int mode = C == 1 ? 0 : ((inner % 4 == 0 && elem_offset % 4 == 0) ? 1 : 2);
        // 8 codes per vector (32-byte aligned y, the vector inside one channel) only where it measured faster on the B200:
        // per-channel int4 (5.63 -> 6.14 TB/s; the channel look-up per vector halves).  Per-tensor and int8 variants are
        // at 6.3-6.4 TB/s with 4-code vectors and lose ~3 % with 8 (key 5 = 2 forces 8 everywhere for experiments).
        const bool fits8 = mode != 2 && aligned32(y) && (code_mode != MCTQ_CODES_INT8 || (reinterpret_cast<uintptr_t>(codes) & 7u) == 0) &&
                           (mode == 0 || (inner % 8 == 0 && elem_offset % 8 == 0));
        const bool v8 = fits8 && ((g_wide && mode == 1 && code_mode == MCTQ_CODES_INT4) || g_wide == 2);
        const int64_t tile = (int64_t)kThreads * 4 * (v8 ? 8 : 4);
        const int64_t tiles = (n + tile - 1) / tile;
        if (tiles > 0x7fffffffLL) return MCTQ_E_BADARG;
        const uint32_t C32 = (uint32_t)C, in32 = (uint32_t)(C == 1 ? 1 : inner);
        const FastDiv di = make_fastdiv(in32), dc = make_fastdiv(C32);
#define MCTQ_DQ(CM, MD, VV) launch_streaming(dequant_affine_vec_kernel<CM, MD, VV>, (unsigned)tiles, 0, st, codes, is_signed, y, n, scale, zp, C32, in32, elem_offset, di, dc)
        if (code_mode == MCTQ_CODES_INT8) {
            if (v8) return mode == 0 ? MCTQ_DQ(MCTQ_CODES_INT8, 0, 8) : MCTQ_DQ(MCTQ_CODES_INT8, 1, 8);
            return mode == 0 ? MCTQ_DQ(MCTQ_CODES_INT8, 0, 4) : mode == 1 ? MCTQ_DQ(MCTQ_CODES_INT8, 1, 4) : MCTQ_DQ(MCTQ_CODES_INT8, 2, 4);
        }
        if (v8) return mode == 0 ? MCTQ_DQ(MCTQ_CODES_INT4, 0, 8) : MCTQ_DQ(MCTQ_CODES_INT4, 1, 8);
        return mode == 0 ? MCTQ_DQ(MCTQ_CODES_INT4, 0, 4) : mode == 1 ? MCTQ_DQ(MCTQ_CODES_INT4, 1, 4) : MCTQ_DQ(MCTQ_CODES_INT4, 2, 4);
#undef MCTQ_DQ
    }
    int64_t blocks = (n + kThreads - 1) / kThreads;
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (code_mode == MCTQ_CODES_INT8)
        dequant_affine_kernel<MCTQ_CODES_INT8><<<(unsigned)blocks, kThreads, 0, st>>>(codes, is_signed, y, n, scale, zp, C, inner, elem_offset);
    else if (code_mode == MCTQ_CODES_INT4)
        dequant_affine_kernel<MCTQ_CODES_INT4><<<(unsigned)blocks, kThreads, 0, st>>>(codes, is_signed, y, n, scale, zp, C, inner, elem_offset);
    else return MCTQ_E_BADARG;
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cuda_rc(cudaGetLastError());
}

int64_t mctq_multi_tile_elems(void) { return kMultiTile; }

int64_t mctq_multi_plan(const MctqTensorDesc* descs, int n_desc, int32_t* tile_starts) {
    if (!descs || !tile_starts || n_desc < 0) return MCTQ_E_BADARG;
    int64_t t = 0;
    for (int k = 0; k < n_desc; ++k) {
        if (descs[k].n < 0) return MCTQ_E_BADARG;
        tile_starts[k] = (int32_t)t;
        t += (descs[k].n + kMultiTile - 1) / kMultiTile;
        if (t > 0x7fffffffLL) return MCTQ_E_BADARG;
    }
    tile_starts[n_desc] = (int32_t)t;
    return t;
}

int mctq_fq_affine_multi(const MctqTensorDesc* descs_dev, const int32_t* tile_starts_dev, int n_desc,
                         int64_t total_tiles, void* stream) {
    MCTQ_NVTX("mctq_fq_affine_multi");
    if (!descs_dev || !tile_starts_dev || n_desc < 1 || total_tiles < 0 || total_tiles > 0x7fffffffLL) return MCTQ_E_BADARG;
    if (total_tiles == 0) return 0;
    return launch_streaming(fq_affine_multi_kernel, (unsigned)total_tiles, 0, (cudaStream_t)stream, descs_dev, tile_starts_dev, n_desc);
}

int mctq_fq_affine_scalar_multi(const MctqSiteDesc* sites, int n_sites, void* stream) {
    MCTQ_NVTX("mctq_fq_affine_scalar_multi");
    if (n_sites < 0 || (n_sites > 0 && !sites)) return MCTQ_E_BADARG;
    for (int k = 0; k < n_sites; ++k) {                     // validate everything before anything is launched
        const MctqSiteDesc& d = sites[k];
        if (d.n < 0 || (d.n > 0 && (!d.x || !d.y))) return MCTQ_E_BADARG;
        if (d.dtype != MCTQ_F32 && d.dtype != MCTQ_BF16 && d.dtype != MCTQ_F16) return MCTQ_E_DTYPE;
        if (d.qmin > d.qmax) return MCTQ_E_RANGE;
        if (d.n > 0 && (!aligned16(d.x) || !aligned16(d.y))) return MCTQ_E_BADARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    SitesParams p;
    int k = 0;
    while (k < n_sites) {
        memset(&p, 0, sizeof(p));
        int64_t tiles = 0;
        int m = 0;
        for (; k < n_sites && m < kSitesCap; ++k) {
            const MctqSiteDesc& d = sites[k];
            if (d.n == 0) continue;
            const int64_t tile_elems = (int64_t)kThreads * 2 * (d.dtype == MCTQ_F32 ? 4 : 8);
            const int64_t t = (d.n + tile_elems - 1) / tile_elems;
            if (tiles + t > 0x7fffffffLL) { if (m == 0) return MCTQ_E_BADARG; break; }       // next launch takes it
            SiteEntry& e = p.e[m];
            e.x = d.x; e.y = d.y; e.n = d.n; e.scale = d.scale; e.zp = d.zp; e.qmin = d.qmin; e.qmax = d.qmax; e.dtype = d.dtype;
            p.starts[m] = (int32_t)tiles;
            tiles += t;
            ++m;
        }
        if (m == 0) continue;
        p.n_sites = m;
        p.starts[m] = (int32_t)tiles;
        int rc = launch_streaming(fq_affine_sites_kernel, (unsigned)tiles, 0, st, p);
        if (rc) return rc;
    }
    return 0;
}

}  // extern "C"
