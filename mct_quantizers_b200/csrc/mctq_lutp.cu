// mctq_lutp.cu -- "prepared" LUT fake-quant: per-channel decision tables in the x domain.
//
// The reference normalises every element (x / (thr + eps)), scales, clips and runs an argmin over the centroids
// (mct_quantizers/pytorch/quantizer_utils.py:95-170).  All of that is a monotone function of x per channel, so the
// result is decided by which of the K - 1 per-channel thresholds X[c][j] = sup{x : normalised(x) <= tau_j} the element
// exceeds.  mctq_lut_prepare computes those thresholds EXACTLY, once per quantizer, by bisection over f32 bit patterns
// through the reference's own arithmetic (IEEE division, optional rounding to bf16 / f16, first-minimum thresholds
// of the search table).  The dequantised output is (lut_sorted[pos] / 2^(bw-s)) * thr_c: the quotients are a small
// channel-independent table, the product is one multiply per element (a first version staged the products per channel,
// which doubled the record: for rows of 64 elements the tables were then more L2 traffic than the data).
// The hot kernel then needs no division and no search loop: an approximate cell index (one saturating FMA)
// selects the single threshold that can still matter, one exact compare decides, one shared-memory load fetches y.
#include <type_traits>
#include <vector>

#include "mctq_common.cuh"
#include "mctq_lut_table.cuh"

namespace mctq {

constexpr uint32_t kPrepMagic = 0x4d515050u;   // 'MQPP'
constexpr int kCellsPerUnit = 4;               // cell width = 1/4 of a step of the normalised grid
constexpr float kCellSlop = 0.02f;             // cells; >> the 1e-4 cell error of the approximate index
constexpr float kMagicRound = 12582912.0f;     // 1.5 * 2^23

struct LutPrepHeader {      // 64 bytes, start of the prepared blob (device memory)
    uint32_t magic;
    int32_t K, P, NC;       // centroids, padded table size, number of cells (power of two)
    int64_t C;
    float mult;
    int32_t round_dtype;    // 0 none, 1 bf16, 2 f16 (activation flavour with half-precision inputs)
    int32_t pos0;           // sorted position of original index 0 (NaN inputs)
    int32_t rec_floats;     // floats per channel record: X[P] s' thr d pad
    int32_t off_tau, off_cq, off_cells, off_orig, off_rec;   // byte offsets into the blob
    int32_t orig_identity;  // sorted position == original LUT index for every reachable position
};
static_assert(sizeof(LutPrepHeader) == 64, "header layout");

struct PrepGeom {
    int P, NC, rec_floats;
    size_t off_tau, off_cq, off_cells, off_orig, off_rec, bytes;
};

static int prep_geometry(int K, int bw, int is_signed, int64_t C, PrepGeom* g) {
    int L;
    if (lut_geometry_from_K(K, &g->P, &L)) return MCTQ_E_LUT;
    if (bw < 1 || bw > 16 || C < 1) return MCTQ_E_LUT;
    const int64_t mult = 1LL << (bw - (is_signed ? 1 : 0));
    int64_t need = kCellsPerUnit * 2 * mult + 8;
    int64_t nc = 64;
    while (nc < need) nc <<= 1;
    if (nc > 4096) return MCTQ_E_RANGE;           // table would not fit comfortably in shared memory
    g->NC = (int)nc;
    // everything the hot kernel stages with 16-byte bulk copies (cells | orig, channel records) starts on a 16-byte
    // boundary and is a multiple of 16 bytes long, also for tables of one or two entries
    g->rec_floats = (g->P + 4 + 3) & ~3;
    size_t o = sizeof(LutPrepHeader);
    g->off_tau = o; o += (size_t)g->P * 4;
    o = (o + 15) & ~(size_t)15;
    g->off_cq = o; o += ((size_t)g->P * 4 + 15) & ~(size_t)15;      // staged block starts here: [cq | cells | orig]
    g->off_cells = o; o += ((size_t)g->NC + 1 + 15) & ~(size_t)15;
    g->off_orig = o; o += ((size_t)g->P + 15) & ~(size_t)15;
    g->off_rec = o; o += (size_t)C * g->rec_floats * 4;
    g->bytes = o;
    return 0;
}

__host__ __device__ inline float round_like(float q, int round_dtype) {
    if (round_dtype == 1) return __bfloat162float(__float2bfloat16_rn(q));
    if (round_dtype == 2) return __half2float(__float2half_rn(q));
    return q;
}

__host__ __device__ inline int32_t f2ord(float f) {
    int32_t i = (int32_t)
#ifdef __CUDA_ARCH__
        __float_as_int(f);
#else
        [](float v) { int32_t r; memcpy(&r, &v, 4); return r; }(f);
#endif
    return i < 0 ? (int32_t)(0x80000000u - (uint32_t)i) : i;
}
__host__ __device__ inline float ord2f(int32_t o) {
    int32_t i = o < 0 ? (int32_t)(0x80000000u - (uint32_t)o) : o;
#ifdef __CUDA_ARCH__
    return __int_as_float(i);
#else
    float f; memcpy(&f, &i, 4); return f;
#endif
}

// ---- prepare kernel: one thread per (channel, table position)
__global__ void __launch_bounds__(kThreads) lut_prepare_kernel(uint8_t* blob, const float* thr, float eps, int scalar_mode,
                                                               float divisor, float thr_f32) {
    const LutPrepHeader h = *reinterpret_cast<const LutPrepHeader*>(blob);
    const float* tau = reinterpret_cast<const float*>(blob + h.off_tau);
    float* rec = reinterpret_cast<float*>(blob + h.off_rec);
    const int64_t total = h.C * h.P;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
        const int64_t c = i / h.P;
        const int j = (int)(i - c * h.P);
        float d, t;
        if (scalar_mode) { d = divisor; t = thr_f32; }
        else { t = thr[c]; d = __fadd_rn(t, eps); }
        float* r = rec + c * h.rec_floats;
        if (j == 0) {
            // approximate cell scale: u = x * s' + 0.5, cell = round(u * NC)
            r[h.P] = __fdiv_rn(__fmul_rn((float)kCellsPerUnit, h.mult), d) / (float)h.NC;
            r[h.P + 1] = t;                                                 // y = cq[pos] * thr_c (thr WITHOUT eps)
            r[h.P + 2] = d;
        }
        float X = INFINITY;
        if (j < h.P - 1) {
            const float tj = tau[j];
            if (tj == -INFINITY) X = -INFINITY;
            else if (tj != INFINITY) {
                // P(x) := round_like(x / d) > tau_j is monotone in x (d > 0); X = largest x for which it is false
                auto above = [&](float x) { return round_like(__fdiv_rn(x, d), h.round_dtype) > tj; };
                const float big = 3.4028234663852886e38f;
                if (above(-big)) X = -INFINITY;
                else if (!above(big)) X = big;
                else {
                    int64_t lo = f2ord(-big), hi = f2ord(big);
                    while (hi - lo > 1) {
                        int64_t mid = lo + (hi - lo) / 2;
                        if (above(ord2f((int32_t)mid))) hi = mid; else lo = mid;
                    }
                    X = ord2f((int32_t)lo);
                }
            }
        }
        r[j] = X;
    }
}

// ---- hot kernel
struct LutPArgs {
    const void* x;
    float* y;
    void* idx;
    int64_t n;
    const uint8_t* blob;
    int32_t P, NC, rec_floats;
    int32_t off_cq, off_cells, off_orig, off_rec;
    int64_t C, inner, elem_offset;
    FastDiv div_inner, div_W;
    uint32_t W, bigrow;
};

__device__ __forceinline__ float fma_sat(float a, float b, float c) {
    float r;
    asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// V: elements per vector.  4 everywhere (8-byte loads of 2-byte types, one 16-byte f32 store), or 8 for 2-byte types when
// the whole vector lies in one channel (CH_PT, CH_VEC with rows that are a multiple of 8): one 16-byte load, one 32-byte
// store (STG.256), half the per-vector bookkeeping per element -- the 2-byte kernels are instruction-issue bound, not HBM bound.
// `phase`: how many tiles this CTA has already staged through sm_bar.  The mbarrier is initialised ONCE per CTA (phase 0)
// and re-armed for every further tile; its parity alternates with the phase (re-initialising a live mbarrier is undefined).
template <typename T, int CHMODE, int CODE, int UNROLL, int V>
__device__ __forceinline__ void lutp_tile(const LutPArgs& a, const int64_t tile_index, const uint32_t phase = 0) {
    constexpr int WORDS_IN = V * sizeof(T) / 4;
    constexpr uint32_t TILE = kThreads * UNROLL * V;
    static_assert(V == 4 || (V == 8 && sizeof(T) == 2 && CHMODE != CH_ELEM), "vector width");
    extern __shared__ __align__(16) float sm_dyn[];                   // [records W * rec_floats][cq P (+pad)][cells NC + 1 (+pad)][orig P]
    __shared__ Window sm_win;
    __shared__ __align__(8) uint64_t sm_bar;

    const uint32_t tid = threadIdx.x;
    const int64_t t0 = tile_index * TILE;
    const int64_t remaining = a.n - t0;
    const bool full = remaining >= (int64_t)TILE;
    const T* xt = reinterpret_cast<const T*>(a.x) + t0;
    if (phase == 0) {
        pdl_wait();
        pdl_launch_dependents();
    }

    uint32_t w[UNROLL][WORDS_IN];
    if (full) {
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) ld_words<WORDS_IN>(xt + (size_t)(j * kThreads + tid) * V, w[j]);
    } else {
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) {
            int64_t l = (int64_t)(j * kThreads + tid) * V;
            if (l + V <= remaining) ld_words<WORDS_IN>(xt + l, w[j]);
            else {
                T tmp[V];
#pragma unroll
                for (int e = 0; e < V; ++e) tmp[e] = (l + e < remaining) ? xt[l + e] : from_f32<T>(0.0f);
                memcpy(w[j], tmp, sizeof(tmp));
            }
        }
    }

    // Stage the decision tables with the bulk-copy engine (1-D TMA): the blob already holds them in the layout the
    // CTA wants -- [cells | orig] contiguous and channel records back to back -- so one elected thread arms the
    // mbarrier and issues at most three cp.async.bulk copies (cells + orig, records c0 .. C-1, wrapped records 0 ..);
    // they complete while every thread's streaming loads of the data tile are in flight.
    const uint32_t Wn = CHMODE == CH_PT ? 1u : a.W;
    float* sm_rec = sm_dyn;
    float* sm_cq = sm_dyn + (size_t)Wn * a.rec_floats;
    uint8_t* sm_cells = reinterpret_cast<uint8_t*>(sm_cq) + (a.off_cells - a.off_cq);
    uint8_t* sm_orig = sm_cells + ((a.NC + 1 + 15) & ~15);
    if (tid == 0) {
        if (phase == 0) mbar_init(&sm_bar, 1);
        uint32_t c0 = 0;
        if (CHMODE != CH_PT) {
            const int64_t g0 = a.elem_offset + t0;
            const int64_t r0 = g0 / a.inner;
            const int64_t off = g0 - r0 * a.inner;
            Window wv;
            wv.off0 = a.bigrow ? 0u : (uint32_t)off;
            const int64_t sp = a.inner - off;
            wv.split = (uint32_t)(sp > (int64_t)TILE ? (int64_t)TILE + 1 : sp);
            sm_win = wv;
            c0 = (uint32_t)(r0 % a.C);
        }
        const uint32_t rec_bytes = (uint32_t)a.rec_floats * 4u;
        const uint32_t tab_bytes = (uint32_t)(a.off_rec - a.off_cq);
        const uint32_t n1 = min(Wn, (uint32_t)a.C - c0);            // records before the channel index wraps
        mbar_arrive_expect_tx(&sm_bar, tab_bytes + Wn * rec_bytes);
        bulk_g2s(sm_cq, a.blob + a.off_cq, tab_bytes, &sm_bar);
        bulk_g2s(sm_rec, a.blob + a.off_rec + (size_t)c0 * rec_bytes, n1 * rec_bytes, &sm_bar);
        if (n1 < Wn) bulk_g2s(reinterpret_cast<char*>(sm_rec) + (size_t)n1 * rec_bytes, a.blob + a.off_rec, (Wn - n1) * rec_bytes, &sm_bar);
    }
    __syncthreads();                                                  // barrier init + window visible to everyone
    Window win;
    if (CHMODE != CH_PT) win = sm_win;
    mbar_wait(&sm_bar, phase & 1u);

    // Everything below addresses shared memory through 32-bit shared-window addresses: the element loop is
    //   u = sat(x * s' + 0.5); cell = low bits of fma(u, NC, 1.5 * 2^23); b = cells[cell]          (candidate threshold)
    //   px = rec + 4 b; X = [px]; pos = b + (x > X); y = cq[pos] * thr_c                           (record = X[P] | s' | thr_c ...)
    const float NCf = (float)a.NC;
    const uint32_t rec_bytes = (uint32_t)a.rec_floats * 4u;
    const uint32_t xbytes = (uint32_t)a.P * 4u;                       // s' sits right behind the P thresholds, thr_c behind it
    const uint32_t rec_base = smem_u32(sm_rec);
    const uint32_t cq_s = smem_u32(sm_cq);
    const uint32_t cells_s = smem_u32(sm_cells);
    const uint32_t orig_s = smem_u32(sm_orig);
    float* yt = a.y + t0;
    // Index emission: when the centroid list is already sorted and free of duplicates the sorted position IS the LUT index
    // (flag in the blob header) and the look-up of the original index disappears from the element loop.
    auto run = [&](auto ident_tag) {
        constexpr bool IDENT = decltype(ident_tag)::value;
    #pragma unroll
        for (int j = 0; j < UNROLL; ++j) {
            const uint32_t l = (uint32_t)(j * kThreads + tid) * V;
            float f[V];
            int code[V];
            Pack<T, V>::unpack(w[j], f);
            uint32_t slot = 0, rem = 0;
            if (CHMODE != CH_PT) locate(l, win, a, slot, rem);
            uint32_t rec = rec_base + slot * rec_bytes;
            float sp = lds_f32(rec + xbytes);
            float th = lds_f32(rec + xbytes + 4u);
            float nan_probe = f[0] + f[1];                             // NaN iff some element is NaN (or inf - inf)
    #pragma unroll
            for (int e = 2; e < V; e += 2) nan_probe += f[e] + f[e + 1];
            // CH_ELEM: a vector of 4 straddles row boundaries.  Rows of at least 4 elements: at most one boundary, so the
            // record pointer / cell scale of the second row are selected per element; shorter rows advance per element.
            const bool two_rows = CHMODE == CH_ELEM && a.inner >= V;
            uint32_t k = V;                                             // elements of this vector in the first row
            uint32_t rec1 = rec;
            float sp1 = sp, th1 = th;
            if (two_rows) {
                if (a.bigrow) {
                    const bool second = l >= win.split;
                    slot = second ? 1u : 0u;
                    rec = rec_base + slot * rec_bytes;
                    sp = lds_f32(rec + xbytes);
                    th = lds_f32(rec + xbytes + 4u);
                    k = second ? (uint32_t)V : min((uint32_t)V, win.split - l);
                } else {
                    k = a.div_inner.d - rem;
                }
                const uint32_t slot1 = (slot + 1 == a.W) ? 0u : slot + 1;
                rec1 = rec_base + slot1 * rec_bytes;
                sp1 = lds_f32(rec1 + xbytes);
                th1 = lds_f32(rec1 + xbytes + 4u);
            }
    #pragma unroll
            for (int e = 0; e < V; ++e) {
                uint32_t r = rec;
                float s = sp, t = th;
                if (CHMODE == CH_ELEM) {
                    if (two_rows) {
                        const bool first = (uint32_t)e < k;
                        r = first ? rec : rec1;
                        s = first ? sp : sp1;
                        t = first ? th : th1;
                    } else {
                        r = rec_base + slot * rec_bytes;
                        s = lds_f32(r + xbytes);
                        t = lds_f32(r + xbytes + 4u);
                    }
                }
                const float x = f[e];
                const float u = fma_sat(x, s, 0.5f);                        // saturates to [0, 1]; NaN -> 0
                const float cf = __fmaf_rn(u, NCf, kMagicRound);            // integer cell index in the low mantissa bits
                const uint32_t cell = __float_as_uint(cf) & 0x1fffu;
                const uint32_t b = lds_u8(cells_s + cell);                  // index of the candidate threshold
                const uint32_t b4 = b << 2;
                const float X = lds_f32(r + b4);
                const bool above = x > X;
                f[e] = __fmul_rn(lds_f32(cq_s + b4 + (above ? 4u : 0u)), t);     // 16 consecutive words: never a bank conflict
                if (CODE != 0) code[e] = IDENT ? (int)(b + (above ? 1u : 0u)) : (int)lds_u8(orig_s + b + (above ? 1u : 0u));
                if (CHMODE == CH_ELEM && !two_rows) {
                    if (++rem == a.div_inner.d) { rem = 0; slot = (slot + 1 == a.W) ? 0 : slot + 1; }
                }
            }
            if (nan_probe != nan_probe) {
                // rare: torch.argmin over all-NaN distances returns index 0
                const uint32_t pos0 = (uint32_t)__ldg(&reinterpret_cast<const LutPrepHeader*>(a.blob)->pos0);
                float g[V];
                Pack<T, V>::unpack(w[j], g);
                uint32_t slot2 = 0, rem2 = 0;
                if (CHMODE != CH_PT) locate(l, win, a, slot2, rem2);
                for (int e = 0; e < V; ++e) {
                    if (CHMODE == CH_ELEM && a.bigrow) {
                        uint32_t jrow = (l + e) >= win.split ? 1u : 0u;
                        slot2 = jrow >= a.W ? jrow - a.W : jrow;
                    }
                    if (g[e] != g[e]) {
                        f[e] = __fmul_rn(lds_f32(cq_s + 4u * pos0), lds_f32(rec_base + slot2 * rec_bytes + xbytes + 4u));
                        if (CODE != 0) code[e] = (int)lds_u8(orig_s + pos0);
                    }
                    if (CHMODE == CH_ELEM && !a.bigrow) {
                        if (++rem2 == a.div_inner.d) { rem2 = 0; slot2 = (slot2 + 1 == a.W) ? 0 : slot2 + 1; }
                    }
                }
            }
            if (full || (int64_t)l + V <= remaining) {
                if (a.y) {
                    uint32_t o[V];
                    Pack<float, V>::pack(f, o);
                    if (V == 8) st_stream256(yt + l, o);        // 32 contiguous bytes per lane in one store
                    else st_words<4>(yt + l, o);
                }
                if (CODE != 0) st_codes<V, CODE>(a.idx, t0 + l, code);
            } else if ((int64_t)l < remaining) {
                const int cnt = (int)(remaining - l);
                for (int e = 0; e < V; ++e) {
                    if (e < cnt) {
                        if (a.y) yt[l + e] = f[e];
                        if (CODE == MCTQ_CODES_INT8) reinterpret_cast<uint8_t*>(a.idx)[t0 + l + e] = (uint8_t)code[e];
                    }
                }
                if (CODE == MCTQ_CODES_INT4) {
                    uint8_t* cp = reinterpret_cast<uint8_t*>(a.idx) + ((t0 + l) >> 1);
                    for (int e = 0; e < V; e += 2) {
                        if (e < cnt) {
                            int hi = (e + 1 < cnt) ? code[e + 1] : 0;
                            cp[e >> 1] = (uint8_t)((code[e] & 0xf) | ((hi & 0xf) << 4));
                        }
                    }
                }
            }
        }
    };
    if (CODE != 0 && __ldg(&reinterpret_cast<const LutPrepHeader*>(a.blob)->orig_identity) != 0) run(std::true_type{});
    else run(std::false_type{});
}

template <typename T, int CHMODE, int CODE, int UNROLL, int V>
__global__ void __launch_bounds__(kThreads) fq_lutp_kernel(const __grid_constant__ LutPArgs a) {
    lutp_tile<T, CHMODE, CODE, UNROLL, V>(a, (int64_t)blockIdx.x);
}

// ---- many tensors, one launch (whole-model LUT weight quantization).  The per-tensor argument blocks and the first tile
// of every tensor travel as KERNEL PARAMETERS (a __grid_constant__ struct of up to 32 KB): a CTA finds its tensor by a binary
// search over the constant bank and reads that tensor's arguments from the constant bank with register-indexed LDC -- a
// few tens of cycles, no global loads and no barrier before the data loads can be issued (a first version kept the table
// in global memory: the dependent look-ups cost every CTA ~1000 cycles of idle residency and 15-20 % of the bandwidth).
// No launch gaps and no per-launch tails between the tensors.
constexpr int kMultiMaxDesc = 180;
struct alignas(16) LutPMultiEntry {
    LutPArgs a;
    int32_t dtype, chmode, v, pad;
};
// The parameter block is copied by every launch (~20 us for 27 KB), so its capacity is a template parameter: 16 / 64 / 180
// tensors (2.4 / 9.5 / 26.7 KB); the host plan keeps full-capacity chunks and the launch copies what is used.
template <int CAP>
struct LutPMultiParamsT {
    int32_t n_desc, span, pad[2];
    int32_t starts[CAP + 4];                    // starts[k] = first span of tensor k; starts[n_desc] = number of spans
    LutPMultiEntry e[CAP];
};
using LutPMultiParams = LutPMultiParamsT<kMultiMaxDesc>;
static_assert(sizeof(LutPMultiParams) <= 32764, "kernel parameter space");

// SPAN consecutive tiles of one tensor per CTA
template <typename T, int CHMODE, int V, int SPAN>
__device__ __forceinline__ void lutp_span(const LutPArgs& a, int64_t span) {
    constexpr int64_t TILE = (int64_t)kThreads * 4 * V;
    const int64_t first = span * SPAN;
#pragma unroll 1
    for (int g = 0; g < SPAN; ++g) {
        if (g && (first + g) * TILE >= a.n) break;
        if (g) __syncthreads();                  // everybody is done with the staged tables and has left the previous phase's wait
        lutp_tile<T, CHMODE, MCTQ_CODES_NONE, 4, V>(a, first + g, (uint32_t)g);
    }
}

template <typename T, int SPAN>
__device__ __forceinline__ void lutp_multi_dispatch(const LutPMultiEntry& e, int64_t span) {
    constexpr int V8 = sizeof(T) == 2 ? 8 : 4;
    if (e.chmode == CH_PT) {
        if (e.v == 8) lutp_span<T, CH_PT, V8, SPAN>(e.a, span);
        else lutp_span<T, CH_PT, 4, SPAN>(e.a, span);
    } else if (e.chmode == CH_VEC) {
        if (e.v == 8) lutp_span<T, CH_VEC, V8, SPAN>(e.a, span);
        else lutp_span<T, CH_VEC, 4, SPAN>(e.a, span);
    } else {
        lutp_span<T, CH_ELEM, 4, SPAN>(e.a, span);
    }
}

template <int SPAN, int CAP>
__global__ void __launch_bounds__(kThreads, (SPAN > 1 ? 4 : 1)) fq_lutp_multi_kernel(const __grid_constant__ LutPMultiParamsT<CAP> p) {
    const int tile = blockIdx.x;
    int lo = 0, hi = p.n_desc;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (p.starts[mid] <= tile) lo = mid; else hi = mid;
    }
    const LutPMultiEntry& e = p.e[lo];
    const int64_t t = (int64_t)(tile - p.starts[lo]);
    if (e.dtype == MCTQ_F32) lutp_multi_dispatch<float, SPAN>(e, t);
    else if (e.dtype == MCTQ_BF16) lutp_multi_dispatch<__nv_bfloat16, SPAN>(e, t);
    else lutp_multi_dispatch<__half, SPAN>(e, t);
}

}  // namespace mctq

using namespace mctq;

namespace {

// variant selection shared by the single-tensor and the multi-tensor entry points
void lutp_variant(const LutPArgs& a, size_t esz, int idx_mode, int* chmode, int* v) {
    // 8-element vectors for 2-byte inputs: 16-byte aligned x, 32-byte aligned y, 8-byte aligned int8 indices, rows a multiple of 8
    bool v8 = g_wide && esz == 2 && aligned16(a.x) && (!a.y || aligned32(a.y)) &&
              (idx_mode != MCTQ_CODES_INT8 || (reinterpret_cast<uintptr_t>(a.idx) & 7u) == 0);
    if (a.C == 1) *chmode = CH_PT;
    else if (a.inner % 4 == 0 && a.elem_offset % 4 == 0) {
        *chmode = CH_VEC;
        v8 = v8 && a.inner % 8 == 0 && a.elem_offset % 8 == 0;
    } else {
        *chmode = CH_ELEM;
        v8 = false;
    }
    *v = v8 ? 8 : 4;
}

// channel-window geometry and dynamic shared memory of one tensor for tiles of kThreads * 4 * v elements
int lutp_finish_args(LutPArgs& a, int chmode, int v, size_t* smem_out) {
    const uint32_t tile = kThreads * 4 * (uint32_t)v;
    uint32_t W = 1;
    if (chmode != CH_PT) { set_window(a, tile); W = a.W; }
    const size_t smem = (size_t)W * a.rec_floats * 4 + (size_t)(a.off_rec - a.off_cq);      // records + [cq | cells | orig]
    if (smem > 64 * 1024) return MCTQ_E_RANGE;            // caller falls back to the generic kernel
    *smem_out = smem;
    return 0;
}

template <typename T, int CHMODE, int CODE, int V>
int launch_lutp_tiles(const LutPArgs& a_in, cudaStream_t st) {
    constexpr int UNROLL = 4;
    constexpr uint32_t TILE = kThreads * UNROLL * V;
    LutPArgs a = a_in;
    size_t smem = 0;
    int rc = lutp_finish_args(a, CHMODE, V, &smem);
    if (rc) return rc;
    rc = ensure_smem(fq_lutp_kernel<T, CHMODE, CODE, UNROLL, V>, smem);
    if (rc) return rc;
    int64_t tiles = (a.n + TILE - 1) / TILE;
    if (tiles > 0x7fffffffLL) return MCTQ_E_BADARG;
    return launch_streaming(fq_lutp_kernel<T, CHMODE, CODE, UNROLL, V>, (unsigned)tiles, smem, st, a);
}

template <typename T, int CHMODE, int V>
int launch_lutp_code(const LutPArgs& a, int idx_mode, cudaStream_t st) {
    switch (idx_mode) {
        case MCTQ_CODES_INT8: return launch_lutp_tiles<T, CHMODE, MCTQ_CODES_INT8, V>(a, st);
        case MCTQ_CODES_INT4: return launch_lutp_tiles<T, CHMODE, MCTQ_CODES_INT4, V>(a, st);
        default: return launch_lutp_tiles<T, CHMODE, MCTQ_CODES_NONE, V>(a, st);
    }
}

template <typename T>
int launch_lutp_typed(const LutPArgs& a, int idx_mode, cudaStream_t st) {
    constexpr int V8 = sizeof(T) == 2 ? 8 : 4;
    int chmode, v;
    lutp_variant(a, sizeof(T), idx_mode, &chmode, &v);
    if (chmode == CH_PT) return v == 8 ? launch_lutp_code<T, CH_PT, V8>(a, idx_mode, st) : launch_lutp_code<T, CH_PT, 4>(a, idx_mode, st);
    if (chmode == CH_VEC) return v == 8 ? launch_lutp_code<T, CH_VEC, V8>(a, idx_mode, st) : launch_lutp_code<T, CH_VEC, 4>(a, idx_mode, st);
    return launch_lutp_code<T, CH_ELEM, 4>(a, idx_mode, st);
}

// argument block of one prepared-LUT tensor (validated); shared by mctq_fq_lut_prepared and the multi-tensor plan
int lutp_make_args(const void* x, float* y, void* idx, int64_t n, int x_dtype, const void* prepared_dev, int K, int bw, int is_signed,
                   int64_t C, int64_t inner, int64_t elem_offset, int idx_mode, LutPArgs* out) {
    if (!x || !prepared_dev || n < 0 || C < 1 || inner < 1 || elem_offset < 0 || (!y && idx_mode == MCTQ_CODES_NONE)) return MCTQ_E_BADARG;
    if (idx_mode != MCTQ_CODES_NONE && !idx) return MCTQ_E_BADARG;
    if (x_dtype < 0 || x_dtype > 2) return MCTQ_E_DTYPE;
    PrepGeom g;
    int rc = prep_geometry(K, bw, is_signed, C, &g);
    if (rc) return rc;
    if (idx_mode == MCTQ_CODES_INT4 && g.P > 16) return MCTQ_E_RANGE;
    const size_t esz = x_dtype == MCTQ_F32 ? 4 : 2;
    bool vec_ok = (reinterpret_cast<uintptr_t>(x) % (4 * esz)) == 0 && (!y || aligned16(y));
    if (idx_mode != MCTQ_CODES_NONE) vec_ok = vec_ok && (reinterpret_cast<uintptr_t>(idx) & 3u) == 0;
    if (!aligned16(prepared_dev)) vec_ok = false;  // the tables are staged with 16-byte bulk copies
    if (!vec_ok) return MCTQ_E_BADARG;             // caller uses the generic entry point for misaligned views
    LutPArgs a;
    memset(&a, 0, sizeof(a));
    a.x = x; a.y = y; a.idx = idx; a.n = n; a.blob = reinterpret_cast<const uint8_t*>(prepared_dev);
    a.P = g.P; a.NC = g.NC; a.rec_floats = g.rec_floats;
    a.off_cq = (int32_t)g.off_cq; a.off_cells = (int32_t)g.off_cells; a.off_orig = (int32_t)g.off_orig; a.off_rec = (int32_t)g.off_rec;
    a.C = C; a.inner = C == 1 ? 1 : inner; a.elem_offset = C == 1 ? 0 : elem_offset;
    *out = a;
    return 0;
}

}  // namespace

extern "C" {

size_t mctq_lut_prepared_bytes(int K, int lut_values_bitwidth, int is_signed, int64_t C) {
    PrepGeom g;
    if (prep_geometry(K, lut_values_bitwidth, is_signed, C, &g)) return 0;
    return g.bytes;
}

int mctq_lut_prepare(const void* table_host, int K, const float* thr_dev, int64_t C, float eps, int scalar_mode,
                     float divisor, float thr_f32, int round_dtype, void* prepared_dev, size_t prepared_bytes, void* stream) {
    if (!table_host || !prepared_dev || C < 1 || (!scalar_mode && !thr_dev) || round_dtype < 0 || round_dtype > 2) return MCTQ_E_BADARG;
    if (scalar_mode && C != 1) return MCTQ_E_BADARG;
    const LutTableHeader* th = reinterpret_cast<const LutTableHeader*>(table_host);
    if (th->magic != kLutMagic || th->K != K) return MCTQ_E_LUT;
    PrepGeom g;
    int rc = prep_geometry(K, th->bw, th->is_signed, C, &g);
    if (rc) return rc;
    if (prepared_bytes < g.bytes) return MCTQ_E_BADARG;
    const int P = g.P, NC = g.NC;
    const float* tau = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(table_host) + sizeof(LutTableHeader));
    const float* cq = tau + (P - 1);
    const uint8_t* orig = reinterpret_cast<const uint8_t*>(cq + P);
    // channel-independent front of the blob, assembled on the host
    std::vector<uint8_t> front(g.off_rec, 0);
    LutPrepHeader* h = reinterpret_cast<LutPrepHeader*>(front.data());
    h->magic = kPrepMagic; h->K = K; h->P = P; h->NC = NC; h->C = C; h->mult = th->mult; h->round_dtype = round_dtype;
    h->pos0 = th->pos_of_idx0; h->rec_floats = g.rec_floats;
    h->orig_identity = 1;
    for (int i = 0; i < th->Ks; ++i) if (orig[i] != i) h->orig_identity = 0;
    h->off_tau = (int32_t)g.off_tau; h->off_cq = (int32_t)g.off_cq; h->off_cells = (int32_t)g.off_cells;
    h->off_orig = (int32_t)g.off_orig; h->off_rec = (int32_t)g.off_rec;
    float* ftau = reinterpret_cast<float*>(front.data() + g.off_tau);
    float* fcq = reinterpret_cast<float*>(front.data() + g.off_cq);
    for (int j = 0; j < P; ++j) { ftau[j] = j < P - 1 ? tau[j] : INFINITY; fcq[j] = cq[j]; }
    memcpy(front.data() + g.off_orig, orig, P);
    // effective thresholds in the unrounded q domain (rounding to bf16 / f16 moves them to the rounding boundary)
    std::vector<double> V(P - 1 > 0 ? P - 1 : 0);
    for (int j = 0; j < P - 1; ++j) {
        float tj = tau[j];
        double e;
        if (tj == INFINITY) e = INFINITY;
        else if (tj == -INFINITY) e = -INFINITY;
        else if (round_dtype == 0) e = tj;
        else {
            // largest f32 q with round_like(q) <= tau_j
            int64_t lo = f2ord(-3.0e38f), hi = f2ord(3.0e38f);
            if (round_like(ord2f((int32_t)lo), round_dtype) > tj) e = -INFINITY;
            else if (!(round_like(ord2f((int32_t)hi), round_dtype) > tj)) e = INFINITY;
            else {
                while (hi - lo > 1) {
                    int64_t mid = lo + (hi - lo) / 2;
                    if (round_like(ord2f((int32_t)mid), round_dtype) > tj) hi = mid; else lo = mid;
                }
                e = ord2f((int32_t)lo);
            }
        }
        V[j] = e * (double)kCellsPerUnit * (double)th->mult + 0.5 * NC;      // position in cell units
    }
    uint8_t* cells = front.data() + g.off_cells;
    for (int k = 0; k <= NC; ++k) {
        int b = 0;
        for (int j = 0; j < P - 1; ++j) if (V[j] < (double)k - 0.5 - kCellSlop) ++b;
        cells[k] = (uint8_t)b;
    }
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemcpyAsync(prepared_dev, front.data(), g.off_rec, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return (int)e;
    e = cudaStreamSynchronize(st);                 // `front` is pageable and goes out of scope; this call is one-off setup
    if (e != cudaSuccess) return (int)e;
    int64_t total = C * P;
    int64_t blocks = (total + kThreads - 1) / kThreads;
    if (blocks > 148 * 16) blocks = 148 * 16;
    lut_prepare_kernel<<<(unsigned)blocks, kThreads, 0, st>>>(reinterpret_cast<uint8_t*>(prepared_dev), thr_dev, eps, scalar_mode, divisor, thr_f32);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cuda_rc(cudaGetLastError());
}

int mctq_fq_lut_prepared(const void* x, float* y, void* idx, int64_t n, int x_dtype, const void* prepared_dev, int K,
                         int lut_values_bitwidth, int is_signed, int64_t C, int64_t inner, int64_t elem_offset,
                         int idx_mode, void* stream) {
    LutPArgs a;
    int rc = lutp_make_args(x, y, idx, n, x_dtype, prepared_dev, K, lut_values_bitwidth, is_signed, C, inner, elem_offset, idx_mode, &a);
    if (rc) return rc;
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    switch (x_dtype) {
        case MCTQ_F32: return launch_lutp_typed<float>(a, idx_mode, st);
        case MCTQ_BF16: return launch_lutp_typed<__nv_bfloat16>(a, idx_mode, st);
        default: return launch_lutp_typed<__half>(a, idx_mode, st);
    }
}

// ---- multi-tensor plan: a host blob [header][LutPMultiParams x n_chunks]; every chunk is one launch of up to kMultiMaxDesc tensors
}  // extern "C"

namespace {
constexpr uint32_t kMultiMagic = 0x4d514c4du;   // 'MQLM'
struct LutPMultiHeader {     // 64 bytes
    uint32_t magic;
    int32_t n_desc, n_chunks, span;
    int64_t total_spans;
    uint64_t smem_bytes;
    int32_t reserved[8];
};
static_assert(sizeof(LutPMultiHeader) == 64, "multi header layout");

template <int SPAN, int CAP>
int launch_multi_chunk(const LutPMultiParams& c, size_t smem, cudaStream_t st) {
    static thread_local LutPMultiParamsT<CAP> p;             // up to 27 KB: not on the stack
    p.n_desc = c.n_desc;
    p.span = c.span;
    memcpy(p.starts, c.starts, ((size_t)c.n_desc + 1) * sizeof(int32_t));
    memcpy(p.e, c.e, (size_t)c.n_desc * sizeof(LutPMultiEntry));
    int rc = ensure_smem(fq_lutp_multi_kernel<SPAN, CAP>, smem);
    if (rc) return rc;
    return launch_streaming(fq_lutp_multi_kernel<SPAN, CAP>, (unsigned)c.starts[c.n_desc], smem, st, p);
}

size_t multi_bytes(int n_desc) {
    const int chunks = (n_desc + kMultiMaxDesc - 1) / kMultiMaxDesc;
    return sizeof(LutPMultiHeader) + (size_t)chunks * sizeof(LutPMultiParams);
}

// validates every tensor; fills the chunks when `blob` is given; returns the total number of spans (or < 0)
int64_t multi_compile(const MctqLutTensorDesc* descs, int n_desc, uint8_t* blob) {
    if (!descs || n_desc < 1) return MCTQ_E_BADARG;
    const int span = g_multi_span == 1 ? 1 : 4;          // measured on Llama-7B shapes: 4 tiles per CTA +5 % (f32) / +2 % (bf16) over 1
    LutPMultiHeader* h = reinterpret_cast<LutPMultiHeader*>(blob);
    LutPMultiParams* chunks = blob ? reinterpret_cast<LutPMultiParams*>(blob + sizeof(LutPMultiHeader)) : nullptr;
    int64_t total = 0;
    size_t smem_max = 0;
    for (int k = 0; k < n_desc; ++k) {
        const MctqLutTensorDesc& d = descs[k];
        if (d.n < 1 || !d.y) return MCTQ_E_BADARG;         // empty tensors do not belong in a plan
        LutPMultiEntry e;
        memset(&e, 0, sizeof(e));
        int rc = lutp_make_args(d.x, d.y, nullptr, d.n, d.dtype, d.prepared_dev, d.K, d.lut_values_bitwidth, d.is_signed, d.C, d.inner,
                                0, MCTQ_CODES_NONE, &e.a);
        if (rc) return rc;
        lutp_variant(e.a, d.dtype == MCTQ_F32 ? 4 : 2, MCTQ_CODES_NONE, &e.chmode, &e.v);
        size_t smem = 0;
        rc = lutp_finish_args(e.a, e.chmode, e.v, &smem);
        if (rc) return rc;
        if (smem > smem_max) smem_max = smem;
        e.dtype = d.dtype;
        const int64_t per_span = (int64_t)kThreads * 4 * e.v * span;
        const int64_t spans = (d.n + per_span - 1) / per_span;
        if (spans > 0x3fffffffLL) return MCTQ_E_BADARG;
        if (chunks) {
            LutPMultiParams& c = chunks[k / kMultiMaxDesc];
            const int j = k % kMultiMaxDesc;
            if (j == 0) { c.n_desc = 0; c.span = span; c.starts[0] = 0; }
            if ((int64_t)c.starts[j] + spans > 0x7fffffffLL) return MCTQ_E_BADARG;
            c.e[j] = e;
            c.starts[j + 1] = c.starts[j] + (int32_t)spans;
            c.n_desc = j + 1;
        }
        total += spans;
    }
    if (h) {
        h->magic = kMultiMagic;
        h->n_desc = n_desc;
        h->n_chunks = (n_desc + kMultiMaxDesc - 1) / kMultiMaxDesc;
        h->span = span;
        h->total_spans = total;
        h->smem_bytes = smem_max;
    }
    return total;
}
}  // namespace

extern "C" {

size_t mctq_lut_multi_plan_bytes(const MctqLutTensorDesc* descs, int n_desc) {
    if (multi_compile(descs, n_desc, nullptr) < 0) return 0;
    return multi_bytes(n_desc);
}

int64_t mctq_lut_multi_plan(const MctqLutTensorDesc* descs, int n_desc, void* plan_host_out, size_t plan_bytes) {
    int64_t t = multi_compile(descs, n_desc, nullptr);
    if (t < 0) return t;
    if (!plan_host_out || plan_bytes < multi_bytes(n_desc)) return MCTQ_E_BADARG;
    memset(plan_host_out, 0, multi_bytes(n_desc));
    return multi_compile(descs, n_desc, reinterpret_cast<uint8_t*>(plan_host_out));
}

int mctq_fq_lut_prepared_multi(const void* plan_host, void* stream) {
    if (!plan_host) return MCTQ_E_BADARG;
    const LutPMultiHeader* h = reinterpret_cast<const LutPMultiHeader*>(plan_host);
    if (h->magic != kMultiMagic || h->n_desc < 1 || h->n_chunks < 1) return MCTQ_E_BADARG;
    const LutPMultiParams* chunks = reinterpret_cast<const LutPMultiParams*>(reinterpret_cast<const uint8_t*>(plan_host) + sizeof(LutPMultiHeader));
    const size_t smem = (size_t)h->smem_bytes;
    for (int c = 0; c < h->n_chunks; ++c) {
        const LutPMultiParams& p = chunks[c];
        int rc;
        if (p.n_desc <= 16) rc = p.span == 1 ? launch_multi_chunk<1, 16>(p, smem, (cudaStream_t)stream) : launch_multi_chunk<4, 16>(p, smem, (cudaStream_t)stream);
        else if (p.n_desc <= 64) rc = p.span == 1 ? launch_multi_chunk<1, 64>(p, smem, (cudaStream_t)stream) : launch_multi_chunk<4, 64>(p, smem, (cudaStream_t)stream);
        else rc = p.span == 1 ? launch_multi_chunk<1, kMultiMaxDesc>(p, smem, (cudaStream_t)stream) : launch_multi_chunk<4, kMultiMaxDesc>(p, smem, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return 0;
}

}  // extern "C"
