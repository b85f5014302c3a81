// mctq_host.cu -- process-wide state, introspection entry points and the host-buffer (staged) operators.
#include "mctq_common.cuh"

namespace mctq {
std::atomic<int64_t> g_launches{0};
int g_unroll = 0;           // 0 = automatic (per-tensor tiles: 2 vectors per thread, per-channel tiles: 4)
int g_force_rint = 0;
int g_force_ieee_div = 0;
int g_pdl = 1;
int g_lut_shfl = 1;
}  // namespace mctq

using namespace mctq;

extern "C" {

int mctq_abi_version(void) { return MCTQ_ABI_VERSION; }

const char* mctq_build_info(void) {
    return "libmctq_sm100 abi=1 arch=sm_100a threads=256 vec=16B unroll={auto,2,4,8} fmad=off";
}

int64_t mctq_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int mctq_set_tuning(int key, int value) {
    int prev;
    switch (key) {
        case 0: prev = g_unroll; if (value != 0 && value != 2 && value != 4 && value != 8) return MCTQ_E_BADARG; g_unroll = value; return prev;
        case 1: prev = g_force_rint; g_force_rint = value ? 1 : 0; return prev;
        case 2: prev = g_force_ieee_div; g_force_ieee_div = value ? 1 : 0; return prev;
        case 3: prev = g_pdl; g_pdl = value ? 1 : 0; return prev;
        case 4: prev = g_lut_shfl; g_lut_shfl = value ? 1 : 0; return prev;
        default: return MCTQ_E_BADARG;
    }
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------- host staging
namespace {
constexpr int kHostStreams = 3;
constexpr size_t kHostChunkBytesIn = 32u << 20;      // largest input chunk (slot size); small tensors use smaller chunks
struct HostCtx {
    int device = -1;
    cudaStream_t st[kHostStreams] = {nullptr, nullptr, nullptr};
    cudaEvent_t params_ready = nullptr;
};
HostCtx g_hctx[16];

int host_ctx(int device, HostCtx** out) {
    if (device < 0 || device >= 16) return MCTQ_E_NODEVICE;
    HostCtx& c = g_hctx[device];
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int)e;
    if (c.device != device) {
        for (int i = 0; i < kHostStreams; ++i) {
            e = cudaStreamCreateWithFlags(&c.st[i], cudaStreamNonBlocking);
            if (e != cudaSuccess) return (int)e;
        }
        e = cudaEventCreateWithFlags(&c.params_ready, cudaEventDisableTiming);
        if (e != cudaSuccess) return (int)e;
        c.device = device;
    }
    *out = &c;
    return 0;
}
size_t dtype_size(int dt) { return dt == MCTQ_F32 ? 4 : 2; }

// chunk length: about an eighth of the tensor so that uploads, kernels and downloads of neighbouring chunks overlap
// even for tensors of a few tens of MB, between 1 MB of input and the slot size, a multiple of 64 Ki elements
// (a sixteenth was measured and is slower: 67 MB 76 -> 71 GB/s, the per-transfer latency of the copy engines dominates)
int64_t pick_chunk_elems(int64_t n, size_t in_elem_bytes, size_t slot_elem_bytes) {
    const int64_t max_elems = (int64_t)(kHostChunkBytesIn / slot_elem_bytes);
    const int64_t min_elems = (int64_t)((1u << 20) / in_elem_bytes);
    int64_t c = (n / 8 + 65535) / 65536 * 65536;
    if (c < min_elems) c = min_elems;
    if (c > max_elems) c = max_elems;
    return c;
}

// Small tensors in PINNED host memory skip the staging pipeline altogether: the streaming kernel reads x and writes y
// directly over PCIe through their unified virtual addresses (one launch, one synchronise, no fill / drain of a
// three-stage pipeline).  Measured on the B200 box: 4 MB 0.215 -> 0.168 ms, 1 MB 0.055 ms; above ~16 MB the copy engines
// win (92 vs 80 GB/s at 1 GB), so the cut-over is 8 MB of input.
constexpr size_t kZeroCopyMaxBytesIn = 8u << 20;

bool is_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost && at.devicePointer == p;
}
bool zero_copy_ok(const void* x_host, const void* y_host, int64_t n, size_t in_es) {
    return (size_t)n * in_es <= kZeroCopyMaxBytesIn && is_pinned(x_host) && is_pinned(y_host);
}

// H2D -> launch(chunk) -> D2H for every chunk, round-robin over the internal streams; returns when y_host is complete.
// `launch(d_in, d_out, count, elem_offset, stream)` enqueues the kernel for one chunk.  `params_on_stream0`: the caller
// has enqueued parameter uploads on stream 0 that the other streams must wait for.
template <class Launch>
int run_pipeline(HostCtx* ctx, uint8_t* slots, const uint8_t* x_host, uint8_t* y_host, int64_t n, size_t in_es, size_t out_es,
                 bool params_on_stream0, Launch launch) {
    cudaError_t e = cudaSuccess;
    int rc = 0;
    const int64_t chunk_elems = pick_chunk_elems(n, in_es, in_es > out_es ? in_es : out_es);
    const int64_t n_chunks = (n + chunk_elems - 1) / chunk_elems;
    const int used = (int)(n_chunks < kHostStreams ? n_chunks : kHostStreams);
    if (params_on_stream0 && used > 1) {
        cudaEventRecord(ctx->params_ready, ctx->st[0]);
        for (int i = 1; i < used; ++i) cudaStreamWaitEvent(ctx->st[i], ctx->params_ready, 0);
    }
    int k = 0;
    for (int64_t off = 0; off < n; off += chunk_elems, ++k) {
        const int64_t cnt = (n - off) < chunk_elems ? (n - off) : chunk_elems;
        cudaStream_t st = ctx->st[k % kHostStreams];
        uint8_t* d_in = slots + (size_t)(k % kHostStreams) * 3 * kHostChunkBytesIn;
        uint8_t* d_out = d_in + kHostChunkBytesIn;
        e = cudaMemcpyAsync(d_in, x_host + off * in_es, cnt * in_es, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) break;
        rc = launch(d_in, d_out, cnt, off, st);
        if (rc) break;
        e = cudaMemcpyAsync(y_host + off * out_es, d_out, cnt * out_es, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) break;
    }
    for (int i = 0; i < kHostStreams; ++i) {
        if (i >= used && !(i == 0 && params_on_stream0)) continue;
        cudaError_t e2 = cudaStreamSynchronize(ctx->st[i]);
        if (e == cudaSuccess && e2 != cudaSuccess) e = e2;
    }
    if (rc) return rc;
    return cuda_rc(e);
}
}  // namespace

extern "C" {

size_t mctq_host_staging_min_bytes(void) {
    // per stream: one input chunk + one f32-sized output chunk (LUT output of a 2-byte input is 2x larger) + parameter area
    return kHostStreams * (kHostChunkBytesIn + 2 * kHostChunkBytesIn) + (4u << 20);
}

int mctq_fq_affine_host(const void* x_host, void* y_host, int64_t n, int x_dtype, const float* scale_host,
                        const int32_t* zp_host, int64_t C, int64_t inner, int32_t qmin, int32_t qmax,
                        void* staging_dev, size_t staging_bytes, int device) {
    if (!x_host || !y_host || !scale_host || !zp_host || !staging_dev || n < 0 || C < 1 || inner < 1) return MCTQ_E_BADARG;
    if (x_dtype < 0 || x_dtype > 2) return MCTQ_E_DTYPE;
    if (staging_bytes < mctq_host_staging_min_bytes() || (size_t)C * 8 > (4u << 20)) return MCTQ_E_BADARG;
    if (n == 0) return 0;
    HostCtx* ctx;
    int rc = host_ctx(device, &ctx);
    if (rc) return rc;
    const size_t es = dtype_size(x_dtype);
    uint8_t* base = reinterpret_cast<uint8_t*>(staging_dev);
    uint8_t* slots = base + (4u << 20);
    const bool zc = zero_copy_ok(x_host, y_host, n, es);
    if (C == 1) {
        // per-tensor: parameters travel by value, nothing to upload
        const float s = scale_host[0];
        const int32_t z = zp_host[0];
        if (zc) {
            rc = mctq_fq_affine_scalar(x_host, y_host, nullptr, n, x_dtype, s, z, qmin, qmax, MCTQ_CODES_NONE, ctx->st[0]);
            cudaError_t e2 = cudaStreamSynchronize(ctx->st[0]);
            return rc ? rc : cuda_rc(e2);
        }
        return run_pipeline(ctx, slots, reinterpret_cast<const uint8_t*>(x_host), reinterpret_cast<uint8_t*>(y_host), n, es, es, false,
                            [&](uint8_t* d_in, uint8_t* d_out, int64_t cnt, int64_t, cudaStream_t st) {
                                return mctq_fq_affine_scalar(d_in, d_out, nullptr, cnt, x_dtype, s, z, qmin, qmax, MCTQ_CODES_NONE, st);
                            });
    }
    float* d_scale = reinterpret_cast<float*>(base);
    int32_t* d_zp = reinterpret_cast<int32_t*>(base + (2u << 20));
    cudaError_t e = cudaMemcpyAsync(d_scale, scale_host, (size_t)C * 4, cudaMemcpyHostToDevice, ctx->st[0]);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpyAsync(d_zp, zp_host, (size_t)C * 4, cudaMemcpyHostToDevice, ctx->st[0]);
    if (e != cudaSuccess) return (int)e;
    if (zc) {
        rc = mctq_fq_affine(x_host, y_host, nullptr, n, x_dtype, d_scale, d_zp, C, inner, 0, qmin, qmax, MCTQ_CODES_NONE, ctx->st[0]);
        cudaError_t e2 = cudaStreamSynchronize(ctx->st[0]);
        return rc ? rc : cuda_rc(e2);
    }
    return run_pipeline(ctx, slots, reinterpret_cast<const uint8_t*>(x_host), reinterpret_cast<uint8_t*>(y_host), n, es, es, true,
                        [&](uint8_t* d_in, uint8_t* d_out, int64_t cnt, int64_t off, cudaStream_t st) {
                            return mctq_fq_affine(d_in, d_out, nullptr, cnt, x_dtype, d_scale, d_zp, C, inner, off, qmin, qmax,
                                                  MCTQ_CODES_NONE, st);
                        });
}

int mctq_fq_lut_host(const void* x_host, float* y_host, int64_t n, int x_dtype, const void* table_host, int K,
                     const float* thr_host, int64_t C, int64_t inner, float eps, int scalar_mode, float divisor,
                     float thr_f32, int round_to_x_dtype, void* staging_dev, size_t staging_bytes, int device) {
    if (!x_host || !y_host || !table_host || !staging_dev || n < 0 || C < 1 || inner < 1) return MCTQ_E_BADARG;
    if (!scalar_mode && !thr_host) return MCTQ_E_BADARG;
    if (x_dtype < 0 || x_dtype > 2) return MCTQ_E_DTYPE;
    const size_t tbytes = mctq_lut_table_bytes(K);
    if (!tbytes || tbytes > (1u << 20)) return MCTQ_E_LUT;
    if (staging_bytes < mctq_host_staging_min_bytes() || (size_t)C * 4 > (2u << 20)) return MCTQ_E_BADARG;
    if (n == 0) return 0;
    HostCtx* ctx;
    int rc = host_ctx(device, &ctx);
    if (rc) return rc;
    const size_t es = dtype_size(x_dtype);
    uint8_t* base = reinterpret_cast<uint8_t*>(staging_dev);
    float* d_thr = reinterpret_cast<float*>(base);
    uint8_t* d_table = base + (2u << 20);
    uint8_t* slots = base + (4u << 20);
    cudaError_t e = cudaSuccess;
    if (!scalar_mode) e = cudaMemcpyAsync(d_thr, thr_host, (size_t)C * 4, cudaMemcpyHostToDevice, ctx->st[0]);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpyAsync(d_table, table_host, tbytes, cudaMemcpyHostToDevice, ctx->st[0]);
    if (e != cudaSuccess) return (int)e;
    if (zero_copy_ok(x_host, y_host, n, es)) {
        if (scalar_mode)
            rc = mctq_fq_lut_scalar(x_host, y_host, nullptr, n, x_dtype, d_table, K, divisor, thr_f32, round_to_x_dtype, MCTQ_CODES_NONE, ctx->st[0]);
        else
            rc = mctq_fq_lut(x_host, y_host, nullptr, n, x_dtype, d_table, K, d_thr, C, inner, 0, eps, MCTQ_CODES_NONE, ctx->st[0]);
        cudaError_t e2 = cudaStreamSynchronize(ctx->st[0]);
        return rc ? rc : cuda_rc(e2);
    }
    return run_pipeline(ctx, slots, reinterpret_cast<const uint8_t*>(x_host), reinterpret_cast<uint8_t*>(y_host), n, es, 4, true,
                        [&](uint8_t* d_in, uint8_t* d_out, int64_t cnt, int64_t off, cudaStream_t st) {
                            float* out = reinterpret_cast<float*>(d_out);
                            if (scalar_mode)
                                return mctq_fq_lut_scalar(d_in, out, nullptr, cnt, x_dtype, d_table, K, divisor, thr_f32, round_to_x_dtype,
                                                          MCTQ_CODES_NONE, st);
                            return mctq_fq_lut(d_in, out, nullptr, cnt, x_dtype, d_table, K, d_thr, C, inner, off, eps, MCTQ_CODES_NONE, st);
                        });
}

}  // extern "C"
