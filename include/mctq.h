/*
 * mctq.h -- C ABI of libmctq_sm100.so, the B200 (sm_100a) fake-quant path that replaces the arithmetic
 * underneath mct_quantizers' PyTorch inferable quantizers.
 *
 * The reference (sony/mct_quantizers 1.6.0) is pure Python and has no FFI of its own: its hot path is
 *   - torch.fake_quantize_per_tensor_affine / torch.fake_quantize_per_channel_affine, called from
 *       mct_quantizers/pytorch/quantizers/weights_inferable_quantizers/weights_symmetric_inferable_quantizer.py:139-151
 *       mct_quantizers/pytorch/quantizers/weights_inferable_quantizers/weights_uniform_inferable_quantizer.py:153-165
 *       mct_quantizers/pytorch/quantizers/activation_inferable_quantizers/activation_symmetric_inferable_quantizer.py:113-117
 *       mct_quantizers/pytorch/quantizers/activation_inferable_quantizers/activation_uniform_inferable_quantizer.py:124-128
 *   - lut_quantizer / int_quantization_with_threshold, mct_quantizers/pytorch/quantizer_utils.py:95-170, called from
 *       .../weights_lut_symmetric_inferable_quantizer.py:114-122 and .../activation_lut_pot_inferable_quantizer.py:86-91
 * Each entry point below names the reference call site(s) it stands in for.  INTEGRATION.md shows the
 * ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success, a positive cudaError_t on a CUDA failure, or a negative
 *     MCTQ_E_* code on a bad argument; nothing throws, nothing allocates device memory, and the
 *     device-pointer entry points never synchronise: work is enqueued on `stream` (a cudaStream_t
 *     passed as void*; NULL = legacy default stream).
 *   - tensors are contiguous and are viewed as [outer][C][inner]; element i of the logical tensor has
 *     channel (i / inner) % C.  C == 1 is per-tensor quantisation.  `elem_offset` is the logical index
 *     of x[0], so a caller can hand in any flat slice of a tensor (batch / channel-block shards,
 *     host-staging chunks) without re-deriving parameters.
 *   - dtype tags: MCTQ_F32 / MCTQ_BF16 / MCTQ_F16 describe x (and y for the affine ops).
 *   - arithmetic contract (bit-exact with CPU torch 2.11):
 *       affine:  inv = 1.0f / s;  q = clamp(rint(x * inv) + zp, qmin, qmax);  y = (q - zp) * s
 *       lut:     t = clip((x / (thr + eps)) * 2^(bw - signed), lo, hi);  idx = first argmin_k |t - lut_k|;
 *                y = (lut[idx] / 2^(bw - signed)) * thr            (y is always f32)
 *     valid for finite inputs, |qmin - zp|, |qmax - zp| < 2^21 (else an rint-based slower variant is
 *     selected automatically) and NaN -> qmin / LUT index 0.
 */
#ifndef MCTQ_H_
#define MCTQ_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCTQ_ABI_VERSION 1

/* dtype tags */
#define MCTQ_F32 0
#define MCTQ_BF16 1
#define MCTQ_F16 2

/* integer-code / LUT-index emission */
#define MCTQ_CODES_NONE 0
#define MCTQ_CODES_INT8 1 /* one byte per element: (uint8_t)(q & 0xff) -- int8 when qmin < 0, uint8 otherwise */
#define MCTQ_CODES_INT4 2 /* two elements per byte, element 2j in the low nibble; n is rounded up to even */

/* argument errors (negative so they cannot collide with cudaError_t) */
#define MCTQ_E_BADARG (-1)
#define MCTQ_E_DTYPE (-2)
#define MCTQ_E_RANGE (-3)   /* qmin > qmax, or codes requested for a range that does not fit the code width */
#define MCTQ_E_LUT (-4)     /* malformed LUT table blob */
#define MCTQ_E_NODEVICE (-5)

int mctq_abi_version(void);
/* compile-time facts a caller can assert on: "sm_100a", tile geometry, ... (static string, never freed) */
const char* mctq_build_info(void);

/* ------------------------------------------------------------------------------------------------------
 * Affine fake-quant, device pointers.
 * Replaces: torch.fake_quantize_per_channel_affine(x, scales, zero_points, axis, qmin, qmax)
 *             (weights_symmetric_inferable_quantizer.py:139-144, weights_uniform_inferable_quantizer.py:153-158)
 *           torch.fake_quantize_per_tensor_affine(x, scales[1], zero_points[1], qmin, qmax)   [tensor qparams]
 *             (weights_symmetric_inferable_quantizer.py:147-151, weights_uniform_inferable_quantizer.py:160-165)
 * x, y       device, n elements of x_dtype; y may be NULL when only codes are wanted
 * codes      device or NULL; layout per code_mode
 * scale, zp  device arrays of C entries (read on the device: no host sync, unlike the reference's
 *            per-channel path which does two .item() round trips per call).  ATen rejects per-channel zero points
 *            outside [qmin, qmax]; this entry point cannot look (no sync) and only needs |q - zp[c]| < 2^22 for every
 *            q in [qmin, qmax] when qmax - qmin < 2^21 (the fast-rounding path); the Python quantizers check the range on
 *            the host at construction time and raise ATen's error.
 */
int mctq_fq_affine(const void* x, void* y, void* codes, int64_t n, int x_dtype,
                   const float* scale, const int32_t* zp, int64_t C, int64_t inner, int64_t elem_offset,
                   int32_t qmin, int32_t qmax, int code_mode, void* stream);

/* Per-tensor affine fake-quant with scalar parameters passed by value.
 * Replaces: torch.fake_quantize_per_tensor_affine(x, scale=float, zero_point=int, qmin, qmax)
 *             (activation_symmetric_inferable_quantizer.py:113-117, activation_uniform_inferable_quantizer.py:124-128)
 * `scale` is the Python float already narrowed to f32 (what ATen does internally). */
int mctq_fq_affine_scalar(const void* x, void* y, void* codes, int64_t n, int x_dtype,
                          float scale, int32_t zp, int32_t qmin, int32_t qmax, int code_mode, void* stream);

/* Prepared per-channel parameters (the fast path for per-channel weights): mctq_affine_prepare turns scale[C] / zp[C]
 * into a device blob holding, per channel, the record {1/s, s, zp, (float)zp} plus the same values as three arrays; the
 * kernels launched by mctq_fq_affine_prepared stage the channel window of every tile from that blob with 1-D TMA bulk
 * copies (cp.async.bulk + mbarrier) instead of per-thread load / divide / store loops.  Same reference call sites and
 * arithmetic as mctq_fq_affine (1/s is the same IEEE f32 reciprocal, computed once).  Neither call synchronises.
 *   prepared_dev   DEVICE buffer, 16-byte aligned, >= mctq_affine_prepared_bytes(C) bytes; valid for as long as
 *                  scale / zp do not change */
size_t mctq_affine_prepared_bytes(int64_t C);
int mctq_affine_prepare(const float* scale, const int32_t* zp, int64_t C, void* prepared_dev, size_t prepared_bytes,
                        void* stream);
int mctq_fq_affine_prepared(const void* x, void* y, void* codes, int64_t n, int x_dtype, const void* prepared_dev,
                            int64_t C, int64_t inner, int64_t elem_offset, int32_t qmin, int32_t qmax, int code_mode,
                            void* stream);

/* Per-tensor affine fake-quant fused with the elementwise producer of the activation.
 * Replaces the pair  <producer>(x[, x2]) -> PytorchActivationQuantizationHolder.forward
 *   (mct_quantizers/pytorch/activation_quantization_holder.py:43-53 after a ReLU / ReLU6 / residual add of the exported
 *   model) with ONE kernel: the intermediate activation is never written to HBM.  Bit-identical to the eager
 *   composition (producer evaluated in f32 and rounded to x_dtype, then the recipe of mctq_fq_affine_scalar).
 * x2 is read only by the ADD flavours (same dtype and length as x). */
#define MCTQ_PRE_RELU 1      /* y = fq(max(x, 0))        */
#define MCTQ_PRE_RELU6 2     /* y = fq(min(max(x, 0), 6)) */
#define MCTQ_PRE_ADD 3       /* y = fq(x + x2)           */
#define MCTQ_PRE_ADD_RELU 4  /* y = fq(max(x + x2, 0))   */
int mctq_fq_affine_scalar_pre(const void* x, const void* x2, void* y, int64_t n, int x_dtype, int pre_op, float scale,
                              int32_t zp, int32_t qmin, int32_t qmax, void* stream);

/* Dequantise codes written by the functions above: y = (q - zp) * s  (f32 out; exact for |q - zp| < 2^22, i.e. for any
 * zero point inside the code range).  No reference call site: this is the consumer side of the code wire format
 * (SURVEY 8f rank 2). */
int mctq_dequant_affine(const void* codes, int code_mode, int is_signed, float* y, int64_t n,
                        const float* scale, const int32_t* zp, int64_t C, int64_t inner, int64_t elem_offset,
                        void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Whole-model weight quantisation: one launch over a table of tensors.
 * Replaces the per-layer loop  for name, w, quantizer in self._weights_vars: quantizer(w)
 *   (mct_quantizers/pytorch/quantize_wrapper.py:228-240 and :260-270) for affine weight quantizers.
 * `descs` and `tile_starts` live in DEVICE memory (the caller uploads them once; weights do not move).
 * tile_starts has n_desc + 1 entries: tile_starts[k] = first tile of tensor k, in units of
 * mctq_multi_tile_elems(dtype) elements; mctq_multi_plan fills it on the host.
 */
typedef struct MctqTensorDesc {
    const void* x;
    void* y;
    void* codes; /* or NULL */
    const float* scale;
    const int32_t* zp;
    int64_t n;
    int64_t C;
    int64_t inner;
    int32_t qmin;
    int32_t qmax;
    int32_t dtype;
    int32_t code_mode;
} MctqTensorDesc;

int64_t mctq_multi_tile_elems(void);
/* host helper: fills tile_starts[0..n_desc] from descs[k].n (both HOST arrays); returns total tiles or <0 */
int64_t mctq_multi_plan(const MctqTensorDesc* descs_host, int n_desc, int32_t* tile_starts_host);
int mctq_fq_affine_multi(const MctqTensorDesc* descs_dev, const int32_t* tile_starts_dev, int n_desc,
                         int64_t total_tiles, void* stream);

/* Many per-tensor (scalar-parameter) sites in ONE launch: the activation holders of a model whose inputs already exist
 * (calibration / analysis passes, batched serving of independent tensors).
 * Replaces: a loop of PytorchActivationQuantizationHolder.forward calls, each one
 *           torch.fake_quantize_per_tensor_affine(x, scale: float, zero_point: int, quant_min, quant_max)
 *           (mct_quantizers/pytorch/activation_quantization_holder.py:43-53 ->
 *            activation_symmetric_inferable_quantizer.py:113-117 / activation_uniform_inferable_quantizer.py:124-128).
 * `sites` is a HOST array (it travels as kernel parameters, 64 sites per launch); x / y are device pointers, 16-byte
 * aligned, y has x's dtype; arithmetic identical to mctq_fq_affine_scalar.  Sites with n == 0 are skipped. */
typedef struct MctqSiteDesc {
    const void* x;
    void* y;
    int64_t n;
    int32_t dtype;
    float scale;
    int32_t zp;
    int32_t qmin;
    int32_t qmax;
    int32_t reserved;
} MctqSiteDesc;
int mctq_fq_affine_scalar_multi(const MctqSiteDesc* sites_host, int n_sites, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * LUT (nearest-centroid) fake-quant.
 * Replaces: lut_quantizer(...)  mct_quantizers/pytorch/quantizer_utils.py:95-139 (which calls
 *           int_quantization_with_threshold :142-170) as used by
 *           weights_lut_symmetric_inferable_quantizer.py:114-122, weights_lut_pot_inferable_quantizer.py:74-104,
 *           activation_lut_pot_inferable_quantizer.py:86-91.
 *
 * The centroid list is first compiled (on the host, once per quantizer) into a search table: sorted
 * unique centroids, the exact f32 decision threshold between each adjacent pair under torch.argmin's
 * first-minimum rule, and the original index of every sorted entry.  The table is a POD blob of
 * mctq_lut_table_bytes(K) bytes that the caller copies to the device; K (the length of the original
 * centroid list, <= 256) travels with it because the blob's geometry is a function of K alone.
 */
size_t mctq_lut_table_bytes(int K);
int mctq_lut_build_table(const float* lut_host, int K, int lut_values_bitwidth, int is_signed,
                         void* table_host_out, size_t table_bytes);

/* weights flavour: thr is a DEVICE f32 array [C]; d_c = thr_c + (float)eps (f32 add), no intermediate
 * rounding, f32 output.  idx (optional) receives the LUT index per element (MCTQ_CODES_INT8: one byte,
 * MCTQ_CODES_INT4: packed nibbles, K <= 16). */
int mctq_fq_lut(const void* x, float* y, void* idx, int64_t n, int x_dtype,
                const void* table_dev, int K, const float* thr, int64_t C, int64_t inner, int64_t elem_offset,
                float eps, int idx_mode, void* stream);

/* activation flavour: threshold and eps are Python floats in the reference, so the divisor is
 * d = (float)(thr + eps) computed in double by the caller and passed by value with thr_f32 = (float)thr;
 * for bf16/f16 inputs the reference's eager ops round the normalised value back to the input dtype
 * before the search (round_to_x_dtype = 1). */
#define MCTQ_LUT_DIVISOR_IS_MULTIPLIER 2   /* OR into round_to_x_dtype: `divisor` is r = (float)(1.0 / (thr + eps)) and the
                                            * normalisation is x * r -- what libtorch's CUDA kernel for `tensor / python_number`
                                            * computes, i.e. what the unmodified reference yields for CUDA tensors (it differs from
                                            * the CPU kernel's true division at rounding ties).  mctq_lut_prepare: scalar_mode = 2 */
int mctq_fq_lut_scalar(const void* x, float* y, void* idx, int64_t n, int x_dtype,
                       const void* table_dev, int K, float divisor, float thr_f32, int round_to_x_dtype,
                       int idx_mode, void* stream);

/* Prepared flavour (the fast path): per-channel decision tables in the x domain.
 * Normalise -> clip -> argmin is a monotone function of x for each channel, so the LUT entry is decided by which of the
 * K - 1 per-channel thresholds X[c][j] the element exceeds.  mctq_lut_prepare computes them exactly (bisection over f32
 * bit patterns through the reference arithmetic: IEEE division by thr + eps, optional rounding to the activation dtype,
 * first-minimum thresholds of the search table) together with the dequantised outputs (lut / 2^(bw - s)) * thr_c, once per
 * quantizer; mctq_fq_lut_prepared then runs without any division or search loop.  Same reference call sites as above.
 *   table_host        the HOST copy of the blob from mctq_lut_build_table
 *   thr_dev           DEVICE f32 [C] (weights flavour) or NULL with scalar_mode = 1 (divisor / thr_f32 by value, C = 1)
 *   round_dtype       0 none; 1 / 2: the normalised value is rounded to bf16 / f16 first (activation flavour)
 *   prepared_dev      DEVICE buffer of mctq_lut_prepared_bytes(K, bw, signed, C) bytes (0 = configuration unsupported)
 * Grids of more than 10 bits use a cell table coarser than the integer grid; mctq_lut_prepare returns MCTQ_E_RANGE when the
 * centroid list is too dense for it (two decision thresholds in one cell): use the generic entry points then.
 * mctq_lut_prepare is one-off setup and synchronises `stream` once.  mctq_fq_lut_prepared needs 16-byte aligned x / y and
 * returns MCTQ_E_RANGE when the channel window of a tile does not fit in shared memory (rows shorter than ~16
 * elements with large K): callers fall back to mctq_fq_lut. */
size_t mctq_lut_prepared_bytes(int K, int lut_values_bitwidth, int is_signed, int64_t C);
int mctq_lut_prepare(const void* table_host, int K, const float* thr_dev, int64_t C, float eps, int scalar_mode,
                     float divisor, float thr_f32, int round_dtype, void* prepared_dev, size_t prepared_bytes,
                     void* stream);
int mctq_fq_lut_prepared(const void* x, float* y, void* idx, int64_t n, int x_dtype, const void* prepared_dev, int K,
                         int lut_values_bitwidth, int is_signed, int64_t C, int64_t inner, int64_t elem_offset,
                         int idx_mode, void* stream);

/* Whole-model LUT weight quantization in ONE launch (the LUT counterpart of mctq_fq_affine_multi; replaces the per-layer
 * loop mct_quantizers/pytorch/quantize_wrapper.py:228-240 over weights_lut_symmetric_inferable_quantizer.py:89-128 /
 * weights_lut_pot_inferable_quantizer.py:74-104 calls).  Every tensor brings its own prepared blob (mctq_lut_prepare).
 * mctq_lut_multi_plan compiles the HOST descriptor array into an opaque plan blob of
 * mctq_lut_multi_plan_bytes(descs, n_desc) bytes in HOST memory (returns the number of CTAs of the launch, or < 0:
 * MCTQ_E_RANGE / MCTQ_E_BADARG mean "this tensor needs the single-tensor / generic entry point").  The launch passes the
 * per-tensor argument blocks as kernel parameters, so there is no device-side copy of the plan; the tensors are grouped by
 * kernel variant (dtype, channel mode, vector width, record kind) and every group runs as launches of <= 64 tensors, each
 * specialised for its variant (a model whose LUT weights share one layout is one launch per 64 tensors).
 * x, y and the prepared blobs must stay valid while the plan is in use. */
typedef struct MctqLutTensorDesc {
    const void* x;            /* device, dtype below, 16-byte aligned */
    float* y;                 /* device f32, 16-byte aligned */
    const void* prepared_dev; /* mctq_lut_prepare blob of this tensor's quantizer */
    int64_t n;
    int64_t C;
    int64_t inner;
    int32_t dtype;
    int32_t K;
    int32_t lut_values_bitwidth;
    int32_t is_signed;
} MctqLutTensorDesc;
size_t mctq_lut_multi_plan_bytes(const MctqLutTensorDesc* descs_host, int n_desc);   /* 0: some tensor is not plannable */
int64_t mctq_lut_multi_plan(const MctqLutTensorDesc* descs_host, int n_desc, void* plan_host_out, size_t plan_bytes);
int mctq_fq_lut_prepared_multi(const void* plan_host, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Host-buffer entry points: the same operators for tensors that live in HOST memory (pinned memory
 * overlaps; pageable memory works but serialises).  The data is streamed through `staging_dev`
 * (device scratch owned by the caller, >= mctq_host_staging_min_bytes()) in chunks on internal streams:
 * H2D copy, kernel and D2H copy of neighbouring chunks overlap.  Parameters are HOST arrays (copied
 * before the call returns; at most 384 Ki channels).  By default these calls return after y_host is
 * complete (they synchronise their internal streams only).
 * They are what a CPU-tensor call of a quantizer maps to: the product has no CPU arithmetic path.
 *
 * Deferred mode (mctq_host_set_deferred(device, 1)): the calls return as soon as their copies and kernels
 * are enqueued, so the three-stage pipeline keeps running across the tensors of a model instead of
 * filling and draining once per tensor; x_host must stay valid and y_host is complete only after
 * mctq_host_wait(device) (or after deferred mode is switched off, which waits as well).
 */
size_t mctq_host_staging_min_bytes(void);
int mctq_fq_affine_host(const void* x_host, void* y_host, int64_t n, int x_dtype,
                        const float* scale_host, const int32_t* zp_host, int64_t C, int64_t inner,
                        int32_t qmin, int32_t qmax, void* staging_dev, size_t staging_bytes, int device);
int mctq_host_set_deferred(int device, int on);
int mctq_host_wait(int device);
int mctq_fq_lut_host(const void* x_host, float* y_host, int64_t n, int x_dtype,
                     const void* table_host, int K, const float* thr_host, int64_t C, int64_t inner,
                     float eps, int scalar_mode, float divisor, float thr_f32, int round_to_x_dtype,
                     void* staging_dev, size_t staging_bytes, int device);

/* ------------------------------------------------------------------------------------------------------
 * Introspection / test hooks (no reference counterpart).
 */
/* launches of this library's kernels since load (what bench.py reports as gpu_launches) */
int64_t mctq_launch_count(void);
/* variant selection for experiments: key 0 = unroll (0 = automatic [default], 2, 4, 8), key 1 = force rint path (0/1),
 * key 2 = force IEEE-division LUT path (0/1), key 3 = programmatic dependent launch: 0 off, 1 wait-then-load (default),
 *   2 = load-then-wait for launches whose input is not an output of a launch of this library still in flight on the stream,
 *   3 = additionally no wait at all (until a CTA's last instruction) for launches whose buffers are disjoint from those of
 *   every launch still in flight -- 2 and 3 are only legal while the stream carries NO work of other libraries (their
 *   kernels may release dependents before their stores are visible; the allocator may recycle their buffers); opt-in, see
 *   mct_quantizers_b200.private_stream(); changing the key makes the next launch on every stream a waiting one,
 * key 4 = warp-shuffle search in the generic LUT kernel for tables of <= 32 entries (default 1),
 * key 5 = wide vectors (8 elements per vector, 256-bit stores) in the kernels that have them (default 1),
 * key 6 = (retired: the multi-tensor LUT launches run one tile per CTA; the key is accepted and ignored),
 * key 7 = xy-record variant of the prepared LUT kernel for per-tensor / long-row launches (default 1),
 * key 8 = kernels stage their prepared parameter tables BEFORE the dependent-launch wait whenever the blob was not written by
 *   a prepare call still in flight on the stream (default 1; the blobs are private to the library, so this is legal whatever
 *   else runs on the stream),
 * key 9 = NVTX ranges (nvtx3, header-only) around every compute entry point, named after the entry point (default 0;
 *   `ncu --nvtx` and timeline tools then attribute kernels to the call of the reference's API they belong to);
 * key 10 = with key 3 = 3: how many launches may be in flight together before one waits again (2..8, default 3);
 * key 11 = data streams (= staging slots) of the host-buffer pipeline (2..6, default 3; mctq_host_staging_min_bytes follows),
 * key 12 = number of chunks a tensor is cut into by the host-buffer pipeline in deferred mode (1..16, default 1: whole staging slots);
 * returns previous value or <0 */
int mctq_set_tuning(int key, int value);
/* device self-test of the 5-op correctly-rounded division used by the LUT kernels against __fdiv_rn
 * on n_pairs pseudo-random (x, d) pairs; *mismatches_dev (device int64) receives the number that differ */
int mctq_selftest_division(int64_t n_pairs, uint64_t seed, int64_t* mismatches_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MCTQ_H_ */
