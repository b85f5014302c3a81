"""torch.library operators of the `mctq` namespace: the boundary between the quantizer objects and the
sm_100a kernels.

    torch.ops.mctq.fq_affine_scalar   <- torch.fake_quantize_per_tensor_affine(x, float, int, qmin, qmax)
    torch.ops.mctq.fq_affine_tensor   <- torch.fake_quantize_per_tensor_affine(x, Tensor[1], Tensor[1], qmin, qmax)
    torch.ops.mctq.fq_affine_channel  <- torch.fake_quantize_per_channel_affine(x, scale, zp, axis, qmin, qmax)
    torch.ops.mctq.fq_lut_tensor      <- lut_quantizer(..., threshold=Tensor, per_channel, channel_axis, input_rank)
    torch.ops.mctq.fq_lut_scalar      <- lut_quantizer(..., threshold=float)            (activations)
    torch.ops.mctq.quantize_affine_*  -> integer codes (+ optional fake-quant values); mctq.dequantize_affine

(reference call sites: see include/mctq.h).  Each op has a CUDA implementation (device pointers + the
caller's current stream into the C ABI, no synchronisation), a CPU-tensor implementation that streams the
host buffer through the GPU (mctq_fq_*_host), and a fake/meta implementation so that fx tracing,
torch.compile and shape propagation work.  There is no eager / CPU arithmetic fallback anywhere.
"""
import ctypes

import numpy as np
import torch

from mct_quantizers_b200 import _native
from mct_quantizers_b200._native import MctqError, c_vp

_DT = {torch.float32: _native.F32, torch.bfloat16: _native.BF16, torch.float16: _native.F16}

_LIB = torch.library.Library("mctq", "DEF")
_LIB.define("fq_affine_scalar(Tensor x, float scale, int zero_point, int quant_min, int quant_max) -> Tensor")
_LIB.define("fq_affine_scalar_pre(Tensor x, Tensor? other, int pre_op, float scale, int zero_point, int quant_min, int quant_max) -> Tensor")
_LIB.define("fq_affine_tensor(Tensor x, Tensor scale, Tensor zero_point, int quant_min, int quant_max) -> Tensor")
_LIB.define("fq_affine_channel(Tensor x, Tensor scale, Tensor zero_point, int axis, int quant_min, int quant_max) -> Tensor")
_LIB.define("fq_lut_tensor(Tensor x, Tensor table, int K, Tensor threshold, bool per_channel, int axis, float eps) -> Tensor")
_LIB.define("fq_lut_scalar(Tensor x, Tensor table, int K, float divisor, float threshold, bool round_to_input_dtype, bool multiply=False) -> Tensor")
_LIB.define("quantize_affine_channel(Tensor x, Tensor scale, Tensor zero_point, int axis, int quant_min, int quant_max, "
            "int code_mode, bool want_values) -> (Tensor, Tensor)")
_LIB.define("dequantize_affine(Tensor codes, int code_mode, bool is_signed, int[] shape, Tensor scale, Tensor zero_point, "
            "int axis) -> Tensor")
_LIB.define("lut_indices(Tensor x, Tensor table, int K, Tensor threshold, bool per_channel, int axis, float eps, "
            "int idx_mode) -> Tensor")


# --------------------------------------------------------------------------------------------- helpers
def _dtype_tag(x):
    try:
        return _DT[x.dtype]
    except KeyError:
        raise NotImplementedError(f'"mctq_fake_quantize" not implemented for \'{x.dtype}\' '
                                  f'(float32, bfloat16 and float16 are supported)') from None


def _ptr(t):
    return t.data_ptr() if t is not None else None      # ctypes converts ints for c_void_p parameters


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_get_device = getattr(torch._C, "_cuda_getDevice", None)
_set_device = getattr(torch._C, "_cuda_setDevice", None)
if _raw_stream is None:                      # private fast paths missing in this torch build: public equivalents
    def _raw_stream(index):
        return torch.cuda.current_stream(index).cuda_stream
if _get_device is None:
    _get_device = torch.cuda.current_device
if _set_device is None:
    _set_device = torch.cuda.set_device


def _stream(device):
    """cudaStream_t (as int) of the current stream of `device`."""
    if _raw_stream is not None and device.index is not None:
        return _raw_stream(device.index)
    return torch.cuda.current_stream(device).cuda_stream


_PLAIN_TENSOR_TYPES = (torch.Tensor, torch.nn.Parameter)


def direct_ok(x):
    """True when the quantizers may call the launch functions below directly instead of going through the
    torch.ops dispatcher (saves ~10 us per call): a plain CUDA tensor, and nobody is tracing / compiling / exporting
    (tracers must see the custom operator)."""
    return (type(x) in _PLAIN_TENSOR_TYPES and x.is_cuda and not torch.jit.is_tracing()
            and not torch.compiler.is_compiling())


class _on_device:
    """Make `device` current for the duration of a launch (kernels launch on the current device)."""
    __slots__ = ("idx", "prev")

    def __init__(self, device):
        self.idx = device.index if device.index is not None else torch.cuda.current_device()

    def __enter__(self):
        self.prev = _get_device() if _get_device is not None else torch.cuda.current_device()
        if self.prev != self.idx:
            torch.cuda.set_device(self.idx)
        return self

    def __exit__(self, *exc):
        if self.prev != self.idx:
            torch.cuda.set_device(self.prev)


def _is_dense_permutation(x):
    """True when x's elements occupy one gap-free block of memory in some permutation of its dims."""
    dims = sorted((d for d in range(x.dim()) if x.shape[d] > 1), key=lambda d: x.stride(d))
    expect = 1
    for d in dims:
        if x.stride(d) != expect:
            return False
        expect *= x.shape[d]
    return True


def _dense(x):
    """x itself when its memory is one dense block (any permutation of strides), else a contiguous copy."""
    if x.is_contiguous() or x.numel() == 0:
        return x
    if _is_dense_permutation(x):
        return x
    return x.contiguous()


def _channel_layout(x, axis):
    """(x_dense, C, inner): the [outer][C][inner] view in MEMORY order.  For a permuted-but-dense tensor
    (e.g. channels_last) the channel axis keeps its meaning; only `inner` changes."""
    nd = x.dim()
    if not -nd <= axis < nd:
        raise IndexError(f"Dimension out of range (expected to be in range of [{-nd}, {nd - 1}], but got {axis})")
    axis %= nd
    x = _dense(x)
    C = x.shape[axis]
    if x.is_contiguous():
        inner = 1
        for s in x.shape[axis + 1:]:
            inner *= int(s)
        return x, int(C), inner
    # dense, permuted: dims that sit "inside" the channel axis in memory are those with a smaller stride
    # -- valid only if they tile the channel stride exactly
    strides, shape = x.stride(), x.shape
    inner = 1
    for d in range(nd):
        if d != axis and shape[d] > 1 and strides[d] < strides[axis]:
            inner *= int(shape[d])
    if C > 1 and strides[axis] != inner:
        x = x.contiguous()
        return _channel_layout(x, axis)
    return x, int(C), inner


def _check_range(qmin, qmax):
    if qmin > qmax:
        raise RuntimeError("`quant_min` should be less than or equal to `quant_max`.")


def _check_channel_params(x, scale, zp, axis):
    if scale.dtype != torch.float32:
        raise RuntimeError(f"Scale must be Float, found {scale.dtype}")
    if zp.dtype != torch.int32:
        raise RuntimeError(f"Zero-point must be Int32, found {zp.dtype}")
    if scale.dim() != 1 or zp.dim() != 1:
        raise RuntimeError("scale and zero-point need to be 1-D tensors")
    nd = x.dim()
    if not -nd <= axis < nd:
        raise RuntimeError("`axis` must be between 0 and number of dimensions of input")
    if scale.numel() != x.shape[axis] or zp.numel() != x.shape[axis]:
        raise RuntimeError("dimensions of scale and zero-point are not consistent with input tensor")


def _ver(t):
    """Version counter of a parameter tensor for cache keys; inference tensors have none (and cannot be edited in place
    outside inference mode), so they get a constant."""
    return -1 if t.is_inference() else t._version


_PARAM_COPIES = {}        # (data_ptr, version, numel, dtype, src device, dst device) -> (source kept alive, copy)


def _param_on(t, device):
    """`t` on `device`.  Cross-device copies are cached: a fresh copy per call would also defeat the prepared-parameter
    caches below (their keys hold the copy's data_ptr) and re-prepare on every call."""
    if t.device == device:
        return t
    key = (t.data_ptr(), _ver(t), t.numel(), t.dtype, str(t.device), str(device))
    hit = _PARAM_COPIES.get(key)
    if hit is None:
        if len(_PARAM_COPIES) >= 1024:
            _PARAM_COPIES.clear()
        hit = (t, t.to(device, non_blocking=True))
        _PARAM_COPIES[key] = hit
    return hit[1]


_staging = {}


def _staging_buffer(device_index):
    buf = _staging.get(device_index)
    nbytes = _native.load().mctq_host_staging_min_bytes()        # depends on the number of pipeline streams (tuning key 11)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(nbytes, dtype=torch.uint8, device=torch.device("cuda", device_index))
        _staging[device_index] = buf
    return buf


def _require_cuda_for_host_path():
    if not torch.cuda.is_available():
        raise MctqError("mct_quantizers_b200 received a CPU tensor but no CUDA device is available: the package "
                        "has no CPU arithmetic path (host tensors are streamed through the GPU)")
    return torch.cuda.current_device()


def _host_array(t, dtype):
    return np.ascontiguousarray(t.detach().cpu().numpy().astype(dtype, copy=False).reshape(-1))


# --------------------------------------------------------------------------------------------- CUDA impls
def affine_scalar_direct(x, scale_f32, zero_point, quant_min, quant_max):
    """Lean launch (no dispatcher, no re-validation): `scale_f32` is already narrowed to f32, the range was validated
    when the quantizer was built."""
    tag = _DT.get(x.dtype)
    if tag is None:
        _dtype_tag(x)
    xd = x if x.is_contiguous() else _dense(x)
    y = torch.empty_like(xd)
    n = xd.numel()
    if n:
        lib = _native._lib or _native.load()
        dev = xd.device
        prev = _get_device()
        if prev != dev.index:
            _set_device(dev.index)
        fast = _native.fast                     # CPython front door (METH_FASTCALL) when built; ctypes otherwise -- same entry point
        rc = (fast.fq_affine_scalar if fast is not None else lib.mctq_fq_affine_scalar)(
            xd.data_ptr(), y.data_ptr(), None, n, tag, scale_f32, zero_point, quant_min, quant_max, 0, _raw_stream(dev.index))
        if prev != dev.index:
            _set_device(prev)
        if rc:
            _native.check(rc, "mctq_fq_affine_scalar")
    return y


def _affine_scalar_cuda(x, scale, zero_point, quant_min, quant_max):
    _dtype_tag(x)
    _check_range(quant_min, quant_max)
    if not quant_min <= zero_point <= quant_max:
        raise RuntimeError("`zero_point` must be between `quant_min` and `quant_max`.")
    return affine_scalar_direct(x, float(np.float32(scale)), int(zero_point), int(quant_min), int(quant_max))


def affine_scalar_pre_direct(x, other, pre_op, scale_f32, zero_point, quant_min, quant_max):
    """Producer (relu / relu6 / add / add+relu) and per-tensor fake-quant in ONE kernel: mctq_fq_affine_scalar_pre.
    Lean launch; arguments are already validated."""
    tag = _DT.get(x.dtype)
    if tag is None:
        _dtype_tag(x)
    xd = x if x.is_contiguous() else _dense(x)
    od = None
    if pre_op in (_native.PRE_ADD, _native.PRE_ADD_RELU):
        if other is None or other.shape != x.shape or other.dtype != x.dtype or other.device != x.device:
            raise RuntimeError("fused add needs a second tensor of the same shape, dtype and device (no broadcasting)")
        od = other if other.stride() == xd.stride() and _is_dense_permutation(other) else None
        if od is None:
            od = other.contiguous()
            if not xd.is_contiguous():
                xd = xd.contiguous()
    y = torch.empty_like(xd)
    n = xd.numel()
    if n:
        lib = _native._lib or _native.load()
        index = xd.device.index
        prev = _get_device()
        if prev != index:
            _set_device(index)
        fast = _native.fast
        rc = (fast.fq_affine_scalar_pre if fast is not None else lib.mctq_fq_affine_scalar_pre)(
            xd.data_ptr(), od.data_ptr() if od is not None else None, y.data_ptr(), n, tag, pre_op, scale_f32, zero_point,
            quant_min, quant_max, _raw_stream(index))
        if prev != index:
            _set_device(prev)
        if rc:
            _native.check(rc, "mctq_fq_affine_scalar_pre")
    return y


def _affine_scalar_pre_cuda(x, other, pre_op, scale, zero_point, quant_min, quant_max):
    _dtype_tag(x)
    _check_range(quant_min, quant_max)
    if not quant_min <= zero_point <= quant_max:
        raise RuntimeError("`zero_point` must be between `quant_min` and `quant_max`.")
    if pre_op not in (_native.PRE_RELU, _native.PRE_RELU6, _native.PRE_ADD, _native.PRE_ADD_RELU):
        raise RuntimeError("pre_op must be 1 (relu), 2 (relu6), 3 (add) or 4 (add + relu)")
    if pre_op in (_native.PRE_ADD, _native.PRE_ADD_RELU) and (other is None or other.shape != x.shape or other.dtype != x.dtype
                                                              or other.device != x.device):
        # broadcasting / type-promoting add: the eager add, then the (relu-)fused fake-quant
        if other is None:
            raise RuntimeError("pre_op add needs `other`")
        t = torch.add(x, other)
        if pre_op == _native.PRE_ADD:
            return _affine_scalar_cuda(t, scale, zero_point, quant_min, quant_max)
        x, other, pre_op = t, None, _native.PRE_RELU
    return affine_scalar_pre_direct(x, other, int(pre_op), float(np.float32(scale)), int(zero_point), int(quant_min), int(quant_max))


# ---- prepared per-channel parameters: {1/s, s, zp} records the kernels stage with TMA bulk copies (mctq_affine_prepare)
_AFFINE_PREPARED = {}     # (scale ptr, zp ptr, versions, C, device) -> (scale, zp kept alive so the key stays unique, blob)


def clear_affine_caches():
    _AFFINE_PREPARED.clear()


def _affine_prepared(lib, scale, zero_point, index):
    """Device blob of prepared parameters for (scale, zero_point), built once per parameter pair; None while a CUDA
    graph is being captured and the blob does not exist yet (the raw-parameter entry point is used instead)."""
    key = (scale.data_ptr(), zero_point.data_ptr(), _ver(scale), _ver(zero_point), scale.numel(), index)
    ent = _AFFINE_PREPARED.get(key)
    if ent is None:
        if torch.cuda.is_current_stream_capturing():
            return None
        if len(_AFFINE_PREPARED) >= 4096:
            _AFFINE_PREPARED.clear()
        C = scale.numel()
        blob = torch.empty(lib.mctq_affine_prepared_bytes(C), dtype=torch.uint8, device=scale.device)
        stream = _raw_stream(index)
        rc = lib.mctq_affine_prepare(scale.data_ptr(), zero_point.data_ptr(), C, blob.data_ptr(), blob.numel(), stream)
        if rc:
            _native.check(rc, "mctq_affine_prepare")
        torch.cuda.current_stream(scale.device).synchronize()     # one-off: the blob may be used from any stream afterwards
        ent = (scale, zero_point, blob)
        _AFFINE_PREPARED[key] = ent
    return ent[2]


def _launch_affine_params(lib, x_ptr, y_ptr, codes_ptr, n, tag, scale, zero_point, C, inner, quant_min, quant_max, code_mode, index):
    """mctq_fq_affine_prepared for per-channel parameters (C > 1), mctq_fq_affine otherwise.  The caller has made
    `index` the current device."""
    fast = _native.fast
    if C > 1 and scale.is_contiguous() and zero_point.is_contiguous():
        blob = _affine_prepared(lib, scale, zero_point, index)
        if blob is not None:
            return (fast.fq_affine_prepared if fast is not None else lib.mctq_fq_affine_prepared)(
                x_ptr, y_ptr, codes_ptr, n, tag, blob.data_ptr(), C, inner, 0, quant_min, quant_max, code_mode,
                _raw_stream(index)), "mctq_fq_affine_prepared"
    return (fast.fq_affine if fast is not None else lib.mctq_fq_affine)(
        x_ptr, y_ptr, codes_ptr, n, tag, scale.data_ptr(), zero_point.data_ptr(), C, inner, 0, quant_min, quant_max, code_mode,
        _raw_stream(index)), "mctq_fq_affine"


def affine_params_direct(x, scale, zero_point, C, inner, quant_min, quant_max):
    """Lean launch of mctq_fq_affine for parameter TENSORS already on x's device and already validated
    (x: plain CUDA tensor; the (C, inner) view refers to x's memory order)."""
    tag = _DT.get(x.dtype)
    if tag is None:
        _dtype_tag(x)
    y = torch.empty_like(x)
    n = x.numel()
    if n:
        lib = _native._lib or _native.load()
        index = x.device.index
        prev = _get_device()
        if prev != index:
            _set_device(index)
        rc, what = _launch_affine_params(lib, x.data_ptr(), y.data_ptr(), None, n, tag, scale, zero_point, C, inner,
                                         quant_min, quant_max, 0, index)
        if prev != index:
            _set_device(prev)
        if rc:
            _native.check(rc, what)
    return y


def contiguous_layout(shape, axis):
    """(C, inner) of a contiguous tensor of `shape` quantized along `axis`."""
    inner = 1
    for d in shape[axis + 1:]:
        inner *= d
    return shape[axis], inner


def _affine_tensor_cuda(x, scale, zero_point, quant_min, quant_max):
    tag = _dtype_tag(x)
    _check_range(quant_min, quant_max)
    if scale.numel() != 1 or zero_point.numel() != 1:
        raise RuntimeError(f"a Tensor with {max(scale.numel(), zero_point.numel())} elements cannot be converted to Scalar")
    if scale.dtype != torch.float32 or zero_point.dtype != torch.int32:
        raise RuntimeError("scale must be Float and zero_point Int32")
    xd = _dense(x)
    y = torch.empty_like(xd)
    if xd.numel():
        lib = _native.load()
        s, z = _param_on(scale, xd.device), _param_on(zero_point, xd.device)
        with _on_device(xd.device):
            rc = lib.mctq_fq_affine(_ptr(xd), _ptr(y), None, xd.numel(), tag, _ptr(s), _ptr(z), 1, 1, 0,
                                    int(quant_min), int(quant_max), 0, _stream(xd.device))
        _native.check(rc, "mctq_fq_affine")
    return y


def _affine_channel_launch(x, scale, zero_point, axis, quant_min, quant_max, code_mode, want_values):
    tag = _dtype_tag(x)
    _check_range(quant_min, quant_max)
    _check_channel_params(x, scale, zero_point, axis)
    xd, C, inner = _channel_layout(x, axis)
    y = torch.empty_like(xd) if want_values else None
    codes = None
    n = xd.numel()
    if code_mode == _native.CODES_INT8:
        codes = torch.empty_like(xd, dtype=torch.int8 if quant_min < 0 else torch.uint8)
    elif code_mode == _native.CODES_INT4:
        codes = torch.empty((n + 1) // 2, dtype=torch.uint8, device=xd.device)
    if n:
        lib = _native.load()
        s, z = _param_on(scale, xd.device), _param_on(zero_point, xd.device)
        with _on_device(xd.device) as cur:
            rc, what = _launch_affine_params(lib, _ptr(xd), _ptr(y), _ptr(codes), n, tag, s.contiguous(), z.contiguous(), C, inner,
                                             int(quant_min), int(quant_max), int(code_mode), cur.idx)
        _native.check(rc, what)
    return y, codes


def _affine_channel_cuda(x, scale, zero_point, axis, quant_min, quant_max):
    return _affine_channel_launch(x, scale, zero_point, axis, quant_min, quant_max, 0, True)[0]


def _quantize_affine_channel_cuda(x, scale, zero_point, axis, quant_min, quant_max, code_mode, want_values):
    if code_mode not in (_native.CODES_INT8, _native.CODES_INT4):
        raise RuntimeError("code_mode must be 1 (int8) or 2 (packed int4)")
    y, codes = _affine_channel_launch(x, scale, zero_point, axis, quant_min, quant_max, code_mode, want_values)
    if y is None:
        y = torch.empty(0, dtype=x.dtype, device=x.device)
    return codes, y


def _dequantize_affine_cuda(codes, code_mode, is_signed, shape, scale, zero_point, axis):
    n = 1
    for s in shape:
        n *= int(s)
    nd = len(shape)
    axis %= max(nd, 1)
    C = int(shape[axis]) if nd else 1
    inner = 1
    for s in shape[axis + 1:]:
        inner *= int(s)
    if scale.numel() != C or zero_point.numel() != C:
        raise RuntimeError("dimensions of scale and zero-point are not consistent with `shape`")
    y = torch.empty(list(shape), dtype=torch.float32, device=codes.device)
    if n:
        lib = _native.load()
        codes = codes.contiguous()
        s, z = _param_on(scale, codes.device), _param_on(zero_point, codes.device)
        with _on_device(codes.device):
            rc = lib.mctq_dequant_affine(_ptr(codes), int(code_mode), int(bool(is_signed)), _ptr(y), n, _ptr(s), _ptr(z),
                                         C, inner, 0, _stream(codes.device))
        _native.check(rc, "mctq_dequant_affine")
    return y


# ---- LUT: per-device copies of the search table and per-(quantizer, device, dtype) prepared decision tables
_LUT_DEV_TABLES = {}      # (table.data_ptr(), device) -> (table_cpu kept alive, table_dev)
_LUT_PREPARED = {}        # key -> (objects kept alive so that data_ptr keys stay unique, prepared_dev or None)
_ROUND_TAG = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


def clear_lut_caches():
    _LUT_DEV_TABLES.clear()
    _LUT_PREPARED.clear()
    _PARAM_COPIES.clear()


def _table_header(table_cpu):
    """(bw, is_signed) out of the search-table blob (LutTableHeader: magic, K, Ks, P, levels, pos0, bw, is_signed, ...)."""
    hdr = table_cpu[:32].numpy().view(np.int32)
    return int(hdr[6]), int(hdr[7])


def _table_on(table, device):
    if table.device == device:
        return table
    key = (table.data_ptr(), str(device))
    hit = _LUT_DEV_TABLES.get(key)
    if hit is None:
        if torch.cuda.is_current_stream_capturing():
            # first use inside a CUDA-graph capture: the upload becomes a node of the graph and must read pinned memory
            # that outlives the graph (kept in the cache entry)
            src = table.pin_memory()
            hit = (table, src.to(device, non_blocking=True), src)
        else:
            hit = (table, table.to(device))
        _LUT_DEV_TABLES[key] = hit
    return hit[1]


def _prepared_for(table, K, device, thr_dev, eps, scalar, divisor, thr_f32, round_dtype):
    """Device blob of per-channel decision tables for this (centroid list, thresholds, device, rounding), built once.
    None when the configuration is outside the prepared path (table not on the host; a grid of more than 10 bits with a
    centroid list too dense for the cell table)."""
    if table.device.type != 'cpu':
        return None
    if scalar:
        key = (table.data_ptr(), int(K), str(device), 's', int(scalar), float(divisor), float(thr_f32), int(round_dtype))
        C = 1
    else:
        key = (table.data_ptr(), int(K), str(device), 't', thr_dev.data_ptr(), _ver(thr_dev), thr_dev.numel(), float(eps))
        C = thr_dev.numel()
    hit = _LUT_PREPARED.get(key)
    if hit is None:
        if torch.cuda.is_current_stream_capturing():
            return None                     # preparation synchronises: not inside a CUDA-graph capture (generic kernel instead)
        if len(_LUT_PREPARED) >= 1024:      # bounded like _AFFINE_PREPARED: entries pin a threshold tensor and a device blob
            _LUT_PREPARED.clear()
        lib = _native.load()
        bw, signed = _table_header(table)
        nbytes = lib.mctq_lut_prepared_bytes(int(K), bw, signed, C)
        blob = None
        if nbytes:
            blob = torch.empty(nbytes, dtype=torch.uint8, device=device)
            with _on_device(device):
                rc = lib.mctq_lut_prepare(_ptr(table), int(K), _ptr(thr_dev) if not scalar else None, C,
                                          float(np.float32(eps)), int(scalar), float(np.float32(divisor)),
                                          float(np.float32(thr_f32)), int(round_dtype), _ptr(blob), nbytes, _stream(device))
            if rc == -3:                                         # MCTQ_E_RANGE: a grid of more than 10 bits whose centroids are too
                blob = None                                      # dense for the coarse cell table -> generic kernel
            else:
                _native.check(rc, "mctq_lut_prepare")
                torch.cuda.current_stream(device).synchronize()  # one-off: the blob may be used from any stream afterwards
        hit = ((table, thr_dev), blob, bw, signed)
        _LUT_PREPARED[key] = hit
    return hit


def _lut_launch(xd, y, idx, idx_mode, table, K, C, inner, thr_dev, eps, scalar, divisor, thr_f32, round_flag):
    """Prepared kernel when possible, generic kernel otherwise."""
    lib = _native.load()
    tag = _dtype_tag(xd)
    n = xd.numel()
    round_dtype = _ROUND_TAG[xd.dtype] if (scalar and round_flag) else 0
    prep = _prepared_for(table, K, xd.device, thr_dev, eps, scalar, divisor, thr_f32, round_dtype)
    with _on_device(xd.device):
        if prep is not None and prep[1] is not None:
            rc = lib.mctq_fq_lut_prepared(_ptr(xd), _ptr(y), _ptr(idx), n, tag, _ptr(prep[1]), int(K), prep[2], prep[3],
                                          C, inner, 0, int(idx_mode), _stream(xd.device))
            if rc == 0:
                return
            if rc not in (-1, -3):                      # BADARG (misaligned view) / RANGE (window too wide): generic path
                _native.check(rc, "mctq_fq_lut_prepared")
        t = _table_on(table, xd.device)
        if scalar:
            rc = lib.mctq_fq_lut_scalar(_ptr(xd), _ptr(y), _ptr(idx), n, tag, _ptr(t), int(K), float(np.float32(divisor)),
                                        float(np.float32(thr_f32)), int(bool(round_flag)) | (2 if scalar == 2 else 0), int(idx_mode),
                                        _stream(xd.device))
        else:
            rc = lib.mctq_fq_lut(_ptr(xd), _ptr(y), _ptr(idx), n, tag, _ptr(t), int(K), _ptr(thr_dev), C, inner, 0,
                                 float(np.float32(eps)), int(idx_mode), _stream(xd.device))
    _native.check(rc, "mctq_fq_lut")


def _lut_tensor_launch(x, table, K, threshold, per_channel, axis, eps, idx_mode, want_values):
    _dtype_tag(x)
    if threshold.dtype != torch.float32:
        raise RuntimeError(f"threshold must be Float, found {threshold.dtype}")
    if per_channel:
        if threshold.numel() != x.shape[axis]:
            raise RuntimeError(f"shape '{[1] * x.dim()}' is invalid for threshold of size {threshold.numel()} "
                               f"(input has {x.shape[axis]} channels on axis {axis})")
        xd, C, inner = _channel_layout(x.contiguous(), axis)     # row-major output like the reference's eager composition
    else:
        if threshold.numel() != 1:
            raise RuntimeError("per-tensor LUT quantization needs a single threshold")
        xd, C, inner = x.contiguous(), 1, 1
    n = xd.numel()
    y = torch.empty_like(xd, dtype=torch.float32) if want_values else None
    idx = None
    if idx_mode == _native.CODES_INT8:
        idx = torch.empty_like(xd, dtype=torch.uint8)
    elif idx_mode == _native.CODES_INT4:
        idx = torch.empty((n + 1) // 2, dtype=torch.uint8, device=xd.device)
    if n:
        thr = _param_on(threshold.contiguous(), xd.device)
        _lut_launch(xd, y, idx, idx_mode, table, K, C, inner, thr, eps, False, 1.0, 1.0, False)
    if y is not None and y.dim() == 0:
        y = y.reshape(1)        # the reference multiplies by the threshold TENSOR of shape (1,): a 0-dim input comes back 1-D
    return y, idx


def lut_weights_direct(x, table, K, threshold, per_channel, axis, eps, cache):
    """Lean launch of the prepared LUT kernel for a weight quantizer (no dispatcher, no per-call validation or cache-key
    building: ~35 us -> ~10 us of host time per call, which matters once a 45 M-element bf16 matrix takes 40 us on the
    device).  `cache` is a dict owned by the quantizer: (shape, dtype, device, thr ptr) -> launch constants.  Anything
    unusual (non-contiguous input, no prepared blob, misaligned view) goes through the general path."""
    if x.is_contiguous() and x.numel() and x.dim():        # (0-dim inputs: general path, which returns them 1-D like the reference)
        key = (x.shape, x.dtype, x.device, threshold.data_ptr(), _ver(threshold))
        hit = cache.get(key)
        if hit is None:
            hit = False
            tag = _DT.get(x.dtype)
            if tag is not None and threshold.dtype == torch.float32 and (not per_channel or threshold.numel() == x.shape[axis]) \
                    and (per_channel or threshold.numel() == 1):
                if per_channel:
                    _, C, inner = _channel_layout(x, axis)
                else:
                    C, inner = 1, 1
                thr = _param_on(threshold.contiguous(), x.device)
                prep = _prepared_for(table, K, x.device, thr, eps, False, 1.0, 1.0, 0)
                if prep is not None and prep[1] is not None:
                    hit = (tag, int(K), prep[2], prep[3], C, inner, prep[1].data_ptr(), (thr, prep), x.device.index)
            if len(cache) > 16:
                cache.clear()
            if hit or not torch.cuda.is_current_stream_capturing():     # a capture only postpones the preparation
                cache[key] = hit
        if hit:
            tag, K_, bw, signed, C, inner, blob_ptr, _, index = hit
            if (x.data_ptr() & 15) == 0:
                y = torch.empty(x.shape, dtype=torch.float32, device=x.device)
                lib = _native._lib or _native.load()
                prev = _get_device()
                if prev != index:
                    _set_device(index)
                fast = _native.fast
                rc = (fast.fq_lut_prepared if fast is not None else lib.mctq_fq_lut_prepared)(
                    x.data_ptr(), y.data_ptr(), None, x.numel(), tag, blob_ptr, K_, bw, signed, C, inner, 0, 0, _raw_stream(index))
                if prev != index:
                    _set_device(prev)
                if rc == 0:
                    return y
    return _lut_tensor_cuda(x, table, K, threshold, per_channel, axis, eps)


def _lut_tensor_cuda(x, table, K, threshold, per_channel, axis, eps):
    return _lut_tensor_launch(x, table, K, threshold, per_channel, axis, eps, 0, True)[0]


def _lut_indices_cuda(x, table, K, threshold, per_channel, axis, eps, idx_mode):
    if idx_mode not in (_native.CODES_INT8, _native.CODES_INT4):
        raise RuntimeError("idx_mode must be 1 (uint8) or 2 (packed 4-bit)")
    return _lut_tensor_launch(x, table, K, threshold, per_channel, axis, eps, idx_mode, False)[1]


def _lut_scalar_cuda(x, table, K, divisor, threshold, round_to_input_dtype, multiply=False):
    """`multiply`: normalise with x * (float)(1.0 / divisor) -- the reference's CUDA flavour of `tensor / python_number`
    (quantizer_utils.reference_arithmetic) -- instead of the true division x / (float)divisor."""
    _dtype_tag(x)
    xd = x.contiguous()        # the reference's eager composition (argmin + gather) returns a row-major tensor whatever x's strides
    y = torch.empty_like(xd, dtype=torch.float32)
    if xd.numel():
        if multiply:
            _lut_launch(xd, y, None, 0, table, K, 1, 1, None, 0.0, 2, float(np.float32(1.0 / divisor)), threshold, round_to_input_dtype)
        else:
            _lut_launch(xd, y, None, 0, table, K, 1, 1, None, 0.0, 1, divisor, threshold, round_to_input_dtype)
    return y


class private_stream:
    """Context manager for back-to-back calls on tensors that ALREADY EXIST (per-layer weight quantization of a model,
    sweeps over recorded activations): inside the block consecutive launches of this package on a stream may OVERLAP.  A
    launch whose buffers are disjoint from those of every launch still in flight does not wait for its predecessor at all
    (at most two such launches in a row), one whose input is merely not produced by a launch in flight issues its loads
    before it waits -- the drain of one kernel and the ramp of the next disappear (programmatic dependent launch, orders
    "free" and "early" in csrc/mctq_common.cuh; every dependency between this package's own launches is detected from
    their address ranges and honoured).

    The caller guarantees that, inside the block, the streams used carry NO OTHER work than this package's calls: the
    library cannot see foreign kernels (a cuDNN kernel may release its dependents before its stores are visible; the
    caching allocator may recycle a foreign kernel's input buffer as one of our outputs).  Entering and leaving the block
    make the next launch on every stream wait for everything before it.  Process-wide setting."""

    def __enter__(self):
        self._prev = _native.load().mctq_set_tuning(3, 3)
        return self

    def __exit__(self, *exc):
        _native.load().mctq_set_tuning(3, self._prev)
        return False


# --------------------------------------------------------------------------------------------- CPU-tensor impls
# A host tensor is streamed through the GPU in chunks (pinned memory overlaps copies and kernels).
class host_pipeline:
    """Context manager: quantizer / holder / wrapper calls on HOST (pinned) tensors inside the block are only
    *enqueued* -- their results are complete when the block exits.  The H2D -> kernel -> D2H pipeline then keeps
    running across the tensors of a model instead of filling and draining once per call (mctq_host_set_deferred /
    mctq_host_wait).  Inputs must not be modified inside the block; the block keeps them (and the outputs) alive.

        with mct_quantizers_b200.host_pipeline():
            outs = [holder(x) for holder, x in zip(holders, pinned_host_tensors)]
        # outs are complete here
    """
    _active = {}          # device index -> innermost active scope

    def __init__(self, device=None):
        self._device = device
        self._keep = []
        self._outer = None

    def __enter__(self):
        dev = _require_cuda_for_host_path() if self._device is None else torch.device(self._device).index
        self._dev = dev
        self._outer = host_pipeline._active.get(dev)
        if self._outer is None:
            _native.check(_native.load().mctq_host_set_deferred(dev, 1), "mctq_host_set_deferred")
        host_pipeline._active[dev] = self
        return self

    def wait(self):
        """Complete everything enqueued so far (the block stays deferred)."""
        _native.check(_native.load().mctq_host_wait(self._dev), "mctq_host_wait")
        self._keep.clear()

    def __exit__(self, *exc):
        try:
            if self._outer is None:
                _native.check(_native.load().mctq_host_set_deferred(self._dev, 0), "mctq_host_set_deferred")   # waits
            else:
                self.wait()
        finally:
            self._keep.clear()
            if self._outer is None:
                host_pipeline._active.pop(self._dev, None)
            else:
                host_pipeline._active[self._dev] = self._outer
        return False


def _host_keepalive(dev, *tensors):
    scope = host_pipeline._active.get(dev)
    if scope is not None:
        scope._keep.extend(tensors)


def _affine_host(x, scale_np, zp_np, C, inner, quant_min, quant_max):
    dev = _require_cuda_for_host_path()
    tag = _dtype_tag(x)
    xc = x.contiguous()
    y = torch.empty_like(xc, pin_memory=xc.is_pinned())
    if xc.numel():
        lib = _native.load()
        stg = _staging_buffer(dev)
        rc = lib.mctq_fq_affine_host(_ptr(xc), _ptr(y), xc.numel(), tag, scale_np.ctypes.data_as(c_vp),
                                     zp_np.ctypes.data_as(c_vp), C, inner, int(quant_min), int(quant_max),
                                     _ptr(stg), stg.numel(), dev)
        _native.check(rc, "mctq_fq_affine_host")
        _host_keepalive(dev, xc, y)
    return y


def _affine_scalar_cpu(x, scale, zero_point, quant_min, quant_max):
    _check_range(quant_min, quant_max)
    if not quant_min <= zero_point <= quant_max:
        raise RuntimeError("`zero_point` must be between `quant_min` and `quant_max`.")
    return _affine_host(x, np.array([scale], dtype=np.float64).astype(np.float32), np.array([zero_point], dtype=np.int32),
                        1, 1, quant_min, quant_max)


def _affine_tensor_cpu(x, scale, zero_point, quant_min, quant_max):
    _check_range(quant_min, quant_max)
    if scale.numel() != 1 or zero_point.numel() != 1:
        raise RuntimeError(f"a Tensor with {max(scale.numel(), zero_point.numel())} elements cannot be converted to Scalar")
    return _affine_host(x, _host_array(scale, np.float32), _host_array(zero_point, np.int32), 1, 1, quant_min, quant_max)


def _affine_channel_cpu(x, scale, zero_point, axis, quant_min, quant_max):
    _check_range(quant_min, quant_max)
    _check_channel_params(x, scale, zero_point, axis)
    xc = x.contiguous()
    _, C, inner = _channel_layout(xc, axis)
    return _affine_host(xc, _host_array(scale, np.float32), _host_array(zero_point, np.int32), C, inner,
                        quant_min, quant_max)


def _lut_host(x, table, K, thr_np, C, inner, eps, scalar_mode, divisor, thr_f32, round_flag):
    dev = _require_cuda_for_host_path()
    tag = _dtype_tag(x)
    xc = x.contiguous()
    y = torch.empty(xc.shape, dtype=torch.float32, pin_memory=xc.is_pinned())
    if xc.numel():
        lib = _native.load()
        stg = _staging_buffer(dev)
        tab = table.detach().cpu().contiguous()
        rc = lib.mctq_fq_lut_host(_ptr(xc), _ptr(y), xc.numel(), tag, _ptr(tab), int(K),
                                  thr_np.ctypes.data_as(c_vp) if thr_np is not None else None, C, inner,
                                  float(np.float32(eps)), int(scalar_mode), float(np.float32(divisor)),
                                  float(np.float32(thr_f32)), int(round_flag), _ptr(stg), stg.numel(), dev)
        _native.check(rc, "mctq_fq_lut_host")
        _host_keepalive(dev, xc, y)
    return y


def _lut_tensor_cpu(x, table, K, threshold, per_channel, axis, eps):
    xc = x.contiguous()
    if per_channel:
        if threshold.numel() != xc.shape[axis]:
            raise RuntimeError("threshold length does not match the channel axis")
        _, C, inner = _channel_layout(xc, axis)
    else:
        C, inner = 1, 1
    return _lut_host(xc, table, K, _host_array(threshold, np.float32), C, inner, eps, 0, 1.0, 1.0, 0)


def _lut_scalar_cpu(x, table, K, divisor, threshold, round_to_input_dtype, multiply=False):
    # host tensors: the reference would run libtorch's CPU kernels, i.e. a true division, whatever `multiply` says
    return _lut_host(x, table, K, None, 1, 1, 0.0, 1, divisor, threshold, round_to_input_dtype)


def _via_device(cuda_impl, keep=()):
    """Host-tensor implementation of the operators that have no chunked host pipeline of their own (integer codes, LUT
    indices, dequantization, the producer-fused holder): every tensor argument is copied to the current CUDA device, the
    device operator runs there and the results come back as host tensors.  Codes are the cheap direction of this trip:
    one byte (or a nibble) per element returns instead of four.  `keep`: positions of tensor arguments that stay on the
    host (the LUT search table is host data by contract)."""
    def impl(*args):
        dev = torch.device("cuda", _require_cuda_for_host_path())

        def up(a):
            if isinstance(a, torch.Tensor) and not a.is_cuda:
                return a.to(dev, non_blocking=a.is_pinned())
            return a

        out = cuda_impl(*[a if k in keep else up(a) for k, a in enumerate(args)])
        if isinstance(out, (tuple, list)):
            return type(out)(o.cpu() if isinstance(o, torch.Tensor) else o for o in out)
        return out.cpu() if isinstance(out, torch.Tensor) else out
    return impl


_LIB.impl("fq_affine_scalar", _affine_scalar_cuda, "CUDA")
_LIB.impl("fq_affine_tensor", _affine_tensor_cuda, "CUDA")
_LIB.impl("fq_affine_channel", _affine_channel_cuda, "CUDA")
_LIB.impl("fq_lut_tensor", _lut_tensor_cuda, "CUDA")
_LIB.impl("fq_lut_scalar", _lut_scalar_cuda, "CUDA")
_LIB.impl("quantize_affine_channel", _quantize_affine_channel_cuda, "CUDA")
_LIB.impl("dequantize_affine", _dequantize_affine_cuda, "CUDA")
_LIB.impl("lut_indices", _lut_indices_cuda, "CUDA")

_LIB.impl("fq_affine_scalar_pre", _affine_scalar_pre_cuda, "CUDA")
_LIB.impl("fq_affine_scalar_pre", _via_device(_affine_scalar_pre_cuda), "CPU")
_LIB.impl("fq_affine_scalar", _affine_scalar_cpu, "CPU")
_LIB.impl("fq_affine_tensor", _affine_tensor_cpu, "CPU")
_LIB.impl("fq_affine_channel", _affine_channel_cpu, "CPU")
_LIB.impl("fq_lut_tensor", _lut_tensor_cpu, "CPU")
_LIB.impl("fq_lut_scalar", _lut_scalar_cpu, "CPU")
_LIB.impl("quantize_affine_channel", _via_device(_quantize_affine_channel_cuda), "CPU")
_LIB.impl("dequantize_affine", _via_device(_dequantize_affine_cuda), "CPU")
_LIB.impl("lut_indices", _via_device(_lut_indices_cuda, keep=(1,)), "CPU")


# --------------------------------------------------------------------------------------------- fake / meta impls
@torch.library.register_fake("mctq::fq_affine_scalar")
def _(x, scale, zero_point, quant_min, quant_max):
    return torch.empty_like(x)


@torch.library.register_fake("mctq::fq_affine_scalar_pre")
def _(x, other, pre_op, scale, zero_point, quant_min, quant_max):
    return torch.empty_like(x)


@torch.library.register_fake("mctq::fq_affine_tensor")
def _(x, scale, zero_point, quant_min, quant_max):
    return torch.empty_like(x)


@torch.library.register_fake("mctq::fq_affine_channel")
def _(x, scale, zero_point, axis, quant_min, quant_max):
    return torch.empty_like(x)


@torch.library.register_fake("mctq::fq_lut_tensor")
def _(x, table, K, threshold, per_channel, axis, eps):
    return torch.empty_like(x, dtype=torch.float32)


@torch.library.register_fake("mctq::fq_lut_scalar")
def _(x, table, K, divisor, threshold, round_to_input_dtype, multiply=False):
    return torch.empty_like(x, dtype=torch.float32)


@torch.library.register_fake("mctq::quantize_affine_channel")
def _(x, scale, zero_point, axis, quant_min, quant_max, code_mode, want_values):
    if code_mode == _native.CODES_INT4:
        codes = x.new_empty(((x.numel() + 1) // 2,), dtype=torch.uint8)
    else:
        codes = torch.empty_like(x, dtype=torch.int8 if quant_min < 0 else torch.uint8)
    return codes, (torch.empty_like(x) if want_values else x.new_empty((0,)))


@torch.library.register_fake("mctq::dequantize_affine")
def _(codes, code_mode, is_signed, shape, scale, zero_point, axis):
    return codes.new_empty(list(shape), dtype=torch.float32)


@torch.library.register_fake("mctq::lut_indices")
def _(x, table, K, threshold, per_channel, axis, eps, idx_mode):
    if idx_mode == _native.CODES_INT4:
        return x.new_empty(((x.numel() + 1) // 2,), dtype=torch.uint8)
    return torch.empty_like(x, dtype=torch.uint8)


# --------------------------------------------------------------------------------------------- whole-model launch
class MultiTensorPlan:
    """One-launch fake-quant of many weight tensors (C ABI: mctq_fq_affine_multi).

    Build once from [(x, scale, zp, axis, qmin, qmax)] (all CUDA, same device): outputs are allocated here and
    reused by every run(); the descriptor table is uploaded once.  run() enqueues ONE kernel on the current
    stream.  Reference loop being replaced: quantize_wrapper.py:228-240 / :260-270."""

    @staticmethod
    def accepts(item):
        """Whether a tensor can be part of a plan: CUDA, supported dtype, non-empty, dense in memory (the plan captures
        pointers, so no hidden copies), parameters of the right length."""
        x, scale, zp, axis, qmin, qmax = item
        if not x.is_cuda or x.dtype not in _DT or x.numel() == 0 or not _is_dense_permutation(x):
            return False
        try:
            if axis is None:
                if scale.numel() != 1 or zp.numel() != 1:
                    return False
                xd = _dense(x)
            else:
                _check_channel_params(x, scale, zp, axis)
                xd = _channel_layout(x, axis)[0]
        except (RuntimeError, IndexError):
            return False
        return xd.data_ptr() == x.data_ptr() and scale.dtype == torch.float32 and zp.dtype == torch.int32

    def __init__(self, items):
        lib = _native.load()
        if not items:
            raise ValueError("MultiTensorPlan needs at least one tensor")
        self.device = items[0][0].device
        self.inputs, self.outputs, self._keep = [], [], []
        descs = (_native.MctqTensorDesc * len(items))()
        for k, (x, scale, zp, axis, qmin, qmax) in enumerate(items):
            if x.device != self.device or not x.is_cuda:
                raise ValueError("all tensors of a MultiTensorPlan must live on one CUDA device")
            _check_range(qmin, qmax)
            if axis is None:
                xd, C, inner = _dense(x), 1, 1
                if scale.numel() != 1 or zp.numel() != 1:
                    raise RuntimeError("per-tensor entry needs one scale / zero-point")
            else:
                _check_channel_params(x, scale, zp, axis)
                xd, C, inner = _channel_layout(x, axis)
            if xd.data_ptr() != x.data_ptr():
                raise ValueError("MultiTensorPlan needs dense tensors (no hidden copies: the plan captures pointers)")
            y = torch.empty_like(xd)
            s, z = _param_on(scale.contiguous(), self.device), _param_on(zp.contiguous(), self.device)
            self._keep += [s, z]
            self.inputs.append(xd)
            self.outputs.append(y)
            d = descs[k]
            d.x, d.y, d.codes, d.scale, d.zp = xd.data_ptr(), y.data_ptr(), None, s.data_ptr(), z.data_ptr()
            d.n, d.C, d.inner = xd.numel(), C, inner
            d.qmin, d.qmax, d.dtype, d.code_mode = int(qmin), int(qmax), _dtype_tag(xd), 0
        starts = (ctypes.c_int32 * (len(items) + 1))()
        total = lib.mctq_multi_plan(ctypes.cast(descs, c_vp), len(items), ctypes.cast(starts, c_vp))
        if total < 0:
            raise MctqError(f"mctq_multi_plan: {total}")
        self.total_tiles = int(total)
        self.n_desc = len(items)
        self._descs_dev = torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8).to(self.device)
        self._starts_dev = torch.frombuffer(bytearray(bytes(starts)), dtype=torch.uint8).to(self.device)

    def run(self):
        lib = _native.load()
        with _on_device(self.device):
            rc = lib.mctq_fq_affine_multi(_ptr(self._descs_dev), _ptr(self._starts_dev), self.n_desc, self.total_tiles,
                                          _stream(self.device))
        _native.check(rc, "mctq_fq_affine_multi")
        return self.outputs


class ScalarSitesPlan:
    """One-launch per-tensor fake-quant of many tensors (C ABI: mctq_fq_affine_scalar_multi).

    Build once from [(x, scale, zero_point, qmin, qmax)] (CUDA tensors of ONE device, dense, 16-byte aligned; scale a Python
    float, narrowed to f32 like ATen does): outputs are allocated here and reused by every run(); the site table is a host
    array that travels as kernel parameters.  run() enqueues ONE kernel per 64 sites on the current stream.
    Reference loop being replaced: one PytorchActivationQuantizationHolder.forward per site
    (activation_quantization_holder.py:43-53)."""

    @staticmethod
    def accepts(x):
        return (x.is_cuda and x.dtype in _DT and x.numel() > 0 and _is_dense_permutation(x) and (x.data_ptr() & 15) == 0)

    def __init__(self, items):
        if not items:
            raise ValueError("ScalarSitesPlan needs at least one tensor")
        self.device = items[0][0].device
        self.inputs, self.outputs = [], []
        self._sites = (_native.MctqSiteDesc * len(items))()
        for k, (x, scale, zp, qmin, qmax) in enumerate(items):
            if x.device != self.device or not self.accepts(x):
                raise ValueError(f"tensor {k} cannot be part of a ScalarSitesPlan (see ScalarSitesPlan.accepts)")
            _check_range(qmin, qmax)
            if not int(qmin) <= int(zp) <= int(qmax):
                raise RuntimeError("`zero_point` must be between `quant_min` and `quant_max`.")
            y = torch.empty_like(x)                         # preserve_format: a dense permutation keeps its strides
            if y.stride() != x.stride():
                raise ValueError(f"tensor {k}: output strides differ from the input's")
            d = self._sites[k]
            d.x, d.y, d.n, d.dtype = x.data_ptr(), y.data_ptr(), x.numel(), _DT[x.dtype]
            d.scale, d.zp, d.qmin, d.qmax = float(np.float32(scale)), int(zp), int(qmin), int(qmax)
            self.inputs.append(x)
            self.outputs.append(y)
        self.n_sites = len(items)
        self._sites_ptr = ctypes.cast(self._sites, c_vp)
        self._sites_addr = ctypes.addressof(self._sites)

    def run(self):
        lib = _native._lib or _native.load()
        index = self.device.index
        prev = _get_device()
        if prev != index:
            _set_device(index)
        fast = _native.fast
        if fast is not None:
            rc = fast.fq_affine_scalar_multi(self._sites_addr, self.n_sites, _raw_stream(index))
        else:
            rc = lib.mctq_fq_affine_scalar_multi(self._sites_ptr, self.n_sites, _raw_stream(index))
        if prev != index:
            _set_device(prev)
        if rc:
            _native.check(rc, "mctq_fq_affine_scalar_multi")
        return self.outputs


class LutMultiPlan:
    """One-launch LUT fake-quant of many weight tensors (C ABI: mctq_lut_multi_plan + mctq_fq_lut_prepared_multi).

    Build once from [(x, table, K, threshold, per_channel, axis, eps)] (all CUDA, one device; `table` is the host search
    table of the quantizer): every tensor's decision tables are prepared (or taken from the per-quantizer cache), outputs
    are allocated here and reused by every run(); the plan travels as kernel parameters.  run() enqueues ONE kernel
    per kernel variant and 64 tensors on the current stream -- no launch gaps and no per-launch tails between the tensors
    (Llama-7B: 96 launches -> 2).
    `accepts(item)` tells whether a tensor can be part of a plan (prepared path available, dense, aligned)."""

    @staticmethod
    def _describe(item, device, y=None):
        x, table, K, threshold, per_channel, axis, eps = item
        if not x.is_cuda or x.device != device or x.dtype not in _DT or x.numel() == 0:
            return None
        if threshold.dtype != torch.float32:
            return None
        if per_channel:
            if threshold.numel() != x.shape[axis]:
                return None
            xd, C, inner = _channel_layout(x, axis)
        else:
            if threshold.numel() != 1:
                return None
            xd, C, inner = _dense(x), 1, 1
        if xd.data_ptr() != x.data_ptr():
            return None                                    # the plan captures pointers: no hidden copies
        thr = _param_on(threshold.contiguous(), device)
        prep = _prepared_for(table, K, device, thr, eps, False, 1.0, 1.0, 0)
        if prep is None or prep[1] is None:
            return None
        d = _native.MctqLutTensorDesc()
        d.x, d.y, d.prepared_dev = xd.data_ptr(), (y.data_ptr() if y is not None else xd.data_ptr()), prep[1].data_ptr()
        d.n, d.C, d.inner = xd.numel(), C, inner
        d.dtype, d.K, d.lut_values_bitwidth, d.is_signed = _DT[xd.dtype], int(K), prep[2], prep[3]
        return d, xd, (thr, prep)

    @staticmethod
    def accepts(item):
        x = item[0]
        if not x.is_cuda:
            return False
        got = LutMultiPlan._describe(item, x.device)
        if got is None:
            return False
        lib = _native.load()
        descs = (_native.MctqLutTensorDesc * 1)(got[0])     # y stands in as x here: fresh outputs are at least as aligned
        return lib.mctq_lut_multi_plan_bytes(ctypes.cast(descs, c_vp), 1) > 0

    def __init__(self, items):
        lib = _native.load()
        if not items:
            raise ValueError("LutMultiPlan needs at least one tensor")
        self.device = items[0][0].device
        self.inputs, self.outputs, self._keep = [], [], []
        descs = (_native.MctqLutTensorDesc * len(items))()
        for k, item in enumerate(items):
            x = item[0]
            if not x.is_cuda or x.device != self.device:
                raise ValueError("all tensors of a LutMultiPlan must live on one CUDA device")
            y = torch.empty(x.shape, dtype=torch.float32, device=self.device)
            if not x.is_contiguous():
                y = torch.empty_like(_dense(x), dtype=torch.float32)
            got = self._describe(item, self.device, y)
            if got is None:
                raise ValueError(f"tensor {k} cannot be part of a LutMultiPlan (see LutMultiPlan.accepts)")
            descs[k] = got[0]
            self.inputs.append(got[1])
            self.outputs.append(y)
            self._keep.append(got[2])
        nb = lib.mctq_lut_multi_plan_bytes(ctypes.cast(descs, c_vp), len(items))
        if nb == 0:
            raise MctqError("mctq_lut_multi_plan_bytes: a tensor of the plan is outside the prepared LUT path")
        self._plan_host = (ctypes.c_uint8 * nb)()
        total = lib.mctq_lut_multi_plan(ctypes.cast(descs, c_vp), len(items), ctypes.cast(self._plan_host, c_vp), nb)
        if total <= 0:
            raise MctqError(f"mctq_lut_multi_plan: {total}")
        self.total_tiles = int(total)
        self.n_desc = len(items)
        self.n_launches = int(np.frombuffer(bytes(self._plan_host[:16]), dtype=np.int32)[2])    # one per kernel variant and 64 tensors

    def run(self):
        lib = _native.load()
        with _on_device(self.device):
            rc = lib.mctq_fq_lut_prepared_multi(ctypes.cast(self._plan_host, c_vp), _stream(self.device))
        _native.check(rc, "mctq_fq_lut_prepared_multi")
        return self.outputs

