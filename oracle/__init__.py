"""CPU oracle for the fake-quant hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``mct_quantizers_b200/`` may import this package.  Legitimate users:
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py``.  Parity status: pinned against fixtures generated from the unmodified reference
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``; checked by ``tests/test_oracle_golden.py``).

Two restatements live here:

* ``mctq_oracle.c`` (this module's ctypes API): plain C, no torch, follows the arithmetic spec of
  SURVEY.md section 8a.
* ``torch_cpu_port.py``: the reference's call sites re-expressed on CPU torch ops (the ATen
  ``fake_quantize_*`` ops the reference delegates to, and the eager LUT composition); it is the
  timed CPU baseline because BASELINE.json's metric is quoted "vs CPU torch".
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "mctq_oracle.c")
_OUT_DIR = os.path.join(_HERE, "_build")
_LIB_PATH = os.path.join(_OUT_DIR, "libmctq_oracle.so")

F32, BF16, F16 = 0, 1, 2
_lib = None


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (seconds).  Returns the .so path."""
    os.makedirs(_OUT_DIR, exist_ok=True)
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(_SRC)):
        return _LIB_PATH
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-fno-fast-math",
           "-std=c11", "-o", _LIB_PATH, _SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        path = build()
        L = ctypes.CDLL(path)
        vp, i64, i32, dbl = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_double
        L.mctq_oracle_fq_affine.argtypes = [vp, vp, vp, i64, ctypes.c_int, vp, vp, i64, i64, i32, i32]
        L.mctq_oracle_fq_affine.restype = ctypes.c_int
        L.mctq_oracle_fq_lut.argtypes = [vp, vp, vp, i64, ctypes.c_int, vp, ctypes.c_int, vp, i64, i64,
                                         ctypes.c_int, ctypes.c_int, dbl, ctypes.c_int, dbl]
        L.mctq_oracle_fq_lut.restype = ctypes.c_int
        L.mctq_oracle_dequant_affine.argtypes = [vp, vp, i64, vp, vp, i64, i64]
        L.mctq_oracle_dequant_affine.restype = ctypes.c_int
        L.mctq_oracle_cvt.argtypes = [vp, vp, i64, ctypes.c_int]
        L.mctq_oracle_uncvt.argtypes = [vp, vp, i64, ctypes.c_int]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _as_raw(x, dtype):
    """x: float32 ndarray (dtype F32) or uint16 bit patterns (BF16 / F16)."""
    want = np.float32 if dtype == F32 else np.uint16
    a = np.ascontiguousarray(x)
    if a.dtype != want:
        raise TypeError(f"oracle expects {want} storage for dtype tag {dtype}, got {a.dtype}")
    return a


def channel_layout(shape, channel_axis):
    """(C, inner) of the [outer][C][inner] view for a contiguous tensor of `shape`."""
    if channel_axis is None:
        return 1, 1
    ax = channel_axis % len(shape)
    inner = 1
    for s in shape[ax + 1:]:
        inner *= int(s)
    return int(shape[ax]), inner


def fq_affine(x, dtype, scale, zp, C, inner, qmin, qmax, want_codes=False):
    """Affine fake-quant.  Returns y (same storage dtype as x) and optionally int32 codes."""
    a = _as_raw(x, dtype)
    scale = np.ascontiguousarray(scale, dtype=np.float32).reshape(-1)
    zp = np.ascontiguousarray(zp, dtype=np.int32).reshape(-1)
    assert scale.size == C and zp.size == C
    y = np.empty_like(a)
    codes = np.empty(a.shape, dtype=np.int32) if want_codes else None
    rc = lib().mctq_oracle_fq_affine(_ptr(a), _ptr(y), _ptr(codes), a.size, dtype, _ptr(scale), _ptr(zp),
                                     C, inner, int(qmin), int(qmax))
    assert rc == 0
    return (y, codes) if want_codes else y


def fq_lut(x, dtype, lut, thr, C, inner, bw, signed, eps, activation_mode=False, want_idx=False, cuda_flavour=False):
    """LUT fake-quant.  `thr`: f32 array [C] (weights mode) or a Python float (activation mode).
    `cuda_flavour` (activation mode): normalise with x * f32(1 / (thr + eps)) -- what the reference computes for CUDA
    tensors -- instead of the CPU kernel's true division.  Returns f32 y (and int32 LUT indices)."""
    a = _as_raw(x, dtype)
    lut = np.ascontiguousarray(lut, dtype=np.float32).reshape(-1)
    y = np.empty(a.shape, dtype=np.float32)
    idx = np.empty(a.shape, dtype=np.int32) if want_idx else None
    if activation_mode:
        thr_arr = np.array([thr], dtype=np.float32)
        thr_scalar = float(thr)
        assert C == 1
    else:
        thr_arr = np.ascontiguousarray(thr, dtype=np.float32).reshape(-1)
        thr_scalar = 0.0
        assert thr_arr.size == C
    rc = lib().mctq_oracle_fq_lut(_ptr(a), _ptr(y), _ptr(idx), a.size, dtype, _ptr(lut), lut.size,
                                  _ptr(thr_arr), C, inner, int(bw), int(bool(signed)), float(eps),
                                  (2 if cuda_flavour else 1) if activation_mode else 0, thr_scalar)
    assert rc == 0
    return (y, idx) if want_idx else y


def dequant_affine(codes, scale, zp, C, inner):
    codes = np.ascontiguousarray(codes, dtype=np.int32)
    scale = np.ascontiguousarray(scale, dtype=np.float32).reshape(-1)
    zp = np.ascontiguousarray(zp, dtype=np.int32).reshape(-1)
    y = np.empty(codes.shape, dtype=np.float32)
    lib().mctq_oracle_dequant_affine(_ptr(codes), _ptr(y), codes.size, _ptr(scale), _ptr(zp), C, inner)
    return y


def f32_to_half_bits(x, dtype):
    a = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty(a.shape, dtype=np.uint16)
    lib().mctq_oracle_cvt(_ptr(a), _ptr(out), a.size, dtype)
    return out


def half_bits_to_f32(x, dtype):
    a = np.ascontiguousarray(x, dtype=np.uint16)
    out = np.empty(a.shape, dtype=np.float32)
    lib().mctq_oracle_uncvt(_ptr(a), _ptr(out), a.size, dtype)
    return out
