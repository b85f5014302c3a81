"""Host-side helpers of the PyTorch quantizers (reference: mct_quantizers/pytorch/quantizer_utils.py).

`lut_quantizer` / `int_quantization_with_threshold` (reference :95-170) are NOT eager torch compositions
here: they go through the `mctq` custom operators, i.e. the fused sm_100a LUT kernel.  Range fixing
(`fix_range_to_include_zero`, reference :60-92) stays on the host -- it runs once per quantizer -- and is
evaluated on CPU f32 tensors with exactly the reference's sequence of f32 operations so that derived scales
and zero points are bit-identical.
"""
import os
from typing import Tuple

import numpy as np
import torch

from mct_quantizers_b200.logger import Logger


def get_working_device():
    """'cuda' (the current device) when a GPU is visible, else 'cpu' (reference :23-31)."""
    return torch.device('cuda' if torch.cuda.is_available() else 'cpu')


def to_torch_tensor(tensor):
    """numpy / list / tuple / python scalar -> torch tensor on the working device (reference :34-57)."""
    device = get_working_device()
    if isinstance(tensor, torch.Tensor):
        return tensor.to(device)
    if isinstance(tensor, list):
        return [to_torch_tensor(t) for t in tensor]
    if isinstance(tensor, tuple):
        return (to_torch_tensor(t) for t in tensor)
    if isinstance(tensor, np.ndarray):
        return torch.from_numpy(tensor.astype(np.float32)).to(device)
    if isinstance(tensor, float):
        return torch.Tensor([tensor]).to(device)
    if isinstance(tensor, int):
        return torch.Tensor([tensor]).int().to(device)
    raise Exception(f'Conversion of type {type(tensor)} to {type(torch.Tensor)} is not supported')


# The reference derives its parameters with torch ops on `get_working_device()`, and ONE of those ops is not the same
# function on the two devices: `tensor / python_number` is a true IEEE division on CPU, but libtorch's CUDA kernel
# multiplies by the f32 reciprocal of the number (aten/native/cuda/BinaryDivTrueKernel.cu).  For `/ (2 ** n_bits - 1)`
# (range fixing, reference :76; WeightsUniform scales, weights_uniform_inferable_quantizer.py:123) the two differ by one
# ulp for most ranges -- and the truncated zero point then differs by one in rare cases.  So the unmodified reference
# gives (slightly) different quantized models on a CUDA machine and on a CPU machine.  This package computes parameters
# on the host and reproduces the CPU flavour by default (what the golden fixtures pin; the same numbers on every
# machine); `reference_arithmetic("cuda")` reproduces, bit for bit, what the reference computes when a GPU is visible
# (checked against the reference running on the B200: tools/differential_fuzz.py), "auto" follows the machine like the
# reference does.  Divisions by powers of two (symmetric / POT quantizers, LUT scaling) are exact either way.  The third
# site is in the inference path itself: ActivationLutPOT normalises with `tensor / (threshold + eps)` on every call
# (reference :145-170 via int_quantization_with_threshold); for CUDA inputs that is x * (float)(1.0 / (thr + eps)), which
# moves exact rounding ties between two centroids (lut_quantizer below; kernels: MCTQ_LUT_DIVISOR_IS_MULTIPLIER).
_REFERENCE_ARITHMETIC = os.environ.get("MCTQ_REFERENCE_ARITHMETIC", "cpu")


def reference_arithmetic(mode: str = None) -> str:
    """Get / set the flavour of the reference's parameter arithmetic that constructors reproduce: "cpu" (default), "cuda",
    or "auto" (cuda when a GPU is visible, like the reference itself).  Returns the previous setting.  Affects quantizers
    constructed afterwards."""
    global _REFERENCE_ARITHMETIC
    prev = _REFERENCE_ARITHMETIC
    if mode is not None:
        if mode not in ("cpu", "cuda", "auto"):
            raise ValueError('reference_arithmetic: mode must be "cpu", "cuda" or "auto"')
        _REFERENCE_ARITHMETIC = mode
    return prev


def _cuda_flavour() -> bool:
    mode = _REFERENCE_ARITHMETIC
    return mode == "cuda" or (mode == "auto" and torch.cuda.is_available())


def div_by_python_number(t: torch.Tensor, k) -> torch.Tensor:
    """`t / k` for an f32 CPU tensor `t` and a Python number `k`, the way the reference's tensors would see it on the
    device flavour selected by reference_arithmetic()."""
    if _cuda_flavour():
        return t * float(np.float32(1.0 / float(k)))           # a * (float)(1.0 / b): reciprocal in double, narrowed to f32
    return t / k


def fix_range_to_include_zero(range_min: torch.Tensor, range_max: torch.Tensor, n_bits: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Shift [min, max] so that 0.0 falls on the quantization grid (f32 tensor arithmetic, round-half-even).

    Three cases per entry, selected with 0/1 masks exactly as the reference does (so that a strictly negative
    range yields max = -0.0 just like it):  min > 0 -> (0, max);  max < 0 -> (min, 0);  otherwise the grid is
    re-anchored at  step * round(min / step).  Note there is no final clamp (the numpy twin of the reference has
    one; the torch path that the quantizers use does not)."""
    lo_is_pos = range_min > 0
    hi_is_neg = range_max < 0
    straddles = torch.logical_and(torch.logical_not(lo_is_pos), torch.logical_not(hi_is_neg)).float()
    lo_is_pos, hi_is_neg = lo_is_pos.float(), hi_is_neg.float()

    step = div_by_python_number(range_max - range_min, 2 ** n_bits - 1)
    lo_adj = step * torch.round(range_min / step)
    hi_adj = range_max - range_min + lo_adj

    lo_adj = lo_adj * straddles + hi_is_neg * range_min
    hi_adj = hi_adj * straddles + lo_is_pos * range_max

    span = range_max - range_min
    if not torch.all(torch.isclose((lo_adj - range_min) / span, torch.tensor(0., device=span.device), atol=1e-6)):
        Logger.warning(f"Adjusting (min_range, max_range) from ({range_min},{range_max}) to ({lo_adj},{hi_adj})")
    return lo_adj, hi_adj


_TABLE_CACHE = {}


def lut_search_table(lut_values, lut_values_bitwidth: int, signed: bool) -> torch.Tensor:
    """Centroid list -> search-table blob (uint8 CPU tensor) consumed by the LUT kernels; cached by content."""
    from mct_quantizers_b200 import _native
    key = (tuple(float(v) for v in np.asarray(lut_values, dtype=np.float32).reshape(-1)), int(lut_values_bitwidth), bool(signed))
    tab = _TABLE_CACHE.get(key)
    if tab is None:
        blob = _native.build_lut_table(key[0], lut_values_bitwidth, signed)
        tab = torch.frombuffer(bytearray(blob), dtype=torch.uint8)
        _TABLE_CACHE[key] = tab
    return tab


def lut_quantizer_export(tensor_data: torch.Tensor, lut_values: torch.Tensor, signed: bool, threshold, lut_values_bitwidth: int,
                         eps: float, per_channel: bool = None, channel_axis: int = None, input_rank: int = None) -> torch.Tensor:
    """The LUT fake-quant spelt in elementary torch ops, for `torch.jit` tracing / ONNX export ONLY (the exporter needs
    ops it knows; the inference path is `lut_quantizer` below = one fused kernel).  Same op sequence as the reference's
    lut_quantizer + int_quantization_with_threshold (quantizer_utils.py:95-170): normalise by threshold + eps, scale to the
    2^lut_values_bitwidth grid, clip (no rounding), nearest centroid by argmin of |t - lut|, scale back by the threshold."""
    if per_channel:
        view = [1] * input_rank
        view[channel_axis] = -1
        threshold = torch.reshape(threshold, view)
    mult = 2 ** (lut_values_bitwidth - int(signed))
    lo, hi = (-2 ** (lut_values_bitwidth - 1), 2 ** (lut_values_bitwidth - 1) - 1) if signed else (0, 2 ** lut_values_bitwidth - 1)
    t = torch.clip((tensor_data / (threshold + eps)) * mult, min=lo, max=hi).unsqueeze(-1)
    nearest = torch.argmin(torch.abs(t - lut_values.reshape([1] * (t.dim() - 1) + [-1])), dim=-1)
    return (lut_values.flatten()[nearest] / mult) * threshold


def lut_quantizer(tensor_data: torch.Tensor,
                  lut_values: torch.Tensor,
                  signed: bool,
                  threshold,
                  lut_values_bitwidth: int,
                  eps: float,
                  per_channel: bool = None,
                  channel_axis: int = None,
                  input_rank: int = None,
                  _table: torch.Tensor = None) -> torch.Tensor:
    """Nearest-centroid fake-quant (same signature as the reference's lut_quantizer, :95-139):
    normalise by (threshold + eps) into the 2^lut_values_bitwidth grid, clip, pick the first nearest LUT
    entry, scale back by threshold.  `threshold` is an f32 tensor (weights) or a Python float (activations).
    One fused kernel; f32 output."""
    from mct_quantizers_b200 import ops
    K = int(lut_values.numel())
    table = _table if _table is not None else lut_search_table(lut_values.detach().cpu().numpy(), lut_values_bitwidth, signed)
    direct = ops.direct_ok(tensor_data)       # plain CUDA tensor and nobody tracing: skip the dispatcher (~10 us)
    if isinstance(threshold, torch.Tensor):
        thr = threshold.reshape(-1)
        if thr.dtype != torch.float32:
            thr = thr.float()
        if per_channel:
            if input_rank is not None and input_rank != tensor_data.dim():
                raise RuntimeError(f"input_rank is {input_rank} but the tensor has {tensor_data.dim()} dimensions")
            fn = ops._lut_tensor_cuda if direct else torch.ops.mctq.fq_lut_tensor
            return fn(tensor_data.detach() if direct else tensor_data, table, K, thr, True, int(channel_axis), float(eps))
        fn = ops._lut_tensor_cuda if direct else torch.ops.mctq.fq_lut_tensor
        return fn(tensor_data.detach() if direct else tensor_data, table, K, thr, False, 0, float(eps))
    # Python-float threshold: the divisor is formed in double and narrowed once, and half-precision inputs keep
    # their dtype through the normalisation (the reference's eager ops round after each step)
    divisor = float(threshold) + float(eps)
    fn = ops._lut_scalar_cuda if direct else torch.ops.mctq.fq_lut_scalar
    # round_to_input_dtype=True is a no-op for f32 inputs; passing the constant keeps the call fx-traceable.
    # Last argument: `tensor / python_number` on a CUDA tensor is a multiplication by the reciprocal in libtorch (see
    # reference_arithmetic above); the CUDA implementation of the operator reproduces exactly that when asked to, the
    # host-tensor implementation keeps the CPU kernel's true division (nothing here looks at the tensor: fx-traceable).
    return fn(tensor_data, table, K, divisor, float(threshold), True, _cuda_flavour())
