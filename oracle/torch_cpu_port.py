"""CPU-torch port of the reference's hot-path call sites -- TEST / BASELINE INFRASTRUCTURE ONLY.

For the seven affine quantizers the reference does no arithmetic of its own: it hands the tensor to
ATen's ``fake_quantize_per_{tensor,channel}_affine`` (third-party, part of libtorch).  Calling the
same ATen ops on CPU tensors therefore *is* the reference's CPU path, and that is what this module
does, one function per reference call site.  The two LUT families are an eager-op composition in
the reference (``pytorch/quantizer_utils.py:95-170``); it is re-expressed here op for op so that the
dtype promotions and rounding points are the ones CPU torch produces.

Used by: tests (second checker next to the C restatement), ``bench.py --impl reference`` and the
``cpu_baseline`` leg.  Never imported by ``mct_quantizers_b200``.
"""
import numpy as np
import torch


# --------------------------------------------------------------------------- parameter derivation
def symmetric_qparams(threshold, num_bits, signed):
    """base_symmetric_inferable_quantizer.py:49-60 -> (scales f64 ndarray, qmin, qmax)."""
    thr = np.asarray(threshold)
    if signed:
        return thr / 2 ** (num_bits - 1), -2 ** (num_bits - 1), 2 ** (num_bits - 1) - 1
    return thr / 2 ** num_bits, 0, 2 ** num_bits - 1


def weights_symmetric_qparams(threshold, num_bits):
    """weights_symmetric_inferable_quantizer.py:114-115 -> (scales f32 [C], zero_points i32 [C])."""
    scales, qmin, qmax = symmetric_qparams(threshold, num_bits, True)
    return (torch.from_numpy(scales.astype(np.float32)), torch.zeros(len(threshold), dtype=torch.int32),
            qmin, qmax)


def range_including_zero(range_min, range_max, n_bits):
    """pytorch/quantizer_utils.py:60-92 on f32 tensors (no final clamp, unlike the numpy twin)."""
    lo_pos = (range_min > 0)
    hi_neg = (range_max < 0)
    straddles = (~lo_pos & ~hi_neg).float()
    step = (range_max - range_min) / (2 ** n_bits - 1)
    lo_adj = step * torch.round(range_min / step)
    hi_adj = range_max - range_min + lo_adj
    lo_adj = lo_adj * straddles + hi_neg.float() * range_min
    hi_adj = hi_adj * straddles + lo_pos.float() * range_max
    return lo_adj, hi_adj


def weights_uniform_qparams(min_range, max_range, num_bits):
    """base_uniform_inferable_quantizer.py:55-66 + weights_uniform_inferable_quantizer.py:119-127."""
    lo = torch.from_numpy(np.asarray(min_range).astype(np.float32))
    hi = torch.from_numpy(np.asarray(max_range).astype(np.float32))
    lo, hi = range_including_zero(lo, hi, num_bits)
    scales = (hi - lo) / (2 ** num_bits - 1)
    zero_points = -(lo / scales).int()          # truncation toward zero, as in the reference
    return lo, hi, scales, zero_points, 0, 2 ** num_bits - 1


def activation_uniform_qparams(min_range, max_range, num_bits):
    """activation_uniform_inferable_quantizer.py:104-108 -> python floats / int."""
    lo = torch.from_numpy(np.asarray(min_range).astype(np.float32))
    hi = torch.from_numpy(np.asarray(max_range).astype(np.float32))
    lo, hi = range_including_zero(lo, hi, num_bits)
    lo, hi = lo[0].item(), hi[0].item()
    scale = float((hi - lo) / (2 ** num_bits - 1))
    zero_point = int(-np.round(lo / scale))
    return lo, hi, scale, zero_point, 0, 2 ** num_bits - 1


# --------------------------------------------------------------------------- affine call sites
def affine_scalar_qparams(x, scale, zero_point, qmin, qmax):
    """activation_symmetric_inferable_quantizer.py:113-117 / activation_uniform...:124-128."""
    with torch.no_grad():
        return torch.fake_quantize_per_tensor_affine(x, scale=float(scale), zero_point=int(zero_point),
                                                     quant_min=qmin, quant_max=qmax)


def affine_tensor_qparams(x, scales, zero_points, qmin, qmax):
    """weights_symmetric_inferable_quantizer.py:147-151 (per-tensor weights, 1-element tensors)."""
    return torch.fake_quantize_per_tensor_affine(x, scales, zero_points, quant_min=qmin, quant_max=qmax)


def affine_per_channel(x, scales, zero_points, axis, qmin, qmax):
    """weights_symmetric_inferable_quantizer.py:139-144 / weights_uniform...:153-158."""
    return torch.fake_quantize_per_channel_affine(x, scales.flatten(), zero_points.flatten(), axis=axis,
                                                  quant_min=qmin, quant_max=qmax)


# --------------------------------------------------------------------------- LUT call sites
def lut_fake_quant(x, lut_values, signed, threshold, lut_values_bitwidth, eps,
                   per_channel=None, channel_axis=None, input_rank=None, want_idx=False):
    """quantizer_utils.py:95-170.  `threshold`: f32 tensor (weights) or Python float (activations)."""
    if per_channel:
        shape = [1] * input_rank
        shape[channel_axis] = -1
        threshold = threshold.reshape(shape)
    frac_bits = lut_values_bitwidth - int(signed)
    if signed:
        lo, hi = -2 ** (lut_values_bitwidth - 1), 2 ** (lut_values_bitwidth - 1) - 1
    else:
        lo, hi = 0, 2 ** lut_values_bitwidth - 1
    t = torch.clip((x / (threshold + eps)) * (2 ** frac_bits), min=lo, max=hi)
    t = t.unsqueeze(-1)
    table = lut_values.reshape([1] * (t.dim() - 1) + [-1])
    idx = torch.argmin(torch.abs(t - table), dim=-1)
    y = (lut_values.flatten()[idx] / (2 ** frac_bits)) * threshold
    return (y, idx) if want_idx else y
