"""ONNX-export shims: autograd Functions whose `forward` is a plain torch formula and whose `symbolic` emits a
node of the `mct_quantizers` custom-op domain.  They run ONLY under `torch.jit` tracing after
`quantizer.enable_custom_impl()` (i.e. during `torch.onnx.export`); the inference hot path never touches them.

Reference: mct_quantizers/pytorch/quantizers/base_quantizer_autograd_function.py:22-59 and the `*F` classes
next to each quantizer (e.g. weights_symmetric_inferable_quantizer.py:159-215).  Note the reference's export
formulas use TRUE division `round(clip(x) / scale) * scale`, which differs from the ATen reciprocal-multiply
used at inference in the last ulp at rounding ties; the same formulas are kept here so exported graphs agree
with the reference's."""
from typing import Any, Dict

import numpy as np
import torch

import mct_quantizers_b200
from mct_quantizers_b200.common.constants import MCTQ_VERSION, ONNX_CUSTOM_OP_DOMAIN


class BaseQuantizerAutogradFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input_tensor, **kwargs):
        raise NotImplemented

    @staticmethod
    def symbolic(g, input_tensor, **kwargs):
        raise NotImplemented

    def backward(ctx: Any, *grad_outputs: Any) -> Any:
        raise NotImplementedError()

    @staticmethod
    def _get_metadata_attributes() -> Dict[str, Any]:
        return {f"{MCTQ_VERSION}_s": mct_quantizers_b200.__version__}


def _as_param(v, like):
    if isinstance(v, np.ndarray):
        return torch.tensor(v, dtype=torch.float32, device=like.device)
    return v


def _per_channel_view(t, x, per_channel, channel_axis):
    if per_channel and isinstance(t, torch.Tensor):
        shape = [1] * x.ndim
        shape[channel_axis] = -1
        return t.reshape(shape)
    return t


def export_symmetric(x, num_bits, threshold, signed, per_channel=False, channel_axis=None):
    """round(clip(x, lo, hi) / scale) * scale with scale = threshold / 2^(n - signed)."""
    threshold = _as_param(threshold, x)
    scale = threshold / (2 ** (num_bits - 1) if signed else 2 ** num_bits)
    lo = -threshold if signed else threshold * 0
    hi = threshold - scale
    lo, hi, scale = (_per_channel_view(v, x, per_channel, channel_axis) for v in (lo, hi, scale))
    if isinstance(lo, torch.Tensor):
        clipped = torch.where(x > hi, hi, torch.where(x < lo, lo, x))
    else:
        clipped = torch.clip(x, lo, hi)
    return torch.round(clipped / scale) * scale
