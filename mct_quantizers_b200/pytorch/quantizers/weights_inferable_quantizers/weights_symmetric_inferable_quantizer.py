"""WeightsSymmetricInferableQuantizer: signed symmetric fake-quant of weights, per-channel or per-tensor.
Reference: .../weights_inferable_quantizers/weights_symmetric_inferable_quantizer.py:76-157 (class),
:32-70 and :159-215 (export formula and ONNX node)."""
from typing import List

import numpy as np
import torch

from mct_quantizers_b200 import ops
from mct_quantizers_b200.common.base_inferable_quantizer import mark_quantizer, QuantizationTarget, QuantizerID
from mct_quantizers_b200.common.constants import ONNX_CUSTOM_OP_DOMAIN
from mct_quantizers_b200.common.quant_info import QuantizationMethod
from mct_quantizers_b200.pytorch.quantizer_utils import to_torch_tensor, get_working_device
from mct_quantizers_b200.pytorch.quantizers.base_quantizer_autograd_function import export_symmetric
from mct_quantizers_b200.pytorch.quantizers.base_symmetric_inferable_quantizer import BaseSymmetricInferableQuantizer
from mct_quantizers_b200.pytorch.quantizers.weights_inferable_quantizers.base_weight_quantizer_autograd_function import \
    BaseWeightQuantizerAutogradFunction


def quantize_sym_weights_torch(input_tensor, num_bits, threshold, per_channel, channel_axis):
    """Export-time formula (true division); not the inference path."""
    return export_symmetric(input_tensor, num_bits, threshold, True, per_channel, channel_axis)


def _validated_geometry(q, shape, scales, zero_points):
    """(C, inner, qmin, qmax) after the argument checks torch.fake_quantize_per_{channel,tensor}_affine performs."""
    qmin, qmax = int(q.min_quantized_domain), int(q.max_quantized_domain)
    if qmin > qmax:
        raise RuntimeError("`quant_min` should be less than or equal to `quant_max`.")
    if q.per_channel:
        nd = len(shape)
        axis = q.channel_axis
        if not -nd <= axis < nd:
            raise RuntimeError("`axis` must be between 0 and number of dimensions of input")
        axis %= nd
        if scales.numel() != shape[axis] or zero_points.numel() != shape[axis]:
            raise RuntimeError("dimensions of scale and zero-point are not consistent with input tensor")
        C, inner = ops.contiguous_layout(tuple(shape), axis)
        return int(C), int(inner), qmin, qmax
    if scales.numel() != 1 or zero_points.numel() != 1:
        raise RuntimeError(f"a Tensor with {max(scales.numel(), zero_points.numel())} elements cannot be converted to Scalar")
    return 1, 1, qmin, qmax


def affine_weights_call(q, inputs, export_fn):
    """Shared `__call__` body of the affine weight quantizers (symmetric, POT, uniform):
    reuse cache -> ONNX tracing branch -> one fused kernel on the input's device and current stream.
    The reference delegates to ATen here and pays two device->host syncs per per-channel call
    (zero-point range check); the range is validated once at construction instead."""
    if q.enable_reuse and not q.quantizer_first_run:
        return q.resue_outputs
    if q._use_custom_impl and torch.jit.is_tracing():
        outputs = export_fn()
    else:
        inputs.requires_grad = False            # same side effect as the reference (raises on non-leaf tensors)
        scales, zero_points = q._on(inputs.device, q.scales, q.zero_points)
        if ops.direct_ok(inputs) and inputs.is_contiguous() and scales.dim() == 1 and scales.dtype == torch.float32 \
                and zero_points.dtype == torch.int32:
            # lean path: no dispatcher; the checks ATen repeats on every call are done here once per shape
            shape = inputs.shape
            key = (shape, q.channel_axis, q.per_channel, scales.numel(), zero_points.numel(), q.min_quantized_domain, q.max_quantized_domain)
            geom = q.__dict__.get('_geom')
            if geom is None or geom[0] != key:
                geom = (key, _validated_geometry(q, shape, scales, zero_points))
                q.__dict__['_geom'] = geom
            C, inner, qmin, qmax = geom[1]
            outputs = ops.affine_params_direct(inputs.detach(), scales, zero_points, C, inner, qmin, qmax)
        elif q.per_channel:
            outputs = torch.ops.mctq.fq_affine_channel(inputs, scales.flatten(), zero_points.flatten(), q.channel_axis,
                                                       q.min_quantized_domain, q.max_quantized_domain)
        else:
            outputs = torch.ops.mctq.fq_affine_tensor(inputs, scales, zero_points,
                                                      q.min_quantized_domain, q.max_quantized_domain)
    if q.enable_reuse and q.quantizer_first_run:
        q.resue_outputs = outputs
        q.quantizer_first_run = False
    return outputs


@mark_quantizer(quantization_target=QuantizationTarget.Weights,
                quantization_method=[QuantizationMethod.SYMMETRIC],
                identifier=QuantizerID.INFERABLE)
class WeightsSymmetricInferableQuantizer(BaseSymmetricInferableQuantizer):
    """Signed symmetric weight quantizer: scale_c = threshold_c / 2^(n-1), zero point 0."""

    def __init__(self, num_bits: int, threshold: List[float], per_channel: bool, channel_axis: int = None):
        super(WeightsSymmetricInferableQuantizer, self).__init__(threshold=threshold, num_bits=num_bits, signed=True)
        if per_channel:
            assert channel_axis is not None, f'Channel axis is missing in per channel quantization'
            assert len(threshold) >= 1, \
                f'In per-channel quantization threshold should be of length >= 1 but is {len(threshold)}'
        else:
            assert len(threshold) == 1, \
                f'In per-tensor quantization threshold should be of length 1 but is {len(threshold)}'
        self.per_channel = per_channel
        self.channel_axis = channel_axis
        # f64 scales -> f32 tensor; int32 zero points (all zero)
        self.scales = to_torch_tensor(self.scales).to(get_working_device())
        self.zero_points = torch.zeros(len(threshold), dtype=torch.int32).to(get_working_device())

    def __call__(self, inputs: torch.Tensor) -> torch.Tensor:
        return affine_weights_call(self, inputs, lambda: WeightsSymmetricF.apply(
            inputs, self.num_bits, self.threshold_np, self.per_channel, self.channel_axis))


class WeightsSymmetricF(BaseWeightQuantizerAutogradFunction):
    @staticmethod
    def forward(ctx, input_tensor, num_bits, threshold, per_channel, channel_axis):
        return quantize_sym_weights_torch(input_tensor, num_bits, threshold, per_channel, channel_axis)

    @staticmethod
    def symbolic(g, input_tensor, num_bits, threshold, per_channel, channel_axis):
        # a per-tensor op must still carry a channel_axis attribute for onnxruntime
        if not per_channel and channel_axis is None:
            channel_axis = 0
        return g.op(f"{ONNX_CUSTOM_OP_DOMAIN}::WeightsSymmetricQuantizer", input_tensor,
                    g.op('Constant', value_t=torch.tensor(threshold, dtype=torch.float32)),
                    num_bits_i=num_bits, per_channel_i=int(per_channel), channel_axis_i=channel_axis,
                    signed_i=int(WeightsSymmetricF.is_signed()),
                    **WeightsSymmetricF._get_metadata_attributes()).setType(input_tensor.type())
