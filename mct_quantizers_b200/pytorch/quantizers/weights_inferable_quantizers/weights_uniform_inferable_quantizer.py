"""WeightsUniformInferableQuantizer: asymmetric (min/max) fake-quant of weights.
Reference: .../weights_inferable_quantizers/weights_uniform_inferable_quantizer.py:83-171."""
from typing import List

import numpy as np
import torch

from mct_quantizers_b200.common.base_inferable_quantizer import mark_quantizer, QuantizationTarget, QuantizerID
from mct_quantizers_b200.common.constants import ONNX_CUSTOM_OP_DOMAIN
from mct_quantizers_b200.common.quant_info import QuantizationMethod
from mct_quantizers_b200.pytorch.quantizer_utils import fix_range_to_include_zero, get_working_device, div_by_python_number
from mct_quantizers_b200.pytorch.quantizers.base_quantizer_autograd_function import _per_channel_view
from mct_quantizers_b200.pytorch.quantizers.base_uniform_inferable_quantizer import BaseUniformInferableQuantizer
from mct_quantizers_b200.pytorch.quantizers.weights_inferable_quantizers.base_weight_quantizer_autograd_function import \
    BaseWeightQuantizerAutogradFunction
from mct_quantizers_b200.pytorch.quantizers.weights_inferable_quantizers.weights_symmetric_inferable_quantizer import \
    affine_weights_call


def quantize_uniform_weights_torch(input_tensor, num_bits, min_range, max_range, per_channel, channel_axis=None):
    """Export-time formula (true division, no zero point); not the inference path."""
    lo = torch.tensor(min_range, dtype=torch.float32, device=input_tensor.device) if isinstance(min_range, np.ndarray) else min_range
    hi = torch.tensor(max_range, dtype=torch.float32, device=input_tensor.device) if isinstance(max_range, np.ndarray) else max_range
    lo, hi = fix_range_to_include_zero(lo, hi, num_bits)
    step = (hi - lo) / (2 ** num_bits - 1)
    lo, hi, step = (_per_channel_view(v, input_tensor, per_channel, channel_axis) for v in (lo, hi, step))
    clipped = torch.where(input_tensor > hi, hi, torch.where(input_tensor < lo, lo, input_tensor))
    return torch.round(clipped / step) * step


@mark_quantizer(quantization_target=QuantizationTarget.Weights,
                quantization_method=[QuantizationMethod.UNIFORM],
                identifier=QuantizerID.INFERABLE)
class WeightsUniformInferableQuantizer(BaseUniformInferableQuantizer):

    def __init__(self, num_bits: int, min_range: List[float], max_range: List[float], per_channel: bool,
                 channel_axis: int = None):
        super(WeightsUniformInferableQuantizer, self).__init__(num_bits=num_bits, min_range=min_range, max_range=max_range)
        if per_channel:
            assert channel_axis is not None, f'Channel axis is missing in per channel quantization'
            assert len(min_range) >= 1, f'In per-channel quantization min_range should be of length >= 1 but is {len(min_range)}'
            assert len(max_range) >= 1, f'In per-channel quantization max_range should be of length >= 1 but is {len(max_range)}'
        else:
            assert len(min_range) == 1, f'In per-tensor quantization min_range should be of length 1 but is {len(min_range)}'
            assert len(max_range) == 1, f'In per-tensor quantization max_range should be of length 1 but is {len(max_range)}'
        self.per_channel = per_channel
        self.channel_axis = channel_axis

        self.adjusted_min_range_np = self.min_range.cpu().numpy()
        self.adjusted_max_range_np = self.max_range.cpu().numpy()

        # step and zero point in f32 on the host; NB the zero point is TRUNCATED toward zero (.int()), not
        # rounded -- that is what the reference does and what the golden vectors pin
        lo, hi = self.min_range.cpu(), self.max_range.cpu()
        scales = div_by_python_number(hi - lo, 2 ** num_bits - 1)     # device flavour of the reference's `/`: quantizer_utils.py
        zero_points = -(lo / scales).int()
        zp_lo, zp_hi = int(zero_points.min()), int(zero_points.max())
        if zp_lo < self.min_quantized_domain or zp_hi > self.max_quantized_domain:
            # ATen raises this at call time (after a device->host sync); the parameters are known here
            self._zero_point_out_of_range = True
        self.scales = scales.to(get_working_device())
        self.zero_points = zero_points.to(get_working_device())

    def __call__(self, inputs: torch.Tensor) -> torch.Tensor:
        if getattr(self, '_zero_point_out_of_range', False) and self.per_channel:
            raise RuntimeError("`zero_point` must be between `quant_min` and `quant_max`.")
        return affine_weights_call(self, inputs, lambda: WeightsUniformF.apply(
            inputs, self.num_bits, self.adjusted_min_range_np, self.adjusted_max_range_np, self.per_channel,
            self.channel_axis))


class WeightsUniformF(BaseWeightQuantizerAutogradFunction):
    @staticmethod
    def forward(ctx, input_tensor, num_bits, min_range, max_range, per_channel, channel_axis):
        return quantize_uniform_weights_torch(input_tensor, num_bits, min_range, max_range, per_channel, channel_axis)

    @staticmethod
    def symbolic(g, input_tensor, num_bits, min_range, max_range, per_channel, channel_axis):
        if not per_channel and channel_axis is None:
            channel_axis = 0
        return g.op(f"{ONNX_CUSTOM_OP_DOMAIN}::WeightsUniformQuantizer", input_tensor,
                    g.op('Constant', value_t=torch.tensor(min_range, dtype=torch.float32)),
                    g.op('Constant', value_t=torch.tensor(max_range, dtype=torch.float32)),
                    num_bits_i=num_bits, per_channel_i=int(per_channel), channel_axis_i=channel_axis,
                    signed_i=WeightsUniformF.is_signed(),
                    **WeightsUniformF._get_metadata_attributes()).setType(input_tensor.type())
