"""ActivationUniformInferableQuantizer: per-tensor asymmetric (min/max) fake-quant of activations.
Reference: .../activation_inferable_quantizers/activation_uniform_inferable_quantizer.py:71-128."""
from typing import List

import numpy as np
import torch

from mct_quantizers_b200 import ops
from mct_quantizers_b200.common.base_inferable_quantizer import mark_quantizer, QuantizationTarget, QuantizerID
from mct_quantizers_b200.common.constants import ONNX_CUSTOM_OP_DOMAIN
from mct_quantizers_b200.common.quant_info import QuantizationMethod
from mct_quantizers_b200.pytorch.quantizers.activation_inferable_quantizers.base_activation_quantizer_autograd_function import \
    BaseActivationQuantizerAutogradFunction
from mct_quantizers_b200.pytorch.quantizers.base_uniform_inferable_quantizer import BaseUniformInferableQuantizer


def _range_including_zero_f64(range_min: float, range_max: float, n_bits: int):
    """Scalar twin of the range fixing used by the export formula (float64, with the final clamp that the
    reference's numpy helper common/quant_utils.py:20-50 applies)."""
    if range_min > 0:
        return 0.0, range_max
    if range_max < 0:
        return range_min, 0.0
    step = (range_max - range_min) / (2 ** n_bits - 1)
    lo = step * np.round(range_min / step)
    hi = range_max - range_min + lo
    return min(lo, 0.0), max(hi, 0.0)


def quantize_uniform_activations_torch(tensor_data, range_min, range_max, n_bits):
    """Export-time formula; not the inference path."""
    a, b = _range_including_zero_f64(range_min, range_max, n_bits)
    step = (b - a) / (2 ** n_bits - 1)
    return step * torch.round((torch.clip(tensor_data, min=a, max=b) - a) / step) + a


@mark_quantizer(quantization_target=QuantizationTarget.Activation,
                quantization_method=[QuantizationMethod.UNIFORM],
                identifier=QuantizerID.INFERABLE)
class ActivationUniformInferableQuantizer(BaseUniformInferableQuantizer):

    def __init__(self, num_bits: int, min_range: List[float], max_range: List[float]):
        super(ActivationUniformInferableQuantizer, self).__init__(num_bits=num_bits, min_range=min_range, max_range=max_range)
        assert isinstance(min_range, list), f'min_range is expected to be a list, but is of type {type(min_range)}'
        assert isinstance(max_range, list), f'max_range is expected to be a list, but is of type {type(max_range)}'
        assert len(min_range) == 1, \
            f'For activation, only per-tensor quantization is supported. Thus, min_range should be of length 1 but is {len(min_range)}'
        assert len(max_range) == 1, \
            f'For activation, only per-tensor quantization is supported. Thus, max_range should be of length 1 but is {len(max_range)}'
        # range-fixed f32 values widened to Python floats; scale in float64; zero point ROUNDED (np.round),
        # unlike the weights flavour which truncates
        self.min_range = self.min_range[0].cpu().item()
        self.max_range = self.max_range[0].cpu().item()
        self.scale = float((self.max_range - self.min_range) / ((2 ** num_bits) - 1))
        self.zero_point = int(-np.round(self.min_range / self.scale))

    def __call__(self, inputs: torch.Tensor):
        if self._use_custom_impl and torch.jit.is_tracing():
            return ActivationUniformF.apply(inputs, self.min_range, self.max_range, self.num_bits)
        if ops.direct_ok(inputs):
            cached = self.__dict__.get('_launch_args')
            if cached is None or cached[0] != (self.scale, self.zero_point, self.min_quantized_domain, self.max_quantized_domain):
                cached = self._validated_launch_args(self.scale, self.zero_point)
            return ops.affine_scalar_direct(inputs, *cached[1])
        return torch.ops.mctq.fq_affine_scalar(inputs.detach(), self.scale, self.zero_point,
                                               self.min_quantized_domain, self.max_quantized_domain)


class ActivationUniformF(BaseActivationQuantizerAutogradFunction):
    @staticmethod
    def forward(ctx, input_tensor, min_range, max_range, num_bits):
        return quantize_uniform_activations_torch(input_tensor, min_range, max_range, num_bits)

    @staticmethod
    def symbolic(g, input_tensor, min_range, max_range, num_bits):
        return g.op(f"{ONNX_CUSTOM_OP_DOMAIN}::ActivationUniformQuantizer", input_tensor, min_range_f=min_range,
                    max_range_f=max_range, num_bits_i=num_bits,
                    **ActivationUniformF._get_metadata_attributes()).setType(input_tensor.type())
