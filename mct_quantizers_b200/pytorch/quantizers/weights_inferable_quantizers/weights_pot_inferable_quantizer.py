"""WeightsPOTInferableQuantizer: symmetric weight quantizer whose thresholds must be powers of two.
Reference: .../weights_inferable_quantizers/weights_pot_inferable_quantizer.py:32-96."""
from typing import List

import numpy as np
import torch

from mct_quantizers_b200.common.base_inferable_quantizer import mark_quantizer, QuantizationTarget, QuantizerID
from mct_quantizers_b200.common.constants import ONNX_CUSTOM_OP_DOMAIN
from mct_quantizers_b200.common.quant_info import QuantizationMethod
from mct_quantizers_b200.pytorch.quantizers.weights_inferable_quantizers.base_weight_quantizer_autograd_function import \
    BaseWeightQuantizerAutogradFunction
from mct_quantizers_b200.pytorch.quantizers.weights_inferable_quantizers.weights_symmetric_inferable_quantizer import \
    WeightsSymmetricInferableQuantizer, quantize_sym_weights_torch, affine_weights_call


def is_power_of_two(values: np.ndarray) -> bool:
    exponents = np.log2(values.flatten())
    return bool(np.all(np.round(exponents) == exponents))


@mark_quantizer(quantization_target=QuantizationTarget.Weights,
                quantization_method=[QuantizationMethod.POWER_OF_TWO],
                identifier=QuantizerID.INFERABLE)
class WeightsPOTInferableQuantizer(WeightsSymmetricInferableQuantizer):

    def __init__(self, num_bits: int, threshold: List[float], per_channel: bool, channel_axis: int = None):
        super(WeightsPOTInferableQuantizer, self).__init__(num_bits=num_bits, threshold=threshold,
                                                           per_channel=per_channel, channel_axis=channel_axis)
        self.num_bits = num_bits
        self.threshold = threshold
        self.per_channel = per_channel
        self.channel_axis = channel_axis
        assert is_power_of_two(self.threshold_np), f'Expected threshold to be power of 2 but is {threshold}'

    def __call__(self, inputs: torch.Tensor) -> torch.Tensor:
        return affine_weights_call(self, inputs, lambda: WeightsPOTF.apply(
            inputs, self.num_bits, self.threshold_np, self.per_channel, self.channel_axis))


class WeightsPOTF(BaseWeightQuantizerAutogradFunction):
    @staticmethod
    def forward(ctx, input_tensor, num_bits, threshold, per_channel, channel_axis):
        return quantize_sym_weights_torch(input_tensor, num_bits, threshold, per_channel, channel_axis)

    @staticmethod
    def symbolic(g, input_tensor, num_bits, threshold, per_channel, channel_axis):
        if not per_channel and channel_axis is None:
            channel_axis = 0
        return g.op(f"{ONNX_CUSTOM_OP_DOMAIN}::WeightsPOTQuantizer", input_tensor,
                    g.op('Constant', value_t=torch.tensor(threshold, dtype=torch.float32)),
                    num_bits_i=num_bits, per_channel_i=int(per_channel), channel_axis_i=channel_axis,
                    signed_i=int(WeightsPOTF.is_signed()),
                    **WeightsPOTF._get_metadata_attributes()).setType(input_tensor.type())
