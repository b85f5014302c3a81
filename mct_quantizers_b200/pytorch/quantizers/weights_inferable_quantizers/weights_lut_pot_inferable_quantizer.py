"""WeightsLUTPOTInferableQuantizer: LUT weight quantizer whose thresholds must be powers of two.
Reference: .../weights_inferable_quantizers/weights_lut_pot_inferable_quantizer.py:35-104."""
from typing import List

import torch

from mct_quantizers_b200.common.base_inferable_quantizer import mark_quantizer, QuantizationTarget, QuantizerID
from mct_quantizers_b200.common.constants import LUT_VALUES_BITWIDTH, EPS
from mct_quantizers_b200.common.quant_info import QuantizationMethod
from mct_quantizers_b200.pytorch.quantizers.weights_inferable_quantizers.base_weight_quantizer_autograd_function import \
    BaseWeightQuantizerAutogradFunction
from mct_quantizers_b200.pytorch.quantizers.weights_inferable_quantizers.weights_lut_symmetric_inferable_quantizer import \
    WeightsLUTSymmetricInferableQuantizer, lut_weights_call, _lut_export_forward, _lut_export_symbolic
from mct_quantizers_b200.pytorch.quantizers.weights_inferable_quantizers.weights_pot_inferable_quantizer import \
    is_power_of_two


@mark_quantizer(quantization_target=QuantizationTarget.Weights,
                quantization_method=[QuantizationMethod.LUT_POT_QUANTIZER],
                identifier=QuantizerID.INFERABLE)
class WeightsLUTPOTInferableQuantizer(WeightsLUTSymmetricInferableQuantizer):

    def __init__(self, num_bits: int, lut_values: List[float], threshold: List[float], per_channel: bool,
                 channel_axis: int = None, input_rank: int = None, lut_values_bitwidth: int = LUT_VALUES_BITWIDTH,
                 eps: float = EPS):
        super(WeightsLUTPOTInferableQuantizer, self).__init__(num_bits=num_bits, threshold=threshold,
                                                              lut_values=lut_values, per_channel=per_channel,
                                                              channel_axis=channel_axis, input_rank=input_rank,
                                                              lut_values_bitwidth=lut_values_bitwidth, eps=eps)
        assert is_power_of_two(self._threshold_np), f'Expected threshold to be power of 2 but is {threshold}'

    def __call__(self, inputs: torch.Tensor) -> torch.Tensor:
        return lut_weights_call(self, inputs, lambda: WeightsLUTPOTF.apply(
            inputs, self.num_bits, self._lut_values_np, self._threshold_np, self.lut_values_bitwidth, self.eps,
            self.per_channel, self.channel_axis, self.input_rank))


class WeightsLUTPOTF(BaseWeightQuantizerAutogradFunction):
    @staticmethod
    def forward(ctx, input_tensor, num_bits, lut_values, threshold, lut_values_bitwidth, eps, per_channel,
                channel_axis, input_rank):
        return _lut_export_forward(input_tensor, lut_values, threshold, lut_values_bitwidth, eps, per_channel,
                                   channel_axis, input_rank)

    @staticmethod
    def symbolic(g, input_tensor, num_bits, lut_values, threshold, lut_values_bitwidth, eps, per_channel,
                 channel_axis, input_rank):
        return _lut_export_symbolic("WeightsLUTPOTQuantizer", WeightsLUTPOTF, g, input_tensor, num_bits, lut_values,
                                    threshold, lut_values_bitwidth, eps, per_channel, channel_axis, input_rank)
