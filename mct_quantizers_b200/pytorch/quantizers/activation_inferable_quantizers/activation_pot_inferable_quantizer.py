"""ActivationPOTInferableQuantizer: symmetric activation quantizer with a power-of-two threshold.
Reference: .../activation_inferable_quantizers/activation_pot_inferable_quantizer.py:31-73."""
from typing import List

import torch

from mct_quantizers_b200.common.base_inferable_quantizer import mark_quantizer, QuantizationTarget, QuantizerID
from mct_quantizers_b200.common.constants import ONNX_CUSTOM_OP_DOMAIN
from mct_quantizers_b200.common.quant_info import QuantizationMethod
from mct_quantizers_b200.pytorch.quantizers.activation_inferable_quantizers.activation_symmetric_inferable_quantizer import \
    ActivationSymmetricInferableQuantizer, quantize_sym_activations_torch
from mct_quantizers_b200.pytorch.quantizers.activation_inferable_quantizers.base_activation_quantizer_autograd_function import \
    BaseActivationQuantizerAutogradFunction
from mct_quantizers_b200.pytorch.quantizers.weights_inferable_quantizers.weights_pot_inferable_quantizer import \
    is_power_of_two


@mark_quantizer(quantization_target=QuantizationTarget.Activation,
                quantization_method=[QuantizationMethod.POWER_OF_TWO],
                identifier=QuantizerID.INFERABLE)
class ActivationPOTInferableQuantizer(ActivationSymmetricInferableQuantizer):

    def __init__(self, num_bits: int, threshold: List[float], signed: bool):
        super(ActivationPOTInferableQuantizer, self).__init__(num_bits=num_bits, signed=signed, threshold=threshold)
        assert is_power_of_two(self.threshold_np), f'Expected threshold to be power of 2 but is {threshold}'

    def __call__(self, inputs):
        if self._use_custom_impl and torch.jit.is_tracing():
            return ActivationPOTF.apply(inputs, self.threshold_np, self.signed, self.num_bits)
        return super(ActivationPOTInferableQuantizer, self).__call__(inputs)


class ActivationPOTF(BaseActivationQuantizerAutogradFunction):
    @staticmethod
    def forward(ctx, input_tensor, threshold, signed, num_bits):
        return quantize_sym_activations_torch(input_tensor, threshold, signed, num_bits)

    @staticmethod
    def symbolic(g, input_tensor, threshold, signed, num_bits):
        return g.op(f"{ONNX_CUSTOM_OP_DOMAIN}::ActivationPOTQuantizer", input_tensor, threshold_f=threshold,
                    signed_i=int(signed), num_bits_i=num_bits,
                    **ActivationPOTF._get_metadata_attributes()).setType(input_tensor.type())
