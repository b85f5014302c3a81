"""Holder variant for quantization-preserving ops (reshape, max-pool ...): quantization can be bypassed.
Reference: mct_quantizers/pytorch/preserving_activation_quantization_holder.py:24-56."""
from mct_quantizers_b200.common.base_inferable_quantizer import BaseInferableQuantizer
from mct_quantizers_b200.pytorch.activation_quantization_holder import PytorchActivationQuantizationHolder


class PytorchPreservingActivationQuantizationHolder(PytorchActivationQuantizationHolder):
    def __init__(self, activation_holder_quantizer: BaseInferableQuantizer, quantization_bypass: bool = False, **kwargs):
        super(PytorchPreservingActivationQuantizationHolder, self).__init__(
            activation_holder_quantizer=activation_holder_quantizer, **kwargs)
        self.quantization_bypass = quantization_bypass

    def forward(self, inputs):
        if self.quantization_bypass:
            return inputs
        return super().forward(inputs)
