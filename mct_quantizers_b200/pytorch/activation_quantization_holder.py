"""nn.Module that applies one activation quantizer to its input.
Reference: mct_quantizers/pytorch/activation_quantization_holder.py:23-63."""
import torch

from mct_quantizers_b200.common.base_inferable_quantizer import BaseInferableQuantizer
from mct_quantizers_b200.common.constants import ACTIVATION_HOLDER_QUANTIZER


class PytorchActivationQuantizationHolder(torch.nn.Module):
    def __init__(self, activation_holder_quantizer: BaseInferableQuantizer, **kwargs):
        super(PytorchActivationQuantizationHolder, self).__init__(**kwargs)
        self.activation_holder_quantizer = activation_holder_quantizer
        self.activation_holder_quantizer.initialize_quantization(None, ACTIVATION_HOLDER_QUANTIZER + "_out", self)

    def forward(self, inputs):
        """One fused fake-quant kernel on the input's device and current stream."""
        return self.activation_holder_quantizer(inputs)

    def convert_to_inferable_quantizers(self):
        """Swap a trainable quantizer (anything exposing convert2inferable) for its inferable twin."""
        convert = getattr(self.activation_holder_quantizer, 'convert2inferable', None)
        if callable(convert):  # pragma: no cover
            self.activation_holder_quantizer = convert()
