"""ActivationSymmetricInferableQuantizer: per-tensor symmetric fake-quant of activations (signed or unsigned).
Reference: .../activation_inferable_quantizers/activation_symmetric_inferable_quantizer.py:59-117."""
from typing import List

import torch

from mct_quantizers_b200 import ops
from mct_quantizers_b200.common.base_inferable_quantizer import mark_quantizer, QuantizationTarget, QuantizerID
from mct_quantizers_b200.common.constants import ONNX_CUSTOM_OP_DOMAIN
from mct_quantizers_b200.common.quant_info import QuantizationMethod
from mct_quantizers_b200.pytorch.quantizers.activation_inferable_quantizers.base_activation_quantizer_autograd_function import \
    BaseActivationQuantizerAutogradFunction
from mct_quantizers_b200.pytorch.quantizers.base_quantizer_autograd_function import export_symmetric
from mct_quantizers_b200.pytorch.quantizers.base_symmetric_inferable_quantizer import BaseSymmetricInferableQuantizer


def quantize_sym_activations_torch(input_tensor, threshold, signed, num_bits):
    """Export-time formula (true division); not the inference path."""
    return export_symmetric(input_tensor, num_bits, threshold, signed)


@mark_quantizer(quantization_target=QuantizationTarget.Activation,
                quantization_method=[QuantizationMethod.SYMMETRIC],
                identifier=QuantizerID.INFERABLE)
class ActivationSymmetricInferableQuantizer(BaseSymmetricInferableQuantizer):

    def __init__(self, num_bits: int, threshold: List[float], signed: bool):
        super(ActivationSymmetricInferableQuantizer, self).__init__(num_bits=num_bits, threshold=threshold, signed=signed)
        assert len(threshold) == 1, \
            f'For activation, only per-tensor quantization is supported. Thus, threshold should be of length 1 but is {len(threshold)}'
        assert self.threshold_np.shape[0] == 1
        self.threshold_np = self.threshold_np[0]          # numpy scalar
        assert len(self.scales) == 1, \
            f'For activation, quantization per channel is not supported and threshold should be of length 1 but is {len(threshold)}'
        self.scales = float(self.scales[0])               # Python float (f64); narrowed to f32 at launch like ATen
        self.zero_points = 0

    def __call__(self, inputs: torch.Tensor):
        if self._use_custom_impl and torch.jit.is_tracing():
            return ActivationSymF.apply(inputs, self.threshold_np, self.signed, self.num_bits)
        # scalar parameters travel by value: one launch on the current stream, no host sync, no autograd graph
        if ops.direct_ok(inputs):
            cached = self.__dict__.get('_launch_args')
            if cached is None or cached[0] != (self.scales, self.zero_points, self.min_quantized_domain, self.max_quantized_domain):
                cached = self._validated_launch_args(self.scales, self.zero_points)
            return ops.affine_scalar_direct(inputs, *cached[1])
        return torch.ops.mctq.fq_affine_scalar(inputs.detach(), self.scales, self.zero_points,
                                               self.min_quantized_domain, self.max_quantized_domain)


class ActivationSymF(BaseActivationQuantizerAutogradFunction):
    @staticmethod
    def forward(ctx, input_tensor, threshold, signed, num_bits):
        return quantize_sym_activations_torch(input_tensor, threshold, signed, num_bits)

    @staticmethod
    def symbolic(g, input_tensor, threshold, signed, num_bits):
        return g.op(f"{ONNX_CUSTOM_OP_DOMAIN}::ActivationSymmetricQuantizer", input_tensor, threshold_f=threshold,
                    signed_i=int(signed), num_bits_i=num_bits,
                    **ActivationSymF._get_metadata_attributes()).setType(input_tensor.type())
