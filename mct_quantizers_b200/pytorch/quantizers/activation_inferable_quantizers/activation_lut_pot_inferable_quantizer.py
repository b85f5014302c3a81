"""ActivationLutPOTInferableQuantizer: per-tensor look-up-table fake-quant of activations, POT threshold.
Reference: .../activation_inferable_quantizers/activation_lut_pot_inferable_quantizer.py:33-91."""
from typing import List

import torch

from mct_quantizers_b200 import ops  # noqa: F401
from mct_quantizers_b200.common.base_inferable_quantizer import mark_quantizer, QuantizationTarget, QuantizerID
from mct_quantizers_b200.common.constants import LUT_VALUES_BITWIDTH, EPS
from mct_quantizers_b200.common.quant_info import QuantizationMethod
from mct_quantizers_b200.pytorch.quantizer_utils import to_torch_tensor, get_working_device, lut_quantizer, lut_quantizer_export, \
    lut_search_table
from mct_quantizers_b200.pytorch.quantizers.base_lut_symmetric_inferable_quantizer import \
    BaseLUTSymmetricInferableQuantizer
from mct_quantizers_b200.pytorch.quantizers.weights_inferable_quantizers.weights_pot_inferable_quantizer import \
    is_power_of_two


@mark_quantizer(quantization_target=QuantizationTarget.Activation,
                quantization_method=[QuantizationMethod.LUT_POT_QUANTIZER],
                identifier=QuantizerID.INFERABLE)
class ActivationLutPOTInferableQuantizer(BaseLUTSymmetricInferableQuantizer):

    def __init__(self, num_bits: int, lut_values: List[float], threshold: List[float], signed: bool,
                 lut_values_bitwidth: int = LUT_VALUES_BITWIDTH, eps: float = EPS):
        super(ActivationLutPOTInferableQuantizer, self).__init__(num_bits=num_bits, lut_values=lut_values,
                                                                 threshold=threshold, signed=signed,
                                                                 lut_values_bitwidth=lut_values_bitwidth, eps=eps)
        assert is_power_of_two(self._threshold_np), f'Expected threshold to be power of 2 but is {threshold}'
        assert len(self.threshold) == 1, \
            f'For activation, quantization per channel is not supported and threshold should be of length 1 but is {len(threshold)}'
        self.threshold = self.threshold[0]                       # Python float
        self.lut_values = to_torch_tensor(self._lut_values_np).to(get_working_device())   # f32 tensor (public attr)
        self._search_table = None

    def __call__(self, inputs: torch.Tensor):
        if self._use_custom_impl and torch.jit.is_tracing():
            # export: the reference has no autograd Function for this quantizer -- it traces into the elementary ops of
            # lut_quantizer (activation_lut_pot_inferable_quantizer.py:76-91), and so does this branch
            return lut_quantizer_export(inputs, lut_values=self.lut_values.to(inputs.device), signed=self.signed, threshold=self.threshold,
                                        lut_values_bitwidth=self.lut_values_bitwidth, eps=self.eps)
        if self.__dict__.get('_search_table') is None:     # also absent in objects unpickled from the reference
            self._search_table = lut_search_table(self._lut_values_np, self.lut_values_bitwidth, self.signed)
        # the table stays on the host: the operator keeps per-device copies and the prepared per-channel decision
        # tables (nothing here touches inputs.device, so the call is fx-traceable)
        return lut_quantizer(inputs.detach(), lut_values=self.lut_values, signed=self.signed, threshold=self.threshold,
                             lut_values_bitwidth=self.lut_values_bitwidth, eps=self.eps, _table=self._search_table)
