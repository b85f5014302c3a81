"""Shared constructor logic of the symmetric quantizers: threshold -> scale and integer domain.
Reference: mct_quantizers/pytorch/quantizers/base_symmetric_inferable_quantizer.py:30-60."""
from typing import List

import numpy as np

from mct_quantizers_b200.common.base_inferable_quantizer import mark_quantizer, QuantizerID
from mct_quantizers_b200.common.quant_info import QuantizationMethod
from mct_quantizers_b200.pytorch.quantizers.base_pytorch_inferable_quantizer import BasePyTorchInferableQuantizer


@mark_quantizer(quantization_target=None,
                quantization_method=[QuantizationMethod.SYMMETRIC],
                identifier=QuantizerID.INFERABLE)
class BaseSymmetricInferableQuantizer(BasePyTorchInferableQuantizer):

    def __init__(self, num_bits: int, threshold: List[float], signed: bool):
        super(BaseSymmetricInferableQuantizer, self).__init__()
        assert isinstance(threshold, list), f'Threshold is expected to be a list, but is of type {type(threshold)}'

        self.signed = signed
        self.threshold_np = np.asarray(threshold)
        self.num_bits = num_bits
        # signed: 2^(n-1) steps on each side of zero; unsigned: 2^n steps above zero.  scales stay float64 here
        levels_per_side = 2 ** (num_bits - 1) if signed else 2 ** num_bits
        self.min_quantized_domain = -levels_per_side if signed else 0
        self.max_quantized_domain = levels_per_side - 1
        self.scales = self.threshold_np / levels_per_side
