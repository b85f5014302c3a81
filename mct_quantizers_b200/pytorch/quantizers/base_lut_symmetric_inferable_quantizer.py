"""Shared constructor logic of the LUT quantizers: validation of the centroid list.
Reference: mct_quantizers/pytorch/quantizers/base_lut_symmetric_inferable_quantizer.py:30-94 (the assertion
messages are part of the contract; the reference's tests compare them verbatim)."""
import warnings
from typing import List

import numpy as np

from mct_quantizers_b200.common.base_inferable_quantizer import mark_quantizer, QuantizerID
from mct_quantizers_b200.common.quant_info import QuantizationMethod
from mct_quantizers_b200.pytorch.quantizers.base_pytorch_inferable_quantizer import BasePyTorchInferableQuantizer


@mark_quantizer(quantization_target=None,
                quantization_method=[QuantizationMethod.LUT_SYM_QUANTIZER],
                identifier=QuantizerID.INFERABLE)
class BaseLUTSymmetricInferableQuantizer(BasePyTorchInferableQuantizer):

    def __init__(self, num_bits: int, lut_values: List[float], threshold: List[float], signed: bool,
                 lut_values_bitwidth: int, eps: float):
        super(BaseLUTSymmetricInferableQuantizer, self).__init__()
        assert isinstance(threshold, list), f'Threshold is expected to be a list, but is of type {type(threshold)}'
        assert isinstance(lut_values, list), f'lut_values is expected to be a list, but is of type {type(lut_values)}'

        self._threshold_np = np.asarray(threshold)
        self._lut_values_np = np.asarray(lut_values)
        lut = self._lut_values_np

        assert len(np.unique(lut)) <= 2 ** num_bits, \
            f'Expected num of lut values to be less or equal than {2 ** num_bits} but got {len(lut)}'
        assert not np.any(lut - lut.astype(int)), f'Expected lut values to be integers'
        if signed:
            half = 2 ** (lut_values_bitwidth - 1)
            assert np.all((-half <= lut) & (lut <= half - 1)), f'Expected lut values in the quantization range'
        else:
            assert np.all(lut <= 2 ** lut_values_bitwidth), f'Expected lut values in the quantization range'
            assert np.all(lut >= 0), f'Expected unsigned lut values in unsigned activation quantization'
        assert num_bits <= lut_values_bitwidth, \
            f'Look-Up-Table bit configuration has {num_bits} bits. It must be less then {lut_values_bitwidth}'
        if num_bits == lut_values_bitwidth:
            warnings.warn("Num of bits equal to multiplier n bits, Please be aware LUT quantizier may be "
                          "inefficient in that case, consider using SymmetricInferableQuantizer instead")

        self.threshold = threshold
        self.lut_values = lut_values
        self.signed = signed
        self.num_bits = num_bits
        self.lut_values_bitwidth = lut_values_bitwidth
        self.eps = eps
