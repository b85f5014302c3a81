"""Shared constructor logic of the uniform (min/max) quantizers: validation + range fixing.
Reference: mct_quantizers/pytorch/quantizers/base_uniform_inferable_quantizer.py:31-66."""
from typing import List

import numpy as np
import torch

from mct_quantizers_b200.common.base_inferable_quantizer import mark_quantizer, QuantizerID
from mct_quantizers_b200.common.quant_info import QuantizationMethod
from mct_quantizers_b200.pytorch.quantizer_utils import get_working_device, fix_range_to_include_zero
from mct_quantizers_b200.pytorch.quantizers.base_pytorch_inferable_quantizer import BasePyTorchInferableQuantizer


@mark_quantizer(quantization_target=None,
                quantization_method=[QuantizationMethod.UNIFORM],
                identifier=QuantizerID.INFERABLE)
class BaseUniformInferableQuantizer(BasePyTorchInferableQuantizer):

    def __init__(self, num_bits: int, min_range: List[float], max_range: List[float]):
        super(BaseUniformInferableQuantizer, self).__init__()
        assert isinstance(min_range, list), f'min_range is expected to be a list, but is of type {type(min_range)}'
        assert isinstance(max_range, list), f'max_range is expected to be a list, but is of type {type(max_range)}'
        for _min, _max in zip(min_range, max_range):
            assert _min < _max, f"Max range must be greater than min value but min is {_min} and max is {_max}"

        # f32 range fixing on the host (bit-identical on any device: IEEE f32 elementwise ops), then to the
        # working device like the reference
        lo = torch.from_numpy(np.asarray(min_range).astype(np.float32))
        hi = torch.from_numpy(np.asarray(max_range).astype(np.float32))
        lo, hi = fix_range_to_include_zero(lo, hi, num_bits)
        self.min_range = lo.to(get_working_device())
        self.max_range = hi.to(get_working_device())

        self.num_bits = num_bits
        self.min_quantized_domain = 0
        self.max_quantized_domain = 2 ** num_bits - 1
