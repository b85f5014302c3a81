"""Quantization method enum (reference: mct_quantizers/common/quant_info.py:19-38; values are part of
the serialized surface, keep them)."""
from enum import Enum


class QuantizationMethod(Enum):
    POWER_OF_TWO = 0          # symmetric, uniform, threshold is a power of two
    LUT_POT_QUANTIZER = 1     # look-up table, power-of-two threshold
    SYMMETRIC = 2             # symmetric, uniform
    UNIFORM = 3               # asymmetric uniform (min/max range)
    LUT_SYM_QUANTIZER = 4     # look-up table, symmetric threshold
