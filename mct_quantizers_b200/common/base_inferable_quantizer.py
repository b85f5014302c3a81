"""Root of the quantizer hierarchy and the @mark_quantizer registry decorator.

Reference: mct_quantizers/common/base_inferable_quantizer.py:21-91.  This package has its OWN root class so
that the reference and the replacement can be imported side by side (parity tests) without the registry's
"exactly one class" rule (get_quantizers.py) seeing two candidates."""
from enum import Enum
from typing import Any, Dict, List

from mct_quantizers_b200.common.quant_info import QuantizationMethod


class QuantizationTarget(Enum):
    Activation = "Activation"
    Weights = "Weights"


class QuantizerID(Enum):
    INFERABLE = "inferable_quantizer_id"


def mark_quantizer(quantization_target: QuantizationTarget = None,
                   quantization_method: List[QuantizationMethod] = None,
                   identifier: Any = None):
    """Class decorator: records what a quantizer class quantizes (target), which methods it implements and
    its family identifier as class attributes; the registry lookup filters on exactly these three."""

    def mark(quantizer_class_object):
        quantizer_class_object.quantization_target = quantization_target
        quantizer_class_object.quantization_method = quantization_method
        quantizer_class_object.identifier = identifier
        return quantizer_class_object

    return mark


class BaseInferableQuantizer:
    """Contract: ``quantizer(tensor) -> tensor`` plus ``initialize_quantization`` (called by the wrapper and
    the holders at construction; inferable quantizers have no variables to create)."""

    def __init__(self):
        pass

    def initialize_quantization(self, tensor_shape: Any, name: str, layer: Any) -> Dict[Any, Any]:
        return {}
