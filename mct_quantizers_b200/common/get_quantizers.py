"""Registry lookup (reference: mct_quantizers/common/get_quantizers.py:22-53): the single inferable
quantizer class under `quantizer_base_class` for a (target, method) pair, else Logger.error raises."""
from mct_quantizers_b200.common.base_inferable_quantizer import QuantizationTarget, QuantizerID
from mct_quantizers_b200.common.constants import QUANTIZATION_TARGET, QUANTIZATION_METHOD, QUANTIZER_ID
from mct_quantizers_b200.common.get_all_subclasses import get_all_subclasses
from mct_quantizers_b200.common.quant_info import QuantizationMethod
from mct_quantizers_b200.logger import Logger


def get_inferable_quantizer_class(quant_target: QuantizationTarget,
                                  quant_method: QuantizationMethod,
                                  quantizer_base_class: type) -> type:
    def matches(q_class):
        methods = getattr(q_class, QUANTIZATION_METHOD)
        return (getattr(q_class, QUANTIZATION_TARGET) == quant_target and methods is not None
                and quant_method in methods and getattr(q_class, QUANTIZER_ID) is QuantizerID.INFERABLE)

    filtered_quantizers = [c for c in get_all_subclasses(quantizer_base_class) if matches(c)]
    if len(filtered_quantizers) != 1:
        Logger.error(f"Found {len(filtered_quantizers)} quantizer for target {quant_target.value} "
                     f"that matches the requested quantization method {quant_method.name} "
                     f"but there should be exactly one."
                     f"The possible quantizers that were found are {filtered_quantizers}.")
    return filtered_quantizers[0]
