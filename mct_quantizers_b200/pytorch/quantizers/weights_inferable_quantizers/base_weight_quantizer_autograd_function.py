"""Reference: .../weights_inferable_quantizers/base_weight_quantizer_autograd_function.py:19-28."""
from mct_quantizers_b200.pytorch.quantizers.base_quantizer_autograd_function import BaseQuantizerAutogradFunction


class BaseWeightQuantizerAutogradFunction(BaseQuantizerAutogradFunction):
    @staticmethod
    def is_signed():
        return True
