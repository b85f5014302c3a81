"""Reference: .../activation_inferable_quantizers/base_activation_quantizer_autograd_function.py:18-25."""
from mct_quantizers_b200.pytorch.quantizers.base_quantizer_autograd_function import BaseQuantizerAutogradFunction


class BaseActivationQuantizerAutogradFunction(BaseQuantizerAutogradFunction):
    pass
