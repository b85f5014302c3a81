"""WeightsLUTSymmetricInferableQuantizer: nearest-centroid (look-up table) fake-quant of weights.
Reference: .../weights_inferable_quantizers/weights_lut_symmetric_inferable_quantizer.py:37-128."""
from typing import List

import numpy as np
import torch

from mct_quantizers_b200 import ops  # noqa: F401
from mct_quantizers_b200.common.base_inferable_quantizer import mark_quantizer, QuantizationTarget, QuantizerID
from mct_quantizers_b200.common.constants import LUT_VALUES_BITWIDTH, EPS, ONNX_CUSTOM_OP_DOMAIN
from mct_quantizers_b200.common.quant_info import QuantizationMethod
from mct_quantizers_b200.pytorch.quantizer_utils import to_torch_tensor, get_working_device, lut_quantizer, lut_quantizer_export, \
    lut_search_table
from mct_quantizers_b200.pytorch.quantizers.base_lut_symmetric_inferable_quantizer import \
    BaseLUTSymmetricInferableQuantizer
from mct_quantizers_b200.pytorch.quantizers.weights_inferable_quantizers.base_weight_quantizer_autograd_function import \
    BaseWeightQuantizerAutogradFunction


def lut_weights_call(q, inputs, export_fn):
    """Shared `__call__` body of the LUT weight quantizers: reuse cache -> ONNX tracing branch -> fused kernel."""
    if q.enable_reuse and not q.quantizer_first_run:
        return q.resue_outputs
    if q._use_custom_impl and torch.jit.is_tracing():
        outputs = export_fn()
    else:
        inputs.requires_grad = False
        if q.__dict__.get('_search_table') is None:   # built lazily (native library loads on first use; absent in objects unpickled from the reference)
            q._search_table = lut_search_table(q._lut_values_np, q.lut_values_bitwidth, True)
        thr, = q._on(inputs.device, q._threshold_torch)
        if ops.direct_ok(inputs) and (not q.per_channel or q.input_rank is None or q.input_rank == inputs.dim()):
            # lean path: launch constants cached per (shape, dtype, device); same kernel as the general path below
            cache = q.__dict__.get('_direct_cache')
            if cache is None:
                cache = q.__dict__['_direct_cache'] = {}
            outputs = ops.lut_weights_direct(inputs.detach(), q._search_table, int(q._lut_values_torch.numel()), thr.reshape(-1),
                                             bool(q.per_channel), int(q.channel_axis) if q.per_channel else 0, float(q.eps), cache)
        else:
            outputs = lut_quantizer(inputs, lut_values=q._lut_values_torch, signed=True, threshold=thr,
                                    lut_values_bitwidth=q.lut_values_bitwidth, eps=q.eps, per_channel=q.per_channel,
                                    channel_axis=q.channel_axis, input_rank=q.input_rank, _table=q._search_table)
    if q.enable_reuse and q.quantizer_first_run:
        q.resue_outputs = outputs
        q.quantizer_first_run = False
    return outputs


@mark_quantizer(quantization_target=QuantizationTarget.Weights,
                quantization_method=[QuantizationMethod.LUT_SYM_QUANTIZER],
                identifier=QuantizerID.INFERABLE)
class WeightsLUTSymmetricInferableQuantizer(BaseLUTSymmetricInferableQuantizer):

    def __init__(self, num_bits: int, lut_values: List[float], threshold: List[float], per_channel: bool,
                 channel_axis: int = None, input_rank: int = None, lut_values_bitwidth: int = LUT_VALUES_BITWIDTH,
                 eps: float = EPS):
        super(WeightsLUTSymmetricInferableQuantizer, self).__init__(threshold=threshold, num_bits=num_bits,
                                                                    lut_values=lut_values, signed=True,
                                                                    lut_values_bitwidth=lut_values_bitwidth, eps=eps)
        self.per_channel = per_channel
        self.channel_axis = channel_axis
        self.input_rank = input_rank
        if per_channel:
            assert channel_axis is not None, f'Channel axis is missing in per channel quantization'
            assert input_rank is not None, f'input_rank is missing in per channel quantization'
            assert len(threshold) >= 1, \
                f'In per-channel quantization threshold should be of length >= 1 but is {len(threshold)}'
        else:
            assert len(threshold) == 1, \
                f'In per-tensor quantization threshold should be of length 1 but is {len(threshold)}'
        self._threshold_torch = to_torch_tensor(self._threshold_np).to(get_working_device())
        self._lut_values_torch = to_torch_tensor(self._lut_values_np).to(get_working_device())
        self._search_table = None

    def __call__(self, inputs: torch.Tensor) -> torch.Tensor:
        return lut_weights_call(self, inputs, lambda: WeightsLUTSymmetricF.apply(
            inputs, self.num_bits, self._lut_values_np, self._threshold_np, self.lut_values_bitwidth, self.eps,
            self.per_channel, self.channel_axis, self.input_rank))


def _lut_export_forward(input_tensor, lut_values, threshold, lut_values_bitwidth, eps, per_channel, channel_axis, input_rank):
    thr = torch.from_numpy(np.asarray(threshold).astype(np.float32)).to(input_tensor.device)
    lut = torch.from_numpy(np.asarray(lut_values).astype(np.float32)).to(input_tensor.device)
    return lut_quantizer_export(input_tensor, lut_values=lut, signed=True, threshold=thr, lut_values_bitwidth=lut_values_bitwidth,
                                eps=eps, per_channel=per_channel, channel_axis=channel_axis, input_rank=input_rank)


def _lut_export_symbolic(op_name, cls, g, input_tensor, num_bits, lut_values, threshold, lut_values_bitwidth, eps,
                         per_channel, channel_axis, input_rank):
    if not per_channel:
        channel_axis = 0 if channel_axis is None else channel_axis
        input_rank = 0 if input_rank is None else input_rank
    return g.op(f"{ONNX_CUSTOM_OP_DOMAIN}::{op_name}", input_tensor,
                g.op('Constant', value_t=torch.tensor(lut_values, dtype=torch.float32)),
                g.op('Constant', value_t=torch.tensor(threshold, dtype=torch.float32)),
                num_bits_i=num_bits, per_channel_i=int(per_channel), channel_axis_i=channel_axis,
                input_rank_i=input_rank, lut_values_bitwidth_i=lut_values_bitwidth, eps_f=eps,
                signed_i=int(cls.is_signed()), **cls._get_metadata_attributes()).setType(input_tensor.type())


class WeightsLUTSymmetricF(BaseWeightQuantizerAutogradFunction):
    @staticmethod
    def forward(ctx, input_tensor, num_bits, lut_values, threshold, lut_values_bitwidth, eps, per_channel,
                channel_axis, input_rank):
        return _lut_export_forward(input_tensor, lut_values, threshold, lut_values_bitwidth, eps, per_channel,
                                   channel_axis, input_rank)

    @staticmethod
    def symbolic(g, input_tensor, num_bits, lut_values, threshold, lut_values_bitwidth, eps, per_channel,
                 channel_axis, input_rank):
        return _lut_export_symbolic("WeightsLUTSymmetricQuantizer", WeightsLUTSymmetricF, g, input_tensor, num_bits,
                                    lut_values, threshold, lut_values_bitwidth, eps, per_channel, channel_axis, input_rank)
