"""Whole-model weight quantization in one launch.

The reference re-quantizes each wrapped weight with its own kernel(s) on every forward
(mct_quantizers/pytorch/quantize_wrapper.py:228-240): 53 launches (+ 106 host syncs) for MobileNetV2.  Here all
affine weight quantizers of a model are gathered into one descriptor table and run as ONE kernel
(mctq_fq_affine_multi); LUT quantizers and user-defined quantizers keep their own call."""
from typing import Dict, List

import torch

from mct_quantizers_b200.ops import MultiTensorPlan


def _is_affine_weight_quantizer(q) -> bool:
    from mct_quantizers_b200.pytorch.quantizers import WeightsSymmetricInferableQuantizer, WeightsUniformInferableQuantizer
    return isinstance(q, (WeightsSymmetricInferableQuantizer, WeightsUniformInferableQuantizer)) and not q._use_custom_impl


def plan_for(weight_vars) -> "WeightPlan":
    return WeightPlan(weight_vars)


class WeightPlan:
    """Pre-built launch plan over (name, weight, quantizer) triples.  Keeps output buffers; run() refreshes them."""

    def __init__(self, weight_vars):
        self.names, self.fused_idx, self.other = [], [], []
        items = []
        for k, (name, w, q) in enumerate(weight_vars):
            self.names.append(name)
            if _is_affine_weight_quantizer(q) and w.is_cuda and w.dtype in (torch.float32, torch.bfloat16, torch.float16):
                w.requires_grad = False
                scales, zps = q._on(w.device, q.scales, q.zero_points)
                items.append((w.detach(), scales.flatten(), zps.flatten(), q.channel_axis if q.per_channel else None,
                              q.min_quantized_domain, q.max_quantized_domain))
                self.fused_idx.append(k)
            else:
                self.other.append((k, w, q))
        self.plan = MultiTensorPlan(items) if items else None
        self.n = len(weight_vars)

    def run(self) -> List[torch.Tensor]:
        out = [None] * self.n
        if self.plan is not None:
            for k, y in zip(self.fused_idx, self.plan.run()):
                out[k] = y
        for k, w, q in self.other:
            out[k] = q(w)
        return out


def quantize_weight_vars(weight_vars) -> Dict[str, torch.Tensor]:
    plan = WeightPlan(weight_vars)
    return dict(zip(plan.names, plan.run()))


def quantize_model_weights(model: torch.nn.Module) -> Dict[str, Dict[str, torch.Tensor]]:
    """{wrapper module name: {weight name: quantized weight}} for every PytorchQuantizationWrapper in `model`,
    all affine quantizers of the whole model in one kernel launch."""
    from mct_quantizers_b200.pytorch.quantize_wrapper import PytorchQuantizationWrapper
    triples, owners = [], []
    for mod_name, mod in model.named_modules():
        if isinstance(mod, PytorchQuantizationWrapper):
            for name, w, q in mod.get_weights_vars():
                triples.append((name, w, q))
                owners.append(mod_name)
    result: Dict[str, Dict[str, torch.Tensor]] = {}
    if not triples:
        return result
    plan = WeightPlan(triples)
    for owner, name, y in zip(owners, plan.names, plan.run()):
        result.setdefault(owner, {})[name] = y
    return result
