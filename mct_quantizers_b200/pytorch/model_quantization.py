"""Whole-model weight quantization in one launch.

The reference re-quantizes each wrapped weight with its own kernel(s) on every forward
(mct_quantizers/pytorch/quantize_wrapper.py:228-240): 53 launches (+ 106 host syncs) for MobileNetV2.  Here all
affine weight quantizers of a model are gathered into one descriptor table and run as ONE kernel
(mctq_fq_affine_multi), all LUT weight quantizers as a second one (mctq_fq_lut_prepared_multi); user-defined quantizers
and tensors outside the prepared paths keep their own call."""
from typing import Dict, List

import torch

from mct_quantizers_b200.ops import MultiTensorPlan, LutMultiPlan


def _is_affine_weight_quantizer(q) -> bool:
    from mct_quantizers_b200.pytorch.quantizers import WeightsSymmetricInferableQuantizer, WeightsUniformInferableQuantizer
    return isinstance(q, (WeightsSymmetricInferableQuantizer, WeightsUniformInferableQuantizer)) and not q._use_custom_impl


def _lut_item(q, w):
    """LutMultiPlan item of a LUT weight quantizer (WeightsLUTSymmetric / WeightsLUTPOT), or None."""
    from mct_quantizers_b200.pytorch.quantizers import WeightsLUTSymmetricInferableQuantizer
    from mct_quantizers_b200.pytorch.quantizer_utils import lut_search_table
    if not isinstance(q, WeightsLUTSymmetricInferableQuantizer) or q._use_custom_impl or not w.is_cuda:
        return None
    if q.per_channel and q.input_rank is not None and q.input_rank != w.dim():
        return None                                   # the per-layer call raises the reference's error
    if q.__dict__.get('_search_table') is None:
        q._search_table = lut_search_table(q._lut_values_np, q.lut_values_bitwidth, True)
    thr, = q._on(w.device, q._threshold_torch)
    thr = thr.reshape(-1)
    item = (w.detach(), q._search_table, int(q._lut_values_torch.numel()), thr, bool(q.per_channel),
            int(q.channel_axis) if q.per_channel else 0, float(q.eps))
    return item if LutMultiPlan.accepts(item) else None


def plan_for(weight_vars) -> "WeightPlan":
    return WeightPlan(weight_vars)


class WeightPlan:
    """Pre-built launch plan over (name, weight, quantizer) triples.  Keeps output buffers; run() refreshes them."""

    def __init__(self, weight_vars):
        self.names, self.fused_idx, self.lut_idx, self.other = [], [], [], []
        items, lut_items = [], []
        for k, (name, w, q) in enumerate(weight_vars):
            self.names.append(name)
            ok_dtype = w.is_cuda and w.dtype in (torch.float32, torch.bfloat16, torch.float16)
            lut_item = _lut_item(q, w) if ok_dtype else None
            if _is_affine_weight_quantizer(q) and ok_dtype:
                w.requires_grad = False
                scales, zps = q._on(w.device, q.scales, q.zero_points)
                items.append((w.detach(), scales.flatten(), zps.flatten(), q.channel_axis if q.per_channel else None,
                              q.min_quantized_domain, q.max_quantized_domain))
                self.fused_idx.append(k)
            elif lut_item is not None:
                w.requires_grad = False
                lut_items.append(lut_item)
                self.lut_idx.append(k)
            else:
                self.other.append((k, w, q))
        self.plan = MultiTensorPlan(items) if items else None
        self.lut_plan = LutMultiPlan(lut_items) if lut_items else None
        self.n = len(weight_vars)

    def run(self) -> List[torch.Tensor]:
        out = [None] * self.n
        if self.plan is not None:
            for k, y in zip(self.fused_idx, self.plan.run()):
                out[k] = y
        if self.lut_plan is not None:
            for k, y in zip(self.lut_idx, self.lut_plan.run()):
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


class ModelWeightPlan:
    """Whole-model weight quantization hoisted out of the per-layer forward.

    The reference re-quantizes every wrapped weight inside every `PytorchQuantizationWrapper.forward`
    (mct_quantizers/pytorch/quantize_wrapper.py:228-240).  A plan gathers the weights of ALL wrappers of a model,
    quantizes them with one multi-tensor launch (`refresh()`), installs the results in the wrapped layers and tells the
    wrappers to skip their own quantization while the plan is active:

        plan = plan_model_weights(model)      # builds the descriptor table once, output buffers are reused
        with plan:                            # refresh(): ONE kernel for every affine weight quantizer of the model
            y = model(x)                      # wrappers run their layers on the installed weights, no per-layer launches

    or, for inference with constant weights, `plan.enable()` once (and `plan.refresh()` again only if the float weights
    change).  Results are identical to the per-layer path (same kernels' arithmetic, checked by tests)."""

    def __init__(self, model: torch.nn.Module):
        from mct_quantizers_b200.pytorch.quantize_wrapper import PytorchQuantizationWrapper
        self.wrappers, triples, self._owner = [], [], []
        for mod in model.modules():
            if isinstance(mod, PytorchQuantizationWrapper) and mod.is_weights_quantization:
                self.wrappers.append(mod)
                for name, w, q in mod.get_weights_vars():
                    triples.append((name, w, q))
                    self._owner.append(mod)
        self.plan = WeightPlan(triples) if triples else None
        self.active = False

    def refresh(self):
        """Quantize every weight (one launch for the affine quantizers) and install the results."""
        if self.plan is None:
            return
        per_wrapper = {}
        for owner, name, y in zip(self._owner, self.plan.names, self.plan.run()):
            per_wrapper.setdefault(id(owner), (owner, {}))[1][name] = y
        for owner, weights in per_wrapper.values():
            owner.set_quantize_weights(weights)

    def enable(self):
        self.refresh()
        for w in self.wrappers:
            w._prequantized = True
        self.active = True
        return self

    def disable(self):
        for w in self.wrappers:
            w._prequantized = False
        self.active = False

    def __enter__(self):
        return self.enable()

    def __exit__(self, *exc):
        self.disable()


def plan_model_weights(model: torch.nn.Module) -> ModelWeightPlan:
    return ModelWeightPlan(model)
