"""Activation holders fused with their elementwise producer (no counterpart in the reference).

In an MCT-exported model a `PytorchActivationQuantizationHolder` directly follows the op whose output it quantizes
(reference: mct_quantizers/pytorch/activation_quantization_holder.py:43-53).  When that op is a ReLU / ReLU6 or a
residual add and the holder's quantizer is a per-tensor affine one (ActivationSymmetric / POT / Uniform), both run in
ONE kernel (`mctq_fq_affine_scalar_pre`): the intermediate activation never goes to HBM (relu -> holder: 16 -> 8 bytes
per f32 element; add -> holder: 20 -> 12).  The result is bit-identical to the unfused pair.

    fused = fuse_activation_producers(model)          # torch.fx pass; returns a GraphModule
"""
import operator

import numpy as np
import torch
import torch.nn.functional as F

from mct_quantizers_b200 import _native, ops
from mct_quantizers_b200.pytorch.activation_quantization_holder import PytorchActivationQuantizationHolder
from mct_quantizers_b200.pytorch.quantizers.activation_inferable_quantizers.activation_symmetric_inferable_quantizer import \
    ActivationSymmetricInferableQuantizer
from mct_quantizers_b200.pytorch.quantizers.activation_inferable_quantizers.activation_uniform_inferable_quantizer import \
    ActivationUniformInferableQuantizer

PRE_OPS = {"relu": _native.PRE_RELU, "relu6": _native.PRE_RELU6, "add": _native.PRE_ADD, "add_relu": _native.PRE_ADD_RELU}


def affine_scalar_params(quantizer):
    """(scale, zero_point, qmin, qmax) of a per-tensor affine activation quantizer, or None for any other quantizer
    (LUT quantizers, user-defined ones)."""
    if isinstance(quantizer, ActivationUniformInferableQuantizer):
        return quantizer.scale, quantizer.zero_point, quantizer.min_quantized_domain, quantizer.max_quantized_domain
    if isinstance(quantizer, ActivationSymmetricInferableQuantizer):          # covers the POT subclass
        return quantizer.scales, quantizer.zero_points, quantizer.min_quantized_domain, quantizer.max_quantized_domain
    return None


class PytorchFusedActivationQuantizationHolder(PytorchActivationQuantizationHolder):
    """`holder(producer(x[, other]))` in one kernel.  `pre_op` is one of 'relu', 'relu6', 'add', 'add_relu'."""

    def __init__(self, activation_holder_quantizer, pre_op: str, **kwargs):
        super().__init__(activation_holder_quantizer, **kwargs)
        if pre_op not in PRE_OPS:
            raise ValueError(f"pre_op must be one of {sorted(PRE_OPS)} but is {pre_op!r}")
        if affine_scalar_params(activation_holder_quantizer) is None:
            raise TypeError("producer fusion needs a per-tensor affine activation quantizer "
                            "(ActivationSymmetric / POT / Uniform InferableQuantizer)")
        self.pre_op = pre_op

    def _unfused(self, inputs, other):
        x = inputs + other if self.pre_op in ("add", "add_relu") else inputs
        if self.pre_op in ("relu", "add_relu"):
            x = F.relu(x)
        elif self.pre_op == "relu6":
            x = F.relu6(x)
        return self.activation_holder_quantizer(x)

    def forward(self, inputs, other=None):
        q = self.activation_holder_quantizer
        two = self.pre_op in ("add", "add_relu")
        if two and other is None:
            raise TypeError(f"pre_op {self.pre_op!r} needs two inputs")
        scale, zp, qmin, qmax = affine_scalar_params(q)
        fusable = inputs.is_cuda and not (q._use_custom_impl and torch.jit.is_tracing()) and \
            int(qmax) - int(qmin) < (1 << 21) and \
            (not two or (torch.is_tensor(other) and other.shape == inputs.shape and other.dtype == inputs.dtype
                         and other.device == inputs.device))
        if not fusable:
            # broadcasting adds, host tensors, ONNX export, ranges of 2^21 codes or more (outside the fused kernel's fast
            # rounding): the plain pair (each op still runs on the GPU)
            return self._unfused(inputs, other)
        code = PRE_OPS[self.pre_op]
        if ops.direct_ok(inputs):
            with torch.no_grad():
                return ops.affine_scalar_pre_direct(inputs, other if two else None, code, float(np.float32(scale)), int(zp),
                                                    int(qmin), int(qmax))
        return torch.ops.mctq.fq_affine_scalar_pre(inputs.detach(), other.detach() if two else None, code, scale, zp, qmin, qmax)


_RELU_FUNCS = {F.relu, torch.relu}
_RELU6_FUNCS = {F.relu6}
_ADD_FUNCS = {operator.add, torch.add}


def _is_inplace(node, modules):
    """relu_(x) spellings: nn.ReLU(inplace=True), F.relu(x, inplace=True) / F.relu(x, True), x.relu_()."""
    if node.op == "call_module":
        return bool(getattr(modules.get(node.target), "inplace", False))
    if node.op == "call_function":
        return bool(node.kwargs.get("inplace", False)) or (len(node.args) > 1 and node.args[1] is True)
    return node.op == "call_method" and str(node.target).endswith("_")


def _kind(node, modules):
    """'relu' / 'relu6' / 'add' when `node` is a fusable producer, else None.  An IN-PLACE relu mutates its input: it is
    only fusable when nobody else can observe that input (the input node has this relu as its only user and is not a
    graph input), because the fused kernel leaves the input untouched."""
    if node.op in ("call_module", "call_function", "call_method") and _is_inplace(node, modules):
        src = node.args[0] if node.args else None
        if not isinstance(src, torch.fx.Node) or len(src.users) != 1 or src.op in ("placeholder", "get_attr"):
            return None
        if node.op == "call_function":
            if set(node.kwargs) - {"inplace"} or len(node.args) > 2:
                return None
            return "relu" if node.target in _RELU_FUNCS else ("relu6" if node.target in _RELU6_FUNCS else None)
    elif node.op == "call_function" and node.kwargs and set(node.kwargs) != {"inplace"}:
        return None
    if node.op == "call_module":
        m = modules.get(node.target)
        if isinstance(m, torch.nn.ReLU):
            return "relu"
        if isinstance(m, torch.nn.ReLU6):
            return "relu6"
        return None
    if node.op == "call_function":
        if node.target in _RELU_FUNCS and len(node.args) in (1, 2):
            return "relu"
        if node.target in _RELU6_FUNCS and len(node.args) in (1, 2):
            return "relu6"
        if node.target in _ADD_FUNCS and len(node.args) == 2 and not node.kwargs and \
                all(isinstance(a, torch.fx.Node) for a in node.args):
            return "add"
        return None
    if node.op == "call_method" and node.target == "relu" and len(node.args) == 1:
        return "relu"
    return None


class _LeafTracer(torch.fx.Tracer):
    """Keeps holders and quantization wrappers as call_module nodes (the default tracer would trace through them)."""

    def is_leaf_module(self, m, qualname):
        from mct_quantizers_b200.pytorch.quantize_wrapper import PytorchQuantizationWrapper
        if isinstance(m, (PytorchActivationQuantizationHolder, PytorchQuantizationWrapper)):
            return True
        return super().is_leaf_module(m, qualname)


def _producer_chain(value, modules):
    """(kind, sources, nodes_to_erase) for the fusable producer chain that ends in `value`, or None.  Every link must
    have the next one as its only consumer."""
    chain = []
    prod = value
    while isinstance(prod, torch.fx.Node) and prod.op == "call_method" and prod.target == "detach" and len(prod.users) == 1:
        chain.append(prod)                      # x.detach() in front of a traced-through quantizer call
        prod = prod.args[0]
    if not isinstance(prod, torch.fx.Node) or len(prod.users) != 1:
        return None
    kind = _kind(prod, modules)
    if kind is None:
        return None
    chain.append(prod)
    srcs = list(prod.args)
    if kind in ("relu", "relu6"):
        inner = prod.args[0]
        if kind == "relu" and isinstance(inner, torch.fx.Node) and len(inner.users) == 1 and _kind(inner, modules) == "add":
            kind, srcs = "add_relu", list(inner.args)        # relu(add(a, b)) -> holder
            chain.append(inner)
        else:
            srcs = [inner]
    return kind, srcs, chain


def fuse_activation_producers(model: torch.nn.Module) -> torch.fx.GraphModule:
    """torch.fx pass: rewrites  relu|relu6|add[->relu] -> activation fake-quant  chains whose intermediate values have no
    other consumer into ONE fused call.  Two spellings of the fake-quant are recognised: a
    `PytorchActivationQuantizationHolder` module call (a plain nn.Module is traced here with holders kept as leaves)
    and, in an already traced GraphModule, the `torch.ops.mctq.fq_affine_scalar` node a traced-through holder leaves
    behind.  Holders with a `quantization_bypass` flag, non-affine quantizers and producers whose result feeds anything
    else are left alone; tensor adds of different shapes fall back to the unfused pair at run time."""
    if isinstance(model, torch.fx.GraphModule):
        gm = model
    else:
        gm = torch.fx.GraphModule(model, _LeafTracer().trace(model))
    modules = dict(gm.named_modules())
    fq_targets = (torch.ops.mctq.fq_affine_scalar, torch.ops.mctq.fq_affine_scalar.default)
    n_fused = 0
    for node in list(gm.graph.nodes):
        if node.op == "call_module":
            holder = modules.get(node.target)
            if type(holder) is not PytorchActivationQuantizationHolder or len(node.args) != 1 or node.kwargs:
                continue
            if affine_scalar_params(holder.activation_holder_quantizer) is None:
                continue
            found = _producer_chain(node.args[0], modules)
            if found is None:
                continue
            kind, srcs, chain = found
            name = f"{node.target}_fused_{kind}"
            gm.add_submodule(name, PytorchFusedActivationQuantizationHolder(holder.activation_holder_quantizer, kind))
            with gm.graph.inserting_before(node):
                new = gm.graph.call_module(name, tuple(srcs))
        elif node.op == "call_function" and node.target in fq_targets and len(node.args) == 5 and not node.kwargs:
            found = _producer_chain(node.args[0], modules)
            if found is None:
                continue
            kind, srcs, chain = found
            other = srcs[1] if len(srcs) == 2 else None
            with gm.graph.inserting_before(node):
                new = gm.graph.call_function(torch.ops.mctq.fq_affine_scalar_pre,
                                             (srcs[0], other, PRE_OPS[kind]) + tuple(node.args[1:]))
        else:
            continue
        node.replace_all_uses_with(new)
        gm.graph.erase_node(node)
        for dead in chain:
            gm.graph.erase_node(dead)
        n_fused += 1
    gm.graph.lint()
    gm.recompile()
    gm.mctq_fused_sites = n_fused
    return gm
