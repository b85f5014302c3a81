"""The ONNX-export branch (`enable_custom_impl()` + `torch.jit` tracing) against fixtures produced by the
UNMODIFIED reference under the same conditions (tests/golden/make_golden_export.py: all nine quantizers, per-channel and
per-tensor, tie-dense inputs).  The traced function must (a) contain the same autograd-Function node the reference's graph
contains (`WeightsSymmetricF`, ..., the name the ONNX symbolic hangs on) and (b) reproduce the reference's output bit for bit.
These formulas are export shims in plain torch ops (true division), not the inference path; they run wherever the tensor
lives, so the comparison is made on CPU here and on the GPU under `-m gpu`.
"""
import json
import os

import numpy as np
import pytest
import torch

import golden_util as G
from mct_quantizers_b200.pytorch import quantizers as Q

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "golden_export.json")) as f:
    MANIFEST = json.load(f)
ARRAYS = np.load(os.path.join(HERE, "golden", "golden_export.npz"))
CASES = {c["name"]: c for c in MANIFEST["cases"]}


def _run(case, device):
    q = getattr(Q, case["cls"])(**case["args"])
    q.enable_custom_impl()
    x = G.to_torch(ARRAYS[f"{case['name']}/x"], case["x_dtype"], device)
    traced = torch.jit.trace(lambda t: q(t), x, check_trace=False)
    y = traced(x)
    ops = sorted({n.pyname() for n in traced.graph.nodes() if n.kind() == "prim::PythonOp"})
    kinds = {n.kind() for n in traced.graph.nodes()}
    return q, x, y, ops, kinds


def test_fixture_covers_all_nine_quantizers():
    assert MANIFEST["reference_version"] == "1.6.0" and len(CASES) == 50
    assert {c["cls"] for c in CASES.values()} == {
        "WeightsSymmetricInferableQuantizer", "WeightsPOTInferableQuantizer", "WeightsUniformInferableQuantizer",
        "WeightsLUTSymmetricInferableQuantizer", "WeightsLUTPOTInferableQuantizer", "ActivationSymmetricInferableQuantizer",
        "ActivationPOTInferableQuantizer", "ActivationUniformInferableQuantizer", "ActivationLutPOTInferableQuantizer"}
    # the export formulas really are a different arithmetic: 21 cases differ from the inference path on their tie-dense inputs
    assert sum(c["differs_from_inference_path"] > 0 for c in CASES.values()) >= 15


@pytest.mark.parametrize("name", sorted(CASES))
def test_traced_custom_impl_matches_reference_cpu(name):
    case = CASES[name]
    q, x, y, ops, kinds = _run(case, "cpu")
    assert ops == case["python_ops"], (ops, case["python_ops"])          # same *F node as the reference's traced graph
    assert not any(k.startswith("mctq::") for k in kinds)               # nothing the ONNX exporter does not know
    assert str(y.dtype).replace("torch.", "") == case["y_dtype"] and list(y.shape) == case["shape"]
    assert G.bits_equal(G.from_torch(y), ARRAYS[f"{name}/y"]), G.mismatch_report(G.from_torch(y), ARRAYS[f"{name}/y"], ARRAYS[f"{name}/x"])


def test_custom_impl_is_only_active_while_tracing():
    """Outside `torch.jit` tracing the flag changes nothing (reference: `self._use_custom_impl and torch.jit.is_tracing()`)."""
    q = Q.WeightsSymmetricInferableQuantizer(8, [1.0, 2.0], True, 0)
    q.enable_custom_impl()
    assert q._use_custom_impl is True
    if not torch.cuda.is_available():
        with pytest.raises(Exception):                 # the inference path needs the GPU: proof that it was taken
            q(torch.zeros(2, 4))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_traced_custom_impl_matches_reference_cuda(name):
    """Same branch with CUDA tensors: the reference's own torch formulas executed by libtorch's CUDA kernels, which are
    not bit-compatible with its CPU kernels -- `a * round(..) + b` is contracted to an FMA (1-ulp differences; measured on
    the B200: 52 of 108 elements of a per-channel uniform case) and `tensor / python_scalar` multiplies by the reciprocal
    (SURVEY 8a hazard 4), which moves elements within an ulp of a rounding tie by ONE quantization step (the fixture
    inputs are tie-dense on purpose: 162 of 4000 elements).  So against the CPU fixture: every element within a few ulp,
    or -- for at most a quarter of the tensor (measured maximum: 11 %) -- one quantization step away."""
    case = CASES[name]
    q, x, y, ops, kinds = _run(case, "cuda:0")
    assert y.is_cuda and ops == case["python_ops"]
    got, want = G.from_torch(y).astype(np.float64), ARRAYS[f"{name}/y"].astype(np.float64)
    close = np.isclose(got, want, rtol=2e-6, atol=1e-7)
    if close.all():
        return
    bits = case["args"].get("lut_values_bitwidth", case["args"]["num_bits"]) if "lut_values" in case["args"] else case["args"]["num_bits"]
    step = float(np.max(np.abs(want))) * 2 / (2 ** bits - 1) if "lut_values" not in case["args"] else float(np.max(np.abs(want)))
    far = np.abs(got - want)[~close]
    assert far.size <= got.size // 4 and float(far.max()) <= 1.01 * step + 1e-6, (far.size, far.max(), step)
