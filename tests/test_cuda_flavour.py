"""The reference's numbers depend on the machine it runs on: `tensor / python_number` is a true division in libtorch's
CPU kernels and a multiplication by the (double -> f32) reciprocal in its CUDA kernels (mct_quantizers_b200/pytorch/
quantizer_utils.py: reference_arithmetic).  tests/golden/golden_cuda_flavour.* holds what the UNMODIFIED reference produced
on a B200 (generator: tests/golden/make_golden_cuda_flavour.py); `reference_arithmetic("cuda")` has to reproduce it bit for
bit -- constructor-derived parameters (CPU test), the oracle's CUDA flavour (CPU test) and the CUDA path (GPU test)."""
import json
import os

import numpy as np
import pytest
import torch

import golden_util as G
import mct_quantizers_b200 as mctq
import oracle
from mct_quantizers_b200.pytorch import quantizers as Q

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "golden_cuda_flavour.json")) as f:
    MANIFEST = json.load(f)
ARR = np.load(os.path.join(HERE, "golden", "golden_cuda_flavour.npz"))
CASES = {c["name"]: c for c in MANIFEST["cases"]}


@pytest.fixture()
def cuda_flavour():
    prev = mctq.reference_arithmetic("cuda")
    yield
    mctq.reference_arithmetic(prev)


def _build(case):
    return getattr(Q, case["cls"])(**case["args"])


def _param(q, name):
    v = getattr(q, name)
    return v.detach().cpu().flatten().numpy() if isinstance(v, torch.Tensor) else v


@pytest.mark.parametrize("name", [n for n in CASES if "_uni_" in n])
def test_constructor_parameters_match_the_reference_on_cuda(name, cuda_flavour):
    case = CASES[name]
    q = _build(case)
    for p, kind in case["params"].items():
        got = _param(q, p)
        if kind == "tensor":
            want = ARR[f"{name}/p/{p}"]
            assert G.bits_equal(np.asarray(got, dtype=want.dtype), want), (name, p)
        else:
            assert got == kind, (name, p, got, kind)


def test_the_two_flavours_really_differ():
    """Guards the fixture: with the default (CPU) flavour a good part of the uniform cases must NOT match the CUDA-machine
    parameters -- otherwise this file would pin nothing."""
    assert mctq.reference_arithmetic() == "cpu"
    differ = 0
    for name, case in CASES.items():
        if not name.startswith("cw_uni_"):
            continue
        q = _build(case)
        differ += not G.bits_equal(_param(q, "scales"), ARR[f"{name}/p/scales"])
    assert differ >= 10


def _oracle(case, name, cuda):
    a = case["args"]
    x = ARR[f"{name}/x"]
    tag = G.DT_TAG[case["x_dtype"]]
    if "_lut_" in name:
        return oracle.fq_lut(x, tag, np.asarray(a["lut_values"], np.float32), a["threshold"][0], 1, 1, a["lut_values_bitwidth"],
                             a["signed"], 1e-8, activation_mode=True, cuda_flavour=cuda)
    q = _build(case)
    if case["cls"].startswith("Weights"):
        sc, zp = _param(q, "scales"), _param(q, "zero_points")
        return oracle.fq_affine(x, tag, sc, zp, len(sc), x.shape[1], 0, 2 ** a["num_bits"] - 1)
    return oracle.fq_affine(x, tag, np.array([q.scale], np.float64).astype(np.float32), np.array([q.zero_point], np.int32), 1, 1,
                            0, 2 ** a["num_bits"] - 1)


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_reproduces_the_reference_on_cuda(name, cuda_flavour):
    case = CASES[name]
    want = ARR[f"{name}/y"]
    got = np.asarray(_oracle(case, name, True)).reshape(want.shape)
    assert G.bits_equal(got, want), name


def test_lut_ties_move_between_the_flavours():
    """At least one LUT fixture must come out differently under the CPU flavour (an exact tie resolved the other way)."""
    moved = 0
    for name, case in CASES.items():
        if "_lut_" in name:
            want = ARR[f"{name}/y"]
            moved += not G.bits_equal(np.asarray(_oracle(case, name, False)).reshape(want.shape), want)
    assert moved >= 1


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_cuda_path_reproduces_the_reference_on_cuda(name, cuda_flavour):
    case = CASES[name]
    q = _build(case)
    x = G.to_torch(ARR[f"{name}/x"], case["x_dtype"], "cuda:0")
    y = q(x)
    want = ARR[f"{name}/y"]
    assert str(y.dtype).replace("torch.", "") == case["y_dtype"]
    assert G.bits_equal(G.from_torch(y).reshape(want.shape), want), name
    if "_lut_" in name:
        # generic kernel too (the prepared one ran above): misaligned view
        buf = torch.empty(x.numel() + 1, dtype=x.dtype, device=x.device)
        buf[1:] = x.flatten()
        y2 = q(buf[1:].view(x.shape))
        assert torch.equal(y2, y)
