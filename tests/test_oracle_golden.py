"""Pin the CPU oracle (C restatement + torch-CPU port + constructor-math restatement) against the
fixtures generated from the unmodified reference.  Bit-exact everywhere."""
import numpy as np
import pytest
import torch

import oracle
from oracle import torch_cpu_port as port

import golden_util as G

CASES = G.case_names()


def test_fixture_inventory():
    m = G.manifest()
    assert m["reference_version"] == "1.6.0"
    classes = {c["cls"] for c in m["cases"]}
    assert classes == {
        "WeightsSymmetricInferableQuantizer", "WeightsPOTInferableQuantizer", "WeightsUniformInferableQuantizer",
        "WeightsLUTSymmetricInferableQuantizer", "WeightsLUTPOTInferableQuantizer",
        "ActivationSymmetricInferableQuantizer", "ActivationPOTInferableQuantizer",
        "ActivationUniformInferableQuantizer", "ActivationLutPOTInferableQuantizer"}
    assert len(m["cases"]) >= 150


@pytest.mark.parametrize("name", CASES)
def test_c_oracle_matches_reference(name):
    case = G.get_case(name)
    out = G.oracle_run(case)
    assert G.bits_equal(out["y"], case["y"]), G.mismatch_report(out["y"], case["y"], case["x"])
    if case["idx"] is not None:
        assert G.bits_equal(out["idx"], case["idx"]), G.mismatch_report(out["idx"], case["idx"], case["x"])


@pytest.mark.parametrize("name", CASES)
def test_derived_params_match_reference(name):
    """The oracle's restatement of the constructor math reproduces the reference's attributes."""
    case = G.get_case(name)
    p = G.derive_params(case)
    meta = case["params"]
    if p["kind"] != "affine":
        return
    if "scales" in meta and meta["scales"]["kind"] == "tensor":
        assert G.bits_equal(p["scale"], case["p"]["scales"].reshape(-1))
        assert np.array_equal(p["zp"], case["p"]["zero_points"].reshape(-1))
    elif "scales" in meta:                       # activation symmetric: python float
        assert p["scale_py"] == meta["scales"]["value"]
    if "scale" in meta:                          # activation uniform: python float / int
        assert p["scale_py"] == meta["scale"]["value"]
        assert p["zp_py"] == meta["zero_point"]["value"]
        assert p["min_range"] == meta["min_range"]["value"] and p["max_range"] == meta["max_range"]["value"]
    if "adjusted_min_range_np" in meta:
        assert G.bits_equal(p["min_range"], case["p"]["adjusted_min_range_np"])
        assert G.bits_equal(p["max_range"], case["p"]["adjusted_max_range_np"])
    assert p["qmin"] == meta["min_quantized_domain"]["value"] and p["qmax"] == meta["max_quantized_domain"]["value"]


@pytest.mark.parametrize("name", CASES)
def test_torch_cpu_port_matches_reference(name):
    case = G.get_case(name)
    p = G.derive_params(case)
    a = case["args"]
    x = G.to_torch(case["x"], case["x_dtype"])
    if p["kind"] == "affine":
        if "scale_py" in p:
            y = port.affine_scalar_qparams(x, p["scale_py"], p.get("zp_py", 0), p["qmin"], p["qmax"])
        elif p["C"] == 1 and not a.get("per_channel"):
            y = port.affine_tensor_qparams(x, torch.from_numpy(p["scale"]), torch.from_numpy(p["zp"]), p["qmin"], p["qmax"])
        else:
            y = port.affine_per_channel(x, torch.from_numpy(p["scale"]), torch.from_numpy(p["zp"]), a["channel_axis"],
                                        p["qmin"], p["qmax"])
        assert G.bits_equal(G.from_torch(y), case["y"])
    else:
        lut = torch.from_numpy(p["lut"])
        if p["act"]:
            y, idx = port.lut_fake_quant(x, lut, p["signed"], p["thr"], p["bw"], p["eps"], want_idx=True)
        else:
            y, idx = port.lut_fake_quant(x, lut, True, torch.from_numpy(p["thr"]), p["bw"], p["eps"],
                                         per_channel=a["per_channel"], channel_axis=a.get("channel_axis"),
                                         input_rank=a.get("input_rank"), want_idx=True)
        assert G.bits_equal(G.from_torch(y), case["y"])
        assert np.array_equal(idx.numpy().astype(np.int32), case["idx"])


def test_range_fix_known_answers():
    """Constants asserted by the reference's own tests
    (tests/pytorch_tests/test_fln_activation_quantizer_holder.py:42-45)."""
    ka = G.manifest()["range_fix_known_answers"]
    lo, hi, scale, zp, _, _ = port.activation_uniform_qparams([ka["min"]], [ka["max"]], ka["num_bits"])
    assert (lo, hi, scale, zp) == (ka["min_range"], ka["max_range"], ka["scale"], ka["zero_point"])
    assert np.isclose(lo, -4.03149606299213) and np.isclose(hi, 3.96850393700787) and np.isclose(scale, 0.062992125984252)


@pytest.mark.parametrize("dtype", ["bfloat16", "float16"])
def test_half_conversions_match_torch(dtype):
    rng = np.random.default_rng(0)
    v = np.concatenate([rng.normal(0, 1, 20000), rng.normal(0, 1e-6, 5000), rng.normal(0, 3e4, 5000),
                        [0.0, -0.0, 65504.0, 65519.9, 65520.0, 1e-8, 6e-8, 2.98e-8, 2.9802322e-8]]).astype(np.float32)
    tag = G.DT_TAG[dtype]
    got = oracle.f32_to_half_bits(v, tag)
    want = G.from_torch(torch.from_numpy(v).to(G.TORCH_DT[dtype]))
    assert np.array_equal(got, want)
    allbits = np.arange(1 << 16, dtype=np.uint16)
    back = oracle.half_bits_to_f32(allbits, tag)
    want_f = G.to_torch(allbits, dtype).float().numpy()
    assert np.array_equal(back.view(np.uint32)[~np.isnan(want_f)], want_f.view(np.uint32)[~np.isnan(want_f)])


def test_codes_dequantise_to_y():
    """codes are not produced by the reference; pin their meaning: dequant(codes) == y bitwise."""
    for name in ("w_sym_b8_pc_6x5x3x3_ax0", "w_uni_b4_pc_4x7x3x5_ax1", "a_uni_b8_straddle"):
        case = G.get_case(name)
        out = G.oracle_run(case)
        p = out["p"]
        y = oracle.dequant_affine(out["codes"], p["scale"], p["zp"], p["C"], p["inner"])
        assert G.bits_equal(y, case["y"])
        assert out["codes"].min() >= p["qmin"] and out["codes"].max() <= p["qmax"]
