"""Host logic of the LUT path: the search table built by the C ABI (mctq_lut_build_table), emulated in numpy
with exact f32 division, must reproduce the reference's argmin indices and outputs on every LUT fixture.
(No GPU needed: the table builder is host code inside libmctq_sm100.so.)"""
import struct

import numpy as np
import pytest

import golden_util as G
from mct_quantizers_b200 import _native

LUT_CASES = [n for n in G.case_names() if n.startswith(("wl_", "al_"))]


def parse_table(blob, K):
    magic, K_, Ks, P, levels, pos0, bw, signed = struct.unpack_from("<Iiiiiiii", blob, 0)
    mult, = struct.unpack_from("<f", blob, 32)
    assert magic == 0x4d514c54 and K_ == K
    off = 64
    tau = np.frombuffer(blob, dtype=np.float32, count=P - 1, offset=off)
    cq = np.frombuffer(blob, dtype=np.float32, count=P, offset=off + 4 * (P - 1))
    orig = np.frombuffer(blob, dtype=np.uint8, count=P, offset=off + 4 * (P - 1) + 4 * P)
    return dict(Ks=Ks, P=P, levels=levels, pos0=pos0, mult=mult, tau=tau, cq=cq, orig=orig)


def emulate(case, p):
    lut = p["lut"]
    blob = _native.build_lut_table(lut, p["bw"], p["signed"])
    t = parse_table(blob, lut.size)
    x = case["x"]
    if case["x_dtype"] != "float32":
        import oracle
        x = oracle.half_bits_to_f32(x, G.DT_TAG[case["x_dtype"]])
    x = x.astype(np.float32)
    if p["act"]:
        d = np.float32(np.float64(p["thr"]) + np.float64(p["eps"]))
        thr = np.float32(p["thr"])
        q = (x / d).astype(np.float32)
        if case["x_dtype"] != "float32":
            import oracle
            tag = G.DT_TAG[case["x_dtype"]]
            q = oracle.half_bits_to_f32(oracle.f32_to_half_bits(q, tag), tag)
        thr_b = thr
    else:
        C, inner = p["C"], p["inner"]
        ch = (np.arange(x.size) // inner) % C
        thr_e = p["thr"].astype(np.float32)[ch].reshape(x.shape)
        d = (thr_e + np.float32(p["eps"])).astype(np.float32)
        q = (x / d).astype(np.float32)
        thr_b = thr_e
    # pos = number of thresholds strictly below q, via the same power-of-two stepping as the kernel
    pos = np.zeros(q.shape, dtype=np.int64)
    step = t["P"] >> 1
    tau = np.concatenate([t["tau"], [np.inf]]).astype(np.float32)
    while step > 0:
        pos = pos + np.where(q > tau[pos + step - 1], step, 0)
        step >>= 1
    pos = np.where(np.isnan(x), t["pos0"], pos)
    idx = t["orig"][pos].astype(np.int32)
    y = (t["cq"][pos] * thr_b).astype(np.float32)
    return y, idx


@pytest.mark.parametrize("name", LUT_CASES)
def test_table_search_reproduces_reference(name):
    case = G.get_case(name)
    p = G.derive_params(case)
    with np.errstate(all="ignore"):
        y, idx = emulate(case, p)
    assert G.bits_equal(idx.reshape(case["idx"].shape), case["idx"]), G.mismatch_report(idx, case["idx"], case["x"])
    assert G.bits_equal(y.reshape(case["y"].shape), case["y"]), G.mismatch_report(y, case["y"], case["x"])


def test_table_rejects_bad_input():
    with pytest.raises(_native.MctqError):
        _native.build_lut_table(np.zeros(300, np.float32), 8, True)
    with pytest.raises(_native.MctqError):
        _native.build_lut_table(np.array([np.nan, 1.0], np.float32), 8, True)
