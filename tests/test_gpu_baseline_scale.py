"""Parity at BASELINE.json's full sizes, CUDA path against the CPU oracle on row / element windows:

C3  WeightsLUTSymmetric 4-bit, K = 16, lut_values_bitwidth = 8, per-channel axis 0 on Llama-7B-shaped (11008, 4096) and
    (4096, 11008) matrices, f32 and bf16 -- through the per-layer call AND through WeightPlan (one multi-tensor launch);
    values and packed 4-bit indices; windows: first rows, last rows, random rows and rows planted with the decision
    boundaries of `lut_quantizer` (every centroid midpoint -3..+3 ulp, in the input dtype's grid).
    Reference semantics: mct_quantizers/pytorch/quantizer_utils.py:95-170.
C4  ActivationSymmetric 8-bit signed, thr 4.0 / 3.7, bf16, (256, 197, 768) and (256, 197, 3072).
C5  the 16 GB point of the size sweep: 2^32 f32 elements (+ a ragged tail) in ONE call, head / 2^31 / 2^32 / tail windows.

The oracle finishes a window in milliseconds; whole tensors are covered by size-independent properties (plan output ==
per-layer output bit for bit, index range, idempotence of the affine path).
"""
import numpy as np
import pytest
import torch

import golden_util as G
import oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
LUT = [float(v) for v in sorted(np.random.default_rng(0).choice(np.arange(-128, 128), size=16, replace=False))]


@pytest.fixture(scope="module")
def Q():
    from mct_quantizers_b200.pytorch import quantizers
    return quantizers


def _tag(dtype):
    return {torch.float32: oracle.F32, torch.bfloat16: oracle.BF16, torch.float16: oracle.F16}[dtype]


def _bits(t):
    return G.from_torch(t)


def _boundary_row(lut, thr, dtype, length, rng):
    """One row (numpy, raw bits of `dtype`) holding, for every pair of adjacent centroids, the input values around the
    point where the nearest centroid switches: x = mid / 2^(bw-1) * thr, -3..+3 ulp in the INPUT grid; the rest random."""
    mids = (np.asarray(lut[:-1], np.float64) + np.asarray(lut[1:], np.float64)) / 2
    edges = np.concatenate([mids, np.asarray(lut, np.float64), [-128.0, 127.0]])
    base = (edges / 128.0 * float(thr)).astype(np.float32)
    row = (rng.standard_normal(length) * 0.3 * thr).astype(np.float32)
    if dtype == torch.float32:
        pts = []
        for k in range(-3, 4):
            pts.append((base.view(np.int32) + k).view(np.float32))
        pts = np.concatenate(pts)
        pts = pts[np.abs(pts) <= thr]
        row[:pts.size] = pts
        return row
    tag = _tag(dtype)
    hb = oracle.f32_to_half_bits(row, tag).astype(np.uint16)
    bb = oracle.f32_to_half_bits(base, tag).astype(np.int32)
    pts = np.concatenate([(bb + k) for k in range(-3, 4)]).astype(np.uint16)
    ok = np.abs(oracle.half_bits_to_f32(pts, tag)) <= thr
    pts = pts[ok]
    hb[:pts.size] = pts
    return hb


def _make_matrix(shape, dtype, seed):
    """N(0, 0.02) weights with 8 boundary-planted rows in the middle; thresholds = row maxima (BASELINE C3)."""
    g = torch.Generator(device=DEV).manual_seed(seed)
    W = torch.empty(shape, device=DEV).normal_(0, 0.02, generator=g).to(dtype)
    rng = np.random.default_rng(seed)
    planted = list(range(shape[0] // 2, shape[0] // 2 + 8))
    for r in planted:
        thr = float(W[r].float().abs().max())
        row = _boundary_row(LUT, thr, dtype, shape[1], rng)
        if dtype == torch.float32:
            W[r] = torch.from_numpy(row).to(DEV)
        else:
            W[r] = torch.from_numpy(row.view(np.int16)).to(DEV).view(dtype)
    thr = [float(v) for v in W.float().abs().amax(1).double().cpu()]
    return W, thr, planted


def _row_windows(C, planted, seed):
    rng = np.random.default_rng(seed)
    r0 = int(rng.integers(8, C - 16))
    return [(0, 8), (C - 8, C), (planted[0], planted[-1] + 1), (r0, r0 + 5)]


def _check_lut_rows(W, thr, y, idx4, lo, hi, what):
    inner = W.shape[1]
    sub = W[lo:hi].contiguous()
    want, want_idx = oracle.fq_lut(_bits(sub), _tag(W.dtype), np.asarray(LUT, np.float32),
                                   np.asarray(thr[lo:hi], np.float64).astype(np.float32), hi - lo, inner, 8, True, 1e-8, want_idx=True)
    got = y[lo:hi].cpu().numpy()
    assert np.array_equal(got.view(np.uint32), np.asarray(want).reshape(got.shape).view(np.uint32)), what
    if idx4 is not None:
        assert inner % 2 == 0
        packed = idx4[lo * inner // 2: hi * inner // 2].cpu().numpy()
        un = np.stack([packed & 0xF, packed >> 4], 1).reshape(-1).astype(np.int32)
        assert np.array_equal(un, np.asarray(want_idx).reshape(-1)), what + " (indices)"


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("shape", [(11008, 4096), (4096, 11008)], ids=["11008x4096", "4096x11008"])
def test_c3_llama_lut_per_layer_and_plan(shape, dtype, Q):
    from mct_quantizers_b200.pytorch.model_quantization import WeightPlan
    from mct_quantizers_b200.pytorch.quantizer_utils import lut_search_table
    W, thr, planted = _make_matrix(shape, dtype, seed=shape[0] + (1 if dtype == torch.float32 else 2))
    q = Q.WeightsLUTSymmetricInferableQuantizer(4, LUT, thr, True, 0, 2)
    y = q(W)                                                           # per-layer call (prepared kernel)
    assert y.dtype == torch.float32 and y.shape == W.shape
    table = lut_search_table(np.asarray(LUT, np.float32), 8, True)
    thr_t = torch.tensor(thr, dtype=torch.float32, device=DEV)
    idx4 = torch.ops.mctq.lut_indices(W, table, 16, thr_t, True, 0, 1e-8, 2)
    idx8 = torch.ops.mctq.lut_indices(W, table, 16, thr_t, True, 0, 1e-8, 1)
    assert int(idx8.max()) <= 15
    # a second matrix of the other Llama shape in the same plan, so that the multi-tensor launch walks both layouts
    other_shape = (shape[1], shape[0])
    W2, thr2, planted2 = _make_matrix(other_shape, dtype, seed=77)
    q2 = Q.WeightsLUTSymmetricInferableQuantizer(4, LUT, thr2, True, 0, 2)
    plan = WeightPlan([("a", W, q), ("b", W2, q2)])
    assert plan.lut_plan is not None and not plan.other
    ya, yb = plan.run()
    torch.cuda.synchronize()
    assert torch.equal(ya.view(torch.int32), y.view(torch.int32))     # whole tensor: one launch == per-layer call
    assert torch.equal(yb.view(torch.int32), q2(W2).view(torch.int32))
    for lo, hi in _row_windows(shape[0], planted, 5):
        _check_lut_rows(W, thr, y, idx4, lo, hi, f"per-layer rows {lo}:{hi} of {shape} {dtype}")
        _check_lut_rows(W, thr, ya, None, lo, hi, f"WeightPlan rows {lo}:{hi} of {shape} {dtype}")
    for lo, hi in _row_windows(other_shape[0], planted2, 6):
        _check_lut_rows(W2, thr2, yb, None, lo, hi, f"WeightPlan rows {lo}:{hi} of {other_shape} {dtype}")
    # indices agree between the two wire formats over the whole tensor
    flat8 = idx8.reshape(-1)
    assert torch.equal(idx4, (flat8[0::2] | (flat8[1::2] << 4)))


@pytest.mark.parametrize("feat", [768, 3072])
@pytest.mark.parametrize("thr", [4.0, 3.7])
def test_c4_vit_activation_symmetric_bf16(feat, thr, Q):
    import mct_quantizers_b200 as mctq
    g = torch.Generator(device=DEV).manual_seed(feat)
    x = torch.empty((256, 197, feat), device=DEV).normal_(0, 1, generator=g).bfloat16()
    q = Q.ActivationSymmetricInferableQuantizer(8, [thr], True)
    h = mctq.PytorchActivationQuantizationHolder(q)
    y = h(x)
    assert y.dtype == torch.bfloat16 and y.shape == x.shape
    assert torch.equal(h(y).view(torch.int16), y.view(torch.int16))             # idempotent over the whole tensor
    scale = np.array([q.scales], np.float64).astype(np.float32)
    n = x.numel()
    xf, yf = x.reshape(-1), y.reshape(-1)
    for lo in (0, (n // 3) | 1, n - 70001):
        hi = min(lo + 70001, n)
        want = oracle.fq_affine(_bits(xf[lo:hi]), oracle.BF16, scale, np.zeros(1, np.int32), 1, 1, -128, 127)
        assert G.bits_equal(_bits(yf[lo:hi]), np.asarray(want).reshape(-1)), (feat, thr, lo)


def test_c5_16gb_f32_point(Q):
    """2^32 f32 elements (16 GiB in, 16 GiB out) + a ragged tail in ONE call: 64-bit tile indexing."""
    n = (1 << 32) + 4099
    free, _ = torch.cuda.mem_get_info()
    if free < 2 * n * 4 + (4 << 30):
        pytest.skip("not enough free device memory for the 16 GB point")
    g = torch.Generator(device=DEV).manual_seed(16)
    x = torch.empty(n, device=DEV)
    step = 1 << 28
    for s in range(0, n, step):
        x[s:s + step].uniform_(-50, 50, generator=g)
    for q, qmin, qmax in ((Q.ActivationSymmetricInferableQuantizer(8, [4.0], True), -128, 127),
                          (Q.ActivationUniformInferableQuantizer(8, [-1.0], [2.3]), 0, 255)):
        y = q(x)
        if qmin < 0:
            scale, zp = np.array([q.scales], np.float64).astype(np.float32), np.zeros(1, np.int32)
        else:
            scale, zp = np.array([q.scale], np.float64).astype(np.float32), np.array([q.zero_point], np.int32)
        for lo in (0, (1 << 31) - 50000, (1 << 32) - 50000, n - 100000):
            hi = min(lo + 100000, n)
            want = oracle.fq_affine(x[lo:hi].cpu().numpy(), oracle.F32, scale, zp, 1, 1, qmin, qmax)
            assert G.bits_equal(y[lo:hi].cpu().numpy(), np.asarray(want).reshape(-1)), (qmin, lo)
        del y
    del x
    torch.cuda.empty_cache()


def test_lut_beyond_2_31_elements(Q):
    """LUT quantizers on 2^31 + 2^20 + 4099 bf16 elements in ONE call (4 GiB in, 8 GiB of f32 out): 64-bit tile indexing of the
    prepared LUT kernels -- the per-tensor activation flavour (every eager op rounds to bf16) and per-channel weights with
    rows of 2^20 elements (channel = tile_start / inner in 64-bit arithmetic)."""
    n = (1 << 31) + (1 << 20) + 4096 + 3
    free, _ = torch.cuda.mem_get_info()
    if free < n * 2 + n * 4 + (6 << 30):
        pytest.skip("not enough free device memory")
    g = torch.Generator(device=DEV).manual_seed(31)
    x = torch.empty(n, device=DEV, dtype=torch.bfloat16)
    step = 1 << 28
    for s in range(0, n, step):
        x[s:s + step].uniform_(-3, 3, generator=g)
    lut = [-128.0, -100.0, -64.0, -30.0, -9.0, 0.0, 7.0, 21.0, 50.0, 90.0, 127.0]
    qa = Q.ActivationLutPOTInferableQuantizer(4, lut, [2.0], True)
    y = qa(x)
    assert y.dtype == torch.float32 and y.numel() == n
    for lo in (0, (1 << 30) - 33333, (1 << 31) - 30000, n - 60000):
        hi = min(lo + 60000, n)
        want = oracle.fq_lut(_bits(x[lo:hi]), oracle.BF16, np.asarray(lut, np.float32), 2.0, 1, 1, 8, True, 1e-8, activation_mode=True)
        assert G.bits_equal(y[lo:hi].cpu().numpy(), np.asarray(want).reshape(-1)), lo
    del y
    torch.cuda.empty_cache()
    # per-channel weights: 2049 rows of 2^20 elements (the tail of x is left out)
    C, L = 2049, 1 << 20
    rng = np.random.default_rng(2049)
    thr = [float(np.float32(v)) for v in rng.uniform(0.5, 3.0, C)]
    w = x[:C * L].view(C, L)
    qw = Q.WeightsLUTSymmetricInferableQuantizer(4, lut, thr, True, 0, 2)
    yw = qw(w)
    thr32 = np.asarray(thr, np.float64).astype(np.float32)
    for c in (0, 1023, 1024, 2047, 2048):
        for lo in (0, L - 4096):
            want = oracle.fq_lut(_bits(w[c, lo:lo + 4096]), oracle.BF16, np.asarray(lut, np.float32), thr32[c:c + 1], 1, 4096, 8, True, 1e-8)
            assert G.bits_equal(yw[c, lo:lo + 4096].cpu().numpy(), np.asarray(want).reshape(-1)), (c, lo)
    del yw, w, x
    torch.cuda.empty_cache()
