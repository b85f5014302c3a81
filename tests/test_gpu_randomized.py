"""Seeded randomized differential test: every quantizer class through the PUBLIC API on the GPU against the CPU-torch
port of the reference's call sites (oracle/torch_cpu_port.py: the ATen ops / eager composition the reference runs),
over random shapes, channel axes, dtypes, bit widths, parameters and memory layouts.  Bit-exact."""
import numpy as np
import pytest
import torch

from oracle import torch_cpu_port as port

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
DTYPES = [torch.float32, torch.bfloat16, torch.float16]


@pytest.fixture(scope="module")
def Q():
    from mct_quantizers_b200.pytorch import quantizers
    return quantizers


def _bits(t):
    t = t.detach().cpu().contiguous()
    return t.view(torch.int32) if t.dtype == torch.float32 else t.view(torch.int16)


def _same(a, b):
    return a.dtype == b.dtype and a.shape == b.shape and torch.equal(_bits(a), _bits(b))


def _shape(rng):
    rank = int(rng.integers(1, 5))
    dims = [int(rng.integers(1, 9)) for _ in range(rank)]
    dims[int(rng.integers(0, rank))] = int(rng.choice([1, 3, 4, 8, 13, 16, 33, 64, 257]))
    return tuple(dims)


def _tensor(rng, shape, dtype, scale):
    x = torch.from_numpy((rng.standard_normal(shape) * scale).astype(np.float32)).to(dtype)
    return x


def _host_variants(x):
    """The same values as HOST tensors: pageable (staged through the GPU in chunks) and pinned (small pinned tensors are
    read / written by the kernel directly over PCIe)."""
    return [x.clone(), x.clone().pin_memory()]


def _layouts(x, rng):
    """x itself, a permuted-but-dense view, and a strided (non-dense) view holding the same values."""
    out = [x]
    if x.dim() >= 2:
        perm = list(rng.permutation(x.dim()))
        inv = [perm.index(i) for i in range(x.dim())]
        out.append(x.permute(perm).contiguous().permute(inv))            # same values, permuted strides
        big = torch.zeros(tuple(2 * s for s in x.shape), dtype=x.dtype)
        view = big[tuple(slice(0, 2 * s, 2) for s in x.shape)]
        view.copy_(x)
        out.append(view)                                                 # every stride doubled: not dense
    return out


@pytest.mark.parametrize("seed", range(12))
def test_weights_symmetric_and_pot(seed, Q):
    rng = np.random.default_rng(1000 + seed)
    for _ in range(6):
        shape = _shape(rng)
        dtype = DTYPES[int(rng.integers(0, 3))]
        bits = int(rng.integers(2, 9))
        per_channel = bool(rng.integers(0, 2))
        axis = int(rng.integers(0, len(shape)))
        C = shape[axis] if per_channel else 1
        pot = bool(rng.integers(0, 2))
        thr = [float(2.0 ** rng.integers(-4, 4)) for _ in range(C)] if pot else [float(v) for v in rng.uniform(0.05, 6.0, size=C)]
        cls = Q.WeightsPOTInferableQuantizer if pot else Q.WeightsSymmetricInferableQuantizer
        q = cls(bits, thr, per_channel, axis if per_channel else None)
        s, z, qmin, qmax = port.weights_symmetric_qparams(thr, bits)
        x = _tensor(rng, shape, dtype, 2.0)
        want = port.affine_per_channel(x, s, z, axis, qmin, qmax) if per_channel else port.affine_tensor_qparams(x, s, z, qmin, qmax)
        for xl in _layouts(x, rng):
            got = q(xl.to(DEV))
            assert _same(got, want), (shape, dtype, bits, per_channel, axis, pot, xl.stride())
        for xh in _host_variants(x):
            got = q(xh)
            assert not got.is_cuda and _same(got, want), ("host", shape, dtype, bits, per_channel, axis, pot)


@pytest.mark.parametrize("seed", range(12))
def test_weights_uniform(seed, Q):
    rng = np.random.default_rng(2000 + seed)
    for _ in range(6):
        shape = _shape(rng)
        dtype = DTYPES[int(rng.integers(0, 3))]
        bits = int(rng.integers(2, 9))
        per_channel = bool(rng.integers(0, 2))
        axis = int(rng.integers(0, len(shape)))
        C = shape[axis] if per_channel else 1
        lo = rng.uniform(-3.0, 1.0, size=C)
        hi = lo + rng.uniform(0.1, 5.0, size=C)
        q = Q.WeightsUniformInferableQuantizer(bits, [float(v) for v in lo], [float(v) for v in hi], per_channel,
                                               axis if per_channel else None)
        _, _, s, z, qmin, qmax = port.weights_uniform_qparams(lo, hi, bits)
        x = _tensor(rng, shape, dtype, 2.0)
        want = port.affine_per_channel(x, s, z, axis, qmin, qmax) if per_channel else port.affine_tensor_qparams(x, s, z, qmin, qmax)
        for xl in _layouts(x, rng):
            got = q(xl.to(DEV))
            assert _same(got, want), (shape, dtype, bits, per_channel, axis, xl.stride())
        for xh in _host_variants(x):
            got = q(xh)
            assert not got.is_cuda and _same(got, want), ("host", shape, dtype, bits, per_channel, axis)


@pytest.mark.parametrize("seed", range(8))
def test_activation_affine(seed, Q):
    rng = np.random.default_rng(3000 + seed)
    for _ in range(8):
        shape = _shape(rng)
        dtype = DTYPES[int(rng.integers(0, 3))]
        bits = int(rng.integers(2, 9))
        kind = int(rng.integers(0, 3))
        x = _tensor(rng, shape, dtype, 3.0)
        if kind == 2:
            lo = float(rng.uniform(-3.0, 1.0))
            hi = lo + float(rng.uniform(0.1, 6.0))
            q = Q.ActivationUniformInferableQuantizer(bits, [lo], [hi])
            _, _, scale, zp, qmin, qmax = port.activation_uniform_qparams([lo], [hi], bits)
        else:
            signed = bool(rng.integers(0, 2))
            thr = float(2.0 ** rng.integers(-3, 4)) if kind == 1 else float(rng.uniform(0.1, 8.0))
            q = (Q.ActivationPOTInferableQuantizer if kind == 1 else Q.ActivationSymmetricInferableQuantizer)(bits, [thr], signed)
            scales, qmin, qmax = port.symmetric_qparams([thr], bits, signed)
            scale, zp = float(scales[0]), 0
        want = port.affine_scalar_qparams(x, scale, zp, qmin, qmax)
        for xl in _layouts(x, rng):
            assert _same(q(xl.to(DEV)), want), (shape, dtype, bits, kind, xl.stride())
        for xh in _host_variants(x):
            got = q(xh)
            assert not got.is_cuda and _same(got, want), ("host", shape, dtype, bits, kind)


@pytest.mark.parametrize("seed", range(8))
def test_lut_weights(seed, Q):
    rng = np.random.default_rng(4000 + seed)
    for _ in range(5):
        shape = _shape(rng)
        dtype = DTYPES[int(rng.integers(0, 3))]
        bits = int(rng.integers(2, 6))
        K = int(rng.integers(1, 2 ** bits + 1))
        lut = [float(v) for v in rng.choice(np.arange(-128, 128), size=K, replace=bool(rng.integers(0, 2)))]
        per_channel = bool(rng.integers(0, 2))
        axis = int(rng.integers(0, len(shape)))
        C = shape[axis] if per_channel else 1
        pot = bool(rng.integers(0, 2))
        thr = [float(2.0 ** rng.integers(-5, 3)) for _ in range(C)] if pot else [float(v) for v in rng.uniform(0.02, 4.0, size=C)]
        cls = Q.WeightsLUTPOTInferableQuantizer if pot else Q.WeightsLUTSymmetricInferableQuantizer
        q = cls(bits, lut, thr, per_channel, axis if per_channel else None, len(shape) if per_channel else None)
        x = _tensor(rng, shape, dtype, 1.0)
        want = port.lut_fake_quant(x, torch.tensor(lut, dtype=torch.float32), True,
                                   torch.from_numpy(np.asarray(thr, np.float64).astype(np.float32)), 8, 1e-8,
                                   per_channel=per_channel, channel_axis=axis, input_rank=len(shape))
        for xl in _layouts(x, rng):
            got = q(xl.to(DEV))
            assert _same(got, want), (shape, dtype, bits, K, per_channel, axis, pot, xl.stride())
        for xh in _host_variants(x):
            got = q(xh)
            assert not got.is_cuda and _same(got, want), ("host", shape, dtype, bits, K, per_channel, axis, pot)


@pytest.mark.parametrize("seed", range(6))
def test_lut_activations(seed, Q):
    rng = np.random.default_rng(5000 + seed)
    for _ in range(6):
        shape = _shape(rng)
        dtype = DTYPES[int(rng.integers(0, 3))]
        bits = int(rng.integers(2, 6))
        signed = bool(rng.integers(0, 2))
        K = int(rng.integers(1, 2 ** bits + 1))
        pool = np.arange(-128, 128) if signed else np.arange(0, 256)
        lut = [float(v) for v in rng.choice(pool, size=K, replace=False)]
        thr = float(2.0 ** rng.integers(-2, 4))              # >= 0.25: the regime where CPU and CUDA torch agree (SURVEY 8a hazard 4)
        q = Q.ActivationLutPOTInferableQuantizer(bits, lut, [thr], signed)
        x = _tensor(rng, shape, dtype, 2.0)
        want = port.lut_fake_quant(x, torch.tensor(lut, dtype=torch.float32), signed, thr, 8, 1e-8)
        for xl in _layouts(x, rng):
            got = q(xl.to(DEV))
            assert _same(got, want), (shape, dtype, bits, K, signed, thr, xl.stride())
        for xh in _host_variants(x):
            got = q(xh)
            assert not got.is_cuda and _same(got, want), ("host", shape, dtype, bits, K, signed, thr)


@pytest.mark.parametrize("seed", range(6))
def test_integer_codes_and_lut_indices(seed, Q):
    """Code / index emission through the torch.library operators against the C oracle: int8 and packed int4."""
    import oracle
    import mct_quantizers_b200  # noqa: F401  (registers torch.ops.mctq)
    from mct_quantizers_b200.pytorch.quantizer_utils import lut_search_table
    rng = np.random.default_rng(6000 + seed)
    tags = {torch.float32: oracle.F32, torch.bfloat16: oracle.BF16, torch.float16: oracle.F16}
    for _ in range(5):
        shape = _shape(rng)
        dtype = DTYPES[int(rng.integers(0, 3))]
        axis = int(rng.integers(0, len(shape)))
        C = shape[axis]
        inner = int(np.prod(shape[axis + 1:]))
        x = _tensor(rng, shape, dtype, 1.5)
        xb = x.numpy() if dtype == torch.float32 else x.view(torch.int16).numpy().view(np.uint16)
        n = x.numel()
        for bits in (8, 4):
            signed = bool(rng.integers(0, 2))
            qmin, qmax = (-(2 ** (bits - 1)), 2 ** (bits - 1) - 1) if signed else (0, 2 ** bits - 1)
            scale = (np.abs(rng.standard_normal(C)) * 0.05 + 0.01).astype(np.float32)
            zp = np.zeros(C, np.int32) if signed else rng.integers(qmin, qmax + 1, size=C).astype(np.int32)
            mode = 1 if bits == 8 else 2
            codes, y = torch.ops.mctq.quantize_affine_channel(x.to(DEV), torch.from_numpy(scale).to(DEV), torch.from_numpy(zp).to(DEV),
                                                               axis, qmin, qmax, mode, True)
            want_y, want_codes = oracle.fq_affine(xb.reshape(-1), tags[dtype], scale, zp, C, inner, qmin, qmax, want_codes=True)
            got_y = y.cpu().reshape(-1)
            got_y = got_y.numpy() if dtype == torch.float32 else got_y.view(torch.int16).numpy().view(np.uint16)
            assert np.array_equal(got_y.view(want_y.dtype), want_y)
            c = codes.cpu().numpy().reshape(-1)
            if mode == 1:
                g = c.view(np.int8).astype(np.int32) if signed else c.view(np.uint8).astype(np.int32)
            else:
                nib = np.stack([c & 0xF, c >> 4], 1).reshape(-1)[:n].astype(np.int32)
                g = np.where(nib >= 8, nib - 16, nib) if signed else nib
            assert np.array_equal(g, want_codes), (shape, dtype, axis, bits, signed)
            # dequantising the codes gives the fake-quantised values (f32)
            yd = torch.ops.mctq.dequantize_affine(codes, mode, signed, list(shape), torch.from_numpy(scale).to(DEV),
                                                  torch.from_numpy(zp).to(DEV), axis)
            ref = ((want_codes - zp[(np.arange(n) // inner) % C]).astype(np.float32) * scale[(np.arange(n) // inner) % C]).astype(np.float32)
            assert np.array_equal(yd.cpu().numpy().reshape(-1).view(np.uint32), ref.view(np.uint32))
        # LUT indices
        K = int(rng.integers(1, 17))
        lut = rng.choice(np.arange(-128, 128), size=K, replace=False).astype(np.float32)
        thr = rng.uniform(0.05, 3.0, size=C).astype(np.float32)
        table = lut_search_table(lut, 8, True)
        for mode in (1, 2):
            idx = torch.ops.mctq.lut_indices(x.to(DEV), table, K, torch.from_numpy(thr).to(DEV), True, axis, 1e-8, mode)
            _, want_idx = oracle.fq_lut(xb.reshape(-1), tags[dtype], lut, thr, C, inner, 8, True, 1e-8, want_idx=True)
            c = idx.cpu().numpy().reshape(-1)
            g = c.astype(np.int32) if mode == 1 else np.stack([c & 0xF, c >> 4], 1).reshape(-1)[:n].astype(np.int32)
            assert np.array_equal(g[:n], want_idx), (shape, dtype, axis, K, mode)
