"""Producer-fused activation holders: the fx pass (CPU, structure only) and, on the GPU, bit-exact
agreement of every fused flavour with the unfused pair, the oracle, and CUDA-graph replay of a quantized block."""
import numpy as np
import pytest
import torch

import mct_quantizers_b200 as mctq
from mct_quantizers_b200.pytorch import quantizers as Q
from mct_quantizers_b200.pytorch.fused_activation_holder import PytorchFusedActivationQuantizationHolder

HAS_CUDA = torch.cuda.is_available()


class Block(torch.nn.Module):
    """conv -> relu -> holder; residual add -> holder; add -> relu -> holder; relu6 -> holder; plus sites that must NOT fuse."""

    def __init__(self, wrap=True):
        super().__init__()
        conv = torch.nn.Conv2d(4, 4, 3, padding=1)
        thr = [float(v) for v in conv.weight.detach().abs().flatten(1).amax(1)]
        # (tracing THROUGH a wrapper quantizes its weights, which needs the GPU: the CPU-only structure test uses a plain conv)
        self.conv = mctq.PytorchQuantizationWrapper(conv, {'weight': Q.WeightsSymmetricInferableQuantizer(8, thr, True, 0)}) if wrap else conv
        self.relu = torch.nn.ReLU()
        self.relu6 = torch.nn.ReLU6()
        self.h1 = mctq.PytorchActivationQuantizationHolder(Q.ActivationUniformInferableQuantizer(8, [0.0], [6.0]))
        self.h2 = mctq.PytorchActivationQuantizationHolder(Q.ActivationSymmetricInferableQuantizer(8, [4.0], True))
        self.h3 = mctq.PytorchActivationQuantizationHolder(Q.ActivationPOTInferableQuantizer(8, [8.0], False))
        self.h4 = mctq.PytorchActivationQuantizationHolder(Q.ActivationUniformInferableQuantizer(4, [-1.0], [5.0]))
        self.h5 = mctq.PytorchActivationQuantizationHolder(Q.ActivationLutPOTInferableQuantizer(
            4, [-8.0, -3.0, 0.0, 1.0, 5.0, 7.0], [4.0], True))
        self.h6 = mctq.PytorchActivationQuantizationHolder(Q.ActivationSymmetricInferableQuantizer(8, [2.0], True))

    def forward(self, x):
        a = self.h1(self.relu(self.conv(x)))           # fuses: relu
        b = self.h2(a + x)                             # fuses: add
        c = self.h3(torch.relu(b + a))                 # fuses: add_relu
        d = self.h4(self.relu6(c))                     # fuses: relu6
        e = self.h5(torch.relu(d))                     # LUT quantizer: stays
        r = torch.relu(e)
        f = self.h6(r)                                 # relu has a second consumer: stays
        return f + r


def test_fx_pass_structure():
    torch.manual_seed(0)
    gm = mctq.fuse_activation_producers(Block())
    assert gm.mctq_fused_sites == 4
    fused = {n: m.pre_op for n, m in gm.named_modules() if isinstance(m, PytorchFusedActivationQuantizationHolder)}
    assert sorted(fused.values()) == ["add", "add_relu", "relu", "relu6"]
    targets = [str(n.target) for n in gm.graph.nodes if n.op == "call_module"]
    assert "h5" in targets and "h6" in targets and "h1" not in targets
    # the traced-through spelling (default symbolic_trace leaves torch.ops.mctq.fq_affine_scalar nodes)
    gm2 = mctq.fuse_activation_producers(torch.fx.symbolic_trace(Block(wrap=False)))
    assert gm2.mctq_fused_sites == 4
    assert sum(1 for n in gm2.graph.nodes if n.op == "call_function" and "fq_affine_scalar_pre" in str(n.target)) == 4


class InplaceBlock(torch.nn.Module):
    """In-place relus: fusable only when nobody else can observe the mutated input."""

    def __init__(self):
        super().__init__()
        self.lin = torch.nn.Linear(8, 8)
        self.relu_ = torch.nn.ReLU(inplace=True)
        self.h1 = mctq.PytorchActivationQuantizationHolder(Q.ActivationUniformInferableQuantizer(8, [0.0], [6.0]))
        self.h2 = mctq.PytorchActivationQuantizationHolder(Q.ActivationUniformInferableQuantizer(8, [0.0], [6.0]))
        self.h3 = mctq.PytorchActivationQuantizationHolder(Q.ActivationUniformInferableQuantizer(8, [0.0], [6.0]))
        self.h4 = mctq.PytorchActivationQuantizationHolder(Q.ActivationUniformInferableQuantizer(8, [0.0], [6.0]))

    def forward(self, x):
        a = self.h1(self.relu_(self.lin(x)))                           # private intermediate, in place: fuses
        t = self.lin(a)
        b = self.h2(torch.nn.functional.relu(t, inplace=True))         # t is read again below (after its mutation): stays
        c = self.h3(torch.nn.functional.relu(self.lin(b), inplace=True))   # private intermediate: fuses
        d = self.h4(self.relu_(x))                                     # would mutate a graph input: stays
        return a + b + c + d + t


def test_fx_pass_leaves_observable_inplace_relus_alone():
    torch.manual_seed(0)
    gm = mctq.fuse_activation_producers(InplaceBlock())
    targets = [str(n.target) for n in gm.graph.nodes if n.op == "call_module"]
    assert gm.mctq_fused_sites == 2
    assert "h2" in targets and "h4" in targets and "h1" not in targets and "h3" not in targets


def test_fused_holder_argument_errors():
    with pytest.raises(ValueError):
        PytorchFusedActivationQuantizationHolder(Q.ActivationSymmetricInferableQuantizer(8, [4.0], True), "gelu")
    with pytest.raises(TypeError):
        PytorchFusedActivationQuantizationHolder(Q.ActivationLutPOTInferableQuantizer(4, [-8.0, 0.0, 7.0], [4.0], True), "relu")


def _bits(t):
    t = t.detach().cpu().contiguous()
    return t.view(torch.int32).numpy() if t.dtype == torch.float32 else t.view(torch.int16).numpy()


def _inputs(rng, n, dtype, dev):
    v = (rng.standard_normal(n) * 3).astype(np.float32)
    if n > 16:
        v[:8] = [np.nan, np.inf, -np.inf, -0.0, 0.0, 6.0, 1e30, -1e30]
    return torch.from_numpy(v).to(dtype).to(dev)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("pre_op", ["relu", "relu6", "add", "add_relu"])
def test_fused_equals_unfused_pair(dtype, pre_op):
    import oracle
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(5)
    quantizers = [Q.ActivationUniformInferableQuantizer(8, [-1.0], [5.0]), Q.ActivationSymmetricInferableQuantizer(8, [3.7], True),
                  Q.ActivationPOTInferableQuantizer(4, [4.0], False)]
    for q in quantizers:
        fused = PytorchFusedActivationQuantizationHolder(q, pre_op)
        for n in (1, 7, 8, 1023, 4096, 8193, 100003):
            x, o = _inputs(rng, n, dtype, dev), _inputs(rng, n, dtype, dev).flip(0).contiguous()
            got = fused(x, o) if "add" in pre_op else fused(x)
            want = fused._unfused(x, o)
            assert got.dtype == dtype and got.shape == x.shape
            assert np.array_equal(_bits(got), _bits(want)), (type(q).__name__, n)
            # and against the CPU oracle fed with the eager producer's output
            t = x + o if "add" in pre_op else x
            t = torch.relu(t) if pre_op in ("relu", "add_relu") else (torch.nn.functional.relu6(t) if pre_op == "relu6" else t)
            scale, zp, qmin, qmax = mctq.pytorch.fused_activation_holder.affine_scalar_params(q)
            tag = {torch.float32: oracle.F32, torch.bfloat16: oracle.BF16, torch.float16: oracle.F16}[dtype]
            tb = t.cpu().numpy() if dtype == torch.float32 else t.cpu().view(torch.int16).numpy().view(np.uint16)
            ref = oracle.fq_affine(tb, tag, np.array([scale], np.float64).astype(np.float32), np.array([zp], np.int32), 1, 1, qmin, qmax)
            finite = torch.isfinite(t).cpu().numpy()
            assert np.array_equal(_bits(got).view(ref.dtype)[finite], ref[finite])
        # misaligned views take the element-per-thread kernel
        x, o = _inputs(rng, 5001, dtype, dev)[1:], _inputs(rng, 5003, dtype, dev)[3:]
        got = fused(x, o) if "add" in pre_op else fused(x)
        assert np.array_equal(_bits(got), _bits(fused._unfused(x, o)))
        # a broadcasting add is not fusable: falls back to the pair
        if "add" in pre_op:
            xb = _inputs(rng, 12, dtype, dev).reshape(3, 4)
            ob = _inputs(rng, 4, dtype, dev)
            assert np.array_equal(_bits(fused(xb, ob)), _bits(fused._unfused(xb, ob)))


@pytest.mark.gpu
def test_fx_pass_preserves_results_and_saves_launches():
    from mct_quantizers_b200 import _native
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = Block().to(dev).eval()
    x = torch.randn(8, 4, 33, 35, device=dev)
    lib = _native.load()
    with torch.no_grad():
        want = model(x)
        c0 = lib.mctq_launch_count()
        model(x)
        unfused_launches = lib.mctq_launch_count() - c0
        for gm in (mctq.fuse_activation_producers(model), mctq.fuse_activation_producers(torch.fx.symbolic_trace(model))):
            got = gm(x)
            assert torch.equal(got.view(torch.int32), want.view(torch.int32))
            c0 = lib.mctq_launch_count()
            gm(x)
            # same number of OUR launches (one fewer when tracing through the wrapper folded the weight quantization) ...
            assert lib.mctq_launch_count() - c0 in (unfused_launches, unfused_launches - 1)
        # ... but four eager producer kernels (relu, add, add + relu, relu6) are gone
        from torch.profiler import profile, ProfilerActivity
        def cuda_kernels(fn):
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                fn(x)
                torch.cuda.synchronize()
            return sum(e.count for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA)
        gm = mctq.fuse_activation_producers(model)
        assert cuda_kernels(gm) <= cuda_kernels(model) - 5


@pytest.mark.gpu
@pytest.mark.parametrize("fuse", [False, True])
def test_cuda_graph_capture_of_a_quantized_block(fuse):
    """Wrappers + holders are capture-safe (no sync, no host round trip): replay == eager, also for new inputs."""
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    model = Block().to(dev).eval()
    if fuse:
        model = mctq.fuse_activation_producers(model)
    static_x = torch.randn(4, 4, 17, 19, device=dev)
    with torch.no_grad():
        for _ in range(3):                       # warm-up on a side stream (prepared parameter blobs are built here)
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                model(static_x)
            torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            static_y = model(static_x)
        for seed in (2, 3):
            xn = torch.randn(4, 4, 17, 19, device=dev, generator=torch.Generator(device=dev).manual_seed(seed))
            static_x.copy_(xn)
            g.replay()
            want = model(xn)
            assert torch.equal(static_y.view(torch.int32), want.view(torch.int32))


@pytest.mark.gpu
@pytest.mark.parametrize("backend", ["eager", "aot_eager"])
def test_torch_compile_sees_the_custom_ops(backend):
    """torch.compile (dynamo capture + fake-tensor propagation through the registered fake kernels) of a model with
    wrappers, holders and fused holders gives the eager result; the ops stay opaque calls into the CUDA library."""
    dev = torch.device("cuda:0")
    torch.manual_seed(2)
    model = Block().to(dev).eval()
    fused = mctq.fuse_activation_producers(model)
    x = torch.randn(4, 4, 17, 19, device=dev)
    with torch.no_grad():
        want = model(x)
        for m in (model, fused):
            compiled = torch.compile(m, backend=backend)
            got = compiled(x)
            assert torch.equal(got.view(torch.int32), want.view(torch.int32))
