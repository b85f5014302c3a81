"""Host-side robustness of the CUDA path: inference tensors, stale launch plans, in-place parameter edits, reuse and
fallbacks of the whole-model plans, wide ranges in the fused holder.  Every result is compared with the per-layer call
and / or the oracle."""
import numpy as np
import pytest
import torch

import golden_util as G
import oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def Q():
    from mct_quantizers_b200.pytorch import quantizers
    return quantizers


def _oracle_sym(w, thr, bits=8):
    C = w.shape[0]
    sc = (np.asarray(thr, np.float64) / 2 ** (bits - 1)).astype(np.float32)
    return oracle.fq_affine(G.from_torch(w), {torch.float32: oracle.F32, torch.bfloat16: oracle.BF16, torch.float16: oracle.F16}[w.dtype],
                            sc, np.zeros(C, np.int32), C, w[0].numel(), -2 ** (bits - 1), 2 ** (bits - 1) - 1)


def test_quantizers_built_and_run_under_inference_mode(Q):
    """Inference tensors carry no version counter: cache keys must not touch `_version` (a per-channel weight call used
    to raise 'Inference tensors do not track version counter')."""
    rng = np.random.default_rng(0)
    with torch.inference_mode():
        w = torch.from_numpy(rng.standard_normal((24, 40)).astype(np.float32)).to(DEV)
        thr = [float(v) for v in w.abs().amax(1)]
        q = Q.WeightsSymmetricInferableQuantizer(8, thr, True, 0)
        assert q.scales.is_inference()
        y = q(w)
        lut = [-8.0, -3.0, 0.0, 1.0, 5.0, 7.0]
        ql = Q.WeightsLUTSymmetricInferableQuantizer(4, lut, thr, True, 0, 2, 4)
        yl = ql(w)
        qa = Q.ActivationUniformInferableQuantizer(8, [-1.0], [2.0])
        ya = qa(w)
    assert G.bits_equal(y.cpu().numpy(), np.asarray(_oracle_sym(w, thr)).reshape(y.shape))
    want = oracle.fq_lut(w.cpu().numpy(), oracle.F32, np.asarray(lut, np.float32), np.asarray(thr, np.float64).astype(np.float32),
                         24, 40, 4, True, 1e-8)
    assert G.bits_equal(yl.cpu().numpy(), np.asarray(want).reshape(yl.shape))
    assert ya.shape == w.shape
    # parameters created under inference mode, used outside it, and (when there is a second GPU) on another device
    y2 = q(w.clone())
    assert torch.equal(y2, y)
    if torch.cuda.device_count() > 1:
        with torch.inference_mode():
            y3 = q(w.to("cuda:1"))
        assert torch.equal(y3.cpu(), y.cpu())


def test_weight_plan_follows_moved_and_converted_weights(Q):
    """WeightPlan / ModelWeightPlan capture pointers: after .half(), a re-assigned parameter or new storage the plan must
    rebuild instead of quantizing the stale storage."""
    import mct_quantizers_b200 as mctq
    from mct_quantizers_b200.pytorch.model_quantization import plan_model_weights
    torch.manual_seed(0)
    convs = [torch.nn.Conv2d(8, 16, 3), torch.nn.Conv2d(16, 8, 1)]
    wrappers = []
    for c in convs:
        thr = [float(v) for v in c.weight.detach().abs().flatten(1).amax(1)]
        wrappers.append(mctq.PytorchQuantizationWrapper(c, {'weight': Q.WeightsSymmetricInferableQuantizer(8, thr, True, 0)}))
    model = torch.nn.Sequential(*wrappers).to(DEV)
    plan = plan_model_weights(model)
    plan.enable()

    def installed():
        return [w.layer.weight.detach().clone() for w in wrappers]

    def per_layer():
        return [w.get_quantized_weights()['weight'] for w in wrappers]

    for got, want in zip(installed(), per_layer()):
        assert torch.equal(got, want)
    # 1. new values in new storage (what load_state_dict / an optimizer step into fresh tensors does)
    with torch.no_grad():
        for w in wrappers:
            for _, p, _ in w.get_weights_vars():
                p.data = (p.data * 0.5).clone()
    plan.refresh()
    for got, want in zip(installed(), per_layer()):
        assert torch.equal(got, want)
    # 2. dtype conversion of the whole model
    model.half()
    plan.refresh()
    for got, want, w in zip(installed(), per_layer(), wrappers):
        assert got.dtype == torch.float16 and torch.equal(got, want)
        p = w.get_weights_vars()[0][1]
        thr = [float(v) for v in w.get_weights_vars()[0][2].threshold_np]
        assert G.bits_equal(G.from_torch(got).reshape(-1), np.asarray(_oracle_sym(p.detach(), thr)).reshape(-1))
    plan.disable()


def test_weight_plan_fallbacks_and_reuse(Q):
    from mct_quantizers_b200.pytorch.model_quantization import WeightPlan
    rng = np.random.default_rng(1)
    w_dense = torch.from_numpy(rng.standard_normal((16, 32)).astype(np.float32)).to(DEV)
    w_view = torch.from_numpy(rng.standard_normal((16, 64)).astype(np.float32)).to(DEV)[:, ::2]      # not dense
    w_reuse = torch.from_numpy(rng.standard_normal((16, 32)).astype(np.float32)).to(DEV)
    thr = lambda w: [float(v) for v in w.abs().amax(1)]  # noqa: E731
    q1 = Q.WeightsSymmetricInferableQuantizer(8, thr(w_dense), True, 0)
    q2 = Q.WeightsSymmetricInferableQuantizer(8, thr(w_view), True, 0)
    q3 = Q.WeightsSymmetricInferableQuantizer(8, thr(w_reuse), True, 0)
    q3.enable_reuse_quantizer()
    plan = WeightPlan([("a", w_dense, q1), ("b", w_view, q2), ("c", w_reuse, q3)])
    assert plan.plan is not None and plan.plan.n_desc == 1 and len(plan.other) == 2      # view and reuse -> per-quantizer calls
    y = plan.run()
    assert torch.equal(y[0], q1(w_dense)) and torch.equal(y[1], q2(w_view.contiguous()).reshape(y[1].shape))
    first = y[2]
    w_reuse.mul_(0.25)
    again = plan.run()[2]
    assert again is first                                  # the reuse contract: later calls return the first output object
    q3.disable_reuse_quantizer()
    fresh = plan.run()[2]                                  # flag change -> plan rebuilt, tensor back in the fused launch
    assert not torch.equal(fresh, first) and torch.equal(fresh, q3(w_reuse))
    if torch.cuda.device_count() > 1:
        w_far = w_dense.to("cuda:1")
        plan2 = WeightPlan([("a", w_dense, q1), ("far", w_far, q1)])
        ya, yf = plan2.run()
        assert yf.device == w_far.device and torch.equal(ya.cpu(), yf.cpu())


def test_inplace_parameter_edits_are_noticed(Q):
    rng = np.random.default_rng(2)
    w = torch.from_numpy((rng.standard_normal((12, 264)) * 0.05).astype(np.float32)).to(DEV)
    thr = [float(v) for v in w.abs().amax(1)]
    lut = [float(v) for v in sorted(rng.choice(np.arange(-128, 128), size=16, replace=False))]
    ql = Q.WeightsLUTSymmetricInferableQuantizer(4, lut, thr, True, 0, 2)
    qs = Q.WeightsSymmetricInferableQuantizer(8, thr, True, 0)
    y0, s0 = ql(w), qs(w)
    ql._threshold_torch.mul_(2.0)                           # in place: same data_ptr, new version
    qs.scales.mul_(2.0)
    y1, s1 = ql(w), qs(w)
    thr2 = (np.asarray(thr, np.float64).astype(np.float32) * np.float32(2.0))
    want = oracle.fq_lut(w.cpu().numpy(), oracle.F32, np.asarray(lut, np.float32), thr2, 12, 264, 8, True, 1e-8)
    assert G.bits_equal(y1.cpu().numpy(), np.asarray(want).reshape(y1.shape))
    assert not torch.equal(y0, y1)
    sc = (np.asarray(thr, np.float64) / 128).astype(np.float32) * np.float32(2.0)
    want_s = oracle.fq_affine(w.cpu().numpy(), oracle.F32, sc, np.zeros(12, np.int32), 12, 264, -128, 127)
    assert G.bits_equal(s1.cpu().numpy(), np.asarray(want_s).reshape(s1.shape)) and not torch.equal(s0, s1)


def test_fused_holder_wide_range_falls_back(Q):
    import mct_quantizers_b200 as mctq
    q = Q.ActivationSymmetricInferableQuantizer(23, [4.0], True)           # 2^23 codes: outside the fused kernel's fast rounding
    fh = mctq.PytorchFusedActivationQuantizationHolder(q, "relu")
    x = torch.randn(5000, device=DEV) * 3
    y = fh(x)
    assert torch.equal(y, q(torch.relu(x)))
