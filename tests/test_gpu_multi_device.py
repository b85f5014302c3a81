"""Inputs on a GPU other than the one the quantizer was built on (the reference pins parameters to the construction-time
current device, `pytorch/quantizer_utils.py:31`, and fails for `cuda:k`, k != 0; here parameters follow the input).
Needs two GPUs: skipped on single-GPU boxes."""
import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")]


def _q_all():
    from mct_quantizers_b200.pytorch import quantizers as Q
    lut = [-128.0, -40.0, -7.0, 0.0, 3.0, 30.0, 127.0]
    C = 6
    thr = [0.5 + 0.1 * k for k in range(C)]
    return [
        ("w_sym_pc", Q.WeightsSymmetricInferableQuantizer(8, thr, True, 1), (4, C, 5)),
        ("w_pot_pt", Q.WeightsPOTInferableQuantizer(4, [2.0], False), (4, C, 5)),
        ("w_uni_pc", Q.WeightsUniformInferableQuantizer(8, [-0.4 - 0.1 * k for k in range(C)], [0.6 + 0.1 * k for k in range(C)], True, 1), (4, C, 5)),
        ("w_lut_pc", Q.WeightsLUTSymmetricInferableQuantizer(4, lut, thr, True, 1, 3), (4, C, 5)),
        ("w_lutpot_pt", Q.WeightsLUTPOTInferableQuantizer(4, lut, [1.0], False), (4, C, 5)),
        ("a_sym", Q.ActivationSymmetricInferableQuantizer(8, [3.7], True), (3, 7, 11)),
        ("a_pot", Q.ActivationPOTInferableQuantizer(8, [4.0], False), (3, 7, 11)),
        ("a_uni", Q.ActivationUniformInferableQuantizer(8, [-1.0], [2.3]), (3, 7, 11)),
        ("a_lutpot", Q.ActivationLutPOTInferableQuantizer(4, lut, [2.0], True), (3, 7, 11)),
    ]


def test_every_quantizer_follows_the_input_device():
    rng = np.random.default_rng(0)
    torch.cuda.set_device(0)
    for name, q, shape in _q_all():
        x = torch.from_numpy(rng.standard_normal(shape).astype(np.float32))
        y0 = q(x.to("cuda:0"))
        y1 = q(x.to("cuda:1"))
        assert y1.device == torch.device("cuda:1"), name
        assert torch.equal(y0.cpu().view(torch.int32), y1.cpu().view(torch.int32)), name
        assert torch.cuda.current_device() == 0, "the call must restore the current device"


def test_modules_moved_to_the_second_device_and_streams():
    import mct_quantizers_b200 as mctq
    from mct_quantizers_b200.pytorch import quantizers as Q
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(3, 8, 3)
    thr = [float(v) for v in conv.weight.detach().abs().flatten(1).amax(1)]
    w = mctq.PytorchQuantizationWrapper(conv, {'weight': Q.WeightsSymmetricInferableQuantizer(8, thr, True, 0)})
    h = mctq.PytorchActivationQuantizationHolder(Q.ActivationUniformInferableQuantizer(8, [-1.0], [2.3]))
    x = torch.randn(2, 3, 9, 9)
    want_w = w.to("cuda:0").get_quantized_weights()['weight'].cpu()
    want_h = h.to("cuda:0")(x.to("cuda:0")).cpu()
    w1, h1 = w.to("cuda:1"), h.to("cuda:1")
    side = torch.cuda.Stream(device="cuda:1")
    with torch.cuda.stream(side):                      # non-default stream of a non-current device
        got_w = w1.get_quantized_weights()['weight']
        got_h = h1(x.to("cuda:1"))
        out = w1(x.to("cuda:1"))
    side.synchronize()
    assert got_w.device.index == 1 and out.device.index == 1
    assert torch.equal(got_w.cpu().view(torch.int32), want_w.view(torch.int32))
    assert torch.equal(got_h.cpu().view(torch.int32), want_h.view(torch.int32))
