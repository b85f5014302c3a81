#!/usr/bin/env python
"""Generates tests/golden/ref_pickles/: modules pickled BY THE UNMODIFIED REFERENCE (torch.save of whole modules, class
paths `mct_quantizers.…`) plus the inputs / outputs the reference computes for them on CPU torch.

    cd /tmp && PYTHONPATH=/root/reference python /root/repo/tests/golden/make_ref_pickles.py

tests/test_compat_alias.py loads them through `mct_quantizers_b200.compat` (import alias) and checks on the GPU that the
B200 objects built from the reference's pickled state give the reference's results."""
import os
import sys

import numpy as np
import torch

import mct_quantizers as ref                      # the reference (PYTHONPATH=/root/reference)
from mct_quantizers.pytorch import quantizers as RQ

assert "b200" not in ref.__file__, "run with PYTHONPATH=/root/reference, without the compat alias"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_pickles")
os.makedirs(OUT, exist_ok=True)
torch.manual_seed(0)
rng = np.random.default_rng(0)
arrays = {}


def boundary_dense(shape, scale):
    x = rng.standard_normal(shape).astype(np.float32) * scale
    flat = x.reshape(-1)
    flat[:8] = [0.0, -0.0, 1e-9, -1e-9, 100.0, -100.0, 0.5, -0.5]
    return torch.from_numpy(x)


LUT = [-128.0, -77.0, -30.0, -9.0, 0.0, 4.0, 21.0, 64.0, 127.0]
holders = {
    "holder_act_symmetric": ref.PytorchActivationQuantizationHolder(RQ.ActivationSymmetricInferableQuantizer(8, [3.7], True)),
    "holder_act_pot_unsigned": ref.PytorchActivationQuantizationHolder(RQ.ActivationPOTInferableQuantizer(8, [2.0], False)),
    "holder_act_uniform": ref.PytorchActivationQuantizationHolder(RQ.ActivationUniformInferableQuantizer(8, [-1.0], [2.3])),
    "holder_act_lut_pot": ref.PytorchActivationQuantizationHolder(RQ.ActivationLutPOTInferableQuantizer(4, LUT, [2.0], True)),
    "holder_fln_bypass_off": ref.PytorchFLNActivationQuantizationHolder(RQ.ActivationSymmetricInferableQuantizer(7, [4.0], True), quantization_bypass=False),
    "holder_preserving_bypass_on": ref.PytorchPreservingActivationQuantizationHolder(RQ.ActivationSymmetricInferableQuantizer(7, [4.0], True), quantization_bypass=True),
}
for name, h in holders.items():
    x = boundary_dense((2, 3, 17, 19), 2.0)
    y = h(x)
    arrays[name + "/x"], arrays[name + "/y"] = x.numpy(), y.detach().numpy()
    torch.save(h, os.path.join(OUT, name + ".pt"))


def thr_of(w, axis=0):
    return [float(v) for v in w.detach().abs().transpose(0, axis).flatten(1).amax(1)]


conv = torch.nn.Conv2d(3, 8, 3)
lin = torch.nn.Linear(16, 8)
convt = torch.nn.ConvTranspose2d(4, 6, 2)
wrappers = {
    "wrapper_conv_w_symmetric_pc": (ref.PytorchQuantizationWrapper(conv, {'weight': RQ.WeightsSymmetricInferableQuantizer(8, thr_of(conv.weight), True, 0)}),
                                    (2, 3, 9, 9)),
    "wrapper_linear_w_uniform_pc": (ref.PytorchQuantizationWrapper(lin, {'weight': RQ.WeightsUniformInferableQuantizer(
        8, [float(v) for v in lin.weight.detach().amin(1) - 0.01], [float(v) for v in lin.weight.detach().amax(1) + 0.01], True, 0)}), (5, 16)),
    "wrapper_convT_w_pot_pc_axis1": (ref.PytorchQuantizationWrapper(convt, {'weight': RQ.WeightsPOTInferableQuantizer(
        8, [float(2.0 ** np.ceil(np.log2(t))) for t in thr_of(convt.weight, 1)], True, 1)}), (2, 4, 5, 5)),
    "wrapper_linear_w_lut_sym_pc": (ref.PytorchQuantizationWrapper(torch.nn.Linear(16, 8), {'weight': RQ.WeightsLUTSymmetricInferableQuantizer(
        4, LUT, [0.31, 0.27, 0.4, 0.25, 0.33, 0.29, 0.5, 0.26], True, 0, 2)}), (5, 16)),
    "wrapper_conv_w_lut_pot_pt": (ref.PytorchQuantizationWrapper(torch.nn.Conv2d(3, 4, 1), {'weight': RQ.WeightsLUTPOTInferableQuantizer(
        4, LUT, [1.0], False)}), (2, 3, 6, 6)),
}
for name, (w, in_shape) in wrappers.items():
    x = torch.from_numpy(rng.standard_normal(in_shape).astype(np.float32))
    y = w(x)
    qw = w.get_quantized_weights()
    arrays[name + "/x"], arrays[name + "/y"] = x.numpy(), y.detach().numpy()
    for k, v in qw.items():
        arrays[name + "/qw/" + k] = v.detach().numpy()
    torch.save(w, os.path.join(OUT, name + ".pt"))

np.savez_compressed(os.path.join(OUT, "expected.npz"), **arrays)
print("wrote", len(holders) + len(wrappers), "pickles and", len(arrays), "arrays to", OUT, "| reference", ref.__version__,
      "| torch", torch.__version__, "| python", sys.version.split()[0])
