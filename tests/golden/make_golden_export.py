#!/usr/bin/env python
"""Fixtures for the ONNX-export branch, from the UNMODIFIED reference (build container, CPU torch):

    python tests/golden/make_golden_export.py        ->  golden_export.npz + golden_export.json

Every one of the nine quantizers is built from the reference, switched to its custom implementation
(`enable_custom_impl()`) and called under `torch.jit.trace` -- the only situation in which the reference runs its `*F`
autograd Functions (`quantizer._use_custom_impl and torch.jit.is_tracing()`, e.g.
weights_symmetric_inferable_quantizer.py:130-136 -> WeightsSymmetricF.forward -> quantize_sym_weights_torch :32-70).
Stored per case: constructor arguments, input, the traced function's output on that input, and the names of the
autograd Functions that appear in the traced graph.  These formulas use true division and differ from the inference
path at rounding ties, so the inputs contain the same tie neighbourhoods as the inference fixtures.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_common import Q, torch, affine_channel_input, build_tensor, lut_channel_input, store, HERE  # noqa: E402

rng = np.random.default_rng(20261101)
arrays, cases = {}, []


def python_ops(graph):
    return sorted({n.pyname() for n in graph.nodes() if n.kind() == "prim::PythonOp"})


def add(name, cls, args, x):
    q = getattr(Q, cls)(**args)
    q.enable_custom_impl()
    traced = torch.jit.trace(lambda t: q(t), x, check_trace=False)
    y = traced(x)
    eager = q(x)                                     # not tracing: the inference path (documented to differ at ties)
    arrays[f"{name}/x"] = store(x)
    arrays[f"{name}/y"] = store(y)
    cases.append({"name": name, "cls": cls, "args": args, "shape": list(x.shape), "x_dtype": str(x.dtype).replace("torch.", ""),
                  "y_dtype": str(y.dtype).replace("torch.", ""), "python_ops": python_ops(traced.graph),
                  "differs_from_inference_path": int((eager != y).sum())})


def weight_input(scales, zps, qmin, qmax, shape, axis):
    C = shape[axis]
    L = int(np.prod(shape)) // C
    return torch.from_numpy(build_tensor([affine_channel_input(rng, scales[c], int(zps[c]), qmin, qmax, L) for c in range(C)], shape, axis))


def thresholds(C, pot=False):
    if pot:
        return [float(2.0 ** int(e)) for e in rng.integers(-4, 3, size=C)]
    return [float(v) for v in np.abs(rng.normal(0, 1.5, size=C)) + 0.05]


# ---- symmetric / POT weights
for cls, pot in (("WeightsSymmetricInferableQuantizer", False), ("WeightsPOTInferableQuantizer", True)):
    for bits in (8, 4):
        for shape, axis in (((6, 4, 3, 3), 0), ((10, 12), 1)):
            thr = thresholds(shape[axis], pot)
            sc = np.asarray(thr, np.float64) / 2 ** (bits - 1)
            x = weight_input(sc.astype(np.float32), np.zeros(len(thr)), -2 ** (bits - 1), 2 ** (bits - 1) - 1, shape, axis)
            add(f"x_{'pot' if pot else 'sym'}_w_b{bits}_pc_{'x'.join(map(str, shape))}_ax{axis}", cls,
                dict(num_bits=bits, threshold=thr, per_channel=True, channel_axis=axis), x)
        thr = thresholds(1, pot)
        sc = np.asarray(thr, np.float64) / 2 ** (bits - 1)
        x = torch.from_numpy(affine_channel_input(rng, np.float32(sc[0]), 0, -2 ** (bits - 1), 2 ** (bits - 1) - 1, 3000).reshape(30, 100))
        add(f"x_{'pot' if pot else 'sym'}_w_b{bits}_pt", cls, dict(num_bits=bits, threshold=thr, per_channel=False), x)

# ---- uniform weights
for bits in (8, 3):
    for shape, axis in (((5, 6, 3, 3), 0), ((12, 9), 1)):
        C = shape[axis]
        lo = rng.normal(-1.0, 1.0, size=C)
        hi = lo + np.abs(rng.normal(0, 2.0, size=C)) + 0.1
        lo[0], hi[0] = 0.3, 2.1
        lo[1], hi[1] = -2.5, -0.4
        lo, hi = [float(v) for v in lo], [float(v) for v in hi]
        probe = Q.WeightsUniformInferableQuantizer(bits, lo, hi, True, axis)
        x = weight_input(probe.scales.numpy(), probe.zero_points.numpy(), 0, 2 ** bits - 1, shape, axis)
        add(f"x_uni_w_b{bits}_pc_{'x'.join(map(str, shape))}_ax{axis}", "WeightsUniformInferableQuantizer",
            dict(num_bits=bits, min_range=lo, max_range=hi, per_channel=True, channel_axis=axis), x)
    probe = Q.WeightsUniformInferableQuantizer(bits, [-0.73], [1.9], False)
    x = torch.from_numpy(affine_channel_input(rng, probe.scales.numpy()[0], int(probe.zero_points.numpy()[0]), 0, 2 ** bits - 1, 2500).reshape(50, 50))
    add(f"x_uni_w_b{bits}_pt", "WeightsUniformInferableQuantizer", dict(num_bits=bits, min_range=[-0.73], max_range=[1.9], per_channel=False), x)

# ---- activations
for cls, pot in (("ActivationSymmetricInferableQuantizer", False), ("ActivationPOTInferableQuantizer", True)):
    for bits in (8, 4):
        for signed in (True, False):
            thr = [4.0] if pot else [3.7]
            levels = 2 ** (bits - 1) if signed else 2 ** bits
            sc = np.float32(thr[0] / levels)
            qmin, qmax = (-levels, levels - 1) if signed else (0, levels - 1)
            x = torch.from_numpy(affine_channel_input(rng, sc, 0, qmin, qmax, 4000).reshape(4, 10, 100))
            add(f"x_{'pot' if pot else 'sym'}_a_b{bits}_{'s' if signed else 'u'}", cls, dict(num_bits=bits, threshold=thr, signed=signed), x)
for bits in (8, 4):
    for lo, hi in (([-1.0], [2.3]), ([0.4], [5.0]), ([-3.0], [-0.5])):
        probe = Q.ActivationUniformInferableQuantizer(bits, lo, hi)
        x = torch.from_numpy(affine_channel_input(rng, np.float32(probe.scale), int(probe.zero_point), 0, 2 ** bits - 1, 4000).reshape(8, 500))
        add(f"x_uni_a_b{bits}_{lo[0]}_{hi[0]}", "ActivationUniformInferableQuantizer", dict(num_bits=bits, min_range=lo, max_range=hi), x)

# ---- LUT quantizers
lut16 = [float(v) for v in sorted(rng.choice(np.arange(-128, 128), size=16, replace=False))]
lut5 = [-8.0, -3.0, 0.0, 2.0, 7.0]
for cls, pot in (("WeightsLUTSymmetricInferableQuantizer", False), ("WeightsLUTPOTInferableQuantizer", True)):
    for lut, nb, bw in ((lut16, 4, 8), (lut5, 3, 4)):
        for shape, axis in (((6, 40), 0), ((7, 5, 3, 3), 0), ((9, 6), 1)):
            thr = thresholds(shape[axis], pot)
            L = int(np.prod(shape)) // shape[axis]
            x = torch.from_numpy(build_tensor([lut_channel_input(rng, np.asarray(lut, np.float32), np.float32(t), bw, True, L) for t in thr], shape, axis))
            add(f"x_lut{'pot' if pot else 'sym'}_w_k{len(lut)}_pc_{'x'.join(map(str, shape))}_ax{axis}", cls,
                dict(num_bits=nb, lut_values=lut, threshold=thr, per_channel=True, channel_axis=axis, input_rank=len(shape), lut_values_bitwidth=bw), x)
        thr = thresholds(1, pot)
        x = torch.from_numpy(lut_channel_input(rng, np.asarray(lut, np.float32), np.float32(thr[0]), bw, True, 2000).reshape(40, 50))
        add(f"x_lut{'pot' if pot else 'sym'}_w_k{len(lut)}_pt", cls,
            dict(num_bits=nb, lut_values=lut, threshold=thr, per_channel=False, lut_values_bitwidth=bw), x)
for signed, lut in ((True, lut5), (False, [0.0, 1.0, 3.0, 6.0, 11.0, 15.0])):
    x = torch.from_numpy(lut_channel_input(rng, np.asarray(lut, np.float32), np.float32(4.0), 4, signed, 3000).reshape(6, 500))
    add(f"x_lutpot_a_{'s' if signed else 'u'}", "ActivationLutPOTInferableQuantizer",
        dict(num_bits=3, lut_values=lut, threshold=[4.0], signed=signed, lut_values_bitwidth=4), x)

np.savez_compressed(os.path.join(HERE, "golden_export.npz"), **arrays)
with open(os.path.join(HERE, "golden_export.json"), "w") as f:
    json.dump({"reference_version": "1.6.0", "torch": torch.__version__, "cases": cases}, f, indent=1)
print(len(cases), "cases;", sum(c["differs_from_inference_path"] > 0 for c in cases), "differ from the inference path somewhere")
print(sorted({op for c in cases for op in c["python_ops"]}))
