#!/usr/bin/env python
"""Fixtures from the UNMODIFIED reference running ON A CUDA MACHINE (the B200 box; `gpurun -- python
tests/golden/make_golden_cuda_flavour.py`, reference from baseline/_ref):

    ->  tests/golden/golden_cuda_flavour.npz + .json      (written under gpurun_out/ on the box, copied here by hand)

Why: the reference derives its parameters and normalises LUT activations with `tensor / python_number`, which libtorch
evaluates as a true division on CPU and as a multiplication by the (double -> f32) reciprocal on CUDA.  So the reference's
numbers depend on the machine it runs on: scales of the uniform quantizers differ by one ulp for most ranges (zero points
by one in rare cases), and ActivationLutPOT resolves some exact rounding ties towards the other centroid.  The other
fixture files pin the CPU flavour (this package's default); this file pins the CUDA flavour, which
`mct_quantizers_b200.reference_arithmetic("cuda")` reproduces.
Stored per case: constructor arguments, the reference's derived parameters, a CUDA input (with ties planted) and the
reference's CUDA output.
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
warnings.filterwarnings("ignore")
import logging  # noqa: E402
logging.disable(logging.WARNING)
import mct_quantizers  # noqa: E402
from mct_quantizers import pytorch_quantizers as Q  # noqa: E402

assert torch.cuda.is_available(), "this generator pins what the reference computes when a GPU is visible"
assert mct_quantizers.__version__ == "1.6.0" and "baseline" in mct_quantizers.__file__
DT = {"float32": torch.float32, "bfloat16": torch.bfloat16, "float16": torch.float16}
rng = np.random.default_rng(20270117)
arrays, cases = {}, []


def store(t):
    t = t.detach().cpu().contiguous()
    if t.dtype in (torch.bfloat16, torch.float16):
        return t.view(torch.int16).numpy().view(np.uint16).copy()
    return t.numpy().copy()


def add(name, cls, kw, x, params):
    q = getattr(Q, cls)(**kw)
    y = q(x.clone())
    arrays[f"{name}/x"] = store(x)
    arrays[f"{name}/y"] = store(y)
    rec = {"name": name, "cls": cls, "args": kw, "shape": list(x.shape), "x_dtype": str(x.dtype).replace("torch.", ""),
           "y_dtype": str(y.dtype).replace("torch.", ""), "params": {}}
    for p in params:
        v = getattr(q, p)
        if isinstance(v, torch.Tensor):
            arrays[f"{name}/p/{p}"] = store(v.flatten())
            rec["params"][p] = "tensor"
        else:
            rec["params"][p] = float(v) if isinstance(v, float) else int(v)
    cases.append(rec)


def tie_input(span, n, dt):
    v = rng.normal(0, span * 0.6, size=n).astype(np.float32)
    k = n // 2
    v[:k] = (rng.integers(-300, 300, size=k).astype(np.float32) + 0.5) * np.float32(span / 128.0) * rng.choice([0.5, 1.0, 2.0], size=k).astype(np.float32)
    rng.shuffle(v)
    return torch.from_numpy(v).to("cuda").to(DT[dt])


# uniform quantizers: parameters derived through `/ (2 ** n_bits - 1)`
for i in range(24):
    bits = int(rng.integers(2, 9))
    C = int(rng.integers(1, 9))
    hi = [float(np.float32(rng.uniform(2.0 ** -5, 12.0))) for _ in range(C)]
    lo = [-h * float(rng.uniform(0.0, 1.0)) for h in hi]
    dt = ("float32", "bfloat16", "float16")[i % 3]
    x = tie_input(float(np.mean(hi)), C * 257, dt).reshape(C, 257)
    add(f"cw_uni_{i}", "WeightsUniformInferableQuantizer", dict(num_bits=bits, min_range=lo, max_range=hi, per_channel=True, channel_axis=0),
        x, ("scales", "zero_points"))
for i in range(12):
    bits = int(rng.integers(2, 9))
    hi = float(np.float32(rng.uniform(2.0 ** -5, 12.0)))
    lo = -hi * float(rng.uniform(0.0, 1.0))
    dt = ("float32", "bfloat16", "float16")[i % 3]
    add(f"ca_uni_{i}", "ActivationUniformInferableQuantizer", dict(num_bits=bits, min_range=[lo], max_range=[hi]),
        tie_input(hi, 3000, dt).reshape(3, 1, 1000), ("scale", "zero_point", "min_range", "max_range"))

# LUT activations: `tensor / (threshold + eps)` on every call; exact ties between two centroids
for i in range(30):
    bits = int(rng.integers(3, 7))
    bw = int(rng.integers(max(bits, 5), 11))
    signed = bool(i % 2 == 0)
    lo_v, hi_v = (-2 ** (bw - 1), 2 ** (bw - 1) - 1) if signed else (0, 2 ** bw - 1)
    lut = [float(v) for v in rng.integers(lo_v, hi_v + 1, size=int(rng.integers(4, 2 ** bits + 1)))]
    thr = float(2.0 ** int(rng.integers(-4, 4)))
    dt = ("float32", "float32", "bfloat16", "float16")[i % 4]
    # inputs whose normalised value is an integer or a half-integer of the centroid grid: ties whenever two centroids are
    # an even / odd distance apart
    g = rng.integers(lo_v - 3, hi_v + 4, size=6000).astype(np.float64) + rng.choice([0.0, 0.5], size=6000)
    xv = (g / 2.0 ** (bw - int(signed)) * thr).astype(np.float32)
    x = torch.from_numpy(xv).to("cuda").to(DT[dt]).reshape(6, 1000)
    add(f"ca_lut_{i}", "ActivationLutPOTInferableQuantizer",
        dict(num_bits=bits, lut_values=lut, threshold=[thr], signed=signed, lut_values_bitwidth=bw), x, ())

out = os.path.join(ROOT, "gpurun_out")
os.makedirs(out, exist_ok=True)
np.savez_compressed(os.path.join(out, "golden_cuda_flavour.npz"), **arrays)
with open(os.path.join(out, "golden_cuda_flavour.json"), "w") as f:
    json.dump({"reference_version": mct_quantizers.__version__, "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0),
               "cases": cases}, f, indent=1)
print(len(cases), "cases ->", out)
