#!/usr/bin/env python
"""Third batch of fixtures from the UNMODIFIED reference (build container, CPU torch):

    python tests/golden/make_golden_v3.py        ->  golden_v3.npz + golden_v3.json

LUT quantizers on grids of MORE than 10 bits (`lut_values_bitwidth` = 12 / 14 / 16): since round 2 the prepared CUDA path
covers them with a cell table coarser than the integer grid, so they need reference outputs of their own -- sparse
centroid lists (what k-means on such a grid produces), lists with two neighbouring integers (too dense for the coarse
table: the generic kernel must take over), weights (per-channel / per-tensor, symmetric and power-of-two thresholds) and
activations (signed / unsigned, half-precision inputs round every eager op to the input dtype).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_common import torch, manifest, add_case, build_tensor, lut_channel_input, save  # noqa: E402

rng = np.random.default_rng(20261217)
manifest["cases"] = []


def narrow(x, dt):
    return x.to({"float32": torch.float32, "bfloat16": torch.bfloat16, "float16": torch.float16}[dt])


def thresholds(C, pot=False):
    if pot:
        return [float(2.0 ** e) for e in rng.integers(-3, 3, size=C)]
    return [float(np.float32(v)) for v in rng.uniform(0.05, 6.0, size=C)]


def sparse_lut(bw, k, signed=True):
    lo, hi = (-2 ** (bw - 1), 2 ** (bw - 1) - 1) if signed else (0, 2 ** bw - 1)
    step = (hi - lo) // (k + 1)
    base = lo + step // 2 + step * np.arange(k)
    v = base + rng.integers(-step // 4, step // 4 + 1, size=k)
    v[rng.integers(0, k)] = 0
    return [float(x) for x in sorted(set(int(t) for t in v))]


LUTS = {12: sparse_lut(12, 16), 14: sparse_lut(14, 16), 16: sparse_lut(16, 16)}
DENSE12 = [float(v) for v in (-2048, -3, -2, -1, 0, 1, 900, 2047)]            # five neighbouring integers

for bw, lut in LUTS.items():
    for dt in ("float32", "bfloat16", "float16"):
        for shape in ((6, 328), (5, 44)):
            C = shape[0]
            thr = thresholds(C)
            vecs = [lut_channel_input(rng, lut, thr[c], bw, True, shape[1]) for c in range(C)]
            add_case(f"wl_sym_bw{bw}_pc_{shape[0]}x{shape[1]}_{dt}", "WeightsLUTSymmetricInferableQuantizer",
                     dict(num_bits=4, lut_values=lut, threshold=thr, per_channel=True, channel_axis=0, input_rank=2, lut_values_bitwidth=bw),
                     narrow(torch.from_numpy(build_tensor(vecs, shape, 0)), dt),
                     lut_info=dict(threshold=torch.tensor(thr, dtype=torch.float32).reshape(-1, 1), bw=bw, signed=True, eps=1e-8))
    thr = thresholds(4, pot=True)
    vecs = [lut_channel_input(rng, lut, thr[c], bw, True, 300) for c in range(4)]
    add_case(f"wl_pot_bw{bw}_pc_4x300", "WeightsLUTPOTInferableQuantizer",
             dict(num_bits=4, lut_values=lut, threshold=thr, per_channel=True, channel_axis=0, input_rank=2, lut_values_bitwidth=bw),
             torch.from_numpy(build_tensor(vecs, (4, 300), 0)),
             lut_info=dict(threshold=torch.tensor(thr, dtype=torch.float32).reshape(-1, 1), bw=bw, signed=True, eps=1e-8))
    t1 = thresholds(1)
    v = lut_channel_input(rng, lut, t1[0], bw, True, 3000)
    add_case(f"wl_sym_bw{bw}_pt_3000", "WeightsLUTSymmetricInferableQuantizer",
             dict(num_bits=4, lut_values=lut, threshold=t1, per_channel=False, lut_values_bitwidth=bw),
             torch.from_numpy(v.reshape(30, 100)), lut_info=dict(threshold=torch.tensor(t1, dtype=torch.float32), bw=bw, signed=True, eps=1e-8))

# dense lists (neighbouring integers): a 12-bit grid still gets one cell per step; on a 14-bit grid the list is outside the
# coarse cell table and the generic kernel takes over
thr = thresholds(5)
for bw in (12, 14):
    vecs = [lut_channel_input(rng, DENSE12, thr[c], bw, True, 400) for c in range(5)]
    for dt in ("float32", "bfloat16"):
        add_case(f"wl_sym_bw{bw}_dense_pc_5x400_{dt}", "WeightsLUTSymmetricInferableQuantizer",
                 dict(num_bits=3, lut_values=DENSE12, threshold=thr, per_channel=True, channel_axis=0, input_rank=2, lut_values_bitwidth=bw),
                 narrow(torch.from_numpy(build_tensor(vecs, (5, 400), 0)), dt),
                 lut_info=dict(threshold=torch.tensor(thr, dtype=torch.float32).reshape(-1, 1), bw=bw, signed=True, eps=1e-8))

# activations (power-of-two thresholds), signed and unsigned, 12- and 16-bit grids
for bw in (12, 16):
    for signed in (True, False):
        lut = sparse_lut(bw, 8, signed)
        for thr in (4.0, 0.25):
            v = lut_channel_input(rng, lut, thr, bw, signed, 4000)
            for dt in ("float32", "bfloat16", "float16"):
                if dt == "float16" and bw == 16 and not signed:
                    continue            # the reference itself raises here: the clip bound 65535 overflows float16
                add_case(f"al_{'s' if signed else 'u'}_lut8_bw{bw}_t{thr}_{dt}", "ActivationLutPOTInferableQuantizer",
                         dict(num_bits=3, lut_values=lut, threshold=[thr], signed=signed, lut_values_bitwidth=bw),
                         narrow(torch.from_numpy(v.reshape(4, 1000)), dt), lut_info=dict(threshold=thr, bw=bw, signed=signed, eps=1e-8))

save("golden_v3")
