#!/usr/bin/env python
"""Differential fuzzing on the B200: the UNMODIFIED reference (baseline/_ref, its own CUDA path: ATen fake_quantize ops /
eager LUT composition) against this package, same constructor arguments, same CUDA inputs, outputs compared bit for bit.

    python tools/differential_fuzz.py [--cases 400] [--seed 0] [--json out.json]

Random over: the nine quantizer classes; 2-8 bits and, for the affine classes, 9-24 bits (LUT: up to 16 centroids on 4-10 bit grids, sorted or shuffled, with
duplicates); per-tensor / per-channel on any axis of 1-4-D shapes with odd sizes; float32 / bfloat16 / float16;
thresholds from 2^-6 to 2^4 (powers of two for the POT classes); contiguous tensors and channels_last / transposed views;
inputs with ties planted at the rounding points, +-0, denormals, +-inf (huge values only where the zero point is 0).  The parity ORACLE stays the reference's CPU path
(tests/golden/*); this tool widens the net with the GPU-vs-GPU comparison, which is the drop-in situation itself.
(NaN inputs are left out: libtorch's CUDA and CPU kernels disagree with each other on them.)

The reference derives its parameters on the CUDA device here, where `tensor / python_number` multiplies by a reciprocal
(see mct_quantizers_b200/pytorch/quantizer_utils.py: reference_arithmetic); the tool therefore runs this package with
reference_arithmetic("cuda") unless told otherwise.
"""
import argparse
import json
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=400)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--json", default=None)
    ap.add_argument("--arithmetic", default="cuda", choices=["cpu", "cuda"],
                    help="flavour of the reference's parameter arithmetic this package reproduces (the reference itself uses the CUDA "
                         "flavour on this machine; \"cpu\" shows what the default setting differs in)")
    ap.add_argument("--thr-exp", type=int, default=0, help="0: thresholds in 2^-6 .. 2^4; E > 0: a third of the cases draws them from 2^-E .. 2^E")
    ap.add_argument("--oracle", action="store_true", help="build container (no GPU): the reference on CPU against the C oracle + the oracle's "
                    "restatement of the constructor math (oracle/, tests/golden_util.py) -- widens the pinning of the oracle beyond the fixtures")
    ap.add_argument("--export", action="store_true", help="compare the ONNX-export branch instead: enable_custom_impl() + torch.jit.trace on both sides")
    ap.add_argument("--cpu", action="store_true", help="with --export: run both sides on CPU tensors (the export formulas are plain torch ops; works without a GPU)")
    ap.add_argument("--show", type=int, default=-1, help="print every detail of this case number")
    ap.add_argument("--dry", action="store_true", help="build container (no GPU): construct both sides, run the reference on CPU only")
    args = ap.parse_args()
    warnings.filterwarnings("ignore")
    import logging
    logging.disable(logging.WARNING)
    import mct_quantizers as ref
    assert os.path.abspath(ref.__file__).startswith(os.path.join(ROOT, "baseline", "_ref")), ref.__file__
    from mct_quantizers import pytorch_quantizers as RQ
    import mct_quantizers_b200
    from mct_quantizers_b200.pytorch import quantizers as BQ
    mct_quantizers_b200.reference_arithmetic(args.arithmetic)
    if args.oracle:
        args.dry = True
        args.arithmetic = "cpu"
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import golden_util as G
    if args.cpu:
        assert args.export, "--cpu only makes sense for the export branch (the inference path has no CPU arithmetic)"
        args.arithmetic = "cpu"
        mct_quantizers_b200.reference_arithmetic("cpu")
    dev = torch.device("cpu" if (args.dry or args.cpu) else "cuda:0")
    rng = np.random.default_rng(args.seed)
    dtypes = [torch.float32, torch.bfloat16, torch.float16]

    def rand_shape():
        r = rng.random()
        if r < 0.03:
            return ()                                   # 0-dim tensor
        if r < 0.06:
            return (0, int(rng.integers(1, 9)))         # empty tensor
        nd = int(rng.integers(1, 5))
        dims = [int(rng.choice([1, 2, 3, 5, 7, 8, 9, 16, 17, 31, 32, 33, 64, 100, 129])) for _ in range(nd)]
        while int(np.prod(dims)) > 600000:
            dims[int(np.argmax(dims))] //= 2
        return tuple(max(d, 1) for d in dims)

    def rand_thr(pot):
        if args.thr_exp and rng.random() < 0.34:
            e = int(rng.integers(-args.thr_exp, args.thr_exp + 1))
            return float(2.0 ** e) if pot else float(np.float32(2.0 ** e * rng.uniform(1.0, 2.0)))
        if pot:
            return float(2.0 ** int(rng.integers(-6, 5)))
        return float(np.float32(rng.uniform(2.0 ** -6, 16.0)))

    def make_input(shape, dtype, span, huge=True, limit=None):
        n = int(np.prod(shape)) if len(shape) else 1
        v = rng.normal(0, span * 0.6, size=n).astype(np.float32)
        k = min(n // 3, 4096)
        if k:
            grid = rng.integers(-300, 300, size=k).astype(np.float32) + 0.5          # ties of an integer grid ...
            v[:k] = grid * np.float32(span / 128.0) * np.float32(rng.choice([1.0, 0.5, 2.0]))
        specials = np.array([0.0, -0.0, 1e-40, -1e-40, span, -span] + ([np.inf, -np.inf, 3e38, -3e38] if huge else []), np.float32)
        m = min(specials.size, n)
        if m:
            v[n - m:] = specials[:m]
        if limit is not None:                 # keep |x / scale| < 2^31 and half-precision inputs finite (see below)
            v = np.clip(v, -limit, limit)
        rng.shuffle(v)
        x = torch.from_numpy(v.reshape(shape)).to(dev).to(dtype)
        layout = "contiguous"
        if x.dim() == 4 and rng.random() < 0.3:
            x = x.contiguous(memory_format=torch.channels_last)
            layout = "channels_last"
        elif x.dim() >= 2 and rng.random() < 0.15:
            x = x.transpose(0, 1).contiguous().transpose(0, 1)          # same shape and values, permuted strides
            layout = "transposed view"
        return x, layout

    kinds = ["w_sym", "w_pot", "w_uni", "w_lut_sym", "w_lut_pot", "a_sym", "a_pot", "a_uni", "a_lut"]
    stats = {k: {"cases": 0, "mismatching_cases": 0, "elements": 0, "mismatching_elements": 0} for k in kinds}
    failures = []
    strides_differ = []
    for it in range(args.cases):
        kind = kinds[it % len(kinds)]
        dtype = dtypes[int(rng.integers(0, 3))]
        shape = rand_shape()
        bits = int(rng.integers(2, 9))
        if "lut" not in kind and rng.random() < 0.25:
            bits = int(rng.choice([9, 10, 12, 16, 20, 24]))      # wide grids: beyond 2^21 codes the kernels switch to the rint path
        degenerate = len(shape) == 0 or 0 in shape
        per_channel = bool(rng.random() < 0.7) and not degenerate
        axis = int(rng.integers(0, len(shape))) if not degenerate else 0
        C = shape[axis] if per_channel else 1
        pot = kind in ("w_pot", "w_lut_pot", "a_pot", "a_lut")
        thr = [rand_thr(pot) for _ in range(C)]
        try:
            if kind in ("w_sym", "w_pot"):
                cls = "WeightsSymmetricInferableQuantizer" if kind == "w_sym" else "WeightsPOTInferableQuantizer"
                kw = dict(num_bits=bits, threshold=thr, per_channel=per_channel, channel_axis=axis if per_channel else None)
            elif kind == "w_uni":
                cls = "WeightsUniformInferableQuantizer"
                lo = [-t * float(rng.uniform(0.0, 1.0)) for t in thr]
                kw = dict(num_bits=bits, min_range=lo, max_range=thr, per_channel=per_channel, channel_axis=axis if per_channel else None)
            elif kind in ("w_lut_sym", "w_lut_pot", "a_lut"):
                bw = int(rng.integers(max(bits, 4), 11))
                signed = True if kind != "a_lut" else bool(rng.random() < 0.5)
                lo_v, hi_v = (-2 ** (bw - 1), 2 ** (bw - 1) - 1) if signed else (0, 2 ** bw - 1)
                k = int(rng.integers(2, 2 ** bits + 1))
                lut = [float(v) for v in rng.integers(lo_v, hi_v + 1, size=k)]
                if rng.random() < 0.5:
                    lut = sorted(lut)
                if kind == "a_lut":
                    cls = "ActivationLutPOTInferableQuantizer"
                    kw = dict(num_bits=bits, lut_values=lut, threshold=[thr[0]], signed=signed, lut_values_bitwidth=bw)
                    thr = [thr[0]]
                else:
                    cls = "WeightsLUTSymmetricInferableQuantizer" if kind == "w_lut_sym" else "WeightsLUTPOTInferableQuantizer"
                    kw = dict(num_bits=bits, lut_values=lut, threshold=thr, per_channel=per_channel,
                              channel_axis=axis if per_channel else None, input_rank=len(shape) if per_channel else None,
                              lut_values_bitwidth=bw)
            elif kind in ("a_sym", "a_pot"):
                cls = "ActivationSymmetricInferableQuantizer" if kind == "a_sym" else "ActivationPOTInferableQuantizer"
                kw = dict(num_bits=bits, threshold=[thr[0]], signed=bool(rng.random() < 0.5))
                thr = [thr[0]]
            else:
                cls = "ActivationUniformInferableQuantizer"
                kw = dict(num_bits=bits, min_range=[-thr[0] * float(rng.uniform(0.0, 1.0))], max_range=[thr[0]])
                thr = [thr[0]]
            qr, qb = getattr(RQ, cls)(**kw), getattr(BQ, cls)(**kw)
            # |x / scale| >= 2^31 is outside the contract (SURVEY 8a hazard 3) and libtorch's own kernels disagree there when
            # the zero point is not 0: the CUDA per-channel kernel wraps in int32, the CPU kernel saturates
            uni = kind in ("w_uni", "a_uni")
            limit = min(2.0 ** 30 * min(thr) / 2.0 ** bits, 60000.0 if dtype == torch.float16 else 3e38) if uni else None
            # (--oracle: libtorch's CPU per-channel kernel converts through int64 and turns +inf / 3e38 into quant_min, its
            # per-tensor kernel saturates; outside the contract either way, the oracle saturates)
            huge = not uni and not (args.oracle and per_channel and kind in ("w_sym", "w_pot"))
            x, layout = make_input(shape, dtype, float(np.mean(thr)), huge=huge, limit=limit)
            with torch.no_grad():
                if args.export:
                    qr.enable_custom_impl()
                    qb.enable_custom_impl()
                    yr = torch.jit.trace(lambda t: qr(t), x.clone(), check_trace=False)(x.clone())
                    yb = yr if args.dry else torch.jit.trace(lambda t: qb(t), x.clone(), check_trace=False)(x.clone())
                elif args.oracle:
                    yr = qr(x.clone()).contiguous()
                    case = {"cls": cls, "args": dict(kw), "shape": list(x.shape), "x_dtype": str(dtype).replace("torch.", "")}
                    got = np.asarray(G.oracle_run(case, G.from_torch(x))["y"]).reshape(yr.shape)
                    yb = G.to_torch(np.ascontiguousarray(got), str(yr.dtype).replace("torch.", "")).reshape(yr.shape)
                else:
                    yr = qr(x.clone())
                    yb = yr if args.dry else qb(x.clone())
            assert yr.dtype == yb.dtype and yr.shape == yb.shape, (yr.dtype, yb.dtype, yr.shape, yb.shape)
            if yr.stride() != yb.stride() and yr.numel() > 1 and all(d > 1 for d in yr.shape):
                strides_differ.append({"case": it, "kind": kind, "shape": list(shape), "layout": layout, "reference": list(yr.stride()), "b200": list(yb.stride())})
            view = {4: torch.int32, 2: torch.int16}[yr.element_size()]
            diff = yr.contiguous().view(view) != yb.contiguous().view(view)
            bad = int(diff.sum())
        except Exception as e:                          # noqa: BLE001 -- a crash of either side is a finding too
            failures.append({"case": it, "kind": kind, "cls": cls, "shape": list(shape), "dtype": str(dtype), "error": repr(e)[:300]})
            print(f"case {it} {kind} {shape} {dtype}: ERROR {e!r}"[:300], flush=True)
            continue
        if it == args.show:
            print("case", it, cls, kw, "shape", shape, dtype, layout)
            for i in diff.flatten().nonzero().flatten().tolist()[:5]:
                xs_ = x.contiguous().flatten()
                print("   x", repr(float(xs_[i])), "reference", repr(float(yr.contiguous().flatten()[i])), "b200", repr(float(yb.contiguous().flatten()[i])))
            print("   reference lut tensor:", getattr(qr, "lut_values", None), "threshold", getattr(qr, "threshold", None))
        st = stats[kind]
        st["cases"] += 1
        st["elements"] += x.numel()
        if bad:
            st["mismatching_cases"] += 1
            st["mismatching_elements"] += bad
            idx = diff.flatten().nonzero()[:3].flatten().tolist()
            xs = x.contiguous().flatten()
            failures.append({"case": it, "kind": kind, "cls": cls, "args": {k: (v if not isinstance(v, list) or len(v) < 9 else v[:8] + ["..."]) for k, v in kw.items()},
                             "shape": list(shape), "dtype": str(dtype), "layout": layout, "mismatching_elements": bad,
                             "examples": [{"x": float(xs[i]), "reference": float(yr.contiguous().flatten()[i]), "b200": float(yb.contiguous().flatten()[i])} for i in idx]})
            print(f"case {it} {kind} {shape} {dtype} {layout}: {bad} of {x.numel()} elements differ", flush=True)
    total = sum(s["cases"] for s in stats.values())
    bad_cases = sum(s["mismatching_cases"] for s in stats.values())
    print(f"{total} cases, {sum(s['elements'] for s in stats.values())} elements, {bad_cases} mismatching cases, "
          f"{len([f for f in failures if 'error' in f])} errors")
    for k, s in stats.items():
        print(f"  {k:10s} {s}")
    print("outputs whose strides differ from the reference's:", len(strides_differ), strides_differ[:4])
    if args.json:
        with open(args.json, "w") as f:
            json.dump({"seed": args.seed, "arithmetic": args.arithmetic, "stats": stats, "failures": failures, "strides_differ": strides_differ, "torch": torch.__version__}, f, indent=1)


if __name__ == "__main__":
    main()
