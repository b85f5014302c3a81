#!/usr/bin/env python
"""The REFERENCE's own CUDA path on the B200 (kernel table: "the number to beat"), next to this package:

    python tools/reference_on_gpu.py [--json out.json]

What the unmodified reference runs when its tensors live on a GPU: ATen's `fake_quantize_per_tensor_affine` /
`fake_quantize_per_channel_affine` CUDA kernels for the seven affine quantizers and ~10 eager ops for `lut_quantizer`.
The quantizer objects come from baseline/_ref (the unmodified package); every shape is one of BASELINE.md's configs.
For each case: algorithmic GB/s (SURVEY 8d bytes per element), microseconds per call with the launch queue kept full, and
the host-side cost of one call on a tiny tensor -- for the reference and for mct_quantizers_b200 on the same inputs, and
a bitwise comparison of the two outputs (GPU libtorch vs this package; the parity oracle is the CPU reference).
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))


def burst_ms(fn, nbuf, burst, reps=5):
    for i in range(3):
        fn(i % nbuf)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(burst):
            fn(i % nbuf)
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) / burst)
    ts.sort()
    return ts[len(ts) // 2]


def wall_us(fn, reps=300):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e6


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    import logging
    import warnings
    warnings.filterwarnings("ignore")
    import numpy as np
    import mct_quantizers as ref                       # baseline/_ref: parameters are created on `cuda` (get_working_device)
    assert os.path.abspath(ref.__file__).startswith(os.path.join(ROOT, "baseline", "_ref")), ref.__file__
    from mct_quantizers import pytorch_quantizers as RQ
    import mct_quantizers_b200 as mctq
    from mct_quantizers_b200.pytorch import quantizers as BQ
    mctq.reference_arithmetic("cuda")       # compare with what the reference derives on THIS machine (see quantizer_utils.py)
    for name in ("MCT Quantizers", "MCT Quantizers B200"):
        logging.getLogger(name).setLevel(logging.ERROR)
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1234)
    rows = []

    def case(config, what, make_ref, make_b200, xs, bytes_per_elem, burst):
        qr, qb = make_ref(), make_b200()
        n = xs[0].numel()
        with torch.no_grad():
            yr, yb = qr(xs[0]), qb(xs[0])
            view = torch.int32 if yr.dtype == torch.float32 else torch.int16
            same = bool(torch.equal(yr.view(view), yb.view(view)))
            mism = int((yr.view(view) != yb.view(view)).sum()) if not same else 0
            del yr, yb
            ms_r = burst_ms(lambda i: qr(xs[i]), len(xs), burst)
            ms_b = burst_ms(lambda i: qb(xs[i]), len(xs), burst)
        row = {"config": config, "what": what, "elements": n, "reference_ms": round(ms_r, 4), "b200_ms": round(ms_b, 4),
               "reference_GBs": round(n * bytes_per_elem / ms_r / 1e6, 1), "b200_GBs": round(n * bytes_per_elem / ms_b / 1e6, 1),
               "speedup": round(ms_r / ms_b, 2), "bitwise_equal_to_reference_on_gpu": same, "mismatching_elements": mism}
        rows.append(row)
        print(f"{config:3s} {what:84s} ref {ms_r * 1e3:10.1f} us {row['reference_GBs']:8.1f} GB/s | b200 {ms_b * 1e3:9.1f} us "
              f"{row['b200_GBs']:8.1f} GB/s  x{row['speedup']:<7} equal={same}" + ("" if same else f" ({mism} differ)"), flush=True)

    # C1: ResNet-18 conv weights (per-channel symmetric), per-layer calls, and the 0.6 MB activation
    from scale_bench import RESNET18_CONVS
    ws = [torch.empty(s, device=dev).normal_(0, (2.0 / (s[0] * s[2] * s[3])) ** 0.5, generator=g) for s in RESNET18_CONVS]
    thrs = [[float(v) for v in w.abs().flatten(1).amax(1)] for w in ws]
    qr = [RQ.WeightsSymmetricInferableQuantizer(8, t, True, 0) for t in thrs]
    qb = [BQ.WeightsSymmetricInferableQuantizer(8, t, True, 0) for t in thrs]
    n_w = sum(w.numel() for w in ws)
    ms_r = burst_ms(lambda i: [q(w) for q, w in zip(qr, ws)], 1, 5)
    ms_b = burst_ms(lambda i: [q(w) for q, w in zip(qb, ws)], 1, 5)
    from mct_quantizers_b200.pytorch.model_quantization import WeightPlan
    plan = WeightPlan([(str(k), w, q) for k, (w, q) in enumerate(zip(ws, qb))])
    ms_p = burst_ms(lambda i: plan.run(), 1, 5)
    eq = all(torch.equal(a(w), b(w)) for a, b, w in zip(qr, qb, ws))
    rows.append({"config": "C1", "what": "ResNet-18 20 conv weights WeightsSymmetric 8-bit per-channel, 20 per-layer calls", "elements": n_w,
                 "reference_ms": round(ms_r, 4), "b200_ms": round(ms_b, 4), "b200_one_launch_ms": round(ms_p, 4),
                 "reference_GBs": round(n_w * 8 / ms_r / 1e6, 1), "b200_GBs": round(n_w * 8 / ms_b / 1e6, 1),
                 "b200_one_launch_GBs": round(n_w * 8 / ms_p / 1e6, 1), "bitwise_equal_to_reference_on_gpu": bool(eq)})
    print(f"C1  ResNet-18 20 conv weights: reference {ms_r * 1e3:.1f} us (20 calls, 2 host syncs each) | b200 per-layer {ms_b * 1e3:.1f} us | "
          f"b200 one launch {ms_p * 1e3:.1f} us  equal={eq}", flush=True)
    x1 = [torch.empty((1, 3, 224, 224), device=dev).normal_(0, 1, generator=g) for _ in range(4)]
    case("C1", "ActivationPOT 8-bit thr=4 signed on 1x3x224x224 f32", lambda: RQ.ActivationPOTInferableQuantizer(8, [4.0], True),
         lambda: BQ.ActivationPOTInferableQuantizer(8, [4.0], True), x1, 8, 50)
    hr = ref.PytorchActivationQuantizationHolder(RQ.ActivationPOTInferableQuantizer(8, [4.0], True))
    hb = mctq.PytorchActivationQuantizationHolder(BQ.ActivationPOTInferableQuantizer(8, [4.0], True))
    host = {"reference_holder_call_us": round(wall_us(lambda: hr(x1[0])), 2), "b200_holder_call_us": round(wall_us(lambda: hb(x1[0])), 2)}
    wq_r, wq_b = qr[5], qb[5]
    host["reference_weight_call_us"] = round(wall_us(lambda: wq_r(ws[5])), 2)
    host["b200_weight_call_us"] = round(wall_us(lambda: wq_b(ws[5])), 2)
    print("host-side cost per call (wall clock, tiny tensors, queue kept full):", host, flush=True)

    # C2: the largest and a mid-size MobileNetV2 activation site, ActivationUniform 8-bit f32
    for shp in ((256, 96, 112, 112), (256, 192, 28, 28)):
        xs = [torch.empty(shp, device=dev).normal_(0, 1, generator=g) for _ in range(2)]
        lo, hi = float(xs[0].min()), float(xs[0].max())
        case("C2", f"ActivationUniform 8-bit [min,max] on {shp} f32", lambda: RQ.ActivationUniformInferableQuantizer(8, [lo], [hi]),
             lambda: BQ.ActivationUniformInferableQuantizer(8, [lo], [hi]), xs, 8, 6)
        del xs
    # C3: Llama-shaped LUT weights
    lut = [float(v) for v in sorted(np.random.default_rng(0).choice(np.arange(-128, 128), size=16, replace=False))]
    for dt in (torch.float32, torch.bfloat16):
        for shp in ((11008, 4096), (4096, 11008)):
            xs = [torch.empty(shp, device=dev).normal_(0, 0.02, generator=g).to(dt) for _ in range(2)]
            thr = [float(v) for v in xs[0].float().abs().amax(1)]
            case("C3", f"WeightsLUTSymmetric 4-bit K=16 per-channel axis 0 on {shp} {str(dt)[6:]}",
                 lambda: RQ.WeightsLUTSymmetricInferableQuantizer(4, lut, thr, True, 0, 2),
                 lambda: BQ.WeightsLUTSymmetricInferableQuantizer(4, lut, thr, True, 0, 2), xs, (4 if dt == torch.float32 else 2) + 4, 4)
            del xs
            torch.cuda.empty_cache()
    # per-channel affine on a Llama-shaped matrix (ATen's per-channel kernel)
    xs = [torch.empty((11008, 4096), device=dev).normal_(0, 0.02, generator=g) for _ in range(2)]
    thr = [float(v) for v in xs[0].abs().amax(1)]
    case("--", "WeightsSymmetric 8-bit per-channel axis 0 on (11008, 4096) f32", lambda: RQ.WeightsSymmetricInferableQuantizer(8, thr, True, 0),
         lambda: BQ.WeightsSymmetricInferableQuantizer(8, thr, True, 0), xs, 8, 6)
    del xs
    # C4: ViT activations bf16
    for shp in ((256, 197, 768), (256, 197, 3072)):
        xs = [torch.empty(shp, device=dev).normal_(0, 1, generator=g).bfloat16() for _ in range(4)]
        case("C4", f"ActivationSymmetric 8-bit thr=3.7 on {shp} bf16", lambda: RQ.ActivationSymmetricInferableQuantizer(8, [3.7], True),
             lambda: BQ.ActivationSymmetricInferableQuantizer(8, [3.7], True), xs, 4, 12)
        del xs
    # C5: 1 GB f32
    xs = [torch.empty(1 << 28, device=dev).uniform_(-50, 50, generator=g) for _ in range(2)]
    case("C5", "ActivationSymmetric 8-bit thr=4 on 1 GiB f32", lambda: RQ.ActivationSymmetricInferableQuantizer(8, [4.0], True),
         lambda: BQ.ActivationSymmetricInferableQuantizer(8, [4.0], True), xs, 8, 4)
    if args.json:
        with open(args.json, "w") as f:
            json.dump({"device": torch.cuda.get_device_name(0), "torch": torch.__version__, "rows": rows, "host_us": host}, f, indent=1)


if __name__ == "__main__":
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    main()
