#!/usr/bin/env python
"""All five BASELINE.json configs through the PUBLIC API (quantizer objects, wrapper, holders) on one GPU.

    python tools/config_bench.py [--json out.json] [--quick]

C1  ResNet-18: 20 conv weights (WeightsSymmetric 8-bit per-channel) + ActivationPOT 8-bit on 1x3x224x224  (latency-bound)
C2  MobileNetV2 batch 256                                                     -> that is bench.py itself
C3  Llama-7B-shaped linears, WeightsLUTSymmetric 4-bit per-channel, f32 and bf16 weights (3 matrices of one layer)
C4  ViT-B/16 activations, ActivationSymmetric 8-bit, bf16, 256 images per GPU: (256,197,768) and (256,197,3072)
C5  size sweep 1 MB .. 4 GB (input bytes), f32 / bf16, ActivationSymmetric thr=4 and ActivationUniform [-1, 2.3]

Timing: CUDA events on the current stream, inputs rotated over >= 3 buffers whenever the working set could be L2
resident, median of `reps`; GB/s = algorithmic bytes / time.  The per-call host overhead (Python + operator
dispatch + ctypes) is reported separately for the small-tensor regime.
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mct_quantizers_b200 as mctq  # noqa: E402
from mct_quantizers_b200.pytorch import quantizers as Q  # noqa: E402

DEV = torch.device("cuda:0")
RESNET18_CONVS = [(64, 3, 7, 7)] + [(64, 64, 3, 3)] * 4 + [(128, 64, 3, 3)] + [(128, 128, 3, 3)] * 3 + [(128, 64, 1, 1)] + \
    [(256, 128, 3, 3)] + [(256, 256, 3, 3)] * 3 + [(256, 128, 1, 1)] + [(512, 256, 3, 3)] + [(512, 512, 3, 3)] * 3 + [(512, 256, 1, 1)]


def time_gpu(fn, reps=20, warm=3):
    for _ in range(warm):
        fn(0)
    torch.cuda.synchronize()
    ts = []
    for i in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn(i)
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def time_burst(fn, burst=20, reps=5):
    """GPU-side rate with the launch queue kept full: `burst` calls between two events."""
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(burst):
            fn(i)
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) / burst)
    ts.sort()
    return ts[len(ts) // 2]


def time_wall(fn, reps=200):
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
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    g = torch.Generator(device=DEV).manual_seed(1234)
    out = {"device": torch.cuda.get_device_name(0), "results": []}
    peak = 6457.7
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass

    def rec(config, what, ms, nbytes, extra=None):
        gbs = nbytes / ms / 1e6
        r = {"config": config, "what": what, "ms": round(ms, 4), "GBs": round(gbs, 1), "pct_of_copy_peak": round(100 * gbs / peak, 1),
             "pct_of_8TBs": round(gbs / 80, 1)}
        r.update(extra or {})
        out["results"].append(r)
        print(f"{config:3s} {what:78s} {ms:9.4f} ms {gbs:8.1f} GB/s {100 * gbs / peak:6.1f}% copy-peak {gbs / 80:5.1f}% of 8TB/s", flush=True)

    # ---------------- C1: ResNet-18 weights + tiny activation
    torch.manual_seed(0)
    wrappers = []
    for shp in RESNET18_CONVS:
        conv = torch.nn.Conv2d(shp[1], shp[0], shp[2], bias=False)
        thr = [float(v) for v in conv.weight.detach().abs().flatten(1).amax(1)]
        wrappers.append(mctq.PytorchQuantizationWrapper(conv, {'weight': Q.WeightsSymmetricInferableQuantizer(8, thr, True, 0)}))
    model = torch.nn.Sequential(*wrappers).to(DEV)
    n_w = sum(w.weight.numel() for w in wrappers)
    ms = time_gpu(lambda i: [w.get_quantized_weights() for w in wrappers])
    rec("C1", f"ResNet-18 20 conv weights, per-layer get_quantized_weights() ({n_w} elems, 20 launches)", ms, n_w * 8)
    from mct_quantizers_b200.pytorch.model_quantization import WeightPlan
    plan = WeightPlan([tv for w in wrappers for tv in w.get_weights_vars()])
    ms = time_gpu(lambda i: plan.run())
    rec("C1", "ResNet-18 20 conv weights, one multi-tensor launch (WeightPlan.run)", ms, n_w * 8)
    # the same 20 tensors with 4-bit LUT quantizers: per-layer calls vs one multi-tensor LUT launch
    import numpy as np
    lut = [float(v) for v in sorted(np.random.default_rng(0).choice(np.arange(-128, 128), size=16, replace=False))]
    lut_vars = []
    for k, w in enumerate(wrappers):
        wt = w.get_weights_vars()[0][1].detach()              # the float weight lives on the wrapper, not on the wrapped layer
        thr = [float(v) + 1e-6 for v in wt.abs().flatten(1).amax(1)]
        lut_vars.append((f"w{k}", wt, Q.WeightsLUTSymmetricInferableQuantizer(4, lut, thr, True, 0, 4)))
    ms = time_gpu(lambda i: [q(wt) for _, wt, q in lut_vars])
    rec("C1", "ResNet-18 20 conv weights, LUT 4-bit, per-layer calls (20 launches)", ms, n_w * 8)
    lplan = WeightPlan(lut_vars)
    assert lplan.lut_plan is not None and not lplan.other
    ms = time_gpu(lambda i: lplan.run())
    rec("C1", "ResNet-18 20 conv weights, LUT 4-bit, one multi-tensor launch (WeightPlan.run)", ms, n_w * 8)
    ms = time_gpu(lambda i: mctq.quantize_model_weights(model), reps=10)
    rec("C1", "ResNet-18 quantize_model_weights(model) incl. plan construction", ms, n_w * 8)
    holder = mctq.PytorchActivationQuantizationHolder(Q.ActivationPOTInferableQuantizer(8, [4.0], True)).to(DEV)
    x = torch.randn(1, 3, 224, 224, device=DEV)
    ms = time_gpu(lambda i: holder(x))
    rec("C1", "ActivationPOT 8-bit on 1x3x224x224 f32 (0.6 MB: launch-latency bound by construction)", ms, x.numel() * 8)
    us = time_wall(lambda: holder(x))
    out["host_overhead_us_per_holder_call"] = round(us, 2)
    print(f"    host-side cost of one holder call (Python + dispatcher + ctypes, wall clock, tiny tensor): {us:.1f} us")
    ref_us = time_wall(lambda: torch.fake_quantize_per_tensor_affine(x, 4.0 / 128, 0, -128, 127))
    out["host_overhead_us_per_aten_call"] = round(ref_us, 2)
    print(f"    same call through ATen's fake_quantize_per_tensor_affine (what the reference does):       {ref_us:.1f} us")

    # ---------------- C3: Llama-7B-shaped LUT weights
    lut = [float(v) for v in sorted(torch.randperm(256, generator=torch.Generator().manual_seed(0))[:16].sub(128).tolist())]
    for dt in (torch.float32, torch.bfloat16):
        for shp in ((11008, 4096), (4096, 11008)):
            Ws = [torch.empty(shp, device=DEV).normal_(0, 0.02, generator=g).to(dt) for _ in range(3)]
            thr = [float(v) for v in Ws[0].float().abs().amax(1)]
            q = Q.WeightsLUTSymmetricInferableQuantizer(4, lut, thr, True, 0, 2)
            q(Ws[0])
            ms = time_gpu(lambda i: q(Ws[i % 3]))
            nb = Ws[0].numel() * (Ws[0].element_size() + 4)
            rec("C3", f"WeightsLUTSymmetric 4-bit K=16 per-channel axis 0 on {shp[0]}x{shp[1]} {str(dt)[6:]} [single call, synced]", ms, nb)
            ms = time_burst(lambda i: q(Ws[i % 3]))
            rec("C3", f"WeightsLUTSymmetric 4-bit K=16 per-channel axis 0 on {shp[0]}x{shp[1]} {str(dt)[6:]} [20 calls queued]", ms, nb)
            del Ws

    # ---------------- C4: ViT-B/16 activations bf16
    for shp in ((256, 197, 768), (256, 197, 3072)):
        xs = [torch.empty(shp, device=DEV).normal_(0, 1, generator=g).bfloat16() for _ in range(4)]
        for thr in (4.0, 3.7):
            h = mctq.PytorchActivationQuantizationHolder(Q.ActivationSymmetricInferableQuantizer(8, [thr], True))
            ms = time_gpu(lambda i: h(xs[i % 4]))
            rec("C4", f"ActivationSymmetric 8-bit thr={thr} on {shp} bf16 [single call, synced]", ms, xs[0].numel() * 4)
            ms = time_burst(lambda i: h(xs[i % 4]))
            rec("C4", f"ActivationSymmetric 8-bit thr={thr} on {shp} bf16 [20 calls queued]", ms, xs[0].numel() * 4)
        del xs

    # ---------------- C5: size sweep
    sizes_mb = [1, 4, 16, 64, 256, 1024] + ([] if args.quick else [4096])
    for dt, es in ((torch.float32, 4), (torch.bfloat16, 2)):
        for mb in sizes_mb:
            n = mb * (1 << 20) // es
            nbuf = max(3, min(16, (512 << 20) // (mb << 20) + 1)) if mb < 512 else 2
            xs = [torch.empty(n, device=DEV).uniform_(-50, 50, generator=g).to(dt) for _ in range(nbuf)]
            for name, q in (("ActivationSymmetric thr=4", Q.ActivationSymmetricInferableQuantizer(8, [4.0], True)),
                            ("ActivationUniform [-1,2.3]", Q.ActivationUniformInferableQuantizer(8, [-1.0], [2.3]))):
                ms = time_gpu(lambda i: q(xs[i % nbuf]), reps=30 if mb <= 64 else 10)
                rec("C5", f"{name} {mb} MB {str(dt)[6:]} (rotating {nbuf} buffers) [single call, synced]", ms, 2 * n * es)
                if mb <= 256:
                    ms = time_burst(lambda i: q(xs[i % nbuf]), burst=max(nbuf, 16))
                    rec("C5", f"{name} {mb} MB {str(dt)[6:]} (rotating {nbuf} buffers) [calls queued]", ms, 2 * n * es)
            del xs
            torch.cuda.empty_cache()
    if args.json:
        with open(args.json, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
