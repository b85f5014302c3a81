#!/usr/bin/env python
"""The hot path IN SITU: forward passes of whole quantized models on one B200, the unmodified reference (baseline/_ref)
against this package, same weights, same thresholds, same inputs.

    python tools/model_forward_bench.py [--json out.json] [--models mobilenet_v2,resnet18] [--batches 1,32,256]

The models are torchvision architectures (random init) quantized the way MCT exports them: every Conv2d / Linear sits in a
`PytorchQuantizationWrapper` with an 8-bit per-channel `WeightsSymmetricInferableQuantizer` (thresholds = per-channel
absmax), every ReLU / ReLU6 is followed by a `PytorchActivationQuantizationHolder` with an 8-bit
`ActivationUniformInferableQuantizer`.  Arms:

  reference        mct_quantizers 1.6.0 from baseline/_ref, its own CUDA path (ATen fake_quantize ops; the wrapper
                   re-quantizes every weight on every forward, two host syncs per per-channel call)
  b200             mct_quantizers_b200, nothing else changed (the drop-in case)
  b200+plan        + `plan_model_weights(model).enable()`: all weights quantized by ONE launch, wrappers skip theirs
  b200+plan+fuse   + `fuse_activation_producers(model)`: ReLU / ReLU6 run inside the holder's kernel

Reported per arm and batch: milliseconds per forward (CUDA events, median), for batch 1 wall-clock per forward with the
queue drained (host-bound regime), for batches <= 32 the replay time of the forward captured in a CUDA graph (the
reference cannot be captured); outputs of every arm are compared with the reference arm's.  Convolutions are
the same cuDNN kernels in every arm, so the differences are the fake-quant path and its host side.
"""
import argparse
import copy
import json
import os
import sys
import time
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))


def quantize_model(model, pkg, Q):
    """In-place: wrap Conv2d / Linear, put a holder behind every ReLU / ReLU6.  `pkg` / `Q`: package and its quantizers."""
    def convert(parent):
        for name, child in list(parent.named_children()):
            if isinstance(child, (torch.nn.Conv2d, torch.nn.Linear)):
                w = child.weight.detach()
                thr = w.abs().reshape(w.shape[0], -1).amax(1).clamp_min(1e-6)
                q = Q.WeightsSymmetricInferableQuantizer(num_bits=8, threshold=[float(t) for t in thr], per_channel=True, channel_axis=0)
                setattr(parent, name, pkg.PytorchQuantizationWrapper(child, {"weight": q}))
            elif isinstance(child, (torch.nn.ReLU, torch.nn.ReLU6)):
                hi = 6.0 if isinstance(child, torch.nn.ReLU6) else 4.0
                holder = pkg.PytorchActivationQuantizationHolder(Q.ActivationUniformInferableQuantizer(8, [0.0], [hi]))
                setattr(parent, name, torch.nn.Sequential(type(child)(inplace=False), holder))
            else:
                convert(child)
    convert(model)
    return model


def events_ms(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def wall_ms(fn, reps):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def graph_replay_ms(model, x, want):
    """Capture one forward in a CUDA graph and time replays.  The reference cannot be captured: its per-channel weights path
    reads scale / zero point back to the host (`.item()`) inside every forward."""
    try:
        with torch.no_grad():
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):
                    model(x)
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                y = model(x)
            g.replay()
            torch.cuda.synchronize()
            ok = bool(torch.equal(y, want))
            ms = events_ms(g.replay, 30)
        return round(ms, 4), ("replay == eager" if ok else "replay DIFFERS from eager")
    except Exception as e:          # noqa: BLE001 -- reported, not hidden
        torch.cuda.synchronize()
        return None, "not capturable (" + str(e).split("\n")[0][:90] + ")"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--models", default="mobilenet_v2,resnet18")
    ap.add_argument("--batches", default="1,32,256")
    args = ap.parse_args()
    warnings.filterwarnings("ignore")
    import logging
    logging.disable(logging.WARNING)
    import torchvision
    import mct_quantizers as ref                       # baseline/_ref: the unmodified reference
    from mct_quantizers.pytorch import quantizers as refQ
    import mct_quantizers_b200 as b2
    from mct_quantizers_b200.pytorch import quantizers as b2Q
    b2.reference_arithmetic("cuda")         # the reference derives its parameters on the GPU here (see quantizer_utils.py)
    assert "baseline" in ref.__file__, ref.__file__
    dev = torch.device("cuda:0")
    torch.backends.cudnn.benchmark = False
    rows = []
    for mname in args.models.split(","):
        torch.manual_seed(0)
        base = getattr(torchvision.models, mname)(weights=None).eval()
        arms = {}
        with torch.no_grad():
            arms["reference"] = quantize_model(copy.deepcopy(base), ref, refQ).to(dev)
            arms["b200"] = quantize_model(copy.deepcopy(base), b2, b2Q).to(dev)
            planned = quantize_model(copy.deepcopy(base), b2, b2Q).to(dev)
            plan = b2.plan_model_weights(planned).enable()
            arms["b200+plan"] = planned
            fused = quantize_model(copy.deepcopy(base), b2, b2Q).to(dev)
            plan2 = b2.plan_model_weights(fused).enable()
            arms["b200+plan+fuse"] = b2.fuse_activation_producers(fused)
        n_w = sum(1 for m in arms["b200"].modules() if isinstance(m, b2.PytorchQuantizationWrapper))
        n_h = sum(1 for m in arms["b200"].modules() if isinstance(m, b2.PytorchActivationQuantizationHolder))
        print(f"== {mname}: {n_w} wrapped layers, {n_h} activation holders", flush=True)
        for batch in [int(b) for b in args.batches.split(",")]:
            x = torch.randn(batch, 3, 224, 224, device=dev)
            want = None
            for arm, model in arms.items():
                with torch.no_grad():
                    y = model(x)
                    if want is None:
                        want = y
                    same = bool(torch.equal(y, want))
                    maxdiff = float((y - want).abs().max())
                    reps = 30 if batch <= 32 else 10
                    ms = events_ms(lambda: model(x), reps)
                    wall = wall_ms(lambda: model(x), 20) if batch == 1 else None
                graph_ms, graph_note = graph_replay_ms(model, x, want) if batch <= 32 else (None, None)
                row = {"model": mname, "batch": batch, "arm": arm, "ms_per_forward": round(ms, 4),
                       "cuda_graph_ms": graph_ms, "cuda_graph_note": graph_note,
                       "wall_ms_batch1": None if wall is None else round(wall, 4), "bit_equal_to_reference": same,
                       "max_abs_diff": maxdiff, "wrapped_layers": n_w, "holders": n_h}
                rows.append(row)
                print(f"{mname:13s} batch {batch:4d}  {arm:15s} {ms:9.3f} ms / forward" +
                      (f"  (wall, drained queue: {wall:7.3f} ms)" if wall is not None else "") +
                      f"  equal={same} maxdiff={maxdiff:.3g}" +
                      ("" if graph_note is None else f"  | CUDA graph: {graph_note if graph_ms is None else f'{graph_ms:.3f} ms / replay'}"), flush=True)
        del plan, plan2
    if args.json:
        with open(args.json, "w") as f:
            json.dump({"rows": rows, "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0)}, f, indent=1)


if __name__ == "__main__":
    main()
