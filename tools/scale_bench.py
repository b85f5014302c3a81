#!/usr/bin/env python
"""BASELINE.json configs C1 / C3 / C4 / C5 at 1 / 2 / 4 / 8 GPUs (C2 is bench.py itself, which also embeds these rows as `per_config`), one process per GPU:

    python tools/scale_bench.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/scale_bench.py [--json out.json]

C3  Llama-7B-shaped linears (32 layers x {gate 11008x4096, up 11008x4096, down 4096x11008}), WeightsLUTSymmetric 4-bit K=16
    per-channel axis 0, f32 and bf16: the 96 matrices are sharded BY LAYER over the ranks (sharding.shard_layers);
    strong scaling (the model is fixed).
C4  ViT-B/16 activations, ActivationSymmetric 8-bit bf16, global batch 2048 sharded BY BATCH (sharding.shard_batch):
    sites (B,197,768) and (B,197,3072); strong scaling.
C5  size sweep 1 MB .. 16 GB (input bytes, whole job) f32 / bf16, ActivationSymmetric thr=4 and ActivationUniform [-1,2.3]:
    every rank takes a contiguous 1/N slice; strong scaling.

No collective on the data path.  Timing: CUDA events per rank around `reps` back-to-back passes over the rank's shard
after warm-up (buffers rotate so that the working set exceeds L2), MAX over ranks (all_reduce), aggregate GB/s =
algorithmic bytes of the whole job / that time.  NCCL is only used for the time reduction and a checksum all-gather.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mct_quantizers_b200 as mctq  # noqa: E402,F401
from mct_quantizers_b200 import sharding  # noqa: E402
from mct_quantizers_b200.pytorch import quantizers as Q  # noqa: E402


RESNET18_CONVS = [(64, 3, 7, 7)] + [(64, 64, 3, 3)] * 4 + [(128, 64, 3, 3)] + [(128, 128, 3, 3)] * 3 + [(128, 64, 1, 1)] + \
    [(256, 128, 3, 3)] + [(256, 256, 3, 3)] * 3 + [(256, 128, 1, 1)] + [(512, 256, 3, 3)] + [(512, 512, 3, 3)] * 3 + [(512, 256, 1, 1)]


def run_configs(dev, rank, world, reps=5, max_gb=16.0, log=sys.stdout):
    """Times C1 / C3 / C4 / C5 on an initialised process group (or a single process) and returns the result rows
    (identical on every rank: times are max-reduced).  bench.py calls this for its `per_config` block."""
    import logging
    logging.getLogger("MCT Quantizers B200").setLevel(logging.ERROR)
    peak = 6451.2
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    rows = []

    def timed(fn, reps):
        """fn(i) enqueues one pass over this rank's shard; returns max-over-ranks milliseconds per pass."""
        for i in range(2):
            fn(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(reps):
            fn(i)
        b.record()
        b.synchronize()
        t = torch.tensor([a.elapsed_time(b) / reps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def report(name, total_bytes, ms, extra=None, scale_div=None):
        gbs = total_bytes / ms / 1e6
        div = world if scale_div is None else scale_div
        row = {"config": name, "n_gpus": world, "ms": round(ms, 4), "GBs": round(gbs, 1), "GBs_per_gpu": round(gbs / div, 1),
               "frac_of_copy_peak_per_gpu": round(gbs / div / peak, 4), "pct_of_8TBs_per_gpu": round(gbs / div / 80.0, 2)}
        if extra:
            row.update(extra)
        rows.append(row)
        if rank == 0 and log is not None:
            print(f"{name:86s} N={world}  {ms:9.4f} ms  {gbs:9.1f} GB/s  ({gbs / div / peak * 100:5.1f}% of copy peak per GPU)", file=log, flush=True)

    g = torch.Generator(device=dev).manual_seed(1234 + rank)

    def timed_median(fn, n=20):
        """median of `n` individually timed (synchronised) calls: the latency-bound C1 rows; max over ranks"""
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        t = torch.tensor([ts[len(ts) // 2]], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # ------------------------------------------------------------------ C1: ResNet-18 conv weights + one tiny activation
    # (every rank runs the whole config: 45 MB of weights are not worth sharding; reported per GPU)
    from mct_quantizers_b200.pytorch.model_quantization import WeightPlan
    gw = torch.Generator(device=dev).manual_seed(0)
    c1_vars = []
    for k, shp in enumerate(RESNET18_CONVS):
        fan_out = shp[0] * shp[2] * shp[3]
        w = torch.empty(shp, device=dev).normal_(0, (2.0 / fan_out) ** 0.5, generator=gw)
        thr = [float(v) for v in w.abs().flatten(1).amax(1).double().cpu()]
        c1_vars.append((f"conv{k}", w, Q.WeightsSymmetricInferableQuantizer(8, thr, True, 0)))
    n_w = sum(w.numel() for _, w, _ in c1_vars)
    c1_plan = WeightPlan(c1_vars)
    ms = timed_median(lambda: c1_plan.run())
    report("C1 ResNet-18 20 conv weights WeightsSymmetric 8-bit per-channel f32, ONE multi-tensor launch (replicated per rank)",
           n_w * 8, ms, {"kernel": "fq_affine_multi_kernel", "scaling": "replica", "latency_bound": True}, scale_div=1)
    ms = timed_median(lambda: [q(w) for _, w, q in c1_vars])
    report("C1 ResNet-18 20 conv weights WeightsSymmetric 8-bit per-channel f32, 20 per-layer calls (replicated per rank)",
           n_w * 8, ms, {"kernel": "fq_affine_kernel<float, CH_VEC|CH_ELEM, prepared>", "scaling": "replica", "latency_bound": True}, scale_div=1)
    x1 = torch.empty((1, 3, 224, 224), device=dev).normal_(0, 1, generator=g)
    h1 = mctq.PytorchActivationQuantizationHolder(Q.ActivationPOTInferableQuantizer(8, [4.0], True))
    ms = timed_median(lambda: h1(x1))
    report("C1 ActivationPOT 8-bit thr=4 on 1x3x224x224 f32 (0.6 MB, one call, replicated per rank)", x1.numel() * 8, ms,
           {"kernel": "fq_affine_kernel<float, CH_PT>", "scaling": "replica", "latency_bound": True}, scale_div=1)
    del c1_plan, c1_vars

    # ------------------------------------------------------------------ C3: Llama-7B linears, layer-sharded
    import numpy as np
    lut = [float(v) for v in sorted(np.random.default_rng(0).choice(np.arange(-128, 128), size=16, replace=False))]
    shapes = []
    for _ in range(32):
        shapes += [(11008, 4096), (11008, 4096), (4096, 11008)]
    mine = sharding.shard_layers([a * b for a, b in shapes], world)[rank]
    for dt in (torch.float32, torch.bfloat16):
        bufs = {}
        for shp in set(shapes):
            bufs[shp] = [torch.empty(shp, device=dev).normal_(0, 0.02, generator=g).to(dt) for _ in range(2)]
        quant = {}
        for shp in set(shapes):          # one quantizer per distinct shape (thresholds = row maxima of buffer 0)
            thr = bufs[shp][0].float().abs().amax(1).double().cpu().tolist()
            quant[shp] = Q.WeightsLUTSymmetricInferableQuantizer(4, lut, thr, True, 0, 2)

        def pass_c3(i):
            out = None
            for k, li in enumerate(mine):
                shp = shapes[li]
                out = quant[shp](bufs[shp][(i + k) & 1])
            return out
        ms = timed(pass_c3, reps)
        es = 4 if dt == torch.float32 else 2
        total = sum(a * b for a, b in shapes) * (es + 4)
        report(f"C3 Llama-7B 96 linears WeightsLUTSymmetric 4-bit K=16 per-channel {str(dt).split('.')[-1]} (layer-sharded)", total, ms,
               {"layers_on_rank0": len(sharding.shard_layers([a * b for a, b in shapes], world)[0]),
                "kernel": "fq_lutx_kernel<%s, CH_VEC> (xy records)" % ("float" if es == 4 else "bf16"), "scaling": "strong"})
        with mctq.private_stream():          # the weights exist before the pass starts: opt-in early order (loads before the wait)
            ms = timed(pass_c3, reps)
        report(f"C3 Llama-7B 96 linears WeightsLUTSymmetric 4-bit K=16 per-channel {str(dt).split('.')[-1]} (layer-sharded, private_stream)", total, ms,
               {"layers_on_rank0": len(sharding.shard_layers([a * b for a, b in shapes], world)[0]),
                "kernel": "fq_lutx_kernel<%s, CH_VEC> (xy records), early order" % ("float" if es == 4 else "bf16"), "scaling": "strong"})
        # the same shard as ONE launch: WeightPlan gathers the rank's LUT weight quantizers into a LutMultiPlan
        # (mctq_fq_lut_prepared_multi); every layer has its own weight tensor here, as in the real model
        layer_w = [torch.empty(shapes[li], device=dev).normal_(0, 0.02, generator=g).to(dt) for li in mine]
        wplan = WeightPlan([(f"layer{li}", w, quant[shapes[li]]) for li, w in zip(mine, layer_w)])
        assert wplan.lut_plan is not None and not wplan.other
        per_layer = quant[shapes[mine[0]]](layer_w[0])
        assert torch.equal(wplan.run()[0], per_layer)
        ms = timed(lambda i: wplan.run(), reps)
        report(f"C3 Llama-7B 96 linears WeightsLUTSymmetric 4-bit K=16 per-channel {str(dt).split('.')[-1]} (layer-sharded, ONE launch per rank)",
               total, ms, {"layers_on_rank0": len(sharding.shard_layers([a * b for a, b in shapes], world)[0]),
                           "kernel": "fq_lut_multi_kernel<%s, CH_VEC, xy> (one launch per 64 tensors)" % ("float" if es == 4 else "bf16"), "scaling": "strong"})
        del bufs, quant, wplan, layer_w, per_layer
        torch.cuda.empty_cache()

    # ------------------------------------------------------------------ C4: ViT-B/16 activations, batch-sharded
    b0, b1 = sharding.shard_batch(2048, world, rank)
    for thr in (4.0, 3.7):
        q = Q.ActivationSymmetricInferableQuantizer(8, [thr], True)
        for feat in (768, 3072):
            xs = [torch.empty((b1 - b0, 197, feat), device=dev).normal_(0, 1, generator=g).bfloat16() for _ in range(2)]
            ms = timed(lambda i: q(xs[i & 1]), reps * 2)
            report(f"C4 ViT-B/16 ActivationSymmetric 8-bit thr={thr} bf16 (2048,197,{feat}) (batch-sharded)", 2048 * 197 * feat * 4, ms,
                   {"rows_per_rank": b1 - b0, "kernel": "fq_affine_kernel<bf16, CH_PT>", "scaling": "strong"})
            if thr == 3.7:
                with mctq.private_stream():
                    ms = timed(lambda i: q(xs[i & 1]), reps * 2)
                report(f"C4 ViT-B/16 ActivationSymmetric 8-bit thr={thr} bf16 (2048,197,{feat}) (batch-sharded, private_stream)", 2048 * 197 * feat * 4, ms,
                       {"rows_per_rank": b1 - b0, "kernel": "fq_affine_kernel<bf16, CH_PT>, early order", "scaling": "strong"})
            del xs
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ C5: size sweep, contiguous slices
    qs = [("ActivationSymmetric thr=4", Q.ActivationSymmetricInferableQuantizer(8, [4.0], True)),
          ("ActivationUniform [-1,2.3]", Q.ActivationUniformInferableQuantizer(8, [-1.0], [2.3]))]
    sizes_mb = [1, 4, 16, 64, 256, 1024, 4096, 16384]
    for dt in (torch.float32, torch.bfloat16):
        es = 4 if dt == torch.float32 else 2
        for mb in sizes_mb:
            if mb / 1024 > max_gb:
                continue
            n_total = mb * (1 << 20) // es
            s0, s1 = sharding.shard_range(n_total, world, rank, align=4096)
            n = s1 - s0
            shard_bytes = n * es
            nbuf = max(2, min(16, int(600e6 // max(shard_bytes, 1)) + 1)) if shard_bytes < 300e6 else 2
            xs = [torch.empty(n, device=dev, dtype=dt).uniform_(-50, 50, generator=g) for _ in range(nbuf)]
            for name, q in qs:
                burst = 20 if shard_bytes < 300e6 else 3

                def pass_c5(i):
                    for k in range(burst):
                        q(xs[(i * burst + k) % nbuf])
                ms = timed(pass_c5, reps) / burst
                report(f"C5 {name} {mb} MB {str(dt).split('.')[-1]} (slice per rank, {burst} calls queued)", n_total * es * 2, ms,
                       {"kernel": "fq_affine_kernel<%s, CH_PT>" % ("float" if es == 4 else "bf16"), "scaling": "strong"})
            del xs
            torch.cuda.empty_cache()

    # verification outside the timed regions (the only NCCL traffic of the tool): every rank fake-quantizes ITS shard of
    # a tensor all ranks can regenerate (same seed), the shards are all-gathered over NVLink and compared bit for bit
    # with the unsharded result -- batch sharding (activations) and channel-block sharding (per-channel weights)
    gv = torch.Generator(device=dev).manual_seed(4321)
    full = torch.empty((64 * world, 197, 768), device=dev).normal_(0, 1, generator=gv).bfloat16()
    qa = Q.ActivationSymmetricInferableQuantizer(8, [3.7], True)
    r0, r1 = sharding.shard_batch(full.shape[0], world, rank)
    mine_out = qa(full[r0:r1].contiguous())
    Wf = torch.empty((64 * world, 1024), device=dev).normal_(0, 0.02, generator=gv)
    thr_all = Wf.abs().amax(1).double().cpu().tolist()
    slices, (c0, c1) = sharding.shard_channel_blocks(Wf.shape, 0, world, rank)
    qw_shard = Q.WeightsSymmetricInferableQuantizer(8, thr_all[c0:c1], True, 0)
    w_out = qw_shard(Wf[slices].contiguous())
    if world > 1:
        gathered = torch.empty_like(full)
        dist.all_gather_into_tensor(gathered, mine_out)
        wg = torch.empty_like(Wf)
        dist.all_gather_into_tensor(wg, w_out)
    else:
        gathered, wg = mine_out, w_out
    ok_a = torch.equal(gathered.view(torch.int16), qa(full).view(torch.int16))
    ok_w = torch.equal(wg.view(torch.int32), Q.WeightsSymmetricInferableQuantizer(8, thr_all, True, 0)(Wf).view(torch.int32))
    assert ok_a and ok_w, ("sharded result differs from the unsharded one", ok_a, ok_w)
    if rank == 0 and log is not None:
        print(f"verification: all-gathered batch shards ({tuple(full.shape)} bf16) and channel-block shards ({tuple(Wf.shape)} f32) "
              f"== unsharded results, N={world}", file=log, flush=True)
    rows.append({"config": "verification all-gather (batch shards + channel-block shards) == unsharded", "n_gpus": world, "ok": True})

    # checksums of one small output, gathered with NCCL
    y = qs[0][1](torch.arange(4096, device=dev, dtype=torch.float32) * 0.01 - 20.0)
    chk = torch.tensor([sharding.checksum64(y)], device=dev, dtype=torch.int64)
    if world > 1:
        outs = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(outs, chk)
        assert len({int(o.item()) for o in outs}) == 1, "ranks disagree on a deterministic result"
    return rows, peak


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--max-gb", type=float, default=16.0, help="largest C5 point (input GB, whole job)")
    args = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rows, peak = run_configs(dev, rank, world, args.reps, args.max_gb)
    if rank == 0 and args.json:
        with open(args.json, "w") as f:
            json.dump({"n_gpus": world, "peak_gbs": peak, "rows": rows}, f, indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
