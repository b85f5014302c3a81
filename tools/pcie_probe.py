#!/usr/bin/env python
"""Host-link ceiling for the e2e number: pinned H2D, D2H and both at once (GB/s), on one GPU or on N GPUs CONCURRENTLY:

    python tools/pcie_probe.py                                             # one GPU (+ the host-staging operator sweep)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/pcie_probe.py [--json out.json]

Under torchrun every rank drives its own GPU with its own pinned buffers; all ranks start each phase together (barrier) and
the aggregate is the total bytes / the slowest rank's time.  "H2D + D2H total" is the ceiling of bench.py's `e2e` value in
algorithmic GB/s (one input byte up and one output byte down per two algorithmic bytes).
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--mb", type=int, default=1024)
    args = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.mb * (1 << 20) // 4
    h_in = torch.empty(n, dtype=torch.float32, pin_memory=True).normal_()
    h_out = torch.empty(n, dtype=torch.float32, pin_memory=True)
    d_a = torch.empty(n, dtype=torch.float32, device=dev)
    d_b = torch.empty(n, dtype=torch.float32, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def t(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = torch.tensor([(time.perf_counter() - t0) / reps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return dt.item()

    def both():
        with torch.cuda.stream(s1):
            d_a.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_b, non_blocking=True)

    gb = n * 4 / 1e9 * world
    res = {"n_gpus": world, "mb_per_gpu": args.mb, "host_cores": os.cpu_count()}
    res["h2d_alone_GBs"] = round(gb / t(lambda: d_a.copy_(h_in, non_blocking=True)), 1)
    res["d2h_alone_GBs"] = round(gb / t(lambda: h_out.copy_(d_b, non_blocking=True)), 1)
    dt = t(both)
    res["both_per_direction_GBs"] = round(gb / dt, 1)
    res["both_total_GBs"] = round(2 * gb / dt, 1)
    if rank == 0:
        print(f"N = {world} GPU(s), {args.mb} MB per GPU and direction, aggregate over all ranks (slowest rank's time):")
        print(f"H2D alone   : {res['h2d_alone_GBs']:7.1f} GB/s")
        print(f"D2H alone   : {res['d2h_alone_GBs']:7.1f} GB/s")
        print(f"H2D + D2H   : {res['both_per_direction_GBs']:7.1f} GB/s per direction, {res['both_total_GBs']:7.1f} GB/s total  "
              f"(the e2e ceiling in algorithmic GB/s)", flush=True)

    # the host-staging operator itself, every rank on its own tensors
    from mct_quantizers_b200.pytorch import quantizers as Q
    q = Q.ActivationUniformInferableQuantizer(8, [-2.5], [3.0])
    rows = []
    for elems in (n, 64 << 20, 16 << 20, 4 << 20, 1 << 20):
        if elems > n:
            continue
        x = h_in[:elems]
        dt = t(lambda: q(x), reps=5)
        rows.append({"mb": elems * 4 / 1e6, "ms": round(dt * 1e3, 3), "GBs": round(2 * elems * 4 * world / dt / 1e9, 1)})
        if rank == 0:
            print(f"host-staged fake-quant of {elems * 4 / 1e6:8.1f} MB per GPU: {dt * 1e3:8.3f} ms  {rows[-1]['GBs']:7.1f} GB/s algorithmic (aggregate)", flush=True)
    res["host_staged"] = rows
    if world == 1:
        # zero-copy: the streaming kernel reads / writes PINNED host memory directly over PCIe (UVA), no staging, no chunk pipeline
        import ctypes
        from mct_quantizers_b200 import _native
        lib = _native.load()
        st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731
        h_y = torch.empty(n, dtype=torch.float32, pin_memory=True)
        for elems in (n, 64 << 20, 16 << 20, 4 << 20, 1 << 20, 1 << 18):
            if elems > n:
                continue

            def zc():
                rc = lib.mctq_fq_affine_scalar(h_in.data_ptr(), h_y.data_ptr(), None, elems, 0, 0.0215, 116, 0, 255, 0, st())
                assert rc == 0
                torch.cuda.current_stream().synchronize()
            dt = t(zc, reps=5)
            print(f"zero-copy fake-quant of {elems * 4 / 1e6:8.1f} MB: {dt * 1e3:8.3f} ms  {2 * elems * 4 / dt / 1e9:6.1f} GB/s algorithmic")
    if rank == 0 and args.json:
        with open(args.json, "w") as f:
            json.dump(res, f, indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
