#!/usr/bin/env python
"""Where the host time of one small holder call goes (wall clock, tiny tensor, GPU idle): total call, its layers, and the
primitive costs (output allocation, pointer / stream queries, the ctypes launch itself)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mct_quantizers_b200 as mctq  # noqa: E402
from mct_quantizers_b200 import _native, ops  # noqa: E402
from mct_quantizers_b200.pytorch import quantizers as Q  # noqa: E402


def bench(fn, n=20000):
    for _ in range(200):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    dt = (time.perf_counter() - t0) / n
    torch.cuda.synchronize()
    return dt * 1e6


def main():
    dev = torch.device("cuda:0")
    x = torch.randn(1, 3, 32, 32, device=dev)
    q = Q.ActivationUniformInferableQuantizer(8, [-1.0], [2.3])
    h = mctq.PytorchActivationQuantizationHolder(q).to(dev)
    lib = _native.load()
    y = torch.empty_like(x)
    xp, yp, n = x.data_ptr(), y.data_ptr(), x.numel()
    st = torch._C._cuda_getCurrentRawStream(0)
    rows = [
        ("holder(x)  [nn.Module.__call__ -> quantizer -> launch]", lambda: h(x)),
        ("quantizer(x)", lambda: q(x)),
        ("ops.affine_scalar_direct(x, ...)", lambda: ops.affine_scalar_direct(x, 0.0129, 77, 0, 255)),
        ("torch.fake_quantize_per_tensor_affine (ATen, for comparison)", lambda: torch.fake_quantize_per_tensor_affine(x, 0.0129, 77, 0, 255)),
        ("torch.empty_like(x)", lambda: torch.empty_like(x)),
        ("x.data_ptr()", lambda: x.data_ptr()),
        ("x.is_contiguous()", lambda: x.is_contiguous()),
        ("x.numel()", lambda: x.numel()),
        ("x.device.index", lambda: x.device.index),
        ("torch._C._cuda_getCurrentRawStream(0)", lambda: torch._C._cuda_getCurrentRawStream(0)),
        ("torch._C._cuda_getDevice()", lambda: torch._C._cuda_getDevice()),
        ("ops.direct_ok(x)", lambda: ops.direct_ok(x)),
        ("ctypes launch alone: lib.mctq_fq_affine_scalar(...)", lambda: lib.mctq_fq_affine_scalar(xp, yp, None, n, 0, 0.0129, 77, 0, 255, 0, st)),
        ("ctypes call of a trivial function: lib.mctq_launch_count()", lambda: lib.mctq_launch_count()),
    ]
    for name, fn in rows:
        print(f"{name:70s} {bench(fn):7.2f} us")


if __name__ == "__main__":
    main()
