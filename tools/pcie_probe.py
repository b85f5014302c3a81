#!/usr/bin/env python
"""Host-link ceiling for the e2e number: pinned H2D, D2H and both at once (GB/s), plus the host-staging operator
on one large and one small tensor."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mct_quantizers_b200.pytorch import quantizers as Q  # noqa: E402

dev = torch.device("cuda:0")
n = 256 << 20   # 1 GiB of f32
h_in = torch.empty(n, dtype=torch.float32, pin_memory=True).normal_()
h_out = torch.empty(n, dtype=torch.float32, pin_memory=True)
d_a = torch.empty(n, dtype=torch.float32, device=dev)
d_b = torch.empty(n, dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def t(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


gb = n * 4 / 1e9
print(f"H2D alone   : {gb / t(lambda: d_a.copy_(h_in, non_blocking=True)):6.1f} GB/s")
print(f"D2H alone   : {gb / t(lambda: h_out.copy_(d_b, non_blocking=True)):6.1f} GB/s")


def both():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


dt = t(both)
print(f"H2D + D2H   : {gb / dt:6.1f} GB/s per direction, {2 * gb / dt:6.1f} GB/s total  (the e2e ceiling in algorithmic GB/s)")
q = Q.ActivationUniformInferableQuantizer(8, [-2.5], [3.0])
for elems in (n, 64 << 20, 16 << 20, 4 << 20, 1 << 20):
    x = h_in[:elems]
    dt = t(lambda: q(x), reps=5)
    print(f"host-staged fake-quant of {elems * 4 / 1e6:8.1f} MB: {dt * 1e3:8.3f} ms  {2 * elems * 4 / dt / 1e9:6.1f} GB/s algorithmic")

# zero-copy: the streaming kernel reads / writes PINNED host memory directly over PCIe (UVA), no staging, no chunk pipeline
import ctypes  # noqa: E402
from mct_quantizers_b200 import _native  # noqa: E402
lib = _native.load()
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731
h_y = torch.empty(n, dtype=torch.float32, pin_memory=True)
for elems in (n, 64 << 20, 16 << 20, 4 << 20, 1 << 20, 1 << 18):
    def zc():
        rc = lib.mctq_fq_affine_scalar(h_in.data_ptr(), h_y.data_ptr(), None, elems, 0, 0.0215, 116, 0, 255, 0, st())
        assert rc == 0
        torch.cuda.current_stream().synchronize()
    dt = t(zc, reps=5)
    print(f"zero-copy fake-quant of {elems * 4 / 1e6:8.1f} MB: {dt * 1e3:8.3f} ms  {2 * elems * 4 / dt / 1e9:6.1f} GB/s algorithmic")
for u in (2, 8):
    lib.mctq_set_tuning(0, u)
    def zc():
        lib.mctq_fq_affine_scalar(h_in.data_ptr(), h_y.data_ptr(), None, n, 0, 0.0215, 116, 0, 255, 0, st())
        torch.cuda.current_stream().synchronize()
    dt = t(zc, reps=3)
    print(f"zero-copy 1 GiB unroll {u}: {2 * n * 4 / dt / 1e9:6.1f} GB/s algorithmic")
lib.mctq_set_tuning(0, 0)
