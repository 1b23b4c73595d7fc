"""Throughput of the CTA-per-row (TMA-staged) path for several row lengths."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quantized_distillation_b200 import _native as N  # noqa: E402

dev = torch.device("cuda", 0)
lib, sp = N.lib(), N.stream_ptr(dev)
n = 1 << 26
x = torch.randn(n, device=dev) * 0.05
g = torch.randn(n, device=dev)
q, go = torch.empty_like(x), torch.empty_like(g)
idx = torch.empty(n, dtype=torch.uint8, device=dev)
pts = torch.linspace(0, 1, 4, device=dev)


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


for bucket in (512, 1024, 2048, 4096, 8192, 16384, 32768, 49152):
    ws = N.workspace(n, bucket, dev)
    t_f = timed(lambda: N.check(lib.qd_uniform_fwd(N.ptr(x), N.ptr(q), None, None, None, None, None, n, bucket, 16, None, 0.0, 0, 0, 0, N.ptr(ws), ws.numel(), sp)))
    t_b = timed(lambda: N.check(lib.qd_uniform_fwd_bwd(N.ptr(x), N.ptr(g), N.ptr(q), N.ptr(go), n, bucket, 16, N.BWD_MINMAX, N.ptr(ws), ws.numel(), sp)))
    t_n = timed(lambda: N.check(lib.qd_nonuniform_fwd(N.ptr(x), N.ptr(pts), 4, N.RULE_MIDPOINT, N.ptr(q), N.ptr(idx), None, None, None, n, bucket, None, 0.0, N.ptr(ws), ws.numel(), sp)))
    print(f"bucket {bucket:6d}: uniform fwd {t_f*1e6:8.1f} us {n*8/t_f/1e9:7.0f} GB/s | fused minmax {t_b*1e6:8.1f} us {n*16/t_b/1e9:7.0f} GB/s | nonuniform K4 {t_n*1e6:8.1f} us {n*9/t_n/1e9:7.0f} GB/s")
