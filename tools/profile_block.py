"""Two launches of the mid-row paths for ncu (warp two-pass at 2048, TMA-staged CTA at 16384)."""
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
q = torch.empty_like(x)
for bucket in (2048, 16384):
    ws = N.workspace(n, bucket, dev)
    for _ in range(2):
        N.check(lib.qd_uniform_fwd(N.ptr(x), N.ptr(q), None, None, None, None, None, n, bucket, 16, None, 0.0, 0, 0, 0, N.ptr(ws), ws.numel(), sp))
torch.cuda.synchronize()
