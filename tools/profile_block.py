"""Launches the block-path kernels (rows of 2048 / 8192 / 49152 floats: staged TMA chunk ring) for ncu:
uniform forward, fused forward + min/max backward, centroid op K=4 -- two launches each.

    ncu --set full --clock-control none --import-source on -k regex:"staged_rows_kernel|block_rows_kernel" \
        -o gpurun_out/prof_r2_block python tools/profile_block.py
"""
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
for bucket in (2048, 8192, 49152):
    ws = N.workspace(n, bucket, dev)
    for _ in range(2):
        N.check(lib.qd_uniform_fwd(N.ptr(x), N.ptr(q), None, None, None, None, None, n, bucket, 16, None, 0.0, 0, 0, 0, N.ptr(ws), ws.numel(), sp))
    for _ in range(2):
        N.check(lib.qd_uniform_fwd_bwd(N.ptr(x), N.ptr(g), N.ptr(q), N.ptr(go), n, bucket, 16, N.BWD_MINMAX, N.ptr(ws), ws.numel(), sp))
    for _ in range(2):
        N.check(lib.qd_nonuniform_fwd(N.ptr(x), N.ptr(pts), 4, N.RULE_MIDPOINT, N.ptr(q), N.ptr(idx), None, None, None, n, bucket, None, 0.0,
                                      N.ptr(ws), ws.numel(), sp))
torch.cuda.synchronize()
print("done")
