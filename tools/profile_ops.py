"""Launches each hot kernel a few times on a 64 Mi-float tensor so that one ncu
invocation can capture all of them (see profiles/README.md for the command)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quantized_distillation_b200 import _native as N  # noqa: E402

n = int(os.environ.get("QD_PROFILE_N", 1 << 26))
reps = int(os.environ.get("QD_PROFILE_REPS", 3))
dev = torch.device("cuda", 0)
lib, sp = N.lib(), N.stream_ptr(dev)
gen = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(n, generator=gen, device=dev) * 0.05
g = torch.randn(n, generator=gen, device=dev)
q, go = torch.empty_like(x), torch.empty_like(g)
idx = torch.empty(n, dtype=torch.uint8, device=dev)
ws = N.workspace(n, 0, dev)
pts = torch.linspace(0, 1, 4, device=dev)
pts16 = torch.sort(torch.rand(16, device=dev))[0]
rows = N.geometry(n, 256)[0]
alpha = torch.empty(rows, device=dev)
beta = torch.empty(rows, device=dev)
gp = torch.empty(4, device=dev)
ops = [
    ("uniform_fwd", lambda: lib.qd_uniform_fwd(N.ptr(x), N.ptr(q), None, None, None, None, None, n, 256, 16, None, 0.0, 0, 0, 0, N.ptr(ws), ws.numel(), sp)),
    ("fwd_bwd_ste", lambda: lib.qd_uniform_fwd_bwd(N.ptr(x), N.ptr(g), N.ptr(q), N.ptr(go), n, 256, 16, N.BWD_STE, N.ptr(ws), ws.numel(), sp)),
    ("fwd_bwd_trunc", lambda: lib.qd_uniform_fwd_bwd(N.ptr(x), N.ptr(g), N.ptr(q), N.ptr(go), n, 256, 16, N.BWD_TRUNCATED, N.ptr(ws), ws.numel(), sp)),
    ("fwd_bwd_minmax", lambda: lib.qd_uniform_fwd_bwd(N.ptr(x), N.ptr(g), N.ptr(q), N.ptr(go), n, 256, 16, N.BWD_MINMAX, N.ptr(ws), ws.numel(), sp)),
    ("nonuniform_fwd_k4", lambda: lib.qd_nonuniform_fwd(N.ptr(x), N.ptr(pts), 4, N.RULE_MIDPOINT, N.ptr(q), N.ptr(idx), None, N.ptr(alpha), N.ptr(beta), n, 256, None, 0.0, N.ptr(ws), ws.numel(), sp)),
    ("nonuniform_fwd_k16_nearest", lambda: lib.qd_nonuniform_fwd(N.ptr(x), N.ptr(pts16), 16, N.RULE_NEAREST, N.ptr(q), N.ptr(idx), None, N.ptr(alpha), N.ptr(beta), n, 256, None, 0.0, N.ptr(ws), ws.numel(), sp)),
    ("bwd_minmax_alone", lambda: lib.qd_uniform_bwd(N.ptr(x), N.ptr(g), N.ptr(go), n, 256, 16, N.BWD_MINMAX, N.ptr(ws), ws.numel(), sp)),
    ("nonuniform_bwd_k4", lambda: lib.qd_nonuniform_bwd(N.ptr(g), N.ptr(idx), None, N.ptr(alpha), 4, N.ptr(gp), n, 256, N.ptr(ws), ws.numel(), sp)),
    ("uniform_fwd_none", lambda: lib.qd_uniform_fwd(N.ptr(x), N.ptr(q), None, None, None, None, None, n, 0, 16, None, 0.0, 0, 0, 0, N.ptr(ws), ws.numel(), sp)),
]
only = os.environ.get("QD_PROFILE_ONLY")
for name, fn in ops:
    if only and name not in only.split(","):
        continue
    for _ in range(reps):
        N.check(fn())
    torch.cuda.synchronize()
    print("ran", name)
