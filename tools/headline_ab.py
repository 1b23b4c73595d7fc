"""A/B of the r_b accumulation in the headline kernel (fused uniform forward + min/max backward, 64 Mi floats,
bucket 256) on ONE box, interleaved: variant B = shared lane-sum routine (division mode hoisted, float32 groups),
variant A = one float64 add per element.  200 back-to-back launches per sample, like bench.py.

    python tools/headline_ab.py
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quantized_distillation_b200 import _native as N  # noqa: E402

dev = torch.device("cuda", 0)
lib, sp = N.lib(), N.stream_ptr(dev)
n = 1 << 26
gen = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(n, generator=gen, device=dev) * 0.05
g = torch.randn(n, generator=gen, device=dev)
q, go = torch.empty_like(x), torch.empty_like(g)


def sample(mode, launches=200):
    fn = (lambda: lib.qd_uniform_fwd_bwd(N.ptr(x), N.ptr(g), N.ptr(q), N.ptr(go), n, 256, 16, mode, None, 0, sp)) if mode is not None else \
         (lambda: lib.qd_uniform_bwd(N.ptr(x), N.ptr(g), N.ptr(go), n, 256, 16, N.BWD_MINMAX, None, 0, sp))
    for _ in range(5):
        N.check(fn())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(launches):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / launches * 1e3


out = {"fused_minmax_us": {"A": [], "B": []}, "bwd_minmax_alone_us": {"A": [], "B": []}, "fused_ste_us": []}
for rep in range(4):
    for name, v in (("B", 0), ("A", 1)):
        N.check(lib.qd_debug_set_tuning(3, v))
        out["fused_minmax_us"][name].append(round(sample(N.BWD_MINMAX), 2))
        out["bwd_minmax_alone_us"][name].append(round(sample(None), 2))
    N.check(lib.qd_debug_set_tuning(3, -1))
    out["fused_ste_us"].append(round(sample(N.BWD_STE), 2))
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "headline_ab.json"), "w"), indent=1)
