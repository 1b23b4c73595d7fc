"""Host pipeline variants (tensor size x chunk size x staging path) through the tuning hook: e2e GB/s of the headline call.
path 0 = staged pipeline (cudaMemcpyAsync both ways), 1 = one launch straight on the pinned host pointers, -1 = built-in rule;
chunk_MiB -1 = built-in choice.  (An earlier version of this tool also measured mixed forms -- DMA one way, the kernel the
other -- and a head/tail ramp of smaller chunks: both lost at every size and were removed from qd_host.cu.)"""
import itertools
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quantized_distillation_b200 import _native as N  # noqa: E402
from quantized_distillation_b200.distributed import numa_local  # noqa: E402


def main(out):
    lib = N.lib()
    nmax = 1 << 26
    with numa_local(0):
        hx = (torch.randn(nmax) * 0.05).pin_memory()
        hg = torch.randn(nmax).pin_memory()
        hq = torch.zeros(nmax).pin_memory()
        hgo = torch.zeros(nmax).pin_memory()
    rows = []
    for n, slots, chunk_mib, path in itertools.product((1 << 18, 1 << 20, 1 << 22, 1 << 23, 1 << 24, 1 << 26), (3,), (-1, 8, 16), (-1, 0, 1)):
        if path != 0 and chunk_mib != -1:
            continue
        lib.qd_debug_set_tuning(5, slots)
        lib.qd_debug_set_tuning(6, chunk_mib << 18 if chunk_mib > 0 else -1)        # MiB of float32 -> elements
        lib.qd_debug_set_tuning(7, path)

        def step():
            N.check(lib.qd_uniform_fwd_bwd_host(N.ptr(hx), N.ptr(hg), N.ptr(hq), N.ptr(hgo), n, 256, 16, N.BWD_MINMAX, 0))

        step()
        torch.cuda.synchronize()
        iters = 4 if n >= (1 << 26) else 20
        t0 = time.perf_counter()
        for _ in range(iters):
            step()
        dt = (time.perf_counter() - t0) / iters
        rows.append({"n": n, "slots": slots, "chunk_MiB": chunk_mib, "path": path, "ms": round(dt * 1e3, 3), "GBps": round(n * 16 / dt / 1e9, 2)})
        print(rows[-1], flush=True)
    for k in (5, 6, 7):
        lib.qd_debug_set_tuning(k, -1)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    json.dump(rows, open(out, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "e2e_variants.json"))
