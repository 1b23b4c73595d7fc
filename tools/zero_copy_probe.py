"""Does the fused kernel run at PCIe rate straight on pinned host memory (UVA: cudaHostAlloc'd buffers are device-addressable)?
Compares the staged host pipeline (qd_uniform_fwd_bwd_host) with ONE launch of the resident-tensor entry point on host pointers."""
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
    sp = N.stream_ptr()
    res = []
    for n in (1 << 22, 1 << 24, 1 << 26):
        with numa_local(0):
            hx = (torch.randn(n) * 0.05).pin_memory()
            hg = torch.randn(n).pin_memory()
            hq = torch.zeros(n).pin_memory()
            hgo = torch.zeros(n).pin_memory()
            rq = torch.zeros(n).pin_memory()
            rgo = torch.zeros(n).pin_memory()
        ws = N.workspace(n, 256, torch.device("cuda", 0))

        def staged():
            N.check(lib.qd_uniform_fwd_bwd_host(N.ptr(hx), N.ptr(hg), N.ptr(rq), N.ptr(rgo), n, 256, 16, N.BWD_MINMAX, 0))

        def direct():
            N.check(lib.qd_uniform_fwd_bwd(N.ptr(hx), N.ptr(hg), N.ptr(hq), N.ptr(hgo), n, 256, 16, N.BWD_MINMAX, N.ptr(ws), ws.numel(), sp))
            torch.cuda.synchronize()

        dx, dg = hx.cuda(), hg.cuda()
        dq, dgo = torch.empty_like(dx), torch.empty_like(dx)

        def read_host():          # host in, device out: the PCIe read side alone (8 B/elt over the link)
            N.check(lib.qd_uniform_fwd_bwd(N.ptr(hx), N.ptr(hg), N.ptr(dq), N.ptr(dgo), n, 256, 16, N.BWD_MINMAX, N.ptr(ws), ws.numel(), sp))
            torch.cuda.synchronize()

        def write_host():         # device in, host out: the PCIe write side alone
            N.check(lib.qd_uniform_fwd_bwd(N.ptr(dx), N.ptr(dg), N.ptr(hq), N.ptr(hgo), n, 256, 16, N.BWD_MINMAX, N.ptr(ws), ws.numel(), sp))
            torch.cuda.synchronize()

        row = {"n": n}
        for name, fn in (("staged", staged), ("direct", direct), ("read_host", read_host), ("write_host", write_host)):
            fn()
            torch.cuda.synchronize()
            iters = 8 if n >= (1 << 26) else 30
            t0 = time.perf_counter()
            for _ in range(iters):
                fn()
            dt = (time.perf_counter() - t0) / iters
            row[name + "_ms"] = round(dt * 1e3, 4)
            row[name + "_GBps"] = round(n * 16 / dt / 1e9, 2)
        row["identical"] = bool(torch.equal(hq, rq) and torch.equal(hgo, rgo))
        res.append(row)
        print(row, flush=True)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "zero_copy_probe.json"))
