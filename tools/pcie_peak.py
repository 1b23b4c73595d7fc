"""Raw host<->device copy rates of this box (pinned, NUMA-local buffers): H2D alone, D2H alone, both directions at once.
The e2e leg of bench.py moves 8 B/elt each way concurrently; this is the ceiling it is judged against."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quantized_distillation_b200.distributed import numa_local  # noqa: E402


def main(out=None, mb=1024, iters=8):
    dev = torch.device("cuda", 0)
    n = mb << 20
    with numa_local(0):
        h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
        h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
        h_in.fill_(1)
        h_out.fill_(0)
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.ones(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def h2d():
        with torch.cuda.stream(s1):
            d_a.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_b, non_blocking=True)

    def both():
        h2d()
        d2h()

    import time

    def wall(fn):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / iters

    res = {"buffer_MiB": mb,
           "h2d_GBps": round(n / wall(h2d) / 1e9, 1),
           "d2h_GBps": round(n / wall(d2h) / 1e9, 1)}
    t = wall(both)
    res["both_each_GBps"] = round(n / t / 1e9, 1)
    res["both_sum_GBps"] = round(2 * n / t / 1e9, 1)
    print(json.dumps(res))
    if out:
        os.makedirs(os.path.dirname(out), exist_ok=True)
        json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "pcie_peak.json"))
