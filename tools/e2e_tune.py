"""e2e (host buffers) throughput of qd_uniform_fwd_bwd_host for the current QD_HOST_* settings."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quantized_distillation_b200 import _native as N  # noqa: E402

n = 1 << 26
hx = (torch.randn(n) * 0.05).pin_memory()
hg = torch.randn(n).pin_memory()
hq = torch.empty(n).pin_memory()
hgo = torch.empty(n).pin_memory()
lib = N.lib()
for _ in range(2):
    N.check(lib.qd_uniform_fwd_bwd_host(N.ptr(hx), N.ptr(hg), N.ptr(hq), N.ptr(hgo), n, 256, 16, N.BWD_MINMAX, 0))
t0 = time.perf_counter()
for _ in range(6):
    N.check(lib.qd_uniform_fwd_bwd_host(N.ptr(hx), N.ptr(hg), N.ptr(hq), N.ptr(hgo), n, 256, 16, N.BWD_MINMAX, 0))
dt = (time.perf_counter() - t0) / 6
print(f"slots={os.environ.get('QD_HOST_SLOTS', '3')} chunk={os.environ.get('QD_HOST_CHUNK_ELEMS', str(4 << 20))}: "
      f"{dt * 1e3:.2f} ms/step, {n * 16 / dt / 1e9:.1f} GB/s algorithmic, {n * 8 / dt / 1e9:.1f} GB/s per PCIe direction")
