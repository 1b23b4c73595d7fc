"""One launch of each helper kernel (inverse scaling, pack, unpack) at 2^28 elements for `ncu --set full`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quantized_distillation_b200 import _native as N  # noqa: E402

lib, sp = N.lib(), N.stream_ptr()
n, bucket = 1 << 28, 256
idx = torch.randint(0, 16, (n,), dtype=torch.uint8, device="cuda")
packed = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
alpha = torch.ones(n // bucket, device="cuda")
beta = torch.zeros(n // bucket, device="cuda")
q = torch.empty(n, device="cuda")
y = torch.rand(n, device="cuda")
pts = torch.linspace(0, 1, 16, device="cuda")
for _ in range(2):
    N.check(lib.qd_pack_indices(N.ptr(idx), N.ptr(packed), n, 4, sp))
    N.check(lib.qd_unpack_dequant_uniform(N.ptr(packed), 4, N.ptr(alpha), N.ptr(beta), N.ptr(q), n, bucket, 16, sp))
    N.check(lib.qd_unpack_dequant_nonuniform(N.ptr(packed), 4, N.ptr(pts), 16, N.ptr(alpha), N.ptr(beta), N.ptr(q), n, bucket, sp))
    N.check(lib.qd_inv_scale_down(N.ptr(y), N.ptr(q), N.ptr(alpha), N.ptr(beta), None, n, bucket, sp))
torch.cuda.synchronize()
