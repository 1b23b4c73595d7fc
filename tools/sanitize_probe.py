"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quantized_distillation_b200.quantization as Q  # noqa: E402
from quantized_distillation_b200 import codec  # noqa: E402
from quantized_distillation_b200.plan import QuantizationPlan  # noqa: E402
from oracle import quant_oracle as O  # noqa: E402

rng = np.random.default_rng(0)
# row lengths: warp path (256, 100, 1024, 7), staged ring with 64 / 128 / 256 / 512 / 1024-thread CTAs, one and two rows in
# flight, aligned and alternating-unaligned rows (1026, 3002), grid path (None)
for n, bucket in ((1000, 256), (5000, 100), (70001, 1024), (20000, 2048), (150000, 49152), (300001, None), (257, 7),
                  (30000, 1026), (40000, 3002), (100000, 8192), (120000, 20000), (200000, 32768)):
    x = (rng.standard_normal(n) * 0.05).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    q, sf = Q.uniformQuantization(xd, 16, bucket_size=bucket)
    ref, _, st = O.uniform_fwd(x, 16, bucket)
    assert np.array_equal(q.cpu().numpy(), ref), (n, bucket)
    Q.uniformQuantization(xd, 16, bucket_size=bucket, stochastic_rounding=True)
    pts = torch.linspace(0, 1, 4).cuda()
    f = Q.nonUniformQuantization_variable(bucket_size=bucket, pre_process_tensors=True, tensor=xd)
    f.forward(None, pts)
    f.backward(torch.randn(n).cuda())
    Q.nonUniformQuantization(xd, torch.linspace(0, 1, 40).cuda(), bucket_size=bucket)
    if bucket is not None:
        g = Q.uniformQuantization_variable(16, bucket_size=bucket)
        g.forward(xd)
        g.backward(torch.randn(n).cuda())
    sfn = Q.ScalingFunction("linear", False, False, bucket, False)
    sfn.inv_scale_down(sfn.scale_down(xd))
    codec.decode(codec.encode_uniform(xd, 16, bucket))
params = [torch.randn(n).cuda() * 0.05 for n in (5000, 10, 93750, 75, 257, 1)]
plan = QuantizationPlan(params, 16, 256)
plan.save_and_quantize_()
plan.restore_master()
plan.backward_([torch.randn_like(p) for p in params], "complicated")
# round 2: fused optimizer step, long-row plan (bucket None), centroid plan, order statistics, multi-tensor norm, a10 extension
from quantized_distillation_b200.plan import CentroidPlan  # noqa: E402
from quantized_distillation_b200.quantization import help_functions as H  # noqa: E402
from quantized_distillation_b200.quantization import quant_functions as QF  # noqa: E402
plan.save_and_quantize_()
for style in ("none", "truncated", "complicated"):
    plan.fused_step_([torch.randn_like(p) for p in params], style, 1e-2, 0.9, 2e-4, True)
big = [torch.randn(n).cuda() * 0.05 for n in (432, 200000, 49153, 16385, 3)]
lplan = QuantizationPlan(big, 4, None)
lplan.save_and_quantize_()
lplan.restore_master()
lplan.backward_([torch.randn_like(p) for p in big], "truncated")
src = [torch.randn(n).cuda() * 0.05 for n in (5000, 10, 93750, 257)]
pts = [torch.sort(torch.rand(k).cuda())[0] for k in (4, 32, 7, 16)]
cplan = CentroidPlan(src, [torch.empty_like(t) for t in src], pts, 256)
cplan.forward_()
cplan.backward_([torch.randn_like(t) for t in src])
v = torch.rand(100003).cuda()
assert torch.equal(H.order_statistics(v, [0, 7, 50000, 100002]), torch.sort(v)[0][[0, 7, 50000, 100002]])
H.gradient_norms(src)
QF.ALLOW_UNPINNED_SCALING = True
for kind in ("absmax", "absnorm"):
    for bucket in (256, 4096, None):
        Q.uniformQuantization(src[2], 8, type_of_scaling=kind, bucket_size=bucket)
        sfa = Q.ScalingFunction(kind, False, False, bucket, False)
        sfa.inv_scale_down(sfa.scale_down(src[2]))
QF.ALLOW_UNPINNED_SCALING = False
# tiled helper kernels past one tile per thread-group (full-tile fast paths + general tails), 1 / 2 / 4 / 8-bit codes
from quantized_distillation_b200 import _native as N  # noqa: E402
xb = torch.randn(70_001).cuda() * 0.05
for s_levels, bucket in ((2, 256), (4, 100), (16, 256), (256, 1000), (16, None)):
    assert torch.equal(codec.decode(codec.encode_uniform(xb, s_levels, bucket)), Q.uniformQuantization(xb, s_levels, bucket_size=bucket)[0])
sfb = Q.ScalingFunction("linear", False, False, 256, False)
sfb.inv_scale_down(sfb.scale_down(xb))
# host entry points: one launch on pinned host pointers (small), chunked pipeline (pageable memory)
for pinned in (True, False):
    m = 300_001
    hx, hg = torch.randn(m) * 0.05, torch.randn(m)
    hq, hgo = torch.zeros(m), torch.zeros(m)
    if pinned:
        hx, hg, hq, hgo = hx.pin_memory(), hg.pin_memory(), hq.pin_memory(), hgo.pin_memory()
    N.check(N.lib().qd_uniform_fwd_bwd_host(N.ptr(hx), N.ptr(hg), N.ptr(hq), N.ptr(hgo), m, 256, 16, N.BWD_MINMAX, 0))
    assert torch.equal(hq, Q.uniformQuantization(hx.cuda(), 16, bucket_size=256)[0].cpu())
torch.cuda.synchronize()
print("sanitize probe ok")
