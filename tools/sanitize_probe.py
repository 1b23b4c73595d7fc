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
for n, bucket in ((1000, 256), (5000, 100), (70001, 1024), (20000, 2048), (150000, 49152), (300001, None), (257, 7)):
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
torch.cuda.synchronize()
print("sanitize probe ok")
