import cProfile, pstats, sys, io, torch
sys.path.insert(0, ".")
import quantized_distillation_b200.quantization as Q
x = torch.randn(5000, device="cuda")
for _ in range(100): Q.uniformQuantization(x, 16, bucket_size=256, modify_in_place=True)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(3000): Q.uniformQuantization(x, 16, bucket_size=256, modify_in_place=True)
t1 = time.perf_counter(); torch.cuda.synchronize()
from quantized_distillation_b200 import _native as N
print(f"uniformQuantization(5000 floats, bucket 256, in place): {(t1 - t0) / 3000 * 1e6:.1f} us per call on the host "
      f"({'compiled front door' if N.fast() is not None else 'ctypes path'}, no profiler)")
_saved, N._fast = N._fast, None
for _ in range(100): Q.uniformQuantization(x, 16, bucket_size=256, modify_in_place=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3000): Q.uniformQuantization(x, 16, bucket_size=256, modify_in_place=True)
t1 = time.perf_counter(); torch.cuda.synchronize(); N._fast = _saved
print(f"same through ctypes: {(t1 - t0) / 3000 * 1e6:.1f} us per call")
pr = cProfile.Profile(); pr.enable()
for _ in range(3000): Q.uniformQuantization(x, 16, bucket_size=256, modify_in_place=True)
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18); print(s.getvalue()[:3500])
