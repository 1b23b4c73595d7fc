"""Training legs of bench.py under torch's convolution settings: TF32 (torch's default for cuDNN convolutions) on / off,
cudnn.benchmark on / off.  Decides what bench.py pins for its train legs."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
rows = []
for kind, steps in (("student", 40), ("wrn", 12)):
    for tf32 in (True, False):
        for autotune in (False, True):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = False
            torch.backends.cudnn.benchmark = autotune
            r = bench.run_train_leg(kind, 1, 0, dev, steps, 8, graph=True, fused=True)
            rows.append({"kind": kind, "conv_tf32": tf32, "cudnn_benchmark": autotune, "steps_per_s": r["steps_per_s"], "captured": r.get("captured")})
            print(rows[-1], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "train_precision_probe.json"), "w"), indent=1)
