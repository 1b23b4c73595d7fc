"""Generates tests/golden/reference_host_logic.json: the reference's HOST-side helpers of the hot path's
callers, run unmodified (antspy/quantized_distillation mounted at /root/reference), on seeded inputs:

  * quantization.help_functions.assign_bits_automatically      (help_functions.py:97-138)
  * quantization.help_functions.huffman_encode                 (:157-172)
  * quantization.help_functions.create_bucket_tensor           (:67-94)
  * cnn_models.help_fun.LearningRateScheduler                  (help_fun.py:172-260)
  * helpers.functions.get_size_reduction                       (functions.py:216-224)
  * helpers.functions.convert_state_dict_{to,from}_data_parallel (:179-205)

Run in the build container only:  python tests/golden/make_golden_host.py
"""
import json
import os
import random
import sys
import warnings
from collections import OrderedDict

import torch

REF = "/root/reference"
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")
import cnn_models.help_fun as HF  # noqa: E402
import helpers.functions as MF  # noqa: E402
import quantization.help_functions as QH  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_host_logic.json")


def main():
    rnd = random.Random(20240924)
    out = {}

    cases = []
    for _ in range(200):
        k = rnd.randint(1, 24)
        norms = [rnd.random() * 10 ** rnd.uniform(-3, 2) for _ in range(k)]
        is_point = rnd.random() < 0.5
        if rnd.random() < 0.5:
            init = rnd.randint(2, 16)
        else:
            init = [rnd.randint(2, 33) for _ in range(k)]
        cases.append({"norms": norms, "initial": init, "input_is_point": is_point,
                      "result": QH.assign_bits_automatically(list(norms), init if isinstance(init, int) else list(init), input_is_point=is_point)})
    out["assign_bits_automatically"] = cases

    cases = []
    for _ in range(60):
        k = rnd.randint(1, 40)
        w = [rnd.random() ** rnd.choice((1, 3)) + 1e-9 for _ in range(k)]
        tot = sum(w)
        freq = {i: v / tot for i, v in enumerate(w)}
        if rnd.random() < 0.3:                       # equal weights: exercises the tie-breaking of the heap
            freq = {i: 1.0 / k for i in range(k)}
        cases.append({"freq": [[s, f] for s, f in freq.items()], "code": [[s, c] for s, c in QH.huffman_encode(dict(freq))]})
    out["huffman_encode"] = cases

    cases = []
    for n in (1, 5, 17, 256, 257, 1000):
        for b in (None, 1, 4, 17, 256, 300):
            for fill in ("last", "nan"):
                t = torch.arange(n, dtype=torch.float32) * 0.5 - 3
                r = QH.create_bucket_tensor(t.clone(), b, fill_values=fill)
                cases.append({"n": n, "bucket": b, "fill": fill, "shape": list(r.shape),
                              "tail": [None if v != v else v for v in r.reshape(-1)[-8:].tolist()]})
    out["create_bucket_tensor"] = cases

    cases = []
    for style in ("generic", "cifar100", "imagenet", "quant_points_cifar100"):
        for trial in range(4):
            lr0 = rnd.choice((0.1, 0.001, 1e-5, 1.0))
            sch = HF.LearningRateScheduler(lr0, style)
            err, trace = 0.9, []
            for epoch in range(220):
                # a validation error that improves, stalls for long stretches, and sometimes gets worse
                if rnd.random() < (0.5 if epoch < 40 else 0.08):
                    err = max(0.01, err - rnd.random() * 0.02)
                elif rnd.random() < 0.2:
                    err = err + rnd.random() * 0.005
                lr, stop = sch.update_learning_rate(epoch, err)
                trace.append([epoch, err, lr, bool(stop)])
                if stop:
                    break
            cases.append({"style": style, "initial": lr0, "trace": trace})
    out["learning_rate_scheduler"] = cases

    out["get_size_reduction"] = [{"bits": b, "bucket": k, "full": f, "result": MF.get_size_reduction(b, bucket_size=k, full_precision_bits=f)}
                                 for b in (1, 2, 2.7, 4, 8) for k in (None, 64, 256) for f in (32, 16)]

    sd = OrderedDict((k, 0) for k in ("conv1.weight", "module.bn.bias", "layer.0.module.weight", "linear.bias"))
    wrapped = MF.convert_state_dict_to_data_parallel(OrderedDict(sd))
    try:
        MF.convert_state_dict_from_data_parallel(OrderedDict(sd))
        unprefixed = "accepted"
    except ValueError:
        unprefixed = "ValueError"
    out["state_dict_prefix"] = {"keys": list(sd), "to": list(wrapped), "from_of_to": list(MF.convert_state_dict_from_data_parallel(wrapped)),
                                "from_with_unprefixed_key": unprefixed}

    with open(OUT, "w") as f:
        json.dump(out, f)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB")


if __name__ == "__main__":
    main()
