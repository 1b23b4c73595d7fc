"""Launch list of ONE data-parallel training step (config 2, FlatDataParallel, whole step captured in a
CUDA graph), taken with torch.profiler (CUPTI; nsys is not in the image): every kernel of the step with
its duration, the NCCL kernel named, its exposed time (the step runs on one stream, so a collective that
does not overlap compute is exposed for its whole duration).

    torchrun --nproc-per-node N tools/profile_ddp_step.py [--out gpurun_out/ddp_step]
"""
import argparse
import csv
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ddp_step"))
    ap.add_argument("--kind", default="student", choices=["student", "wrn"])
    ap.add_argument("--graph", type=int, default=1)
    args = ap.parse_args()
    os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
    import bench
    from quantized_distillation_b200 import distributed as D
    from quantized_distillation_b200.cnn_models import conv_forward_model as cfm
    from quantized_distillation_b200.cnn_models import help_fun as hf
    from torch.profiler import ProfilerActivity, profile

    world, rank, dev = D.init_distributed()
    st = bench.leg_settings(args.kind, world)
    student, teacher = bench.build_models(args.kind, dev)
    model = D.wrap_data_parallel(student, dev)
    warm, prof_steps = 10, 3
    total = warm + prof_steps
    data = hf.synthetic_cifar_loader(total, st["per_gpu_batch"], seed=100 + rank)
    prof = profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA])
    marks = {}

    def hook(i, loss):
        if i == warm:
            torch.cuda.synchronize(dev)
            prof.start()
            marks["t0"] = torch.cuda.Event(enable_timing=True)
            marks["t0"].record()
        if i == total:
            marks["t1"] = torch.cuda.Event(enable_timing=True)
            marks["t1"].record()
            torch.cuda.synchronize(dev)
            prof.stop()

    info = cfm.train_model_quantized(model, data, data, numBits=st["bits"], bucket_size=256, use_distillation_loss=True,
                                     teacher_model=teacher, epochs_to_train=1, print_every=1, verbose=False, evaluate=False,
                                     max_steps=total, step_hook=hook, cuda_graph_step=bool(args.graph), **st["kw"])[1]
    step_ms = marks["t0"].elapsed_time(marks["t1"]) / prof_steps
    if rank == 0:
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.name and "Memcpy" not in e.name
               and "Memset" not in e.name]
        evs.sort(key=lambda e: e.time_range.start)
        per_step = len(evs) // prof_steps
        one = evs[per_step:2 * per_step] if per_step else evs            # the middle step
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        tag = f"{args.out}_{args.kind}_n{world}"
        with open(tag + ".csv", "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["order", "kernel", "start_us_in_step", "duration_us"])
            t0 = one[0].time_range.start if one else 0
            for i, e in enumerate(one):
                w.writerow([i, e.name[:160], round(e.time_range.start - t0, 2), round(e.time_range.end - e.time_range.start, 2)])
        nccl = [e for e in one if "nccl" in e.name.lower()]
        ours = [e for e in one if e.name.startswith("qd::") or "qd::" in e.name]
        busy = sum(e.time_range.end - e.time_range.start for e in one)
        # exposed = the part of each NCCL kernel during which no other kernel of the step is running
        others = sorted((e.time_range.start, e.time_range.end) for e in one if "nccl" not in e.name.lower())

        def uncovered(a, b):
            left = b - a
            for s0, s1 in others:
                lo, hi = max(a, s0), min(b, s1)
                if hi > lo:
                    left -= hi - lo
            return max(left, 0.0)

        exposed = sum(uncovered(e.time_range.start, e.time_range.end) for e in nccl)
        summary = {"kind": args.kind, "n_gpus": world, "cuda_graph_step": bool(info.get("cuda_graph_step", False)),
                   "step_ms_events": round(step_ms, 4), "kernels_per_step": len(one),
                   "kernel_busy_us_per_step": round(busy, 1),
                   "nccl_kernels": [{"name": e.name[:120], "duration_us": round(e.time_range.end - e.time_range.start, 2)} for e in nccl],
                   "nccl_total_us_per_step": round(sum(e.time_range.end - e.time_range.start for e in nccl), 2),
                   "nccl_exposed_us_per_step": round(exposed, 2),
                   "nccl_exposed_share_of_step": round(exposed / (step_ms * 1e3), 4),
                   "gradient_buckets": len(getattr(model, "_buckets", [])),
                   "qd_kernels": [{"name": e.name[:120], "duration_us": round(e.time_range.end - e.time_range.start, 2)} for e in ours],
                   "note": "exposed = NCCL kernel time not covered by any other kernel of the step (one bucket: nothing can "
                           "overlap it; several buckets: reduced on a side stream while the backward pass continues)"}
        with open(tag + ".json", "w") as f:
            json.dump(summary, f, indent=1)
        print(json.dumps(summary))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
