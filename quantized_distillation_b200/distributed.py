"""Data-parallel plumbing: one process per GPU, ``DistributedDataParallel`` with NCCL
gradient all-reduce over NVLink / NVSwitch.

The reference's multi-GPU story is single-process ``nn.DataParallel``
(cifar10_wideResNet.py:68-69, 97, 117); the quantization op itself needs no
collective -- every rank quantizes its full replica, and because the kernels are
deterministic the replicas stay bit-identical.  The only collective in a step is
DDP's bucketed gradient all-reduce, overlapped with backward.  The gradient
fix-up kernels run after it, on the already reduced gradients, like the reference
runs them on its single reduced copy (conv_forward_model.py:315)."""
from __future__ import annotations

import os
from collections import OrderedDict

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def init_distributed(backend=None):
    """Initialises the default process group from the torchrun environment.
    Returns (world, rank, device)."""
    world, rank, local_rank = env_world()
    use_cuda = torch.cuda.is_available()
    if use_cuda:
        torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank) if use_cuda else torch.device("cpu")
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if use_cuda else "gloo")
        kwargs = {"device_id": device} if backend == "nccl" else {}
        dist.init_process_group(backend, **kwargs)
    return world, rank, device


def wrap_ddp(model, device):
    """DDP wrapper.  Parameters are only ever modified IN PLACE by the quantization
    plan, so the parameter objects DDP registered its hooks on stay the live ones."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return model
    from torch.nn.parallel import DistributedDataParallel as DDP
    if device.type == "cuda":
        return DDP(model, device_ids=[device.index], output_device=device.index, broadcast_buffers=True,
                   gradient_as_bucket_view=True)
    return DDP(model)


def shard_batches(batches, rank, world):
    """Batch-dimension partitioning of a list of (inputs, labels): rank r gets the
    r-th contiguous slice of every global batch (the reference's DataParallel
    scatter, cifar10_wideResNet.py:142)."""
    if world == 1:
        return batches
    out = []
    for x, y in batches:
        if x.size(0) % world != 0:
            raise ValueError("Batch size: {} must be a multiple of the number of gpus: {}".format(x.size(0), world))
        per = x.size(0) // world
        out.append((x[rank * per:(rank + 1) * per], y[rank * per:(rank + 1) * per]))
    return out


def convert_state_dict_to_data_parallel(state_dict):
    """Adds the ``module.`` prefix a (Distributed)DataParallel wrapper expects
    (reference: helpers/functions.py:179-189)."""
    return OrderedDict((k if k.startswith("module.") else "module." + k, v) for k, v in state_dict.items())


def convert_state_dict_from_data_parallel(state_dict):
    """Strips the ``module.`` prefix (reference: helpers/functions.py:191-205)."""
    return OrderedDict((k[len("module."):] if k.startswith("module.") else k, v) for k, v in state_dict.items())


def max_over_ranks(value: float, device) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
