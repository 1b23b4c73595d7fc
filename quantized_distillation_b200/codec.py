"""Packed integer codec and size accounting.

The reference's deliverable is a *small* model, but it only ever computes the size it would
have (helpers/functions.py:216-262: Huffman mean bit length x parameter count + 8 bytes per
bucket); the weights themselves stay float32.  Here a quantized tensor can actually be stored:
bit-packed codes (1/2/4/8 bits) + (alpha, beta) per bucket, produced and decoded on the GPU,
and decoding reproduces the fake-quantized float tensor bit for bit."""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

from . import _native as N
from .quantization import help_functions as qhf


def bits_for(levels: int) -> int:
    b = max(1, math.ceil(math.log2(levels)))
    return 1 if b <= 1 else 2 if b <= 2 else 4 if b <= 4 else 8


@dataclass
class PackedTensor:
    packed: torch.Tensor        # uint8, ceil(n*bits/8) bytes
    alpha: torch.Tensor         # float32 [rows]
    beta: torch.Tensor          # float32 [rows]
    shape: torch.Size
    bits: int
    levels: int                 # uniform: s; non-uniform: number of points
    bucket_size: object
    points: object = None       # centroid table for the non-uniform codec

    @property
    def nbytes(self) -> int:
        extra = 0 if self.points is None else self.points.numel() * 4
        return self.packed.numel() + (self.alpha.numel() + self.beta.numel()) * 4 + extra


def _rows(n, bucket_size):
    return N.geometry(n, 0 if bucket_size is None else int(bucket_size))[0]


def encode_uniform(tensor: torch.Tensor, s: int, bucket_size=None) -> PackedTensor:
    """uniformQuantization (quant_functions.py:155-194) straight to packed codes: the float
    fake-quantized tensor is never written."""
    N.require_cuda()
    if s > 256:
        raise ValueError("the packed codec stores at most 8 bits per weight")
    x = tensor.detach().cuda().contiguous().float()
    n = x.numel()
    b = 0 if bucket_size is None else int(bucket_size)
    rows = _rows(n, bucket_size)
    alpha = torch.empty(rows, device=x.device)
    beta = torch.empty(rows, device=x.device)
    idx = torch.empty(n, dtype=torch.uint8, device=x.device)
    ws = N.workspace(n, b, x.device)
    sp = N.stream_ptr(x.device)
    N.check(N.lib().qd_uniform_fwd(N.ptr(x), None, N.ptr(idx), N.ptr(alpha), N.ptr(beta), None, None, n, b, int(s), None, 0.0,
                                   0, 0, 0, N.ptr(ws), ws.numel(), sp))
    bits = bits_for(s)
    packed = torch.empty((n * bits + 7) // 8, dtype=torch.uint8, device=x.device)
    N.check(N.lib().qd_pack_indices(N.ptr(idx), N.ptr(packed), n, bits, sp))
    return PackedTensor(packed, alpha, beta, tensor.shape, bits, int(s), bucket_size)


def encode_nonuniform(tensor: torch.Tensor, points, bucket_size=None, rule="nearest") -> PackedTensor:
    N.require_cuda()
    x = tensor.detach().cuda().contiguous().float()
    pts = torch.as_tensor(points, dtype=torch.float32).detach().to(x.device).contiguous()
    n = x.numel()
    b = 0 if bucket_size is None else int(bucket_size)
    rows = _rows(n, bucket_size)
    alpha = torch.empty(rows, device=x.device)
    beta = torch.empty(rows, device=x.device)
    idx = torch.empty(n, dtype=torch.uint8, device=x.device)
    ws = N.workspace(n, b, x.device)
    sp = N.stream_ptr(x.device)
    N.check(N.lib().qd_nonuniform_fwd(N.ptr(x), N.ptr(pts), pts.numel(), N.RULE_MIDPOINT if rule == "midpoint" else N.RULE_NEAREST,
                                      None, N.ptr(idx), None, N.ptr(alpha), N.ptr(beta), n, b, None, 0.0, N.ptr(ws), ws.numel(), sp))
    bits = bits_for(pts.numel())
    packed = torch.empty((n * bits + 7) // 8, dtype=torch.uint8, device=x.device)
    N.check(N.lib().qd_pack_indices(N.ptr(idx), N.ptr(packed), n, bits, sp))
    return PackedTensor(packed, alpha, beta, tensor.shape, bits, pts.numel(), bucket_size, points=pts)


def decode(pt: PackedTensor) -> torch.Tensor:
    """Packed codes -> the fake-quantized float32 tensor (bit-identical to the fused op)."""
    N.require_cuda()
    n = 1
    for d in pt.shape:
        n *= int(d)
    out = torch.empty(n, dtype=torch.float32, device=pt.packed.device)
    b = 0 if pt.bucket_size is None else int(pt.bucket_size)
    sp = N.stream_ptr(out.device)
    if pt.points is None:
        N.check(N.lib().qd_unpack_dequant_uniform(N.ptr(pt.packed), pt.bits, N.ptr(pt.alpha), N.ptr(pt.beta), N.ptr(out), n, b,
                                                  pt.levels, sp))
    else:
        N.check(N.lib().qd_unpack_dequant_nonuniform(N.ptr(pt.packed), pt.bits, N.ptr(pt.points), pt.points.numel(),
                                                     N.ptr(pt.alpha), N.ptr(pt.beta), N.ptr(out), n, b, sp))
    return out.view(pt.shape)


def get_size_reduction(effective_number_bits, bucket_size=256, full_precision_bits=32):
    """Compression factor of b-bit weights with two full-precision scalars per bucket
    (reference: helpers/functions.py:216-224)."""
    if bucket_size is None:
        return full_precision_bits / effective_number_bits
    f, k, b = full_precision_bits, bucket_size, effective_number_bits
    return (k * f) / (k * b + 2 * f)


def get_size_quantized_model(model, numBits, quantization_functions, bucket_size=256, type_quantization="uniform",
                             quantizeFirstLastLayer=True):
    """Model size in MB with Huffman-coded indices (reference: helpers/functions.py:226-262)."""
    params = list(model.parameters())
    if numBits is None:
        return sum(p.numel() for p in params) * 4 / 1000000
    quantized = params if quantizeFirstLastLayer is True else params[1:-1]
    unquantized = [] if quantizeFirstLastLayer is True else [params[0], params[-1]]
    count_q = sum(p.numel() for p in quantized)
    count_u = sum(p.numel() for p in unquantized)
    mean_bits = qhf.get_huffman_encoding_mean_bit_length(iter(quantized), quantization_functions, type_quantization,
                                                         s=2 ** numBits)
    size = count_u * 4 + mean_bits * count_q / 8
    if bucket_size is not None:
        size += count_q / bucket_size * 8
    return size / 1000000
