"""ORACLE (test infrastructure, NOT product code) -- CPU restatement of the ARITHMETIC of the engine's weight operands, so that the GPU GEMMs can be
checked against a definition instead of only against a tolerance.

The reference (loralib.Linear.forward, vit_pytorch_face/vit_face.py:326-338) multiplies fp32 activations with fp32 weights.  The engine feeds the tensor
cores fp16 activations and one of three weight encodings (include/gslora.h GslConfig.precision):

    fast    B = fp16(W)                                                     y = A B^T
    split   B_hi = fp16(W), B_lo = fp16(W - B_hi)                           y = A B_hi^T + A B_lo^T
    split8  B_hi = fp16(W 2^s), B_lo8 = e4m3(W 2^s - B_hi)                   y = 2^-s (A B_hi^T + e5m2(A) B_lo8^T)

all with exact products and fp32 (here: fp64) accumulation.  This file restates the encodings with torch's own fp16 / float8 casts
(round-to-nearest-even, saturating to the largest finite value) and evaluates the three products in fp64, so a test can ask the kernel for agreement to
accumulation-order level (~1e-6) -- far tighter than the distance of any mode to the fp32 reference.

Only tests/ may import this module.  It is never on the product path.
"""
from __future__ import annotations

import torch


def encode_split(W: torch.Tensor):
    """precision "split": (fp16(W), fp16(W - fp16(W)))"""
    hi = W.float().half()
    lo = (W.float() - hi.float()).half()
    return hi, lo


def encode_split8(W: torch.Tensor, shift: int = 12):
    """precision "split8": (fp16(W 2^shift), e4m3(W 2^shift - hi) as raw bytes).  The scaling by a power of two is exact."""
    v = W.float() * (2.0 ** shift)
    hi = v.half()
    lo8 = (v - hi.float()).to(torch.float8_e4m3fn)
    return hi, lo8.view(torch.uint8)


def decode_e4m3(b: torch.Tensor) -> torch.Tensor:
    return b.view(torch.float8_e4m3fn).float()


def e5m2(A: torch.Tensor) -> torch.Tensor:
    """the activation copy the split8 GEMM's converter warps make (cvt.rn.satfinite.e5m2x2.f16x2): fp16 -> e5m2, nearest-even, saturating"""
    a = A.float().clamp(-57344.0, 57344.0)
    return a.to(torch.float8_e5m2).float()


def gemm(A16: torch.Tensor, mode: str, W: torch.Tensor, shift: int = 12) -> torch.Tensor:
    """y = A W^T as the engine's GEMM family evaluates it in `mode` (fp64 accumulation of the exact operand products); A16 is the fp16 activation."""
    A = A16.double()
    if mode == "fast":
        return A @ W.float().half().double().t()
    if mode == "split":
        hi, lo = encode_split(W)
        return A @ hi.double().t() + A @ lo.double().t()
    if mode == "split8":
        hi, lo8 = encode_split8(W, shift)
        return (A @ hi.double().t() + e5m2(A16).double() @ decode_e4m3(lo8).double().t()) * (2.0 ** -shift)
    raise ValueError(mode)
