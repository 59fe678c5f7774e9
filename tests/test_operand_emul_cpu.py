"""CPU: the oracle's restatement of the engine's three weight encodings (oracle/operand_emul.py) and the error each leaves against the FP32 product --
the numbers DESIGN.md section 2 argues with, reproduced without a GPU.  The GPU counterpart (tests/test_kernels_gpu.py) holds the kernels to this
restatement at accumulation-order level."""
import torch

from oracle import operand_emul as E


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def test_weight_encodings_keep_the_stated_number_of_bits():
    torch.manual_seed(0)
    W = torch.randn(512, 768) * 0.05
    hi, lo = E.encode_split(W)
    assert rel(hi.float(), W) > 1.2e-4                      # one fp16 rounding: 2^-11 / sqrt(3) ~ 1.6e-4 for this distribution
    assert rel(hi.float() + lo.float(), W) < 2e-6           # fp16 pair: ~22 significant bits (lo's exponent range limits tiny weights)
    h8, l8 = E.encode_split8(W, 12)
    assert torch.equal(h8, (W * 4096).half()) and l8.dtype == torch.uint8
    r8 = rel((h8.float() + E.decode_e4m3(l8)) * 2.0 ** -12, W)
    assert 4e-6 < r8 < 2.5e-5                               # fp16 + e4m3 residual: ~15 significant bits, 8-30x below one fp16 rounding
    # the pre-scale is what makes the e4m3 residual representable: without it everything underflows to zero codes
    _, l0 = E.encode_split8(W, 0)
    assert (l0 & 0x7F).max() <= 1


def test_products_rank_fast_split8_split_against_fp32():
    torch.manual_seed(1)
    A = (torch.randn(300, 1024) * 0.5).half()
    W = torch.randn(256, 1024) * 0.05
    exact = A.double() @ W.double().t()
    e_fast, e_s8, e_s = (rel(E.gemm(A, m, W), exact) for m in ("fast", "split8", "split"))
    assert e_fast > 8e-5 and e_s8 < e_fast / 10 and e_s < e_s8 / 2, (e_fast, e_s8, e_s)
    # the e5m2 copy of the activations only touches the residual term: 2 mantissa bits (~5 % rms) of a term that is 2e-4 of the product -> ~1e-5,
    # RANDOM per element (the fp16 rounding of the activations themselves is 15x larger); the e4m3 residual leaves ~5e-6 of SYSTEMATIC weight error
    hi, lo8 = E.encode_split8(W)
    exact_a = (A.double() @ hi.double().t() + A.double() @ E.decode_e4m3(lo8).double().t()) * 2.0 ** -12
    assert rel(E.gemm(A, "split8", W), exact_a) < 2e-5 and rel(exact_a, exact) < 8e-6


def test_e5m2_cast_saturates_and_rounds_to_nearest_even():
    a = torch.tensor([0.0, 1.0, 1.1, 1.125, 1.375, 3.0e5, -7.0e4, 2.0 ** -16, 2.0 ** -18]).half()
    got = E.e5m2(a)
    assert got.tolist()[:5] == [0.0, 1.0, 1.0, 1.0, 1.5]    # 2 mantissa bits: 1.125 and 1.375 are ties -> even
    assert got[5] == 57344.0 and got[6] == -57344.0         # fp16 inf / large -> largest finite e5m2
    assert got[7] == 2.0 ** -16 and got[8] == 0.0           # smallest subnormal kept, half of it rounds to zero (tie -> even)
