"""Kernel-level parity (GPU): each C-ABI op against a plain PyTorch fp32 restatement of the same op on identical
inputs.  Tolerances are for fp16 operands with fp32 accumulation."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F():
    from gslora import _ffi
    _ffi.lib()
    return _ffi


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


GEMM_CASES = [
    # M, N, K, epi, cta_group, block_n
    (256, 256, 128, 0, 1, 128), (512, 512, 512, 0, 2, 256), (1000, 384, 528, 0, 1, 128), (1000, 1536, 528, 0, 2, 256),
    (300, 512, 2064, 1, 2, 256), (520, 2048, 528, 2, 1, 256), (520, 2048, 528, 2, 2, 256), (520, 2048, 528, 3, 2, 256),
    (777, 512, 2064, 4, 2, 256), (777, 128, 272, 4, 2, 128), (26 * 5, 128, 192, 5, 1, 128), (197 * 3, 512, 192, 5, 2, 256),
    (9456, 512, 2064, 4, 0, 0), (9456, 2048, 528, 2, 0, 0),
]


@pytest.mark.parametrize("M,N,K,epi,cg,bn", GEMM_CASES)
def test_gemm_epilogues(F, M, N, K, epi, cg, bn):
    torch.manual_seed(M + N + K + epi)
    dev = "cuda"
    A = (torch.randn(M, K, device=dev) * 0.5).half()
    B = (torch.randn(N, K, device=dev) * 0.05).half()
    bias = torch.randn(N, device=dev)
    acc = A.float() @ B.float().t()
    out1 = aux = None
    period = 0
    want1 = None
    if epi == F.EPI_F16:
        out0 = torch.empty(M, N, device=dev, dtype=torch.half); want0 = acc + bias
    elif epi == F.EPI_F32:
        out0 = torch.empty(M, N, device=dev); out1 = torch.empty(M, N, device=dev, dtype=torch.half); want0 = want1 = acc + bias
    elif epi == F.EPI_GELU:
        out0 = torch.empty(M, N, device=dev, dtype=torch.half); out1 = torch.empty(M, N, device=dev, dtype=torch.half)
        h = (acc + bias).requires_grad_(True); want1 = torch.nn.functional.gelu(h); want1.sum().backward()
        want0 = h.grad; want1 = want1.detach()       # out0 = gelu'(h) (what the backward multiplies by), out1 = gelu(h)
    elif epi == F.EPI_GELU_BWD:
        aux = torch.randn(M, N, device=dev).half(); out0 = torch.empty(M, N, device=dev, dtype=torch.half)
        bias = None; want0 = acc * aux.float()
    elif epi == F.EPI_RES_F32:
        aux = torch.randn(M, N, device=dev); out0 = torch.empty(M, N, device=dev)
        want0 = acc + bias + aux
    else:
        period = 26 if M % 26 == 0 else 197
        aux = torch.randn(period, N, device=dev); out0 = torch.empty(M, N, device=dev); bias = None
        want0 = acc + aux.repeat(M // period, 1)
    F.gemm_f16(A, B, epi=epi, bias=bias, out0=out0, out1=out1, aux=aux, aux_period=period, cta_group=cg, block_n=bn)
    torch.cuda.synchronize()
    tol0 = 2e-3 if out0.dtype == torch.half else 2e-5      # fp16 output rounding vs fp32 output
    assert (out0.float() - want0).abs().max() <= tol0 * want0.abs().max() + 1e-3 * (out0.dtype == torch.half)
    if want1 is not None:
        assert (out1.float() - want1).abs().max() <= 2e-3 * want1.abs().max() + 1e-3


@pytest.mark.parametrize("M,D", [(197 * 3, 512), (1000, 128), (77, 768), (64, 1024)])
def test_layernorm_fwd_bwd(F, M, D):
    torch.manual_seed(0)
    x = torch.randn(M, D, device="cuda") * 2 + 0.5
    g = 1 + 0.1 * torch.randn(D, device="cuda"); b = 0.1 * torch.randn(D, device="cuda")
    y16 = torch.zeros(M, D + 16, device="cuda", dtype=torch.half)
    mean = torch.empty(M, device="cuda"); rstd = torch.empty(M, device="cuda")
    F.check(F.lib().gsl_layernorm_fwd(F.ptr(x), D, F.ptr(g), F.ptr(b), 1e-5, F.ptr(y16), D + 16, F.ptr(mean), F.ptr(rstd), M, D, F.cur_stream()))
    xr = x.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (D,), g, b, 1e-5)
    assert (y16[:, :D].float() - ref).abs().max() < 4e-3
    assert (y16[:, D:] == 0).all()
    dy = torch.randn(M, D, device="cuda"); dres = torch.randn(M, D, device="cuda")
    ref.backward(dy)
    dx = dres.clone(); dx16 = torch.empty(M, D + 16, device="cuda", dtype=torch.half)
    F.check(F.lib().gsl_layernorm_bwd(F.ptr(dy), D, F.ptr(x), D, F.ptr(mean), F.ptr(rstd), F.ptr(g), F.ptr(dx), D, F.ptr(dx), D, F.ptr(dx16), D + 16, M, D, F.cur_stream()))
    want = xr.grad + dres
    assert rel(dx, want) < 1e-5
    assert rel(dx16[:, :D].float(), want) < 1e-3


def _e4m3_decode(b):
    """uint8 tensor of e4m3 codes -> fp32 (bias 7, no infinities)."""
    b = b.to(torch.int32)
    sign = torch.where((b & 0x80) != 0, -1.0, 1.0)
    e = (b >> 3) & 0xF
    m = (b & 7).float()
    val = torch.where(e == 0, m / 8 * 2.0 ** -6, (1 + m / 8) * torch.pow(2.0, (e - 7).float()))
    return sign * val


@pytest.mark.parametrize("M,N,K,epi,bn", [(512, 512, 512, 1, 256), (1000, 1536, 576, 1, 256), (777, 512, 2048, 4, 256), (300, 384, 192, 1, 128),
                                          (9456, 512, 2048, 4, 0), (9456, 2048, 512, 2, 0), (9456, 512, 1536, 0, 0)])
def test_gemm_split8_fp8_residual(F, M, N, K, epi, bn):
    """Precision mode split8: B = fp16(W 2^s), B_lo8 = e4m3(W 2^s - B); the kernel adds e5m2(A) B_lo8^T on the FP8 tensor path and un-scales.
    The product must sit far below the single-rounding floor (the fp16 weight error 2^-12 shrinks by the e4m3 / e5m2 precision, ~2^-4)."""
    torch.manual_seed(M + N + K + 1)
    shift = 12
    A = (torch.randn(M, K, device="cuda") * 0.5).half()
    W = torch.randn(N, K, device="cuda") * 0.05
    hi = torch.empty(N, K, device="cuda", dtype=torch.half); lo8 = torch.empty(N, K, device="cuda", dtype=torch.uint8)
    F.check(F.lib().gsl_cast_f32_to_f16_split8(F.ptr(W), K, F.ptr(hi), F.ptr(lo8), K, N, K, shift, 0, F.cur_stream()))
    assert torch.equal(hi, (W * 2.0 ** shift).half())
    want8 = (W * 2.0 ** shift - hi.float()).to(torch.float8_e4m3fn).view(torch.uint8)      # torch's round-to-nearest-even e4m3 cast of the same residual
    assert (lo8 == want8).float().mean() > 0.9999 and ((lo8 & 0x7F).int() - (want8 & 0x7F).int()).abs().max() <= 1
    recon = (hi.float() + _e4m3_decode(lo8)) * 2.0 ** -shift
    assert rel(recon, W) < 2.5e-5 and rel(hi.float() * 2.0 ** -shift, W) > 1.5e-4       # ~15 significant bits vs 11
    hiT = torch.empty(K, N, device="cuda", dtype=torch.half); lo8T = torch.empty(K, N, device="cuda", dtype=torch.uint8)
    F.check(F.lib().gsl_cast_f32_to_f16_split8(F.ptr(W), K, F.ptr(hiT), F.ptr(lo8T), N, N, K, shift, 1, F.cur_stream()))
    assert torch.equal(hiT, hi.t().contiguous()) and torch.equal(lo8T, lo8.t().contiguous())
    bias = torch.randn(N, device="cuda")
    exact = (A.double() @ W.double().t()).float() + bias
    one_w = W.half()
    if epi == F.EPI_F32:
        out0 = torch.empty(M, N, device="cuda"); one = torch.empty(M, N, device="cuda")
        F.gemm_f16(A, hi, B_lo8=lo8, lo8_shift=shift, epi=epi, bias=bias, out0=out0, block_n=bn)
        F.gemm_f16(A, one_w, epi=epi, bias=bias, out0=one, block_n=bn)
        # the oracle's restatement of this arithmetic (fp16 x fp16 + e5m2 x e4m3, exact products, fp64 accumulation): agreement to accumulation order
        from oracle import operand_emul as E
        emul = E.gemm(A.cpu(), "split8", W.cpu(), shift).float().cuda() + bias
        print(f"split8 {M}x{N}x{K}: {rel(out0, exact):.2e} vs FP32 (single rounding {rel(one, exact):.2e}), {rel(out0, emul):.2e} vs the operand emulation")
        assert rel(out0, emul) < 2e-6
        assert rel(out0, exact) < 3e-5
        assert rel(one, exact) > 5 * rel(out0, exact)
    elif epi == F.EPI_RES_F32:
        res = torch.randn(M, N, device="cuda"); out0 = torch.empty(M, N, device="cuda")
        F.gemm_f16(A, hi, B_lo8=lo8, lo8_shift=shift, epi=epi, bias=bias, out0=out0, aux=res, block_n=bn)
        assert rel(out0, exact + res) < 3e-5
    elif epi == F.EPI_F16:
        out0 = torch.empty(M, N, device="cuda", dtype=torch.half)
        F.gemm_f16(A, hi, B_lo8=lo8, lo8_shift=shift, epi=epi, bias=bias, out0=out0, block_n=bn)
        assert rel(out0.float(), exact) < 4e-4      # fp16 output rounding
    else:       # GELU: fp16 outputs, checked to fp16 rounding
        gp = torch.empty(M, N, device="cuda", dtype=torch.half); g = torch.empty(M, N, device="cuda", dtype=torch.half)
        F.gemm_f16(A, hi, B_lo8=lo8, lo8_shift=shift, epi=epi, bias=bias, out0=gp, out1=g, block_n=bn)
        want = torch.nn.functional.gelu(exact)
        assert (g.float() - want).abs().max() <= 1e-3 * want.abs().max() + 1e-4



@pytest.mark.parametrize("M,K,r", [(1000, 512, 8), (197 * 5, 2048, 8), (333, 1024, 16), (50, 128, 8), (100, 272 - 16, 8)])
def test_lora_down(F, M, K, r):
    torch.manual_seed(1)
    X = torch.randn(M, K + 16, device="cuda").half()
    A = torch.zeros(16, K, device="cuda", dtype=torch.half); A[:r] = (torch.randn(r, K, device="cuda") * 0.05).half()
    X[:, K:] = 7.0
    F.check(F.lib().gsl_lora_down(F.ptr(X), K + 16, F.ptr(A), K, F.ptr(X[:, K:]), K + 16, M, K, r, F.cur_stream()))
    want = X[:, :K].float() @ A.float().t()
    assert (X[:, K:].float() - want).abs().max() <= 2e-3 * want.abs().max() + 1e-4
    assert (X[:, K + r:] == 0).all()


@pytest.mark.parametrize("M,N,r,tr", [(1000, 512, 8, 0), (197 * 7, 2048, 8, 1), (300, 256, 16, 1), (65, 128, 8, 0)])
def test_skinny_tn(F, M, N, r, tr):
    torch.manual_seed(2)
    L = torch.randn(M, N + 16, device="cuda").half(); R = torch.randn(M, 16, device="cuda").half()
    nb = F.lib().gsl_skinny_tn_workspace(M, N, r)
    ws = torch.empty(nb // 4 + 16, device="cuda")
    out = torch.full((r, N) if tr else (N, r), 0.5, device="cuda")
    want = 0.25 * (L[:, :N].float().t() @ R[:, :r].float())
    for acc in (0, 1):
        F.check(F.lib().gsl_skinny_tn(F.ptr(L), N + 16, F.ptr(R), 16, F.ptr(out), N if tr else r, tr, 0.25, acc, M, N, r, F.ptr(ws), nb, F.cur_stream()))
        got = out.t() if tr else out
        assert rel(got, want * (1 + acc)) < 1e-5


@pytest.mark.parametrize("M,N,r,tr", [(197 * 9, 2048, 8, 1), (197 * 9, 2048, 8, 0), (5000, 3072, 8, 0), (37, 256, 8, 1), (1, 512, 8, 0),
                                      (300, 1024, 16, 1), (200, 320, 8, 0)])
def test_lora_side_fused_pass(F, M, N, r, tr):
    """One pass over L gives T = L P^T (fp16) and out = scale * L^T R; rank 16 / N % 256 != 0 take the two-kernel fallback."""
    torch.manual_seed(4)
    L = torch.randn(M, N, device="cuda").half()
    P = torch.zeros(16, N, device="cuda", dtype=torch.half); P[:r] = (torch.randn(r, N, device="cuda") * 0.05).half()
    R = torch.zeros(M, 16, device="cuda", dtype=torch.half); R[:, :r] = torch.randn(M, r, device="cuda").half()
    T = torch.full((M, 16), 7.0, device="cuda", dtype=torch.half)
    nb = F.lib().gsl_lora_side_workspace(M, N, r)
    ws = torch.empty(nb // 4 + 16, device="cuda")
    out = torch.full((r, N) if tr else (N, r), 0.5, device="cuda")
    want_t = L.float() @ P[:r].float().t()
    want_q = 0.25 * (L.float().t() @ R[:, :r].float())
    for acc in (0, 1):
        F.check(F.lib().gsl_lora_side(F.ptr(L), N, F.ptr(P), N, F.ptr(T), 16, F.ptr(R), 16, F.ptr(out), N if tr else r, tr, 0.25, acc,
                                      M, N, r, F.ptr(ws), nb, F.cur_stream()))
        got = out.t() if tr else out
        assert rel(got, want_q * (1 + acc)) < 1e-5
        assert (T[:, :r].float() - want_t).abs().max() <= 2e-3 * want_t.abs().max() + 1e-4
        assert (T[:, r:] == 0).all()


@pytest.mark.parametrize("B,N,heads,scale", [(3, 197, 8, 512 ** -0.5), (2, 26, 2, 128 ** -0.5), (1, 197, 12, 0.125), (5, 50, 4, 0.2)])
def test_attention_fwd_bwd(F, B, N, heads, scale):
    torch.manual_seed(3)
    D = heads * 64
    qkv = torch.randn(B * N, 3 * D, device="cuda").half()
    out = torch.empty(B * N, D, device="cuda", dtype=torch.half); lse = torch.empty(B * heads * N, device="cuda")
    F.check(F.lib().gsl_attention_fwd(F.ptr(qkv), 3 * D, F.ptr(out), D, F.ptr(lse), B, N, heads, scale, F.cur_stream()))
    q32 = qkv.float().requires_grad_(True)
    q, k, v = [t.reshape(B, N, heads, 64).permute(0, 2, 1, 3) for t in q32.chunk(3, dim=-1)]
    dots = torch.einsum("bhid,bhjd->bhij", q, k) * scale
    ref = torch.einsum("bhij,bhjd->bhid", dots.softmax(-1), v).permute(0, 2, 1, 3).reshape(B * N, D)
    assert rel(out.float(), ref) < 2e-3
    assert (lse.view(B, heads, N) - torch.logsumexp(dots, -1)).abs().max() < 2e-3
    dout = torch.randn(B * N, D, device="cuda").half()
    ref.backward(dout.float())
    dqkv = torch.empty(B * N, 3 * D, device="cuda", dtype=torch.half)
    F.check(F.lib().gsl_attention_bwd(F.ptr(qkv), 3 * D, F.ptr(out), D, F.ptr(dout), D, F.ptr(lse), F.ptr(dqkv), 3 * D, B, N, heads, scale, F.cur_stream()))
    for i, name in enumerate("qkv"):
        assert rel(dqkv[:, i * D:(i + 1) * D].float(), q32.grad[:, i * D:(i + 1) * D]) < 4e-3, name


@pytest.mark.parametrize("B,N,heads,scale,split", [(3, 197, 8, 512 ** -0.5, True), (2, 26, 2, 128 ** -0.5, False), (5, 50, 4, 0.2, True)])
def test_rowdot_epilogue_feeds_attention_bwd(F, B, N, heads, scale, split):
    """EPI_F16_ROWDOT: dO = dY Wo^T with delta = rowsum(dO * O) per (image, head, token) as two partial sums; gsl_attention_bwd_rowdot with those
    equals gsl_attention_bwd computing delta itself (autograd of Attention.forward, vit_face.py:358-379)."""
    torch.manual_seed(11)
    D = heads * 64
    M = B * N
    qkv = torch.randn(M, 3 * D, device="cuda").half()
    out = torch.empty(M, D, device="cuda", dtype=torch.half); lse = torch.empty(B * heads * N, device="cuda")
    F.check(F.lib().gsl_attention_fwd(F.ptr(qkv), 3 * D, F.ptr(out), D, F.ptr(lse), B, N, heads, scale, F.cur_stream()))
    dy = (torch.randn(M, D, device="cuda") * 0.5).half()
    w = torch.randn(D, D, device="cuda") * 0.05
    w_hi = w.half(); w_lo = (w - w_hi.float()).half()
    do16 = torch.empty(M, D, device="cuda", dtype=torch.half)
    parts = torch.full((2, B, heads, N), float("nan"), device="cuda")
    F.gemm_f16(dy, w_hi, B_lo=w_lo if split else None, epi=F.EPI_F16_ROWDOT, out0=do16, out1=parts, aux=out, aux_period=N)
    do_ref = dy.float() @ (w if split else w_hi.float()).t()
    assert rel(do16.float(), do_ref) < 1e-3
    delta_ref = (do_ref * out.float()).view(B, N, heads, 64).sum(-1).permute(0, 2, 1)          # [B, heads, N]
    assert not torch.isnan(parts).any()
    assert rel(parts.sum(0), delta_ref) < 1e-4
    half_ref = (do_ref * out.float()).view(B, N, heads, 2, 32).sum(-1).permute(3, 0, 2, 1)    # [2, B, heads, N]
    assert rel(parts, half_ref) < 1e-4
    dq_a = torch.empty(M, 3 * D, device="cuda", dtype=torch.half); dq_b = torch.empty_like(dq_a)
    F.check(F.lib().gsl_attention_bwd(F.ptr(qkv), 3 * D, F.ptr(out), D, F.ptr(do16), D, F.ptr(lse), F.ptr(dq_a), 3 * D, B, N, heads, scale, F.cur_stream()))
    F.check(F.lib().gsl_attention_bwd_rowdot(F.ptr(qkv), 3 * D, F.ptr(do16), D, F.ptr(lse), F.ptr(parts), F.ptr(dq_b), 3 * D, B, N, heads, scale,
                                             F.cur_stream()))
    assert rel(dq_b.float(), dq_a.float()) < 1e-3
    q32 = qkv.float().requires_grad_(True)
    q, k, v = [t.reshape(B, N, heads, 64).permute(0, 2, 1, 3) for t in q32.chunk(3, dim=-1)]
    ref = torch.einsum("bhij,bhjd->bhid", (torch.einsum("bhid,bhjd->bhij", q, k) * scale).softmax(-1), v).permute(0, 2, 1, 3).reshape(M, D)
    ref.backward(do16.float())
    for i, name in enumerate("qkv"):
        assert rel(dq_b[:, i * D:(i + 1) * D].float(), q32.grad[:, i * D:(i + 1) * D]) < 4e-3, name


def test_grouplasso_adamw_matches_torch(F):
    torch.manual_seed(4)
    G, n_per = 6, 40960
    p = torch.randn(G * n_per, device="cuda") * 0.05
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([ref], lr=1e-2, weight_decay=0.05)
    m = torch.zeros_like(p); v = torch.zeros_like(p)
    offs = (torch.arange(G + 1, device="cuda", dtype=torch.int32) * n_per).contiguous()
    norms = torch.empty(G, device="cuda")
    alpha = 1e-2
    for step in range(1, 4):
        g = torch.randn_like(p) * 1e-3
        opt.zero_grad()
        loss = (ref * g).sum() + alpha * ref.view(G, -1).norm(dim=1).sum()
        loss.backward(); opt.step()
        F.check(F.lib().gsl_grouplasso_adamw_step(F.ptr(p), F.ptr(g), F.ptr(m), F.ptr(v), F.ptr(offs), G, p.numel(), 1e-2, 0.05, 0.9, 0.999,
                                                  1e-8, alpha, 1.0, step, F.ptr(norms), F.cur_stream()))
        assert rel(p, ref.detach()) < 2e-6
    tn = torch.empty(G, device="cuda")
    F.check(F.lib().gsl_tensor_norms(F.ptr(p), F.ptr(offs), G, 1, F.ptr(tn), F.cur_stream()))
    assert rel(tn, p.view(G, -1).abs().sum(1)) < 1e-5


@pytest.mark.parametrize("Br,Bf,D,C,bnd", [(5, 3, 512, 100, 50.0), (4, 4, 128, 10, 1e-4), (7, 0, 768, 100, 18.0)])
def test_prototype_kl_matches_reference_expression(F, Br, Bf, D, C, bnd):
    """gsl_prototype_kl_fwd / _grad vs engine_cl.get_prototype_loss' expression (engine_cl.py:571-603) and its use at :97-101."""
    import torch.nn.functional as Fn
    torch.manual_seed(5)
    B = Br + Bf
    emb = torch.randn(B, D, device="cuda"); proto = torch.randn(C, D, device="cuda") * 2; lab = torch.randint(0, C, (B,), device="cuda")
    w_f, w_r = 0.7, 0.3
    e = emb.clone().requires_grad_(True)
    def kl(x, y):
        return Fn.kl_div(Fn.log_softmax(x, 1), Fn.log_softmax(proto[y], 1), reduction="batchmean", log_target=True)
    kr = kl(e[:Br], lab[:Br])
    kf = kl(e[Br:], lab[Br:]) if Bf else torch.zeros((), device="cuda")
    (w_f * torch.relu(bnd - kf) + w_r * kr).backward() if Bf else (w_r * kr).backward()
    klv = torch.empty(B, device="cuda")
    F.check(F.lib().gsl_prototype_kl_fwd(F.ptr(emb), F.ptr(lab), F.ptr(proto), B, D, F.ptr(klv), F.cur_stream()))
    ce = torch.zeros(B, device="cuda"); sums = torch.zeros(8, device="cuda")
    F.check(F.lib().gsl_loss_sums(F.ptr(ce), None, F.ptr(klv), Br, B, F.ptr(sums), F.cur_stream()))
    assert abs(float(sums[6] / max(Br, 1)) - float(kr)) < 1e-4 * max(1.0, abs(float(kr)))
    if Bf:
        assert abs(float(sums[7] / Bf) - float(kf)) < 1e-4 * max(1.0, abs(float(kf)))
    demb = torch.empty(B, D, device="cuda")
    F.check(F.lib().gsl_prototype_kl_grad(F.ptr(emb), F.ptr(lab), F.ptr(proto), F.ptr(sums), Br, B, D, w_f, w_r, bnd, F.ptr(demb), F.cur_stream()))
    assert (demb - e.grad).abs().max() < 1e-5 * max(1.0, float(e.grad.abs().max()))


# ------------------------------------------------------------------------------------------------ split-precision operands (GslConfig.precision = 1)
def _split(F, W, transpose=False):
    """(hi, lo) fp16 pair of an fp32 matrix through the library's own cast (gsl_cast_f32_to_f16_split)."""
    R, C = W.shape
    shape = (C, R) if transpose else (R, C)
    hi = torch.empty(shape, device="cuda", dtype=torch.half); lo = torch.empty(shape, device="cuda", dtype=torch.half)
    F.check(F.lib().gsl_cast_f32_to_f16_split(F.ptr(W), C, F.ptr(hi), F.ptr(lo), shape[1], R, C, 1.0, 1 if transpose else 0, F.cur_stream()))
    return hi, lo


@pytest.mark.parametrize("M,N,K,epi,cg,bn", [(512, 512, 512, 1, 2, 256), (1000, 1536, 528, 1, 2, 256), (777, 512, 2064, 4, 2, 256), (520, 2048, 528, 1, 1, 256),
                                             (300, 384, 192, 1, 1, 128), (9456, 512, 2048, 4, 0, 0), (9456, 2048, 512, 2, 0, 0)])
def test_gemm_split_weight_is_exact_in_the_weight(F, M, N, K, epi, cg, bn):
    """A (B_hi + B_lo)^T against the fp32 product with the UNROUNDED weight: the only rounding left is the fp16 A operand, which the test
    applies to the reference too -- so fp32-output epilogues must agree to accumulation-order level, 100x below one fp16 weight rounding."""
    torch.manual_seed(M + N + K)
    A = (torch.randn(M, K, device="cuda") * 0.5).half()
    W = torch.randn(N, K, device="cuda") * 0.05
    hi, lo = _split(F, W)
    assert torch.equal(hi, W.half()) and rel(hi.float() + lo.float(), W) < 2e-6
    hiT, loT = _split(F, W, transpose=True)
    assert torch.equal(hiT, hi.t().contiguous()) and torch.equal(loT, lo.t().contiguous())
    bias = torch.randn(N, device="cuda")
    exact = (A.double() @ W.double().t()).float() + bias
    if epi == F.EPI_F32:
        out0 = torch.empty(M, N, device="cuda")
        F.gemm_f16(A, hi, B_lo=lo, epi=epi, bias=bias, out0=out0, cta_group=cg, block_n=bn)
        one = torch.empty(M, N, device="cuda")
        F.gemm_f16(A, hi, epi=epi, bias=bias, out0=one, cta_group=cg, block_n=bn)
        assert rel(out0, exact) < 3e-6
        assert rel(one, exact) > 20 * rel(out0, exact)                 # the single-rounding product sits at the fp16 weight floor (~2e-4)
    elif epi == F.EPI_RES_F32:
        res = torch.randn(M, N, device="cuda"); out0 = torch.empty(M, N, device="cuda")
        F.gemm_f16(A, hi, B_lo=lo, epi=epi, bias=bias, out0=out0, aux=res, cta_group=cg, block_n=bn)
        assert rel(out0, exact + res) < 1e-5
    else:       # GELU: fp16 outputs, checked to fp16 rounding
        gp = torch.empty(M, N, device="cuda", dtype=torch.half); g = torch.empty(M, N, device="cuda", dtype=torch.half)
        F.gemm_f16(A, hi, B_lo=lo, epi=epi, bias=bias, out0=gp, out1=g, cta_group=cg, block_n=bn)
        want = torch.nn.functional.gelu(exact)
        assert (g.float() - want).abs().max() <= 1e-3 * want.abs().max() + 1e-4


@pytest.mark.parametrize("M,K,r", [(1000, 512, 8), (197 * 5, 2048, 8), (333, 1024, 16), (300, 512, 4), (300, 768, 12), (100, 272 - 16, 5)])
def test_lora_down_split_and_odd_ranks(F, M, K, r):
    """T = X P^T with P = hi + lo (32-row operand): exact in P; ranks 1..16 ride on the zero-padded 16-row tiles."""
    torch.manual_seed(3)
    X = torch.randn(M, K, device="cuda").half()
    P = torch.randn(r, K, device="cuda") * 0.05
    P32 = torch.zeros(32, K, device="cuda", dtype=torch.half)
    P32[:r] = P.half(); P32[16:16 + r] = (P - P.half().float()).half()
    T = torch.full((M, 16), 7.0, device="cuda", dtype=torch.half)
    F.check(F.lib().gsl_lora_down_split(F.ptr(X), K, F.ptr(P32), K, F.ptr(T), 16, M, K, r, F.cur_stream()))
    want = (X.double() @ P.double().t()).float()
    assert (T[:, :r].float() - want).abs().max() <= 1e-3 * want.abs().max() + 1e-5      # fp16 store rounding only
    assert (T[:, r:] == 0).all()
    T1 = torch.full((M, 16), 7.0, device="cuda", dtype=torch.half)
    F.check(F.lib().gsl_lora_down(F.ptr(X), K, F.ptr(P32), K, F.ptr(T1), 16, M, K, r, F.cur_stream()))
    want1 = X.float() @ P32[:r].float().t()
    assert (T1[:, :r].float() - want1).abs().max() <= 1e-3 * want1.abs().max() + 1e-5
    assert (T1[:, r:] == 0).all()


@pytest.mark.parametrize("M,N,r,tr", [(197 * 9, 2048, 8, 1), (5000, 3072, 8, 0), (300, 1024, 16, 1), (333, 2048, 6, 0), (200, 320, 8, 0)])
def test_lora_side_split(F, M, N, r, tr):
    torch.manual_seed(5)
    L = torch.randn(M, N, device="cuda").half()
    P = torch.randn(r, N, device="cuda") * 0.05
    P32 = torch.zeros(32, N, device="cuda", dtype=torch.half)
    P32[:r] = P.half(); P32[16:16 + r] = (P - P.half().float()).half()
    R = torch.zeros(M, 16, device="cuda", dtype=torch.half); R[:, :r] = torch.randn(M, r, device="cuda").half()
    T = torch.full((M, 16), 3.0, device="cuda", dtype=torch.half)
    nb = F.lib().gsl_lora_side_workspace(M, N, r)
    ws = torch.empty(nb // 4 + 16, device="cuda")
    out = torch.zeros((r, N) if tr else (N, r), device="cuda")
    F.check(F.lib().gsl_lora_side_split(F.ptr(L), N, F.ptr(P32), N, F.ptr(T), 16, F.ptr(R), 16, F.ptr(out), N if tr else r, tr, 0.5, 0, M, N, r,
                                        F.ptr(ws), nb, F.cur_stream()))
    wantT = (L.double() @ P.double().t()).float()
    assert (T[:, :r].float() - wantT).abs().max() <= 1e-3 * wantT.abs().max() + 1e-5
    assert (T[:, r:] == 0).all()
    wantQ = 0.5 * (L.float().t() @ R[:, :r].float())
    assert rel(out.t() if tr else out, wantQ) < 1e-5


def test_grouplasso_adamw_skips_groups_with_nonfinite_gradients(F):
    """Overflow guard of the loss-scaled fp16 gradient stream: a group with an inf / NaN gradient keeps its parameters and moments and reports
    n_g = NaN (engine_cl.StepResult raises FloatingPointError on it); the other groups step normally."""
    torch.manual_seed(6)
    G, n_per = 3, 4096
    p = torch.randn(G * n_per, device="cuda") * 0.05
    g = torch.randn_like(p) * 1e-3
    g[n_per + 17] = float("inf")
    m = torch.zeros_like(p); v = torch.zeros_like(p)
    offs = (torch.arange(G + 1, device="cuda", dtype=torch.int32) * n_per).contiguous()
    norms = torch.empty(G, device="cuda")
    p0 = p.clone()
    F.check(F.lib().gsl_grouplasso_adamw_step(F.ptr(p), F.ptr(g), F.ptr(m), F.ptr(v), F.ptr(offs), G, p.numel(), 1e-2, 0.05, 0.9, 0.999, 1e-8, 1e-2, 1.0, 1,
                                              F.ptr(norms), F.cur_stream()))
    assert torch.isnan(norms[1]) and not torch.isnan(norms[0]) and not torch.isnan(norms[2])
    assert torch.equal(p[n_per:2 * n_per], p0[n_per:2 * n_per]) and float(m[n_per:2 * n_per].abs().max()) == 0.0
    assert not torch.equal(p[:n_per], p0[:n_per]) and not torch.equal(p[2 * n_per:], p0[2 * n_per:])
    assert torch.isfinite(p).all()
