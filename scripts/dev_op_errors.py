"""Dev (GPU): relative L2 error of each hot op against an fp64 evaluation of the same fp16 inputs, next to the error an ideal implementation
with the same fp16 materialisation points would have (rounding emulated in fp64) -- finds ops that lose more than their roundings explain."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "gs-lora_b200"), ROOT]
import torch
from gslora import _ffi as F
torch.manual_seed(0)
dev = "cuda"
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
h16 = lambda t: t.half().double()

# ---------------- attention fwd / bwd at the P8S8 shape with realistic magnitudes (q, k ~ LN(x) W: std ~0.45; scale 512^-0.5)
B, N, heads = 16, 197, 8
D = heads * 64
for std, scale in ((0.45, 512 ** -0.5), (1.0, 0.125), (2.0, 512 ** -0.5)):
    qkv = (torch.randn(B * N, 3 * D, device=dev) * std).half()
    out = torch.empty(B * N, D, device=dev, dtype=torch.half); lse = torch.empty(B * heads * N, device=dev)
    F.check(F.lib().gsl_attention_fwd(F.ptr(qkv), 3 * D, F.ptr(out), D, F.ptr(lse), B, N, heads, scale, F.cur_stream()))
    q64 = qkv.double().requires_grad_(True)
    q, k, v = [t.reshape(B, N, heads, 64).permute(0, 2, 1, 3) for t in q64.chunk(3, dim=-1)]
    P = (torch.einsum("bhid,bhjd->bhij", q, k) * scale).softmax(-1)
    ref = torch.einsum("bhij,bhjd->bhid", P, v).permute(0, 2, 1, 3).reshape(B * N, D)
    ideal = h16(torch.einsum("bhij,bhjd->bhid", h16(P), v).permute(0, 2, 1, 3).reshape(B * N, D))
    print(f"attention fwd  std {std} scale {scale:.4f}: kernel {rel(out, ref):.2e}   ideal (P, O fp16) {rel(ideal, ref):.2e}   O-rounding only {rel(h16(ref), ref):.2e}")
    dout = (torch.randn(B * N, D, device=dev) * 0.1).half()
    ref.backward(dout.double())
    dqkv = torch.empty(B * N, 3 * D, device=dev, dtype=torch.half)
    F.check(F.lib().gsl_attention_bwd(F.ptr(qkv), 3 * D, F.ptr(out), D, F.ptr(dout), D, F.ptr(lse), F.ptr(dqkv), 3 * D, B, N, heads, scale, F.cur_stream()))
    # ideal: P fp16 for dV, delta from fp16 O, dS fp16, outputs fp16
    with torch.no_grad():
        do = dout.double().reshape(B, N, heads, 64).permute(0, 2, 1, 3)
        o16 = out.double().reshape(B, N, heads, 64).permute(0, 2, 1, 3)
        delta = (do * o16).sum(-1, keepdim=True)
        dP = torch.einsum("bhid,bhjd->bhij", do, v)
        dS = h16(P * (dP - delta)) * scale
        dq = torch.einsum("bhij,bhjd->bhid", dS, k); dk = torch.einsum("bhij,bhid->bhjd", dS, q); dv = torch.einsum("bhij,bhid->bhjd", h16(P), do)
        idl = torch.cat([t.permute(0, 2, 1, 3).reshape(B * N, D) for t in (dq, dk, dv)], dim=1)
    for i, nm in enumerate("qkv"):
        sl = slice(i * D, (i + 1) * D)
        print(f"   bwd d{nm}: kernel {rel(dqkv[:, sl], q64.grad[:, sl]):.2e}   ideal {rel(h16(idl[:, sl]), q64.grad[:, sl]):.2e}   out-rounding only {rel(h16(q64.grad[:, sl]), q64.grad[:, sl]):.2e}")

# ---------------- LayerNorm fwd / bwd
M, Dm = 197 * 16, 512
x = torch.randn(M, Dm, device=dev) * 1.5 + 0.2
g = 1 + 0.05 * torch.randn(Dm, device=dev); b = 0.02 * torch.randn(Dm, device=dev)
y16 = torch.empty(M, Dm, device=dev, dtype=torch.half); mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
F.check(F.lib().gsl_layernorm_fwd(F.ptr(x), Dm, F.ptr(g), F.ptr(b), 1e-5, F.ptr(y16), Dm, F.ptr(mean), F.ptr(rstd), M, Dm, F.cur_stream()))
x64 = x.double().requires_grad_(True)
ref = torch.nn.functional.layer_norm(x64, (Dm,), g.double(), b.double(), 1e-5)
print(f"layernorm fwd: kernel {rel(y16, ref):.2e}   out-rounding only {rel(h16(ref), ref):.2e}")
dy = torch.randn(M, Dm, device=dev) * 0.01; dres = torch.randn(M, Dm, device=dev) * 0.01
ref.backward(dy.double())
dx = torch.empty(M, Dm, device=dev); dx16 = torch.empty(M, Dm, device=dev, dtype=torch.half)
F.check(F.lib().gsl_layernorm_bwd(F.ptr(dy), Dm, F.ptr(x), Dm, F.ptr(mean), F.ptr(rstd), F.ptr(g), F.ptr(dres), Dm, F.ptr(dx), Dm, F.ptr(dx16), Dm, M, Dm, F.cur_stream()))
print(f"layernorm bwd: fp32 out {rel(dx, x64.grad + dres.double()):.2e}   fp16 copy {rel(dx16, x64.grad + dres.double()):.2e}")

# ---------------- GEMM epilogues (fp16 in, fp16 out): error beyond the output rounding
M2, K2, N2 = 197 * 16, 512, 2048
A = (torch.randn(M2, K2, device=dev)).half(); W = (torch.randn(N2, K2, device=dev) * 0.02).half(); bias = torch.randn(N2, device=dev) * 0.02
gp = torch.empty(M2, N2, device=dev, dtype=torch.half); gg = torch.empty(M2, N2, device=dev, dtype=torch.half)
F.gemm_f16(A, W, epi=F.EPI_GELU, bias=bias, out0=gp, out1=gg)
h = (A.double() @ W.double().t() + bias.double()).requires_grad_(True)
gr = torch.nn.functional.gelu(h); gr.sum().backward()
print(f"fc1 GELU epilogue: g kernel {rel(gg, gr):.2e} (rounding only {rel(h16(gr), gr):.2e});  g' kernel {rel(gp, h.grad):.2e} (rounding only {rel(h16(h.grad), h.grad):.2e})")
dyy = (torch.randn(M2, K2, device=dev) * 0.01).half(); WT = (torch.randn(N2, K2, device=dev) * 0.02).half()
dh = torch.empty(M2, N2, device=dev, dtype=torch.half)
F.gemm_f16(dyy, WT, epi=F.EPI_GELU_BWD, out0=dh, aux=gp)
refdh = (dyy.double() @ WT.double().t()) * gp.double()
print(f"dH epilogue: kernel {rel(dh, refdh):.2e} (rounding only {rel(h16(refdh), refdh):.2e})")
q16 = torch.empty(M2, N2, device=dev, dtype=torch.half)
F.gemm_f16(A, W, epi=F.EPI_F16, bias=bias, out0=q16)
print(f"F16 epilogue: kernel {rel(q16, h.detach()):.2e} (rounding only {rel(h16(h.detach()), h.detach()):.2e})")

# ---------------- cls attention (last block)
qkv = (torch.randn(B * N, 3 * D, device=dev) * 0.45).half()
