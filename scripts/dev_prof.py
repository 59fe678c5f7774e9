"""Dev harness (GPU box): launch each hot kernel once at the config-2 shape (for ncu captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "gs-lora_b200"), ROOT]
import torch
from gslora import _ffi as F
dev = "cuda"
B, N, heads = int(os.environ.get("B", "1024")), 197, 8
M, D, H = B * N, 512, 2048
torch.manual_seed(0)
what = sys.argv[1:] or ["gelu", "res", "f16", "attn", "skinny"]
reps = int(os.environ.get("REPS", "1"))
def timeit(fn, name, flops=None, bytes_=None):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    extra = (f" {flops/ms/1e9:.0f} TFLOP/s" if flops else "") + (f" {bytes_/ms/1e6:.0f} GB/s" if bytes_ else "")
    print(f"{name}: {ms:.3f} ms{extra}", flush=True)
if "gelu" in what:
    A = (torch.randn(M, D, device=dev) * 0.5).half(); W = (torch.randn(H, D, device=dev) * 0.05).half(); bias = torch.randn(H, device=dev)
    o0 = torch.empty(M, H, device=dev, dtype=torch.half); o1 = torch.empty(M, H, device=dev, dtype=torch.half)
    timeit(lambda: F.gemm_f16(A, W, epi=F.EPI_GELU, bias=bias, out0=o0, out1=o1), "fc1 GELU gemm", flops=2.0 * M * H * D)
    timeit(lambda: F.gemm_f16(A, W, epi=F.EPI_GELU, bias=bias, out0=o0, out1=o1, drop_p=0.1, drop_seed=1234), "fc1 GELU gemm, dropout 0.1", flops=2.0 * M * H * D)
    dy = (torch.randn(M, D, device=dev) * 0.5).half(); WT = (torch.randn(H, D, device=dev) * 0.05).half()
    timeit(lambda: F.gemm_f16(dy, WT, epi=F.EPI_GELU_BWD, out0=o1, aux=o0), "dH gemm (x saved gelu')", flops=2.0 * M * H * D)
    del A, W, o0, o1, dy, WT
if "res" in what:
    G = (torch.randn(M, H, device=dev) * 0.5).half(); W2 = (torch.randn(D, H, device=dev) * 0.05).half(); bias = torch.randn(D, device=dev)
    x = torch.randn(M, D, device=dev); y = torch.empty(M, D, device=dev)
    timeit(lambda: F.gemm_f16(G, W2, epi=F.EPI_RES_F32, bias=bias, out0=y, aux=x), "fc2 RES gemm", flops=2.0 * M * D * H)
    timeit(lambda: F.gemm_f16(G, W2, epi=F.EPI_RES_F32, bias=bias, out0=y, aux=x, drop_p=0.1, drop_seed=99), "fc2 RES gemm, dropout 0.1", flops=2.0 * M * D * H)
    y16 = torch.empty(M, D, device=dev, dtype=torch.half)
    timeit(lambda: F.gemm_f16(G, W2, epi=F.EPI_F16, out0=y16), "dXn F16 gemm (K=2048)", flops=2.0 * M * D * H)
    del G, W2, x, y, y16
if "f16" in what:
    xn = (torch.randn(M, D, device=dev) * 0.5).half(); Wq = (torch.randn(3 * D, D, device=dev) * 0.05).half()
    qkv = torch.empty(M, 3 * D, device=dev, dtype=torch.half)
    timeit(lambda: F.gemm_f16(xn, Wq, epi=F.EPI_F16, out0=qkv), "QKV F16 gemm", flops=2.0 * M * 3 * D * D)
    del xn, Wq, qkv
if "split" in what:
    # precision mode "split": the same GEMMs with a (hi, lo) weight pair -- two MMAs per k-step on one A tile
    def pair(n, k, std=0.05):
        w = torch.randn(n, k, device=dev) * std
        hi = w.half()
        return hi, (w - hi.float()).half()
    A = (torch.randn(M, D, device=dev) * 0.5).half(); bias = torch.randn(H, device=dev)
    W, Wl = pair(H, D)
    o0 = torch.empty(M, H, device=dev, dtype=torch.half); o1 = torch.empty(M, H, device=dev, dtype=torch.half)
    timeit(lambda: F.gemm_f16(A, W, B_lo=Wl, epi=F.EPI_GELU, bias=bias, out0=o0, out1=o1, drop_p=0.1, drop_seed=1234), "SPLIT fc1 GELU gemm, dropout 0.1", flops=2.0 * M * H * D)
    timeit(lambda: F.gemm_f16(A, W, B_lo=Wl, epi=F.EPI_GELU_BWD, out0=o1, aux=o0), "SPLIT dH gemm (x saved gelu')", flops=2.0 * M * H * D)
    Wq, Wql = pair(3 * D, D)
    qkv = torch.empty(M, 3 * D, device=dev, dtype=torch.half)
    timeit(lambda: F.gemm_f16(A, Wq, B_lo=Wql, epi=F.EPI_F16, out0=qkv), "SPLIT QKV F16 gemm", flops=2.0 * M * 3 * D * D)
    timeit(lambda: F.gemm_f16(qkv, Wq.t().contiguous(), B_lo=Wql.t().contiguous(), epi=F.EPI_F16, out0=A), "SPLIT dLN1 F16 gemm (K=1536)", flops=2.0 * M * 3 * D * D)
    del qkv, Wq, Wql
    W2, W2l = pair(D, H)
    x = torch.randn(M, D, device=dev); y = torch.empty(M, D, device=dev); b2 = torch.randn(D, device=dev)
    timeit(lambda: F.gemm_f16(o1, W2, B_lo=W2l, epi=F.EPI_RES_F32, bias=b2, out0=y, aux=x, drop_p=0.1, drop_seed=99), "SPLIT fc2 RES gemm, dropout 0.1", flops=2.0 * M * D * H)
    y16 = torch.empty(M, D, device=dev, dtype=torch.half)
    timeit(lambda: F.gemm_f16(o1, W2, B_lo=W2l, epi=F.EPI_F16, out0=y16), "SPLIT dXn F16 gemm (K=2048)", flops=2.0 * M * D * H)
    Wo, Wol = pair(D, D)
    timeit(lambda: F.gemm_f16(A, Wo, B_lo=Wol, epi=F.EPI_RES_F32, bias=b2, out0=y, aux=x), "SPLIT out-proj RES gemm (K=512)", flops=2.0 * M * D * D)
    timeit(lambda: F.gemm_f16(A, Wo, epi=F.EPI_RES_F32, bias=b2, out0=y, aux=x), "plain out-proj RES gemm (K=512)", flops=2.0 * M * D * D)
    del A, W, Wl, o0, o1, W2, W2l, x, y, y16
if "split8" in what:
    # precision mode "split8": fp16(W 2^12) + e4m3 residual, the residual term on the FP8 tensor path
    def pair8(n, k, std=0.05):
        w = torch.randn(n, k, device=dev) * std
        hi = torch.empty(n, k, device=dev, dtype=torch.half); lo8 = torch.empty(n, k, device=dev, dtype=torch.uint8)
        F.check(F.lib().gsl_cast_f32_to_f16_split8(F.ptr(w), k, F.ptr(hi), F.ptr(lo8), k, n, k, 12, 0, F.cur_stream()))
        return hi, lo8
    A = (torch.randn(M, D, device=dev) * 0.5).half(); bias = torch.randn(H, device=dev)
    W, Wl = pair8(H, D)
    o0 = torch.empty(M, H, device=dev, dtype=torch.half); o1 = torch.empty(M, H, device=dev, dtype=torch.half)
    kw = dict(lo8_shift=12)
    w32 = torch.randn(H, D, device=dev) * 0.05; w16 = w32.half(); w16lo = (w32 - w16.float()).half()
    timeit(lambda: F.gemm_f16(A, w16, B_lo=w16lo, epi=F.EPI_GELU, bias=bias, out0=o0, out1=o1, drop_p=0.1, drop_seed=1234), "fc1 GELU gemm, dropout 0.1, fp16 residual (as the engine runs it in split8 mode)", flops=2.0 * M * H * D)
    timeit(lambda: F.gemm_f16(A, W, B_lo8=Wl, epi=F.EPI_GELU, bias=bias, out0=o0, out1=o1, drop_p=0.1, drop_seed=1234, **kw), "SPLIT8 fc1 GELU gemm, dropout 0.1 (not used by the engine)", flops=2.0 * M * H * D)
    timeit(lambda: F.gemm_f16(A, W, B_lo8=Wl, epi=F.EPI_GELU_BWD, out0=o1, aux=o0, **kw), "SPLIT8 dH gemm (x saved gelu')", flops=2.0 * M * H * D)
    Wq, Wql = pair8(3 * D, D)
    qkv = torch.empty(M, 3 * D, device=dev, dtype=torch.half)
    timeit(lambda: F.gemm_f16(A, Wq, B_lo8=Wql, epi=F.EPI_F16, out0=qkv, **kw), "SPLIT8 QKV F16 gemm", flops=2.0 * M * 3 * D * D)
    WqT, WqlT = pair8(D, 3 * D)
    timeit(lambda: F.gemm_f16(qkv, WqT, B_lo8=WqlT, epi=F.EPI_F16, out0=A, **kw), "SPLIT8 dLN1 F16 gemm (K=1536)", flops=2.0 * M * 3 * D * D)
    del qkv, Wq, Wql, WqT, WqlT
    W2, W2l = pair8(D, H)
    x = torch.randn(M, D, device=dev); y = torch.empty(M, D, device=dev); b2 = torch.randn(D, device=dev)
    timeit(lambda: F.gemm_f16(o1, W2, B_lo8=W2l, epi=F.EPI_RES_F32, bias=b2, out0=y, aux=x, drop_p=0.1, drop_seed=99, **kw), "SPLIT8 fc2 RES gemm, dropout 0.1", flops=2.0 * M * D * H)
    y16 = torch.empty(M, D, device=dev, dtype=torch.half)
    timeit(lambda: F.gemm_f16(o1, W2, B_lo8=W2l, epi=F.EPI_F16, out0=y16, **kw), "SPLIT8 dXn F16 gemm (K=2048)", flops=2.0 * M * D * H)
    Wo, Wol = pair8(D, D)
    timeit(lambda: F.gemm_f16(A, Wo, B_lo8=Wol, epi=F.EPI_RES_F32, bias=b2, out0=y, aux=x, **kw), "SPLIT8 out-proj RES gemm (K=512)", flops=2.0 * M * D * D)
    del A, W, Wl, o0, o1, W2, W2l, x, y, y16
if "attn" in what:
    qkv = torch.randn(M, 3 * D, device=dev).half(); out = torch.empty(M, D, device=dev, dtype=torch.half); lse = torch.empty(B * heads * N, device=dev)
    sc = 512 ** -0.5
    timeit(lambda: F.check(F.lib().gsl_attention_fwd(F.ptr(qkv), 3 * D, F.ptr(out), D, F.ptr(lse), B, N, heads, sc, F.cur_stream())), "attention fwd", flops=4.0 * B * heads * N * N * 64)
    dout = torch.randn(M, D, device=dev).half(); dqkv = torch.empty(M, 3 * D, device=dev, dtype=torch.half)
    timeit(lambda: F.check(F.lib().gsl_attention_bwd(F.ptr(qkv), 3 * D, F.ptr(out), D, F.ptr(dout), D, F.ptr(lse), F.ptr(dqkv), 3 * D, B, N, heads, sc, F.cur_stream())), "attention bwd", flops=14.0 * B * heads * N * N * 64)
    del qkv, out, dout, dqkv
if "skinny" in what:
    L = torch.randn(M, H, device=dev).half(); R = torch.randn(M, 16, device=dev).half()
    nb = F.lib().gsl_lora_side_workspace(M, H, 8); ws = torch.empty(nb // 4 + 16, device=dev); out = torch.zeros(H, 8, device=dev)
    timeit(lambda: F.check(F.lib().gsl_skinny_tn(F.ptr(L), H, F.ptr(R), 16, F.ptr(out), 8, 0, 1.0, 0, M, H, 8, F.ptr(ws), nb, F.cur_stream())), "skinny_tn N=2048", bytes_=M * H * 2.0)
    A16 = torch.randn(16, H, device=dev).half(); T = torch.empty(M, 16, device=dev, dtype=torch.half)
    timeit(lambda: F.check(F.lib().gsl_lora_down(F.ptr(L), H, F.ptr(A16), H, F.ptr(T), 16, M, H, 8, F.cur_stream())), "lora_down K=2048", bytes_=M * H * 2.0)
    timeit(lambda: F.check(F.lib().gsl_lora_side(F.ptr(L), H, F.ptr(A16), H, F.ptr(T), 16, F.ptr(R), 16, F.ptr(out), 8, 0, 1.0, 0, M, H, 8, F.ptr(ws), nb, F.cur_stream())), "lora_side N=2048 (both, one pass)", bytes_=M * H * 2.0)
    X = torch.randn(M, D, device=dev).half(); A5 = torch.randn(16, D, device=dev).half()
    timeit(lambda: F.check(F.lib().gsl_lora_down(F.ptr(X), D, F.ptr(A5), D, F.ptr(T), 16, M, D, 8, F.cur_stream())), "lora_down K=512", bytes_=M * D * 2.0)
    outd = torch.zeros(D, 8, device=dev)
    timeit(lambda: F.check(F.lib().gsl_skinny_tn(F.ptr(X), D, F.ptr(R), 16, F.ptr(outd), 8, 0, 1.0, 0, M, D, 8, F.ptr(ws), nb, F.cur_stream())), "skinny_tn N=512", bytes_=M * D * 2.0)
