"""Dev (GPU box): phase timeline of attention_fwd_tc_kernel's CTA 0 from a -DGSL_ATTN_TRACE build of the library.
build: make -C gs-lora_b200/csrc NVFLAGS_EXTRA=-DGSL_ATTN_TRACE OUT=../lib/libgslora_trace.so OBJDIR=../lib/obj_trace"""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L = ctypes.CDLL(os.path.join(ROOT, "gs-lora_b200", "lib", "libgslora_trace.so"))
B, N, heads = 1024, 197, 8
D = heads * 64
qkv = torch.randn(B * N, 3 * D, device="cuda").half()
out = torch.empty(B * N, D, device="cuda", dtype=torch.half); lse = torch.empty(B * heads * N, device="cuda")
P = ctypes.c_void_p
L.gsl_attention_fwd.argtypes = [P, ctypes.c_int64, P, ctypes.c_int64, P, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, P]
for _ in range(2):
    rc = L.gsl_attention_fwd(qkv.data_ptr(), 3 * D, out.data_ptr(), D, lse.data_ptr(), B, N, heads, 0.125, None)
    torch.cuda.synchronize()
assert rc == 0
buf = (ctypes.c_longlong * (16 * 64 * 4))()
sym = getattr(L, "gsl_debug_attn_trace", None) or getattr(L, "_ZN3gsl20gsl_debug_attn_traceEPx")
assert sym(buf) == 0
t = torch.tensor(list(buf)).view(16, 64, 4)
t0 = int(t[13, 0, 0])
rel = lambda x: int(x) - t0
print("item | MMA: loop_top S_issue pv_issue | grp warp0/4: wait_S S_full turn exp_done | writer: O_full O_free")
for j in range(24):
    w = 0 if j % 2 == 0 else 4
    print(f"{j:3d} | {rel(t[13,j,0]):7d} {rel(t[13,j,1]):7d} {rel(t[13,j,2]):7d} | {rel(t[w,j,0]):7d} {rel(t[w,j,1]):7d} {rel(t[w,j,2]):7d} {rel(t[w,j,3]):7d} | {rel(t[8,j,0]):7d} {rel(t[8,j,1]):7d}")
