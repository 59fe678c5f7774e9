import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "gs-lora_b200"), ROOT]
from gslora import _ffi as F
B, N, heads, scale = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), 0.125
D = heads * 64
torch.manual_seed(3)
qkv = torch.randn(B * N, 3 * D, device="cuda").half()
out = torch.empty(B * N, D, device="cuda", dtype=torch.half); lse = torch.empty(B * heads * N, device="cuda")
F.check(F.lib().gsl_attention_fwd(F.ptr(qkv), 3 * D, F.ptr(out), D, F.ptr(lse), B, N, heads, scale, F.cur_stream()))
torch.cuda.synchronize()
q, k, v = [t.float().reshape(B, N, heads, 64).permute(0, 2, 1, 3) for t in qkv.chunk(3, dim=-1)]
dots = torch.einsum("bhid,bhjd->bhij", q, k) * scale
ref = torch.einsum("bhij,bhjd->bhid", dots.softmax(-1), v).permute(0, 2, 1, 3).reshape(B * N, D)
print("rel err", float((out.float() - ref).norm() / ref.norm()), "lse err", float((lse.view(B, heads, N) - torch.logsumexp(dots, -1)).abs().max()))
