import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "gs-lora_b200"), ROOT]
from gslora import _ffi as F
torch.manual_seed(0)
M, N, K = 4096, 2048, 512
A = (torch.randn(M, K, device="cuda") * 1.0).half(); B = ((torch.rand(N, K, device="cuda") * 2 - 1) / K ** 0.5).half(); bias = torch.randn(N, device="cuda") * 0.05
o0 = torch.empty(M, N, device="cuda", dtype=torch.half); o1 = torch.empty(M, N, device="cuda", dtype=torch.half)
F.gemm_f16(A, B, epi=F.EPI_GELU, bias=bias, out0=o0, out1=o1)
h = A.double() @ B.double().t() + bias.double()
g = torch.nn.functional.gelu(h)
rel = lambda a, b: float((a.double() - b).norm() / b.norm())
print("G rel err vs fp64:", rel(o1, g), " fp16-rounded exact:", rel(g.half(), g))
x = torch.randn(M, 512, device="cuda"); y = torch.empty(M, 512, device="cuda")
A2 = (torch.randn(M, 2048, device="cuda") * 0.5).half(); B2 = ((torch.rand(512, 2048, device="cuda") * 2 - 1) / 2048 ** 0.5).half(); b2 = torch.randn(512, device="cuda") * 0.05
F.gemm_f16(A2, B2, epi=F.EPI_RES_F32, bias=b2, out0=y, aux=x)
ref = A2.double() @ B2.double().t() + b2.double() + x.double()
print("RES rel err:", rel(y, ref))
