"""Dev harness (GPU box): run each GEMM case in its own subprocess so a trap in one configuration does
not take the others down.  python scripts/dev_gemm_check.py [case ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gs-lora_b200"))

CASES = {
    # name: (M, N, K, epi, cta_group, block_n)
    "cg1_f16_small": (256, 256, 128, 0, 1, 128),
    "cg1_f16_bn256": (384, 512, 512, 0, 1, 256),
    "cg2_f16_small": (256, 256, 128, 0, 2, 128),
    "cg2_f16_bn256": (512, 512, 512, 0, 2, 256),
    "cg1_f16_ragged": (1000, 384, 528, 0, 1, 128),
    "cg2_f16_ragged": (1000, 1536, 528, 0, 2, 256),
    "cg1_f32": (300, 512, 2064, 1, 1, 256),
    "cg2_f32_o1": (300, 512, 2064, 1, 2, 256),
    "cg1_gelu": (520, 2048, 528, 2, 1, 256),
    "cg2_gelu": (520, 2048, 528, 2, 2, 256),
    "cg1_gelu_bwd": (520, 2048, 528, 3, 1, 256),
    "cg2_gelu_bwd": (520, 2048, 528, 3, 2, 256),
    "cg1_res": (777, 512, 2064, 4, 1, 256),
    "cg2_res": (777, 512, 2064, 4, 2, 256),
    "cg2_res_bn128": (777, 128, 272, 4, 2, 128),
    "cg1_periodic": (26 * 5, 128, 192, 5, 1, 128),
    "cg2_periodic": (197 * 3, 512, 192, 5, 2, 256),
    "cg2_big": (100864, 2048, 528, 2, 2, 256),
    "cg1_big": (100864, 2048, 528, 2, 1, 256),
}


def run_case(name):
    import torch
    from gslora import _ffi as F
    M, N, K, epi, cg, bn = CASES[name]
    torch.manual_seed(0)
    dev = "cuda"
    A = (torch.randn(M, K, device=dev) * 0.5).half()
    B = (torch.randn(N, K, device=dev) * 0.05).half()
    bias = torch.randn(N, device=dev)
    ref = A.float() @ B.float().t() + bias
    out1 = None
    aux = None
    period = 0
    if epi == 0:
        out0 = torch.full((M, N), 7.0, device=dev, dtype=torch.half)
        want0, want1 = ref, None
    elif epi == 1:
        out0 = torch.full((M, N), 7.0, device=dev)
        out1 = torch.full((M, N), 7.0, device=dev, dtype=torch.half) if "o1" in name else None
        want0, want1 = ref, ref if out1 is not None else None
    elif epi == 2:
        out0 = torch.empty(M, N, device=dev, dtype=torch.half)
        out1 = torch.empty(M, N, device=dev, dtype=torch.half)
        want0, want1 = ref, torch.nn.functional.gelu(ref)
    elif epi == 3:
        aux = torch.randn(M, N, device=dev).half()
        out0 = torch.empty(M, N, device=dev, dtype=torch.half)
        h = aux.float().requires_grad_(True)
        torch.nn.functional.gelu(h).sum().backward()
        bias = None
        want0, want1 = (A.float() @ B.float().t()) * h.grad, None
    elif epi == 4:
        aux = torch.randn(M, N, device=dev)
        out0 = torch.empty(M, N, device=dev)
        want0 = ref + aux
        want1 = None
    elif epi == 5:
        period = 26 if M % 26 == 0 else 197
        aux = torch.randn(period, N, device=dev)
        out0 = torch.empty(M, N, device=dev)
        bias = None
        want0 = (A.float() @ B.float().t()) + aux.repeat(M // period, 1)
        want1 = None
    F.gemm_f16(A, B, epi=epi, bias=bias, out0=out0, out1=out1, aux=aux, aux_period=period, cta_group=cg, block_n=bn)
    torch.cuda.synchronize()
    e0 = (out0.float() - want0).abs().max().item()
    s0 = want0.abs().max().item()
    msg = f"{name}: out0 maxerr {e0:.3e} (scale {s0:.2f})"
    ok = e0 <= 2e-3 * s0 + 1e-3
    if want1 is not None:
        e1 = (out1.float() - want1).abs().max().item()
        msg += f" out1 maxerr {e1:.3e}"
        ok = ok and e1 <= 2e-3 * want1.abs().max().item() + 1e-3
    if M >= 50000:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            F.gemm_f16(A, B, epi=epi, bias=bias, out0=out0, out1=out1, aux=aux, aux_period=period, cta_group=cg, block_n=bn)
        ev0.record()
        for _ in range(10):
            F.gemm_f16(A, B, epi=epi, bias=bias, out0=out0, out1=out1, aux=aux, aux_period=period, cta_group=cg, block_n=bn)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / 10
        msg += f" | {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s"
    print(("PASS " if ok else "FAIL ") + msg, flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        sys.exit(run_case(sys.argv[2]))
    names = sys.argv[1:] or list(CASES)
    bad = 0
    for n in names:
        try:
            r = subprocess.run([sys.executable, __file__, "--one", n], timeout=120, capture_output=True, text=True)
            out = (r.stdout + r.stderr).strip().splitlines()
            tail = [l for l in out if l.startswith(("PASS", "FAIL", "gslora"))] or out[-6:]
            print("\n".join(tail[:8]), flush=True)
            if r.returncode != 0:
                bad += 1
                if not any(l.startswith("FAIL") for l in tail):
                    print(f"CRASH {n}: rc={r.returncode}", flush=True)
        except subprocess.TimeoutExpired:
            bad += 1
            print(f"TIMEOUT {n}", flush=True)
    print(f"{len(names) - bad}/{len(names)} cases passed")
    sys.exit(1 if bad else 0)
