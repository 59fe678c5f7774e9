"""Dev: the P8S8 parity numbers (logits, per-tensor LoRA gradient errors) of tests/test_engine_gpu.py for both precision modes and the
five weight seeds, printed.   SEEDS="1337 1 2 3 4" MODES="split fast" B=32 python tests/dev_parity.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "gs-lora_b200"), ROOT, os.path.join(ROOT, "tests")]
import torch
from oracle import vit_oracle as O
from test_engine_gpu import build_model, rel
torch.backends.cuda.matmul.allow_tf32 = False
cfg = O.P8S8
B = int(os.environ.get("B", "32"))
verbose = os.environ.get("VERBOSE", "0") == "1"
for seed in [int(s) for s in os.environ.get("SEEDS", "1337 1 2 3 4").split()]:
    sd = O.init_state_dict(cfg, seed=seed)
    gen = torch.Generator().manual_seed(7)
    xr, xf = torch.rand(B, 3, 112, 112, generator=gen).cuda(), torch.rand(B, 3, 112, 112, generator=gen).cuda()
    yr, yf = torch.randint(0, 100, (B,), generator=gen).cuda(), torch.randint(0, 100, (B,), generator=gen).cuda()
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    ref, ref_grads = O.unlearn_grads(sd_gpu, cfg, xr, yr, xf, yf, beta=0.15, alpha=1e-4, BND=105.0, include_structure=False)
    for mode in os.environ.get("MODES", "split fast").split():
        model = build_model(cfg, sd)
        model.gsl_precision = mode
        crit = torch.nn.CrossEntropyLoss()
        out_r, _ = model(xr, yr)
        out_f, _ = model(xf, yf)
        total = torch.relu(105.0 - crit(out_f, yf)) * 0.15 + crit(out_r, yr)
        total.backward()
        names = O.lora_param_list(cfg)
        per = {n: rel(model.get_parameter(n).grad, ref_grads[n]) for n in names}
        allrel = rel(torch.cat([model.get_parameter(n).grad.flatten() for n in names]), torch.cat([ref_grads[n].flatten() for n in names]))
        print(f"seed {seed:5d} {mode:5s} B={B}: logits {rel(out_r, ref['logits_r']):.3e}/{rel(out_f, ref['logits_f']):.3e}  grads all {allrel:.3e}  "
              f"worst {max(per.values()):.3e}", flush=True)
        if verbose:
            for n in names:
                print(f"  {per[n]:.3e}  |g|={float(ref_grads[n].norm()):.3e}  {n}")
        del model
        torch.cuda.empty_cache()
