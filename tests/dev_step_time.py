"""Dev harness (GPU box): parity table + step timing of the fused unlearning step at the config-2 shape."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "gs-lora_b200"), ROOT]
import torch
from oracle import vit_oracle as O
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_engine_gpu import build_model, rel
import engine_cl

torch.backends.cuda.matmul.allow_tf32 = False
cfg = O.P8S8
sd = O.init_state_dict(cfg, seed=1337)
names = O.lora_param_list(cfg)

def parity(B):
    gen = torch.Generator().manual_seed(7)
    xr, xf = torch.rand(B, 3, 112, 112, generator=gen).cuda(), torch.rand(B, 3, 112, 112, generator=gen).cuda()
    yr, yf = torch.randint(0, 100, (B,), generator=gen).cuda(), torch.randint(0, 100, (B,), generator=gen).cuda()
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    ref, rg = O.unlearn_grads(sd_gpu, cfg, xr, yr, xf, yf, beta=0.15, alpha=1e-4, BND=105.0, include_structure=False)
    model = build_model(cfg, sd)
    crit = torch.nn.CrossEntropyLoss()
    out_r, _ = model(xr, yr); out_f, _ = model(xf, yf)
    (torch.relu(105.0 - crit(out_f, yf)) * 0.15 + crit(out_r, yr)).backward()
    per = {n: rel(model.get_parameter(n).grad, rg[n]) for n in names}
    allrel = rel(torch.cat([model.get_parameter(n).grad.flatten() for n in names]), torch.cat([rg[n].flatten() for n in names]))
    print(f"B={B}: logits {rel(out_r, ref['logits_r']):.2e}/{rel(out_f, ref['logits_f']):.2e} grads all {allrel:.2e} worst {max(per.values()):.2e} mean {sum(per.values())/len(per):.2e}")
    for n in names: print(f"    {n[19:]:40s} {per[n]:.2e}  |g| {float(rg[n].norm()):.3e}")
    gmax = max(float(model.get_parameter(n).grad.abs().max()) for n in names)
    print("    finite:", all(torch.isfinite(model.get_parameter(n).grad).all() for n in names), "max|g|", gmax)
    del model
    torch.cuda.empty_cache()

def timing(B, steps=int(os.environ.get('STEPS', '5'))):
    model = build_model(cfg, sd)
    gen = torch.Generator().manual_seed(9)
    xr, xf = torch.rand(B, 3, 112, 112, generator=gen).cuda(), torch.rand(B, 3, 112, 112, generator=gen).cuda()
    yr, yf = torch.randint(0, 100, (B,), generator=gen).cuda(), torch.randint(0, 100, (B,), generator=gen).cuda()
    hp = dict(lr=1e-2, wd=0.05)
    for _ in range(int(os.environ.get('WARM', '2'))):
        engine_cl.unlearn_step(model, xr, yr, xf, yf, beta=0.15, alpha=1e-4, BND=105.0, hparams=hp)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e[0].record()
    for _ in range(steps):
        out = engine_cl.unlearn_step(model, xr, yr, xf, yf, beta=0.15, alpha=1e-4, BND=105.0, hparams=hp)
    e[1].record(); torch.cuda.synchronize()
    ms = e[0].elapsed_time(e[1]) / steps
    fl = O.flops_per_image(cfg)["total"] * 2 * B
    print(f"fused step B={B}+{B}: {ms:.2f} ms/step  {2*B/ms*1e3:.0f} img/s  {fl/ms/1e9:.1f} TFLOP/s  loss {out['total']:.4f}")
    # phase split
    eng = model._engine
    img = torch.cat([xr, xf]); lab = torch.cat([yr, yf])
    dl = torch.randn(2 * B, 100, device="cuda") * 1e-3
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record(); eng.forward(img, lab, 0, True); ev[1].record(); eng.backward(0, dl, None, False); ev[2].record()
    eng.optimizer_step(1e-2, 0.05, 1e-4); ev[3].record(); torch.cuda.synchronize()
    print(f"   forward {ev[0].elapsed_time(ev[1]):.2f} ms  backward {ev[1].elapsed_time(ev[2]):.2f} ms  optimizer+repack {ev[2].elapsed_time(ev[3]):.3f} ms")

if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "parity"):
        for B in (8, 64):
            parity(B)
    if what in ("all", "time"):
        timing(int(os.environ.get("B", "512")))
