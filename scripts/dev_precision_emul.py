"""Dev (CPU): which operand roundings set the LoRA-gradient error of the fp16-operand engine?

Emulates the engine's arithmetic on the oracle: every dense contraction takes operands rounded to an 11-bit significand (fp16 without
the exponent range, i.e. what the loss-scaled engine sees) and accumulates in fp32.  Weight rounding and activation / gradient rounding
are switched independently, per GEMM family and per direction, and the LoRA gradients are compared with the FP32 oracle.

    python scripts/dev_precision_emul.py [B]
"""
import os, sys, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import torch.nn.functional as F
from oracle import vit_oracle as O

torch.set_num_threads(os.cpu_count())


def r11(t):
    m, e = torch.frexp(t)
    return torch.ldexp(torch.round(m * 2048.0) / 2048.0, e)


class RoundST(torch.autograd.Function):
    """value: rounded (if fwd); gradient: rounded (if bwd)"""
    @staticmethod
    def forward(ctx, x, fwd, bwd):
        ctx.bwd = bwd
        return r11(x) if fwd else x.clone()

    @staticmethod
    def backward(ctx, g):
        return (r11(g) if ctx.bwd else g), None, None


class QLinear(torch.autograd.Function):
    """y = xq Wf^T (+ b), dx = dyq Wb; dW = dyq^T xq flows to the (differentiable) merged weight -> dA / dB."""
    @staticmethod
    def forward(ctx, x, W, b, w_fwd, w_bwd, act):
        act_f, act_b = (act if isinstance(act, tuple) else (act, act))
        act = act_b
        xq = r11(x) if act_f else x
        Wf = r11(W) if w_fwd else W
        ctx.save_for_backward(xq, W)
        ctx.w_bwd, ctx.act, ctx.has_b = w_bwd, act, b is not None
        y = xq @ Wf.t()
        return y + b if b is not None else y

    @staticmethod
    def backward(ctx, dy):
        xq, W = ctx.saved_tensors
        dyq = r11(dy) if ctx.act else dy
        Wb = r11(W) if ctx.w_bwd else W
        dx = dyq @ Wb
        dW = dyq.reshape(-1, dyq.shape[-1]).t() @ xq.reshape(-1, xq.shape[-1])
        return dx, dW, (dy.sum(tuple(range(dy.dim() - 1))) if ctx.has_b else None), None, None, None


def forward(sd, cfg, img, label, knobs):
    """knobs: dict family -> (w_fwd, w_bwd) for qkv/out/fc1/fc2/patch; 'act': bool"""
    D = cfg.dim
    act = knobs["act"]
    x = O.patchify(img.float(), cfg.patch_size)
    x = QLinear.apply(x, sd["patch_to_embedding.weight"], sd["patch_to_embedding.bias"], *knobs["patch"], act)
    b, n, _ = x.shape
    x = torch.cat((sd["cls_token"].expand(b, -1, -1), x), dim=1) + sd["pos_embedding"][:, : n + 1]
    s = cfg.lora_scaling
    for i in range(cfg.depth):
        xn = F.layer_norm(x, (D,), sd[O.blk(i, "0.fn.norm.weight")], sd[O.blk(i, "0.fn.norm.bias")], cfg.ln_eps)
        qkv = QLinear.apply(xn, sd[O.blk(i, "0.fn.fn.to_qkv.weight")], None, *knobs["qkv"], act)
        af, ab = (act if isinstance(act, tuple) else (act, act))
        qkv = RoundST.apply(qkv, af, ab)
        q, k, v = [t.reshape(b, n + 1, cfg.heads, -1).permute(0, 2, 1, 3) for t in qkv.chunk(3, dim=-1)]
        dots = torch.einsum("bhid,bhjd->bhij", q, k) * cfg.attn_scale
        attn = RoundST.apply(dots.softmax(dim=-1), af, ab)
        out = torch.einsum("bhij,bhjd->bhid", attn, v).permute(0, 2, 1, 3).reshape(b, n + 1, -1)
        x = QLinear.apply(out, sd[O.blk(i, "0.fn.fn.to_out.0.weight")], sd[O.blk(i, "0.fn.fn.to_out.0.bias")], *knobs["out"], act) + x
        xn = F.layer_norm(x, (D,), sd[O.blk(i, "1.fn.norm.weight")], sd[O.blk(i, "1.fn.norm.bias")], cfg.ln_eps)
        W1 = sd[O.blk(i, "1.fn.fn.net.0.weight")] + s * sd[O.blk(i, "1.fn.fn.net.0.lora_B")] @ sd[O.blk(i, "1.fn.fn.net.0.lora_A")]
        W2 = sd[O.blk(i, "1.fn.fn.net.3.weight")] + s * sd[O.blk(i, "1.fn.fn.net.3.lora_B")] @ sd[O.blk(i, "1.fn.fn.net.3.lora_A")]
        h = QLinear.apply(xn, W1, sd[O.blk(i, "1.fn.fn.net.0.bias")], *knobs["fc1"], act)
        g = F.gelu(h)
        x = QLinear.apply(g, W2, sd[O.blk(i, "1.fn.fn.net.3.bias")], *knobs["fc2"], act) + x
    emb = F.layer_norm(x[:, 0], (D,), sd["mlp_head.0.weight"], sd["mlp_head.0.bias"], cfg.ln_eps)
    return O.cosface(emb, sd["loss.weight"], label, cfg.cos_s, cfg.cos_m)


def grads(sd, cfg, batch, knobs):
    xr, yr, xf, yf = batch
    names = O.lora_param_list(cfg)
    work = {k: v.detach().clone() for k, v in sd.items()}
    for nme in names:
        work[nme].requires_grad_(True)
    lr_ = forward(work, cfg, xr, yr, knobs)
    lf_ = forward(work, cfg, xf, yf, knobs)
    total = F.cross_entropy(lr_, yr) + 0.15 * F.relu(105.0 - F.cross_entropy(lf_, yf))
    gs = torch.autograd.grad(total, [work[nme] for nme in names])
    return lr_.detach(), {nme: g for nme, g in zip(names, gs)}


def rel(a, b):
    return float((a - b).norm() / b.norm())


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    seed = int(os.environ.get("SEED", "1337"))
    cfg = O.P8S8
    sd = O.init_state_dict(cfg, seed=seed)
    gen = torch.Generator().manual_seed(7)
    batch = (torch.rand(B, 3, 112, 112, generator=gen), torch.randint(0, 100, (B,), generator=gen),
             torch.rand(B, 3, 112, 112, generator=gen), torch.randint(0, 100, (B,), generator=gen))
    fams = ["patch", "qkv", "out", "fc1", "fc2"]
    exact = {f: (False, False) for f in fams}
    exact["act"] = False
    ref_logits, ref = grads(sd, cfg, batch, exact)
    names = O.lora_param_list(cfg)

    def report(tag, knobs):
        lg, g = grads(sd, cfg, batch, knobs)
        per = [rel(g[n], ref[n]) for n in names]
        allr = rel(torch.cat([g[n].flatten() for n in names]), torch.cat([ref[n].flatten() for n in names]))
        print(f"{tag:44s} logits {rel(lg, ref_logits):.2e}  grads all {allr:.2e}  worst {max(per):.2e}  mean {sum(per)/len(per):.2e}", flush=True)

    allq = {f: (True, True) for f in fams}; allq["act"] = True
    report("everything rounded (engine today)", allq)
    k = {f: (False, False) for f in fams}; k["act"] = True
    report("weights exact, activations rounded", k)
    k = {f: (True, True) for f in fams}; k["act"] = False
    report("weights rounded fwd+bwd, activations exact", k)
    k = {f: (False, True) for f in fams}; k["act"] = True
    report("weights exact in fwd only", k)
    k = {f: (True, False) for f in fams}; k["act"] = True
    report("weights exact in bwd only", k)
    if os.environ.get("PERFAMILY"):
        k = {f: (False, False) for f in fams}; k["act"] = True
        report("weights exact everywhere (split mode)", k)
        for fam in ["qkv", "out", "fc1", "fc2", "patch"]:
            for which, tag in (((True, False), "FWD"), ((False, True), "BWD")):
                k = {f: (False, False) for f in fams}; k["act"] = True
                k[fam] = which
                report(f"single rounding only in {tag} of {fam}", k)
        return
    if os.environ.get("ACTSPLIT"):
        k = {f: (False, False) for f in fams}; k["act"] = (True, False)
        report("weights exact, FORWARD activations rounded only", k)
        k = {f: (False, False) for f in fams}; k["act"] = (False, True)
        report("weights exact, BACKWARD gradients rounded only", k)
        return
    if os.environ.get("QUICK"):
        return
    for fam in ["qkv", "out", "fc1", "fc2"]:
        k = {f: (True, True) for f in fams}; k["act"] = True
        k[fam] = (False, True)
        report(f"fwd-exact weights for {fam} only", k)
    for combo in [("qkv", "fc1"), ("qkv", "out", "fc1"), ("fc1", "fc2"), ("qkv", "fc1", "fc2")]:
        k = {f: (True, True) for f in fams}; k["act"] = True
        for fam in combo:
            k[fam] = (False, True)
        report(f"fwd-exact weights for {'+'.join(combo)}", k)


if __name__ == "__main__":
    main()
