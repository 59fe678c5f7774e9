"""Dev (CPU): engine-faithful emulation of the SPLIT-precision arithmetic (weights exact, every fp16 tensor the engine materialises rounded to
an 11-bit significand) with one switch per rounding site, to rank what sets the remaining LoRA-gradient error.

    SEED=1 python scripts/dev_precision_emul2.py [B]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import torch.nn.functional as F
from oracle import vit_oracle as O

torch.set_num_threads(os.cpu_count())
SW = {}          # rounding switches, set per run


def r11(t):
    m, e = torch.frexp(t)
    return torch.ldexp(torch.round(m * 2048.0) / 2048.0, e)


def rnd(t, key):
    return r11(t) if SW[key] else t


class RoundGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, key):
        ctx.key = key
        return x.clone()

    @staticmethod
    def backward(ctx, g):
        return rnd(g, ctx.key), None


class RoundVal(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, key):
        return rnd(x, key)

    @staticmethod
    def backward(ctx, g):
        return g, None


class FC(torch.autograd.Function):
    """y = xq W'^T + b with W' = W + s B A exact; backward as the engine: dx = dyq W', U = fp16(dyq B), dA = s U^T xq, T = fp16(xq A^T), dB = s dyq^T T"""
    @staticmethod
    def forward(ctx, x, W, b, A, B, s, kx, kdy):
        xq = rnd(x, kx)
        ctx.save_for_backward(xq, W, A if A is not None else torch.zeros(0), B if B is not None else torch.zeros(0))
        ctx.s, ctx.kdy, ctx.lora, ctx.has_b = s, kdy, A is not None, b is not None
        Wm = W + s * (B @ A) if A is not None else W
        y = xq @ Wm.t()
        return y + b if b is not None else y

    @staticmethod
    def backward(ctx, dy):
        xq, W, A, B = ctx.saved_tensors
        dyq = rnd(dy, ctx.kdy)
        Wm = W + ctx.s * (B @ A) if ctx.lora else W
        dx = dyq @ Wm
        dA = dB = None
        if ctx.lora:
            x2, d2 = xq.reshape(-1, xq.shape[-1]), dyq.reshape(-1, dyq.shape[-1])
            U = rnd(d2 @ B, "tu")
            T = rnd(x2 @ A.t(), "tu")
            dA = ctx.s * (U.t() @ x2)
            dB = ctx.s * (d2.t() @ T)
        return dx, None, (dy.sum(tuple(range(dy.dim() - 1))) if ctx.has_b else None), dA, dB, None, None, None


class Gelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h):
        ctx.save_for_backward(h)
        return F.gelu(h)

    @staticmethod
    def backward(ctx, dg):
        (h,) = ctx.saved_tensors
        cdf = 0.5 * (1 + torch.erf(h * 0.7071067811865476))
        gp = cdf + h * torch.exp(-0.5 * h * h) * 0.3989422804014327
        return dg * rnd(gp, "gp")


class Attn(torch.autograd.Function):
    """q, k, v [b, h, n, d] (already fp16 values).  Engine: P fp16 for P V and dV, O fp16, dO fp16, delta from the fp16 O / dO, dS fp16."""
    @staticmethod
    def forward(ctx, q, k, v, scale):
        P = (torch.einsum("bhid,bhjd->bhij", q, k) * scale).softmax(dim=-1)
        Pq = rnd(P, "p")
        Oo = torch.einsum("bhij,bhjd->bhid", Pq, v)
        ctx.save_for_backward(q, k, v, P, Pq, Oo)
        ctx.scale = scale
        return Oo

    @staticmethod
    def backward(ctx, dO):
        q, k, v, P, Pq, Oo = ctx.saved_tensors
        dOq = rnd(dO, "do")
        Oq = rnd(Oo, "o_delta")
        delta = (dOq * Oq).sum(-1, keepdim=True)
        dP = torch.einsum("bhid,bhjd->bhij", dOq, v)
        dS = rnd((Pq if SW["p_bwd"] else P) * (dP - delta), "ds") * ctx.scale
        dq = torch.einsum("bhij,bhjd->bhid", dS, k)
        dk = torch.einsum("bhij,bhid->bhjd", dS, q)
        dv = torch.einsum("bhij,bhid->bhjd", Pq if SW["p_bwd"] else P, dOq)
        return dq, dk, dv, None


def forward(sd, cfg, img, label):
    D = cfg.dim
    x = O.patchify(img.float(), cfg.patch_size)
    x = FC.apply(x, sd["patch_to_embedding.weight"], sd["patch_to_embedding.bias"], None, None, 0.0, "x", "dy")
    b, n, _ = x.shape
    x = torch.cat((sd["cls_token"].expand(b, -1, -1), x), dim=1) + sd["pos_embedding"][:, : n + 1]
    s = cfg.lora_scaling
    for i in range(cfg.depth):
        xn = F.layer_norm(x, (D,), sd[O.blk(i, "0.fn.norm.weight")], sd[O.blk(i, "0.fn.norm.bias")], cfg.ln_eps)
        xn = RoundGrad.apply(xn, "dxn")
        qkv = FC.apply(xn, sd[O.blk(i, "0.fn.fn.to_qkv.weight")], None, None, None, 0.0, "x", "dy")
        qkv = RoundGrad.apply(RoundVal.apply(qkv, "x"), "dy")
        q, k, v = [t.reshape(b, n + 1, cfg.heads, -1).permute(0, 2, 1, 3) for t in qkv.chunk(3, dim=-1)]
        out = Attn.apply(q, k, v, cfg.attn_scale).permute(0, 2, 1, 3).reshape(b, n + 1, -1)
        x = FC.apply(out, sd[O.blk(i, "0.fn.fn.to_out.0.weight")], sd[O.blk(i, "0.fn.fn.to_out.0.bias")], None, None, 0.0, "x", "dy") + x
        xn = F.layer_norm(x, (D,), sd[O.blk(i, "1.fn.norm.weight")], sd[O.blk(i, "1.fn.norm.bias")], cfg.ln_eps)
        xn = RoundGrad.apply(xn, "dxn")
        h = FC.apply(xn, sd[O.blk(i, "1.fn.fn.net.0.weight")], sd[O.blk(i, "1.fn.fn.net.0.bias")], sd[O.blk(i, "1.fn.fn.net.0.lora_A")],
                     sd[O.blk(i, "1.fn.fn.net.0.lora_B")], s, "x", "dy")
        g = Gelu.apply(h)
        x = FC.apply(g, sd[O.blk(i, "1.fn.fn.net.3.weight")], sd[O.blk(i, "1.fn.fn.net.3.bias")], sd[O.blk(i, "1.fn.fn.net.3.lora_A")],
                     sd[O.blk(i, "1.fn.fn.net.3.lora_B")], s, "x", "dy") + x
    emb = F.layer_norm(x[:, 0], (D,), sd["mlp_head.0.weight"], sd["mlp_head.0.bias"], cfg.ln_eps)
    return O.cosface(emb, sd["loss.weight"], label, cfg.cos_s, cfg.cos_m)


def grads(sd, cfg, batch):
    xr, yr, xf, yf = batch
    names = O.lora_param_list(cfg)
    work = {k: v.detach().clone() for k, v in sd.items()}
    for nme in names:
        work[nme].requires_grad_(True)
    lr_ = forward(work, cfg, xr, yr)
    lf_ = forward(work, cfg, xf, yf)
    total = F.cross_entropy(lr_, yr) + 0.15 * F.relu(105.0 - F.cross_entropy(lf_, yf))
    gs = torch.autograd.grad(total, [work[nme] for nme in names])
    return lr_.detach(), {nme: g for nme, g in zip(names, gs)}


def rel(a, b):
    return float((a - b).norm() / b.norm())


KEYS = ["x", "dy", "tu", "gp", "p", "p_bwd", "do", "o_delta", "ds", "dxn"]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    seed = int(os.environ.get("SEED", "1"))
    cfg = O.P8S8
    sd = O.init_state_dict(cfg, seed=seed)
    gen = torch.Generator().manual_seed(7)
    batch = (torch.rand(B, 3, 112, 112, generator=gen), torch.randint(0, 100, (B,), generator=gen),
             torch.rand(B, 3, 112, 112, generator=gen), torch.randint(0, 100, (B,), generator=gen))
    names = O.lora_param_list(cfg)
    SW.update({k: False for k in KEYS})
    ref_logits, ref = grads(sd, cfg, batch)

    def report(tag):
        lg, g = grads(sd, cfg, batch)
        per = {n: rel(g[n], ref[n]) for n in names}
        allr = rel(torch.cat([g[n].flatten() for n in names]), torch.cat([ref[n].flatten() for n in names]))
        wn = max(per, key=per.get)
        a1 = [per[n] for n in names if n.endswith("net.0.lora_A")]
        print(f"{tag:34s} logits {rel(lg, ref_logits):.2e}  grads all {allr:.2e}  worst {per[wn]:.2e} ({wn.split('.')[2]}.{wn.split('.')[-2]}.{wn.split('.')[-1]})  "
              f"mean fc1.lora_A {sum(a1) / len(a1):.2e}", flush=True)

    SW.update({k: True for k in KEYS})
    report("all engine roundings")
    for k in KEYS:
        SW.update({kk: True for kk in KEYS})
        SW[k] = False
        report(f"all but '{k}'")
    for k in ["x", "dy", "tu"]:
        SW.update({kk: False for kk in KEYS})
        SW[k] = True
        report(f"only '{k}'")


if __name__ == "__main__":
    main()
