"""Size-independent properties at BASELINE.json's FULL config-2 size (ViT-P8S8, 512 + 512 images per step), where the FP32 oracle is too
slow / too large to run beside the engine: batch invariance of the forward, additivity of the selective backward over the batch, the
remain / forget split of the fused step, idempotence of a repeated forward, and ragged / single-image batches."""
import pytest
import torch

from oracle import vit_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.fixture(scope="module")
def p8s8():
    from test_engine_gpu import build_model
    cfg = O.P8S8
    sd = O.init_state_dict(cfg, seed=1337)
    model = build_model(cfg, sd)
    gen = torch.Generator().manual_seed(99)
    x = torch.rand(1024, 3, 112, 112, generator=gen).cuda()
    y = torch.randint(0, 100, (1024,), generator=gen).cuda()
    return cfg, sd, model, x, y


def test_forward_is_batch_invariant_and_idempotent_at_full_size(p8s8):
    """Row b of a 1024-image forward == the same image alone / in a ragged 37-image batch, bit for bit (no cross-sample reduction anywhere in
    the forward, tiles never mix rows); running the full batch twice gives identical bits (deterministic kernels)."""
    cfg, sd, model, x, y = p8s8
    with torch.no_grad():
        full_logits, full_emb = model(x, y)
        again_logits, again_emb = model(x, y)
        assert torch.equal(full_logits, again_logits) and torch.equal(full_emb, again_emb)
        for lo, hi in ((0, 1), (1023, 1024), (500, 537), (3, 260)):
            l, e = model(x[lo:hi].contiguous(), y[lo:hi].contiguous())
            assert torch.equal(l, full_logits[lo:hi]) and torch.equal(e, full_emb[lo:hi]), (lo, hi)
    assert torch.isfinite(full_logits).all() and torch.isfinite(full_emb).all()
    assert full_logits.shape == (1024, cfg.num_class) and full_emb.shape == (1024, cfg.dim)


def test_backward_is_additive_over_the_batch_at_full_size(p8s8):
    """The LoRA gradient of sum_b w_b CE_b over 1024 images == the sum of the gradients of its two halves (fp32 reductions in a different order
    plus the fp16 gradient stream: 2e-3, the tolerance of test_engine_gpu.py), and scales linearly with the upstream gradient."""
    cfg, sd, model, x, y = p8s8
    crit = torch.nn.CrossEntropyLoss(reduction="sum")

    def grads(lo, hi, scale=1.0):
        model.zero_grad(set_to_none=True)
        logits, _ = model(x[lo:hi].contiguous(), y[lo:hi].contiguous())
        (crit(logits, y[lo:hi]) * (scale / 1024.0)).backward()
        return torch.cat([p.grad.flatten() for p in model.lora_parameters()]).clone()

    g_all = grads(0, 1024)
    g_a, g_b = grads(0, 512), grads(512, 1024)
    assert torch.isfinite(g_all).all() and float(g_all.norm()) > 0
    assert rel(g_a + g_b, g_all) < 2e-3
    assert rel(grads(0, 512, scale=4.0), 4.0 * g_a) < 1e-3          # power-of-two scale: only the fp16 gradient stream's rounding differs


def test_fused_step_split_matches_autograd_loop_at_full_size(p8s8):
    """engine_cl.unlearn_step on 512 remain + 512 forget == the reference's own loop shape (two forwards, torch CE, relu(BND - CE_f), backward)
    driven through the autograd seam of the same engine, at BASELINE's size: scalars to 1e-5, LoRA gradient to 2e-3."""
    import engine_cl
    from test_engine_gpu import build_model
    cfg, sd, model, x, y = p8s8
    xr, yr, xf, yf = x[:512].contiguous(), y[:512].contiguous(), x[512:].contiguous(), y[512:].contiguous()
    beta, BND = 0.15, 105.0
    crit = torch.nn.CrossEntropyLoss()
    model.zero_grad(set_to_none=True)
    out_r, _ = model(xr, yr)
    out_f, _ = model(xf, yf)
    ce_r, ce_f = crit(out_r, yr), crit(out_f, yf)
    (torch.relu(BND - ce_f) * beta + ce_r).backward()
    g_loop = torch.cat([p.grad.flatten() for p in model.lora_parameters()]).clone()
    fused = build_model(cfg, sd)
    out = engine_cl.unlearn_step(fused, xr, yr, xf, yf, beta=beta, alpha=0.0, BND=BND, hparams=dict(lr=0.0, wd=0.0))
    assert abs(out["loss_remain"] - float(ce_r)) < 1e-5 * abs(float(ce_r))
    assert abs(out["ce_forget"] - float(ce_f)) < 1e-5 * abs(float(ce_f))
    assert abs(out["loss_forget"] - max(BND - float(ce_f), 0.0)) < 1e-3
    # lr = 0: parameters must not move; the gradient buffer holds the step's LoRA gradient
    for p, q in zip(fused.lora_parameters(), model.lora_parameters()):
        assert torch.equal(p.data, q.data)
    g_step = fused._engine.grad_flat.clone()
    assert rel(g_step, g_loop) < 2e-3
