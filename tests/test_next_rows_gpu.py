"""SURVEY 8f rows on the GPU: the uint8 input pipeline (8f-4) and device-side class prototypes (8f-1), each against a plain PyTorch
restatement of the reference's host code (transforms.ToTensor / Normalize, util/utils.py:502-549).  Integer / selection work and the
fp32 sums taken in the reference's order are held to bit equality; the rest to the tolerances of test_engine_gpu.py."""
import ctypes
import os

import pytest
import torch

from oracle import vit_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F():
    from gslora import _ffi
    _ffi.lib()
    return _ffi


def rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).norm() / (b.double().cpu().norm() + 1e-30))


def _model(cfg, seed=5, device="cuda"):
    from test_engine_gpu import build_model
    sd = O.init_state_dict(cfg, seed=seed)
    return build_model(cfg, sd, device), sd


IMAGENET_NORM = ([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])       # train/train_own_forget_cl.py:138-139


@pytest.mark.parametrize("B,C,S,patch,order", [(3, 3, 112, 8, 0), (2, 3, 64, 16, 1), (1, 3, 40, 8, 0), (5, 1, 32, 8, 1)])
@pytest.mark.parametrize("norm", [None, IMAGENET_NORM])
@pytest.mark.parametrize("nhwc", [False, True])
def test_patchify_u8_equals_host_totensor_normalize_bitwise(F, B, C, S, patch, order, norm, nhwc):
    """uint8 pixels -> ToTensor (/255) -> Normalize -> patchify: the kernel's fp16 rows must equal, bit for bit, the fp32 patchify of the
    host-transformed image (what the reference's loader hands to the model)."""
    g = torch.Generator().manual_seed(B * 131 + S)
    u8 = torch.randint(0, 256, (B, C, S, S), dtype=torch.uint8, generator=g)
    mean, std = (norm[0][:C], norm[1][:C]) if norm is not None else (None, None)
    ref = O.pixels_to_tensor(u8, mean, std)        # transforms.ToTensor [+ Normalize]; pinned bit-equal to torchvision in test_oracle_golden.py
    P = (S // patch) ** 2
    pd = C * patch * patch
    out_ref = torch.full((B * (P + 1), pd), 7.0, dtype=torch.half, device="cuda")
    out_u8 = torch.full_like(out_ref, 9.0)
    L = F.lib()
    F.check(L.gsl_patchify_f16(F.ptr(ref.cuda().contiguous()), F.ptr(out_ref), pd, B, C, S, patch, order, F.cur_stream()))
    src = (u8.permute(0, 2, 3, 1) if nhwc else u8).contiguous().cuda()
    F.check(L.gsl_patchify_u8_f16(F.ptr(src), 1 if nhwc else 0, F.host_floats(mean), F.host_floats(std), F.ptr(out_u8), pd, B, C, S, patch, order,
                                  F.cur_stream()))
    torch.cuda.synchronize()
    assert torch.equal(out_u8.view(torch.int16), out_ref.view(torch.int16))
    assert float(out_u8.view(B, P + 1, pd)[:, 0].abs().max()) == 0.0      # cls slot rows stay zero


def test_patchify_u8_rejects_bad_arguments(F):
    L = F.lib()
    img = torch.zeros(1, 3, 16, 16, dtype=torch.uint8, device="cuda")
    out = torch.zeros(5, 192, dtype=torch.half, device="cuda")
    assert L.gsl_patchify_u8_f16(F.ptr(img), 2, None, None, F.ptr(out), 192, 1, 3, 16, 8, 0, F.cur_stream()) != 0          # layout
    assert L.gsl_patchify_u8_f16(F.ptr(img), 0, F.host_floats([0.5] * 3), None, F.ptr(out), 192, 1, 3, 16, 8, 0, F.cur_stream()) != 0   # mean w/o std
    assert L.gsl_patchify_u8_f16(F.ptr(img), 0, F.host_floats([0.5] * 3), F.host_floats([1, 0, 1]), F.ptr(out), 192, 1, 3, 16, 8, 0,
                                 F.cur_stream()) != 0                                                                        # std == 0
    assert b"std" in L.gsl_last_error()


def test_uint8_forward_and_unlearn_step_equal_the_fp32_path():
    """Same pixels as uint8 (NCHW and NHWC) or as ToTensor output: identical logits / embeddings, and an identical fused unlearning step."""
    import engine_cl
    cfg = O.TINY
    g = torch.Generator().manual_seed(11)
    S = cfg.image_size
    u_r = torch.randint(0, 256, (5, 3, S, S), dtype=torch.uint8, generator=g).cuda()
    u_f = torch.randint(0, 256, (3, 3, S, S), dtype=torch.uint8, generator=g).cuda()
    y_r = torch.randint(0, cfg.num_class, (5,), generator=g).cuda()
    y_f = torch.randint(0, cfg.num_class, (3,), generator=g).cuda()
    m32, _ = _model(cfg)
    m8, _ = _model(cfg)
    with torch.no_grad():
        l32, e32 = m32(u_r.float().div(255), y_r)
        l8, e8 = m8(u_r, y_r)
        l8c, e8c = m8(u_r.permute(0, 2, 3, 1).contiguous(), y_r)
    assert torch.equal(l32, l8) and torch.equal(e32, e8) and torch.equal(l32, l8c) and torch.equal(e32, e8c)
    kw = dict(beta=0.15, alpha=1e-3, BND=105.0, hparams=dict(lr=1e-2, wd=0.05))
    o32 = engine_cl.unlearn_step(m32, u_r.float().div(255), y_r, u_f.float().div(255), y_f, **kw)
    o8 = engine_cl.unlearn_step(m8, u_r, y_r, u_f, y_f, **kw)
    assert o32 == o8
    for p, q in zip(m32.lora_parameters(), m8.lora_parameters()):
        assert torch.equal(p, q)
    with pytest.raises(TypeError):
        engine_cl.unlearn_step(m8, u_r, y_r, u_f.float(), y_f, **kw)
    # Normalize in flight (ImageNet runs): equals the host-normalised fp32 input
    m8.input_pixel_norm = IMAGENET_NORM
    mean, std = (torch.tensor(v, device="cuda").view(1, 3, 1, 1) for v in IMAGENET_NORM)
    with torch.no_grad():
        ln, _ = m8(u_r, y_r)
        lr_, _ = m32(u_r.float().div(255).sub(mean).div(std), y_r)
    assert torch.equal(ln, lr_)


@pytest.mark.parametrize("B,D,C", [(37, 512, 100), (300, 128, 10), (1, 768, 100), (64, 1024, 7)])
def test_class_sums_follow_the_reference_loop_bitwise(F, B, D, C):
    """util/utils.py:535-547 restated: sums in dataset order, fp32, then / count.  Two batches to cover the carried accumulators."""
    g = torch.Generator().manual_seed(B + D)
    emb = torch.randn(B, D, generator=g).cuda()
    lab = torch.randint(0, C, (B,), generator=g).cuda()
    sums = torch.zeros(C, D, device="cuda")
    counts = torch.zeros(C, device="cuda")
    L = F.lib()
    cut = B // 3
    for lo, hi in ((0, cut), (cut, B)):
        F.check(L.gsl_class_sums(F.ptr(emb[lo:hi].contiguous()), F.ptr(lab[lo:hi].contiguous()), hi - lo, D, C, F.ptr(sums), F.ptr(counts),
                                 F.cur_stream()))
    means = torch.empty_like(sums)
    F.check(L.gsl_class_means(F.ptr(sums), F.ptr(counts), C, D, F.ptr(means), F.cur_stream()))
    ref_sum, ref_n = {}, {}
    emb_h, lab_h = emb, lab.cpu()                      # sums and the final division run on the DEVICE in the reference (a * (1.0f / n) in ATen)
    for e, l in zip(emb_h, lab_h):                     # the reference's per-sample loop
        k = int(l)
        ref_sum[k] = ref_sum.get(k, 0) + e
        ref_n[k] = ref_n.get(k, 0) + 1
    means_h, counts_h = means.cpu(), counts.cpu()
    for k in range(C):
        if k in ref_sum:
            assert counts_h[k] == ref_n[k]
            assert torch.equal(means_h[k], (ref_sum[k] / ref_n[k]).cpu())
        else:
            assert counts_h[k] == 0 and float(means_h[k].abs().max()) == 0.0


def test_calculate_prototypes_matches_reference_function_semantics():
    """util.utils.calculate_prototypes on an engine-backed model == the reference's loop over the same model's eval-mode embeddings; feeds
    engine_cl.get_prototype_loss / the fused GS-LoRA++ step unchanged."""
    from util.utils import calculate_prototypes
    import engine_cl
    cfg = O.TINY
    model, _ = _model(cfg, seed=9)
    g = torch.Generator().manual_seed(21)
    N = 23
    imgs = torch.rand(N, 3, cfg.image_size, cfg.image_size, generator=g)
    labs = torch.randint(0, 4, (N,), generator=g)                       # 4 of the classes only: absent classes must not appear
    ds = torch.utils.data.TensorDataset(imgs, labs)
    protos = calculate_prototypes(model, ds, batch_size=8, device="cuda")
    assert not model.training                                            # the reference leaves the backbone in eval mode
    ref_sum, ref_n = {}, {}
    with torch.no_grad():
        for lo in range(0, N, 8):
            _, emb = model(imgs[lo:lo + 8].cuda(), labs[lo:lo + 8].cuda())
            for e, l in zip(emb, labs[lo:lo + 8]):
                ref_sum[int(l)] = ref_sum.get(int(l), 0) + e
                ref_n[int(l)] = ref_n.get(int(l), 0) + 1
    assert sorted(protos) == sorted(ref_sum)
    for k in ref_sum:
        assert protos[k].device.type == "cpu" and protos[k].shape == (cfg.dim,)
        assert torch.equal(protos[k], (ref_sum[k] / ref_n[k]).cpu())
    model.train()
    out = engine_cl.unlearn_step(model, imgs[:6].cuda(), labs[:6].cuda(), imgs[6:10].cuda(), labs[6:10].cuda(), beta=0.15, alpha=1e-3, BND=105.0,
                                 hparams=dict(lr=1e-2, wd=0.05), use_prototype=True, prototype_dict=protos, prototype_weight_forget=0.5,
                                 prototype_weight_remain=0.5, BND_pro=1.0)
    assert out["proto_remain"] >= 0.0 and out["total"] == out["total"]


def test_reinitialize_lora_parameters_is_seen_by_the_engine():
    """util/utils.py:428-441 mutates lora_A / lora_B in place between tasks (train_own_forget_cl.py:536): the engine must pick the new values up
    (lora_B = 0 => the LoRA branch vanishes: logits equal the frozen model's)."""
    from util.utils import reinitialize_lora_parameters
    cfg = O.TINY
    model, sd = _model(cfg, seed=4)
    x = torch.rand(4, 3, cfg.image_size, cfg.image_size).cuda()
    y = torch.randint(0, cfg.num_class, (4,)).cuda()
    with torch.no_grad():
        before, _ = model(x, y)
        reinitialize_lora_parameters(model)
        after, _ = model(x, y)
    assert all(float(p.detach().abs().max()) == 0.0 for n, p in model.named_parameters() if "lora_B" in n)
    sd0 = {k: (torch.zeros_like(v) if "lora_B" in k else v) for k, v in sd.items()}
    from test_engine_gpu import build_model
    frozen_only = build_model(cfg, sd0)
    with torch.no_grad():
        base, _ = frozen_only(x, y)
    assert torch.equal(after, base) and not torch.equal(before, after)


# ------------------------------------------------------------------------------------------------ engine.py twin
def _loaders(cfg, n_forget, n_remain, seed=17):
    gen = torch.Generator().manual_seed(seed)
    S = cfg.image_size
    forget = [(torch.rand(3, 3, S, S, generator=gen), torch.randint(0, cfg.num_class, (3,), generator=gen)) for _ in range(n_forget)]
    remain = [(torch.rand(4, 3, S, S, generator=gen), torch.randint(0, cfg.num_class, (4,), generator=gen)) for _ in range(n_remain)]
    return forget, remain


@pytest.mark.parametrize("few_shot,alpha_epoch,group_type", [(True, 0, "block"), (False, 0, "matrix"), (True, 5, "lora")])
def test_engine_py_train_one_epoch_loader_swap_and_alpha_gate(few_shot, alpha_epoch, group_type):
    """engine.train_one_epoch (engine.py:13-434): with the longer forget loader and cfg["few_shot"] the FORGET loader drives the epoch and the
    remain loader is recycled (engine.py:53-57); epoch < ALPHA_EPOCH switches the structure term off (engine.py:82-90); GROUP_TYPE picks the
    group-lasso grouping.  Result must equal the same sequence of synchronous fused steps."""
    import engine
    import engine_cl
    from engine_cl import AverageMeter
    cfg = O.TINY
    forget, remain = _loaders(cfg, 6, 4)
    hp = dict(lr=1e-2, wd=0.05, beta=0.15, alpha=1e-2, BND=105.0)
    run_cfg = {"few_shot": few_shot, "ALPHA_EPOCH": alpha_epoch, "GROUP_TYPE": group_type, "GROUP_POS": "FFN", "NUM_LAYERS": cfg.depth,
               "WORK_PATH": "/tmp", "BACKBONE_NAME": "VIT", "MULTI_GPU": False}

    def fresh():
        m, _ = _model(cfg, seed=6)
        opt = torch.optim.AdamW([p for p in m.parameters() if p.requires_grad], lr=hp["lr"], weight_decay=hp["wd"])
        return m, opt

    m1, opt1 = fresh()
    alpha_eff = 0.0 if 0 < alpha_epoch else hp["alpha"]
    if few_shot:    # forget (6 batches) drives, remain (4) recycled
        pairs = [(remain[i % len(remain)], forget[i]) for i in range(len(forget))]
    else:           # remain drives, forget recycled
        pairs = [(remain[i], forget[i % len(forget)]) for i in range(len(remain))]
    for (xr, yr), (xf, yf) in pairs:
        out = engine_cl.unlearn_step(m1, xr.cuda(), yr.cuda(), xf.cuda(), yf.cuda(), beta=hp["beta"], alpha=alpha_eff, BND=hp["BND"], optimizer=opt1,
                                     group_type=group_type)
    m2, opt2 = fresh()
    mt = [AverageMeter() for _ in range(8)]
    ret = engine.train_one_epoch(m2, forget, remain, torch.device("cuda"), torch.nn.CrossEntropyLoss(), opt2, 0, mt[0], mt[1], mt[2], mt[3], mt[4],
                                 mt[5], hp["beta"], hp["alpha"], hp["BND"], 0, None, None, 0.0, 0.0, run_cfg,
                                 losses_prototype_forget=mt[6], losses_prototype_remain=mt[7])
    assert ret[0] == len(pairs) and len(ret) == 10
    for p, q in zip(m1.lora_parameters(), m2.lora_parameters()):
        assert torch.equal(p.data, q.data)
    if alpha_eff == 0.0:
        assert ret[7].sum == 0.0          # structure meter: alpha * loss with the term gated off
    # the structure loss of the twin's signature equals engine_cl's for the same grouping
    a = engine.get_structure_loss(m2, num_layers=cfg.depth, group_type=group_type, group_pos="FFN")
    b = engine_cl.get_structure_loss(m2, group_type=group_type)
    assert float(a) == float(b)
    with pytest.raises(ValueError):
        engine.get_structure_loss(m2, num_layers=cfg.depth, group_type="block", group_pos="Attention")
    with pytest.raises(ValueError):
        engine.get_structure_loss(m2, num_layers=cfg.depth + 1)


def test_engine_py_eval_and_checkpoint_match_a_merged_deepcopy(tmp_path):
    """engine.eval_data / evaluate (engine.py:436-529) evaluate `copy.deepcopy(model).eval()` in the reference; the twin evaluates the live weights
    and saves a merged state_dict without touching the training model."""
    import copy
    import engine
    cfg = O.TINY
    model, _ = _model(cfg, seed=8)
    gen = torch.Generator().manual_seed(2)
    S = cfg.image_size
    test = [(torch.rand(6, 3, S, S, generator=gen), torch.randint(0, cfg.num_class, (6,), generator=gen)) for _ in range(3)]
    w_name = "transformer.layers.0.1.fn.fn.net.0.weight"
    w0 = model.get_parameter(w_name).detach().clone()
    acc = engine.eval_data(model, test, torch.device("cuda"), "forget")
    assert model.training and torch.equal(model.get_parameter(w_name), w0)          # model untouched, still in train mode
    ref = copy.deepcopy(model).eval()
    hits = total = 0
    with torch.no_grad():
        for x, y in test:
            logits, _ = ref(x.cuda(), y.cuda())
            hits += int((logits.argmax(1).cpu() == y).sum())
            total += len(y)
    assert abs(acc - 100 * hits / total) < 1e-9
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-2)
    run_cfg = {"WORK_PATH": str(tmp_path), "BACKBONE_NAME": "VIT", "MULTI_GPU": False}
    h = engine.evaluate(model, test, test, torch.device("cuda"), 0, 0, forget_acc_before=100.0, highest_H_mean=-1.0, cfg=run_cfg, optimizer=opt)
    files = [f for f in os.listdir(tmp_path) if f.endswith(".pth")]
    assert len(files) == 1 and h > -1.0
    saved = torch.load(os.path.join(tmp_path, files[0]))
    ref_sd = ref.state_dict()
    assert set(saved) == set(ref_sd)
    for k in saved:
        assert torch.allclose(saved[k].cpu(), ref_sd[k].cpu(), rtol=0, atol=1e-6), k
    assert not torch.equal(saved[w_name].cpu(), w0.cpu())                            # the LoRA delta is in the saved weight


# ------------------------------------------------------------------------------------------------ 8f-2: LoRA on attention (lora_pos "Attention")
def _attn_model(cfg, sd):
    import loralib as lora
    from vit_pytorch_face import ViT_face
    m = ViT_face(loss_type="CosFace", GPU_ID=[0], num_class=cfg.num_class, image_size=cfg.image_size, patch_size=cfg.patch_size, dim=cfg.dim,
                 depth=cfg.depth, heads=cfg.heads, mlp_dim=cfg.mlp_dim, dim_head=cfg.dim_head, dropout=0.0, emb_dropout=0.0,
                 lora_rank=cfg.lora_rank, lora_pos="Attention")
    m.load_state_dict(sd, strict=True)
    lora.mark_only_lora_as_trainable(m)
    assert sorted(n for n, p in m.named_parameters() if p.requires_grad) == sorted(O.lora_param_list(cfg))
    return m.cuda().train()


def test_attention_lora_matches_unmodified_reference_golden(golden_dir):
    """ViT_face(lora_pos="Attention") on the engine -- merged to_qkv operand with per-slice (A_g, B_g), attention backward down to block 0,
    dA_g / dB_g side products -- against the unmodified reference's records (tests/golden/make_golden_attn.py): autograd path (two forwards,
    torch losses, engine.get_structure_loss(group_pos="Attention")), then the fused step, then the norm report and merge / un-merge."""
    import engine
    import engine_cl
    from util.cal_norm import get_norm_of_lora
    g = torch.load(os.path.join(golden_dir, "tiny3_attn_lora.pt"), weights_only=False)
    cfg, hp = O.VitConfig(**g["cfg"]), g["hp"]
    xr, yr, xf, yf = [g[k].cuda() for k in ("img_r", "lab_r", "img_f", "lab_f")]
    names = O.lora_param_list(cfg)
    rec = g["steps"][0]
    model = _attn_model(cfg, g["state_dict"])
    crit = torch.nn.CrossEntropyLoss()
    out_r, emb_r = model(xr, yr)
    out_f, _ = model(xf, yf)
    s_loss = engine.get_structure_loss(model, num_layers=cfg.depth, group_type="block", group_pos="Attention")
    total = torch.relu(hp["BND"] - crit(out_f, yf)) * hp["beta"] + crit(out_r, yr) + s_loss * hp["alpha"]
    total.backward()
    assert rel(out_r, rec["logits_r"]) < 1e-3 and rel(out_f, rec["logits_f"]) < 1e-3 and rel(emb_r, rec["emb_r"]) < 1e-3
    assert abs(float(s_loss) - rec["structure"]) < 1e-5 * rec["structure"] and abs(float(total) - rec["total"]) < 2e-3 * abs(rec["total"])
    per = {n: rel(model.get_parameter(n).grad, rec["grads"][n]) for n in names}
    allrel = rel(torch.cat([model.get_parameter(n).grad.flatten() for n in names]), torch.cat([rec["grads"][n].flatten() for n in names]))
    print(f"attention LoRA tiny3: logits {rel(out_r, rec['logits_r']):.2e} grads all {allrel:.2e} worst {max(per.values()):.2e}")
    assert allrel < 1e-3 and max(per.values()) < 1.5e-3
    with pytest.raises(ValueError):
        engine.get_structure_loss(model, num_layers=cfg.depth, group_pos="FFN")
    # fused steps, teacher-forced onto the reference's parameters between them
    model = _attn_model(cfg, g["state_dict"])
    for rec in g["steps"]:
        before = {n: model.get_parameter(n).detach().clone() for n in names}
        out = engine_cl.unlearn_step(model, xr, yr, xf, yf, beta=hp["beta"], alpha=hp["alpha"], BND=hp["BND"], hparams=dict(lr=hp["lr"], wd=hp["wd"]))
        for key in ("loss_remain", "ce_forget", "loss_forget", "structure", "total"):
            assert abs(out[key] - rec[key]) <= 2e-3 * max(1.0, abs(rec[key])), (key, out[key], rec[key])
        eng = model._engine
        got, ref = [], []
        for l, grp in enumerate(O.lora_names(cfg)):
            gn = torch.sqrt(sum((before[n] ** 2).sum() for n in grp))
            for w, n in enumerate(grp):
                got.append((eng.lora_view(eng.grad_flat, l, w) + hp["alpha"] * before[n] / gn).flatten())
                ref.append(rec["grads"][n].flatten())
        assert rel(torch.cat(got), torch.cat(ref)) < 1e-3
        for n in names:
            gref = rec["grads"][n].cuda()
            p, pref = model.get_parameter(n).data, rec["params_after"][n].cuda()
            well = gref.abs() > 0.05 * gref.abs().mean()
            assert (p - pref)[well].abs().max() < 0.05 * hp["lr"], n
            p.copy_(pref)
        model.sync_engine(force_lora=True)
    norms = get_norm_of_lora(model, type="L2", group_num=cfg.depth, group_pos="Attention")
    for a, b in zip(norms, g["norm_of_lora_L2"]):
        assert abs(float(a) - b) < 2e-3 * abs(b)
    with pytest.raises(KeyError):
        get_norm_of_lora(model, type="L2", group_num=cfg.depth, group_pos="FFN")
    # loralib merge semantics on to_qkv: eval() folds s B_g A_g into the weight slice by slice, the function does not change, train() restores
    with torch.no_grad():
        lt, _ = model(xr, yr)
        w0 = model.get_parameter(O.blk(0, "0.fn.fn.to_qkv.weight")).clone()
        model.eval()
        le, _ = model(xr, yr)
        assert not torch.equal(model.get_parameter(O.blk(0, "0.fn.fn.to_qkv.weight")), w0)
        model.train()
        lt2, _ = model(xr, yr)
    assert rel(le, lt) < 1e-3 and rel(lt2, lt) < 1e-4 and rel(le, g["eval_logits_r"].cuda()) < 5e-2


@pytest.mark.parametrize("mode", ["split8", "split", "fast"])
def test_attention_lora_p8s8_vs_oracle_fp32(mode):
    """Config-2 widths with LoRA r = 8 on to_qkv, bs 32+32: logits and the 12 to_qkv LoRA gradients against the oracle in FP32 on the same GPU."""
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = O.VitConfig(**{**O.P8S8.to_dict(), "lora_pos": "Attention"})
    sd = O.init_state_dict(cfg, seed=1337)
    gen = torch.Generator().manual_seed(7)
    B = 32
    xr, xf = torch.rand(B, 3, 112, 112, generator=gen).cuda(), torch.rand(B, 3, 112, 112, generator=gen).cuda()
    yr, yf = torch.randint(0, 100, (B,), generator=gen).cuda(), torch.randint(0, 100, (B,), generator=gen).cuda()
    ref, ref_grads = O.unlearn_grads({k: v.cuda() for k, v in sd.items()}, cfg, xr, yr, xf, yf, beta=0.15, alpha=1e-4, BND=105.0, include_structure=False)
    model = _attn_model(cfg, sd)
    model.gsl_precision = mode
    crit = torch.nn.CrossEntropyLoss()
    out_r, _ = model(xr, yr)
    out_f, _ = model(xf, yf)
    (torch.relu(105.0 - crit(out_f, yf)) * 0.15 + crit(out_r, yr)).backward()
    names = O.lora_param_list(cfg)
    per = {n: rel(model.get_parameter(n).grad, ref_grads[n]) for n in names}
    allrel = rel(torch.cat([model.get_parameter(n).grad.flatten() for n in names]), torch.cat([ref_grads[n].flatten() for n in names]))
    print(f"attention LoRA P8S8 bs32 [{mode}]: logits {rel(out_r, ref['logits_r']):.2e} grads all {allrel:.2e} worst {max(per.values()):.2e}")
    assert rel(out_r, ref["logits_r"]) < 1e-3 and rel(out_f, ref["logits_f"]) < 1e-3
    assert allrel < (1e-3 if mode != "fast" else 3e-3) and max(per.values()) < (1.5e-3 if mode != "fast" else 5e-3)
