"""Pin the oracle (oracle/vit_oracle.py) against golden vectors produced by the UNMODIFIED
reference (tests/golden/make_golden.py).  CPU only."""
import os

import pytest
import torch

from oracle import vit_oracle as O

CASES = ["tiny6_b4", "tiny6_b4_proto", "tiny6_b3_lowbnd", "p8s8_b2"]


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def load_case(golden_dir, name):
    g = torch.load(os.path.join(golden_dir, name + ".pt"), weights_only=False)
    cfg = O.VitConfig(**g["cfg"])
    sd = g.get("state_dict") or O.init_state_dict(cfg, seed=g["seed"])
    for k, v in g["state_dict_checksum"].items():
        assert abs(float(sd[k].double().abs().sum()) - v) <= 1e-9 * max(1.0, abs(v)), f"weight regen drift: {k}"
    return g, cfg, {k: v.clone() for k, v in sd.items()}


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(golden_dir, name):
    g, cfg, sd = load_case(golden_dir, name)
    hp = g["hp"]
    kw = {}
    if hp.get("use_proto"):
        kw = dict(prototypes=g["prototypes"], w_pf=hp["w_pf"], w_pr=hp["w_pr"], BND_pro=hp["BND_pro"])
    state = {}
    for rec in g["steps"]:
        out, grads = O.unlearn_step(sd, cfg, state, g["img_r"], g["lab_r"], g["img_f"], g["lab_f"],
                                    lr=hp["lr"], wd=hp["wd"], beta=hp["beta"], alpha=hp["alpha"], BND=hp["BND"], **kw)
        assert rel(out["logits_r"], rec["logits_r"]) < 2e-6
        assert rel(out["logits_f"], rec["logits_f"]) < 2e-6
        assert rel(out["emb_r"], rec["emb_r"]) < 2e-6
        for key in ("loss_remain", "ce_forget", "loss_forget", "structure", "total", "proto_forget", "proto_remain"):
            assert abs(float(out[key]) - rec[key]) <= 2e-5 * max(1.0, abs(rec[key])), key
        sub = (lambda t: t.flatten()[::37]) if name == "p8s8_b2" else (lambda t: t)
        for n in O.lora_param_list(cfg):
            ref = rec["grads"][n]
            assert (sub(grads[n]).double() - ref.double()).norm() <= 2e-5 * ref.double().norm() + 1e-9, n
            assert rel(sub(sd[n]), rec["params_after"][n]) < 1e-5, n
    l2 = O.norm_of_lora(sd, cfg, "L2")
    l1 = O.norm_of_lora(sd, cfg, "L1")
    for a, b in zip(l2, g["norm_of_lora_L2"]):
        assert abs(float(a) - b) < 1e-4 * abs(b)
    for a, b in zip(l1, g["norm_of_lora_L1"]):
        assert abs(float(a) - b) < 1e-4 * abs(b)


def test_lowbnd_gate_closed(golden_dir):
    g, cfg, sd = load_case(golden_dir, "tiny6_b3_lowbnd")
    assert g["steps"][0]["loss_forget"] == 0.0      # relu(BND - CE) closed => no forget gradient


def test_merge_semantics_eval_forward(golden_dir):
    """loralib eval(): W += B@A/r ; merged forward == unmerged forward."""
    g, cfg, sd = load_case(golden_dir, "tiny6_b4")
    # replay the two optimizer steps so sd matches the golden model's final state
    hp = g["hp"]
    state = {}
    for _ in g["steps"]:
        O.unlearn_step(sd, cfg, state, g["img_r"], g["lab_r"], g["img_f"], g["lab_f"], lr=hp["lr"], wd=hp["wd"],
                       beta=hp["beta"], alpha=hp["alpha"], BND=hp["BND"])
    with torch.no_grad():
        logits, _ = O.vit_forward(sd, cfg, g["img_r"], g["lab_r"])
    assert rel(logits, g["eval_logits_r"]) < 1e-5
    n0 = O.blk(0, "1.fn.fn.net.0.")
    merged = sd[n0 + "weight"] + (sd[n0 + "lora_B"] @ sd[n0 + "lora_A"]) * cfg.lora_scaling
    assert rel(merged[:4, :8], g["eval_merged_fc1_w0"]) < 1e-6


def test_flop_table_matches_baseline_md():
    f = O.flops_per_image(O.P8S8)
    assert abs(f["fwd"] / 1e9 - 8.049) < 0.01
    assert abs(f["bwd"] / 1e9 - 7.599) < 0.01


def test_class_prototypes_match_the_unmodified_reference_function(golden_dir):
    """util.utils.calculate_prototypes of the UNMODIFIED reference (tests/golden/make_golden_prototypes.py) vs the oracle's restatement."""
    g = torch.load(os.path.join(golden_dir, "tiny6_prototypes.pt"), weights_only=False)
    cfg = O.VitConfig(**g["cfg"])
    sd = O.init_state_dict(cfg, seed=g["seed"])
    for k, v in g["state_dict_checksum"].items():
        assert abs(float(sd[k].double().abs().sum()) - v) <= 1e-9 * max(1.0, abs(v)), f"weight regen drift: {k}"
    got = O.class_prototypes(sd, cfg, g["images"], g["labels"], batch_size=g["batch_size"])
    assert sorted(got) == sorted(g["prototypes"]) == sorted(set(g["labels"].tolist()))      # only the classes that occur
    for k, want in g["prototypes"].items():
        assert got[k].shape == want.shape == (cfg.dim,)
        assert rel(got[k], want) < 2e-6, k
    # the per-class mean does not depend on how the dataset was batched (up to fp32 matmul blocking)
    other = O.class_prototypes(sd, cfg, g["images"], g["labels"], batch_size=5)
    for k in got:
        assert rel(other[k], got[k]) < 2e-6


def test_pixels_to_tensor_is_torchvisions_totensor_normalize():
    """oracle.pixels_to_tensor vs torchvision's own ToTensor / Normalize on PIL images (the reference's loader transforms,
    train/train_own_forget_cl.py:130-146): bit-equal."""
    import numpy as np
    from PIL import Image
    import torchvision.transforms as T
    rng = np.random.default_rng(0)
    arrs = [rng.integers(0, 256, (24, 24, 3), dtype=np.uint8) for _ in range(3)]
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    u8 = torch.stack([torch.from_numpy(a).permute(2, 0, 1) for a in arrs])
    plain = torch.stack([T.ToTensor()(Image.fromarray(a)) for a in arrs])
    normed = torch.stack([T.Compose([T.ToTensor(), T.Normalize(mean, std)])(Image.fromarray(a)) for a in arrs])
    assert torch.equal(O.pixels_to_tensor(u8), plain)
    assert torch.equal(O.pixels_to_tensor(u8, mean, std), normed)


@pytest.mark.parametrize("depth", [6, 3])
def test_groupings_match_the_unmodified_reference_engine_py_and_cal_norm(golden_dir, depth):
    """engine.get_structure_loss(group_type) and util.cal_norm.get_norm_of_lora(group_type) of the UNMODIFIED reference
    (tests/golden/make_golden_groups.py) vs the oracle's group-lasso value and vs the group ORDER the drop-in util.cal_norm reports in."""
    import sys
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gs-lora_b200")
    if pkg not in sys.path:
        sys.path.insert(0, pkg)
    from util import cal_norm
    rec = torch.load(os.path.join(golden_dir, "tiny_groupings.pt"), weights_only=False)[f"depth{depth}"]
    cfg = O.VitConfig(**rec["cfg"])
    sd = O.init_state_dict(cfg, seed=rec["seed"])
    blocks = O.lora_names(cfg)                       # per block [fc1.A, fc1.B, fc2.A, fc2.B] = the engine's flat tensor order
    for gt in ("block", "lora", "matrix"):
        assert abs(float(O.structure_loss(sd, cfg, gt)) - rec["structure"][gt]) < 1e-5 * rec["structure"][gt], gt
        groups = cal_norm._ffn_groups(depth, gt)     # [(block, which)] per reported group
        for typ, fn in (("L2", lambda t: t.norm(p=2)), ("L1", lambda t: t.abs().sum())):
            got = [float(sum(fn(sd[blocks[i][w]]) for i, w in grp)) for grp in groups]
            want = rec["norms"][f"{gt}_{typ}"]
            assert len(got) == len(want)
            for a, b in zip(got, want):
                assert abs(a - b) < 1e-5 * abs(b), (gt, typ)


def test_torchvision_family_reference_is_plain_torchvision_plus_loralib(golden_dir):
    """The GPU parity tests of configs 4 / 5 (tests/test_engine_gpu.py::_tv_pair) use torchvision's VisionTransformer + the loralib restatement,
    FP32 eager, as their reference.  This pins that construction against the UNMODIFIED reference wrapper (ModifiedViT + replace_ffn_with_lora,
    tests/golden/make_golden_tv.py): same state_dict keys, logits, cls embedding (= encoder output row 0, after encoder.ln) and LoRA gradients."""
    from torchvision.models.vision_transformer import VisionTransformer
    from oracle import loralib_restated as olora
    g = torch.load(os.path.join(golden_dir, "tv_small_b3.pt"), weights_only=False)
    sh = g["shape"]
    ref = VisionTransformer(**sh)
    for blk in ref.encoder.layers.children():
        blk.mlp[0] = olora.Linear(sh["hidden_dim"], sh["mlp_dim"], r=g["rank"])
        blk.mlp[3] = olora.Linear(sh["mlp_dim"], sh["hidden_dim"], r=g["rank"])
    assert set(ref.state_dict().keys()) == set(g["state_dict"].keys())
    ref.load_state_dict(g["state_dict"], strict=True)
    olora.mark_only_lora_as_trainable(ref)
    ref.train()
    names = [n for n, p in ref.named_parameters() if p.requires_grad]
    assert names == g["trainable"] and len(names) == 4 * sh["num_layers"]
    # torchvision's forward up to the head, keeping the cls row the reference returns as the embedding (modified_VIT.py:26-37)
    x = ref._process_input(g["x"])
    x = torch.cat([ref.class_token.expand(x.shape[0], -1, -1), x], dim=1)
    emb = ref.encoder(x)[:, 0]
    logits = ref.heads(emb)
    assert torch.equal(logits, ref(g["x"]))                       # == torchvision's own forward()
    loss = torch.nn.functional.cross_entropy(logits, g["y"])
    loss.backward()
    assert rel(logits, g["logits"]) < 1e-6 and rel(emb, g["emb"]) < 1e-6 and abs(float(loss) - g["loss"]) < 1e-6
    for n in names:
        assert rel(ref.get_parameter(n).grad, g["grads"][n]) < 1e-5, n


def test_attention_lora_oracle_matches_reference_golden(golden_dir):
    """lora_pos "Attention" (SURVEY 8f-2): the oracle against the UNMODIFIED ViT_face(lora_pos="Attention") + engine.get_structure_loss(group_pos=
    "Attention") + cal_norm(group_pos="Attention") records of tests/golden/make_golden_attn.py -- logits, losses, every to_qkv LoRA gradient,
    parameters after two AdamW steps, the norm report and the merged eval forward."""
    g = torch.load(os.path.join(golden_dir, "tiny3_attn_lora.pt"), weights_only=False)
    cfg, hp = O.VitConfig(**g["cfg"]), g["hp"]
    assert cfg.lora_pos == "Attention"
    sd = {k: v.clone() for k, v in g["state_dict"].items()}
    regen = O.init_state_dict(cfg, seed=g["seed"])
    assert set(regen) == set(sd) and all(torch.equal(regen[k], sd[k]) for k in sd)
    assert not any("net.0.lora" in k or "net.3.lora" in k for k in sd) and sum("to_qkv.lora_" in k for k in sd) == 2 * cfg.depth
    state = {}
    for rec in g["steps"]:
        out, grads = O.unlearn_step(sd, cfg, state, g["img_r"], g["lab_r"], g["img_f"], g["lab_f"], lr=hp["lr"], wd=hp["wd"], beta=hp["beta"],
                                    alpha=hp["alpha"], BND=hp["BND"])
        assert rel(out["logits_r"], rec["logits_r"]) < 2e-6 and rel(out["logits_f"], rec["logits_f"]) < 2e-6
        for key in ("loss_remain", "ce_forget", "loss_forget", "structure", "total"):
            assert abs(float(out[key]) - rec[key]) <= 2e-5 * max(1.0, abs(rec[key])), key
        for n in O.lora_param_list(cfg):
            assert rel(grads[n], rec["grads"][n]) < 2e-5, n
            assert rel(sd[n], rec["params_after"][n]) < 1e-5, n
    for a, b in zip(O.norm_of_lora(sd, cfg, "L2"), g["norm_of_lora_L2"]):
        assert abs(float(a) - b) < 1e-4 * abs(b)
    for a, b in zip(O.norm_of_lora(sd, cfg, "L1"), g["norm_of_lora_L1"]):
        assert abs(float(a) - b) < 1e-4 * abs(b)
    with torch.no_grad():       # merged == un-merged function; the golden's merged to_qkv rows are W + s B_g A_g slice by slice
        logits_e, _ = O.vit_forward(sd, cfg, g["img_r"], g["lab_r"])
    assert rel(logits_e, g["eval_logits_r"]) < 1e-5
    W = sd[O.blk(0, "0.fn.fn.to_qkv.weight")] + cfg.lora_scaling * O.merged_qkv_delta(sd[O.blk(0, "0.fn.fn.to_qkv.lora_A")],
                                                                                     sd[O.blk(0, "0.fn.fn.to_qkv.lora_B")], cfg.lora_rank)
    assert rel(W[::37, :16], g["eval_merged_qkv_w"]) < 1e-6
