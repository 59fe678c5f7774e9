"""Host-side logic of the drop-in loops on CPU (no GPU, no kernels): the engine step is replaced by a recording fake so that what is
checked is exactly what this repo adds around it -- loader pairing / recycling (engine_cl.py:51-58,226-231; engine.py:53-57), the meters'
arithmetic (engine_cl.py:68-121), the ALPHA_EPOCH gate (engine.py:82-90), display resets, the return tuple, and the scalar algebra of
StepResult (what the reference reads with `.item()`)."""
import types

import pytest
import torch

import engine
import engine_cl
from engine_cl import AverageMeter, StepResult


class _FakeEvent:
    def synchronize(self):
        pass


def _result(ce_r, n_r, ce_f, n_f, hit_r, hit_f, kl_r, kl_f, structure, **consts):
    c = dict(beta=0.15, alpha=1e-2, BND=105.0, BND_pro=0.0, pwf=0.0, pwr=0.0, use_prototype=False)
    c.update(consts)
    host = torch.tensor([ce_r * n_r, n_r, ce_f * n_f, n_f, hit_r, hit_f, kl_r * n_r, kl_f * n_f, structure], dtype=torch.float32)
    return StepResult(host, _FakeEvent(), 9, c)


def test_step_result_reproduces_the_reference_scalars():
    r = _result(2.0, 4, 100.0, 3, 3, 1, 0.0, 0.0, 7.5)
    assert r["loss_remain"] == pytest.approx(2.0) and r["ce_forget"] == pytest.approx(100.0)
    assert r["loss_forget"] == pytest.approx(5.0)                         # relu(BND - CE_f), engine_cl.py:78
    assert r["top1_remain"] == pytest.approx(75.0) and r["top1_forget"] == pytest.approx(100.0 / 3)
    assert r["total"] == pytest.approx(0.15 * 5.0 + 2.0 + 1e-2 * 7.5)    # engine_cl.py:118-121
    closed = _result(2.0, 4, 110.0, 3, 0, 0, 0.0, 0.0, 1.0)               # CE_f above the bound: the forget term is switched off
    assert closed["loss_forget"] == 0.0 and closed["total"] == pytest.approx(2.0 + 1e-2)
    proto = _result(1.0, 2, 50.0, 2, 0, 0, 0.25, 3.0, 0.0, use_prototype=True, BND_pro=18.0, pwf=0.5, pwr=2.0)
    assert proto["proto_forget"] == pytest.approx(3.0) and proto["proto_remain"] == pytest.approx(0.25)
    assert proto["total"] == pytest.approx(0.15 * 55.0 + 1.0 + 0.5 * (18.0 - 3.0) + 2.0 * 0.25)      # engine_cl.py:97-101
    empty = _result(0.0, 0, 0.0, 0, 0, 0, 0.0, 0.0, 0.0)                  # a rank with an empty shard divides by max(n, 1)
    assert empty["loss_remain"] == 0.0 and empty["top1_forget"] == 0.0


class _Recorder:
    def __init__(self):
        self.calls = []

    def __call__(self, model, xr, yr, xf, yf, **kw):
        self.calls.append((int(xr[0, 0]), int(xf[0, 0]), xr.shape[0], xf.shape[0], kw))
        return _result(1.0, xr.shape[0], 100.0, xf.shape[0], 1, 0, 0.0, 0.0, 4.0, beta=kw["beta"], alpha=kw["alpha"], BND=kw["BND"])


class _CpuPrefetcher:
    def __init__(self, loader, device):
        self.it = iter(loader)

    def next(self):
        return next(self.it, (None, None))


@pytest.fixture
def fake(monkeypatch):
    rec = _Recorder()
    for mod in (engine_cl, engine):
        monkeypatch.setattr(mod, "unlearn_step_async", rec)
    monkeypatch.setattr(engine_cl, "_Prefetcher", _CpuPrefetcher)
    monkeypatch.setattr(engine_cl, "engine_fresh_optimizer", lambda m, o: False)
    monkeypatch.setattr(engine_cl, "sync_optimizer_state", lambda m, o: None)
    return rec


def _loader(tag, n, bs):
    """batches whose first element identifies them: value = tag + index"""
    return [(torch.full((bs, 1), float(tag + i)), torch.zeros(bs, dtype=torch.long)) for i in range(n)]


def _meters(n=8):
    return [AverageMeter() for _ in range(n)]


def test_engine_cl_epoch_pairs_remain_batches_with_a_recycled_forget_loader(fake, capsys):
    remain, forget = _loader(100, 7, 4), _loader(200, 3, 2)
    lf, lr, lt, ls, tf, tr, lpf, lpr = _meters()
    ret = engine_cl.train_one_epoch(torch.nn.Linear(1, 1), forget, remain, "cpu", torch.nn.CrossEntropyLoss(), None, 0, lf, lr, lt, ls, tf, tr,
                                    0.15, 1e-2, 105.0, 0, None, None, 0.0, 0.0, {}, 3, False, None, 0.0, 0.0, lpf, lpr)
    assert [(c[0], c[1]) for c in fake.calls] == [(100, 200), (101, 201), (102, 202), (103, 200), (104, 201), (105, 202), (106, 200)]
    assert ret[0] == 7 and len(ret) == 10 and ret[1] == 0.0
    # the display at batch 5 (batch index 4) printed the running averages and reset the meters: the returned ones hold steps 5 and 6
    out = capsys.readouterr().out
    assert "Task 3 Epoch 1 Batch 5" in out
    losses_forget, losses_remain, losses_total, losses_structure = ret[2], ret[3], ret[6], ret[7]
    assert losses_remain.count == 8 and losses_forget.count == 4         # 2 steps x 4 remain / 2 forget images
    assert losses_forget.avg == pytest.approx(0.15 * 5.0) and losses_structure.avg == pytest.approx(1e-2 * 4.0)
    assert losses_total.avg == pytest.approx(0.15 * 5.0 + 1.0 + 1e-2 * 4.0)


@pytest.mark.parametrize("few_shot", [True, False])
def test_engine_py_epoch_swaps_the_driving_loader_only_for_few_shot(fake, few_shot):
    remain, forget = _loader(100, 2, 4), _loader(200, 5, 3)               # forget loader is the longer one
    m = _meters()
    cfg = {"few_shot": few_shot, "ALPHA_EPOCH": 0, "GROUP_TYPE": "lora", "GROUP_POS": "FFN"}
    ret = engine.train_one_epoch(torch.nn.Linear(1, 1), forget, remain, "cpu", torch.nn.CrossEntropyLoss(), None, 0, m[0], m[1], m[2], m[3], m[4],
                                 m[5], 0.15, 1e-2, 105.0, 0, None, None, 0.0, 0.0, cfg, losses_prototype_forget=m[6], losses_prototype_remain=m[7])
    pairs = [(c[0], c[1]) for c in fake.calls]
    if few_shot:        # engine.py:53-57: the forget loader drives, remain is prefetched and recycled
        assert pairs == [(100, 200), (101, 201), (100, 202), (101, 203), (100, 204)]
    else:               # engine.py:236-: remain drives, forget recycled
        assert pairs == [(100, 200), (101, 201)]
    assert ret[0] == len(pairs)
    assert all(c[4]["group_type"] == "lora" for c in fake.calls)
    assert all(c[2] == 4 and c[3] == 3 for c in fake.calls)               # remain / forget tensors never swap roles


def test_engine_py_alpha_epoch_gate_and_unsupported_group_pos(fake):
    remain, forget = _loader(100, 2, 4), _loader(200, 2, 3)
    m = _meters()
    cfg = {"few_shot": False, "ALPHA_EPOCH": 3, "GROUP_TYPE": "block", "GROUP_POS": "FFN"}
    args = (torch.nn.Linear(1, 1), forget, remain, "cpu", torch.nn.CrossEntropyLoss(), None)
    engine.train_one_epoch(*args, 2, m[0], m[1], m[2], m[3], m[4], m[5], 0.15, 1e-2, 105.0, 0, None, None, 0.0, 0.0, cfg)
    assert all(c[4]["alpha"] == 0.0 for c in fake.calls)                  # epoch 2 < ALPHA_EPOCH 3: structure term off (engine.py:82-90)
    fake.calls.clear()
    ret = engine.train_one_epoch(*args, 3, *_meters(6), 0.15, 1e-2, 105.0, 0, None, None, 0.0, 0.0, cfg)
    assert all(c[4]["alpha"] == 1e-2 for c in fake.calls)
    assert ret[7].avg == pytest.approx(1e-2 * 4.0)
    with pytest.raises(ValueError):     # GROUP_POS must name where the model's LoRA lives (this model: FFN); the reference would sum an empty group list
        engine.train_one_epoch(*args, 0, *_meters(6), 0.15, 1e-2, 105.0, 0, None, None, 0.0, 0.0, dict(cfg, GROUP_POS="Attention"))


def test_prototype_loss_torch_form_matches_the_reference_expression():
    """engine_cl.get_prototype_loss (engine_cl.py:571-603): dict of per-class CPU tensors or a [C, D] table, 'kl' and 'l2'."""
    g = torch.Generator().manual_seed(0)
    emb = torch.randn(6, 16, generator=g)
    labels = torch.tensor([0, 2, 2, 1, 0, 1])
    protos = {k: torch.randn(16, generator=g) for k in range(3)}
    stacked = torch.stack([protos[int(l)] for l in labels])
    want = torch.nn.functional.kl_div(torch.log_softmax(emb, 1), torch.log_softmax(stacked, 1), reduction="batchmean", log_target=True)
    assert torch.allclose(engine_cl.get_prototype_loss(emb, labels, protos), want)
    table = torch.stack([protos[k] for k in range(3)])
    assert torch.allclose(engine_cl.get_prototype_loss(emb, labels, table), want)
    assert torch.allclose(engine_cl.get_prototype_loss(emb, labels, protos, distance="l2"), torch.mean((emb - stacked) ** 2))


# ------------------------------------------------------------------------------------------------ against the UNMODIFIED reference epoch
class _OracleStep:
    """unlearn_step_async stand-in that executes the step with the CPU oracle (so the arithmetic is the reference's) -- what is under test is
    engine_cl.train_one_epoch's own host logic around it, against a golden recorded from the unmodified reference loop."""

    def __init__(self, sd, cfg, hp, prototypes):
        from oracle import vit_oracle as O
        self.O, self.sd, self.cfg, self.hp, self.protos, self.state = O, sd, cfg, hp, prototypes, {}

    def __call__(self, model, xr, yr, xf, yf, *, beta, alpha, BND, optimizer=None, use_prototype=False, prototype_dict=None,
                 prototype_weight_forget=0.0, prototype_weight_remain=0.0, BND_pro=0.0, **kw):
        pk = dict(prototypes=self.protos, w_pf=prototype_weight_forget, w_pr=prototype_weight_remain, BND_pro=BND_pro) if use_prototype else {}
        if "group_type" in kw:
            pk["group_type"] = kw["group_type"]
        out, _ = self.O.unlearn_step(self.sd, self.cfg, self.state, xr, yr, xf, yf, lr=self.hp["lr"], wd=self.hp["wd"], beta=beta, alpha=alpha,
                                     BND=BND, **pk)
        vals = dict(loss_remain=float(out["loss_remain"]), ce_forget=float(out["ce_forget"]), loss_forget=float(out["loss_forget"]),
                    structure=float(out["structure"]), top1_remain=float(out["top1_r"]), top1_forget=float(out["top1_f"]),
                    proto_forget=float(out["proto_forget"]), proto_remain=float(out["proto_remain"]), total=float(out["total"]))
        return types.SimpleNamespace(wait=lambda: vals)


@pytest.mark.parametrize("case", ["plain", "proto"])
def test_epoch_loop_reproduces_the_unmodified_reference_epoch(golden_dir, monkeypatch, case):
    """tests/golden/make_golden_epoch.py ran the UNMODIFIED engine_cl.train_one_epoch (7 remain batches, 3 recycled forget batches, display reset
    at batch 5, with and without the prototype term).  The drop-in's loop, stepping through the oracle, must return the same batch counter,
    the same eight meters (val / avg / sum / count) and the same LoRA parameters."""
    import os
    from oracle import vit_oracle as O
    g = torch.load(os.path.join(golden_dir, "tiny6_epoch.pt"), weights_only=False)[case]
    cfg = O.VitConfig(**g["cfg"])
    sd = O.init_state_dict(cfg, seed=g["seed"])
    for k, v in g["state_dict_checksum"].items():
        assert abs(float(sd[k].double().abs().sum()) - v) <= 1e-9 * max(1.0, abs(v)), f"weight regen drift: {k}"
    gen = torch.Generator().manual_seed(g["loader_seed"])
    S = cfg.image_size
    remain = [(torch.rand(4, 3, S, S, generator=gen), torch.randint(0, cfg.num_class, (4,), generator=gen)) for _ in range(g["n_remain"])]
    forget = [(torch.rand(3, 3, S, S, generator=gen), torch.randint(0, cfg.num_class, (3,), generator=gen)) for _ in range(g["n_forget"])]
    hp = g["hp"]
    step = _OracleStep(sd, cfg, hp, g["prototypes"])
    monkeypatch.setattr(engine_cl, "unlearn_step_async", step)
    monkeypatch.setattr(engine_cl, "_Prefetcher", _CpuPrefetcher)
    monkeypatch.setattr(engine_cl, "engine_fresh_optimizer", lambda m, o: False)
    monkeypatch.setattr(engine_cl, "sync_optimizer_state", lambda m, o: None)
    lf, lr, lt, ls, tf, tr, lpf, lpr = _meters()
    proto_dict = {i: g["prototypes"][i] for i in range(cfg.num_class)}
    ret = engine_cl.train_one_epoch(torch.nn.Linear(1, 1), forget, remain, "cpu", torch.nn.CrossEntropyLoss(), None, 0, lf, lr, lt, ls, tf, tr,
                                    hp["beta"], hp["alpha"], hp["BND"], 0, None, None, 0.0, 0.0, {"BND_pro": hp["BND_pro"]}, 2, g["use_proto"], proto_dict,
                                    hp["w_pf"], hp["w_pr"], lpf, lpr)
    assert ret[0] == g["batch"] and ret[1] == g["highest_H_mean"]
    names = ["losses_forget", "losses_remain", "top1_forget", "top1_remain", "losses_total", "losses_structure", "losses_prototype_forget",
             "losses_prototype_remain"]
    for n, meter in zip(names, ret[2:]):
        want = g["meters"][n]
        assert meter.count == want["count"], n
        for field in ("val", "avg", "sum"):
            assert getattr(meter, field) == pytest.approx(want[field], rel=2e-4, abs=1e-6), (n, field)
    for n in O.lora_param_list(cfg):
        a, b = sd[n], g["params_after"][n]
        assert float((a.double() - b.double()).norm() / b.double().norm()) < 1e-3, n


@pytest.mark.parametrize("case", ["few_shot_lora", "gated_matrix", "open_block"])
def test_single_step_twin_reproduces_the_unmodified_reference_engine_py_epoch(golden_dir, monkeypatch, case):
    """tests/golden/make_golden_epoch_single.py ran the UNMODIFIED engine.train_one_epoch: the few-shot branch (forget loader drives, remain
    recycled, GROUP_TYPE lora), the ordinary branch with the structure term gated off by ALPHA_EPOCH (GROUP_TYPE matrix), and the ordinary
    branch with a longer forget loader but few_shot off at epoch == ALPHA_EPOCH (GROUP_TYPE block).  gs-lora_b200/engine.py, stepping through
    the oracle, must return the same batch counter, meters and LoRA parameters."""
    import os
    from oracle import vit_oracle as O
    g = torch.load(os.path.join(golden_dir, "tiny3_epoch_single.pt"), weights_only=False)[case]
    cfg = O.VitConfig(**g["cfg"])
    sd = O.init_state_dict(cfg, seed=g["seed"])
    for k, v in g["state_dict_checksum"].items():
        assert abs(float(sd[k].double().abs().sum()) - v) <= 1e-9 * max(1.0, abs(v)), f"weight regen drift: {k}"
    gen = torch.Generator().manual_seed(g["loader_seed"])
    S = cfg.image_size
    remain = [(torch.rand(4, 3, S, S, generator=gen), torch.randint(0, cfg.num_class, (4,), generator=gen)) for _ in range(g["n_remain"])]
    forget = [(torch.rand(3, 3, S, S, generator=gen), torch.randint(0, cfg.num_class, (3,), generator=gen)) for _ in range(g["n_forget"])]
    hp = g["hp"]
    step = _OracleStep(sd, cfg, hp, None)
    monkeypatch.setattr(engine, "unlearn_step_async", step)
    monkeypatch.setattr(engine_cl, "_Prefetcher", _CpuPrefetcher)
    monkeypatch.setattr(engine_cl, "engine_fresh_optimizer", lambda m, o: False)
    monkeypatch.setattr(engine_cl, "sync_optimizer_state", lambda m, o: None)
    m = _meters()
    ret = engine.train_one_epoch(torch.nn.Linear(1, 1), forget, remain, "cpu", torch.nn.CrossEntropyLoss(), None, g["epoch"], m[0], m[1], m[2], m[3],
                                 m[4], m[5], hp["beta"], hp["alpha"], hp["BND"], 0, None, None, 0.0, 0.0, g["run_cfg"],
                                 losses_prototype_forget=m[6], losses_prototype_remain=m[7])
    assert ret[0] == g["batch"] and ret[1] == g["highest_H_mean"]
    names = ["losses_forget", "losses_remain", "top1_forget", "top1_remain", "losses_total", "losses_structure", "losses_prototype_forget",
             "losses_prototype_remain"]
    for n, meter in zip(names, ret[2:]):
        want = g["meters"][n]
        assert meter.count == want["count"], n
        for field in ("val", "avg", "sum"):
            assert getattr(meter, field) == pytest.approx(want[field], rel=2e-4, abs=1e-6), (n, field)
    for n in O.lora_param_list(cfg):
        a, b = sd[n], g["params_after"][n]
        assert float((a.double() - b.double()).norm() / b.double().norm()) < 1e-3, n


# ------------------------------------------------------------------------------------------------ evaluate(): H-mean and checkpoint rotation
def _run_evaluate(mod, workdir, accs, task_kw, monkeypatch):
    """Drive `mod.evaluate` through a sequence of (forget_acc, remain_acc) results; returns (H-mean trajectory, surviving checkpoint names)."""
    import itertools
    import os
    seq = iter(accs)
    cur = {}

    def fake_eval_data(model, loader, device, mode, batch=0):
        if mode.startswith("forget"):
            cur["pair"] = next(seq)
            return cur["pair"][0]
        return cur["pair"][1]

    tick = itertools.count()
    monkeypatch.setattr(mod, "eval_data", fake_eval_data)
    monkeypatch.setattr(mod, "get_time", lambda: f"t{next(tick):03d}")
    model = torch.nn.Linear(2, 2)
    opt = torch.optim.AdamW(model.parameters(), lr=0.01)
    cfg = {"WORK_PATH": str(workdir), "BACKBONE_NAME": "VIT", "MULTI_GPU": False}
    open(os.path.join(workdir, "config.txt"), "w").write("x")
    h, traj = -1.0, []
    for i in range(len(accs)):
        h = mod.evaluate(model, None, None, "cpu", batch=i, epoch=0, forget_acc_before=90.0, highest_H_mean=h, cfg=cfg, optimizer=opt, **task_kw)
        traj.append(h)
        for f in os.listdir(workdir):                      # make the mtime order unambiguous
            os.utime(os.path.join(workdir, f), None) if f.endswith(f"t{i:03d}_checkpoint.pth") else None
    return traj, sorted(f for f in os.listdir(workdir) if f.endswith(".pth"))


ACCS = [(80.0, 70.0), (60.0, 72.0), (65.0, 71.0), (30.0, 69.0), (10.0, 75.0), (90.0, 75.0)]      # improving, a dip, improving, then forget_drop = 0


def test_evaluate_hmean_and_checkpoint_rotation(tmp_path, monkeypatch):
    """engine_cl.evaluate (engine_cl.py:247-315): H = 2 d r / (d + r + 1e-8) with d = forget_acc_before - forget_acc; a checkpoint is written
    only when H improves and the oldest .pth is dropped once the directory holds 4 entries (config.txt + 3 checkpoints -> 2 survive)."""
    (tmp_path / "a").mkdir()
    traj, files = _run_evaluate(engine_cl, tmp_path / "a", ACCS, dict(task_i=1), monkeypatch)
    want, h = [], -1.0
    for fa, ra in ACCS:
        d = 90.0 - fa
        h = max(h, 2 * d * ra / (d + ra + 1e-8))
        want.append(h)
    assert traj == pytest.approx(want)
    assert len(files) == 2 and files[-1].startswith("Backbone_VIT_Epoch_1_Batch_5_")         # the 5th call (batch index 4) was the last improvement
    # the single-step twin keeps one checkpoint fewer (engine.py:486: `>= 3`) and divides unguarded
    (tmp_path / "b").mkdir()
    traj2, files2 = _run_evaluate(engine, tmp_path / "b", ACCS[:5], {}, monkeypatch)
    assert traj2 == pytest.approx([2 * (90 - fa) * ra / ((90 - fa) + ra) for fa, ra in [ACCS[0], ACCS[1], ACCS[1], ACCS[3], ACCS[4]]])
    assert len(files2) == 1


def test_evaluate_matches_the_reference_functions_live(tmp_path, monkeypatch):
    """Same driver against the UNMODIFIED reference engine_cl.evaluate / engine.evaluate when the reference tree is present."""
    import importlib.util
    import os
    import sys
    ref_root = "/root/reference"
    if not os.path.isfile(os.path.join(ref_root, "engine_cl.py")):
        pytest.skip("reference tree not present")
    shims = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "shims")
    # the reference modules must see the reference's own `util` package, not the drop-in overlay this process has already imported
    before = set(sys.modules)
    for k in [k for k in sys.modules if k == "util" or k.startswith("util.")]:
        monkeypatch.delitem(sys.modules, k)
    monkeypatch.setattr(sys, "path", [shims, ref_root] + sys.path)
    monkeypatch.setenv("WANDB_MODE", "disabled")
    monkeypatch.setitem(sys.modules, "image_iter", types.SimpleNamespace(CustomSubset=type("CustomSubset", (), {})))

    def load(name):
        spec = importlib.util.spec_from_file_location("_ref_" + name, os.path.join(ref_root, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        assert mod.util.__file__.startswith(ref_root)
        return mod

    for name, ours, kw, accs in (("engine_cl", engine_cl, dict(task_i=1), ACCS), ("engine", engine, {}, ACCS[:5])):
        ref = load(name)
        (tmp_path / (name + "_ref")).mkdir()
        (tmp_path / (name + "_ours")).mkdir()
        t_ref, f_ref = _run_evaluate(ref, tmp_path / (name + "_ref"), accs, kw, monkeypatch)
        t_ours, f_ours = _run_evaluate(ours, tmp_path / (name + "_ours"), accs, kw, monkeypatch)
        assert t_ours == pytest.approx(t_ref) and f_ours == f_ref, name
    for k in set(sys.modules) - before:                 # reference modules imported on the way: leave no trace for later tests
        if k == "util" or k.startswith("util.") or k.startswith("_ref_"):
            sys.modules.pop(k, None)


def test_merged_state_dict_equals_an_eval_mode_deepcopy():
    """engine.evaluate saves `copy.deepcopy(model).eval().state_dict()` in the reference (engine.py:449-476): loralib merges W += B A / r in
    eval().  The twin builds the same dict without copying the model and without touching its mode."""
    import copy
    from oracle import vit_oracle as O
    from vit_pytorch_face import ViT_face
    cfg = O.TINY
    m = ViT_face(loss_type="CosFace", GPU_ID=[0], num_class=cfg.num_class, image_size=cfg.image_size, patch_size=cfg.patch_size, dim=cfg.dim,
                 depth=cfg.depth, heads=cfg.heads, mlp_dim=cfg.mlp_dim, lora_rank=cfg.lora_rank)
    m.load_state_dict(O.init_state_dict(cfg, seed=3), strict=True)
    m.train()
    before = {k: v.clone() for k, v in m.state_dict().items()}
    got = engine._merged_state_dict(m)
    want = copy.deepcopy(m).eval().state_dict()
    assert set(got) == set(want)
    for k in want:
        assert torch.equal(got[k], want[k]), k
    assert m.training and all(torch.equal(v, before[k]) for k, v in m.state_dict().items())
    w = "transformer.layers.0.1.fn.fn.net.0.weight"
    assert not torch.equal(got[w], before[w])                        # the LoRA delta is in the saved weight
    assert torch.equal(engine._merged_state_dict(m.eval())[w], got[w])      # already merged: nothing is added twice


# ------------------------------------------------------------------------------------------------ data parallel sharding of the epoch loops
class _FakeDist:
    def __init__(self, rank, world):
        self.rank, self.world = rank, world

    def get_rank(self):
        return self.rank

    def get_world_size(self):
        return self.world


@pytest.mark.parametrize("mod_name", ["engine_cl", "engine"])
def test_epoch_loops_shard_the_global_batch_by_rank(fake, monkeypatch, mod_name, capsys):
    """under torchrun every rank iterates the SAME seeded loaders; rank r of w must step on samples r, r + w, ... of each global
    batch (remain and forget alike) while the meters keep weighing by the global batch sizes.  The union over ranks is the global batch."""
    x = torch.arange(10, dtype=torch.float32).view(10, 1)
    y = torch.arange(10)
    assert engine_cl.shard_batch(x, y)[0] is x                                       # no process group: untouched
    seen = {}
    for rank in range(3):
        monkeypatch.setattr(engine_cl, "_dist", lambda r=rank: _FakeDist(r, 3))
        xs, ys = engine_cl.shard_batch(x, y)
        assert xs[:, 0].tolist() == list(range(rank, 10, 3)) and ys.tolist() == list(range(rank, 10, 3))
        seen[rank] = set(ys.tolist())
    assert set().union(*seen.values()) == set(range(10)) and sum(len(v) for v in seen.values()) == 10
    # through the loops: 4-image remain batches, 2-image forget batches, rank 1 of 2
    monkeypatch.setattr(engine_cl, "_dist", lambda: _FakeDist(1, 2))
    remain, forget = _loader(100, 3, 4), _loader(200, 2, 2)
    m = _meters()
    if mod_name == "engine_cl":
        ret = engine_cl.train_one_epoch(torch.nn.Linear(1, 1), forget, remain, "cpu", torch.nn.CrossEntropyLoss(), None, 0, m[0], m[1], m[2], m[3], m[4],
                                        m[5], 0.15, 1e-2, 105.0, 0, None, None, 0.0, 0.0, {}, 0, False, None, 0.0, 0.0, m[6], m[7])
    else:
        ret = engine.train_one_epoch(torch.nn.Linear(1, 1), forget, remain, "cpu", torch.nn.CrossEntropyLoss(), None, 0, m[0], m[1], m[2], m[3], m[4],
                                     m[5], 0.15, 1e-2, 105.0, 0, None, None, 0.0, 0.0, {"few_shot": False}, losses_prototype_forget=m[6],
                                     losses_prototype_remain=m[7])
    assert [(c[2], c[3]) for c in fake.calls] == [(2, 1)] * 3                          # each step saw half of each global batch
    assert ret[0] == 3 and ret[3].count == 12 and ret[2].count == 6                    # meters: global sizes (3 steps x 4 remain / 2 forget)
    assert capsys.readouterr().out == ""                                               # rank 1 does not print


def test_prototype_table_is_built_once_and_flags_missing_labels():
    """engine_cl._cached_prototype_table: the driver's {label: CPU tensor} dict (util/utils.py:546-549) becomes ONE dense table per dict object
    (no per-class copies per step); classes without a prototype are recorded so that a batch label without one can be reported (the reference's
    dict lookup raises KeyError, engine_cl.py:585-590)."""
    g = torch.Generator().manual_seed(1)
    protos = {k: torch.randn(8, generator=g) for k in (0, 2, 5)}
    holder = types.SimpleNamespace()
    t1, present = engine_cl._cached_prototype_table(holder, protos, 6, 8, "cpu")
    assert t1.shape == (6, 8) and present.tolist() == [True, False, True, False, False, True]
    assert torch.equal(t1[2], protos[2]) and float(t1[1].abs().sum()) == 0.0
    t2, _ = engine_cl._cached_prototype_table(holder, protos, 6, 8, "cpu")
    assert t2 is t1                                                       # same dict object: cached
    other = dict(protos)
    t3, _ = engine_cl._cached_prototype_table(holder, other, 6, 8, "cpu")
    assert t3 is not t1                                                   # a new dict (next task): rebuilt
    with pytest.raises(KeyError):
        engine_cl._prototype_tensor({7: torch.zeros(8)}, 6, 8, "cpu")     # prototype label outside [0, num_class)
    with pytest.raises(KeyError):
        engine_cl.get_prototype_loss(torch.randn(2, 8), torch.tensor([0, 1]), protos)     # label 1 has no prototype
    r = StepResult(torch.tensor([2.0, 1, 3.0, 1, 0, 0, 0, 0, 1.0, 1.0]), _FakeEvent(), 10, dict(beta=0.1, alpha=0.0, BND=5.0, BND_pro=0.0, pwf=0.0, pwr=0.0,
                                                                                            use_prototype=True))
    with pytest.raises(KeyError):
        r.wait()                                                          # the device-side "missing prototype" flag of the fused step
    nan = StepResult(torch.tensor([2.0, 1, 3.0, 1, 0, 0, 0, 0, float("nan")]), _FakeEvent(), 9, dict(beta=0.1, alpha=0.0, BND=5.0, BND_pro=0.0, pwf=0.0,
                                                                                                       pwr=0.0, use_prototype=False))
    with pytest.raises(FloatingPointError):
        nan.wait()                                                        # gsl_grouplasso_adamw_step skipped a group with a non-finite gradient
