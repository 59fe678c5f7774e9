"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol include/gslora.h declares, the drop-in module
surface matches the reference's (state_dict names, loralib merge semantics, freezing), and the data-parallel formulation
(all-reduced loss sums -> per-sample weights -> summed LoRA gradients) equals the single-process step, checked with
two gloo ranks."""
import copy
import os
import re
import socket

import types

import pytest
import torch
import torch.multiprocessing as mp

from oracle import vit_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol():
    from gslora import _ffi
    L = _ffi.lib()
    header = open(os.path.join(ROOT, "include", "gslora.h")).read()
    declared = set(re.findall(r"\b(gsl_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(L, name), f"libgslora.so does not export {name}"
    assert set(_ffi.EXPORTS) == declared
    assert L.gsl_version() >= 100
    # workspace sizing is host arithmetic and must work without a device
    import ctypes
    cfg = _ffi.GslConfig(image_size=112, patch_size=8, channels=3, dim=512, depth=6, heads=8, mlp_dim=2048, num_class=100, lora_rank=8,
                         max_batch=64, num_slots=2, patch_order=0, attn_scale=512 ** -0.5, ln_eps=1e-5, cos_s=64, cos_m=0.35,
                         lora_scaling=0.125, grad_scale=1024)
    assert L.gsl_engine_workspace_bytes(ctypes.byref(cfg)) > 1 << 28
    bad = _ffi.GslConfig(image_size=112, patch_size=8, channels=3, dim=500, depth=6, heads=8, mlp_dim=2048, num_class=100, lora_rank=8,
                         max_batch=64, num_slots=2)
    assert L.gsl_engine_workspace_bytes(ctypes.byref(bad)) == 0 and b"dim" in L.gsl_last_error()


def make_model(cfg=O.TINY, **kw):
    from vit_pytorch_face import ViT_face
    return ViT_face(loss_type="CosFace", GPU_ID=[0], num_class=cfg.num_class, image_size=cfg.image_size, patch_size=cfg.patch_size,
                    dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, mlp_dim=cfg.mlp_dim, lora_rank=cfg.lora_rank, **kw)


def test_precision_modes_by_name_and_default(monkeypatch):
    """GslConfig.precision values behind the Python surface (include/gslora.h): split8 is the default, the environment can select any mode,
    an unknown name fails loudly instead of falling back."""
    import pytest
    from gslora import _ffi as F
    assert F.PRECISION_BY_NAME == {"fast": 0, "split": 1, "split8": 2}
    monkeypatch.delenv("GSLORA_PRECISION", raising=False)
    assert F.default_precision() == F.PRECISION_SPLIT8
    for name, val in F.PRECISION_BY_NAME.items():
        monkeypatch.setenv("GSLORA_PRECISION", name.upper())
        assert F.default_precision() == val
    monkeypatch.setenv("GSLORA_PRECISION", "bf16")
    with pytest.raises(F.GslError):
        F.default_precision()


def test_state_dict_names_and_counts_match_reference_surface():
    import loralib as lora
    m = make_model(O.P8S8, dropout=0.1, emb_dropout=0.1)
    sd = O.init_state_dict(O.P8S8)
    assert set(m.state_dict().keys()) == set(sd.keys())
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
    assert sum(p.numel() for p in m.parameters()) == 19403264                 # SURVEY 8b [probed]
    lora.mark_only_lora_as_trainable(m)
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == 245760  # train_own_forget_cl.py:483 ratio
    assert all(("lora_" in n) == p.requires_grad for n, p in m.named_parameters())
    assert m.load_state_dict(sd, strict=True).missing_keys == []


def test_loralib_merge_unmerge_semantics():
    import loralib as lora
    torch.manual_seed(0)
    lin = lora.Linear(32, 48, r=8)
    assert lin.scaling == 1 / 8 and not lin.weight.requires_grad and lin.lora_A.shape == (8, 32) and lin.lora_B.shape == (48, 8)
    assert torch.count_nonzero(lin.lora_B) == 0                                # loralib init: B = 0, A kaiming
    with torch.no_grad():
        lin.lora_B.normal_()
    w0 = lin.weight.detach().clone()
    gen = lin._gsl_generation
    lin.eval()
    assert lin.merged and torch.allclose(lin.weight, w0 + (lin.lora_B @ lin.lora_A) / 8, atol=1e-6) and lin._gsl_generation == gen + 1
    lin.eval()
    assert lin._gsl_generation == gen + 1                                      # idempotent
    lin.train()
    assert not lin.merged and torch.allclose(lin.weight, w0, atol=1e-6)
    ml = lora.MergedLinear(32, 96, r=0, enable_lora=[True, True, True], bias=False)
    assert ml.bias is None and not hasattr(ml, "lora_A")
    # r > 0: loralib.MergedLinear's per-slice pairs (lora_pos "Attention", vit_face.py:349-355) -- same merge arithmetic as the oracle restatement
    from oracle import loralib_restated as olora
    torch.manual_seed(3)
    mq = lora.MergedLinear(32, 96, r=4, enable_lora=[True, True, True], bias=False)
    oq = olora.MergedLinear(32, 96, r=4, enable_lora=[True, True, True], bias=False)
    assert mq.lora_A.shape == oq.lora_A.shape == (12, 32) and mq.lora_B.shape == oq.lora_B.shape == (96, 4) and not mq.weight.requires_grad
    with torch.no_grad():
        mq.lora_B.normal_(0, 0.1)
        oq.load_state_dict(mq.state_dict())
    w0 = mq.weight.detach().clone()
    mq.eval(); oq.eval()
    assert mq.merged and torch.allclose(mq.weight, oq.weight, atol=1e-7) and not torch.allclose(mq.weight, w0)
    assert torch.allclose(mq.weight[32:64], w0[32:64] + (mq.lora_B[32:64] @ mq.lora_A[4:8]) / 4, atol=1e-6)     # slice k uses (A_k, B_k) only
    mq.train()
    assert not mq.merged and torch.allclose(mq.weight, w0, atol=1e-6)
    with pytest.raises(NotImplementedError):
        lora.MergedLinear(32, 96, r=8, enable_lora=[True, False, True])


def test_model_is_a_regular_module_and_refuses_cpu_execution():
    from gslora import _ffi
    m = make_model()
    c = copy.deepcopy(m)
    assert c is not m and c._engine is None
    assert "ViT_face" in repr(m) and "CosFace" in repr(m)
    m.eval(); m.train()
    with pytest.raises(_ffi.GslError):
        m(torch.rand(2, 3, 40, 40))
    with pytest.raises(RuntimeError):
        m.transformer(torch.rand(1, 26, 128))


def test_cal_norm_groupings():
    from util import cal_norm
    assert len(cal_norm._ffn_groups(6, "block")) == 6 and len(cal_norm._ffn_groups(6, "lora")) == 12 and len(cal_norm._ffn_groups(6, "matrix")) == 24
    assert cal_norm._ffn_groups(2, "block")[1] == [(1, 0), (1, 1), (1, 2), (1, 3)]


def test_cal_norm_imagenet_always_reports_twelve_block_groups():
    """util/cal_norm.py:82-99: with imagenet=True the reference ignores group_num / group_type (the driver passes group_num=vit_depth=6 for
    ViT-B/16) and reports the first 12 encoder blocks."""
    from util import cal_norm

    class FakeEngine:
        class spec:
            depth = 12

        def tensor_norms(self, type):
            return torch.arange(48, dtype=torch.float32)

    model = types.SimpleNamespace(_engine=FakeEngine(), sync_engine=lambda: None)
    out = cal_norm.get_norm_of_lora(model, type="L2", group_num=6, group_type="matrix", imagenet=True)
    assert len(out) == 12 and float(out[0]) == 0 + 1 + 2 + 3 and float(out[11]) == 44 + 45 + 46 + 47
    out6 = cal_norm.get_norm_of_lora(model, type="L1", group_num=6, group_type="lora")
    assert len(out6) == 12 and float(out6[0]) == 0 + 1 and float(out6[6]) == 2 + 3      # (A, B) of net.0 for all blocks, then of net.3
    FakeEngine.spec.depth = 6
    with pytest.raises(KeyError):
        cal_norm.get_norm_of_lora(model, type="L2", group_num=6, imagenet=True)


# ---------------------------------------------------------------------------------------------- data parallel (gloo, 2 ranks)
def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _dp_worker(rank, world, port, ret):
    import torch.distributed as dist
    import torch.nn.functional as Fn
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    cfg = O.TINY
    sd = O.init_state_dict(cfg, seed=21)
    g = torch.Generator().manual_seed(22)
    Bg = 6
    xr, xf = torch.rand(Bg, 3, 40, 40, generator=g), torch.rand(Bg, 3, 40, 40, generator=g)
    yr, yf = torch.randint(0, 10, (Bg,), generator=g), torch.randint(0, 10, (Bg,), generator=g)
    beta, BND = 0.15, 105.0
    sl = slice(rank, Bg, world)                                  # rank-strided slices of the global batches
    names = O.lora_param_list(cfg)
    work = {k: v.clone() for k, v in sd.items()}
    for n in names:
        work[n].requires_grad_(True)
    lr_, _ = O.vit_forward(work, cfg, xr[sl], yr[sl])
    lf_, _ = O.vit_forward(work, cfg, xf[sl], yf[sl])
    ce_r, ce_f = Fn.cross_entropy(lr_, yr[sl], reduction="none"), Fn.cross_entropy(lf_, yf[sl], reduction="none")
    sums = torch.tensor([ce_r.sum().item(), float(len(ce_r)), ce_f.sum().item(), float(len(ce_f))])
    dist.all_reduce(sums)                                        # what engine_cl.unlearn_step does with gsl_loss_sums' output
    gate = 1.0 if sums[2] / sums[3] < BND else 0.0
    local = ce_r.sum() / sums[1] - beta * gate * ce_f.sum() / sums[3]   # per-sample weights of gsl_unlearn_ce_grad
    grads = torch.autograd.grad(local, [work[n] for n in names])
    flat = torch.cat([t.flatten() for t in grads])
    dist.all_reduce(flat)                                        # the one flat LoRA-gradient allreduce
    if rank == 0:
        _, ref = O.unlearn_grads(sd, cfg, xr, yr, xf, yf, beta, 0.0, BND, include_structure=False)
        want = torch.cat([ref[n].flatten() for n in names])
        ret["rel"] = float((flat - want).norm() / want.norm())
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_formulation_two_gloo_ranks():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_dp_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret["rel"] < 1e-5, ret["rel"]


def test_driver_import_lines_resolve_on_the_overlay():
    """Every name the reference's drivers import from the shadowed modules exists with the reference's signature
    (train/train_own_forget_cl.py:15-55, train/train_own_forget.py:40)."""
    import inspect
    import engine
    import engine_cl
    import loralib as lora
    import vit_pytorch_face as vpf
    from util.cal_norm import get_norm_of_lora  # noqa: F401
    from util.utils import AverageMeter, calculate_prototypes, get_time, reinitialize_lora_parameters  # noqa: F401
    for name in ("train_one_epoch", "eval_data", "train_one_epoch_regularzation", "evaluate", "get_structure_loss", "get_prototype_loss"):
        assert callable(getattr(engine_cl, name)), name
    for name in ("train_one_epoch", "eval_data", "evaluate", "get_structure_loss", "get_prototype_loss"):
        assert callable(getattr(engine, name)), name
    for name in ("ViT_face", "ViT_face_low", "ViT_face_up", "ViTs_face", "ModifiedViT"):
        assert hasattr(vpf, name), name
    for name in ("Linear", "MergedLinear", "mark_only_lora_as_trainable"):
        assert hasattr(lora, name), name
    # positional order of the reference's train_one_epoch signatures (engine_cl.py:12-43, engine.py:13-43)
    cl = list(inspect.signature(engine_cl.train_one_epoch).parameters)
    assert cl[:8] == ["model", "dataloader_forget", "dataloader_remain", "device", "criterion", "optimizer", "epoch", "losses_forget"]
    assert cl[-8:] == ["task_i", "use_prototype", "prototype_dict", "prototype_weight_forget", "prototype_weight_remain",
                       "losses_prototype_forget", "losses_prototype_remain", "dataloader_open"]
    sg = list(inspect.signature(engine.train_one_epoch).parameters)
    assert sg[:6] == cl[:6] and sg[21:] == ["cfg", "dataloader_open", "prototype_weight_forget", "prototype_weight_remain", "use_prototype",
                                            "prototype_dict", "losses_prototype_forget", "losses_prototype_remain"]
    assert list(inspect.signature(engine.get_structure_loss).parameters) == ["model", "num_layers", "group_type", "group_pos"]
    assert list(inspect.signature(calculate_prototypes).parameters) == ["backbone", "dataset", "batch_size", "device", "aug_num"]
    # the baselines' loop is the reference's own code: without the reference tree it must say so instead of silently doing something else
    import sys
    if not any(os.path.isfile(os.path.join(p, "baselines", "LIRFtrain.py")) for p in sys.path if p):
        with pytest.raises(NotImplementedError):
            engine_cl.get_reg_loss()


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's CPU path via the oracle port) prints ONE JSON line with the contract's keys; a tiny
    sample keeps this test to seconds (the driver's run uses the default 96 + 96 images per step)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-batch", "2"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"] == "unlearn-step images/sec ViT-P8S8 112px bs512" and d["config"]["workload"] == "p8s8_bs512"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # non-zero ranks of a torchrun launch exit silently
    env = dict(os.environ, RANK="1")
    out1 = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--ref-batch", "2"],
                          capture_output=True, text=True, timeout=600, cwd=root, env=env)
    assert out1.returncode == 0 and out1.stdout.strip() == ""


@pytest.mark.parametrize("with_reference", [True, False])
def test_image_iter_overlay_defuses_the_customsubset_landmine(with_reference):
    """SURVEY 8b landmine 2: the reference's CustomSubset cannot be constructed on torch >= 2.1.  The overlay executes the reference's own
    image_iter.py (when present) and adds `__getitems__`; without the reference tree it still provides CustomSubset."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = "/root/reference"
    if with_reference and not os.path.isfile(os.path.join(ref, "image_iter.py")):
        pytest.skip("reference tree not present")
    code = r'''
import sys, torch
sys.path[:0] = [%r] + ([%r, %r] if %r else [])      # drop-in first; shims for the reference's IPython / mxnet imports; the reference tree
import image_iter
assert image_iter.__file__.startswith(%r), image_iter.__file__
assert image_iter._ref_loaded == %r, getattr(image_iter, "_ref_error", None)
class DS(torch.utils.data.Dataset):
    targets = [0, 1, 2, 0]; classes = ["a", "b", "c"]
    def __getitem__(self, i): return torch.full((2,), float(i)), self.targets[i]
    def __len__(self): return 4
s = image_iter.CustomSubset(DS(), [3, 0, 2])
assert len(s) == 3 and s.classes == ["a", "b", "c"] and s.targets == [0, 1, 2, 0]
x, y = s[0]
assert float(x[0]) == 3.0 and y == 0
xb, yb = next(iter(torch.utils.data.DataLoader(s, batch_size=3)))
assert xb[:, 0].tolist() == [3.0, 0.0, 2.0] and yb.tolist() == [0, 0, 2]
if %r:
    assert hasattr(image_iter, "CLDatasetWrapper") and hasattr(image_iter, "ImageNet900Dataset")
print("ok")
''' % (os.path.join(root, "gs-lora_b200"), os.path.join(root, "oracle", "shims"), ref, with_reference, os.path.join(root, "gs-lora_b200"),
       with_reference, with_reference)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-3000:]


@pytest.mark.parametrize("driver", ["train/train_own_forget_cl.py", "train/train_own_forget.py"])
def test_unmodified_driver_import_block_resolves_on_the_overlay(driver):
    """Every top-level import of the UNMODIFIED reference drivers executes with `gs-lora_b200` first on sys.path (then the stand-ins for the
    third-party packages this image lacks, then the reference tree), and the hot-path names bind to this repo's modules."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = "/root/reference"
    if not os.path.isfile(os.path.join(ref, driver)):
        pytest.skip("reference tree not present")
    code = r'''
import ast, os, sys
pkg, shims, ref, driver = %r, %r, %r, %r
sys.path[:0] = [pkg, shims, ref]
os.environ["WANDB_MODE"] = "disabled"
tree = ast.parse(open(os.path.join(ref, driver)).read())
imports = [n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom))]
ns = {}
exec(compile(ast.Module(body=imports, type_ignores=[]), driver, "exec"), ns)
mine = {"ViT_face": "vit_pytorch_face", "ModifiedViT": "vit_pytorch_face", "train_one_epoch": None, "eval_data": None,
        "get_norm_of_lora": "util/cal_norm.py", "CustomSubset": "image_iter.py", "calculate_prototypes": "util/utils.py"}
for name, where in mine.items():
    if name not in ns:
        continue
    f = sys.modules[ns[name].__module__].__file__
    assert f.startswith(pkg), (name, f)
assert ns["lora"].__file__.startswith(pkg)
assert ns["train_one_epoch"].__module__ in ("engine_cl", "engine")
print("ok", len(imports))
''' % (os.path.join(root, "gs-lora_b200"), os.path.join(root, "oracle", "shims"), ref, driver)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd="/tmp")
    assert out.returncode == 0 and out.stdout.strip().splitlines()[-1].startswith("ok"), out.stderr[-3000:]


def _auto_dist_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import engine_cl
    import torch.distributed as dist
    assert not dist.is_initialized()
    d = engine_cl._dist()                                  # what torchrun + the unchanged driver relies on: created on first use
    assert d is not None and dist.is_initialized() and d.get_world_size() == world and d.get_rank() == rank
    x = torch.arange(7, dtype=torch.float32).view(7, 1)
    xs, ys = engine_cl.shard_batch(x, torch.arange(7))
    t = torch.zeros(7)
    t[ys] = 1.0
    d.all_reduce(t)                                        # every sample is owned by exactly one rank
    if rank == 0:
        ret["cover"] = t.tolist()
        ret["mine"] = ys.tolist()
    d.barrier()
    d.destroy_process_group()


def test_process_group_is_created_on_first_use_under_torchrun_env():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_auto_dist_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert ret["cover"] == [1.0] * 7 and ret["mine"] == [0, 2, 4, 6]
    # a single process (no torchrun environment) never creates a group
    import engine_cl
    assert engine_cl._dist() is None
