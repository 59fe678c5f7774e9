"""Free-running trajectory parity (GPU): the engine steps on its own for 50 optimizer steps -- nothing is teacher-forced -- from loralib's
start (lora_B = 0), lr 1e-2, with the ALPHA_EPOCH switch to a structure weight large enough that the group lasso collapses most blocks and a
forget bound the run reaches.  This is the regime the merged-weight design must survive: W' = W + s B A is re-rounded every step while the
delta starts at exactly zero and most groups are driven back TOWARDS zero.

  * tiny6: the drop-in engine.train_one_epoch against the records of the UNMODIFIED reference loop (tests/golden/tiny6_trajectory.pt,
    tests/golden/make_golden_trajectory.py; the oracle is pinned to the same file on CPU by tests/test_trajectory_cpu.py)
  * P8S8 at bs 32+32: engine_cl.unlearn_step against the oracle stepping beside it in FP32 on the same GPU
Bars: every loss record within 1 %, final per-group norms within 1 %, the same set of collapsed groups, CE_forget first reaches BND at the
same step +- 1."""
import os

import pytest
import torch

from oracle import vit_oracle as O
from trajectory_common import group_norms, oracle_trajectory, trajectory_loaders, windows
from test_engine_gpu import build_model

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["split", "fast"])
def test_tiny6_epoch_loop_free_running_vs_unmodified_reference_records(golden_dir, mode):
    import engine
    import engine_cl
    from engine_cl import AverageMeter
    g = torch.load(os.path.join(golden_dir, "tiny6_trajectory.pt"), weights_only=False)
    cfg, hp = O.VitConfig(**g["cfg"]), g["hp"]
    sd = O.init_state_dict(cfg, seed=hp["seed"], lora_b_std=0.0)
    model = build_model(cfg, sd)
    model.gsl_precision = mode
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=hp["lr"], weight_decay=hp["wd"])
    records = []
    orig = engine_cl._wandb_log
    engine_cl._wandb_log = lambda d: records.append({k: float(v) for k, v in d.items()})
    try:
        run_cfg = {"few_shot": False, "ALPHA_EPOCH": hp["alpha_epoch"], "NUM_LAYERS": cfg.depth, "GROUP_TYPE": "block", "GROUP_POS": "FFN",
                   "WORK_PATH": "/tmp", "BACKBONE_NAME": "VIT", "MULTI_GPU": False}
        batch, norms = 0, []
        for epoch, n in enumerate(hp["steps"]):
            remain, forget = trajectory_loaders(cfg, hp["seed"] + 10, n, hp["batch"])
            m = [AverageMeter() for _ in range(8)]
            ret = engine.train_one_epoch(model, forget, remain, torch.device("cuda"), torch.nn.CrossEntropyLoss(), opt, epoch, m[0], m[1], m[2], m[3],
                                         m[4], m[5], hp["beta"], hp["alpha"], hp["BND"], batch, None, None, 0.0, 0.0, run_cfg,
                                         losses_prototype_forget=m[6], losses_prototype_remain=m[7])
            batch = int(ret[0])
            norms.append(group_norms({n_: model.get_parameter(n_).detach().cpu() for n_ in O.lora_param_list(cfg)}, cfg))
    finally:
        engine_cl._wandb_log = orig
    assert batch == g["batch"] and len(records) == len(g["records"])
    worst_rec = max(abs(r[k] - q[k]) / max(1.0, abs(q[k])) for r, q in zip(records, g["records"])
                    for k in ("epoch_loss_forget", "epoch_loss_remain", "epoch_loss_total", "epoch_loss_structure"))
    worst_norm = max(abs(a - b) / b for e in range(2) for a, b in zip(norms[e], g["group_norms"][e]))
    print(f"tiny6 trajectory [{mode}]: worst loss record {worst_rec:.2e}, worst final group norm {worst_norm:.2e}; norms {[round(x, 4) for x in norms[1]]}")
    assert worst_rec < 1e-2 and worst_norm < 1e-2
    collapsed = [n1 < 0.25 * n0 for n0, n1 in zip(*norms)]
    assert collapsed == [n1 < 0.25 * n0 for n0, n1 in zip(*g["group_norms"])] and any(collapsed) and not all(collapsed)


@pytest.mark.parametrize("mode", ["split", "fast"])
def test_p8s8_bs32_free_running_steps_vs_oracle(mode):
    import engine_cl
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = O.P8S8
    hp = dict(lr=1e-2, wd=0.05, beta=0.15, alpha=0.5, BND=0.0, alpha_epoch=1, steps=(15, 35), batch=32, seed=1337)
    sd = O.init_state_dict(cfg, seed=hp["seed"], lora_b_std=0.0)

    def loaders(epoch, n):
        return trajectory_loaders(cfg, hp["seed"] + 10, n, hp["batch"])
    # bound: CE_forget starts ~1 below it, so the gate relu(BND - CE_f) closes within the run
    (xr, yr), (xf, yf) = [l[0] for l in loaders(0, 1)]
    with torch.no_grad():
        out0 = O.unlearn_losses({k: v.cuda() for k, v in sd.items()}, cfg, xr.cuda(), yr.cuda(), xf.cuda(), yf.cuda(), beta=hp["beta"], alpha=0.0, BND=1e9)
    hp["BND"] = float(out0["ce_forget"]) + 1.0
    ref_steps, ref_norms, _ = oracle_trajectory(cfg, sd, hp, device="cuda", loaders=loaders)
    model = build_model(cfg, sd)
    model.gsl_precision = mode
    steps, norms = [], []
    for epoch, n in enumerate(hp["steps"]):
        remain, forget = loaders(epoch, n)
        alpha = 0.0 if epoch < hp["alpha_epoch"] else hp["alpha"]
        for (a, b), (c, d) in zip(remain, forget):
            out = engine_cl.unlearn_step(model, a.cuda(), b.cuda(), c.cuda(), d.cuda(), beta=hp["beta"], alpha=alpha, BND=hp["BND"],
                                         hparams=dict(lr=hp["lr"], wd=hp["wd"]))
            steps.append(out)
        norms.append(group_norms({n_: model.get_parameter(n_).detach().cpu() for n_ in O.lora_param_list(cfg)}, cfg))
    worst_loss = max(abs(s["total"] - r["total"]) / abs(r["total"]) for s, r in zip(steps, ref_steps))
    worst_norm = max(abs(a - b) / b for a, b in zip(norms[1], ref_norms[1]))

    def first_cross(seq):
        return next((i for i, s in enumerate(seq) if s["ce_forget"] >= hp["BND"]), None)
    print(f"P8S8 trajectory [{mode}]: worst per-step loss {worst_loss:.2e}, worst final group norm {worst_norm:.2e}, CE_f reaches BND at step "
          f"{first_cross(steps)} (oracle {first_cross(ref_steps)}); norms {[round(x, 3) for x in norms[1]]} vs {[round(x, 3) for x in ref_norms[1]]}")
    assert worst_loss < 1e-2 and worst_norm < 1e-2
    assert first_cross(ref_steps) is not None and abs(first_cross(steps) - first_cross(ref_steps)) <= 1
