"""Free-running trajectory parity (GPU): the engine steps on its own for 50 optimizer steps -- nothing is teacher-forced -- from loralib's
start (lora_B = 0), lr 1e-2, with the ALPHA_EPOCH switch to a structure weight large enough that the group lasso collapses most blocks and a
forget bound the run reaches.  This is the regime the merged-weight design must survive: W' = W + s B A is re-rounded every step while the
delta starts at exactly zero and most groups are driven back TOWARDS zero.

How close can ANY implementation that meets the per-step gradient tolerance stay?  Adam turns every element's gradient into a ~lr-sized
update, so additive gradient error far below the 1e-3 bar re-directs the elements whose true gradient is smaller still: the FP32 oracle itself,
perturbed by 3e-4 * rms(g) of additive noise per step, ends 50 steps later with group norms 13 % away from its unperturbed self and per-step
losses 3.7 % away (measured, scripts in DESIGN.md section 2) -- while multiplicative error of the same size moves them by 4e-4.  The bars here
are therefore relative to that envelope, computed live: the oracle perturbed at the north-star tolerance (1e-3) with three noise seeds.
  * 5-step-window total loss, per-step total loss, final per-group norms: within max(1 %, 1.5 x envelope)
  * the same set of collapsed groups as the unperturbed reference
  * CE_forget first reaches BND at the same step +- 1 (or inside the envelope's spread)

  tiny6: the drop-in engine.train_one_epoch against the records of the UNMODIFIED reference loop (tests/golden/tiny6_trajectory.pt,
         tests/golden/make_golden_trajectory.py; the oracle is pinned to the same file on CPU by tests/test_trajectory_cpu.py)
  P8S8 at bs 32+32: engine_cl.unlearn_step against the oracle stepping beside it in FP32 on the same GPU"""
import os

import pytest
import torch

from oracle import vit_oracle as O
from trajectory_common import COLLAPSED, deviation, group_norms, noise_envelope, oracle_trajectory, trajectory_loaders, windows
from test_engine_gpu import build_model

pytestmark = pytest.mark.gpu
EPS = 1e-3          # the north-star per-step gradient tolerance the envelope is drawn at


def _check(tag, dev, env, ref_dev):
    print(f"{tag}: window total {dev['window_total']:.2e} (envelope {env['window_total']:.2e}), per-step total {dev['step_total']:.2e} "
          f"({env['step_total']:.2e}), final group norm {dev['final_norm']:.2e} ({env['final_norm']:.2e}), CE_f reaches BND at step "
          f"{dev['first_cross']} (reference {ref_dev['first_cross']}, envelope {env['first_cross']}), collapsed {dev['collapsed']}")
    for m in ("window_total", "step_total", "final_norm"):
        assert dev[m] <= max(1e-2, 1.5 * env[m]), (m, dev[m], env[m])
    assert dev["collapsed"] == ref_dev["collapsed"]
    if ref_dev["first_cross"] is not None:
        lo = min([ref_dev["first_cross"]] + env["first_cross"]) - 1
        hi = max([ref_dev["first_cross"]] + env["first_cross"]) + 1
        assert dev["first_cross"] is not None and lo <= dev["first_cross"] <= hi


@pytest.mark.parametrize("mode", ["split8", "split", "fast"])
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
    # envelope of the FP32 oracle (pinned to the same golden on CPU) under additive gradient noise at the tolerance
    ref_steps, ref_norms, _ = oracle_trajectory(cfg, sd, hp, device="cuda")
    for a, b in zip(ref_norms[-1], g["group_norms"][-1]):
        assert abs(a - b) <= 5e-3 * b                       # the GPU oracle reproduces the unmodified reference loop's end state
    env = noise_envelope(cfg, sd, hp, EPS, "cuda", ref=(ref_steps, ref_norms))
    ref_dev = deviation(ref_steps, ref_norms, ref_steps, ref_norms, hp)
    win_total = max(abs(r["epoch_loss_total"] - q["epoch_loss_total"]) / abs(q["epoch_loss_total"]) for r, q in zip(records, g["records"]))
    final_norm = max(abs(a - b) / b for a, b in zip(norms[-1], g["group_norms"][-1]))
    collapsed = [n1 < COLLAPSED * n0 for n0, n1 in zip(norms[0], norms[-1])]
    print(f"tiny6 trajectory [{mode}] vs unmodified reference loop: window total {win_total:.2e} (envelope {env['window_total']:.2e}), final group "
          f"norm {final_norm:.2e} ({env['final_norm']:.2e}); norms {[round(x, 4) for x in norms[-1]]} vs {[round(x, 4) for x in g['group_norms'][-1]]}")
    assert win_total <= max(1e-2, 1.5 * env["window_total"]) and final_norm <= max(1e-2, 1.5 * env["final_norm"])
    assert collapsed == ref_dev["collapsed"] == [n1 < COLLAPSED * n0 for n0, n1 in zip(*g["group_norms"])] and any(collapsed) and not all(collapsed)


@pytest.mark.parametrize("mode", ["split8", "split", "fast"])
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
    env = noise_envelope(cfg, sd, hp, EPS, "cuda", loaders=loaders, seeds=(0, 1), ref=(ref_steps, ref_norms))
    model = build_model(cfg, sd)
    model.gsl_precision = mode
    steps, norms = [], []
    for epoch, n in enumerate(hp["steps"]):
        remain, forget = loaders(epoch, n)
        alpha = 0.0 if epoch < hp["alpha_epoch"] else hp["alpha"]
        for (a, b), (c, d) in zip(remain, forget):
            out = engine_cl.unlearn_step(model, a.cuda(), b.cuda(), c.cuda(), d.cuda(), beta=hp["beta"], alpha=alpha, BND=hp["BND"],
                                         hparams=dict(lr=hp["lr"], wd=hp["wd"]))
            steps.append(dict(out, alpha=alpha))
        norms.append(group_norms({n_: model.get_parameter(n_).detach().cpu() for n_ in O.lora_param_list(cfg)}, cfg))
    print(f"P8S8 [{mode}] norms {[round(x, 3) for x in norms[-1]]} vs oracle {[round(x, 3) for x in ref_norms[-1]]}")
    _check(f"P8S8 bs32 trajectory [{mode}]", deviation(steps, norms, ref_steps, ref_norms, hp), env, deviation(ref_steps, ref_norms, ref_steps, ref_norms, hp))
