"""Free-running trajectory, CPU half: the oracle, stepping on its own for 50 optimizer steps (nothing teacher-forced), against the records of
the UNMODIFIED reference loop (tests/golden/tiny6_trajectory.pt, made by tests/golden/make_golden_trajectory.py from engine.train_one_epoch:
loralib's lora_B = 0 start, ALPHA_EPOCH switch, group-lasso collapse of most blocks, forget bound engaged).  Pins the oracle as the
trajectory reference the GPU test (tests/test_trajectory_gpu.py) compares the engine with."""
import os

import torch

from oracle import vit_oracle as O
from trajectory_common import COLLAPSED, oracle_trajectory, windows


def test_oracle_free_running_trajectory_matches_unmodified_reference_loop(golden_dir):
    g = torch.load(os.path.join(golden_dir, "tiny6_trajectory.pt"), weights_only=False)
    cfg, hp = O.VitConfig(**g["cfg"]), g["hp"]
    sd = O.init_state_dict(cfg, seed=hp["seed"], lora_b_std=0.0)
    for k, v in g["state_dict_checksum"].items():
        assert abs(float(sd[k].double().abs().sum()) - v) <= 1e-9 * max(1.0, abs(v)), k
    steps, norms, sd_end = oracle_trajectory(cfg, sd, hp)
    assert len(steps) == g["batch"] == sum(hp["steps"])
    win = windows(steps, hp)
    assert len(win) == len(g["records"])
    for w, r in zip(win, g["records"]):
        for k in ("epoch_loss_forget", "epoch_loss_remain", "epoch_loss_total", "epoch_loss_structure"):
            assert abs(w[k] - r[k]) <= 2e-3 * max(1.0, abs(r[k])), (k, w[k], r[k])
    for e in range(2):
        for a, b in zip(norms[e], g["group_norms"][e]):
            assert abs(a - b) <= 2e-3 * b, (e, a, b)
    # the scenario is the interesting one: the structure term collapses most groups, the data term keeps at least one alive
    collapsed = [n1 < COLLAPSED * n0 for n0, n1 in zip(*g["group_norms"])]
    assert any(collapsed) and not all(collapsed)
    # the forget bound engages inside the run: CE_f starts below BND and is at or above it at some later step
    assert steps[0]["ce_forget"] < hp["BND"] and any(s["ce_forget"] >= hp["BND"] for s in steps)
