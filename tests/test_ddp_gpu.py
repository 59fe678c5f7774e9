"""Data-parallel step on the GPU box's single GPU: two ranks share cuda:0 over gloo (GSLORA_DIST_BACKEND=gloo -- NCCL needs one device per
rank; the code path is otherwise the one torchrun + NCCL takes).  Each rank runs the unchanged call sequence on the SAME global batches,
engine_cl shards them by rank on the host, all-reduces the loss sums and the flat LoRA gradient; the tail batches hold ONE image, so rank 1's
share is empty and it has to join the collectives with zero sums and a zero gradient.  Checked: both ranks end with bit-identical LoRA
parameters and they match a single-process run over the same global batches."""
import os
import socket

import pytest
import torch

from oracle import vit_oracle as O

pytestmark = pytest.mark.gpu


def _batches(cfg):
    g = torch.Generator().manual_seed(5)
    S = cfg.image_size
    mk = lambda n: (torch.rand(n, 3, S, S, generator=g), torch.randint(0, cfg.num_class, (n,), generator=g))
    remain = [mk(6), mk(5), mk(1)]          # the tail batch has ONE image: rank 1 of 2 gets nothing of it ...
    forget = [mk(4), mk(3), mk(1)]          # ... nor of the forget tail it is paired with
    return remain, forget


def _run(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [os.path.join(root, "gs-lora_b200"), root, os.path.join(root, "tests")]
    if world > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK="0", GSLORA_DIST_BACKEND="gloo")
    import engine_cl
    from engine_cl import AverageMeter
    from test_engine_gpu import build_model
    cfg = O.VitConfig(**{**O.TINY.to_dict(), "depth": 3})
    model = build_model(cfg, O.init_state_dict(cfg, seed=9))
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-2, weight_decay=0.05)
    remain, forget = _batches(cfg)
    m = [AverageMeter() for _ in range(8)]
    ret = engine_cl.train_one_epoch(model, forget, remain, torch.device("cuda"), torch.nn.CrossEntropyLoss(), opt, 0, m[0], m[1], m[2], m[3], m[4], m[5],
                                    0.15, 1e-2, 105.0, 0, None, None, 0.0, 0.0, {"WORK_PATH": out_dir, "BACKBONE_NAME": "VIT"}, 0, False, None, 0.0, 0.0,
                                    m[6], m[7])
    assert ret[0] == 3
    if world > 1:
        assert engine_cl._dist() is not None and engine_cl._dist().get_world_size() == world
    flat = torch.cat([p.detach().flatten() for p in model.lora_parameters()]).cpu()
    torch.save(dict(flat=flat, remain_meter=(ret[3].sum, ret[3].count)), os.path.join(out_dir, f"w{world}_r{rank}.pt"))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def test_two_ranks_with_an_empty_share_match_the_single_process_run(tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_run, args=(r, 2, port, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0, "a rank failed or hung (an empty rank that skips a collective hangs the other one)"
    single = ctx.Process(target=_run, args=(0, 1, 0, str(tmp_path)))
    single.start(); single.join(300)
    assert single.exitcode == 0
    r0, r1, one = [torch.load(os.path.join(str(tmp_path), f)) for f in ("w2_r0.pt", "w2_r1.pt", "w1_r0.pt")]
    assert torch.equal(r0["flat"], r1["flat"])                                   # every rank applied the same update
    rel = float((r0["flat"] - one["flat"]).norm() / one["flat"].norm())
    print(f"2 ranks vs 1 process after 3 steps: rel diff of the LoRA parameters {rel:.2e}")
    assert rel < 2e-3                  # same global batches; fp16 rounding differs with the split + Adam's sign-like early steps
    assert r0["remain_meter"][1] == one["remain_meter"][1]                       # meters weigh by the GLOBAL batch sizes
