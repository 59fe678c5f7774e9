"""Multi-rank run of the driver call sequence (tests/driver_replay.py) the way INTEGRATION.md launches the unmodified driver:

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 scripts/ddp_replay.py

Every rank runs the SAME script with the same seeds (the driver knows nothing about torch.distributed): engine_cl creates the NCCL group on
first use, binds cuda:LOCAL_RANK, shards each global batch on the host (rank r keeps samples r, r + world, ...), all-reduces the loss sums and
the flat LoRA gradient.  Checked at the end: every rank holds bit-identical LoRA parameters, and they equal (to fp32 summation order) the
parameters of a single-process run over the same global batches."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "gs-lora_b200"), ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def build():
    import loralib as lora
    from vit_pytorch_face import ViT_face
    torch.manual_seed(11)
    m = ViT_face(loss_type="CosFace", GPU_ID=[0], num_class=20, image_size=112, patch_size=8, dim=512, depth=3, heads=8, mlp_dim=2048,
                 dropout=0.0, emb_dropout=0.0, lora_rank=8)
    with torch.no_grad():
        m.pos_embedding.mul_(0.02)
        m.cls_token.mul_(0.02)
    lora.mark_only_lora_as_trainable(m)
    return m


def main():
    from driver_replay import replay
    import engine_cl
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    work = tempfile.mkdtemp(prefix=f"replay_rank{rank}_")
    out = replay(build(), image_size=112, num_class=20, device=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")) if world > 1 else 0),
                 work_path=work, num_tasks=2, epochs=2, batch_size=8, per_class=4, prototype=True, average_weight=False)
    m = out["model"]
    flat = torch.cat([p.detach().flatten() for p in m.lora_parameters()])
    d = engine_cl._dist()
    if world > 1:
        assert d is not None and d.get_world_size() == world and torch.cuda.current_device() == int(os.environ["LOCAL_RANK"])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        for r in range(world):
            assert torch.equal(gathered[r], gathered[0]), f"rank {r} diverged from rank 0"
    if rank == 0:
        ref_path = os.environ.get("GSLORA_REPLAY_REF")
        msg = f"ddp_replay world={world}: tasks={len(out['tasks'])} steps={[t['steps'] for t in out['tasks']]} |lora|={float(flat.norm()):.6f}"
        if ref_path and world == 1:
            torch.save(flat.cpu(), ref_path)
        elif ref_path and os.path.exists(ref_path):
            ref = torch.load(ref_path)
            rel = float((flat.cpu() - ref).norm() / ref.norm())
            msg += f"  rel diff vs single-process run {rel:.2e}"
            assert rel < 5e-3, rel       # same global batches; per-rank fp16 rounding noise + Adam (see tests/test_trajectory_gpu.py on conditioning)
        print(msg, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
