"""bench.py -- GS-LoRA unlearning step throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # gslora-b200 arm
    python bench.py --impl reference [...]                         # reference arm (CPU, oracle port of the reference)
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload ("p8s8_bs512"): BASELINE.json configs[1] -- ViT-P8S8 (112 px, patch 8, dim 512, depth 6, heads 8, mlp 2048), CASIA-100 head,
LoRA r = 8 on every FFN Linear; one step = engine_cl.py:59-125: forward(remain 512) + forward(forget 512) -> CE, relu(BND - CE_f),
group-lasso structure loss -> selective backward (LoRA grads only) -> [NCCL allreduce of the flat LoRA gradient] -> AdamW.
Synthetic data (torch.rand images, random labels), seeded synthetic weights.  images/s = (512 + 512) * N / t_step.
`value`: inputs resident in HBM.  `e2e`: the same step through the public API (engine_cl.unlearn_step) from pinned HOST
buffers: the H2D copy of both batches and the D2H read of the step's loss scalars are inside the timed region.
Dropout 0.1 / emb_dropout 0.1 as in the reference's ViT-P8S8 construction (train/train_own_forget_cl.py:217-218), train mode.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "gs-lora_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

BATCH = 512          # per stream (remain and forget each) per GPU
DROPOUT = 0.1        # train/train_own_forget_cl.py:217-218
HP = dict(lr=1e-2, wd=0.05, beta=0.15, alpha=1e-4, BND=105.0)      # scripts/run_cl_forget.sh:225-233
METRIC = "unlearn-step images/sec ViT-P8S8 112px bs512"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], burst=d["bf16_tflops"], sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm=6650.0, burst=1590.0, sustained=1400.0, source="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        mx = max((float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()), default=0.0)
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx or None, reasons=sorted(reasons), samples=len(self.rows))


FLOPS_PER_IMAGE = 15.648e9     # fwd 8.049 G + selective bwd 7.599 G, ViT-P8S8 d6 r8 (2mnk per GEMM)


class P8S8:
    image_size, patch_size, dim, depth, heads, mlp_dim, num_class, lora_rank, tokens = 112, 8, 512, 6, 8, 2048, 100, 8, 197


class VITB16:       # BASELINE.json configs[3]: torchvision vit_b_16 through ModifiedViT, LoRA r = 8 (scripts/run_cl_forget_image.sh: bs 48)
    image_size, patch_size, dim, depth, heads, mlp_dim, num_class, lora_rank, tokens = 224, 16, 768, 12, 12, 3072, 100, 8, 197


class VITL16:       # BASELINE.json configs[4]: ViT-L/16 widths, LoRA r = 16, global bs 256 over 8 GPUs = 32 per stream per GPU
    image_size, patch_size, dim, depth, heads, mlp_dim, num_class, lora_rank, tokens = 224, 16, 1024, 24, 16, 4096, 100, 16, 197


# extra workloads (parity-test configs of BASELINE.json, measurable with --workload; the driver's default line is p8s8_bs512)
WORKLOADS = {
    "p8s8_bs512": dict(cfg=P8S8, batch=512, flops=15.648e9, metric=METRIC, dropout=DROPOUT,
                       model="ViT-P8S8 depth 6 dim 512 heads 8 mlp 2048, 112x112, LoRA r=8 on FFN, CosFace 100 classes"),
    "vitb16_bs48": dict(cfg=VITB16, batch=48, flops=70.226e9, metric="unlearn-step images/sec ViT-B/16 224px bs48", dropout=0.0,
                        model="torchvision ViT-B/16 (ModifiedViT), 224x224, LoRA r=8 on mlp.0 / mlp.3, Linear head 100 classes"),
    "vitl16_bs32": dict(cfg=VITL16, batch=32, flops=250.748e9, metric="unlearn-step images/sec ViT-L/16 224px bs32", dropout=0.0,
                        model="torchvision ViT-L/16 (ModifiedViT), 224x224, LoRA r=16 on mlp.0 / mlp.3, Linear head 100 classes"),
}


def build_torchvision_model(device, cfg):
    """vit_b_16 / vit_l_16 (weights=None: no network) wrapped as the reference does (train_own_forget_cl.py:226-242) with every encoder MLP
    Linear swapped for a fresh loralib.Linear (util.utils.replace_ffn_with_lora, util/utils.py:552-576); lora_B ~ N(0, 0.02)."""
    import loralib as lora
    from torchvision.models import vit_b_16, vit_l_16
    from vit_pytorch_face import ModifiedViT
    torch.manual_seed(1337)
    tv = (vit_b_16 if cfg.dim == 768 else vit_l_16)(weights=None, num_classes=cfg.num_class)
    m = ModifiedViT(tv)
    for blk in m.encoder.layers.children():
        blk.mlp[0] = lora.Linear(cfg.dim, cfg.mlp_dim, r=cfg.lora_rank)
        blk.mlp[3] = lora.Linear(cfg.mlp_dim, cfg.dim, r=cfg.lora_rank)
    with torch.no_grad():
        for fc1, fc2 in m.lora_layers():
            fc1.lora_B.normal_(0, 0.02)
            fc2.lora_B.normal_(0, 0.02)
    lora.mark_only_lora_as_trainable(m)
    return m.to(device).train(), cfg


def build_model(device):
    """ViT-P8S8 with module-default random init under SEED 1337 (config.py:8); lora_B ~ N(0, 0.02) so both LoRA factors are live."""
    import loralib as lora
    from vit_pytorch_face import ViT_face
    torch.manual_seed(1337)
    m = ViT_face(loss_type="CosFace", GPU_ID=[0], num_class=100, image_size=112, patch_size=8, dim=512, depth=6, heads=8, mlp_dim=2048,
                 dropout=DROPOUT, emb_dropout=DROPOUT, lora_rank=8)
    with torch.no_grad():
        m.pos_embedding.mul_(0.02)
        m.cls_token.mul_(0.02)
        for fc1, fc2 in m.lora_layers():
            fc1.lora_B.normal_(0, 0.02)
            fc2.lora_B.normal_(0, 0.02)
    lora.mark_only_lora_as_trainable(m)
    return m.to(device).train(), P8S8


def run_gslora(args):
    import torch.distributed as dist
    import engine_cl
    from gslora import _ffi as F
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- gslora-b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = WORKLOADS[args.workload]
    BATCH = wl["batch"]
    # north-star denominator first (its 80 GB of fp32 autograd activations are freed before the engine's workspace is allocated)
    gpu_ref = None
    if world == 1 and args.workload == "p8s8_bs512" and not args.no_gpu_reference:
        gpu_ref = gpu_reference(dev, BATCH)
    model, cfg = build_model(dev) if args.workload == "p8s8_bs512" else build_torchvision_model(dev, wl["cfg"])
    model.gsl_precision = args.precision
    g = torch.Generator().manual_seed(100 + rank)
    S = cfg.image_size
    host = [torch.rand(BATCH, 3, S, S, generator=g).pin_memory(), torch.randint(0, 100, (BATCH,), generator=g).pin_memory(),
            torch.rand(BATCH, 3, S, S, generator=g).pin_memory(), torch.randint(0, 100, (BATCH,), generator=g).pin_memory()]
    devt = [t.to(dev) for t in host]
    alpha = HP["alpha"] if args.alpha is None else args.alpha
    step_kw = dict(beta=HP["beta"], alpha=alpha, BND=HP["BND"], hparams=dict(lr=HP["lr"], wd=HP["wd"]))

    # engine_cl.unlearn_step_async is what engine_cl.train_one_epoch calls: the step's scalars come back through a queued D2H copy into
    # pinned memory (every step, inside the timed region) and are read by the host one step later, so the launch queue never drains.
    def step_resident():
        return engine_cl.unlearn_step_async(model, devt[0], devt[1], devt[2], devt[3], **step_kw)

    # e2e: every step's inputs come from pinned HOST memory; as in the reference's util/data_prefetcher.py (and engine_cl._Prefetcher) the
    # H2D copy of step i+1 is issued on a side stream while step i computes, and step i waits for its own copy before it starts.
    copy_stream = torch.cuda.Stream(device=dev)
    staged = {"next": None}

    def issue_copy():
        with torch.cuda.stream(copy_stream):
            tens = [t.to(dev, non_blocking=True) for t in host]
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        staged["next"] = (tens, ev)

    def step_e2e():
        if staged["next"] is None:
            issue_copy()
        tens, ev = staged["next"]
        torch.cuda.current_stream().wait_event(ev)
        for t in tens:
            t.record_stream(torch.cuda.current_stream())
        issue_copy()                                   # next step's inputs (154 MB) travel while this step runs
        return engine_cl.unlearn_step_async(model, *tens, **step_kw)     # ends with the queued D2H copy of the loss scalars

    # the same e2e step fed with RAW uint8 pixels: 1 byte / pixel over PCIe, ToTensor's /255 applied inside the patchify kernel
    host_u8 = [torch.randint(0, 256, (BATCH, 3, S, S), dtype=torch.uint8, generator=g).pin_memory(), host[1],
               torch.randint(0, 256, (BATCH, 3, S, S), dtype=torch.uint8, generator=g).pin_memory(), host[3]]
    staged_u8 = {"next": None}

    def issue_copy_u8():
        with torch.cuda.stream(copy_stream):
            tens = [t.to(dev, non_blocking=True) for t in host_u8]
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        staged_u8["next"] = (tens, ev)

    def step_e2e_u8():
        if staged_u8["next"] is None:
            issue_copy_u8()
        tens, ev = staged_u8["next"]
        torch.cuda.current_stream().wait_event(ev)
        for t in tens:
            t.record_stream(torch.cuda.current_stream())
        issue_copy_u8()
        return engine_cl.unlearn_step_async(model, *tens, **step_kw)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = F.lib().gsl_launch_count()
        e0.record()
        prev = None
        for _ in range(steps):
            out = fn()
            if prev is not None:
                prev.wait()         # read step i-1's scalars while step i runs (what train_one_epoch does)
            prev = out
        out.wait()                  # the last step's read-back is inside the timed region too
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        launches = (F.lib().gsl_launch_count() - n0) // steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, out

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, launches, out = timed(step_resident, args.steps, args.warmup)
    # same W warm-up steps as the resident leg: the first e2e steps grow the caching allocator's pool on the copy stream (cudaMalloc of the
    # staged 77 MB image blocks synchronises the device) and must not leak into the timed region
    ms_e2e, _, _ = timed(step_e2e, args.steps, args.warmup)
    # --no-u8-leg: the ncu launch-list pass (scripts/gpu_round.sh) skips the extra uint8 leg -- every kernel costs seconds under ncu
    ms_e2e_u8 = timed(step_e2e_u8, args.steps, args.warmup)[0] if not args.no_u8_leg else float("nan")
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join(timeout=3)

    images = 2 * BATCH * world
    value = images / ms * 1e3
    e2e_value = images / ms_e2e * 1e3
    # the other precision modes beside it (resident leg only): "split8" (default: fp16 + e4m3 weight pair, the residual on the FP8 tensor path),
    # "split" (fp16 + fp16 pair, two fp16 MMAs per k-step), "fast" (round 1's arithmetic: one fp16 rounding per weight, misses the 1e-3 gradient bar)
    others = []
    if not args.single_mode:
        for other_mode in [m for m in ("split8", "split", "fast") if m != args.precision]:
            model.gsl_precision = other_mode                # the next step re-creates the engine (optimizer state carried over)
            ms_o, _, _ = timed(step_resident, args.steps, args.warmup)
            others.append({"precision": other_mode, "value": round(images / ms_o * 1e3, 1), "unit": "images/s", "ms_per_step": round(ms_o, 3)})
        model.gsl_precision = args.precision
    other = next((o for o in others if o["precision"] == "fast"), others[0] if others else None)
    result = None
    if rank == 0:
        peaks = load_peaks()
        step_tflops = value * wl["flops"] / 1e12 / world
        roof = kernel_roofline(dev, cfg, peaks, BATCH, wl["dropout"], mode=args.precision)
        if not args.single_mode:
            roof["other_modes"] = []
            for other_mode in [m for m in ("split8", "split", "fast") if m != args.precision]:
                o = kernel_roofline(dev, cfg, peaks, BATCH, wl["dropout"], mode=other_mode)
                roof["other_modes"].append({k: o[k] for k in ("kernel", "achieved", "frac", "ms_per_launch_pair", "executed_frac")})
            roof["other_mode"] = roof["other_modes"][-1]        # ("fast", kept under its round-1 key)
        # executed FLOPs: the last block's out-proj / FFN / attention run on the B cls rows only (exact dead-code elimination, gsl_engine.cu
        # forward / backward), so the tensor cores execute less than the algorithmic count the reference's autograd would
        exe = executed_flops_per_image(cfg)
        exe_tflops = value * exe / 1e12 / world
        roof["step"] = dict(achieved=round(step_tflops, 1), peak=peaks["sustained"], unit="TFLOP/s", frac=round(step_tflops / peaks["sustained"], 4),
                            executed=round(exe_tflops, 1), executed_frac=round(exe_tflops / peaks["sustained"], 4),
                            executed_gflop_per_image=round(exe / 1e9, 3),
                            note=f"whole step: `achieved` = algorithmic FLOPs per image {wl['flops'] / 1e9:.3f} G (BASELINE.md section 3), `executed` = the dense "
                                 "FLOPs the engine really issues (last block on cls rows only; split mode's second MMA per k-step NOT counted), both vs "
                                 "sustained bf16 peak (" + peaks["source"] + ")")
        h2d = sum(t.numel() * t.element_size() for t in host)
        result = {
            "metric": wl["metric"], "value": round(value, 1), "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
            "data": "synthetic",
            "config": {"workload": args.workload, "model": wl["model"],
                       "per_gpu_batch": f"{BATCH} remain + {BATCH} forget", "global_batch": images, "parallelism": f"dp{world}",
                       "precision": args.precision, "alpha": alpha,
                       "arithmetic": ARITHMETIC[args.precision],
                       "dropout": wl["dropout"],
                       "cache": f"inputs_larger_than_l2 ({h2d / 1e6:.0f} MB images + multi-GB activations per step vs 126 MB L2)",
                       "loss": out["total"]},
            "clocks": sampler.summary(),
            "e2e": {"value": round(e2e_value, 1), "unit": "images/s", "ms_per_step": round(ms_e2e, 3), "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 9 * 4,
                    "input_format": "fp32 NCHW in pinned host memory (the reference loader's transforms.ToTensor() output)",
                    "uint8_pipeline": None if args.no_u8_leg else {
                        "value": round(images / ms_e2e_u8 * 1e3, 1), "unit": "images/s", "ms_per_step": round(ms_e2e_u8, 3),
                        "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in host_u8),
                        "note": "same step from raw uint8 pixels (transforms.PILToTensor()); /255 runs in the patchify kernel"}},
            "gpu_launches": int(launches),
            "roofline": roof,
            "other_precision_mode": other,
            "other_precision_modes": others,
            "gpu_reference": gpu_ref,
        }
        if gpu_ref is not None:
            gpu_ref["ratio"] = round(gpu_ref["ms_per_step"] / ms, 2)
            gpu_ref["ratio_e2e"] = round(gpu_ref["ms_per_step"] / ms_e2e, 2)
        if world == 1 and not args.no_cpu_baseline and args.workload == "p8s8_bs512":
            result["cpu_baseline"] = cpu_baseline(sample_batch=96, steps=3)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(result), flush=True)


def executed_flops_per_image(cfg):
    """Dense FLOPs (2mnk) per image the engine issues: as BASELINE.md section 3 / oracle.flops_per_image, minus what the last block skips --
    its attention is one query per (image, head), its out-proj / FFN and their backward run on the single cls row (gsl_engine.cu)."""
    N, D, H, r, L = cfg.tokens, cfg.dim, cfg.mlp_dim, cfg.lora_rank, cfg.depth
    inner = cfg.heads * 64
    qkv, qk, out, fc, lora = 2 * N * D * 3 * inner, 2 * N * N * inner, 2 * N * inner * D, 2 * N * D * H, 2 * N * r * (D + H)
    patch = 2 * (N - 1) * (3 * cfg.patch_size ** 2) * D
    tok = 1.0 / N                                                   # share of the cls row
    fwd = patch + (L - 1) * (qkv + 2 * qk + out + 2 * fc + 2 * lora) + (qkv + tok * (2 * qk + out + 2 * fc + 2 * lora))
    ffn_bwd, attn_bwd = 2 * fc + 4 * lora, qkv + out + 4 * qk
    dense_blocks = max(L - 2, 0)                                    # blocks 1 .. L-2: full FFN + attention backward; block 0: FFN only, no fc1 dX
    bwd = dense_blocks * (ffn_bwd + attn_bwd) + (ffn_bwd - fc if L > 1 else 0) + tok * ffn_bwd + (qkv + tok * (out + 4 * qk) if L > 1 else 0)
    return float(fwd + bwd)


def gpu_reference(dev, B):
    """BASELINE.md section 4 / north_star denominator: the reference step (oracle restatement of engine_cl.py:59-125: two forwards, CE + bounded
    forget loss + structure loss, autograd backward, AdamW) as stock PyTorch FP32 eager, TF32 off, on THIS GPU at the headline batch, timed with
    CUDA events outside the repo arm's timed region."""
    from oracle import vit_oracle as O
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = O.P8S8
    sd = {k: v.to(dev) for k, v in O.init_state_dict(cfg, seed=1337).items()}
    g = torch.Generator().manual_seed(3)
    xr, xf = torch.rand(B, 3, 112, 112, generator=g).to(dev), torch.rand(B, 3, 112, 112, generator=g).to(dev)
    yr, yf = torch.randint(0, 100, (B,), generator=g).to(dev), torch.randint(0, 100, (B,), generator=g).to(dev)
    state = {}

    def step():
        O.unlearn_step(sd, cfg, state, xr, yr, xf, yf, lr=HP["lr"], wd=HP["wd"], beta=HP["beta"], alpha=HP["alpha"], BND=HP["BND"])
    try:
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        out = dict(ms_per_step=round(ms, 2), images_per_s=round(2 * B / ms * 1e3, 1), batch=f"{B}+{B}",
                   what="reference step, PyTorch FP32 eager (TF32 off, dropout 0) on the same B200: oracle/vit_oracle.py unlearn_step, 2 warm-up + 5 timed")
    except torch.cuda.OutOfMemoryError as e:       # never take the box down for a side measurement
        out = dict(error=f"out of memory: {str(e)[:120]}")
    del sd, state, xr, xf
    torch.cuda.empty_cache()
    return out


ARITHMETIC = {
    "split8": "fp16 activations x (fp16 hi + e4m3 lo) frozen weights [the lo term on the FP8 tensor path against an in-kernel e5m2 copy of the "
              "activations], (fp16 hi + fp16 lo) LoRA factors, fp32 accumulate / residual stream / loss: logits and every LoRA gradient within "
              "1e-3 of FP32 (tests/test_engine_gpu.py)",
    "split": "fp16 activations x (fp16 hi + fp16 lo) frozen weights and LoRA factors, fp32 accumulate / residual stream / loss: "
             "logits and every LoRA gradient within 1e-3 of FP32 (tests/test_engine_gpu.py)",
    "fast": "fp16 operands (one rounding per weight), fp32 accumulate / residual stream / loss: LoRA gradients 1-2.4e-3 of FP32",
}


def kernel_roofline(dev, cfg, peaks, BATCH=BATCH, DROPOUT=DROPOUT, mode="split"):
    """Fused FFN+LoRA GEMM pair at the step's own shape (M = 1024 * 197 rows), timed live with CUDA events on the launching stream:
      fc1: x W1'^T + b1 -> G = Dropout(gelu(h)) and mask * gelu'(h)        (W' = W + s B A: the LoRA branch of loralib.Linear folded in)
      fc2: G W2'^T + b2, Dropout, + residual x
    Algorithmic FLOPs per launch pair: 2 * M * (D*H*2 + 2r(D+H)) (SURVEY 8d -- what the reference's two lora.Linear layers compute);
    peak = measured cuBLAS bf16 burst."""
    from gslora import _ffi as F
    M, D, H, r = 2 * BATCH * cfg.tokens, cfg.dim, cfg.mlp_dim, cfg.lora_rank
    x = (torch.randn(M, D, device=dev) * 0.5).half()
    w1f, w2f = torch.randn(H, D, device=dev) * 0.05, torch.randn(D, H, device=dev) * 0.02
    split = mode == "split"
    w1, w2 = w1f.half(), w2f.half()
    w1lo = (w1f - w1.float()).half() if split else None      # precision mode "split": second term of each weight (gsl_gemm_f16_split)
    w2lo = (w2f - w2.float()).half() if split else None
    kw1, kw2 = dict(B_lo=w1lo), dict(B_lo=w2lo)
    if mode == "split8":                                     # fp16(W 2^12) + e4m3 residual (gsl_gemm_f16_split8)
        def pair8(w):
            hi = torch.empty_like(w, dtype=torch.half); lo8 = torch.empty_like(w, dtype=torch.uint8)
            F.check(F.lib().gsl_cast_f32_to_f16_split8(F.ptr(w), w.shape[1], F.ptr(hi), F.ptr(lo8), w.shape[1], w.shape[0], w.shape[1], 12, 0, F.cur_stream()))
            return hi, lo8
        w2, l2 = pair8(w2f)                                  # as in the engine: fc1 (epilogue-bound) keeps the fp16 residual, fc2 takes the e4m3 one
        w1lo = (w1f - w1.float()).half()
        kw1, kw2 = dict(B_lo=w1lo), dict(B_lo8=l2, lo8_shift=12)
    b1, b2 = torch.randn(H, device=dev), torch.randn(D, device=dev)
    gp = torch.empty(M, H, device=dev, dtype=torch.half)
    g = torch.empty(M, H, device=dev, dtype=torch.half)
    res = torch.randn(M, D, device=dev)
    y = torch.empty(M, D, device=dev)

    def pair():
        F.gemm_f16(x, w1, epi=F.EPI_GELU, bias=b1, out0=gp, out1=g, drop_p=DROPOUT, drop_seed=17, **kw1)
        F.gemm_f16(g, w2, epi=F.EPI_RES_F32, bias=b2, out0=y, aux=res, drop_p=DROPOUT, drop_seed=18, **kw2)
    for _ in range(3):
        pair()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        pair()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 2.0 * M * (D * H * 2 + 2 * r * (D + H))
    ach = flops / ms / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")      # dram bytes of the same two launches from the committed ncu --set full capture
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        tj = tj.get(mode, tj.get("split", tj) if mode == "split8" else tj)
        traffic, traffic_src = tj.get("ffn_pair_bytes"), tj.get("source")
    # fc1 reads x, writes G and mask*gelu'(h) (fp16); fc2 reads G and the fp32 residual, writes the fp32 stream; weights 4 x D x H fp16
    alg_bytes = 2 * M * D + 2 * 2 * M * H + 2 * M * H + 2 * 4 * M * D + 4 * D * H
    if (M, D, H) != (2 * 512 * 197, 512, 2048):
        traffic = traffic_src = None                                   # the committed ncu capture is of the P8S8 shape
    # tensor-pipe time really issued, in fp16-MMA units: split = two fp16 MMAs per k-step, split8 = one fp16 MMA + one FP8 MMA at twice the rate
    mma = 2.0 * M * D * H * 2 * {"fast": 1.0, "split": 2.0, "split8": 1.75}[mode]      # (split8: fc1 at 2.0, fc2 at 1.5)
    kname = {"fast": "gemm_tcgen05_kernel<2,256,EPI_GELU,0,2> + <2,256,EPI_RES_F32,0,2>",
             "split": "gemm_tcgen05_kernel<2,256,EPI_GELU,SPLIT,2> + <2,256,EPI_RES_F32,SPLIT,2>",
             "split8": "gemm_tcgen05_kernel<2,256,EPI_GELU,SPLIT,2> [fc1 keeps the fp16 residual] + <2,256,EPI_RES_F32,SPLIT8,1> [e4m3 residual, one epilogue group]"}[mode]
    return dict(bound="tensor", kernel=f"{kname} (FFN+LoRA pair, precision {mode}, dropout {DROPOUT})",
                achieved=round(ach, 1), peak=peaks["burst"], unit="TFLOP/s", frac=round(ach / peaks["burst"], 4),
                executed_frac=round(mma / ms / 1e9 / peaks["burst"], 4), traffic=traffic,
                traffic_source=traffic_src, algorithmic_bytes=int(alg_bytes), ms_per_launch_pair=round(ms, 4),
                peak_source=peaks["source"] + " cuBLAS bf16 burst")


def cpu_baseline(sample_batch=16, steps=2, warm_batch=None):
    """The reference's CPU path (oracle port: oracle/vit_oracle.py restates vit_face.py / engine_cl.py:59-125; the reference tree itself
    does not travel to the GPU box) on all host cores, bounded sample: bs `sample_batch`+`sample_batch`, FP32."""
    from oracle import vit_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.P8S8
    sd = O.init_state_dict(cfg, seed=1337)
    g = torch.Generator().manual_seed(3)
    B = sample_batch
    xr, xf = torch.rand(B, 3, 112, 112, generator=g), torch.rand(B, 3, 112, 112, generator=g)
    yr, yf = torch.randint(0, 100, (B,), generator=g), torch.randint(0, 100, (B,), generator=g)
    state = {}
    w = warm_batch or B         # warm-up (thread pool, allocator): a small batch is enough when one full step costs tens of seconds
    O.unlearn_step(sd, cfg, state, xr[:w], yr[:w], xf[:w], yf[:w], lr=HP["lr"], wd=HP["wd"], beta=HP["beta"], alpha=HP["alpha"], BND=HP["BND"])
    t0 = time.perf_counter()
    for _ in range(steps):
        O.unlearn_step(sd, cfg, state, xr, yr, xf, yf, lr=HP["lr"], wd=HP["wd"], beta=HP["beta"], alpha=HP["alpha"], BND=HP["BND"])
    dt = (time.perf_counter() - t0) / steps
    return dict(value=round(2 * B / dt, 2), unit="images/s", cores=cores, kind="port",
                sample=f"oracle port of the reference step (PyTorch FP32 eager, dropout 0), bs {B}+{B}, 1 warm-up (bs {w}+{w}) + {steps} timed steps, {dt:.2f} s/step")


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port of engine_cl.py:59-125 over vit_face.py; the reference
    tree itself does not travel to the GPU box), all host cores, FP32, at the HEADLINE batch 512+512 -- the same config as the repo arm --
    when host memory allows (autograd keeps ~80 MB of fp32 activations per image), else the largest halving that fits.  One step is tens of
    seconds on CPU, so the run is 1 small warm-up step + at most 2 timed steps whatever --steps says."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.ref_batch
    try:
        import psutil
        avail = psutil.virtual_memory().available
        while B > 16 and 2 * B * 80e6 * 1.3 > avail:
            B //= 2
    except Exception:
        pass
    steps = max(1, min(args.steps, 2))
    cb = cpu_baseline(sample_batch=B, steps=steps, warm_batch=8)
    same = B == BATCH
    out = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "images/s", "n_gpus": args.gpus, "steps": steps,
           "warmup": 1, "ms_per_step": round(2 * B / cb["value"] * 1e3, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": "p8s8_bs512", "per_gpu_batch": f"{B} remain + {B} forget", "global_batch": 2 * B,
                      "sample": ("the full 512+512 step" if same else f"bs {B}+{B} per step (host memory bounds the sample of the 512+512 workload)"),
                      "device": "host CPU", "dropout": 0.0},
           "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gslora", choices=["gslora", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="p8s8_bs512", choices=sorted(WORKLOADS))
    ap.add_argument("--no-u8-leg", action="store_true")
    ap.add_argument("--precision", default="split8", choices=["split8", "split", "fast"],
                    help="split (default): fp16 hi+lo weights, gradients within the 1e-3 parity bar; fast: one fp16 rounding per weight")
    ap.add_argument("--alpha", type=float, default=None, help="group-Lasso weight of the step (default 1e-4; BASELINE config 5 sweeps 0 / 1e-4 / 1e-3 / 1e-2)")
    ap.add_argument("--single-mode", action="store_true", help="skip the resident leg of the other precision mode")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip the FP32-eager reference step on the GPU (north-star denominator)")
    ap.add_argument("--ref-batch", type=int, default=BATCH, help="--impl reference: images per stream of the CPU step (default: the headline 512)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "gslora":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gslora(args)


if __name__ == "__main__":
    main()
