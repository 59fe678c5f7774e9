"""gslora-b200 `engine_cl` -- drop-in for the reference's continual-forgetting inner loop (engine_cl.py of
bjzhb666/GS-LoRA): `train_one_epoch` (engine_cl.py:12-244), `evaluate` (:247), `eval_data` (:318),
`get_structure_loss` (:349) and `get_prototype_loss` (:571) with the reference's signatures and return values.

One unlearning step (engine_cl.py:59-125) is executed as ONE fused pass of the native engine:
  remain and forget batches are concatenated -> one forward -> device-side CE sums and the bounded-forget gate
  relu(BND - CE_f) -> one selective backward (LoRA gradients only) -> [flat NCCL allreduce of the LoRA gradient
  buffer under torch.distributed] -> fused group-Lasso + AdamW kernel -> ONE device-to-host copy of the scalars the
  reference fetches with >= 8 `.item()` calls.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch
import torch.nn.functional as F_t

from gslora import _ffi as F

try:  # logging is host-side orchestration; keep the reference's wandb calls when wandb is importable
    import wandb
except Exception:  # pragma: no cover
    class _NoWandb:
        @staticmethod
        def log(*a, **k):
            return None
    wandb = _NoWandb()

try:
    from util.utils import AverageMeter, get_time  # the reference's own helpers when its tree is on sys.path
except Exception:
    import datetime

    class AverageMeter:
        def __init__(self):
            self.reset()

        def reset(self):
            self.val = self.avg = self.sum = self.count = 0

        def update(self, val, n=1):
            self.val = val
            self.sum += val * n
            self.count += n
            self.avg = self.sum / self.count

    def get_time():
        return (str(datetime.datetime.now())[:-10]).replace(" ", "-").replace(":", "-")


def _wandb_log(d):
    """The reference logs unconditionally (its driver has called wandb.init); stay silent when no run is active."""
    if getattr(wandb, "run", None) is not None or not hasattr(wandb, "run"):
        wandb.log(d)


def _unwrap(model):
    return model.module if isinstance(model, (torch.nn.DataParallel, torch.nn.parallel.DistributedDataParallel)) else model


def _dist():
    """torch.distributed when this process is one of several data-parallel ranks, else None.  The reference drivers know nothing about
    torch.distributed (they use nn.DataParallel, train_own_forget_cl.py:494-497); launched unchanged under `torchrun` (RANK / WORLD_SIZE /
    MASTER_* in the environment) the process group is created here on first use -- NCCL on GPUs, gloo otherwise.  GSLORA_AUTO_DIST=0 opts out."""
    import torch.distributed as dist
    if not dist.is_available():
        return None
    if not dist.is_initialized():
        if int(os.environ.get("WORLD_SIZE", "1")) <= 1 or "RANK" not in os.environ or os.environ.get("GSLORA_AUTO_DIST", "1") == "0":
            return None
        backend = os.environ.get("GSLORA_DIST_BACKEND", "nccl" if torch.cuda.is_available() else "gloo")
        if torch.cuda.is_available():
            # one rank per GPU: without CUDA_VISIBLE_DEVICES=$LOCAL_RANK every rank would otherwise land on cuda:0
            local = int(os.environ.get("LOCAL_RANK", "0"))
            if torch.cuda.device_count() > 1:
                torch.cuda.set_device(local % torch.cuda.device_count())
        if backend == "nccl":
            dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
        else:       # gloo also moves CUDA tensors (through the host): lets several ranks share one GPU in tests
            dist.init_process_group(backend)
    return dist if dist.get_world_size() > 1 else None


def shard_batch(x, y):
    """Data parallel under torchrun: every rank builds the same seeded loaders (the driver is unchanged), so each step's
    GLOBAL batch is identical on all ranks; rank r keeps the samples r, r + world, ... and the step's all-reduces (loss sums / counts, flat LoRA
    gradient) put the global batch back together.  No process group (or world size 1): the batch is returned as is."""
    d = _dist()
    if d is None or x is None:
        return x, y
    r, w = d.get_rank(), d.get_world_size()
    return x[r::w], y[r::w]


def _adamw_hparams(optimizer, params):
    ids = {id(p) for p in params}
    for g in optimizer.param_groups:
        if any(id(p) in ids for p in g["params"]):
            if not isinstance(optimizer, (torch.optim.AdamW,)):
                raise NotImplementedError("gslora-b200 fused step implements torch.optim.AdamW (what timm create_optimizer builds for the reference)")
            return dict(lr=g["lr"], wd=g["weight_decay"], betas=tuple(g["betas"]), eps=g["eps"])
    raise RuntimeError("optimizer does not hold the model's LoRA parameters")


def _prototype_tensor(prototype_dict, num_class, dim, device):
    """Dense [num_class, dim] table (+ a per-class `present` mask, None for a caller-supplied tensor) of the driver's {label: CPU tensor} dict
    (util/utils.py:546-549), built with ONE stacked host-to-device copy."""
    if torch.is_tensor(prototype_dict):
        return prototype_dict.to(device), None
    keys = sorted(int(k) for k in prototype_dict)
    if keys and (keys[0] < 0 or keys[-1] >= num_class):
        raise KeyError(f"prototype label {keys[0] if keys[0] < 0 else keys[-1]} outside [0, {num_class})")
    t = torch.zeros(num_class, dim)
    present = torch.zeros(num_class, dtype=torch.bool)
    if keys:
        idx = torch.tensor(keys, dtype=torch.long)
        t[idx] = torch.stack([torch.as_tensor(prototype_dict[k]).detach().float().cpu().reshape(dim) for k in keys])
        present[idx] = True
    return t.to(device), present.to(device)


def _cached_prototype_table(m, prototype_dict, num_class, dim, device):
    """The table is rebuilt only when the driver hands over a different dict object (it computes the prototypes once per task,
    train_own_forget_cl.py:1026-1040): the step itself issues no host-to-device copy for it."""
    key = (id(prototype_dict), len(prototype_dict) if hasattr(prototype_dict, "__len__") else -1, str(device))
    cache = getattr(m, "_gsl_proto_cache", None)
    if cache is None or cache[0] != key:
        table, present = _prototype_tensor(prototype_dict, num_class, dim, device)
        cache = (key, table.float().contiguous(), present, prototype_dict)      # the dict is kept alive so its id cannot be recycled
        m._gsl_proto_cache = cache
    return cache[1], cache[2]


def get_prototype_loss(output, labels, prototype_dict, distance="kl"):
    """engine_cl.get_prototype_loss (engine_cl.py:571-603) without the per-sample `.item()` loop: the prototypes are gathered on
    device.  Differentiable torch expression for callers that run their own autograd loop; the fused step (unlearn_step) uses the
    gsl_prototype_kl_fwd / _grad kernels instead."""
    if torch.is_tensor(prototype_dict):
        pt = prototype_dict[labels.long()].to(output.device)
    else:
        table, present = _prototype_tensor(prototype_dict, max(int(k) for k in prototype_dict) + 1, output.shape[1], output.device)
        lab = labels.long()
        if bool((lab >= table.shape[0]).any()) or not bool(present[lab.clamp(max=table.shape[0] - 1)].all()):
            raise KeyError("get_prototype_loss: a label of the batch has no prototype")      # the reference's dict lookup raises too
        pt = table[lab]
    if distance == "l2":
        return torch.mean((output - pt) ** 2)
    return F_t.kl_div(F_t.log_softmax(output, dim=1), F_t.log_softmax(pt, dim=1), reduction="batchmean", log_target=True)


class _StructureLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, group_type, *lora_params):
        eng = model._engine
        offs = eng.group_offsets_by_type[group_type]
        G = offs.numel() - 1
        norms = torch.empty(G, dtype=torch.float32, device=eng.device)
        F.check(F.lib().gsl_tensor_norms(F.ptr(eng.lora_flat), F.ptr(offs), G, 0, F.ptr(norms), F.cur_stream()), "gsl_tensor_norms")
        ctx.model, ctx.norms, ctx.offs = model, norms, offs
        return norms.sum()

    @staticmethod
    def backward(ctx, g):
        eng = ctx.model._engine
        inv = torch.where(ctx.norms > 0, 1.0 / ctx.norms, torch.zeros_like(ctx.norms))
        sizes = (ctx.offs[1:] - ctx.offs[:-1]).long()
        flat = eng.lora_flat * torch.repeat_interleave(inv, sizes) * g
        out = [eng.lora_view(flat, l, w) for l in range(eng.spec.depth) for w in range(eng.spec.tensors_per_block)]
        return (None, None, *out)


def get_structure_loss(model: torch.nn.Module, imagenet=False, group_type: str = "block"):
    """engine_cl.get_structure_loss (engine_cl.py:349-432): sum over Transformer blocks of the L2 norm of the block's four
    LoRA matrices.  Differentiable w.r.t. the LoRA parameters (gradient P / ||g||, 0 at ||g|| = 0 where the reference NaNs).
    `group_type` adds the groupings of engine.get_structure_loss (engine.py:532-687, group_pos "FFN"): "lora" (one (A, B) pair per
    group) and "matrix" (every LoRA matrix on its own)."""
    m = _unwrap(model)
    # `imagenet` only selects parameter NAMES in the reference (encoder.layers.encoder_layer_{i}.mlp.{0,3}.lora_{A,B}, 12 groups hard-coded,
    # engine_cl.py:395-402); here the groups are the engine's per-block LoRA slices of whichever engine-backed model is passed.
    m.ensure_engine(1)
    m.sync_engine()
    return _StructureLossFn.apply(m, group_type, *m.lora_parameters())


class StepResult:
    """Scalars of one unlearning step (what the reference reads with >= 8 `.item()` calls, engine_cl.py:68-121).  The device-to-host
    copy is queued on the step's stream into pinned memory; nothing blocks until a value is read (`result["total"]`, `.wait()`), so
    the host can queue the next step while this one runs."""

    _KEYS = ("loss_remain", "ce_forget", "loss_forget", "structure", "top1_remain", "top1_forget", "proto_forget", "proto_remain", "total")

    def __init__(self, pinned, event, n_vals, consts):
        self._pinned, self._event, self._n, self._c = pinned, event, n_vals, consts
        self._vals = None

    def wait(self) -> Dict[str, float]:
        if self._vals is None:
            self._event.synchronize()
            host = self._pinned[:self._n].tolist()
            c = self._c
            s_ce_r, n_r, s_ce_f, n_f, hit_r, hit_f, s_kl_r, s_kl_f, structure = host[:9]
            if structure != structure:       # NaN: gsl_grouplasso_adamw_step found inf / NaN in a group's gradient and skipped that group
                raise FloatingPointError("unlearn_step: non-finite LoRA gradient; the affected groups were not updated.  Either the loss-scaled fp16 "
                                         "gradient stream overflowed -- lower GSLORA_GRAD_SCALE (default 1024) -- or, in precision mode split8, a frozen "
                                         "weight reaches |W| >= 16 and its 2^12-scaled fp16 operand overflowed -- use GSLORA_PRECISION=split")
            if self._n > 9 and host[9] != 0.0:
                raise KeyError("unlearn_step: a label of this step's batch has no entry in prototype_dict (engine_cl.py:571-603 looks every "
                               "label up in the dict)")
            loss_remain = s_ce_r / max(n_r, 1.0)
            ce_forget = s_ce_f / max(n_f, 1.0)
            loss_forget = max(c["BND"] - ce_forget, 0.0)
            pf_v, pr_v = s_kl_f / max(n_f, 1.0), s_kl_r / max(n_r, 1.0)
            proto_total = c["pwf"] * max(c["BND_pro"] - pf_v, 0.0) + c["pwr"] * pr_v if c["use_prototype"] else 0.0
            self._vals = dict(loss_remain=loss_remain, ce_forget=ce_forget, loss_forget=loss_forget, structure=structure,
                              top1_remain=100.0 * hit_r / max(n_r, 1.0), top1_forget=100.0 * hit_f / max(n_f, 1.0),
                              proto_forget=pf_v, proto_remain=pr_v,
                              total=c["beta"] * loss_forget + loss_remain + c["alpha"] * structure + proto_total)
            self._pinned = None
        return self._vals

    def __getitem__(self, k):
        return self.wait()[k]

    def keys(self):
        return self._KEYS


class _PinnedRing:
    """Small ring of pinned host buffers for the per-step scalar read-back (cudaHostAlloc per step would serialise the host)."""

    def __init__(self, n=8, width=16):
        self.bufs = [torch.empty(width, dtype=torch.float32).pin_memory() for _ in range(n)]
        self.pending = [None] * n
        self.i = 0

    def take(self):
        i = self.i
        self.i = (i + 1) % len(self.bufs)
        if self.pending[i] is not None:
            self.pending[i].wait()          # an unread result still owns this buffer: materialise it first
        return i, self.bufs[i]


_RING = None

class _GraphedStep:
    """One unlearning step (forward, loss sums, CE gradient, selective backward, fused group-lasso AdamW, LoRA repack) captured as a CUDA graph for
    a fixed (engine, batch split, input dtype, beta / alpha / BND, AdamW constants, grouping, dropout on / off) key and replayed with ONE launch:
    the ~150 kernel launches of a step otherwise leave ~1 ms of gaps between kernels (profiles/r02h_gaps.log).  What changes from replay to replay
    -- the dropout base seed, the AdamW step count and the learning rate -- lives in a 16-byte device "step state" the kernels read
    (include/gslora.h gsl_engine_forward_dev / gsl_grouplasso_adamw_step_dev); inputs are copied into the graph's static buffers (the copy that
    torch.cat made before).  Single-process steps without the prototype term only; everything else takes the eager path."""

    def __init__(self, m, eng, xr, xf, Br, Bf, beta, alpha, BND, hp, group_type, dropout_on):
        import struct
        self._struct = struct
        dev = xr.device
        B = Br + Bf
        self.eng, self.Br, self.B, self.dropout_on = eng, Br, B, dropout_on
        self.img = torch.empty((B,) + tuple(xr.shape[1:]), dtype=xr.dtype if xr.dtype == torch.uint8 else torch.float32, device=dev)
        self.lab = torch.empty(B, dtype=torch.int64, device=dev)
        self.dlogits = torch.empty(B, eng.spec.num_class, dtype=torch.float32, device=dev)
        self.state_dev = torch.zeros(16, dtype=torch.uint8, device=dev)
        self.state_host = [torch.zeros(16, dtype=torch.uint8).pin_memory() for _ in range(8)]
        self.state_i = 0
        self.packed = torch.zeros(9, dtype=torch.float32, device=dev)
        self.slot = 0
        kw = m.image_kwargs(self.img)
        if kw.get("pixel_norm") is not None:
            raise RuntimeError("graph capture: Normalize-in-kernel inputs take the eager path")
        L = F.lib()
        step0 = eng.opt_step
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        n0 = L.gsl_launch_count()
        with torch.cuda.graph(self.graph):
            eng.forward(self.img, self.lab, self.slot, use_lora=True, dropout_seed=1 if dropout_on else 0, seed_dev=self.state_dev,
                        channels_last=kw.get("channels_last", False))
            sums = eng.loss_sums(self.slot, Br, B, None)
            eng.unlearn_ce_grad(self.slot, self.lab, Br, B, beta, BND, self.dlogits)
            eng.backward(self.slot, self.dlogits, None, accumulate=False)
            eng.optimizer_step(lr=0.0, wd=hp["wd"], alpha=alpha, betas=hp.get("betas", (0.9, 0.999)), eps=hp.get("eps", 1e-8), group_type=group_type,
                               state_dev=self.state_dev)
            self.packed.copy_(torch.cat([sums, eng.group_norms[:eng.num_groups].sum().view(1)]))
        self.launches = int(L.gsl_launch_count() - n0)
        eng.opt_step = step0                      # capturing executed nothing on the device

    def run(self, m, xr, yr, xf, yf, seed, lr):
        Br = self.Br
        self.img[:Br].copy_(xr, non_blocking=True)
        self.img[Br:].copy_(xf, non_blocking=True)
        self.lab[:Br].copy_(yr, non_blocking=True)
        self.lab[Br:].copy_(yf, non_blocking=True)
        eng = self.eng
        eng.opt_step += 1
        host = self.state_host[self.state_i]
        self.state_i = (self.state_i + 1) % len(self.state_host)
        host.copy_(torch.frombuffer(bytearray(self._struct.pack("<QIf", seed & 0xFFFFFFFFFFFFFFFF, eng.opt_step, float(lr))), dtype=torch.uint8))
        self.state_dev.copy_(host, non_blocking=True)
        m._slot_stamp[self.slot] += 1             # the graph's slot now holds THIS step's activations
        self.graph.replay()
        F.lib().gsl_count_launches(self.launches)
        return self.packed


_GRAPH_AFTER = 2        # eager steps with an unchanged key before the step is captured


def _graph_step(m, eng, xr, yr, xf, yf, beta, alpha, BND, hp, group_type, seed):
    """The captured step for this call's key (or None: not yet / not eligible)."""
    if os.environ.get("GSLORA_CUDA_GRAPH", "1") == "0":
        return None
    key = (getattr(eng, "serial", id(eng)), tuple(xr.shape), tuple(xf.shape), xr.dtype, xf.dtype, float(beta), float(alpha), float(BND), float(hp["wd"]),
           tuple(hp.get("betas", (0.9, 0.999))), float(hp.get("eps", 1e-8)), group_type, seed != 0, m.input_pixel_norm is None)
    st = m.__dict__.setdefault("_gsl_graph", dict(key=None, seen=0, step=None, failed=False))
    if st["key"] != key:
        st.update(key=key, seen=0, step=None)
    if st["failed"] or m.input_pixel_norm is not None and xr.dtype == torch.uint8:
        return None
    if st["step"] is None:
        st["seen"] += 1
        if st["seen"] <= _GRAPH_AFTER:
            return None
        try:
            st["step"] = _GraphedStep(m, eng, xr, xf, int(xr.shape[0]), int(xf.shape[0]), beta, alpha, BND, hp, group_type, seed != 0)
        except Exception as e:     # capture is an optimisation: fall back to eager launches, loudly, once
            st["failed"] = True
            import warnings
            warnings.warn(f"gslora-b200: CUDA-graph capture of the unlearning step failed ({e!r}); continuing with eager launches")
            return None
    return st["step"]



def unlearn_step_async(model, inputs_remain, labels_remain, inputs_forget, labels_forget, *, beta: float, alpha: float, BND: float,
                       optimizer=None, hparams: Optional[dict] = None, use_prototype: bool = False, prototype_dict=None,
                       prototype_weight_forget: float = 0.0, prototype_weight_remain: float = 0.0, BND_pro: float = 0.0,
                       dropout_seed: Optional[int] = None, group_type: str = "block") -> StepResult:
    """One step of engine_cl.train_one_epoch (engine_cl.py:59-125), fused; returns a StepResult whose scalars arrive asynchronously."""
    global _RING
    m = _unwrap(model)
    Br, Bf = int(inputs_remain.shape[0]), int(inputs_forget.shape[0])
    B = Br + Bf
    dist = _dist()
    if B == 0 and dist is None:
        raise ValueError("unlearn_step: empty batch")
    dev = inputs_remain.device
    eng = m.ensure_engine(max(B, 1))
    m.sync_engine()
    if m._merged():
        raise RuntimeError("unlearn_step needs model.train() (un-merged LoRA)")
    missing = None
    hp = hparams if hparams is not None else _adamw_hparams(optimizer, m.lora_parameters())
    seed = m.dropout_seed() if dropout_seed is None else int(dropout_seed)
    graphed = None
    if B > 0 and Br > 0 and Bf > 0 and dist is None and not use_prototype and inputs_remain.dtype == inputs_forget.dtype:
        graphed = _graph_step(m, eng, inputs_remain, labels_remain, inputs_forget, labels_forget, beta, alpha, BND, hp, group_type, seed)
    if graphed is not None:
        packed = graphed.run(m, inputs_remain, labels_remain, inputs_forget, labels_forget, seed, hp["lr"])
        m.mark_lora_updated_by_engine()
        return _queue_readback(packed, dict(beta=beta, alpha=alpha, BND=BND, BND_pro=BND_pro, pwf=prototype_weight_forget,
                                            pwr=prototype_weight_remain, use_prototype=use_prototype))
    if B > 0:
        if inputs_remain.dtype == torch.uint8 or inputs_forget.dtype == torch.uint8:     # raw pixels: ToTensor [+ Normalize] runs in the patchify kernel
            if inputs_remain.dtype != inputs_forget.dtype:
                raise TypeError("unlearn_step: remain and forget images must both be uint8 (raw pixels) or both floating point (ToTensor output)")
            img = torch.cat([inputs_remain, inputs_forget], dim=0).contiguous()
        else:
            img = torch.cat([inputs_remain.float(), inputs_forget.float()], dim=0).contiguous()
        lab = torch.cat([labels_remain.to(torch.int64), labels_forget.to(torch.int64)], dim=0).contiguous()
        slot = m._take_slot()
        eng.forward(img, lab, slot, use_lora=True, dropout_seed=seed, **m.image_kwargs(img))
        table = kl = None
        if use_prototype:                                       # GS-LoRA++ (engine_cl.py:97-101): per-sample KL to the class prototype, on device
            table, present = _cached_prototype_table(m, prototype_dict, eng.spec.num_class, eng.spec.dim, dev)
            lab_kl = lab.clamp(0, eng.spec.num_class - 1)       # the KL kernels index the table by label: never out of bounds
            if present is not None:                             # a label without a prototype is a KeyError in the reference: flagged on device,
                missing = ((lab != lab_kl) | ~present[lab_kl]).any().float().view(1)                 # raised when the step's scalars are read
            kl = eng.prototype_kl(slot, lab_kl, table, B)
        sums = eng.loss_sums(slot, Br, B, kl)
    else:
        # this rank's share of the global batch is empty (drop_last=False tails, few-shot forget sets smaller than the world size): it still
        # joins both collectives -- with zero sums and a zero gradient -- and applies the same optimizer step as every other rank
        sums = eng.sums.zero_()
    if dist is not None:
        dist.all_reduce(sums)                                   # global CE / KL sums, counts, hits (8 floats)
    if B > 0:
        dlogits = torch.empty(B, eng.spec.num_class, dtype=torch.float32, device=dev)
        eng.unlearn_ce_grad(slot, lab, Br, B, beta, BND, dlogits)
        demb = eng.prototype_kl_grad(slot, lab_kl, table, Br, B, prototype_weight_forget, prototype_weight_remain, BND_pro) if use_prototype else None
        eng.backward(slot, dlogits, demb, accumulate=False)
    else:
        eng.grad_flat.zero_()
    if dist is not None:
        dist.all_reduce(eng.grad_flat)                          # the one flat LoRA-gradient allreduce (0.98 MB for ViT-P8S8 r=8)
    eng.optimizer_step(lr=hp["lr"], wd=hp["wd"], alpha=alpha, betas=hp.get("betas", (0.9, 0.999)), eps=hp.get("eps", 1e-8),
                       group_type=group_type)                   # cfg["GROUP_TYPE"] of engine.py:82-90
    m.mark_lora_updated_by_engine()
    # one D2H copy (queued, pinned) for everything the reference reads with .item()
    packed = torch.cat([sums, eng.group_norms[:eng.num_groups].sum().view(1)] + ([missing] if missing is not None else []))
    return _queue_readback(packed, dict(beta=beta, alpha=alpha, BND=BND, BND_pro=BND_pro, pwf=prototype_weight_forget,
                                        pwr=prototype_weight_remain, use_prototype=use_prototype))


def _queue_readback(packed, consts) -> StepResult:
    """one D2H copy (queued, pinned) for everything the reference reads with .item()"""
    global _RING
    if _RING is None:
        _RING = _PinnedRing()
    slot_i, pinned = _RING.take()
    pinned[:packed.numel()].copy_(packed, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    res = StepResult(pinned, ev, packed.numel(), consts)
    _RING.pending[slot_i] = res
    return res


def unlearn_step(model, inputs_remain, labels_remain, inputs_forget, labels_forget, **kw) -> Dict[str, float]:
    """Synchronous form of unlearn_step_async: returns the dict of scalars the reference logs."""
    return unlearn_step_async(model, inputs_remain, labels_remain, inputs_forget, labels_forget, **kw).wait()


def sync_optimizer_state(model, optimizer):
    """Expose the engine's fused AdamW moments through `optimizer.state` (views, no copies) so state_dict()/resume see them."""
    m = _unwrap(model)
    eng = m._engine
    if eng is None or optimizer is None:
        return
    tpb = eng.spec.tensors_per_block
    for l in range(eng.spec.depth):
        for w, p in zip(range(tpb), m.lora_parameters()[tpb * l:tpb * l + tpb]):
            optimizer.state[p] = {"step": torch.tensor(float(eng.opt_step)), "exp_avg": eng.lora_view(eng.exp_avg, l, w),
                                  "exp_avg_sq": eng.lora_view(eng.exp_avg_sq, l, w)}


class _Prefetcher:
    """Side-stream H2D prefetch of the forget loader (the reference's util/data_prefetcher.py:10-58 behaviour).  Under torch.distributed the
    batch is sharded on the HOST first (`shards = True`: what next() returns is already this rank's share), so each rank copies 1/world of the
    global batch over PCIe; `global_n` is the size of the global batch the last next() came from (the meters weigh by it)."""
    shards = True

    def __init__(self, loader, device):
        self.loader, self.device = iter(loader), device
        self.stream = torch.cuda.Stream(device=device)
        self.global_n = self._n = 0
        self._preload()

    def _preload(self):
        try:
            s, t = next(self.loader)
        except StopIteration:
            self.s = self.t = None
            return
        self._n = int(s.shape[0])
        s, t = shard_batch(s, t)
        with torch.cuda.stream(self.stream):
            self.s, self.t = s.to(self.device, non_blocking=True), t.to(self.device, non_blocking=True)

    def next(self):
        torch.cuda.current_stream().wait_stream(self.stream)
        s, t = self.s, self.t
        self.global_n = self._n
        if s is not None:
            s.record_stream(torch.cuda.current_stream())
            t.record_stream(torch.cuda.current_stream())
        self._preload()
        return s, t


def _own_share(prefetcher, x, y):
    """(this rank's share of a prefetched batch, size of the global batch)"""
    if getattr(prefetcher, "shards", False):
        return (x, y), prefetcher.global_n
    return shard_batch(x, y), int(x.shape[0])


def train_one_epoch(model, dataloader_forget, dataloader_remain, device, criterion, optimizer, epoch, losses_forget, losses_remain,
                    losses_total, losses_structure, top1_forget, top1_remain, beta, alpha, BND, batch, testloader_forget, testloader_remain,
                    forget_acc_before, highest_H_mean, cfg, task_i, use_prototype, prototype_dict, prototype_weight_forget,
                    prototype_weight_remain, losses_prototype_forget, losses_prototype_remain, dataloader_open=None):
    """Same contract as engine_cl.train_one_epoch (engine_cl.py:12-244); `criterion` must be nn.CrossEntropyLoss (mean)."""
    model.train()
    criterion.train()
    m = _unwrap(model)
    if engine_fresh_optimizer(m, optimizer):
        m.ensure_engine(1)
        m._engine.reset_optimizer()
    prefetcher = _Prefetcher(dataloader_forget, device)
    inputs_forget, labels_forget = prefetcher.next()
    DISP_FREQ, VER_FREQ = 5, 100
    rank0 = _dist() is None or _dist().get_rank() == 0
    pending = None      # (StepResult, n_remain, n_forget) of the step whose scalars have not been folded into the meters yet

    def absorb(p):
        out, nr, nf = p[0].wait(), p[1], p[2]
        losses_remain.update(out["loss_remain"], nr)
        top1_remain.update(out["top1_remain"], nr)
        losses_forget.update(beta * out["loss_forget"], nf)
        top1_forget.update(out["top1_forget"], nf)
        losses_structure.update(alpha * out["structure"], nr)
        # the reference updates this meter unconditionally (engine_cl.py:103-108): without prototypes KL_f is 0 and the logged value is w_f * BND_pro
        losses_prototype_forget.update(prototype_weight_forget * max(cfg.get("BND_pro", 0.0) - out["proto_forget"], 0.0), nr)
        losses_prototype_remain.update(out["proto_remain"] * prototype_weight_remain, nr)
        losses_total.update(out["total"], nr)

    for inputs_remain, labels_remain in iter(dataloader_remain):
        n_r = int(inputs_remain.size(0))                               # the meters weigh by the GLOBAL batch sizes
        xr, yr = shard_batch(inputs_remain, labels_remain)             # sharded on the host: each rank copies only its share over PCIe
        xr, yr = xr.to(device), yr.to(device)
        (xf, yf), n_f = _own_share(prefetcher, inputs_forget, labels_forget)
        res = unlearn_step_async(model, xr, yr, xf, yf, beta=beta, alpha=alpha, BND=BND,
                                 optimizer=optimizer, use_prototype=use_prototype, prototype_dict=prototype_dict,
                                 prototype_weight_forget=prototype_weight_forget, prototype_weight_remain=prototype_weight_remain,
                                 BND_pro=cfg.get("BND_pro", 0.0) if use_prototype else 0.0)
        # the scalars of step i are folded into the meters after step i+1 has been queued (the host never waits for the GPU in the
        # steady state); at the steps where the reference prints or evaluates, the meters are brought fully up to date first
        if pending is not None:
            absorb(pending)
        pending = (res, n_r, n_f)
        if (((batch + 1) % DISP_FREQ == 0) or ((batch + 1) % VER_FREQ == 0)) and batch != 0:
            absorb(pending)
            pending = None

        if ((batch + 1) % DISP_FREQ == 0) and batch != 0:
            if rank0:
                _wandb_log({
                    "epoch_loss_forget-{}".format(task_i): losses_forget.avg, "epoch_loss_remain-{}".format(task_i): losses_remain.avg,
                    "epoch_acc_forget-{}".format(task_i): top1_forget.avg, "epoch_acc_remain-{}".format(task_i): top1_remain.avg,
                    "epoch_loss_total-{}".format(task_i): losses_total.avg, "epoch_loss_structure-{}".format(task_i): losses_structure.avg,
                    "epoch_loss_prototype_forget-{}".format(task_i): losses_prototype_forget.avg,
                    "epoch_loss_prototype_remain-{}".format(task_i): losses_prototype_remain.avg})
                print("Task {} Epoch {} Batch {}\t"
                      "Training forget Loss {lf.val:.4f} ({lf.avg:.4f})\tTraining remain Loss {lr.val:.4f} ({lr.avg:.4f})\t"
                      "Training forget prototype Loss {pf.val:.4f}\tTraining remain prototype Loss {pr.val:.4f}\t"
                      "Training structure Loss {ls.val:.4f} ({ls.avg:.4f})\tTraining total Loss {lt.val:.4f} ({lt.avg:.4f})\t"
                      "Training forget Prec@1 {tf.val:.3f} ({tf.avg:.3f})\tTraining remain Prec@1 {tr.val:.3f} ({tr.avg:.3f})".format(
                          task_i, epoch + 1, batch + 1, lf=losses_forget, lr=losses_remain, pf=losses_prototype_forget,
                          pr=losses_prototype_remain, ls=losses_structure, lt=losses_total, tf=top1_forget, tr=top1_remain))
            losses_forget, losses_remain, top1_forget, top1_remain = AverageMeter(), AverageMeter(), AverageMeter(), AverageMeter()
            losses_total, losses_structure = AverageMeter(), AverageMeter()
            losses_prototype_forget, losses_prototype_remain = AverageMeter(), AverageMeter()

        with torch.no_grad():
            if ((batch + 1) % VER_FREQ == 0) and batch != 0:
                highest_H_mean = evaluate(model, testloader_forget=testloader_forget, testloader_remain=testloader_remain, device=device,
                                          batch=batch, epoch=epoch, task_i=task_i, forget_acc_before=forget_acc_before,
                                          highest_H_mean=highest_H_mean, cfg=cfg, optimizer=optimizer, testloader_open=dataloader_open)
                model.train()
        batch += 1
        inputs_forget, labels_forget = prefetcher.next()
        if inputs_forget is None:
            prefetcher = _Prefetcher(dataloader_forget, device)
            inputs_forget, labels_forget = prefetcher.next()
    if pending is not None:
        absorb(pending)
    sync_optimizer_state(model, optimizer)
    return (batch, highest_H_mean, losses_forget, losses_remain, top1_forget, top1_remain, losses_total, losses_structure,
            losses_prototype_forget, losses_prototype_remain)


def engine_fresh_optimizer(m, optimizer) -> bool:
    """A new torch optimizer object (the driver re-creates it per task, train_own_forget_cl.py:811-813) resets the fused moments."""
    key = id(optimizer)
    if getattr(m, "_gsl_optimizer_id", None) != key:
        m._gsl_optimizer_id = key
        return True
    return False


def evaluate(model, testloader_forget, testloader_remain, device, batch, epoch, forget_acc_before, highest_H_mean, cfg, optimizer, task_i,
             testloader_open=None):
    """engine_cl.evaluate (engine_cl.py:247-315): eval-mode (merged) accuracies, H-mean, rolling best checkpoint."""
    model.eval()
    lr = optimizer.param_groups[0]["lr"]
    print("current learning rate:{:.7f}".format(lr))
    print("Perfom evaluation on test set and save checkpoints...")
    forget_acc = eval_data(model, testloader_forget, device, "forget-{}".format(task_i), batch)
    remain_acc = eval_data(model, testloader_remain, device, "remain-{}".format(task_i), batch)
    if testloader_open is not None:
        eval_data(model, testloader_open, device, "open-{}".format(task_i), batch)
    forget_drop = forget_acc_before - forget_acc
    Hmean = 2 * forget_drop * remain_acc / (forget_drop + remain_acc + 1e-8)
    rank0 = _dist() is None or _dist().get_rank() == 0
    if Hmean > highest_H_mean:
        highest_H_mean = Hmean
        if rank0:
            path = os.path.join(cfg["WORK_PATH"], "Backbone_{}_Epoch_{}_Batch_{}_Time_{}_checkpoint.pth".format(
                cfg["BACKBONE_NAME"], epoch + 1, batch + 1, get_time()))
            torch.save(_unwrap(model).state_dict(), path)     # written in eval mode: `weight` holds the merged LoRA delta, as in the reference
            if len(os.listdir(cfg["WORK_PATH"])) >= 4:
                ckpts = sorted((f for f in os.listdir(cfg["WORK_PATH"]) if f.endswith(".pth")),
                               key=lambda f: os.path.getmtime(os.path.join(cfg["WORK_PATH"], f)))
                os.remove(os.path.join(cfg["WORK_PATH"], ckpts[0]))
    return highest_H_mean


def eval_data(model, dataloader, device, mode: str, batch: int = 0):
    """engine_cl.eval_data (engine_cl.py:318-346): top-1 accuracy (0-100) in eval mode; hits are counted by the head kernel."""
    m = _unwrap(model)
    hits = torch.zeros((), dtype=torch.int64, device=device)
    total = 0
    model.eval()
    with torch.no_grad():
        for images, labels in dataloader:
            images = m.prepare_images(images.to(device))
            labels = labels.to(device).long().contiguous()
            for slot, _, B in m.inference_slots(images, labels):          # chunks of the engine's capacity: eval never grows the workspace
                hits += m._engine.slot_tensor(slot, F.SLOT_CORRECT, B).sum()
            total += labels.size(0)
    accuracy = 100 * int(hits.item()) / max(total, 1)
    print("Test {} Accuracy:{:2f}%".format(mode, accuracy))
    _wandb_log({"Test {} Accuracy".format(mode): accuracy})
    return accuracy


# ------------------------------------------------------------------------------------------------ baselines' loop (OUT of the hot path)
def _reference_engine_cl():
    """The reference's own engine_cl.py, loaded under a private name (EWC / MAS / L2 regularisation baselines, engine_cl.py:435-568: full
    autograd loops that are not the GS-LoRA path -- SURVEY section 2 row 3).  They run on the engine-backed model through its autograd seam."""
    import importlib.util
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for r in [os.environ.get("GSLORA_REFERENCE_ROOT", "")] + list(sys.path):
        f = os.path.join(r, "engine_cl.py") if r else ""
        if f and os.path.isfile(f) and os.path.abspath(r) != here:
            spec = importlib.util.spec_from_file_location("_gslora_ref_engine_cl", f)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod
    raise NotImplementedError("gslora-b200: train_one_epoch_regularzation / get_reg_loss belong to the reference's EWC / MAS / L2 baselines "
                              "(outside the GS-LoRA hot path); put the reference tree on sys.path or set GSLORA_REFERENCE_ROOT to use them")


def train_one_epoch_regularzation(*args, **kwargs):
    """engine_cl.train_one_epoch_regularzation (engine_cl.py:435-568), imported by the driver (train_own_forget_cl.py:41): delegated to the
    reference's implementation."""
    return _reference_engine_cl().train_one_epoch_regularzation(*args, **kwargs)


def get_reg_loss(*args, **kwargs):
    return _reference_engine_cl().get_reg_loss(*args, **kwargs)
