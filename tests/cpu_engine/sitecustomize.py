"""TEST INFRASTRUCTURE (never on the product path, never shipped): lets the UNMODIFIED reference driver run to completion in the authoring
container, which has no GPU, by standing an ORACLE-backed engine in for the native one inside the driver's subprocess.

How it is used: tests/test_driver_full_cpu.py puts THIS directory first on the subprocess' PYTHONPATH; Python imports `sitecustomize` at start-up,
which swaps gslora.engine.VitEngine for OracleEngine (oracle/vit_oracle.py arithmetic on CPU tensors, dropout ignored) and replaces the three
CUDA-only host utilities of engine_cl (pinned ring, CUDA event, side-stream prefetcher).  Everything else the driver touches -- the module
surface, loralib merge / un-merge, engine_cl.train_one_epoch / evaluate / eval_data, util.cal_norm, calculate_prototypes, checkpoints, the
per-task optimizer reset -- is the repo's own host code, unchanged.  What this proves is the HOST side of "the driver drops in unchanged"; the
device side of the same call sequence runs on the GPU in tests/test_driver_replay_gpu.py.  GSLORA_TRACE=<file> records the sequence of
drop-in entry points the driver calls (one name per line)."""
import os
import sys

if os.environ.get("GSLORA_CPU_ORACLE_ENGINE") == "1":
    import torch

    ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if ROOT not in sys.path:
        sys.path.append(ROOT)                      # for `oracle`
    from oracle import vit_oracle as O
    import gslora.engine as _eng
    import gslora.model_base as _mb
    from gslora import _ffi as F

    _TRACE = os.environ.get("GSLORA_TRACE")

    def trace(name):
        if _TRACE:
            with open(_TRACE, "a") as f:
                f.write(name + "\n")

    _GLOBAL = ["pos_embedding", "cls_token", "patch_to_embedding.weight", "patch_to_embedding.bias", "mlp_head.0.weight", "mlp_head.0.bias",
               "loss.weight", None]
    _BLOCK = ["0.fn.norm.weight", "0.fn.norm.bias", "0.fn.fn.to_qkv.weight", None, "0.fn.fn.to_out.0.weight", "0.fn.fn.to_out.0.bias",
              "1.fn.norm.weight", "1.fn.norm.bias", "1.fn.fn.net.0.weight", "1.fn.fn.net.0.bias", "1.fn.fn.net.3.weight", "1.fn.fn.net.3.bias"]

    class OracleEngine:
        """VitEngine's Python surface (gslora/engine.py) on CPU tensors with the oracle's arithmetic."""

        def __init__(self, spec, device, max_batch, num_slots=1):
            assert spec.head_type == 0 and spec.patch_order == 0, "the CPU stand-in restates ViT_face only"
            self.spec, self.device, self.max_batch, self.num_slots = spec, device, int(max_batch), int(num_slots)
            self.precision = 1
            self.cfg = O.VitConfig(image_size=spec.image_size, patch_size=spec.patch_size, dim=spec.dim, depth=spec.depth, heads=spec.heads,
                                   mlp_dim=spec.mlp_dim, num_class=spec.num_class, channels=spec.channels, lora_rank=spec.lora_rank,
                                   cos_s=spec.cos_s, cos_m=spec.cos_m, ln_eps=spec.ln_eps, lora_pos="Attention" if spec.lora_pos == 1 else "FFN")
            n = spec.depth * spec.lora_block_elems
            self.lora_flat, self.grad_flat = torch.zeros(n), torch.zeros(n)
            self.exp_avg, self.exp_avg_sq = torch.zeros(n), torch.zeros(n)
            self.opt_step = 0
            offs = []
            for l in range(spec.depth):
                o = l * spec.lora_block_elems
                for shp in spec.lora_shapes():
                    offs.append(o)
                    o += shp[0] * shp[1]
            offs.append(n)
            self.tensor_offsets_host = offs
            tpb = spec.tensors_per_block
            blocks = offs[0::tpb]
            self.group_offsets_by_type = {"block": blocks, "lora": offs[0::2] if spec.lora_pos == 0 else blocks, "matrix": offs if spec.lora_pos == 0 else blocks}
            self.group_norms = torch.zeros(4 * spec.depth)
            self.num_groups = spec.depth
            self.sums = torch.zeros(8)
            self.slots = [None] * self.num_slots
            self.frozen = None

        lora_view = _eng.VitEngine.lora_view

        def bind(self, frozen):
            self.frozen = list(frozen)

        def refresh_frozen(self):
            pass

        def refresh_lora(self):
            pass

        def _sd(self, leaves):
            sd = {}
            for name, t in zip(_GLOBAL, self.frozen[:8]):
                if name is not None and t is not None:
                    sd[name] = t
            for l in range(self.spec.depth):
                for name, t in zip(_BLOCK, self.frozen[8 + 12 * l: 20 + 12 * l]):
                    if name is not None:
                        sd[O.blk(l, name)] = t
            for n, t in zip(O.lora_param_list(self.cfg), leaves):
                sd[n] = t
            return sd

        def forward(self, img, labels, slot=0, use_lora=True, dropout_seed=0, pixel_norm=None, channels_last=False):
            assert img.dtype == torch.float32 and pixel_norm is None
            B = img.shape[0]
            tpb = self.spec.tensors_per_block
            leaves = [self.lora_view(self.lora_flat, l, w).detach().clone().requires_grad_(True) for l in range(self.spec.depth) for w in range(tpb)]
            with torch.enable_grad():
                emb = O.vit_embed(self._sd(leaves), self.cfg, img, use_lora=bool(use_lora))
                logits = O.cosface(emb, self.frozen[6], labels, self.cfg.cos_s, self.cfg.cos_m) if labels is not None else None
            rec = dict(B=B, leaves=leaves, emb=emb, logits=logits)
            if labels is not None:
                rec["ce"] = torch.nn.functional.cross_entropy(logits.detach(), labels, reduction="none")
                rec["correct"] = (logits.detach().argmax(1) == labels).to(torch.int32)
            self.slots[slot] = rec
            return B

        def slot_tensor(self, slot, what, B):
            r = self.slots[slot]
            return {F.SLOT_EMB: r["emb"].detach(), F.SLOT_LOGITS: None if r["logits"] is None else r["logits"].detach(), F.SLOT_CE: r.get("ce"),
                    F.SLOT_CORRECT: r.get("correct")}[what]

        def class_sums(self, slot, labels, B, sums, counts):
            emb = self.slots[slot]["emb"].detach()
            for e, l in zip(emb, labels):                    # util/utils.py:535-541 order
                sums[int(l)] += e
                counts[int(l)] += 1

        def backward(self, slot, dlogits, demb, accumulate=False):
            r = self.slots[slot]
            outs, gos = [], []
            if dlogits is not None:
                outs.append(r["logits"]); gos.append(dlogits)
            if demb is not None:
                outs.append(r["emb"]); gos.append(demb)
            grads = torch.autograd.grad(outs, r["leaves"], gos, allow_unused=True)
            tpb = self.spec.tensors_per_block
            for i, g in enumerate(grads):
                view = self.lora_view(self.grad_flat, i // tpb, i % tpb)
                g = torch.zeros_like(view) if g is None else g
                view.copy_(view + g if accumulate else g)

        def loss_sums(self, slot, n_remain, B, kl=None):
            r = self.slots[slot]
            ce, ok = r["ce"], r["correct"].float()
            kl = torch.zeros(B) if kl is None else kl
            self.sums = torch.stack([ce[:n_remain].sum(), torch.tensor(float(n_remain)), ce[n_remain:].sum(), torch.tensor(float(B - n_remain)),
                                     ok[:n_remain].sum(), ok[n_remain:].sum(), kl[:n_remain].sum(), kl[n_remain:].sum()]).float()
            return self.sums

        def prototype_kl(self, slot, labels, proto, B):
            emb = self.slots[slot]["emb"].detach()
            lp, lq = torch.log_softmax(proto[labels], 1), torch.log_softmax(emb, 1)
            return (lp.exp() * (lp - lq)).sum(1)

        def prototype_kl_grad(self, slot, labels, proto, n_remain, B, w_f, w_r, BND_pro):
            emb = self.slots[slot]["emb"].detach()
            s = self.sums
            d = torch.softmax(emb, 1) - torch.softmax(proto[labels], 1)            # d KL_b / d emb_b
            w = torch.zeros(B)
            if s[1] > 0:
                w[:n_remain] = w_r / s[1]
            if s[3] > 0 and s[7] / s[3] < BND_pro:
                w[n_remain:] = -w_f / s[3]
            return d * w[:, None]

        def unlearn_ce_grad(self, slot, labels, n_remain, B, beta, BND, out):
            logits = self.slots[slot]["logits"].detach()
            s = self.sums
            w = torch.zeros(B)
            if s[1] > 0:
                w[:n_remain] = 1.0 / s[1]
            if s[3] > 0 and s[2] / s[3] < BND:
                w[n_remain:] = -beta / s[3]
            out.copy_((torch.softmax(logits, 1) - torch.nn.functional.one_hot(labels, logits.shape[1]).float()) * w[:, None])

        def optimizer_step(self, lr, wd, alpha, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0, group_type="block"):
            self.opt_step += 1
            offs = self.group_offsets_by_type[group_type]
            self.num_groups = len(offs) - 1
            t, (b1, b2) = self.opt_step, betas
            for gi in range(self.num_groups):
                sl = slice(offs[gi], offs[gi + 1])
                p = self.lora_flat[sl]
                norm = p.norm()
                self.group_norms[gi] = norm
                g = self.grad_flat[sl] * grad_scale + (alpha * p / norm if (alpha != 0 and norm > 0) else 0)
                m, v = self.exp_avg[sl], self.exp_avg_sq[sl]
                m.mul_(b1).add_(g, alpha=1 - b1)
                v.mul_(b2).addcmul_(g, g, value=1 - b2)
                p.mul_(1 - lr * wd)
                p.addcdiv_(m, (v.sqrt() / (1 - b2 ** t) ** 0.5).add_(eps), value=-lr / (1 - b1 ** t))

        def reset_optimizer(self):
            trace("reset_optimizer")
            self.exp_avg.zero_(); self.exp_avg_sq.zero_(); self.opt_step = 0

        def tensor_norms(self, type="L2"):
            o = self.tensor_offsets_host
            return torch.stack([self.lora_flat[o[i]:o[i + 1]].norm() if type == "L2" else self.lora_flat[o[i]:o[i + 1]].abs().sum() for i in range(len(o) - 1)])

    _eng.VitEngine = OracleEngine
    _mb.VitEngine = OracleEngine

    _orig_ensure = _mb.EngineBackedModel.ensure_engine

    def _ensure_engine_cpu(self, batch, slots=None):
        """EngineBackedModel.ensure_engine without its CUDA-only guard (the guard itself is covered by tests/test_driver_dropin_cpu.py)"""
        slots = slots or int(os.environ.get("GSLORA_SLOTS", "2"))
        e = self._engine
        if e is None or e.max_batch < batch or e.num_slots < slots:
            carry = None if e is None else (e.exp_avg.clone(), e.exp_avg_sq.clone(), e.opt_step)
            self._engine = OracleEngine(self._spec(), next(self.parameters()).device, max(batch, e.max_batch if e else 0), slots)
            if carry is not None:
                self._engine.exp_avg.copy_(carry[0]); self._engine.exp_avg_sq.copy_(carry[1]); self._engine.opt_step = carry[2]
            self._frozen_sig = self._lora_sig = None
            self._slot_stamp = [0] * slots
            self._slot_next = 0
        return self._engine
    _mb.EngineBackedModel.ensure_engine = _ensure_engine_cpu

    class _Event:
        def record(self, *a):
            pass

        def synchronize(self):
            pass
    torch.cuda.Event = _Event

    import engine_cl as _cl
    import gslora.prototypes as _pr

    class _Ring:
        def __init__(self, n=8, width=16):
            self.bufs = [torch.empty(width) for _ in range(n)]
            self.pending = [None] * n
            self.i = 0
        take = _cl._PinnedRing.take
    _cl._PinnedRing = _Ring

    class _CpuPrefetcher:
        shards = True

        def __init__(self, loader, device):
            self.it = iter(loader)
            self.global_n = 0

        def next(self):
            s, t = next(self.it, (None, None))
            if s is not None:
                self.global_n = int(s.shape[0])
                s, t = _cl.shard_batch(s, t)
            return s, t
    _cl._Prefetcher = _CpuPrefetcher

    def _class_prototype_table(backbone, batches, device="cuda"):
        m = _pr._unwrap(backbone)
        backbone.eval()
        sums = counts = None
        with torch.no_grad():
            for images, labels in batches:
                images, labels = m.prepare_images(images), labels.long().contiguous()
                for slot, lo, B in m.inference_slots(images, labels):
                    eng = m._engine
                    if sums is None:
                        sums, counts = torch.zeros(eng.spec.num_class, eng.spec.dim), torch.zeros(eng.spec.num_class)
                    eng.class_sums(slot, labels[lo:lo + B], B, sums, counts)
        if sums is None:
            return None, None
        return torch.where(counts[:, None] > 0, sums / counts.clamp(min=1)[:, None], torch.zeros_like(sums)), counts
    _pr.class_prototype_table = _class_prototype_table

    # ---- call trace of the drop-in entry points (what tests/driver_replay.py must re-enact)
    def _wrap(mod, name, label=None):
        fn = getattr(mod, name)

        def w(*a, **k):
            trace(label or name)
            return fn(*a, **k)
        w.__name__ = name
        setattr(mod, name, w)
    for _n in ("train_one_epoch", "eval_data", "evaluate"):
        _wrap(_cl, _n)
    _wrap(_pr, "calculate_prototypes")
    import util.cal_norm as _cn
    _wrap(_cn, "get_norm_of_lora")
    _orig_save = torch.save

    def _save(obj, f, *a, **k):
        if isinstance(f, str) and "task-level" in f:
            trace("save_task_checkpoint")
        return _orig_save(obj, f, *a, **k)
    torch.save = _save
    import util.utils as _uu
    if hasattr(_uu, "reinitialize_lora_parameters"):
        _wrap(_uu, "reinitialize_lora_parameters")
