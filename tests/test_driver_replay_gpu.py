"""The driver's call sequence after its first engine call, on the GPU, for both backbones the continual driver builds (-n VIT, -n VIT_B16):
tests/driver_replay.py re-enacts train/train_own_forget_cl.py:494-536, 633-646, 807-820, 899-937, 1000-1106, 1696-1705 on the drop-in surface
(2 tasks, prototypes on, EMA on).  Checked: the run completes; task checkpoints carry the reference's key set with MERGED weights next to the
lora_* tensors; task t+1 starts from exactly the function task t saved; LoRA re-initialisation and the per-task optimizer reset are seen by
the engine; evaluation batches 5x the training batch do not grow the engine's workspace."""
import os

import pytest
import torch

from oracle import vit_oracle as O
from driver_replay import replay

pytestmark = pytest.mark.gpu


def _vit_face():
    import loralib as lora
    from vit_pytorch_face import ViT_face
    torch.manual_seed(11)
    m = ViT_face(loss_type="CosFace", GPU_ID=[0], num_class=20, image_size=112, patch_size=8, dim=512, depth=3, heads=8, mlp_dim=2048,
                 dropout=0.1, emb_dropout=0.1, lora_rank=8)                     # train_own_forget_cl.py:206-221
    with torch.no_grad():
        m.pos_embedding.mul_(0.02)
        m.cls_token.mul_(0.02)
    lora.mark_only_lora_as_trainable(m)                                         # :316
    return m


def _vit_b16():
    import loralib as lora
    from torchvision.models.vision_transformer import VisionTransformer
    from vit_pytorch_face import ModifiedViT
    torch.manual_seed(12)
    # full depth 12: with imagenet=True the reference's norm report names encoder_layer_0 .. 11 whatever --vit_depth says (util/cal_norm.py:82-99)
    tv = VisionTransformer(image_size=224, patch_size=16, num_layers=12, num_heads=12, hidden_dim=768, mlp_dim=3072, num_classes=20)
    m = ModifiedViT(tv)                                                         # :226-242
    for blk in m.encoder.layers.children():                                     # util.utils.replace_ffn_with_lora (util/utils.py:552-576)
        blk.mlp[0] = lora.Linear(768, 3072, r=8)
        blk.mlp[3] = lora.Linear(3072, 768, r=8)
    with torch.no_grad():
        m.heads.head.weight.normal_(0, 0.02)
    lora.mark_only_lora_as_trainable(m)
    return m


@pytest.mark.parametrize("kind", ["VIT", "VIT_B16"])
def test_driver_call_sequence_two_tasks(tmp_path, kind):
    model = _vit_face() if kind == "VIT" else _vit_b16()
    ref_keys = set(model.state_dict().keys())
    log = []
    out = replay(model, image_size=112 if kind == "VIT" else 224, num_class=20, device=torch.device("cuda"), work_path=str(tmp_path), num_tasks=2,
                 epochs=2, batch_size=8, per_class=4, imagenet=kind == "VIT_B16", prototype=True, average_weight=True, log=log)
    assert [e[0] for e in log].count("task_done") == 2 and ("reload+reinit", 1) in log
    m = out["model"]
    # evaluation ran with 5x batches (40 images) but the workspace was never sized beyond what training + GSLORA_EVAL_CHUNK need
    assert m._engine.max_batch <= 128
    for t, rec in enumerate(out["tasks"]):
        sd = torch.load(rec["ckpt"])
        assert set(sd.keys()) == ref_keys                                       # same key set the reference saves (weights + lora_A / lora_B)
        assert rec["steps"] > 0 and rec["total"] is not None and all(n == n and n >= 0 for n in rec["norms"])
        assert rec["opt_state_keys"] == 4 * (3 if kind == "VIT" else 12)        # sync_optimizer_state exposed the fused moments of every LoRA tensor
        for k, v in rec["acc"].items():
            assert 0.0 <= v <= 100.0, (k, v)
    # the task-0 checkpoint was written in eval mode: its FFN weights contain the LoRA delta of that moment (loralib merge, SURVEY A-10)
    sd0 = torch.load(out["tasks"][0]["ckpt"])
    wname = next(k for k in sd0 if k.endswith("net.0.weight") or k.endswith("mlp.0.weight"))
    base = wname[: -len("weight")]
    A, B = sd0[base + "lora_A"], sd0[base + "lora_B"]
    assert float(B.abs().max()) > 0                                             # task 0 trained: B left zero
    fresh = (_vit_face() if kind == "VIT" else _vit_b16()).state_dict()[wname]
    delta = sd0[wname].cpu() - fresh
    want = (B.cpu() @ A.cpu()) / 8.0
    assert float((delta - want).norm() / want.norm()) < 1e-4
    # task 1 started from the function task 0 saved: reload + zeroed lora_B == the merged eval-mode model
    import loralib as lora
    again = (_vit_face() if kind == "VIT" else _vit_b16()).cuda()
    again.load_state_dict(sd0)
    from driver_replay import reinitialize_lora_parameters
    reinitialize_lora_parameters(again)
    again.eval()
    with torch.no_grad():
        got = again(out["probe_x"].cuda(), out["probe_y"].cuda())
    got = (got[0] if isinstance(got, tuple) else got).cpu()
    want_logits = out["tasks"][0]["probe_logits"]
    assert float((got - want_logits).norm() / want_logits.norm()) < 1e-3
    # EMA copy is a working engine-backed model of its own
    assert out["ema"] is not None and out["ema"]._engine is not None and out["ema"]._engine is not m._engine
