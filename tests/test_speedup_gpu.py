"""North-star denominator: the reference's step as stock PyTorch FP32 eager ON THE SAME B200 (the oracle restatement
of engine_cl.py:59-125: two forwards, CE + bounded forget loss + structure loss, autograd backward, torch AdamW) timed beside the fused
step at the headline batch (512 + 512).  Prints the ratio; asserts the north-star's >= 10x."""
import json
import os

import pytest
import torch

from oracle import vit_oracle as O

pytestmark = pytest.mark.gpu


def _time(fn, warmup, steps):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def test_fused_step_vs_reference_fp32_eager_on_b200():
    import engine_cl
    import loralib as lora
    from vit_pytorch_face import ViT_face
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    B = int(os.environ.get("GSLORA_SPEEDUP_BATCH", "512"))
    cfg = O.P8S8
    sd = O.init_state_dict(cfg, seed=1337)
    gen = torch.Generator().manual_seed(3)
    xr, xf = torch.rand(B, 3, 112, 112, generator=gen).cuda(), torch.rand(B, 3, 112, 112, generator=gen).cuda()
    yr, yf = torch.randint(0, 100, (B,), generator=gen).cuda(), torch.randint(0, 100, (B,), generator=gen).cuda()
    # reference arm: FP32 eager on the GPU
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    state = {}
    ref_ms = _time(lambda: O.unlearn_step(sd_gpu, cfg, state, xr, yr, xf, yf, lr=1e-2, wd=0.05, beta=0.15, alpha=1e-4, BND=105.0), 2, 5)
    del sd_gpu, state
    torch.cuda.empty_cache()
    # fused arm
    model = ViT_face(loss_type="CosFace", GPU_ID=[0], num_class=100, image_size=112, patch_size=8, dim=512, depth=6, heads=8, mlp_dim=2048,
                     dropout=0.0, emb_dropout=0.0, lora_rank=8)
    model.load_state_dict(sd, strict=True)
    lora.mark_only_lora_as_trainable(model)
    model = model.cuda().train()
    kw = dict(beta=0.15, alpha=1e-4, BND=105.0, hparams=dict(lr=1e-2, wd=0.05))
    ours_ms = _time(lambda: engine_cl.unlearn_step(model, xr, yr, xf, yf, **kw), 3, 10)
    line = {"batch": f"{B}+{B}", "reference_fp32_eager_ms": round(ref_ms, 2), "gslora_b200_ms": round(ours_ms, 2), "speedup": round(ref_ms / ours_ms, 2),
            "images_per_s": {"reference": round(2 * B / ref_ms * 1e3, 1), "gslora_b200": round(2 * B / ours_ms * 1e3, 1)}}
    print("SPEEDUP " + json.dumps(line))
    out = os.environ.get("GSLORA_SPEEDUP_OUT")
    if out:
        with open(out, "w") as f:
            json.dump(line, f)
    assert ref_ms / ours_ms >= 10.0, line
