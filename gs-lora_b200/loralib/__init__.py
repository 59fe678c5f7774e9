"""gslora-b200 `loralib` -- drop-in for the `loralib==0.1.2` surface the reference imports
(`import loralib as lora`; vit_pytorch_face/vit_face.py:330-355, util/utils.py:573, train/train_own_forget_cl.py:316).

The layers are parameter holders with loralib's semantics (frozen `weight`, trainable `lora_A [r, in]` /
`lora_B [out, r]`, `scaling = lora_alpha / r`, merge on `eval()` / un-merge on `train()`); the arithmetic of the
forward/backward runs in libgslora.so (fused into the FFN GEMMs when the layer lives inside `ViT_face`)."""
from .layers import Linear, LoRALayer, MergedLinear, mark_only_lora_as_trainable, lora_state_dict  # noqa: F401

__version__ = "0.1.2+gslora_b200"
