"""`util.cal_norm.get_norm_of_lora` of the reference (util/cal_norm.py:4-146) for the gslora-b200 model.

Per group the SUM of the per-matrix norms ||P||_F (type 'L2') or ||P||_1 ('L1') -- a different quantity from the
group-lasso norm sqrt(sum ||P||_F^2) used by the structure loss.  For an engine-backed ViT_face with the block / lora /
matrix FFN groupings the 4*depth per-tensor norms come from one CUDA kernel (gsl_tensor_norms) and are summed per group."""
import torch


def _ffn_groups(group_num, group_type):
    a1, b1, a2, b2 = 0, 1, 2, 3
    if group_type == "block":
        return [[(i, a1), (i, b1), (i, a2), (i, b2)] for i in range(group_num)]
    if group_type == "lora":
        return [[(i, a1), (i, b1)] for i in range(group_num)] + [[(i, a2), (i, b2)] for i in range(group_num)]
    if group_type == "matrix":
        return ([[(i, a1)] for i in range(group_num)] + [[(i, b1)] for i in range(group_num)] +
                [[(i, a2)] for i in range(group_num)] + [[(i, b2)] for i in range(group_num)])
    raise ValueError("group_type should be block, lora or matrix")


_NAMES = ["transformer.layers.{}.1.fn.fn.net.0.lora_A", "transformer.layers.{}.1.fn.fn.net.0.lora_B",
          "transformer.layers.{}.1.fn.fn.net.3.lora_A", "transformer.layers.{}.1.fn.fn.net.3.lora_B"]


_NAMES_TV = ["encoder.layers.encoder_layer_{}.mlp.0.lora_A", "encoder.layers.encoder_layer_{}.mlp.0.lora_B",
             "encoder.layers.encoder_layer_{}.mlp.3.lora_A", "encoder.layers.encoder_layer_{}.mlp.3.lora_B"]


def get_norm_of_lora(model, type="L2", group_num=6, group_type: str = "block", group_pos: str = "FFN", imagenet: bool = False):
    if type not in ("L1", "L2"):
        raise ValueError("type should be L1 or L2")
    if group_pos == "Attention":       # util/cal_norm.py:100-113: one (to_qkv.lora_A, to_qkv.lora_B) group per block, group_type ignored
        if getattr(model, "lora_pos", "FFN") != "Attention":
            raise KeyError("get_norm_of_lora(group_pos='Attention'): the model carries no LoRA on to_qkv")
        names_a = ["transformer.layers.{}.0.fn.fn.to_qkv.lora_A", "transformer.layers.{}.0.fn.fn.to_qkv.lora_B"]
        print("\033[31mgroup_layers_names\033[0m\n", [[n.format(i) for n in names_a] for i in range(group_num)])
        with torch.no_grad():
            eng = getattr(model, "_engine", None) or model.ensure_engine(1)
            model.sync_engine()
            per_tensor = eng.tensor_norms(type)          # [2 * depth]
            if group_num > eng.spec.depth:
                raise KeyError(f"get_norm_of_lora: group_num={group_num} but the model has {eng.spec.depth} blocks")
            return [per_tensor[2 * i] + per_tensor[2 * i + 1] for i in range(group_num)]
    if group_pos != "FFN":
        raise ValueError("group_pos should be FFN or Attention")
    if getattr(model, "lora_pos", "FFN") != "FFN":
        raise KeyError("get_norm_of_lora(group_pos='FFN'): the model carries its LoRA on to_qkv (lora_pos='Attention')")
    # imagenet=True: the reference ignores group_num and group_type and always reports the first 12 encoder blocks (util/cal_norm.py:82-99;
    # the driver passes group_num=args.vit_depth=6 for ViT-B/16, train_own_forget_cl.py:1100-1105)
    groups = _ffn_groups(12, "block") if imagenet else _ffn_groups(group_num, group_type)
    names = _NAMES_TV if imagenet else _NAMES
    print("\033[31mgroup_layers_names\033[0m\n", [[names[w].format(i) for i, w in g] for g in groups])
    with torch.no_grad():
        eng = getattr(model, "_engine", None)
        if eng is None and hasattr(model, "ensure_engine"):
            eng = model.ensure_engine(1)
        model.sync_engine()
        per_tensor = eng.tensor_norms(type)          # [4 * depth] on device
        if max(i for g in groups for i, _ in g) >= eng.spec.depth:
            raise KeyError(f"get_norm_of_lora: the grouping names block {max(i for g in groups for i, _ in g)} but the model has {eng.spec.depth}")
        return [sum(per_tensor[4 * i + w] for i, w in g) for g in groups]
