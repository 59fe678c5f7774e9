"""Drop-in overlay for the reference's `util.utils` (the driver does `from util.utils import (..., calculate_prototypes, ...,
reinitialize_lora_parameters, ...)`, train/train_own_forget_cl.py:15-29).

Everything that is host-side orchestration (dataset splits, few-shot sampling, head surgery, meters, verification) is the reference's own
code: when the reference tree is importable its `util/utils.py` is executed into this module unchanged.  Only the two functions that touch
the hot path's data are gslora-b200's:
  calculate_prototypes  -> gslora.prototypes (batched eval forward + gsl_class_sums / gsl_class_means kernels, no per-image `.item()`)
Without the reference tree (e.g. on a benchmark box) the small helpers engine_cl needs are defined here."""
import datetime
import math
import os
import sys

import torch
import torch.nn as nn

_ref_file = None
for _d in sys.modules[__package__].__path__[1:] if __package__ else []:
    if os.path.isfile(os.path.join(_d, "utils.py")):
        _ref_file = os.path.join(_d, "utils.py")
        break

_ref_loaded = False
if _ref_file is not None:
    try:
        with open(_ref_file) as _f:
            exec(compile(_f.read(), _ref_file, "exec"), globals())
        _ref_loaded = True
    except Exception as _e:        # a third-party import of the reference's host-side helpers is missing (mxnet, matplotlib, IPython ...)
        _ref_error = _e

if not _ref_loaded:
    def get_time():
        return (str(datetime.datetime.now())[:-10]).replace(" ", "-").replace(":", "-")

    class AverageMeter(object):
        def __init__(self):
            self.reset()

        def reset(self):
            self.val = self.avg = self.sum = self.count = 0

        def update(self, val, n=1):
            self.val = val
            self.sum += val * n
            self.count += n
            self.avg = self.sum / self.count

    def train_accuracy(output, target, topk=(1,)):
        """precision@k in percent for the FIRST k of `topk` (util/utils.py:354-368 returns res[0])"""
        maxk = max(topk)
        _, pred = output.topk(maxk, 1, True, True)
        correct = pred.t().eq(target.view(1, -1).expand_as(pred.t()))
        return correct[:topk[0]].reshape(-1).float().sum(0).mul_(100.0 / target.size(0))

    def count_trainable_parameters(model):
        return sum(p.numel() for p in model.parameters() if p.requires_grad)

    def reinitialize_lora_parameters(model):
        """util/utils.py:428-441: lora_A <- kaiming_uniform(a = sqrt(50)), lora_B <- 0 (in place; the engine sees the version bump)."""
        with torch.no_grad():
            for name, param in model.named_parameters():
                if "lora_A" in name:
                    nn.init.kaiming_uniform_(param, a=math.sqrt(50))
                elif "lora_B" in name:
                    nn.init.zeros_(param)

_reference_calculate_prototypes = globals().get("calculate_prototypes")


def calculate_prototypes(backbone, dataset, batch_size=32, device="cuda", aug_num=0):
    """util/utils.py:502-549 for engine-backed models (device-side class sums); any other module goes to the reference's own function."""
    from gslora.model_base import EngineBackedModel
    from gslora.prototypes import calculate_prototypes as _device_prototypes, _unwrap
    if isinstance(_unwrap(backbone), EngineBackedModel):
        return _device_prototypes(backbone, dataset, batch_size=batch_size, device=device, aug_num=aug_num)
    if _reference_calculate_prototypes is None:
        raise NotImplementedError("calculate_prototypes: not an engine-backed model and the reference's util/utils.py is not importable")
    return _reference_calculate_prototypes(backbone, dataset, batch_size=batch_size, device=device, aug_num=aug_num)
