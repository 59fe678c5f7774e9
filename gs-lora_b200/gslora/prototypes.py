"""Class prototypes on device -- util.utils.calculate_prototypes of the reference (util/utils.py:502-549).

The reference runs an eval-mode forward per batch and then walks the batch on the host, one `label.item()` device sync per image
(`embeds_sum[label.item()] += embed`).  Here the per-class accumulation is the gsl_class_sums kernel (batch order = the reference's fp32
summation order, so equal embeddings give bit-equal prototypes) and the only device-to-host traffic is one copy of the finished table."""
from __future__ import annotations

import torch

from . import _ffi as F
from .model_base import EngineBackedModel


def _unwrap(model):
    return model.module if isinstance(model, (torch.nn.DataParallel, torch.nn.parallel.DistributedDataParallel)) else model


def class_prototype_table(backbone, batches, device="cuda"):
    """(means [C, D] fp32, counts [C] fp32) on `device` for an iterable of (images, labels) batches; eval-mode (merged) forward."""
    m = _unwrap(backbone)
    if not isinstance(m, EngineBackedModel):
        raise TypeError("class_prototype_table needs a gslora-b200 engine-backed model (ViT_face / ModifiedViT)")
    backbone.eval()
    backbone.to(device)
    sums = counts = None
    with torch.no_grad():
        for images, labels in batches:
            images = m.prepare_images(images.to(device))
            labels = labels.to(device).long().contiguous()
            for slot, lo, B in m.inference_slots(images, labels):         # chunks in dataset order: the per-class fp32 summation order is unchanged
                eng = m._engine
                if sums is None:
                    sums = torch.zeros(eng.spec.num_class, eng.spec.dim, dtype=torch.float32, device=eng.device)
                    counts = torch.zeros(eng.spec.num_class, dtype=torch.float32, device=eng.device)
                eng.class_sums(slot, labels[lo:lo + B].contiguous(), B, sums, counts)
    if sums is None:
        return None, None
    means = torch.empty_like(sums)
    F.check(F.lib().gsl_class_means(F.ptr(sums), F.ptr(counts), sums.shape[0], sums.shape[1], F.ptr(means), F.cur_stream()), "gsl_class_means")
    return means, counts


def calculate_prototypes(backbone, dataset, batch_size=32, device="cuda", aug_num=0):
    """Same contract as util.utils.calculate_prototypes (util/utils.py:502-549): {label (int): mean 512-d embedding (CPU tensor)} over `dataset`,
    eval mode; aug_num > 0 = 20 RandAugment(2, aug_num) passes over the dataset.  The embeddings do not depend on the batch size and the class
    sums are taken in dataset order whatever it is, so batches are at least as large as the engine's current capacity."""
    from torch.utils.data import ConcatDataset, DataLoader
    if aug_num == 0:
        repeated = dataset
    else:
        import torchvision.transforms as transforms
        tf = transforms.Compose([transforms.RandAugment(num_ops=2, magnitude=aug_num), transforms.ToTensor()])
        dataset.transform = tf
        repeated = ConcatDataset([dataset] * 20)
        repeated.transform = tf
    m = _unwrap(backbone)
    eng = getattr(m, "_engine", None)
    bs = max(int(batch_size), int(eng.max_batch) if eng is not None else 0)
    loader = DataLoader(repeated, batch_size=bs, shuffle=False)
    means, counts = class_prototype_table(backbone, loader, device)
    if means is None:
        return {}
    means_h, counts_h = means.cpu(), counts.cpu()             # the one device-to-host copy
    return {int(c): means_h[int(c)].clone() for c in torch.nonzero(counts_h > 0).flatten().tolist()}
