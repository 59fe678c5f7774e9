"""Drop-in overlay for the reference's `image_iter` (the drivers do `from image_iter import CLDatasetWrapper, CustomSubset, ImageNet900Dataset`,
train/train_own_forget_cl.py:13).  Dataset plumbing is the reference's own code (out of the hot path): when the
reference tree is importable its `image_iter.py` is executed into this module unchanged.  Two host-side landmines of that file on this software
stack are defused (landmine 2):
  * `CustomSubset` overrides `__getitem__` but not `__getitems__`; torch >= 2.1 refuses to construct such a `Subset` subclass
    (image_iter.py:124-137) -- the batched accessor is added, with the same per-index semantics;
  * `import mxnet` (only used by the .rec face loaders, never by the GS-LoRA runs) is satisfied by an empty stand-in when mxnet is absent.
"""
import os
import sys
import types

from torch.utils.data import Subset

_here = os.path.dirname(os.path.abspath(__file__))
_ref_file = None
for _r in [os.environ.get("GSLORA_REFERENCE_ROOT", "")] + list(sys.path):
    _f = os.path.join(_r, "image_iter.py") if _r else ""
    if _f and os.path.isfile(_f) and os.path.abspath(_r) != _here:
        _ref_file = _f
        break

_ref_loaded = False
if _ref_file is not None:
    try:
        import mxnet  # noqa: F401
    except Exception:
        _mx = types.ModuleType("mxnet")
        for _sub in ("ndarray", "io", "recordio"):
            _m = types.ModuleType("mxnet." + _sub)
            setattr(_mx, _sub, _m)
            sys.modules.setdefault("mxnet." + _sub, _m)
        sys.modules.setdefault("mxnet", _mx)
    try:
        with open(_ref_file) as _fh:
            exec(compile(_fh.read(), _ref_file, "exec"), globals())
        _ref_loaded = True
    except Exception as _e:        # e.g. cv2 missing: fall back to the one class the GS-LoRA path needs
        _ref_error = _e

if not _ref_loaded:
    class CustomSubset(Subset):
        """Subset that keeps the parent's `targets` / `classes` (image_iter.py:124-137)."""

        def __init__(self, dataset, indices):
            super().__init__(dataset, indices)
            self.targets = dataset.targets
            self.classes = dataset.classes


def _getitems(self, indices):
    return [self.__getitem__(i) for i in indices]


if getattr(CustomSubset, "__getitems__", None) is Subset.__getitems__ and CustomSubset.__getitem__ is not Subset.__getitem__:  # noqa: F821
    CustomSubset.__getitems__ = _getitems  # noqa: F821
