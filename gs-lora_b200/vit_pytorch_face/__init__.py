"""Drop-in for the reference package `vit_pytorch_face` (its __init__.py:1-3 exports ViT_face, ViT_face_low, ViT_face_up,
ViTs_face, ModifiedViT).  `ViT_face` and `ModifiedViT` are the gslora-b200 engine-backed models; the other three are outside the GS-LoRA hot
path (SURVEY.md section 2 rows 6, 15 and the LIRF halves) and are re-exported from the reference's own files when the reference tree is
importable (the unmodified driver builds all backbones eagerly, train_own_forget_cl.py:206-242)."""
import importlib.util
import os
import sys

from .vit_face import ViT_face, CosFace  # noqa: F401
from .modified_VIT import ModifiedViT  # noqa: F401


def _reference_dir():
    roots = [os.environ.get("GSLORA_REFERENCE_ROOT", "")] + list(sys.path)
    here = os.path.dirname(os.path.abspath(__file__))
    for r in roots:
        if not r:
            continue
        d = os.path.join(r, "vit_pytorch_face")
        if os.path.isfile(os.path.join(d, "vits_face.py")) and os.path.abspath(d) != here:
            return d
    return None


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _missing(name):
    class _Missing:
        def __init__(self, *a, **k):
            raise NotImplementedError(f"gslora-b200: {name} is outside the GS-LoRA hot path; put the reference tree on sys.path "
                                      "(or set GSLORA_REFERENCE_ROOT) to use the reference's own implementation")
    _Missing.__name__ = name
    return _Missing


_ref = _reference_dir()
try:
    if _ref is None:
        raise ImportError
    _rvf = _load(os.path.join(_ref, "vit_face.py"), "_gslora_ref_vit_face")
    ViT_face_low, ViT_face_up = _rvf.ViT_face_low, _rvf.ViT_face_up
    ViTs_face = _load(os.path.join(_ref, "vits_face.py"), "_gslora_ref_vits_face").ViTs_face
except Exception:  # reference not importable here (e.g. on the GPU box)
    ViT_face_low, ViT_face_up = _missing("ViT_face_low"), _missing("ViT_face_up")
    ViTs_face = _missing("ViTs_face")
