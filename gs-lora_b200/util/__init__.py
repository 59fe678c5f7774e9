"""Drop-in overlay for the reference's `util` package: gslora-b200 provides `util.cal_norm`; every other submodule
(`util.utils`, `util.args`, `util.data_prefetcher`, ...) resolves to the reference's own file when the reference tree is on
sys.path (host-side orchestration the hot path does not touch)."""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
for _r in [os.environ.get("GSLORA_REFERENCE_ROOT", "")] + list(sys.path):
    if not _r:
        continue
    _d = os.path.join(_r, "util")
    if os.path.isfile(os.path.join(_d, "args.py")) and os.path.abspath(_d) != _here and _d not in __path__:
        __path__.append(_d)
        break
