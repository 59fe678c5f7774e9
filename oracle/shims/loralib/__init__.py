"""`import loralib as lora` for the unmodified reference -> oracle restatement of loralib 0.1.2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.loralib_restated import Linear, MergedLinear, LoRALayer, mark_only_lora_as_trainable  # noqa: E402,F401
