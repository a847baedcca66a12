"""Drop-in `wan` package (surface of ZulutionAI/MoviiGen1.1's wan/): scripts/inference/generate.py runs against it
unchanged with PYTHONPATH=moviigen1.1_b200.  Only the DiT sampling hot path and the WanVAE decoder are implemented
(B200-native); see DESIGN.md for what is deliberately out of scope."""
from . import configs, distributed, modules  # noqa: F401
from .text2video import WanT2V  # noqa: F401
