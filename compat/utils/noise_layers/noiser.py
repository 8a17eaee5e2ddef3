"""utils/noise_layers/noiser.py of the reference -> CUDA-backed Noiser / distorsion_unit."""
from aqualora_b200.noise_layers import Noiser, distorsion_unit  # noqa: F401
