"""utils/lora_modules.py of the reference -> fused sm_100a kernels (same four names, same signatures)."""
from aqualora_b200.lora_modules import (CustomLoRACompatibleConvforward, CustomLoRACompatibleLinearforward,  # noqa: F401
                                        CustomLoRAConv2dLayerforward, CustomLoRALinearLayerforward)

try:  # text-encoder LoRA patching stays diffusers' own (out of scope: utils/lora_modules.py:65-146)
    from diffusers.loaders import LoraLoaderMixin as CustomLoraLoaderMixin  # noqa: F401
except ImportError:  # diffusers absent: the name still resolves for scripts that only import it
    class CustomLoraLoaderMixin:  # type: ignore
        pass
