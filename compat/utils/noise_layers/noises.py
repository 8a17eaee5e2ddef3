from aqualora_b200.noise_layers import ColorJitter, CropandResize, GaussianBlur, GaussianNoise, random_int  # noqa: F401
