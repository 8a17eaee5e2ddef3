from aqualora_b200.noise_layers import Identity  # noqa: F401
