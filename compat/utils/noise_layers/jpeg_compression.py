from aqualora_b200.noise_layers import JpegCompression  # noqa: F401
