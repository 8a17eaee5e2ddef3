"""utils/models.py of the reference -> CUDA-backed drop-ins with the same constructors and state-dict keys."""
from aqualora_b200.decoder import SecretDecoder  # noqa: F401
from aqualora_b200.models import MapperNet, SecretEncoder  # noqa: F401
