"""The C-ABI boundary on CPU: libaqualora_b200.so builds for sm_100a without a GPU, loads, and exports exactly the
entry points include/aqualora_b200.h declares.  No compute call is made here (those are the -m gpu tests)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from aqualora_b200 import _lib, build

    build.build()
    return _lib.load()


def _declared():
    text = open(os.path.join(ROOT, "include", "aqualora_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(aq_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(lib):
    from aqualora_b200 import _lib

    declared = _declared()
    assert len(declared) >= 15
    assert declared == _lib.exported_symbols()


def test_every_declared_symbol_is_exported(lib):
    for name in _declared():
        assert isinstance(getattr(lib, name), ctypes._CFuncPtr), name


def test_version_arch_and_error_string(lib):
    assert lib.aq_version() >= 1
    assert lib.aq_arch() == 100
    assert isinstance(lib.aq_last_error(), bytes)


def test_argument_validation_needs_no_gpu(lib):
    """Shape checks run before any CUDA call, so bad arguments are rejected identically with and without a device."""
    from aqualora_b200 import _lib

    rc = lib.aq_mapper_fwd(None, None, None, 0, 48, 64, 0, None)
    assert rc == -1 and b"bad shape" in lib.aq_last_error()
    rc = lib.aq_flat_sumsq(None, 0, None, None)
    assert rc == -1
    assert _lib.last_error()


def test_product_refuses_cpu_tensors(lib):
    """No CPU fallback: the drop-in forwards raise on CPU tensors instead of computing in PyTorch."""
    import torch

    from aqualora_b200 import lora_modules
    from aqualora_b200._lib import AqualoraError

    lin = lora_modules.LoRACompatibleLinear(16, 16).to(torch.bfloat16)
    with pytest.raises(AqualoraError):
        lin(torch.zeros(1, 4, 16, dtype=torch.bfloat16))


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under aqualora_b200/ may import it."""
    pkg = os.path.join(ROOT, "aqualora_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)


def test_decoder_container_matches_torchvision_layout(lib):
    """SecretDecoder keeps the reference checkpoint's names (torchvision efficientnet_b1 under `model.`), and the Python
    packer and the C walker agree on the packed size."""
    import torch

    from aqualora_b200.decoder import SecretDecoder, pack_state_dict
    from oracle import models_oracle as MO

    sd, module = MO.make_decoder_state(48, seed=0)
    dec = SecretDecoder(48)
    assert list(dec.state_dict().keys()) == list(sd.keys())
    assert all(dec.state_dict()[k].shape == v.shape for k, v in sd.items())
    packed = pack_state_dict(sd, 96)
    assert packed.dtype == torch.float32 and packed.numel() == lib.aq_effnetb1_packed_floats(96)
    assert lib.aq_effnetb1_workspace_bytes(2) > 2 * 6291456 * 4


def test_compat_shim_exports_reference_names(lib):
    """compat/utils mirrors the module paths the reference scripts import (train/ppft_train.py:49-57,
    train/latent_wm_pretrain.py:36)."""
    import importlib
    import sys

    sys.path.insert(0, os.path.join(ROOT, "compat"))
    saved = {k: v for k, v in sys.modules.items() if k == "utils" or k.startswith("utils.")}
    for k in saved:
        del sys.modules[k]
    try:
        lm = importlib.import_module("utils.lora_modules")
        for name in ("CustomLoraLoaderMixin", "CustomLoRAConv2dLayerforward", "CustomLoRACompatibleConvforward",
                     "CustomLoRALinearLayerforward", "CustomLoRACompatibleLinearforward"):
            assert hasattr(lm, name)
        import inspect

        assert list(inspect.signature(lm.CustomLoRALinearLayerforward).parameters) == ["self", "hidden_states", "scale"]
        m = importlib.import_module("utils.models")
        assert all(hasattr(m, n) for n in ("SecretEncoder", "SecretDecoder", "MapperNet"))
        n = importlib.import_module("utils.noise_layers.noiser")
        assert hasattr(n, "Noiser") and hasattr(n, "distorsion_unit")
    finally:
        sys.path.remove(os.path.join(ROOT, "compat"))
        for k in [k for k in sys.modules if k == "utils" or k.startswith("utils.")]:
            del sys.modules[k]
        sys.modules.update(saved)
