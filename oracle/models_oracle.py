"""CPU oracle for hot path (ii), model part: SecretEncoder and SecretDecoder.  TEST INFRASTRUCTURE ONLY.

Functional (state-dict driven) restatements in plain PyTorch so every arithmetic step is visible; each function cites
the reference lines it follows.

Parity status
  * secret_encoder_*: PINNED -- tests/golden/models_small.pt holds outputs of the reference's own
    utils/models.py:SecretEncoder (tools/gen_golden.py); tests/test_oracle_golden.py compares.
  * efficientnet_b1_forward / secret_decoder_forward: the reference's decoder IS torchvision's `efficientnet_b1`
    (utils/models.py:88; torchvision==0.15.2 pinned in requirements.txt:30, 0.26 installed here -- same architecture and
    state-dict layout).  The restatement is pinned against the installed torchvision module on seeded random weights
    in tests/test_oracle_golden.py (the ImageNet weights the reference starts from cannot be downloaded offline).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# SecretEncoder  (utils/models.py:51-81)
# ----------------------------------------------------------------------------------------------


def secret_encoder_encode(msg, sd, base_res=32, resolution=64):
    """utils/models.py:57-64,70-72: Linear -> SiLU -> View(-1,1,b,b) -> Repeat(4,1,1) -> Upsample(nearest) -> Conv3x3."""
    h = F.silu(F.linear(msg, sd["secret_scaler.0.weight"], sd["secret_scaler.0.bias"]))
    h = h.view(-1, 1, base_res, base_res).repeat(1, 4, 1, 1)
    f = resolution // base_res
    h = F.interpolate(h, scale_factor=(f, f), mode="nearest")
    return F.conv2d(h, sd["secret_scaler.5.weight"], sd["secret_scaler.5.bias"], padding=1)


def secret_encoder_forward(x, msg, sd, base_res=32, resolution=64):
    """utils/models.py:74-81: c = bilinear(encode(msg), x.shape[2:]); returns (x + c, c)."""
    c = secret_encoder_encode(msg, sd, base_res, resolution)
    c = F.interpolate(c, size=(x.shape[2], x.shape[3]), mode="bilinear")
    return x + c, c


# ----------------------------------------------------------------------------------------------
# EfficientNet-B1 (torchvision.models.efficientnet: MBConv / SqueezeExcitation / Conv2dNormActivation), eval mode
# ----------------------------------------------------------------------------------------------
# (expand_ratio, kernel, stride, in_ch, out_ch, layers) after B1's width 1.0 / depth 1.1 scaling
B1_STAGES = [
    (1, 3, 1, 32, 16, 2),
    (6, 3, 2, 16, 24, 3),
    (6, 5, 2, 24, 40, 3),
    (6, 3, 2, 40, 80, 4),
    (6, 5, 1, 80, 112, 4),
    (6, 5, 2, 112, 192, 5),
    (6, 3, 1, 192, 320, 2),
]
BN_EPS = 1e-5


def _bn(x, sd, p, eps=BN_EPS):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.0, eps)


def _conv_bn_act(x, sd, p, stride=1, groups=1, act=True):
    w = sd[p + ".0.weight"]
    pad = (w.shape[-1] - 1) // 2
    x = _bn(F.conv2d(x, w, None, stride, pad, 1, groups), sd, p + ".1")
    return F.silu(x) if act else x


def _mbconv(x, sd, p, expand, k, stride, cin, cout):
    """torchvision MBConv.forward (eval: StochasticDepth is the identity)."""
    inp = x
    i = 0
    if expand != 1:
        x = _conv_bn_act(x, sd, f"{p}.block.{i}")           # 1x1 expand + BN + SiLU
        i += 1
    cexp = cin * expand
    x = _conv_bn_act(x, sd, f"{p}.block.{i}", stride=stride, groups=cexp)   # depthwise + BN + SiLU
    i += 1
    # SqueezeExcitation: scale = sigmoid(fc2(silu(fc1(avgpool(x)))))
    s = x.mean(dim=(2, 3), keepdim=True)
    s = F.silu(F.conv2d(s, sd[f"{p}.block.{i}.fc1.weight"], sd[f"{p}.block.{i}.fc1.bias"]))
    s = torch.sigmoid(F.conv2d(s, sd[f"{p}.block.{i}.fc2.weight"], sd[f"{p}.block.{i}.fc2.bias"]))
    x = x * s
    i += 1
    x = _conv_bn_act(x, sd, f"{p}.block.{i}", act=False)    # 1x1 project + BN
    if stride == 1 and cin == cout:
        x = x + inp
    return x


def efficientnet_b1_features(x, sd, prefix=""):
    x = _conv_bn_act(x, sd, prefix + "features.0", stride=2)                     # stem 3x3 s2
    for si, (expand, k, stride, cin, cout, layers) in enumerate(B1_STAGES):
        for li in range(layers):
            x = _mbconv(x, sd, f"{prefix}features.{si + 1}.{li}", expand, k, stride if li == 0 else 1,
                        cin if li == 0 else cout, cout)
    return _conv_bn_act(x, sd, prefix + "features.8")                            # head 1x1 320 -> 1280


def efficientnet_b1_forward(x, sd, prefix=""):
    x = efficientnet_b1_features(x, sd, prefix)
    x = x.mean(dim=(2, 3))                                                       # avgpool + flatten
    return F.linear(x, sd[prefix + "classifier.1.weight"], sd[prefix + "classifier.1.bias"])   # Dropout is identity in eval


def secret_decoder_forward(x, sd, output_size, prefix="model."):
    """utils/models.py:91-96 (same as evaluation/utils_eval.py:149-154): bilinear to 512x512, EfficientNet-B1 with a
    Linear(1280, 2*bits) head, view(-1, bits, 2)."""
    x = F.interpolate(x, size=(512, 512), mode="bilinear")
    return efficientnet_b1_forward(x, sd, prefix).view(-1, output_size, 2)


def decode_bits(logits):
    """evaluation/utils_eval.py:194-198: bit = argmax over the last axis."""
    return logits.argmax(dim=-1)


def make_decoder_state(output_size=48, seed=0):
    """Seeded random decoder weights in the reference's msgdecoder.pt layout (keys under `model.`), with non-trivial
    BN statistics so the BN folding of the CUDA path is exercised."""
    import torchvision.models.efficientnet as efficientnet

    torch.manual_seed(seed)
    m = efficientnet.efficientnet_b1(weights=None)
    m.classifier[1] = torch.nn.Linear(m.classifier[1].in_features, output_size * 2, bias=True)
    g = torch.Generator().manual_seed(seed + 1)
    sd = m.state_dict()
    for k, v in sd.items():
        if k.endswith("running_mean"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
        elif k.endswith("running_var"):
            v.copy_(torch.rand(v.shape, generator=g) * 0.5 + 0.75)
        elif k.endswith(".1.weight") and v.dim() == 1:
            v.copy_(torch.rand(v.shape, generator=g) * 0.5 + 0.75)
        elif k.endswith(".1.bias") and v.dim() == 1 and "classifier" not in k:
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
    return {"model." + k: v.clone() for k, v in sd.items()}, m


# ----------------------------------------------------------------------------------------------
# bit accuracy / detection threshold  (evaluation/utils_eval.py:131-140, 193-211)
# ----------------------------------------------------------------------------------------------
def calculate_fpr(tau: int, k: int) -> float:
    """evaluation/utils_eval.py:131-134: P[Binomial(k, 1/2) > tau]."""
    from math import comb

    sum_combinations = sum(comb(k, i) for i in range(tau + 1, k + 1))
    return 1 / (2 ** k) * sum_combinations


def get_threshold(k: int, fpr: float) -> int:
    """evaluation/utils_eval.py:136-140: smallest tau whose false-positive rate is <= fpr."""
    tau = 0
    while calculate_fpr(tau, k) > fpr:
        tau += 1
    return tau


def bit_accuracy(bits, msg_gt):
    """evaluation/utils_eval.py:202-204: fraction of matching bits per image."""
    return (bits == msg_gt).float().mean(dim=-1)
