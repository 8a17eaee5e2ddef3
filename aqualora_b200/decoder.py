"""Drop-in `SecretDecoder` (utils/models.py:84-96, duplicated at evaluation/utils_eval.py:142-154).

The reference wraps torchvision's `efficientnet_b1` and replaces its classifier by `Linear(1280, 2 * bits)`; checkpoints
(`msgdecoder.pt`, the `sec_decoder` entry of the pretrain checkpoint) are that module's state-dict under `model.`.
This class keeps exactly those parameter / buffer names (so `load_state_dict` of a reference checkpoint works unchanged)
but owns no torchvision code.  eval(): the fp32 NHWC kernel chain in csrc/decoder.cu with every BatchNorm folded into its
convolution (PPFT and evaluation run the decoder in eval(): train/ppft_train.py:974, evaluation/utils_eval.py:168).
train() (train/latent_wm_pretrain.py:160-217, rob_enhance_finetune.py:995-1040): batch-statistics BatchNorm, StochasticDepth("row")
and Dropout(0.2) as torchvision's `efficientnet_b1` defines them, differentiable w.r.t. the image and every parameter, on this
repository's kernels: stem / depthwise / pointwise convolutions with their input and weight gradients (csrc/decoder_train.cu,
pointwise forward + input gradient on the tensor cores through csrc/decoder_pw.cu), BatchNorm(+SiLU) forward / backward
(csrc/bn_train.cu).  What stays PyTorch glue: the squeeze-excitation MLP on [B, C] vectors, the average pools, the row masks of
StochasticDepth / Dropout and the final Linear (each a few kFLOP per image).  CPU tensors raise in both modes: no CPU fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from ._lib import AqualoraError

# (expand_ratio, kernel, stride, in_ch, out_ch, layers): torchvision efficientnet_b1 (width 1.0, depth 1.1)
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


def _conv_bn(cin, cout, k, stride=1, groups=1):
    """torchvision Conv2dNormActivation: [0] conv (no bias), [1] BatchNorm2d, [2] activation (no parameters)."""
    return nn.Sequential(nn.Conv2d(cin, cout, k, stride, (k - 1) // 2, groups=groups, bias=False), nn.BatchNorm2d(cout, eps=BN_EPS))


class _SE(nn.Module):
    def __init__(self, c, sq):
        super().__init__()
        self.fc1 = nn.Conv2d(c, sq, 1)
        self.fc2 = nn.Conv2d(sq, c, 1)


class _MBConv(nn.Module):
    def __init__(self, cin, cout, expand, k, stride):
        super().__init__()
        cexp = cin * expand
        layers = []
        if expand != 1:
            layers.append(_conv_bn(cin, cexp, 1))
        layers.append(_conv_bn(cexp, cexp, k, stride, groups=cexp))
        layers.append(_SE(cexp, max(1, cin // 4)))
        layers.append(_conv_bn(cexp, cout, 1))
        self.block = nn.Sequential(*layers)


class _EfficientNetB1Params(nn.Module):
    """Parameter container with torchvision's module paths: features.{0..8}, classifier.{0,1}."""

    def __init__(self, out_features):
        super().__init__()
        feats = [_conv_bn(3, 32, 3, 2)]
        for expand, k, stride, cin, cout, layers in B1_STAGES:
            feats.append(nn.Sequential(*[_MBConv(cin if i == 0 else cout, cout, expand, k, stride if i == 0 else 1) for i in range(layers)]))
        feats.append(_conv_bn(320, 1280, 1))
        self.features = nn.Sequential(*feats)
        self.classifier = nn.Sequential(nn.Dropout(0.2), nn.Linear(1280, out_features))


class _BnActFn(torch.autograd.Function):
    """Batch-statistics BatchNorm2d (+ SiLU) on the kernels of csrc/bn_train.cu; saves only the convolution output and (mean, rstd)."""

    @staticmethod
    def forward(ctx, z, gamma, beta, running_mean, running_var, eps, momentum, act):
        z = z.contiguous(memory_format=torch.channels_last)
        y, mean_rstd = ops.bn_train_fwd(z, gamma.detach().contiguous(), beta.detach().contiguous(), running_mean, running_var, eps, momentum, act)
        ctx.save_for_backward(z, gamma, beta, mean_rstd)
        ctx.act = act
        return y

    @staticmethod
    def backward(ctx, gy):
        z, gamma, beta, mean_rstd = ctx.saved_tensors
        gz, g_gamma, g_beta = ops.bn_train_bwd(gy.contiguous(memory_format=torch.channels_last), z, gamma.detach().contiguous(),
                                               beta.detach().contiguous(), mean_rstd, ctx.act)
        return gz, g_gamma, g_beta, None, None, None, None, None


class _DwConvFn(torch.autograd.Function):
    """Depthwise k x k convolution of the train path (csrc/decoder_train.cu)."""

    @staticmethod
    def forward(ctx, x, w, stride):
        x = x.contiguous(memory_format=torch.channels_last)
        ctx.save_for_backward(x, w)
        ctx.stride = stride
        return ops.dwconv_fwd(x, w, stride)

    @staticmethod
    def backward(ctx, gz):
        x, w = ctx.saved_tensors
        gx, gw = ops.dwconv_bwd(gz.contiguous(memory_format=torch.channels_last), x, w, ctx.stride, ctx.needs_input_grad[0])
        return gx, gw, None


class _PwConvFn(torch.autograd.Function):
    """Pointwise convolution y = (x (.) se) W^T of the train path: forward and input gradient on the tensor cores (3 x TF32,
    csrc/decoder_pw.cu), weight gradient on csrc/decoder_train.cu.  `se` [B, K] is the squeeze-excitation gate of the project
    convolution (torchvision MBConv: block = [..., SE, project]), applied while the A tile is staged."""

    @staticmethod
    def forward(ctx, x, w, se):
        x = x.contiguous(memory_format=torch.channels_last)
        B, K, H, W = x.shape
        N = w.shape[0]
        x2d = x.permute(0, 2, 3, 1).reshape(B * H * W, K)
        w2d = w.reshape(N, K)
        zero_n = torch.zeros(N, dtype=torch.float32, device=x.device)
        se2d = None if se is None else se.reshape(B, K).contiguous()
        y2d = ops.conv1x1_tf32x3(x2d, w2d, zero_n, se=se2d, hw=H * W)
        ctx.save_for_backward(x, w, se2d)
        return y2d.view(B, H, W, N).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, gy):
        x, w, se2d = ctx.saved_tensors
        B, K, H, W = x.shape
        N = w.shape[0]
        gy2d = gy.contiguous(memory_format=torch.channels_last).permute(0, 2, 3, 1).reshape(B * H * W, N)
        x2d = x.permute(0, 2, 3, 1).reshape(B * H * W, K)
        gx = g_se = None
        if ctx.needs_input_grad[0] or (se2d is not None and ctx.needs_input_grad[2]):
            wt = w.reshape(N, K).t().contiguous()
            ge2d = ops.conv1x1_tf32x3(gy2d, wt, torch.zeros(K, dtype=torch.float32, device=x.device), hw=H * W)   # d / d (x (.) se)
            ge = ge2d.view(B, H * W, K)
            if se2d is None:
                gx = ge2d.view(B, H, W, K).permute(0, 3, 1, 2)
            else:
                g_se = (ge * x2d.view(B, H * W, K)).sum(dim=1).view(B, K, 1, 1)
                gx = (ge * se2d.view(B, 1, K)).view(B, H, W, K).permute(0, 3, 1, 2)
        gw = ops.conv1x1_wgrad(gy2d, x2d, se2d, H * W).view_as(w) if ctx.needs_input_grad[1] else None
        return gx, gw, g_se


class _StemConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w):
        x = x.contiguous()
        ctx.save_for_backward(x, w)
        return ops.stem_conv_fwd(x, w)

    @staticmethod
    def backward(ctx, gz):
        x, w = ctx.saved_tensors
        gx, gw = ops.stem_conv_bwd(gz.contiguous(memory_format=torch.channels_last), x, w, ctx.needs_input_grad[0])
        return gx, gw


def _train_conv(conv: nn.Conv2d, x: torch.Tensor, se=None) -> torch.Tensor:
    """One bias-free convolution of the train path on this repository's kernels; shapes outside what they cover (a 512 x 512 decoder
    never produces one) fall to the library convolution."""
    w = conv.weight
    k = w.shape[-1]
    if conv.groups == 1 and k == 1 and x.shape[1] % 8 == 0 and w.shape[0] % 4 == 0 and (x.shape[2] * x.shape[3]) % 128 == 0:
        return _PwConvFn.apply(x, w, se)
    if se is not None:
        x = x * se
    if conv.groups == conv.in_channels and conv.groups > 1 and k in (3, 5) and conv.stride[0] in (1, 2) and x.shape[1] % 4 == 0:
        return _DwConvFn.apply(x, w, conv.stride[0])
    if conv.groups == 1 and k == 3 and conv.stride[0] == 2 and tuple(w.shape[:2]) == (32, 3):
        return _StemConvFn.apply(x, w)
    return conv(x)


def _fold(sd, conv_key, bn_key):
    w = sd[conv_key + ".weight"].double()
    g, b = sd[bn_key + ".weight"].double(), sd[bn_key + ".bias"].double()
    mu, var = sd[bn_key + ".running_mean"].double(), sd[bn_key + ".running_var"].double()
    s = g / torch.sqrt(var + BN_EPS)
    return (w * s.view(-1, 1, 1, 1)).float(), (b - mu * s).float()


def _split_tf32(w: torch.Tensor):
    """w = hi + lo exactly, hi = w with the low 13 mantissa bits cleared (a TF32 number): the operands of the 3-term
    tensor-core product in csrc/decoder_pw.cu."""
    w = w.contiguous().float()
    hi = (w.view(torch.int32) & -8192).view(torch.float32)          # -8192 == 0xFFFFE000
    return hi, w - hi


def pack_state_dict(sd: dict, out_features: int, prefix: str = "model.") -> torch.Tensor:
    """Fold every BatchNorm (eval) into its convolution and lay the fp32 parameters out in execution order, each segment
    padded to a multiple of 4 floats -- the layout csrc/decoder.cu walks:
        stem  w [(ky, kx, ci), 32], b [32]
        block [expand: W_hi [cexp, cin], W_lo [cexp, cin], b] , dw w [k*k, cexp], b , se w1 [sq, cexp], b1, w2^T [sq, cexp], b2 ,
              project W_hi [cout, cexp], W_lo [cout, cexp], b
        head  W_hi [1280, 320], W_lo, b ; fc w [out, 1280], b
    The fold is computed in float64 and rounded once; pointwise weights are stored as an exact hi + lo TF32 split."""
    out = []

    def put(t):
        t = t.reshape(-1).float()
        padn = (-t.numel()) % 4
        out.append(torch.cat([t, t.new_zeros(padn)]) if padn else t)

    def put_pointwise(w, b):
        hi, lo = _split_tf32(w.flatten(1))                              # [cout, cin], K-major like the conv weight itself
        put(hi); put(lo); put(b)

    p = prefix + "features."
    w, b = _fold(sd, p + "0.0", p + "0.1")
    put(w.permute(2, 3, 1, 0)); put(b)
    for si, (expand, k, stride, cin, cout, layers) in enumerate(B1_STAGES):
        for li in range(layers):
            q = f"{p}{si + 1}.{li}.block."
            i = 0
            if expand != 1:
                put_pointwise(*_fold(sd, f"{q}{i}.0", f"{q}{i}.1"))
                i += 1
            w, b = _fold(sd, f"{q}{i}.0", f"{q}{i}.1")
            put(w.flatten(1).t()); put(b)                     # [cexp, 1, k, k] -> [k*k, cexp]
            i += 1
            put(sd[f"{q}{i}.fc1.weight"].flatten(1)); put(sd[f"{q}{i}.fc1.bias"])
            put(sd[f"{q}{i}.fc2.weight"].flatten(1).t()); put(sd[f"{q}{i}.fc2.bias"])       # [cexp, sq] -> [sq, cexp]
            i += 1
            put_pointwise(*_fold(sd, f"{q}{i}.0", f"{q}{i}.1"))
    put_pointwise(*_fold(sd, p + "8.0", p + "8.1"))
    put(sd[prefix + "classifier.1.weight"]); put(sd[prefix + "classifier.1.bias"])
    return torch.cat(out).contiguous()


class SecretDecoder(nn.Module):
    """`SecretDecoder(output_size)`: logits [B, output_size, 2] for images [B, 3, H, W] in [-1, 1] (resized to 512 x 512 with the
    reference's bilinear interpolate when needed).  `decode_bits(x)` returns the argmax bits as uint8."""

    def __init__(self, output_size=64):
        super().__init__()
        self.output_size = output_size
        self.model = _EfficientNetB1Params(output_size * 2)
        self._packed = None
        self._packed_key = None

    def _packed_weights(self, device):
        key = (str(device),) + tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers()))
        if self._packed is None or self._packed_key != key:
            sd = {k: v.detach() for k, v in self.state_dict().items()}
            self._packed = pack_state_dict(sd, self.output_size * 2).to(device)
            self._packed_key = key
        return self._packed

    # -- train mode: torchvision's EfficientNet-B1 dataflow on the parameter container (SURVEY.md 8(a) "MBConv dataflow") ---------
    stochastic_depth_prob = 0.2        # torchvision efficientnet_b1 default
    dropout_p = 0.2                    # classifier[0]

    def _train_forward(self, x):
        m = self.model
        total_blocks = float(sum(layers for *_, layers in B1_STAGES))

        def cba(seq, x, act=True, se=None):
            # conv (no bias: csrc/decoder_train.cu / decoder_pw.cu) -> BatchNorm2d with batch statistics + SiLU (csrc/bn_train.cu)
            bn = seq[1]
            z = _train_conv(seq[0], x, se)
            if z.shape[1] % 4 != 0 or bn.momentum is None:
                y = bn(z)
                return F.silu(y) if act else y
            y = _BnActFn.apply(z, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, bn.momentum, act)
            bn.num_batches_tracked += 1
            return y

        x = cba(m.features[0], x)
        block_id = 0
        for si, (expand, k, stride, cin, cout, layers) in enumerate(B1_STAGES):
            for li in range(layers):
                blk = m.features[si + 1][li].block
                h, i = x, 0
                if expand != 1:
                    h = cba(blk[i], h)
                    i += 1
                h = cba(blk[i], h)
                i += 1
                se = blk[i]
                scale = torch.sigmoid(se.fc2(F.silu(se.fc1(F.adaptive_avg_pool2d(h, 1)))))      # [B, C, 1, 1]: a few kFLOP per image
                i += 1
                h = cba(blk[i], h, act=False, se=scale)      # the gate rides in the project convolution's A-tile staging
                if (stride if li == 0 else 1) == 1 and (cin if li == 0 else cout) == cout:
                    p = self.stochastic_depth_prob * block_id / total_blocks     # StochasticDepth(p, "row")
                    if p > 0.0:
                        keep = torch.empty(h.shape[0], 1, 1, 1, dtype=h.dtype, device=h.device).bernoulli_(1.0 - p) / (1.0 - p)
                        h = h * keep
                    h = h + x
                x = h
                block_id += 1
        x = cba(m.features[8], x)
        x = F.adaptive_avg_pool2d(x, 1).flatten(1)
        x = F.dropout(x, self.dropout_p, training=True)
        return m.classifier[1](x)

    def _run(self, x, want_bits):
        if not x.is_cuda:
            raise AqualoraError("SecretDecoder: CPU tensor passed; aqualora_b200 has no CPU fallback")
        if self.training:
            x = x.float()
            if tuple(x.shape[-2:]) != (512, 512):
                from .noise_layers import crop_resize

                H, W = x.shape[-2:]
                x = crop_resize(x, 0, 0, H, W, 512, 512, (512, 512))         # F.interpolate(x, (512, 512), 'bilinear'), differentiable
            logits = self._train_forward(x.contiguous(memory_format=torch.channels_last))
            bits = logits.view(-1, self.output_size, 2).argmax(-1).to(torch.uint8) if want_bits else None
            return logits.view(-1, self.output_size, 2), bits
        x = x.float()
        if tuple(x.shape[-2:]) != (512, 512):
            H, W = x.shape[-2:]
            x = ops.noise_crop_resize(x, 0, 0, H, W, 512, 512, (512, 512))     # F.interpolate(x, (512, 512), 'bilinear')
        with torch.no_grad():
            logits, bits = ops.effnetb1_fwd(x, self._packed_weights(x.device), self.output_size * 2, want_bits)
        return logits.view(-1, self.output_size, 2), bits

    def forward(self, x):
        return self._run(x, False)[0]

    def decode_bits(self, x):
        return self._run(x, True)[1]
