"""Drop-in replacement for the reference's `utils/lora_modules.py` (the plugin boundary).

The four `Custom*forward` functions keep the reference's names, signatures and semantics
(utils/lora_modules.py:9-62) and are meant to be monkey-patched onto diffusers' `LoRACompatibleLinear/Conv` and
`LoRALinearLayer/LoRAConv2dLayer` exactly as `train/ppft_train.py:681-689` does:

    module.forward = types.MethodType(CustomLoRACompatibleLinearforward, module)

They only read the attributes the reference reads (`weight`, `bias`, `lora_layer`, `down.weight`, `up.weight`,
`network_alpha`, `rank`, conv `stride/padding/dilation/groups`).  Behind them sits ONE fused sm_100a kernel per
projection (csrc/lora_gemm.cu) plus the backward kernels; there is no PyTorch fallback: CPU tensors, missing
library or unsupported dtypes raise `AqualoraError`.

`diffusers` is not required: minimal container classes with the diffusers field names are defined here for the
in-repo U-Net harness (aqualora_b200/unet.py) and for tests.
"""
from __future__ import annotations

import contextlib
import os
import types
import weakref
from typing import Dict, Iterable, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from ._lib import AqualoraError

# ------------------------------------------------------------------------------------------------
# operand caches: bf16 (+ transposed) copies of fp32 master LoRA weights, W^T of frozen projections
# ------------------------------------------------------------------------------------------------
_PACK_CACHE: Dict[int, tuple] = {}
_WT_CACHE: Dict[int, tuple] = {}
_PACK_EPOCH = 0          # bumped by invalidate_packed(): kernels that update parameters in place bypass `_version`
_LORA_DISABLED = False


def invalidate_packed() -> None:
    """Mark every cached bf16 operand copy stale (call after an in-place optimizer step on the flat buffers)."""
    global _PACK_EPOCH
    _PACK_EPOCH += 1


@contextlib.contextmanager
def lora_disabled():
    """Run the patched modules as if `lora_layer is None` (utils/lora_modules.py:47-52,57-59).  Used for the PPFT
    clean pass, whose all-zero scale makes the LoRA branch contribute exactly 0 (train/ppft_train.py:1026-1029)."""
    global _LORA_DISABLED
    prev, _LORA_DISABLED = _LORA_DISABLED, True
    try:
        yield
    finally:
        _LORA_DISABLED = prev


def _packed(param: torch.Tensor, rows: int, cols: int):
    """(bf16 [rows, cols], bf16 [cols, rows]) of a fp32/bf16 parameter, refreshed when the parameter changes."""
    key = id(param)
    ver = (param._version, param.data_ptr(), _PACK_EPOCH)
    hit = _PACK_CACHE.get(key)
    if hit is not None and (hit[3]() is not param or hit[1].shape != (rows, cols)):
        hit = None                                   # id() of a collected tensor reused by another parameter
    if hit is not None and hit[0] == ver:
        return hit[1], hit[2]
    src = param.detach().reshape(rows, cols)
    if hit is not None:
        d, dt = hit[1], hit[2]                       # reuse the buffers (stable addresses, CUDA-graph friendly)
    else:
        d = torch.empty((rows, cols), dtype=torch.bfloat16, device=param.device)
        dt = torch.empty((cols, rows), dtype=torch.bfloat16, device=param.device)
    if src.dtype == torch.float32:
        ops.cast_transpose_bf16(src.contiguous(), d, dt)
    elif src.dtype == torch.bfloat16:
        d.copy_(src)
        dt.copy_(ops.transpose_bf16(src))
    else:
        raise AqualoraError(f"LoRA weights must be fp32 or bf16, got {src.dtype}")
    if len(_PACK_CACHE) > 4096:                      # temporaries (padded ranks) leave dead entries behind: drop them
        for k in [k for k, v in _PACK_CACHE.items() if v[3]() is None]:
            del _PACK_CACHE[k]
    _PACK_CACHE[key] = (ver, d, dt, weakref.ref(param))
    return d, dt


_BATCH_TABLES: dict = {}


def refresh_packed(params) -> None:
    """Refresh the bf16 (+ transposed) operand copies of many fp32 2-D parameters (the 384 LoRA matrices after an optimizer
    step) in ONE launch, and mark them fresh.  The first call packs them one by one (that allocates the stable buffers)."""
    params = list(params)
    key = tuple(id(p) for p in params)
    hit = _BATCH_TABLES.get(key)
    if hit is None or any(_PACK_CACHE.get(id(p)) is None or _PACK_CACHE[id(p)][3]() is not p for p in params):
        rows_cols = []
        for p in params:
            r = p.shape[0]
            c = p.numel() // r
            _packed(p, r, c)
            rows_cols.append((r, c))
        table, tiles = [], 0
        for p, (r, c) in zip(params, rows_cols):
            if p.dtype != torch.float32:
                raise AqualoraError("refresh_packed: fp32 master parameters expected")
            _, d, dt, _ = _PACK_CACHE[id(p)]
            tx = (c + 31) // 32
            table.append([p.data_ptr(), d.data_ptr(), dt.data_ptr(), r, c, tiles, tx])
            tiles += tx * ((r + 31) // 32)
        hit = (torch.tensor(table, dtype=torch.int64, device=params[0].device), tiles, [p.data_ptr() for p in params])
        _BATCH_TABLES[key] = hit
        return                                          # _packed() just refreshed every copy
    jobs, tiles, ptrs = hit
    if any(p.data_ptr() != q for p, q in zip(params, ptrs)):
        del _BATCH_TABLES[key]
        return refresh_packed(params)
    ops.cast_transpose_bf16_batched(jobs, tiles)
    for p in params:
        _, d, dt, ref = _PACK_CACHE[id(p)]
        _PACK_CACHE[id(p)] = ((p._version, p.data_ptr(), _PACK_EPOCH), d, dt, ref)


def _weight_t(weight: torch.Tensor, rows: int, cols: int) -> torch.Tensor:
    """W^T [cols, rows] bf16 of a frozen projection weight (needed by the dX contraction), cached."""
    key = id(weight)
    ver = (weight._version, weight.data_ptr())
    hit = _WT_CACHE.get(key)
    if hit is not None and hit[0] == ver and hit[2]() is weight and hit[1].shape == (cols, rows):
        return hit[1]
    wt = ops.transpose_bf16(weight.detach().reshape(rows, cols))
    _WT_CACHE[key] = (ver, wt, weakref.ref(weight))
    return wt


def clear_caches() -> None:
    _WGRAD_QUEUE.clear()
    _PACK_CACHE.clear()
    _WT_CACHE.clear()
    _BATCH_TABLES.clear()
    _W16_CACHE.clear()


# ------------------------------------------------------------------------------------------------
# deferred weight gradients: dUp / dDn of a layer feed nothing downstream of its backward, so layers that accumulate straight into
# a flat gradient buffer (`_aq_grad`) queue their contraction and several layers leave in ONE launch (aq_lora_wgrad_batch): 192
# launches of a few microseconds each against a 5 - 8 us launch floor become 12 per step (measured: 3.50 -> 2.33 ms of kernel time per
# step, profiles/r02_bench_wgrad_batch_ab.txt).
# ------------------------------------------------------------------------------------------------
_WGRAD_QUEUE: list = []
_WGRAD_FLUSH_SCHEDULED = False
WGRAD_BATCH = max(1, int(os.environ.get("AQ_WGRAD_BATCH", "32")))      # layers per launch group (1 = launch with every layer)


def flush_wgrad_queue() -> None:
    """Launch every queued weight-gradient job.  Runs by itself at the end of each backward pass (autograd engine callback) and when
    the queue is full; call it by hand only if gradients are read in the middle of a backward pass."""
    global _WGRAD_FLUSH_SCHEDULED
    _WGRAD_FLUSH_SCHEDULED = False
    if _WGRAD_QUEUE:
        jobs = list(_WGRAD_QUEUE)
        _WGRAD_QUEUE.clear()
        ops.lora_wgrad_batch(jobs)


def _enqueue_wgrad(job, deferrable: bool) -> None:
    global _WGRAD_FLUSH_SCHEDULED
    _WGRAD_QUEUE.append(job)
    if not deferrable or len(_WGRAD_QUEUE) >= WGRAD_BATCH:
        flush_wgrad_queue()          # gradients handed back to autograd must be complete in stream order: no deferral for those
    elif not _WGRAD_FLUSH_SCHEDULED:
        _WGRAD_FLUSH_SCHEDULED = True
        torch.autograd.Variable._execution_engine.queue_callback(flush_wgrad_queue)


def _grad_target(tgt, shape, device) -> tuple[torch.Tensor, bool]:
    """fp32 buffer the backward kernels accumulate into.  A parameter may carry `_aq_grad` (a view of a flat
    gradient buffer, see aqualora_b200/ppft.py): then gradients are accumulated there and autograd gets None."""
    if tgt is not None:
        return tgt.view(shape), True
    return torch.zeros(shape, dtype=torch.float32, device=device), False


class _FusedLoraProjection(torch.autograd.Function):
    """y = x W^T + b + ((x Dn^T) (.) s) Up^T  on [M, din] bf16 rows; see include/aqualora_b200.h."""

    @staticmethod
    def forward(ctx, x2d, weight, bias, down, up, scale_eff, tokens, residual=None):
        dout, din = weight.shape[0], x2d.shape[1]
        w2d = weight.reshape(dout, din)
        need_grad = any(ctx.needs_input_grad[:6])     # (grad mode itself is off inside Function.forward)
        if down is not None:
            r = down.shape[0]
            dn16, _ = _packed(down, r, din)
            up16, _ = _packed(up, dout, r)
            y, h = ops.lora_linear_fwd(x2d, w2d, bias, dn16, up16, scale_eff.detach(), tokens, save_h=need_grad, residual=residual)
        else:
            y, h = ops.lora_linear_fwd(x2d, w2d, bias, None, None, None, tokens, residual=residual)
        ctx.has_residual = residual is not None
        ctx.tokens = tokens
        ctx.has_lora = down is not None
        # python-side handles: `_aq_grad` (direct accumulation target) lives on the caller's tensor objects
        ctx.grad_targets = tuple(getattr(t, "_aq_grad", None) for t in (down, up, scale_eff))
        ctx.save_for_backward(x2d, weight, down, up, scale_eff, h)
        return y

    @staticmethod
    def backward(ctx, gy):
        x2d, weight, down, up, scale_eff, h = ctx.saved_tensors
        dout, din = weight.shape[0], x2d.shape[1]
        gy = gy if gy.stride(1) == 1 and gy.stride(0) % 8 == 0 else gy.contiguous()
        need_dx = ctx.needs_input_grad[0]
        w_t = _weight_t(weight, dout, din) if need_dx else None
        g_res = gy if ctx.has_residual and ctx.needs_input_grad[7] else None      # y = ... + residual: identity
        if not ctx.has_lora:
            gx = None
            if need_dx:
                gx, _ = ops.lora_linear_fwd(gy, w_t, None, None, None, None, ctx.tokens)
            return gx, None, None, None, None, None, None, g_res
        r = down.shape[0]
        _, dn16_t = _packed(down, r, din)
        _, up16_t = _packed(up, dout, r)
        t_down, t_up, t_scale = ctx.grad_targets
        g_down, down_direct = _grad_target(t_down, (r, din), gy.device)
        g_up, up_direct = _grad_target(t_up, (dout, r), gy.device)
        g_scale, scale_direct = None, False
        if ctx.needs_input_grad[5]:
            g_scale, scale_direct = _grad_target(t_scale, tuple(scale_eff.shape), gy.device)
        gx, ws = ops.lora_linear_bwd_dx(gy, w_t, dn16_t, up16_t, scale_eff.detach(), h, g_scale, ctx.tokens)
        _enqueue_wgrad((gy, x2d, ws, g_down, g_up), deferrable=down_direct and up_direct)
        return (gx, None, None,
                None if down_direct else g_down.view_as(down).to(down.dtype),
                None if up_direct else g_up.view_as(up).to(up.dtype),
                None if (g_scale is None or scale_direct) else g_scale, None, g_res)


class _GroupedLoraProjection(torch.autograd.Function):
    """n projections of the SAME rows in one launch (aq_lora_linear_fwd_grouped): y_i = x W_i^T + b_i + ((x Dn_i^T) (.) s) Up_i^T.
    `flat` = (weight, bias, down, up) per projection; down/up are given for all projections or for none.  The backward runs
    the per-projection kernels (the G_i arrive as separate tensors) and sums the dX contributions."""

    @staticmethod
    def forward(ctx, x2d, scale_eff, tokens, n, *flat):
        din = x2d.shape[1]
        need_grad = any(ctx.needs_input_grad)
        ws, bs, dns, ups = flat[0::4], flat[1::4], flat[2::4], flat[3::4]
        has_lora = dns[0] is not None
        projections = []
        for w, b, dn, up in zip(ws, bs, dns, ups):
            dout = w.shape[0]
            if has_lora:
                r = dn.shape[0]
                projections.append((w.reshape(dout, din), b, _packed(dn, r, din)[0], _packed(up, dout, r)[0]))
            else:
                projections.append((w.reshape(dout, din), b, None, None))
        outs = ops.lora_linear_fwd_grouped(x2d, projections, scale_eff.detach() if has_lora else None, tokens,
                                           save_h=need_grad and has_lora)
        ctx.tokens, ctx.n, ctx.has_lora = tokens, n, has_lora
        ctx.grad_targets = tuple((getattr(dn, "_aq_grad", None), getattr(up, "_aq_grad", None)) for dn, up in zip(dns, ups))
        ctx.scale_target = getattr(scale_eff, "_aq_grad", None)
        ctx.save_for_backward(x2d, scale_eff, *ws, *dns, *ups, *[h for _, h in outs])
        return tuple(y for y, _ in outs)

    @staticmethod
    def backward(ctx, *gys):
        n = ctx.n
        saved = ctx.saved_tensors
        x2d, scale_eff = saved[0], saved[1]
        ws, dns, ups, hs = (saved[2 + k * n: 2 + (k + 1) * n] for k in range(4))
        din = x2d.shape[1]
        need_dx = ctx.needs_input_grad[0]
        gx_total = None
        g_scale, scale_direct = None, False
        if ctx.has_lora and ctx.needs_input_grad[1]:
            g_scale, scale_direct = _grad_target(ctx.scale_target, tuple(scale_eff.shape), x2d.device)
        flat_grads = []
        for i in range(n):
            gy = gys[i]
            if gy is None:
                flat_grads.extend([None, None, None, None])
                continue
            gy = gy if gy.stride(1) == 1 and gy.stride(0) % 8 == 0 else gy.contiguous()
            weight, down, up = ws[i], dns[i], ups[i]
            dout = weight.shape[0]
            w_t = _weight_t(weight, dout, din) if need_dx else None
            if not ctx.has_lora:
                gx = ops.lora_linear_fwd(gy, w_t, None, None, None, None, ctx.tokens)[0] if need_dx else None
                flat_grads.extend([None, None, None, None])
            else:
                r = down.shape[0]
                _, dn16_t = _packed(down, r, din)
                _, up16_t = _packed(up, dout, r)
                g_down, down_direct = _grad_target(ctx.grad_targets[i][0], (r, din), gy.device)
                g_up, up_direct = _grad_target(ctx.grad_targets[i][1], (dout, r), gy.device)
                gx, side = ops.lora_linear_bwd_dx(gy, w_t, dn16_t, up16_t, scale_eff.detach(), hs[i], g_scale, ctx.tokens)
                _enqueue_wgrad((gy, x2d, side, g_down, g_up), deferrable=down_direct and up_direct)
                flat_grads.extend([None, None, None if down_direct else g_down.view_as(down).to(down.dtype),
                                   None if up_direct else g_up.view_as(up).to(up.dtype)])
            if gx is not None:
                gx_total = gx if gx_total is None else gx_total.add_(gx)
        return (gx_total, None if (g_scale is None or scale_direct) else g_scale, None, None, *flat_grads)


def _effective_scale(scale, lora_layer, nsamples: int, r: int, device, compute_dtype) -> torch.Tensor:
    """[B, r] fp32 diagonal the kernel applies between down and up (utils/lora_modules.py:15-25): the tensor scale
    (rounded to the compute dtype, as the reference's `.to(weight_dtype)` + autocast matmul do), or the float
    multiplier broadcast, times network_alpha / rank."""
    a = 1.0
    if getattr(lora_layer, "network_alpha", None) is not None:
        a = float(lora_layer.network_alpha) / float(lora_layer.rank)
    if isinstance(scale, torch.Tensor):
        if scale.dim() != 2 or scale.shape[1] != r:
            raise AqualoraError(f"tensor scale must be [batch, rank={r}], got {tuple(scale.shape)}")
        if getattr(scale, "_aq_grad", None) is not None and scale.dtype == torch.float32 and a == 1.0:
            return scale                                       # pre-rounded leaf managed by the PPFT harness
        s = scale.to(compute_dtype).to(torch.float32)
        return s * a if a != 1.0 else s
    return torch.full((1, r), float(scale) * a, dtype=torch.float32, device=device)


def _check_input(x: torch.Tensor, what: str) -> torch.Tensor:
    """The kernels compute in bf16 (the BASELINE precision: `accelerate --mixed_precision bf16`, train/ppft_train.py:569-581).  fp16 /
    fp32 activations -- the reference's README recipe trains in fp16 (train/README.md:34-48) -- are accepted by casting AT THE BOUNDARY:
    the input (and a frozen fp16 / fp32 base weight, once, cached) is rounded to bf16, the result is cast back to the caller's dtype
    by the four forwards.  bf16 keeps 8 mantissa bits against fp16's 11: results agree with an fp16 run to bf16 rounding (2e-2 of the
    tensor's max, tests/test_lora_gpu.py::test_fp16_and_fp32_activations_cast_at_the_boundary), not to fp16 rounding."""
    if not x.is_cuda:
        raise AqualoraError(f"{what}: CPU tensor passed; aqualora_b200 runs on sm_100a only and has no CPU fallback "
                            "(patch oracle.lora_oracle forwards in tests that need a CPU reference)")
    if x.dtype == torch.bfloat16:
        return x
    if x.dtype in (torch.float16, torch.float32):
        return x.to(torch.bfloat16)
    raise AqualoraError(f"{what}: activations must be bf16, fp16 or fp32, got {x.dtype}")


_W16_CACHE: Dict[int, tuple] = {}


def _bf16_weight(weight: torch.Tensor) -> torch.Tensor:
    """The frozen base weight as bf16: itself, or a cached rounded copy of an fp16 / fp32 weight."""
    if weight.dtype == torch.bfloat16:
        return weight
    if weight.dtype not in (torch.float16, torch.float32):
        raise AqualoraError(f"frozen projection weights must be bf16, fp16 or fp32, got {weight.dtype}")
    if weight.requires_grad:
        raise AqualoraError("the base projection weight is frozen in PPFT (train/ppft_train.py:562-565); a trainable non-bf16 base weight is not supported")
    key = id(weight)
    ver = (weight._version, weight.data_ptr())
    hit = _W16_CACHE.get(key)
    if hit is not None and hit[0] == ver and hit[2]() is weight:
        return hit[1]
    w16 = weight.detach().to(torch.bfloat16)
    _W16_CACHE[key] = (ver, w16, weakref.ref(weight))
    return w16


def _rows_view(x: torch.Tensor, din: int) -> torch.Tensor:
    x2d = x.reshape(-1, din)
    if x2d.stride(1) != 1 or (x2d.shape[0] > 1 and x2d.stride(0) % 8 != 0) or x2d.data_ptr() % 16 != 0:
        x2d = x2d.contiguous()
    return x2d


def _project_rows(x2d, weight, bias, down, up, lora_meta, scale, compute_dtype, residual=None):
    """Shared tail of the four forwards: x2d [M, din] bf16 rows -> [M, dout] through the fused kernel (+ residual [M, dout] rows
    added in the tile epilogue when given)."""
    M = x2d.shape[0]
    weight = _bf16_weight(weight)
    if bias is not None and bias.dtype != torch.bfloat16:
        bias = bias.detach().to(torch.bfloat16)
    if down is None:
        return _FusedLoraProjection.apply(x2d, weight, bias, None, None, None, M, residual)
    r = down.shape[0]
    if r % 8 != 0:
        # ranks that are not a multiple of 8 (ppft_train.py's default --rank is 4): zero rows / columns pad the LoRA operands and
        # the diagonal up to the next multiple of 8 -- they contribute exactly 0 and F.pad's backward drops their gradients
        pad = (-r) % 8
        din = down.numel() // r
        down = F.pad(down.reshape(r, din), (0, 0, 0, pad))
        up = F.pad(up.reshape(up.shape[0], r), (0, pad))
        if isinstance(scale, torch.Tensor):
            scale = F.pad(scale, (0, pad))
        r += pad
    if isinstance(scale, torch.Tensor):
        nsamp = scale.shape[0]
        if M % nsamp != 0:
            raise AqualoraError(f"{M} rows cannot be split over a scale batch of {nsamp}")
        tokens = M // nsamp
    else:
        nsamp, tokens = 1, M
    s_eff = _effective_scale(scale, lora_meta, nsamp, r, x2d.device, torch.bfloat16)
    return _FusedLoraProjection.apply(x2d, weight, bias, down, up, s_eff, tokens, residual)


_ZERO_BASE: Dict[tuple, torch.Tensor] = {}


def _zero_base(dout: int, din: int, device) -> torch.Tensor:
    """Zero base weight for a stand-alone LoRA layer call (the fused kernel always carries a base operand)."""
    key = (dout, din, str(device))
    if key not in _ZERO_BASE:
        _ZERO_BASE[key] = torch.zeros((dout, din), dtype=torch.bfloat16, device=device)
    return _ZERO_BASE[key]


# ------------------------------------------------------------------------------------------------
# the reference's four forwards (names and signatures: utils/lora_modules.py:9,28,46,56)
# ------------------------------------------------------------------------------------------------
def CustomLoRALinearLayerforward(self, hidden_states: torch.Tensor, scale: float = 1.0):
    """up(diag(scale) down(x)) [* alpha/rank] [* float scale]  -- utils/lora_modules.py:9-26.
    Stand-alone call (the compatible-linear forward below fuses this into the base projection instead)."""
    orig_dtype = hidden_states.dtype
    hidden_states = _check_input(hidden_states, "LoRALinearLayer")
    down, up = self.down.weight, self.up.weight
    x2d = _rows_view(hidden_states, hidden_states.shape[-1])
    y = _project_rows(x2d, _zero_base(up.shape[0], down.shape[1], x2d.device), None, down, up, self, scale, hidden_states.dtype)
    return y.view(*hidden_states.shape[:-1], up.shape[0]).to(orig_dtype)


def CustomLoRACompatibleLinearforward(self, hidden_states: torch.Tensor, scale: float = 1.0):
    """Linear(x) [+ lora_layer(x, scale)]  -- utils/lora_modules.py:56-62, as one fused kernel."""
    orig_dtype = hidden_states.dtype
    hidden_states = _check_input(hidden_states, "LoRACompatibleLinear")
    x2d = _rows_view(hidden_states, hidden_states.shape[-1])
    lora = None if _LORA_DISABLED else self.lora_layer
    if lora is None:
        y = _project_rows(x2d, self.weight, self.bias, None, None, None, scale, hidden_states.dtype)
    else:
        y = _project_rows(x2d, self.weight, self.bias, lora.down.weight, lora.up.weight, lora, scale, hidden_states.dtype)
    return y.view(*hidden_states.shape[:-1], self.weight.shape[0]).to(orig_dtype)


def _conv_as_rows(hidden_states: torch.Tensor):
    B, C, H, W = hidden_states.shape
    x = hidden_states.contiguous(memory_format=torch.channels_last)      # no-op inside the channels_last U-Net
    return _rows_view(x.permute(0, 2, 3, 1).reshape(B * H * W, C), C), (B, H, W)


def _is_pointwise(conv) -> bool:
    def one(v, want):
        return all(int(t) == want for t in (v if isinstance(v, (tuple, list)) else (v, v)))

    return one(conv.kernel_size, 1) and one(conv.stride, 1) and one(conv.padding, 0) and one(conv.dilation, 1) and conv.groups == 1


def CustomLoRAConv2dLayerforward(self, hidden_states: torch.Tensor, scale: float = 1.0):
    """up(down(x) * scale[:, :, None, None])  -- utils/lora_modules.py:28-44 (1x1 down only: the SD1.5 proj_in/out).
    A pointwise conv over NCHW is the row projection over the channels_last view [B*H*W, C]."""
    if not _is_pointwise(self.down):
        raise AqualoraError("LoRAConv2dLayer: only 1x1 stride-1 down convolutions are implemented (the unet_keys.json targets)")
    orig_dtype = hidden_states.dtype
    hidden_states = _check_input(hidden_states, "LoRAConv2dLayer")
    x2d, (B, H, W) = _conv_as_rows(hidden_states)
    down, up = self.down.weight, self.up.weight
    y = _project_rows(x2d, _zero_base(up.shape[0], down.shape[1], x2d.device), None, down, up, self, scale, hidden_states.dtype)
    return y.view(B, H, W, -1).permute(0, 3, 1, 2).to(orig_dtype)


def CustomLoRACompatibleConvforward(self, hidden_states: torch.Tensor, scale: float = 1.0):
    """conv2d(x) [+ lora_layer(x, scale)]  -- utils/lora_modules.py:46-54."""
    if self.lora_layer is None or (_LORA_DISABLED and not _is_pointwise(self)):
        return F.conv2d(hidden_states, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
    orig_dtype = hidden_states.dtype
    if _LORA_DISABLED:
        hidden_states = _check_input(hidden_states, "LoRACompatibleConv")
        x2d, (B, H, W) = _conv_as_rows(hidden_states)
        y = _project_rows(x2d, self.weight, self.bias, None, None, None, scale, hidden_states.dtype)
        return y.view(B, H, W, -1).permute(0, 3, 1, 2).to(orig_dtype)
    if not (_is_pointwise(self) and _is_pointwise(self.lora_layer.down)):
        raise AqualoraError("LoRACompatibleConv with a LoRA layer: only 1x1 stride-1 convolutions are implemented "
                            "(proj_in / proj_out, the only conv targets in utils/unet_keys.json)")
    hidden_states = _check_input(hidden_states, "LoRACompatibleConv")
    x2d, (B, H, W) = _conv_as_rows(hidden_states)
    lora = self.lora_layer
    # [Co, C, 1, 1] / [r, C, 1, 1] / [Co, r, 1, 1] are row-major matrices already: the kernel reads them in place
    y = _project_rows(x2d, self.weight, self.bias, lora.down.weight, lora.up.weight, lora, scale, hidden_states.dtype)
    return y.view(B, H, W, -1).permute(0, 3, 1, 2).to(orig_dtype)


# ------------------------------------------------------------------------------------------------
# projection + residual add in one kernel (the adds around the LoRA-target GEMMs of a Transformer2D block)
# ------------------------------------------------------------------------------------------------
def _residual_rows(residual: torch.Tensor, M: int, dout: int):
    """[M, dout] bf16 row view of the residual stream, or None when it cannot ride in the epilogue (then the caller adds it)."""
    if residual is None or not residual.is_cuda or residual.dtype != torch.bfloat16 or residual.numel() != M * dout:
        return None
    r2d = residual.reshape(M, dout)
    if r2d.stride(1) != 1 or r2d.stride(0) % 8 != 0 or r2d.data_ptr() % 16 != 0:
        return None
    return r2d


def linear_with_residual(module, hidden_states: torch.Tensor, scale, residual: torch.Tensor):
    """`module(hidden_states, scale) + residual` for a LoRA-compatible linear patched with this library's forward: the add happens in
    the GEMM's tile epilogue (aq_lora_linear_fwd_residual) instead of a separate pass over [M, dout].  Any other module, dtype or
    layout takes the plain two-op form -- same arithmetic up to one bf16 rounding (the sum is rounded once instead of twice)."""
    ours = isinstance(module, nn.Linear) and getattr(module.forward, "__func__", None) is CustomLoRACompatibleLinearforward
    if ours and hidden_states.is_cuda and hidden_states.dtype == torch.bfloat16:
        x2d = _rows_view(hidden_states, hidden_states.shape[-1])
        r2d = _residual_rows(residual, x2d.shape[0], module.weight.shape[0])
        if r2d is not None:
            lora = None if _LORA_DISABLED else module.lora_layer
            if lora is None:
                y = _project_rows(x2d, module.weight, module.bias, None, None, None, scale, hidden_states.dtype, r2d)
            else:
                y = _project_rows(x2d, module.weight, module.bias, lora.down.weight, lora.up.weight, lora, scale, hidden_states.dtype, r2d)
            return y.view(*hidden_states.shape[:-1], module.weight.shape[0])
    return module(hidden_states, scale) + residual


def conv1x1_with_residual(module, hidden_states: torch.Tensor, scale, residual: torch.Tensor):
    """`module(hidden_states, scale) + residual` for the 1x1 proj_out convolution of an SD1.5 Transformer2D block (channels_last rows)."""
    ours = (isinstance(module, nn.Conv2d) and getattr(module.forward, "__func__", None) is CustomLoRACompatibleConvforward and _is_pointwise(module)
            and (module.lora_layer is None or _is_pointwise(module.lora_layer.down)))
    if (ours and hidden_states.is_cuda and hidden_states.dtype == torch.bfloat16 and residual.dtype == torch.bfloat16
            and residual.shape[1] == module.weight.shape[0] and residual.is_contiguous(memory_format=torch.channels_last)):
        x2d, (B, H, W) = _conv_as_rows(hidden_states)
        r2d = _residual_rows(residual.permute(0, 2, 3, 1), x2d.shape[0], module.weight.shape[0])
        if r2d is not None:
            lora = None if _LORA_DISABLED else module.lora_layer
            if lora is None:
                y = _project_rows(x2d, module.weight, module.bias, None, None, None, scale, hidden_states.dtype, r2d)
            else:
                y = _project_rows(x2d, module.weight, module.bias, lora.down.weight, lora.up.weight, lora, scale, hidden_states.dtype, r2d)
            return y.view(B, H, W, -1).permute(0, 3, 1, 2)
    return module(hidden_states, scale) + residual


# ------------------------------------------------------------------------------------------------
# several projections of one input in one launch; the attention processor that uses it
# ------------------------------------------------------------------------------------------------
def _alpha_over_rank(lora) -> float:
    a = getattr(lora, "network_alpha", None)
    return 1.0 if a is None else float(a) / float(lora.rank)


def project_group(modules, hidden_states: torch.Tensor, scale=1.0):
    """`[m(hidden_states, scale) for m in modules]` for LoRA-compatible linears that read the same input -- q / k / v of a
    self-attention, k / v of a cross-attention, or the K / V projections of every cross-attention of a U-Net (all read the text
    context) -- as ONE grouped launch of the fused kernel (aq_lora_linear_fwd_grouped).  Each module keeps the semantics of
    CustomLoRACompatibleLinearforward (utils/lora_modules.py:56-62).  Modules that cannot share a launch (different input width
    or rank, LoRA on some but not all, more than 32) go through their own fused launch instead."""
    modules = list(modules)
    din = hidden_states.shape[-1]
    ours = all(isinstance(m, nn.Linear) and getattr(m.forward, "__func__", None) is CustomLoRACompatibleLinearforward for m in modules)
    loras = [None if _LORA_DISABLED else getattr(m, "lora_layer", None) for m in modules]
    with_lora = [l is not None for l in loras]
    groupable = (ours and 1 < len(modules) <= 32 and hidden_states.dtype == torch.bfloat16
                 and all(m.weight.shape[1] == din and m.weight.dtype == torch.bfloat16 for m in modules)
                 and (all(with_lora) or not any(with_lora)))
    if groupable and all(with_lora):
        r = loras[0].down.weight.shape[0]
        a = _alpha_over_rank(loras[0])
        # one grouped launch covers ranks up to 64 that are multiples of 8; other ranks take the per-module path (chunked / padded)
        groupable = r <= 64 and r % 8 == 0 and all(l.down.weight.shape[0] == r and _alpha_over_rank(l) == a for l in loras)
    if not groupable:
        return [m(hidden_states, scale) for m in modules]
    _check_input(hidden_states, "project_group")
    x2d = _rows_view(hidden_states, din)
    M = x2d.shape[0]
    flat = []
    if all(with_lora):
        if isinstance(scale, torch.Tensor):
            nsamp = scale.shape[0]
            if M % nsamp != 0:
                raise AqualoraError(f"{M} rows cannot be split over a scale batch of {nsamp}")
            tokens = M // nsamp
        else:
            nsamp, tokens = 1, M
        s_eff = _effective_scale(scale, loras[0], nsamp, r, x2d.device, hidden_states.dtype)
        for m, l in zip(modules, loras):
            flat.extend([m.weight, m.bias, l.down.weight, l.up.weight])
    else:
        tokens = M
        s_eff = torch.empty(0, device=x2d.device)
        for m in modules:
            flat.extend([m.weight, m.bias, None, None])
    ys = _GroupedLoraProjection.apply(x2d, s_eff, tokens, len(modules), *flat)
    return [y.view(*hidden_states.shape[:-1], m.weight.shape[0]) for y, m in zip(ys, modules)]


def precompute_cross_kv(attentions, encoder_hidden_states: torch.Tensor, scale=1.0) -> None:
    """Run `to_k` / `to_v` of every cross-attention in `attentions` on the text context in one grouped launch and park the
    results on the modules (`_aq_kv`) for AquaLoRAAttnProcessor, which consumes them.  diffusers calls these 2 x 16 projections
    one by one inside each AttnProcessor although all of them read the same `encoder_hidden_states`."""
    attentions = list(attentions)
    mods = [m for a in attentions for m in (a.to_k, a.to_v)]
    for i in range(0, len(mods), 32):
        outs = project_group(mods[i:i + 32], encoder_hidden_states, scale)
        for j in range(0, len(outs), 2):
            attentions[(i + j) // 2]._aq_kv = (outs[j], outs[j + 1])


def _flash_for_short_kv():
    try:
        from torch.nn.attention import SDPBackend, sdpa_kernel

        return lambda: sdpa_kernel([SDPBackend.FLASH_ATTENTION, SDPBackend.EFFICIENT_ATTENTION, SDPBackend.MATH])
    except ImportError:      # older torch: keep the default dispatch
        return None


_FLASH_FOR_SHORT_KV = _flash_for_short_kv()


class AquaLoRAAttnProcessor:
    """diffusers-style attention processor (`__call__(attn, hidden_states, encoder_hidden_states, attention_mask, temb, scale)`,
    the signature of diffusers 0.24 `AttnProcessor2_0`, which threads `cross_attention_kwargs["scale"]` into `attn.to_q/k/v/
    to_out[0]` -- train/ppft_train.py:1028,1034).  Same arithmetic as calling the four patched projections one by one, but
    q / k / v of a self-attention (and k / v of a cross-attention) leave in one grouped launch."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale=1.0, residual=None):
        kv = getattr(attn, "_aq_kv", None)
        if kv is not None:
            attn._aq_kv = None
            q = attn.to_q(hidden_states, scale)
            k, v = kv
        elif encoder_hidden_states is None:
            q, k, v = project_group((attn.to_q, attn.to_k, attn.to_v), hidden_states, scale)
        else:
            q = attn.to_q(hidden_states, scale)
            k, v = project_group((attn.to_k, attn.to_v), encoder_hidden_states.to(hidden_states.dtype), scale)
        B, N, C = q.shape
        h = attn.heads
        q = q.view(B, N, h, C // h).transpose(1, 2)
        k = k.view(B, k.shape[1], h, C // h).transpose(1, 2)
        v = v.view(B, v.shape[1], h, C // h).transpose(1, 2)
        if getattr(attn, "upcast", False) or getattr(attn, "upcast_attention", False):
            o = F.scaled_dot_product_attention(q.float(), k.float(), v.float(), attn_mask=attention_mask).to(v.dtype)
        elif k.shape[2] <= 128 and q.is_cuda and q.dtype == torch.bfloat16 and attention_mask is None and _FLASH_FOR_SHORT_KV is not None:
            # cross-attention (77 text tokens): the library's flash backend beats its cuDNN default for short K / V on B200
            # (profiles/r02_sdpa_backend_probe.log: 4096 x 77, d = 40: 406 vs 584 us forward + backward)
            with _FLASH_FOR_SHORT_KV():
                o = F.scaled_dot_product_attention(q, k, v)
        else:
            o = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask)
        o = o.transpose(1, 2).reshape(B, N, C)
        if residual is not None:
            o = linear_with_residual(attn.to_out[0], o, scale, residual)      # x + to_out(...) in the projection's epilogue
        else:
            o = attn.to_out[0](o, scale)
        if len(attn.to_out) > 1:
            o = attn.to_out[1](o)      # diffusers: Dropout(p = 0)
        return o


# ------------------------------------------------------------------------------------------------
# diffusers-shaped containers (field names per diffusers 0.24 models/lora.py, which the reference relies on)
# ------------------------------------------------------------------------------------------------
class LoRALinearLayer(nn.Module):
    def __init__(self, in_features: int, out_features: int, rank: int = 4, network_alpha: Optional[float] = None,
                 device=None, dtype=None):
        super().__init__()
        self.down = nn.Linear(in_features, rank, bias=False, device=device, dtype=dtype)
        self.up = nn.Linear(rank, out_features, bias=False, device=device, dtype=dtype)
        self.network_alpha = network_alpha
        self.rank = rank
        self.in_features, self.out_features = in_features, out_features
        nn.init.normal_(self.down.weight, std=1 / rank)
        nn.init.zeros_(self.up.weight)

    forward = CustomLoRALinearLayerforward


class LoRAConv2dLayer(nn.Module):
    def __init__(self, in_features: int, out_features: int, rank: int = 4, kernel_size=(1, 1), stride=(1, 1), padding=0,
                 network_alpha: Optional[float] = None):
        super().__init__()
        self.down = nn.Conv2d(in_features, rank, kernel_size=kernel_size, stride=stride, padding=padding, bias=False)
        self.up = nn.Conv2d(rank, out_features, kernel_size=(1, 1), stride=(1, 1), bias=False)
        self.network_alpha = network_alpha
        self.rank = rank
        nn.init.normal_(self.down.weight, std=1 / rank)
        nn.init.zeros_(self.up.weight)

    forward = CustomLoRAConv2dLayerforward


class LoRACompatibleLinear(nn.Linear):
    def __init__(self, *args, lora_layer: Optional[LoRALinearLayer] = None, **kwargs):
        super().__init__(*args, **kwargs)
        self.lora_layer = lora_layer

    def set_lora_layer(self, lora_layer):
        self.lora_layer = lora_layer

    forward = CustomLoRACompatibleLinearforward


class LoRACompatibleConv(nn.Conv2d):
    def __init__(self, *args, lora_layer: Optional[LoRAConv2dLayer] = None, **kwargs):
        super().__init__(*args, **kwargs)
        self.lora_layer = lora_layer

    def set_lora_layer(self, lora_layer):
        self.lora_layer = lora_layer

    forward = CustomLoRACompatibleConvforward


# ------------------------------------------------------------------------------------------------
# injection / patching / state-dict naming (train/ppft_train.py:443-471, :620-689)
# ------------------------------------------------------------------------------------------------
def resolve(root: nn.Module, key: str) -> nn.Module:
    mod = root
    for sub in key.split("."):
        mod = getattr(mod, sub)
    return mod


def inject_lora(unet: nn.Module, keys: Iterable[str], rank: int, network_alpha: Optional[float] = None):
    """Build one LoRA layer per target module and attach it (train/ppft_train.py:635-678).  Returns the list of
    (key, target_module, lora_layer)."""
    out = []
    for key in keys:
        target = resolve(unet, key)
        if isinstance(target, nn.Conv2d):
            lora = LoRAConv2dLayer(target.in_channels, target.out_channels, rank=rank, kernel_size=target.kernel_size,
                                   stride=target.stride, padding=target.padding, network_alpha=network_alpha)
        elif isinstance(target, nn.Linear):
            lora = LoRALinearLayer(target.in_features, target.out_features, rank, network_alpha=network_alpha)
        else:
            raise ValueError(f"Module {key} is not a LoRACompatibleConv or LoRACompatibleLinear module.")
        lora.to(target.weight.device)
        target.set_lora_layer(lora) if hasattr(target, "set_lora_layer") else setattr(target, "lora_layer", lora)
        out.append((key, target, lora))
    return out


def patch_unet(unet: nn.Module, linear_cls=None, conv_cls=None, linear_layer_cls=None, conv_layer_cls=None) -> int:
    """Monkey-patch every LoRA-compatible module and its lora_layer (train/ppft_train.py:681-689).  With no class
    arguments this patches the in-repo containers; pass diffusers' classes to patch a diffusers U-Net."""
    linear_cls = linear_cls or LoRACompatibleLinear
    conv_cls = conv_cls or LoRACompatibleConv
    n = 0
    for _, module in unet.named_modules():
        if isinstance(module, conv_cls):
            module.forward = types.MethodType(CustomLoRACompatibleConvforward, module)
            if module.lora_layer is not None:
                module.lora_layer.forward = types.MethodType(CustomLoRAConv2dLayerforward, module.lora_layer)
            n += 1
        elif isinstance(module, linear_cls):
            module.forward = types.MethodType(CustomLoRACompatibleLinearforward, module)
            if module.lora_layer is not None:
                module.lora_layer.forward = types.MethodType(CustomLoRALinearLayerforward, module.lora_layer)
            n += 1
    return n


def lora_state_dict_key(key: str) -> str:
    """Module path -> key stem inside pytorch_lora_weights.safetensors (train/ppft_train.py:459-467)."""
    k = key.replace(".proj_in", ".proj_in.lora").replace(".proj_out", ".proj_out.lora")
    k = k.replace(".to_q", ".processor.to_q_lora").replace(".to_k", ".processor.to_k_lora")
    k = k.replace(".to_v", ".processor.to_v_lora").replace(".to_out.0", ".processor.to_out_lora")
    if "ff" in k:
        k = k + ".lora"
    return k


def unet_attn_processors_state_dict(unet: nn.Module, keys: Iterable[str]) -> Dict[str, torch.Tensor]:
    """train/ppft_train.py:443-471: {'<renamed key>.down.weight' / '.up.weight': tensor} for every target."""
    out: Dict[str, torch.Tensor] = {}
    for key in keys:
        target = resolve(unet, key)
        for pk, p in target.state_dict().items():
            if "lora_layer" in pk:
                out[f"{lora_state_dict_key(key)}.{pk.replace('lora_layer.', '')}"] = p
    return out
