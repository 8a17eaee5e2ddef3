"""Tensor-level wrappers over the C ABI: argument checking, stream plumbing, workspace handling.

PyTorch is used for device memory and streams only; every function here launches hand-written sm_100a kernels
from libaqualora_b200.so and raises if the library or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

_BF16 = torch.bfloat16
_F32 = torch.float32


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def _need(t: torch.Tensor, dtype, name: str, ndim: int | None = None) -> None:
    if not t.is_cuda:
        raise _lib.AqualoraError(f"{name} must be a CUDA tensor (aqualora_b200 has no CPU path)")
    if t.dtype != dtype:
        raise _lib.AqualoraError(f"{name} must be {dtype}, got {t.dtype}")
    if ndim is not None and t.dim() != ndim:
        raise _lib.AqualoraError(f"{name} must be {ndim}-D, got shape {tuple(t.shape)}")


def _rows(t: torch.Tensor, name: str) -> int:
    """Leading dimension (in elements) of a 2-D row-major view."""
    if t.stride(1) != 1:
        raise _lib.AqualoraError(f"{name} must be row-major (unit inner stride), got strides {t.stride()}")
    return t.stride(0) if t.shape[0] > 1 else t.shape[1]


def set_tuning(block_n: int = 0, group_size: int = 0) -> None:
    _lib.call("aq_lora_set_tuning", block_n, group_size)


def lora_linear_fwd(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None, down: torch.Tensor | None,
                    up: torch.Tensor | None, scale: torch.Tensor | None, tokens_per_sample: int, save_h: bool = False,
                    out: torch.Tensor | None = None, residual: torch.Tensor | None = None):
    """y = x @ w.T + bias + ((x @ down.T) * scale[row // tokens]) @ up.T ; returns (y, h or None).

    x [M, din] bf16, w [dout, din] bf16, bias [dout] bf16|None, down [r, din] bf16|None, up [dout, r] bf16,
    scale [M // tokens, r] fp32.  (utils/lora_modules.py:9-26,56-62)
    """
    _need(x, _BF16, "x", 2)
    _need(w, _BF16, "w", 2)
    M, din = x.shape
    dout = w.shape[0]
    if w.shape[1] != din or not w.is_contiguous():
        raise _lib.AqualoraError(f"w must be contiguous [dout, {din}], got {tuple(w.shape)}")
    r = 0
    h = None
    if down is not None:
        _need(down, _BF16, "down", 2)
        _need(up, _BF16, "up", 2)
        _need(scale, _F32, "scale", 2)
        r = down.shape[0]
        if tuple(down.shape) != (r, din) or tuple(up.shape) != (dout, r) or not down.is_contiguous() or not up.is_contiguous():
            raise _lib.AqualoraError(f"down/up must be contiguous [r, din]/[dout, r], got {tuple(down.shape)}/{tuple(up.shape)}")
        nsamp = (M + tokens_per_sample - 1) // tokens_per_sample
        if tuple(scale.shape) != (nsamp, r) or not scale.is_contiguous():
            raise _lib.AqualoraError(f"scale must be contiguous [{nsamp}, {r}], got {tuple(scale.shape)}")
        if save_h:
            h = torch.empty((M, r), dtype=_BF16, device=x.device)
    if bias is not None:
        _need(bias, _BF16, "bias", 1)
    y = out if out is not None else torch.empty((M, dout), dtype=_BF16, device=x.device)
    if residual is not None:
        _need(residual, _BF16, "residual", 2)
        if tuple(residual.shape) != (M, dout):
            raise _lib.AqualoraError(f"residual must be [{M}, {dout}], got {tuple(residual.shape)}")
        _lib.call("aq_lora_linear_fwd_residual", x.data_ptr(), _rows(x, "x"), w.data_ptr(), _ptr(bias), _ptr(down), _ptr(up), _ptr(scale),
                  residual.data_ptr(), _rows(residual, "residual"), y.data_ptr(), _rows(y, "y"), _ptr(h), M,
                  max(int(tokens_per_sample), 1), din, dout, r, _stream())
        return y, h
    _lib.call("aq_lora_linear_fwd", x.data_ptr(), _rows(x, "x"), w.data_ptr(), _ptr(bias), _ptr(down), _ptr(up), _ptr(scale),
              y.data_ptr(), _rows(y, "y"), _ptr(h), M, max(int(tokens_per_sample), 1), din, dout, r, _stream())
    return y, h


def lora_linear_fwd_grouped(x: torch.Tensor, projections, scale: torch.Tensor | None, tokens_per_sample: int, save_h: bool = False):
    """Several projections of the same rows x [M, din] in ONE launch (csrc/lora_gemm.cu, grouped work items).
    projections: list of (w [dout, din], bias [dout] | None, down [r, din] | None, up [dout, r] | None), all bf16, `down` given for
    all or none.  Returns [(y, h | None), ...] like lora_linear_fwd."""
    _need(x, _BF16, "x", 2)
    M, din = x.shape
    n = len(projections)
    if not 1 <= n <= 32:
        raise _lib.AqualoraError(f"grouped projection: 1 ... 32 projections per launch, got {n}")
    has_lora = projections[0][2] is not None
    r = projections[0][2].shape[0] if has_lora else 0
    if has_lora:
        _need(scale, _F32, "scale", 2)
        nsamp = (M + tokens_per_sample - 1) // tokens_per_sample
        if tuple(scale.shape) != (nsamp, r) or not scale.is_contiguous():
            raise _lib.AqualoraError(f"scale must be contiguous [{nsamp}, {r}], got {tuple(scale.shape)}")
    arr = (_lib.LoraProjection * n)()
    outs = []
    for i, (w, bias, down, up) in enumerate(projections):
        _need(w, _BF16, "w", 2)
        dout = w.shape[0]
        if w.shape[1] != din or not w.is_contiguous():
            raise _lib.AqualoraError(f"w must be contiguous [dout, {din}], got {tuple(w.shape)}")
        if (down is not None) != has_lora:
            raise _lib.AqualoraError("grouped projection: LoRA operands must be given for all projections or for none")
        h = None
        if has_lora:
            _need(down, _BF16, "down", 2)
            _need(up, _BF16, "up", 2)
            if tuple(down.shape) != (r, din) or tuple(up.shape) != (dout, r) or not down.is_contiguous() or not up.is_contiguous():
                raise _lib.AqualoraError(f"down/up must be contiguous [r, din]/[dout, r], got {tuple(down.shape)}/{tuple(up.shape)}")
            if save_h:
                h = torch.empty((M, r), dtype=_BF16, device=x.device)
        if bias is not None:
            _need(bias, _BF16, "bias", 1)
        y = torch.empty((M, dout), dtype=_BF16, device=x.device)
        arr[i].w, arr[i].bias, arr[i].down, arr[i].up = w.data_ptr(), _ptr(bias), _ptr(down), _ptr(up)
        arr[i].y, arr[i].ldy, arr[i].h_save, arr[i].dout = y.data_ptr(), y.stride(0), _ptr(h), dout
        outs.append((y, h))
    _lib.call("aq_lora_linear_fwd_grouped", x.data_ptr(), _rows(x, "x"), ctypes.addressof(arr), n, _ptr(scale),
              M, max(int(tokens_per_sample), 1), din, r, _stream())
    return outs


def lora_linear_bwd(gy: torch.Tensor, x: torch.Tensor, w_t: torch.Tensor | None, down_t: torch.Tensor, up_t: torch.Tensor,
                    scale: torch.Tensor, h: torch.Tensor, g_down: torch.Tensor, g_up: torch.Tensor,
                    g_scale: torch.Tensor | None, tokens_per_sample: int):
    """Backward of lora_linear_fwd.  Accumulates into g_down [r, din], g_up [dout, r], g_scale [B, r] (fp32) and returns
    gx [M, din] bf16 (None when w_t is None)."""
    _need(gy, _BF16, "gy", 2)
    _need(x, _BF16, "x", 2)
    _need(h, _BF16, "h", 2)
    _need(down_t, _BF16, "down_t", 2)
    _need(up_t, _BF16, "up_t", 2)
    _need(scale, _F32, "scale", 2)
    _need(g_down, _F32, "g_down", 2)
    _need(g_up, _F32, "g_up", 2)
    M, dout = gy.shape
    din = x.shape[1]
    r = h.shape[1]
    if tuple(down_t.shape) != (din, r) or tuple(up_t.shape) != (r, dout):
        raise _lib.AqualoraError(f"down_t/up_t must be [din, r]/[r, dout], got {tuple(down_t.shape)}/{tuple(up_t.shape)}")
    if tuple(g_down.shape) != (r, din) or tuple(g_up.shape) != (dout, r):
        raise _lib.AqualoraError("g_down/g_up must be [r, din]/[dout, r]")
    for t, n in ((down_t, "down_t"), (up_t, "up_t"), (h, "h"), (g_down, "g_down"), (g_up, "g_up"), (scale, "scale")):
        if not t.is_contiguous():
            raise _lib.AqualoraError(f"{n} must be contiguous")
    gx = None
    if w_t is not None:
        _need(w_t, _BF16, "w_t", 2)
        if tuple(w_t.shape) != (din, dout) or not w_t.is_contiguous():
            raise _lib.AqualoraError(f"w_t must be contiguous [{din}, {dout}]")
        gx = torch.empty((M, din), dtype=_BF16, device=gy.device)
    if g_scale is not None:
        _need(g_scale, _F32, "g_scale", 2)
    nbytes = _lib.load().aq_lora_linear_bwd_workspace_bytes(M, r)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=gy.device)
    _lib.call("aq_lora_linear_bwd", gy.data_ptr(), _rows(gy, "gy"), x.data_ptr(), _rows(x, "x"), _ptr(w_t), down_t.data_ptr(),
              up_t.data_ptr(), scale.data_ptr(), h.data_ptr(), _ptr(gx), din if gx is not None else 0, g_down.data_ptr(),
              g_up.data_ptr(), _ptr(g_scale), M, int(tokens_per_sample), din, dout, r, ws.data_ptr(), nbytes, _stream())
    return gx


def lora_linear_bwd_dx(gy: torch.Tensor, w_t: torch.Tensor | None, down_t: torch.Tensor, up_t: torch.Tensor, scale: torch.Tensor,
                       h: torch.Tensor, g_scale: torch.Tensor | None, tokens_per_sample: int):
    """First half of lora_linear_bwd: returns (gx or None, ws) -- `ws` holds dH / Hs for the weight-gradient job of this layer
    (`lora_wgrad_batch`) and must stay alive until that job has been launched."""
    _need(gy, _BF16, "gy", 2)
    _need(h, _BF16, "h", 2)
    _need(down_t, _BF16, "down_t", 2)
    _need(up_t, _BF16, "up_t", 2)
    _need(scale, _F32, "scale", 2)
    M, dout = gy.shape
    din, r = down_t.shape
    if tuple(up_t.shape) != (r, dout) or tuple(h.shape) != (M, r):
        raise _lib.AqualoraError(f"up_t / h must be [r, dout] / [M, r], got {tuple(up_t.shape)} / {tuple(h.shape)}")
    for t, n in ((down_t, "down_t"), (up_t, "up_t"), (h, "h"), (scale, "scale")):
        if not t.is_contiguous():
            raise _lib.AqualoraError(f"{n} must be contiguous")
    gx = None
    if w_t is not None:
        _need(w_t, _BF16, "w_t", 2)
        if tuple(w_t.shape) != (din, dout) or not w_t.is_contiguous():
            raise _lib.AqualoraError(f"w_t must be contiguous [{din}, {dout}]")
        gx = torch.empty((M, din), dtype=_BF16, device=gy.device)
    if g_scale is not None:
        _need(g_scale, _F32, "g_scale", 2)
    nbytes = _lib.load().aq_lora_linear_bwd_workspace_bytes(M, r)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=gy.device)
    _lib.call("aq_lora_linear_bwd_dx", gy.data_ptr(), _rows(gy, "gy"), _ptr(w_t), down_t.data_ptr(), up_t.data_ptr(), scale.data_ptr(),
              h.data_ptr(), _ptr(gx), din if gx is not None else 0, _ptr(g_scale), M, int(tokens_per_sample), din, dout, r, ws.data_ptr(),
              nbytes, _stream())
    return gx, ws


def lora_wgrad_batch(jobs) -> None:
    """jobs: list of (gy [M, dout], x [M, din], ws, g_down [r, din] fp32, g_up [dout, r] fp32): the weight gradients of several layers,
    accumulated in as few launches as possible (aq_lora_wgrad_batch)."""
    for i in range(0, len(jobs), 64):
        part = jobs[i:i + 64]
        arr = (_lib.WgradJob * len(part))()
        for k, (gy, x, ws, g_down, g_up) in enumerate(part):
            _need(gy, _BF16, "gy", 2)
            _need(x, _BF16, "x", 2)
            _need(g_down, _F32, "g_down", 2)
            _need(g_up, _F32, "g_up", 2)
            r, din = g_down.shape
            dout = g_up.shape[0]
            if not g_down.is_contiguous() or not g_up.is_contiguous() or gy.shape[0] != x.shape[0]:
                raise _lib.AqualoraError("lora_wgrad_batch: contiguous gradient buffers and equal row counts expected")
            arr[k].gy, arr[k].ldgy, arr[k].x, arr[k].ldx = gy.data_ptr(), _rows(gy, "gy"), x.data_ptr(), _rows(x, "x")
            arr[k].ws, arr[k].g_down, arr[k].g_up = ws.data_ptr(), g_down.data_ptr(), g_up.data_ptr()
            arr[k].M, arr[k].din, arr[k].dout, arr[k].r = gy.shape[0], din, dout, r
        _lib.call("aq_lora_wgrad_batch", ctypes.addressof(arr), len(part), _stream())


def wgrad_tn(p: torch.Tensor, q: torch.Tensor, c: torch.Tensor, transpose_out: bool = False) -> None:
    """c[i, j] += sum_m p[m, i] * q[m, j]   (c is [J, I] when transpose_out)."""
    _need(p, _BF16, "p", 2)
    _need(q, _BF16, "q", 2)
    _need(c, _F32, "c", 2)
    M, I = p.shape
    J = q.shape[1]
    _lib.call("aq_wgrad_tn", p.data_ptr(), _rows(p, "p"), q.data_ptr(), _rows(q, "q"), c.data_ptr(), c.stride(0), M, I, J,
              int(transpose_out), _stream())


def mapper_fwd(msg: torch.Tensor, emb: torch.Tensor, round_bf16: bool = True) -> torch.Tensor:
    _need(msg, _F32, "msg", 2)
    _need(emb, _F32, "emb", 2)
    B, bits = msg.shape
    r = emb.shape[1]
    out = torch.empty((B, r), dtype=_F32, device=msg.device)
    _lib.call("aq_mapper_fwd", msg.contiguous().data_ptr(), emb.contiguous().data_ptr(), out.data_ptr(), B, bits, r,
              int(round_bf16), _stream())
    return out


def mapper_bwd(msg: torch.Tensor, g_scale: torch.Tensor, g_emb: torch.Tensor) -> None:
    _need(msg, _F32, "msg", 2)
    _need(g_scale, _F32, "g_scale", 2)
    _need(g_emb, _F32, "g_emb", 2)
    B, bits = msg.shape
    r = g_scale.shape[1]
    _lib.call("aq_mapper_bwd", msg.contiguous().data_ptr(), g_scale.contiguous().data_ptr(), g_emb.data_ptr(), B, bits, r,
              _stream())


def cast_transpose_bf16(src: torch.Tensor, dst: torch.Tensor | None, dst_t: torch.Tensor | None) -> None:
    _need(src, _F32, "src", 2)
    rows, cols = src.shape
    _lib.call("aq_cast_transpose_bf16", src.data_ptr(), _ptr(dst), _ptr(dst_t), rows, cols, _stream())


def cast_transpose_bf16_batched(jobs: torch.Tensor, total_tiles: int) -> None:
    """jobs: int64 [n, 7] device table of (src, dst, dst_t, rows, cols, tile_begin, tiles_x) -- `aq_cast_job`."""
    _need(jobs, torch.int64, "jobs", 2)
    if jobs.shape[1] != 7 or not jobs.is_contiguous():
        raise _lib.AqualoraError("jobs must be a contiguous int64 [n, 7] table")
    _lib.call("aq_cast_transpose_bf16_batched", jobs.data_ptr(), jobs.shape[0], int(total_tiles), _stream())


def transpose_bf16(src: torch.Tensor) -> torch.Tensor:
    _need(src, _BF16, "src", 2)
    rows, cols = src.shape
    dst = torch.empty((cols, rows), dtype=_BF16, device=src.device)
    _lib.call("aq_transpose_bf16", src.contiguous().data_ptr(), dst.data_ptr(), rows, cols, _stream())
    return dst


def flat_sumsq(g: torch.Tensor, out: torch.Tensor) -> None:
    _need(g, _F32, "g", 1)
    _need(out, _F32, "norm_sq")
    _lib.call("aq_flat_sumsq", g.data_ptr(), g.numel(), out.data_ptr(), _stream())


def flat_clip_adamw(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, norm_sq: torch.Tensor, *,
                    grad_scale: float, max_norm: float, lr: float, beta1: float, beta2: float, eps: float,
                    weight_decay: float, step: int) -> None:
    for t, n in ((p, "p"), (g, "g"), (m, "m"), (v, "v")):
        _need(t, _F32, n, 1)
    _lib.call("aq_flat_clip_adamw", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), norm_sq.data_ptr(),
              grad_scale, max_norm, lr, beta1, beta2, eps, weight_decay, step, _stream())


def secret_encoder_fwd(msg: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, wc: torch.Tensor, bc: torch.Tensor,
                       x: torch.Tensor | None, hw: tuple[int, int], base: int = 32, res: int = 64):
    """SecretEncoder forward (utils/models.py:74-81): returns (x + c or None, c), c [B, 4, H, W] fp32."""
    for t, n in ((msg, "msg"), (w1, "w1"), (b1, "b1"), (wc, "wc"), (bc, "bc")):
        _need(t, _F32, n)
    B, bits = msg.shape
    H, W = hw
    c = torch.empty((B, 4, H, W), dtype=_F32, device=msg.device)
    xo = None
    if x is not None:
        _need(x, _F32, "x", 4)
        x = x.contiguous()
        xo = torch.empty_like(x)
    ws = torch.empty(_lib.load().aq_secret_encoder_workspace_bytes(B, res), dtype=torch.uint8, device=msg.device)
    _lib.call("aq_secret_encoder_fwd", msg.contiguous().data_ptr(), w1.contiguous().data_ptr(), b1.contiguous().data_ptr(),
              wc.contiguous().data_ptr(), bc.contiguous().data_ptr(), _ptr(x), c.data_ptr(), _ptr(xo), B, bits, base, res, H, W,
              ws.data_ptr(), _stream())
    return xo, c


# ------------------------------------------------------------------------------------------------------------------
# noise_layers (utils/noise_layers/*): [B, 3, H, W] fp32 images in [-1, 1]
# ------------------------------------------------------------------------------------------------------------------
def _image(x: torch.Tensor, name: str = "image") -> torch.Tensor:
    _need(x, _F32, name, 4)
    if x.shape[1] != 3:
        raise _lib.AqualoraError(f"{name} must be [B, 3, H, W], got {tuple(x.shape)}")
    return x.contiguous()


def noise_jpeg(x: torch.Tensor) -> torch.Tensor:
    x = _image(x)
    B, _, H, W = x.shape
    y = torch.empty_like(x)
    _lib.call("aq_noise_jpeg", x.data_ptr(), y.data_ptr(), B, H, W, _stream())
    return y


def noise_crop_resize(x: torch.Tensor, top: int, left: int, crop_h: int, crop_w: int, resize_h: int, resize_w: int,
                      out_hw=(512, 512)) -> torch.Tensor:
    x = _image(x)
    B, _, H, W = x.shape
    y = torch.empty((B, 3, out_hw[0], out_hw[1]), dtype=_F32, device=x.device)
    _lib.call("aq_noise_crop_resize", x.data_ptr(), y.data_ptr(), B, H, W, int(top), int(left), int(crop_h), int(crop_w),
              int(resize_h), int(resize_w), int(out_hw[0]), int(out_hw[1]), _stream())
    return y


def noise_gauss_blur(x: torch.Tensor, sigmas: torch.Tensor, ksize=(3, 9)) -> torch.Tensor:
    x = _image(x)
    _need(sigmas, _F32, "sigmas", 1)
    B, _, H, W = x.shape
    if sigmas.shape[0] != B:
        raise _lib.AqualoraError(f"sigmas must have one entry per sample ({B}), got {tuple(sigmas.shape)}")
    y = torch.empty_like(x)
    _lib.call("aq_noise_gauss_blur", x.data_ptr(), y.data_ptr(), sigmas.contiguous().data_ptr(), B, H, W, int(ksize[0]),
              int(ksize[1]), _stream())
    return y


def noise_gauss_noise(x: torch.Tensor | None, std: float, seed: int, offset: int = 0, shape=None, device=None) -> torch.Tensor:
    """x + std * N(0, 1) with the in-kernel Philox stream (seed, offset); x = None returns std * noise of `shape`."""
    if x is not None:
        _need(x, _F32, "x")
        x = x.contiguous()
        y = torch.empty_like(x)
    else:
        y = torch.empty(shape, dtype=_F32, device=device)
    _lib.call("aq_noise_gauss_noise", _ptr(x), y.data_ptr(), y.numel(), float(std), int(seed), int(offset), _stream())
    return y


def noise_color_jiggle(x: torch.Tensor, params: torch.Tensor, order) -> torch.Tensor:
    """params [B, 4] = (brightness, contrast, saturation, hue); order = permutation of (0, 1, 2, 3)."""
    import ctypes

    x = _image(x)
    _need(params, _F32, "params", 2)
    B, _, H, W = x.shape
    if tuple(params.shape) != (B, 4):
        raise _lib.AqualoraError(f"params must be [{B}, 4], got {tuple(params.shape)}")
    y = torch.empty_like(x)
    arr = (ctypes.c_int * 4)(*[int(o) for o in order])
    _lib.call("aq_noise_color_jiggle", x.data_ptr(), y.data_ptr(), params.contiguous().data_ptr(), arr, B, H, W, _stream())
    return y


def secret_encoder_bwd(g_c: torch.Tensor, msg: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, wc: torch.Tensor,
                       g_w1: torch.Tensor, g_b1: torch.Tensor, g_wc: torch.Tensor, g_bc: torch.Tensor, base: int = 32, res: int = 64) -> None:
    """Accumulate the SecretEncoder parameter gradients for g_c = dL/dc [B, 4, H, W]."""
    for t, n in ((g_c, "g_c"), (msg, "msg"), (w1, "w1"), (b1, "b1"), (wc, "wc"), (g_w1, "g_w1"), (g_b1, "g_b1"), (g_wc, "g_wc"), (g_bc, "g_bc")):
        _need(t, _F32, n)
        if not t.is_contiguous():
            raise _lib.AqualoraError(f"{n} must be contiguous")
    B, bits = msg.shape
    H, W = g_c.shape[2:]
    nbytes = _lib.load().aq_secret_encoder_bwd_workspace_bytes(B, base, res)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=g_c.device)
    _lib.call("aq_secret_encoder_bwd", g_c.data_ptr(), msg.data_ptr(), w1.data_ptr(), b1.data_ptr(), wc.data_ptr(), g_w1.data_ptr(),
              g_b1.data_ptr(), g_wc.data_ptr(), g_bc.data_ptr(), B, bits, base, res, H, W, ws.data_ptr(), nbytes, _stream())


def noise_jpeg_bwd(gy: torch.Tensor) -> torch.Tensor:
    gy = _image(gy, "gy")
    B, _, H, W = gy.shape
    gx = torch.empty_like(gy)
    _lib.call("aq_noise_jpeg_bwd", gy.data_ptr(), gx.data_ptr(), B, H, W, _stream())
    return gx


def noise_crop_resize_bwd(gy: torch.Tensor, in_hw, top: int, left: int, crop_h: int, crop_w: int, resize_h: int, resize_w: int) -> torch.Tensor:
    gy = _image(gy, "gy")
    B, _, oh, ow = gy.shape
    H, W = in_hw
    gx = torch.empty((B, 3, H, W), dtype=_F32, device=gy.device)
    _lib.call("aq_noise_crop_resize_bwd", gy.data_ptr(), gx.data_ptr(), B, H, W, int(top), int(left), int(crop_h), int(crop_w),
              int(resize_h), int(resize_w), int(oh), int(ow), _stream())
    return gx


def noise_gauss_blur_bwd(gy: torch.Tensor, sigmas: torch.Tensor, ksize=(3, 9)) -> torch.Tensor:
    gy = _image(gy, "gy")
    _need(sigmas, _F32, "sigmas", 1)
    B, _, H, W = gy.shape
    gx = torch.empty_like(gy)
    _lib.call("aq_noise_gauss_blur_bwd", gy.data_ptr(), gx.data_ptr(), sigmas.contiguous().data_ptr(), B, H, W, int(ksize[0]),
              int(ksize[1]), _stream())
    return gx


def noise_color_jiggle_bwd(x: torch.Tensor, gy: torch.Tensor, params: torch.Tensor, order) -> torch.Tensor:
    import ctypes

    x = _image(x)
    gy = _image(gy, "gy")
    _need(params, _F32, "params", 2)
    B, _, H, W = x.shape
    gx = torch.empty_like(x)
    arr = (ctypes.c_int * 4)(*[int(o) for o in order])
    _lib.call("aq_noise_color_jiggle_bwd", x.data_ptr(), gy.data_ptr(), gx.data_ptr(), params.contiguous().data_ptr(), arr, B, H, W, _stream())
    return gx


def prvl_loss_fwd(img1: torch.Tensor, img2: torch.Tensor):
    """PRVL_loss (train/latent_wm_pretrain.py:42-50): returns (loss [1] fp32, state [1] int64 for the backward)."""
    img1, img2 = _image(img1, "img1"), _image(img2, "img2")
    if img1.shape != img2.shape:
        raise _lib.AqualoraError("prvl_loss: images must have equal shapes")
    B, _, H, W = img1.shape
    loss = torch.empty(1, dtype=_F32, device=img1.device)
    state = torch.empty(1, dtype=torch.int64, device=img1.device)
    nbytes = _lib.load().aq_prvl_workspace_bytes(B, H, W)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=img1.device)
    _lib.call("aq_prvl_loss_fwd", img1.data_ptr(), img2.data_ptr(), loss.data_ptr(), state.data_ptr(), B, H, W, ws.data_ptr(), nbytes, _stream())
    return loss, state


def prvl_loss_bwd(img1: torch.Tensor, img2: torch.Tensor, state: torch.Tensor, g_loss: torch.Tensor, want1: bool, want2: bool):
    img1, img2 = _image(img1, "img1"), _image(img2, "img2")
    B, _, H, W = img1.shape
    g1 = torch.empty_like(img1) if want1 else None
    g2 = torch.empty_like(img2) if want2 else None
    g = g_loss.reshape(1).to(_F32).contiguous()
    _lib.call("aq_prvl_loss_bwd", img1.data_ptr(), img2.data_ptr(), state.data_ptr(), g.data_ptr(), _ptr(g1), _ptr(g2), B, H, W, _stream())
    return g1, g2


def bce_logits(logits: torch.Tensor, targets: torch.Tensor, want_grad: bool = True):
    """binary_cross_entropy_with_logits, mean reduction: (loss [1], d loss / d logits or None)."""
    _need(logits, _F32, "logits")
    _need(targets, _F32, "targets")
    if logits.shape != targets.shape:
        raise _lib.AqualoraError("bce_logits: logits and targets must have equal shapes")
    x, y = logits.contiguous(), targets.contiguous()
    loss = torch.empty(1, dtype=_F32, device=x.device)
    g = torch.empty_like(x) if want_grad else None
    _lib.call("aq_bce_logits", x.data_ptr(), y.data_ptr(), loss.data_ptr(), _ptr(g), x.numel(), _stream())
    return loss, g


# ------------------------------------------------------------------------------------------------------------------
# message decoder (EfficientNet-B1, utils/models.py:84-96)
# ------------------------------------------------------------------------------------------------------------------
_DECODER_WS: dict = {}


def effnetb1_fwd(x: torch.Tensor, packed: torch.Tensor, out_features: int, want_bits: bool = True):
    """x [B, 3, 512, 512] fp32 -> (logits [B, out_features] fp32, bits [B, out_features // 2] uint8 or None)."""
    x = _image(x, "x")
    _need(packed, _F32, "packed", 1)
    B = x.shape[0]
    if tuple(x.shape[2:]) != (512, 512):
        raise _lib.AqualoraError(f"decoder input must be 512 x 512 (resize first), got {tuple(x.shape)}")
    lib = _lib.load()
    if packed.numel() != lib.aq_effnetb1_packed_floats(out_features):
        raise _lib.AqualoraError(f"packed decoder weights: {packed.numel()} floats, expected {lib.aq_effnetb1_packed_floats(out_features)}")
    nbytes = lib.aq_effnetb1_workspace_bytes(B)
    key = (x.device.index, torch.cuda.current_stream().cuda_stream)
    ws = _DECODER_WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)   # reused across calls on the same stream
        _DECODER_WS[key] = ws
    logits = torch.empty((B, out_features), dtype=_F32, device=x.device)
    bits = torch.empty((B, out_features // 2), dtype=torch.uint8, device=x.device) if want_bits else None
    _lib.call("aq_effnetb1_fwd", x.data_ptr(), packed.data_ptr(), logits.data_ptr(), _ptr(bits), B, out_features, ws.data_ptr(),
              ws.numel(), _stream())
    return logits, bits


def _nhwc_f32(x: torch.Tensor, name: str):
    """(M, C) of a 4-D fp32 tensor stored channels_last (memory = [B, H, W, C] rows)."""
    _need(x, _F32, name, 4)
    if not x.is_contiguous(memory_format=torch.channels_last):
        raise _lib.AqualoraError(f"{name} must be stored channels_last, got strides {x.stride()}")
    return x.shape[0] * x.shape[2] * x.shape[3], x.shape[1]


def bn_train_fwd(z: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, running_mean: torch.Tensor | None,
                 running_var: torch.Tensor | None, eps: float, momentum: float, act: bool):
    """Batch-statistics BatchNorm2d (+ SiLU) over a channels_last fp32 tensor; returns (y, mean_rstd [C, 2]); running stats updated."""
    M, C = _nhwc_f32(z, "z")
    y = torch.empty_like(z)
    mean_rstd = torch.empty(C, 2, dtype=_F32, device=z.device)
    nbytes = _lib.load().aq_bn_train_workspace_bytes(C)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=z.device)
    _lib.call("aq_bn_train_fwd", z.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _ptr(running_mean), _ptr(running_var), mean_rstd.data_ptr(),
              y.data_ptr(), M, C, float(eps), float(momentum), int(act), ws.data_ptr(), nbytes, _stream())
    return y, mean_rstd


def bn_train_bwd(gy: torch.Tensor, z: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, mean_rstd: torch.Tensor, act: bool):
    """Returns (gz, g_gamma, g_beta)."""
    M, C = _nhwc_f32(z, "z")
    if tuple(_nhwc_f32(gy, "gy")) != (M, C):
        raise _lib.AqualoraError("gy must have the shape and layout of z")
    gz = torch.empty_like(z)
    g_gamma = torch.zeros(C, dtype=_F32, device=z.device)
    g_beta = torch.zeros(C, dtype=_F32, device=z.device)
    nbytes = _lib.load().aq_bn_train_workspace_bytes(C)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=z.device)
    _lib.call("aq_bn_train_bwd", gy.data_ptr(), z.data_ptr(), gamma.data_ptr(), beta.data_ptr(), mean_rstd.data_ptr(), gz.data_ptr(),
              g_gamma.data_ptr(), g_beta.data_ptr(), M, C, int(act), ws.data_ptr(), nbytes, _stream())
    return gz, g_gamma, g_beta


def dwconv_fwd(x: torch.Tensor, w: torch.Tensor, stride: int) -> torch.Tensor:
    """Depthwise k x k convolution (no bias) of a channels_last fp32 tensor; w [C, 1, k, k].  Returns z channels_last."""
    _, C = _nhwc_f32(x, "x")
    _need(w, _F32, "w", 4)
    B, _, H, W = x.shape
    k = w.shape[-1]
    p = (k - 1) // 2
    Ho, Wo = (H + 2 * p - k) // stride + 1, (W + 2 * p - k) // stride + 1
    z = torch.empty((B, C, Ho, Wo), dtype=_F32, device=x.device).contiguous(memory_format=torch.channels_last)
    _lib.call("aq_dwconv_fwd", x.data_ptr(), w.contiguous().data_ptr(), z.data_ptr(), B, H, W, C, k, int(stride), _stream())
    return z


def dwconv_bwd(gz: torch.Tensor, x: torch.Tensor, w: torch.Tensor, stride: int, need_gx: bool = True):
    """Returns (gx or None, gw [C, 1, k, k])."""
    _, C = _nhwc_f32(x, "x")
    _nhwc_f32(gz, "gz")
    B, _, H, W = x.shape
    k = w.shape[-1]
    gx = torch.empty_like(x) if need_gx else None
    gw = torch.zeros_like(w, memory_format=torch.contiguous_format)
    _lib.call("aq_dwconv_bwd", gz.data_ptr(), x.data_ptr(), w.contiguous().data_ptr(), _ptr(gx), gw.data_ptr(), B, H, W, C, k, int(stride), _stream())
    return gx, gw


def conv1x1_wgrad(gz2d: torch.Tensor, x2d: torch.Tensor, se: torch.Tensor | None, hw: int) -> torch.Tensor:
    """gw [N, K] = gz2d^T [N, M] @ (x2d * se[row // hw]) [M, K]  (fp32)."""
    _need(gz2d, _F32, "gz", 2)
    _need(x2d, _F32, "x", 2)
    M, N = gz2d.shape
    K = x2d.shape[1]
    if not gz2d.is_contiguous() or not x2d.is_contiguous() or x2d.shape[0] != M:
        raise _lib.AqualoraError("conv1x1_wgrad: gz [M, N] and x [M, K] must be contiguous with equal row counts")
    gw = torch.zeros((N, K), dtype=_F32, device=x2d.device)
    _lib.call("aq_conv1x1_wgrad", gz2d.data_ptr(), x2d.data_ptr(), _ptr(se), gw.data_ptr(), M, K, N, int(hw), _stream())
    return gw


def stem_conv_fwd(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """3 x 3 stride-2 stem (3 -> 32, no bias): x [B, 3, H, W] NCHW fp32 -> z [B, 32, Ho, Wo] channels_last."""
    x = _image(x, "x")
    _need(w, _F32, "w", 4)
    B, _, H, W = x.shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    z = torch.empty((B, 32, Ho, Wo), dtype=_F32, device=x.device).contiguous(memory_format=torch.channels_last)
    _lib.call("aq_stem_conv_fwd", x.data_ptr(), w.contiguous().data_ptr(), z.data_ptr(), B, H, W, _stream())
    return z


def stem_conv_bwd(gz: torch.Tensor, x: torch.Tensor, w: torch.Tensor, need_gx: bool):
    x = _image(x, "x")
    _nhwc_f32(gz, "gz")
    B, _, H, W = x.shape
    gx = torch.empty_like(x) if need_gx else None
    gw = torch.zeros_like(w, memory_format=torch.contiguous_format)
    _lib.call("aq_stem_conv_bwd", gz.data_ptr(), x.data_ptr(), w.contiguous().data_ptr(), _ptr(gx), gw.data_ptr(), B, H, W, _stream())
    return gx, gw


def conv1x1_tf32x3(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, se: torch.Tensor | None = None,
                   residual: torch.Tensor | None = None, hw: int = 0, epi: int = 0) -> torch.Tensor:
    """Pointwise convolution over NHWC pixels on the tensor cores with the fp32-faithful 3-term TF32 split
    (csrc/decoder_pw.cu).  x [M, K] fp32, w [N, K] fp32 (split here), bias [N]; epi 0 none / 1 SiLU / 2 + residual /
    3 SiLU + per-image column sums (returns [M // hw, N])."""
    _need(x, _F32, "x", 2)
    _need(w, _F32, "w", 2)
    M, K = x.shape
    N = w.shape[0]
    w = w.contiguous()
    hi = (w.view(torch.int32) & -8192).view(torch.float32)
    lo = w - hi
    hw = hw or M
    y = torch.zeros((M // hw, N), dtype=_F32, device=x.device) if epi == 3 else torch.empty((M, N), dtype=_F32, device=x.device)
    _lib.call("aq_conv1x1_tf32x3", x.contiguous().data_ptr(), hi.data_ptr(), lo.data_ptr(), bias.contiguous().data_ptr(), _ptr(se),
              _ptr(residual), y.data_ptr(), M, K, N, hw, epi, _stream())
    return y


def depthwise_silu(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, k: int, stride: int):
    """SiLU(depthwise k x k conv + bias) over NHWC x [B, H, H, C]; w [k * k, C].  Returns (y [B, Ho, Ho, C], pooled sums [B, C])."""
    _need(x, _F32, "x", 4)
    B, H, _, C = x.shape
    ho = (H + 2 * ((k - 1) // 2) - k) // stride + 1
    y = torch.empty((B, ho, ho, C), dtype=_F32, device=x.device)
    pooled = torch.zeros((B, C), dtype=_F32, device=x.device)
    _lib.call("aq_depthwise_silu", x.contiguous().data_ptr(), w.contiguous().data_ptr(), bias.contiguous().data_ptr(), y.data_ptr(),
              pooled.data_ptr(), B, H, C, k, stride, _stream())
    return y, pooled


def expand_dw_fused(x: torch.Tensor, w_e: torch.Tensor, b_e: torch.Tensor, w_d: torch.Tensor, b_d: torch.Tensor, k: int, stride: int):
    """Front half of an MBConv block in one kernel (csrc/decoder_fused.cu): SiLU(depthwise_k,s(SiLU(x W_e^T + b_e)) + b_d) over NHWC
    x [B, H, H, cin]; w_e [cexp, cin] (split into TF32 hi / lo here), w_d [k * k, cexp].  Returns (y [B, Ho, Ho, cexp], pooled sums
    [B, cexp]).  Only the early EfficientNet-B1 shapes have a fused kernel; others raise AqualoraError (BAD_SHAPE)."""
    _need(x, _F32, "x", 4)
    _need(w_e, _F32, "w_e", 2)
    B, H, _, cin = x.shape
    cexp = w_e.shape[0]
    ho = (H + 2 * ((k - 1) // 2) - k) // stride + 1
    w_e = w_e.contiguous()
    hi = (w_e.view(torch.int32) & -8192).view(torch.float32)
    lo = w_e - hi
    y = torch.empty((B, ho, ho, cexp), dtype=_F32, device=x.device)
    pooled = torch.zeros((B, cexp), dtype=_F32, device=x.device)
    _lib.call("aq_expand_dw_fused", x.contiguous().data_ptr(), hi.data_ptr(), lo.data_ptr(), b_e.contiguous().data_ptr(),
              w_d.contiguous().data_ptr(), b_d.contiguous().data_ptr(), y.data_ptr(), pooled.data_ptr(), B, H, cin, cexp, k, stride, _stream())
    return y, pooled


def lora_fold_down(down: torch.Tensor, m: torch.Tensor, scale: float) -> torch.Tensor:
    """down' = (down * m[:, None...]) * scale for a [r, ...] fp32 LoRA down weight (scripts/create_wm_lora.py:30-37)."""
    _need(down, _F32, "down")
    _need(m, _F32, "m", 1)
    r = down.shape[0]
    if m.numel() != r:
        raise _lib.AqualoraError(f"mapper output has {m.numel()} entries, LoRA rank is {r}")
    d = down.contiguous()
    out = torch.empty_like(d)
    _lib.call("aq_lora_fold_down", d.data_ptr(), m.contiguous().data_ptr(), out.data_ptr(), r, d.numel() // r, float(scale), _stream())
    return out


def lora_merge_(w: torch.Tensor, up: torch.Tensor, down: torch.Tensor, coef: float) -> torch.Tensor:
    """In place: w [dout, din] += coef * (up [dout, r] @ down [r, din]), fp32 (scripts/merge_lora.py:107-120)."""
    _need(w, _F32, "w")
    _need(up, _F32, "up")
    _need(down, _F32, "down")
    dout, r = up.shape[0], down.shape[0]
    din = down.numel() // r
    if not w.is_contiguous() or w.numel() != dout * din or up.numel() != dout * r:
        raise _lib.AqualoraError(f"merge shapes: w {tuple(w.shape)}, up {tuple(up.shape)}, down {tuple(down.shape)}")
    _lib.call("aq_lora_merge", w.data_ptr(), up.contiguous().data_ptr(), down.contiguous().data_ptr(), dout, din, r, float(coef), _stream())
    return w


# ------------------------------------------------------------------------------------------------
# GroupNorm (+ add, + SiLU) on channels-last bf16 rows, GEGLU
# ------------------------------------------------------------------------------------------------
def _nhwc_rows(x: torch.Tensor, name: str) -> tuple[int, int, int]:
    """(B, HW, C) of an NCHW-shaped tensor stored channels_last, or of a [B, HW, C] contiguous tensor."""
    _need(x, _BF16, name)
    if x.dim() == 4:
        if not x.is_contiguous(memory_format=torch.channels_last):
            raise _lib.AqualoraError(f"{name} must be stored channels_last, got strides {x.stride()}")
        return x.shape[0], x.shape[2] * x.shape[3], x.shape[1]
    if x.dim() == 3 and x.is_contiguous():
        return x.shape[0], x.shape[1], x.shape[2]
    raise _lib.AqualoraError(f"{name} must be [B, C, H, W] channels_last or contiguous [B, HW, C], got {tuple(x.shape)}")


def group_norm_nhwc_fwd(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, groups: int, eps: float, silu: bool,
                        add_bc: torch.Tensor | None = None):
    """y = [silu](group_norm(x + add_bc[:, :, None, None])) ; returns (y, mean_rstd [B, G, 2] fp32)."""
    B, HW, C = _nhwc_rows(x, "x")
    _need(gamma, _BF16, "gamma", 1)
    _need(beta, _BF16, "beta", 1)
    if gamma.shape[0] != C or beta.shape[0] != C:
        raise _lib.AqualoraError(f"gamma / beta must be [{C}]")
    if add_bc is not None:
        _need(add_bc, _BF16, "add_bc", 2)
        if tuple(add_bc.shape) != (B, C) or not add_bc.is_contiguous():
            raise _lib.AqualoraError(f"add_bc must be contiguous [{B}, {C}], got {tuple(add_bc.shape)}")
    y = torch.empty_like(x)
    stats = torch.empty(B, groups, 2, dtype=_F32, device=x.device)
    nbytes = _lib.load().aq_group_norm_workspace_bytes(B, groups)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    _lib.call("aq_group_norm_nhwc_fwd", x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _ptr(add_bc), y.data_ptr(), stats.data_ptr(),
              B, HW, C, groups, float(eps), int(silu), ws.data_ptr(), nbytes, _stream())
    return y, stats


def group_norm_nhwc_bwd(dy: torch.Tensor, x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, stats: torch.Tensor,
                        groups: int, eps: float, silu: bool, add_bc: torch.Tensor | None = None) -> torch.Tensor:
    B, HW, C = _nhwc_rows(x, "x")
    if tuple(_nhwc_rows(dy, "dy")) != (B, HW, C) or dy.stride() != x.stride():
        raise _lib.AqualoraError("dy must have the shape and strides of x")
    _need(stats, _F32, "mean_rstd", 3)
    dx = torch.empty_like(x)
    nbytes = _lib.load().aq_group_norm_workspace_bytes(B, groups)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    _lib.call("aq_group_norm_nhwc_bwd", dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _ptr(add_bc), stats.data_ptr(),
              dx.data_ptr(), B, HW, C, groups, float(eps), int(silu), ws.data_ptr(), nbytes, _stream())
    return dx


def geglu_fwd(proj: torch.Tensor) -> torch.Tensor:
    """proj [..., 2F] bf16 -> proj[..., :F] * gelu(proj[..., F:])."""
    _need(proj, _BF16, "proj")
    F2 = proj.shape[-1]
    p2 = proj.reshape(-1, F2)
    ldp = _rows(p2, "proj")
    out = torch.empty(*proj.shape[:-1], F2 // 2, dtype=_BF16, device=proj.device)
    _lib.call("aq_geglu_fwd", p2.data_ptr(), ldp, out.data_ptr(), p2.shape[0], F2 // 2, _stream())
    return out


def geglu_bwd(proj: torch.Tensor, g_out: torch.Tensor) -> torch.Tensor:
    _need(proj, _BF16, "proj")
    _need(g_out, _BF16, "g_out")
    F2 = proj.shape[-1]
    p2 = proj.reshape(-1, F2)
    ldp = _rows(p2, "proj")
    if not g_out.is_contiguous():
        raise _lib.AqualoraError("g_out must be contiguous")
    g_proj = torch.empty(proj.shape, dtype=_BF16, device=proj.device)
    _lib.call("aq_geglu_bwd", p2.data_ptr(), ldp, g_out.data_ptr(), g_proj.data_ptr(), p2.shape[0], F2 // 2, _stream())
    return g_proj


def layer_norm_fwd(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, save_stats: bool):
    """LayerNorm over the last dimension of contiguous bf16 rows; returns (y, mean_rstd [M, 2] fp32 or None)."""
    _need(x, _BF16, "x")
    _need(gamma, _BF16, "gamma", 1)
    _need(beta, _BF16, "beta", 1)
    C = x.shape[-1]
    if not x.is_contiguous() or gamma.shape[0] != C or beta.shape[0] != C:
        raise _lib.AqualoraError(f"layer_norm: x must be contiguous [..., {C}] with gamma / beta [{C}]")
    M = x.numel() // C
    y = torch.empty_like(x)
    stats = torch.empty(M, 2, dtype=_F32, device=x.device) if save_stats else None
    _lib.call("aq_layer_norm_fwd", x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), _ptr(stats), M, C, float(eps), _stream())
    return y, stats


def layer_norm_bwd(dy: torch.Tensor, x: torch.Tensor, gamma: torch.Tensor, stats: torch.Tensor) -> torch.Tensor:
    _need(dy, _BF16, "dy")
    _need(x, _BF16, "x")
    _need(stats, _F32, "mean_rstd", 2)
    C = x.shape[-1]
    if not dy.is_contiguous() or not x.is_contiguous() or dy.shape != x.shape:
        raise _lib.AqualoraError("layer_norm_bwd: dy and x must be contiguous and of equal shape")
    dx = torch.empty_like(x)
    _lib.call("aq_layer_norm_bwd", dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), stats.data_ptr(), dx.data_ptr(), x.numel() // C, C, _stream())
    return dx


def add_bias_nhwc(a: torch.Tensor, b: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """a + b + bias[None, :, None, None] for channels_last bf16 [B, C, H, W] tensors of equal shape."""
    B, HW, C = _nhwc_rows(a, "a")
    if tuple(_nhwc_rows(b, "b")) != (B, HW, C) or b.stride() != a.stride():
        raise _lib.AqualoraError("add_bias_nhwc: a and b must have the same shape and strides")
    _need(bias, _BF16, "bias", 1)
    if bias.shape[0] != C:
        raise _lib.AqualoraError(f"bias must be [{C}]")
    out = torch.empty_like(a)
    _lib.call("aq_add_bias_rows", a.data_ptr(), b.data_ptr(), bias.data_ptr(), out.data_ptr(), B * HW, C, _stream())
    return out
