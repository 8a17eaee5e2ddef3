"""PPFT step (train/ppft_train.py:987-1068) replayed as a CUDA graph vs the same step issued eagerly: same inputs, same
initial state, two consecutive steps (the second one exercises the in-graph refresh of the bf16 operand copies after AdamW).
The kernels of this repository are identical in both runs; the library calls in between (cuDNN convolutions, SDPA, cuBLAS) may
pick other algorithms on the capture stream, and fp32 atomics land in a different order.  The loss is an MSE between two nearly
equal bf16 forward passes, so it moves by a few 1e-3 relative under such reordering (measured 2.6e-3)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _trainer(dev):
    from aqualora_b200 import ppft
    from aqualora_b200.unet import UNetConfig
    from oracle import lora_oracle as O

    cfg = UNetConfig.tiny(16)
    unet = ppft.build_unet(cfg, dev, seed=3)
    emb = O.mapper_init(48, 8, generator=torch.Generator().manual_seed(5))
    return ppft.PPFTTrainer(unet, ppft.PPFTConfig(rank=8, msg_bits=48), emb, dev, lora_up_std=0.05, seed=1), cfg


def _batch(cfg, dev, seed, B=2):
    g = torch.Generator().manual_seed(seed)
    s = cfg.sample_size
    bf = torch.bfloat16
    f = lambda x: x.to(dev)
    return (f((torch.randn(B, 4, s, s, generator=g) * 0.18215).to(bf)), f((torch.randn(B, 4, s, s, generator=g) * 0.004).to(bf)),
            f(torch.randn(B, 4, s, s, generator=g).to(bf)), f(torch.randint(0, 1000, (B,), generator=g)),
            f(torch.randn(B, 77, cfg.cross_attention_dim, generator=g).to(bf)), f(torch.randint(0, 2, (B, 48), generator=g).float()))


def test_graphed_step_matches_eager(cuda_device):
    from aqualora_b200 import lora_modules

    lora_modules.clear_caches()
    eager, cfg = _trainer(cuda_device)
    p0 = eager.state.param.clone()
    batches = [_batch(cfg, cuda_device, 11), _batch(cfg, cuda_device, 12)]
    losses_e = [eager.step(*b).item() for b in batches]
    p_eager = eager.state.param.clone()

    lora_modules.clear_caches()
    graphed, _ = _trainer(cuda_device)
    assert torch.equal(graphed.state.param, p0)
    n = graphed.capture(graphed.forward_backward, batches[0], warmup_steps=1)
    assert n > 200                                           # 192 LoRA targets x (fwd + bwd) + glue kernels sit inside the graph
    # capture() warmed up with real optimizer steps: rewind the state so both runs start from the same point
    graphed.state.param.copy_(p0)
    graphed.state.exp_avg.zero_(); graphed.state.exp_avg_sq.zero_(); graphed.state.grad.zero_()
    graphed.global_step = 0
    graphed.refresh_operands()                               # the replay cannot notice that the master parameters were rewound
    losses_g = [graphed.step_graphed(*b).item() for b in batches]
    p_graph = graphed.state.param.clone()

    assert losses_g == pytest.approx(losses_e, rel=1e-2)
    moved = (p_eager - p0).abs().max().item()
    assert moved > 1e-5                                      # the two steps did update the parameters
    # AdamW's first steps move every element by ~lr * sign(g): elements whose gradient is at the noise level may flip, so the
    # updates are compared as directions, not element by element
    cos = torch.nn.functional.cosine_similarity((p_graph - p0).flatten(), (p_eager - p0).flatten(), dim=0).item()
    assert cos > 0.95, cos


FRO_TOL, COS_TOL, NORM_TOL = 5e-2, 0.999, 6e-2


def test_graph_replay_gradient_equals_eager_gradient(cuda_device):
    """The captured graph must produce the SAME flat gradient as the eager forward + backward on the same batch and parameters --
    compared on the gradient itself (before AdamW turns noise-level entries into +-lr steps, which is why the parameter-level test
    above can only compare directions).  This library's kernels are the same launches on the same operands in both paths; what differs
    is the library part of the U-Net (cuDNN convolution / attention algorithm selection on the capture stream vs eagerly, and their
    atomics-based backward) in bf16, plus the order of our fp32 atomics.  Measured on a B200: relative Frobenius error 2.5e-2, cosine
    0.9997, worst per-matrix norm deviation 2.9e-2 -- the level at which two eager runs differ too.  Bounds: Frobenius <= 5e-2,
    cosine >= 0.999, every one of the 384 per-matrix gradient norms within 6 %, loss within 1 %."""
    from aqualora_b200 import lora_modules

    lora_modules.clear_caches()
    tr, cfg = _trainer(cuda_device)
    batch = _batch(cfg, cuda_device, 21)
    tr.state.grad.zero_()
    if tr.g_scale is not None:
        tr.g_scale.zero_()
    loss_e = tr.forward_backward(*batch).item()
    torch.cuda.synchronize()
    g_eager = tr.state.grad.clone()
    assert g_eager.abs().max().item() > 0

    tr.capture(tr.forward_backward, batch, warmup_steps=1)
    tr.state.grad.zero_()
    if tr.g_scale is not None:
        tr.g_scale.zero_()
    for dst, src in zip(tr.static_inputs, batch):
        dst.copy_(src)
    tr._graph.replay()
    torch.cuda.synchronize()
    g_graph = tr.state.grad.clone()
    loss_g = float(tr._static_loss)

    fro = ((g_graph - g_eager).norm() / g_eager.norm()).item()
    cos = torch.nn.functional.cosine_similarity(g_graph, g_eager, dim=0).item()
    worst = 0.0
    off = 0
    for p_ in tr.state.params:
        n = p_.numel()
        a, b = g_graph[off:off + n].norm().item(), g_eager[off:off + n].norm().item()
        if b > 0:
            worst = max(worst, abs(a / b - 1))
        off += n
    print("graph vs eager gradient:", {"loss": (loss_g, loss_e), "fro": fro, "cos": cos, "worst_norm_dev": worst})
    assert loss_g == pytest.approx(loss_e, rel=1e-2)
    assert fro <= FRO_TOL and cos >= COS_TOL, (fro, cos)
    assert worst <= NORM_TOL, worst
