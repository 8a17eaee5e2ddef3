"""Host-side logic of the N > 1 PPFT path on CPU: world_size-2 `gloo` processes exercise the flat gradient buffer layout,
the single allreduce (PPFTTrainer.exchange_gradients) and the per-process scheduler horizon (train/ppft_train.py:896-901).
The CUDA kernels themselves are covered by the -m gpu tests."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from aqualora_b200 import ppft
        from aqualora_b200.unet import UNetConfig

        cfg = UNetConfig.tiny(16)
        unet = ppft.build_unet(cfg, "cpu", dtype=torch.float32, seed=3)
        emb = torch.randn(48, 8, generator=torch.Generator().manual_seed(5))
        tr = ppft.PPFTTrainer(unet, ppft.PPFTConfig(rank=8, lr_warmup_steps=2, max_train_steps=10), emb, "cpu", lora_up_std=0.05, seed=1)
        assert tr.world == world
        st = tr.state
        # identical replicas: parameters live in ONE flat buffer, in unet_keys.json order, mapper last
        assert st.n_lora == sum(p.numel() for p in st.params) and len(st.params) == 2 * 192
        assert all(p.data_ptr() >= st.param.data_ptr() for p in st.params)
        # rank-dependent gradients written through the per-parameter views the kernels accumulate into
        for i, p in enumerate(st.params):
            p._aq_grad.fill_(float(rank + 1) * (1 + (i % 3)))
        st.mapper_grad.fill_(10.0 * (rank + 1))
        gs = tr.exchange_gradients()
        assert gs == 1.0 / world
        want = sum(r + 1 for r in range(world))
        for i, p in enumerate(st.params):
            assert torch.all(p._aq_grad == want * (1 + (i % 3)))
        assert torch.all(st.mapper_grad == 10.0 * want)
        # every rank holds the same buffer after the exchange
        ref = st.grad.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(ref, st.grad)
        # accelerate steps the scheduler once per process: horizon and warm-up are scaled by the world size
        tr.global_step = 1
        lr1 = tr.lr()
        assert abs(lr1 - tr.cfg.learning_rate * ppft.cosine_lr_factor(world, 2 * world, 10 * world)) < 1e-12
        out[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: "ok", 1: "ok"}


def test_rank_sharded_synthetic_batches_are_disjoint():
    """bench.py gives every rank its own data shard (seed offset by rank), like accelerate's sharded DataLoader."""
    import bench
    from aqualora_b200.unet import UNetConfig

    cfg = UNetConfig.tiny(16)
    a = bench.synth_batch(2, cfg, 1234 + 0 + 1000 * 0)
    b = bench.synth_batch(2, cfg, 1234 + 0 + 1000 * 1)
    assert not torch.equal(a[0], b[0]) and not torch.equal(a[4], b[4])
    assert torch.equal(a[0], bench.synth_batch(2, cfg, 1234)[0])
