"""Which scaled_dot_product_attention backend is fastest for the SD1.5 attention shapes of the PPFT step (library call, outside the
SURVEY 8(a) rows; 20 ms of the 77 ms step in round 1)?  Forward + backward, bf16, B = 16."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from torch.nn.attention import SDPBackend, sdpa_kernel

dev = torch.device("cuda:0")
B, H = 16, 8
shapes = [(4096, 4096, 40), (4096, 77, 40), (1024, 1024, 80), (1024, 77, 80), (256, 256, 160), (256, 77, 160), (64, 64, 160), (64, 77, 160)]
backends = {"cudnn": SDPBackend.CUDNN_ATTENTION, "flash": SDPBackend.FLASH_ATTENTION, "efficient": SDPBackend.EFFICIENT_ATTENTION, "math": SDPBackend.MATH}


def bench(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for (nq, nk, d) in shapes:
    row = []
    for layout in ("bhnd_view", "contiguous"):
        qb = torch.randn(B, nq, H * d, device=dev, dtype=torch.bfloat16, requires_grad=True)
        kb = torch.randn(B, nk, H * d, device=dev, dtype=torch.bfloat16, requires_grad=True)
        vb = torch.randn(B, nk, H * d, device=dev, dtype=torch.bfloat16, requires_grad=True)
        for name, be in backends.items():
            if name == "math" and nq * nk > 4096 * 1024:
                continue

            def step():
                q = qb.view(B, nq, H, d).transpose(1, 2)
                k = kb.view(B, nk, H, d).transpose(1, 2)
                v = vb.view(B, nk, H, d).transpose(1, 2)
                if layout == "contiguous":
                    q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
                o = F.scaled_dot_product_attention(q, k, v)
                o.transpose(1, 2).reshape(B, nq, H * d).backward(qb.detach())

            try:
                with sdpa_kernel([be]):
                    ms = bench(step)
                row.append(f"{layout[:4]}/{name}={ms * 1e3:.0f}us")
            except Exception as e:
                row.append(f"{layout[:4]}/{name}=n/a")
        # default dispatch
        def step_default():
            q = qb.view(B, nq, H, d).transpose(1, 2)
            k = kb.view(B, nk, H, d).transpose(1, 2)
            v = vb.view(B, nk, H, d).transpose(1, 2)
            o = F.scaled_dot_product_attention(q, k, v)
            o.transpose(1, 2).reshape(B, nq, H * d).backward(qb.detach())
        if layout == "bhnd_view":
            row.append(f"default={bench(step_default) * 1e3:.0f}us")
    print(f"nq={nq:5d} nk={nk:5d} d={d:3d}: " + "  ".join(row), flush=True)
