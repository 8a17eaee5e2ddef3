"""Checkpoint artefacts (train/ppft_train.py:443-471, :699-748, :1203-1229) and the bit-accuracy / TPR metric
(evaluation/utils_eval.py:131-140, :193-211) of the product, on CPU tensors (host-side plumbing, no kernels)."""
import os
import random

import pytest
import torch

from aqualora_b200 import checkpoint, evaluation, ppft
from aqualora_b200.unet import UNetConfig
from oracle import deploy_oracle as DO
from oracle import models_oracle as MO


def _trainer(seed=3, rank=8):
    cfg = UNetConfig.tiny(16)
    unet = ppft.build_unet(cfg, "cpu", dtype=torch.float32, seed=seed)
    emb = torch.randn(48, rank, generator=torch.Generator().manual_seed(seed))
    return ppft.PPFTTrainer(unet, ppft.PPFTConfig(rank=rank), emb, "cpu", lora_up_std=0.05, seed=seed)


def test_save_writes_the_reference_file_formats(tmp_path, golden_dir):
    from safetensors.torch import load_file

    tr = _trainer()
    tr.save(tmp_path)
    sd = load_file(os.path.join(tmp_path, "pytorch_lora_weights.safetensors"))
    # 16 transformer blocks x 12 targets x (down, up), every key under `unet.` with the diffusers AttnProcessor naming
    assert len(sd) == 2 * len(tr.keys) == 384
    assert all(k.startswith("unet.") and k.endswith((".down.weight", ".up.weight")) for k in sd)
    assert "unet.down_blocks.0.attentions.0.transformer_blocks.0.attn1.processor.to_q_lora.down.weight" in sd
    assert "unet.mid_block.attentions.0.transformer_blocks.0.ff.net.2.lora.up.weight" in sd
    assert "unet.up_blocks.1.attentions.2.proj_in.lora.down.weight" in sd
    assert sd["unet.up_blocks.1.attentions.2.proj_in.lora.down.weight"].dim() == 4           # SD1.5 proj_in is a 1x1 conv
    assert all(v.dtype == torch.float32 for v in sd.values())
    mp = torch.load(os.path.join(tmp_path, "mapper.pt"))
    assert list(mp) == ["bit_embeddings.weight"] and tuple(mp["bit_embeddings.weight"].shape) == (48, 8)
    # the file is what the reference's create_wm_lora.py consumes: fold a message with the oracle restatement of that script
    bits = "".join(random.Random(0).choice("01") for _ in range(48))
    folded = DO.fold_message(sd, mp["bit_embeddings.weight"], bits, 1.03)
    assert set(folded) == set(sd)
    k = "unet.down_blocks.0.attentions.0.transformer_blocks.0.attn1.processor.to_q_lora"
    assert torch.equal(folded[k + ".up.weight"], sd[k + ".up.weight"]) and not torch.equal(folded[k + ".down.weight"], sd[k + ".down.weight"])


def test_save_load_roundtrip_and_resume(tmp_path):
    a = _trainer(seed=3)
    a.state.exp_avg.normal_(generator=torch.Generator().manual_seed(1))
    a.state.exp_avg_sq.uniform_(generator=torch.Generator().manual_seed(2))
    a.global_step = 7
    path = a.save_state(os.fspath(tmp_path))
    assert os.path.basename(path) == "checkpoint-7"
    b = _trainer(seed=4)
    assert not torch.equal(a.state.param, b.state.param)
    assert b.load_state(os.path.join(tmp_path, "latest")) == 7
    assert torch.equal(a.state.param, b.state.param)            # loaded IN PLACE into the flat buffer (parameters stay views of it)
    assert torch.equal(a.state.exp_avg, b.state.exp_avg) and torch.equal(a.state.exp_avg_sq, b.state.exp_avg_sq)
    assert b.state.params[0].data_ptr() == b.state.param.data_ptr()
    # --checkpoints_total_limit rotation (train/ppft_train.py:1083-1099)
    for step in (8, 9, 10):
        a.global_step = step
        a.save_state(os.fspath(tmp_path), total_limit=2)
    assert sorted(d for d in os.listdir(tmp_path) if d.startswith("checkpoint-")) == ["checkpoint-10", "checkpoint-9"]
    assert checkpoint.latest_checkpoint(os.fspath(tmp_path)).endswith("checkpoint-10")


def test_load_rejects_foreign_or_incomplete_files(tmp_path):
    from safetensors.torch import load_file, save_file

    a = _trainer()
    a.save(tmp_path)
    f = os.path.join(tmp_path, "pytorch_lora_weights.safetensors")
    sd = load_file(f)
    k = next(iter(sd))
    short = {kk: v for kk, v in sd.items() if kk != k}
    with pytest.raises(KeyError):
        checkpoint.load_lora_into_unet(short, a.unet, a.keys)
    assert checkpoint.load_lora_into_unet(short, a.unet, a.keys, strict=False) == 383
    bad = dict(sd)
    bad[k] = torch.zeros(3, 3)
    with pytest.raises(ValueError):
        checkpoint.load_lora_into_unet(bad, a.unet, a.keys)


def test_threshold_known_answers_and_oracle_agreement():
    # get_threshold(48, 1e-6) = 40, (48, 1e-3) = 35
    assert evaluation.get_threshold(48, 1e-6) == 40 and evaluation.get_threshold(48, 1e-3) == 35
    for k in (16, 32, 48, 64):
        for fpr in (1e-2, 1e-3, 1e-6):
            assert evaluation.get_threshold(k, fpr) == MO.get_threshold(k, fpr)
            tau = evaluation.get_threshold(k, fpr)
            assert evaluation.calculate_fpr(tau, k) <= fpr < evaluation.calculate_fpr(tau - 1, k)


def test_bit_accuracy_matches_the_reference_loop():
    g = torch.Generator().manual_seed(0)
    gt = torch.randint(0, 2, (48,), generator=g)
    pred = gt.repeat(50, 1)
    flip = torch.rand(50, 48, generator=g) < torch.linspace(0, 0.5, 50)[:, None]
    pred = pred ^ flip.long()
    acc, tpr, per = evaluation.bit_accuracy(pred, "".join(map(str, gt.tolist())), 1e-3)
    # evaluation/utils_eval.py:197-211 restated literally
    tau = evaluation.get_threshold(48, 1e-3) / 48
    accs, tp = [], 0
    gts = "".join(map(str, gt.tolist()))
    for row in pred.tolist():
        msg = "".join(map(str, row))
        a = sum(1 for i in range(48) if msg[i] == gts[i]) / 48
        accs.append(a)
        tp += a >= tau
    assert acc == pytest.approx(sum(accs) / 50, abs=1e-12) and tpr == tp / 50
    assert per.tolist() == pytest.approx(accs)
    # per-image ground truth
    acc2, tpr2, _ = evaluation.bit_accuracy(pred, pred.clone(), 1e-6)
    assert acc2 == 1.0 and tpr2 == 1.0
