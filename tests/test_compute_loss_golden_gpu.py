"""GPU suite: the CUDA kernels of the loss path against outputs of the UNMODIFIED reference compute_loss
(tests/golden/compute_loss_golden.npz, oracle/gen_golden_compute_loss.py; timer1_trainer.py:575-777).

Chain under test, all through the C-ABI: completion texts as byte tokens -> t1_detokenize -> t1_parse_answers ->
t1_iou_reward (rewards of main.py:184-239) -> t1_group_advantage; t1_eos_mask; fp32 logits -> t1_logprob_entropy ->
t1_grpo_loss (loss, dL/dlogp, logged metrics) -> t1_logits_backward (bf16 dL/dlogits).
Bars: mask / rewards bit-exact; loss, advantage and metrics 2e-6 abs (fp32 reductions in another order than torch's);
dL/dlogits to bf16 rounding (2^-8 relative) because the kernel emits bf16.
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
from test_compute_loss_golden_cpu import CASES, oracle_rewards  # noqa: E402

DEV = "cuda"


def _byte_tokens(texts, C_text, eos=257, pad=256):
    ids = np.full((len(texts), C_text), pad, np.int32)
    for g, t in enumerate(texts):
        b = np.frombuffer(t.encode("utf-8"), np.uint8)
        ids[g, :len(b)] = b
        ids[g, len(b)] = eos
    return torch.from_numpy(ids).to(DEV)


@pytest.mark.parametrize("c", CASES, ids=lambda c: f"case{c['case']}")
def test_cuda_loss_path_matches_unmodified_compute_loss(c):
    from time_r1_b200 import ops, rewards
    from time_r1_b200.detok import DetokTable

    G, C, V = c["G"], c["C"], c["V"]
    # ---- rewards on the device from the decoded texts the reference scored
    table = DetokTable.from_byte_strings([bytes([b]) for b in range(256)] + [b"<|endoftext|>", b"<|im_end|>"], [256, 257], torch.device(DEV))
    funcs = [getattr(rewards, n) for n in c["reward_funcs"]]
    dr = rewards.DeviceRewards(funcs, table)
    rpf = dr(_byte_tokens(c["texts"], max(len(t.encode()) for t in c["texts"]) + 2), c["solution"], c["durations"])
    want = oracle_rewards(c)
    assert torch.equal(rpf.cpu(), want), (rpf.cpu(), want)                  # bit-exact fp32 rewards
    rew, adv, mean, std = ops.group_advantage(rpf, G)
    # ---- mask, log-probs, loss
    ids = torch.from_numpy(c["completion_ids"]).to(DEV)
    mask, lengths = ops.eos_mask(ids, c["eos_token_id"])
    from oracle import qwen_vl_oracle as O

    assert torch.equal(mask.cpu(), O.completion_mask(ids.cpu().long(), c["eos_token_id"]))
    logits = torch.from_numpy(c["logits"]).to(DEV).reshape(G * C, V).contiguous()
    ref_logits = torch.from_numpy(c["ref_logits"]).to(DEV).reshape(G * C, V).contiguous()
    tgt = ids.reshape(-1).contiguous()
    lp, ent, lse = ops.logprob_entropy(logits, tgt)
    rlp, _, _ = ops.logprob_entropy(ref_logits, tgt)
    loss, dlogp, stats = ops.grpo_loss(lp.view(G, C), rlp.view(G, C) if c["beta"] else None, ent.view(G, C), adv, mask, c["beta"], 0.2, 0.2,
                                       c["use_grpo"])
    dlogits = ops.logits_backward(logits, tgt, lse, dlogp.reshape(-1).contiguous())
    torch.cuda.synchronize()
    assert abs(loss.item() - c["loss"]) < 2e-6, (loss.item(), c["loss"])
    gm = c["metrics"]
    s = stats.tolist()
    assert abs(s[3] - gm["completion_length"]) < 1e-6
    assert abs(s[2] - gm["generation_entropy"]) < 2e-6
    if c["beta"]:
        assert abs(s[1] - gm["kl"]) < 2e-6
    assert abs(rew.mean().item() - gm["reward"]) < 1e-6 and abs(std.mean().item() - gm["reward_std"]) < 1e-6
    for i, name in enumerate(c["reward_funcs"]):
        assert abs(rpf[:, i].mean().item() - gm[f"rewards/{name}"]) < 1e-6
    got = dlogits.float().cpu().view(G, C, V).numpy()
    np.testing.assert_allclose(got, c["dlogits"], atol=1e-7, rtol=2.0 ** -7)
