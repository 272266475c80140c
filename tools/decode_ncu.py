"""Eager (un-captured) decode steps at 7B dims for an ncu launch list: python tools/decode_ncu.py [groups] [impl]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from time_r1_b200 import config, engine, indexing, lib, weights  # noqa: E402


def make_engine(cfg, mw, run, dev):
    eng = engine.Engine.__new__(engine.Engine)      # decode only: skip the optimizer-state allocation
    eng.cfg, eng.w, eng.run, eng.dev = cfg, mw, run, dev
    eng.ref = None
    eng._decode_cache = {}
    eng.metrics_launches = 0
    eng.replay_hook = None
    hd = cfg.head_dim
    eng.inv_freq_llm = (1.0 / (cfg.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd))).to(dev)
    eng.axis_llm = torch.from_numpy(indexing.mrope_channel_axes(cfg)).to(dev)
    return eng


def main():
    n_groups = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    impl = sys.argv[2] if len(sys.argv) > 2 else "stream"
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    dev = torch.device("cuda", 0)
    cfg = config.PRESETS["qwen2.5-vl-7b"]()
    G = 8
    mw = weights.ModelWeights(cfg, dev, world_size=1)
    mw.init_random(seed=0)
    rng = np.random.default_rng(0)
    nv = 4096
    prompts, embs = [], []
    for i in range(n_groups):
        text = rng.integers(0, 151643, 110)
        ids = np.concatenate([[151644, 872, 198, 151652], np.full(nv, cfg.video_token_id), [151653], text, [151645, 198, 151644, 77091, 198]]).astype(np.int64)
        prompts.append(engine.Prompt(input_ids=ids, pixel_values=None, grid_thw=[(16, 32, 32)], second_per_grid_ts=[1.0]))
        embs.append((torch.randn(nv, cfg.hidden_size, device=dev) * 0.02).to(torch.bfloat16))
    run = config.RunConfig(num_generations=G, max_completion_length=200, beta=0.0, decode_impl=impl, decode_pdl=False)
    eng = make_engine(cfg, mw, run, dev)
    eng.rollout_multi(prompts, embs, G, 200, 1.0, 50, seed=1, use_graph=True)      # fills the KV cache to step 199
    st = eng._last_decode_state
    torch.cuda.synchronize()
    st["step"].fill_(100)
    torch.cuda.cudart().cudaProfilerStart()
    for _ in range(steps):
        eng._decode_step(st, G, 1.0, 50)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
