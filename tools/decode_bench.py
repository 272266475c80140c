"""Decode-step benchmark at 7B dims: ms per captured decode step (all 28 layers + lm_head + sampler) for the legacy and
stream/PDL implementations, plus in-graph per-GEMM costs.  Usage: python tools/decode_bench.py [groups] [C]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from time_r1_b200 import config, engine, lib, ops, weights  # noqa: E402


def time_replays(st, reps=30, step_val=100):
    g = st["graph"]
    ts = []
    for _ in range(reps):
        st["step"].fill_(step_val)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 8)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    n_groups = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    C = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    dev = torch.device("cuda", 0)
    cfg = config.PRESETS["qwen2.5-vl-7b"]()
    G = 8
    mw = weights.ModelWeights(cfg, dev, world_size=1)
    mw.init_random(seed=0)
    rng = np.random.default_rng(0)
    nv = int(os.environ.get("T1_BENCH_NV", "4096"))       # vision tokens of the prompt (cfg2 4096, cfg3 8192, cfg5 16384)
    prompts, embs = [], []
    for i in range(n_groups):
        text = rng.integers(0, 151643, 110)
        ids = np.concatenate([[151644, 872, 198, 151652], np.full(nv, cfg.video_token_id), [151653], text, [151645, 198, 151644, 77091, 198]]).astype(np.int64)
        prompts.append(engine.Prompt(input_ids=ids, pixel_values=None, grid_thw=[(nv // 256, 32, 32)], second_per_grid_ts=[1.0]))
        embs.append((torch.randn(nv, cfg.hidden_size, device=dev) * 0.02).to(torch.bfloat16))
    variants = [("legacy", False, 0, 1), ("stream", False, 0, 1), ("stream", True, 0, 1), ("stream", True, 4, 1), ("stream", True, 10, 1),
                ("stream", True, 5, 2), ("stream", True, 3, 2)]
    if os.environ.get("T1_BENCH_MAX_CTAS"):
        variants = [("stream", True, 0, 1)]
        lib.load().t1_stream_gemm_max_ctas(int(os.environ["T1_BENCH_MAX_CTAS"]))
    # T1_BENCH_VARIANTS='[{"decode_attn": "wide"}, {"decode_prefetch_mb": 64}]': RunConfig overrides, one stream/PDL run each
    overrides = [None] * len(variants)
    if os.environ.get("T1_BENCH_VARIANTS"):
        overrides = json.loads(os.environ["T1_BENCH_VARIANTS"])
        variants = [("stream", True, 0, 1)] * len(overrides)
    for (impl, pdl, stages, ctas), ov in zip(variants, overrides):
        run = config.RunConfig(num_generations=G, max_completion_length=C, beta=0.0, decode_impl=impl, decode_pdl=pdl)
        for k, v in (ov or {}).items():
            if k == "env":          # environment switches read by the library at launch / capture time
                os.environ.update({a: str(b) for a, b in v.items()})
                continue
            assert hasattr(run, k), k
            setattr(run, k, v)
        if os.environ.get("T1_DECODE_FOLD"):
            run.decode_fold_norm = os.environ["T1_DECODE_FOLD"] == "1"
        eng = engine.Engine.__new__(engine.Engine)      # decode only: skip the optimizer-state allocation
        eng.cfg, eng.w, eng.run, eng.dev = cfg, mw, run, dev
        eng.ref = None
        eng._decode_cache = {}
        eng.metrics_launches = 0
        eng.replay_hook = None
        hd = cfg.head_dim
        eng.inv_freq_llm = (1.0 / (cfg.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd))).to(dev)
        from time_r1_b200 import indexing
        eng.axis_llm = torch.from_numpy(indexing.mrope_channel_axes(cfg)).to(dev)
        lib.load().t1_stream_gemm_tune(stages, ctas)
        outs = eng.rollout_multi(prompts, embs, G, C, 1.0, 50, seed=1)
        st = eng._last_decode_state
        ms = time_replays(st)
        wbytes = 2.0 * (sum(p.numel() for n, p in mw.named() if n.startswith("model.layers.") and n.endswith("weight") and p.dim() == 2) + cfg.vocab_size * cfg.hidden_size) if hasattr(mw, "named") else 14.14e9
        print(json.dumps({"overrides": ov, "impl": impl, "pdl": pdl, "stages": stages, "ctas_per_sm": ctas, "groups": n_groups, "ms_per_decode_step": round(ms, 4),
                          "weight_gbs": round(wbytes / ms / 1e6, 1), "launches_per_step": st["launches_per_step"],
                          "tok0": [int(o[0, 1]) for o in outs]}), flush=True)
        del eng, st
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
