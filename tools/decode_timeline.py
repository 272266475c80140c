"""Timeline of the stream GEMMs inside one captured decode step (globaltimer stamps of CTA 0 of every launch):
python tools/decode_timeline.py [groups] [pdl 0|1] [stages] [ctas]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from time_r1_b200 import config, engine, lib, weights  # noqa: E402
from tools.decode_ncu import make_engine  # noqa: E402


def main():
    n_groups = int(sys.argv[1]) if len(sys.argv) > 1 else 2
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
    nl = 4 * cfg.num_hidden_layers + 1
    for pdl, stages, ctas in ([(1, 0, 1)] if os.environ.get("T1_TIMELINE_PAIR") else [(0, 0, 1), (1, 0, 1), (1, 10, 1)]):
        run = config.RunConfig(num_generations=G, max_completion_length=200, beta=0.0, decode_impl="stream", decode_pdl=bool(pdl),
                               decode_pair=os.environ.get("T1_TIMELINE_PAIR", "0") == "1")     # paired: one record per GEMM all the same
        # (phase-B record of a pair: start = this CTA arrived at the counter, issued = z tiles requested, dep_ok = counter complete)
        eng = make_engine(cfg, mw, run, dev)
        lib.load().t1_stream_gemm_tune(stages, ctas)
        dbg = torch.zeros(nl * 4, device=dev, dtype=torch.int64)
        eng.rollout_multi(prompts, embs, G, 200, 1.0, 50, seed=1, use_graph=False, check_every=1000)     # warm KV, no graph yet
        st = eng._last_decode_state
        # capture ONE decode step with the debug ring attached: launch i of the step writes record i
        lib.load().t1_stream_gemm_debug(dbg.data_ptr(), nl)
        st["step"].fill_(100)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            eng._decode_step(st, G, 1.0, 50)
        lib.load().t1_stream_gemm_debug(0, 0)
        for _ in range(3):
            st["step"].fill_(100)
            g.replay()
        torch.cuda.synchronize()
        t = dbg.cpu().numpy().reshape(nl, 4).astype(np.int64)
        t0 = t[0, 0]
        rel = (t - t0) / 1e3
        names = ["qkv", "o", "gate_up", "down"]
        print(json.dumps({"pdl": pdl, "stages": stages, "ctas": ctas, "step_us": round(float(rel[-1, 3]), 1)}))
        for i in list(range(4, 12)) + [nl - 1]:
            nm = names[i % 4] if i < nl - 1 else "lm_head"
            prev_done = rel[i - 1, 3]
            print(f"  L{i // 4:02d} {nm:8s} start {rel[i, 0]:9.2f} issued {rel[i, 1]:9.2f} dep_ok {rel[i, 2]:9.2f} done {rel[i, 3]:9.2f} | start-prev_done {rel[i, 0] - prev_done:7.2f} dur_after_dep {rel[i, 3] - rel[i, 2]:7.2f} gap_prev_done->dep_ok {rel[i, 2] - prev_done:7.2f}")
        # per-type averages over layers 1..27
        for k, nm in enumerate(names):
            idx = np.arange(4 + k, nl - 1, 4)
            print(f"  avg {nm:8s} dur_after_dep {np.mean(rel[idx, 3] - rel[idx, 2]):7.2f}  gap_prev_gemm_done->dep_ok {np.mean(rel[idx, 2] - rel[idx - 1, 3]):7.2f}  start_lead {np.mean(rel[idx, 2] - rel[idx, 0]):7.2f}")
        del eng, st, g
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
