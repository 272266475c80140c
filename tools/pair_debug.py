"""Debug: paired decode launches vs the five-launch chain, graph vs eager, on the toy config."""
import dataclasses
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_engine_gpu import build  # noqa: E402
from time_r1_b200 import engine  # noqa: E402

O, cfg, ocfg, w_or, eng, prompt, pix, ids, grid = build(beta=0.0)
G, C = 4, 24
emb, _, _ = eng.vit_forward(prompt)
res = {}
for pair in (False, True):
    for graph in (True, False):
        e2 = engine.Engine(cfg, eng.w, dataclasses.replace(eng.run, decode_fold_norm=True, decode_pair=pair), with_ref=False)
        for rep in range(2):
            out = e2.rollout(prompt, emb, G, C, temperature=0.0, top_k=0, seed=1, use_graph=graph).cpu()
            res[(pair, graph, rep)] = out
            print("pair", pair, "graph", graph, "rep", rep, out[0].tolist(), "rows equal", bool((out == out[0:1]).all()), flush=True)
base = res[(False, True, 0)]
for k, v in res.items():
    print(k, "== base" if torch.equal(v, base) else "DIFFERS from base at %s" % (v != base).nonzero()[:3].tolist())
