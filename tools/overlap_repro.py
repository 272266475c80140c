"""Repro: which side-stream kernel cannot run concurrently with the captured decode step?  python tools/overlap_repro.py [attn|gemm|norm]"""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from time_r1_b200 import config, engine, lib, ops, weights
from tools.decode_ncu import make_engine
from test_kernels_gpu import make_mask, rnd

what = sys.argv[1] if len(sys.argv) > 1 else "attn"
dev = torch.device("cuda", 0)
cfg = config.PRESETS["qwen2.5-vl-7b"]()
G = 8
mw = weights.ModelWeights(cfg, dev, world_size=1); mw.init_random(seed=0)
rng = np.random.default_rng(0)
nv = 4096
prompts, embs = [], []
for i in range(2):
    text = rng.integers(0, 151643, 110)
    ids = np.concatenate([[151644, 872, 198, 151652], np.full(nv, cfg.video_token_id), [151653], text, [151645, 198, 151644, 77091, 198]]).astype(np.int64)
    prompts.append(engine.Prompt(input_ids=ids, pixel_values=None, grid_thw=[(16, 32, 32)], second_per_grid_ts=[1.0]))
    embs.append((torch.randn(nv, cfg.hidden_size, device=dev) * 0.02).to(torch.bfloat16))
run = config.RunConfig(num_generations=G, max_completion_length=200, beta=0.0)
eng = make_engine(cfg, mw, run, dev)
eng.rollout_multi(prompts, embs, G, 200, 1.0, 50, seed=1)
st = eng._last_decode_state
T, Hq, Hkv, hd = 4216, 28, 4, 128
qkv = rnd(T, (Hq + 2 * Hkv) * hd, seed=1)
q, k, v = qkv[:, :Hq * hd], qkv[:, Hq * hd:(Hq + Hkv) * hd], qkv[:, (Hq + Hkv) * hd:]
lo, hi, shared = make_mask("causal", T)
out = torch.empty(T, Hq * hd, device=dev, dtype=torch.bfloat16)
a = rnd(T, 3584, seed=2); w = rnd(4608, 3584, seed=3, scale=0.02); o2 = torch.empty(T, 4608, device=dev, dtype=torch.bfloat16)
nw = rnd(3584, seed=4)
side = torch.cuda.Stream()
torch.cuda.synchronize()
st["step"].fill_(100)
with torch.cuda.stream(side):
    for _ in range(200):
        if what == "attn":
            ops.attn_fwd_tc(q, k, v, lo, hi, shared, Hq, Hkv, hd, hd ** -0.5, out=out, need_lse=False)
        elif what == "gemm":
            ops.gemm(a, w, out=o2)
        else:
            ops.rmsnorm(a, nw, 1e-6)
for _ in range(60):
    st["graph"].replay()
torch.cuda.synchronize()
print(json.dumps({"side": what, "ok": True}))
