import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from time_r1_b200 import config, engine, weights
cfg = config.toy()
mw = weights.ModelWeights(cfg, "cuda"); mw.init_random(seed=3, std=0.05)
G, C = 4, 9
run = config.RunConfig(num_generations=G, max_completion_length=C, beta=0.04)
eng = engine.Engine(cfg, mw, run)
rng = np.random.default_rng(0)
grid = (2, 8, 12); N = 192
ids = np.array([5, 6, cfg.vision_start_token_id] + [cfg.video_token_id] * (N // 4) + [9] + list(rng.integers(10, 2000, 5)) + [7])
pix = torch.from_numpy(rng.standard_normal((N, cfg.patch_dim)).astype(np.float32)).cuda()
p = engine.Prompt(input_ids=ids, pixel_values=pix, grid_thw=[grid], second_per_grid_ts=[1.0])
comp = torch.from_numpy(rng.integers(10, 2000, (G, C))).int().cuda()
rw = rng.random((G, 2)).astype(np.float32)
def cs():
    return (mw.train.buf.float().sum().item(), mw.frozen.buf.float().sum().item(), eng.ref.buf.float().sum().item())
print("weights before", cs())
o1 = eng.grpo_step(p, lambda c: rw, completion_override=comp, loss_scale=0.5)
g1 = {k: v.float().clone() for k, v in eng.grads.views.items()}
print("weights after 1", cs())
o2 = eng.grpo_step(p, lambda c: rw, completion_override=comp, loss_scale=0.5)
g2 = {k: v.float().clone() for k, v in eng.grads.views.items()}
print("weights after 2", cs())
for k in ("logp", "ref_logp", "entropy", "dlogp", "advantages"):
    print(k, (o1[k] - o2[k]).abs().max().item(), o1[k].abs().max().item())
print("loss", o1["loss"].item(), o2["loss"].item())
for k in ("model.norm.weight", "lm_head.weight", "model.layers.1.mlp.down_proj.weight"):
    d = g2[k] - 2 * g1[k]
    i = d.abs().argmax().item()
    print(k, "max dev", d.abs().max().item(), "at", i, "g1", g1[k].flatten()[i].item(), "g2", g2[k].flatten()[i].item(), "|g1|max", g1[k].abs().max().item())
