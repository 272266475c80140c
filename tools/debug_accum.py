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
prompts = []
for grid in [(2, 8, 12), (2, 8, 8)]:
    N = grid[0] * grid[1] * grid[2]
    ids = np.array([5, 6, cfg.vision_start_token_id] + [cfg.video_token_id] * (N // 4) + [9] + list(rng.integers(10, 2000, 5 + len(prompts) * 3)) + [7])
    pix = torch.from_numpy(rng.standard_normal((N, cfg.patch_dim)).astype(np.float32)).cuda()
    prompts.append(engine.Prompt(input_ids=ids, pixel_values=pix, grid_thw=[grid], second_per_grid_ts=[1.0]))
embs = [eng.vit_forward(p)[0] for p in prompts]
joint = eng.rollout_multi(prompts, embs, G, C, temperature=0.0, top_k=0, seed=1)
rw = [rng.random((G, 2)).astype(np.float32) for _ in prompts]
def grads():
    return {k: v.float().clone() for k, v in eng.grads.views.items()}
eng.grpo_multi_step(prompts, [lambda c, r=r: r for r in rw], completion_overrides=joint, loss_scale=0.5)
gj = grads()
eng.grad_fresh = True
eng.grpo_step(prompts[0], lambda c: rw[0], completion_override=joint[0], loss_scale=0.5); g0 = grads()
eng.grad_fresh = True
eng.grpo_step(prompts[1], lambda c: rw[1], completion_override=joint[1], loss_scale=0.5); g1 = grads()
eng.grad_fresh = True
eng.grpo_step(prompts[1], lambda c: rw[1], completion_override=joint[1], loss_scale=0.5); g1b = grads()
for k in gj:
    ref = g0[k] + g1[k]
    e = (gj[k] - ref).abs().max().item() / (ref.abs().max().item() + 1e-12)
    rep = (g1[k] - g1b[k]).abs().max().item() / (g1[k].abs().max().item() + 1e-12)
    if e > 5e-3 or rep > 1e-3:
        print(f"{k:50s} joint-vs-sum {e:.4f}  repeatability {rep:.4f}  |ref|max {ref.abs().max().item():.3e}")
print("done")
print("---- isolate: gemm bf16 accumulate")
from time_r1_b200 import ops
a = torch.randn(300, 256, device="cuda").to(torch.bfloat16); b = torch.randn(512, 256, device="cuda").to(torch.bfloat16)
o1 = ops.gemm(a, b); o2 = o1.clone(); ops.gemm(a, b, out=o2, accumulate=True)
print("gemm acc err", ((o2.float() - 2 * o1.float()).abs().max() / o1.float().abs().max()).item())
src = torch.randn(1000, device="cuda"); dst = torch.ones(1000, device="cuda", dtype=torch.bfloat16)
ops.f32_to_bf16(src, dst, accumulate=True); print("f32_to_bf16 acc err", (dst.float() - (src + 1)).abs().max().item())
print("---- same group twice")
eng.grad_fresh = True
eng.grpo_step(prompts[0], lambda c: rw[0], completion_override=joint[0], loss_scale=0.5); ga = grads()
eng.grpo_step(prompts[0], lambda c: rw[0], completion_override=joint[0], loss_scale=0.5); gb = grads()
for k in ga:
    e = (gb[k] - 2 * ga[k]).abs().max().item() / (2 * ga[k].abs().max().item() + 1e-12)
    if e > 5e-3: print(f"{k:50s} twice-vs-2x {e:.4f}")
print("---- g0 then g1 accumulate")
eng.grad_fresh = True
eng.grpo_step(prompts[0], lambda c: rw[0], completion_override=joint[0], loss_scale=0.5)
eng.grpo_step(prompts[1], lambda c: rw[1], completion_override=joint[1], loss_scale=0.5); g01 = grads()
for k in g01:
    ref = g0[k] + g1[k]
    e = (g01[k] - ref).abs().max().item() / (ref.abs().max().item() + 1e-12)
    e2 = (g01[k] - gj[k]).abs().max().item() / (ref.abs().max().item() + 1e-12)
    if e > 5e-3: print(f"{k:50s} seq-accum-vs-sum {e:.4f}   seq-accum-vs-joint {e2:.4f}")
