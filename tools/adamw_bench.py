"""AdamW kernel alone at the 7B shard size (7.67 B parameters): ms and HBM GB/s (16 B read + 14 B written per parameter)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from time_r1_b200 import ops
n = 7_667_998_720
dev = "cuda"
master = torch.zeros(n, device=dev); m = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev)
g = torch.zeros(n, device=dev, dtype=torch.bfloat16); p = torch.zeros(n, device=dev, dtype=torch.bfloat16)
gsq = torch.ones(1, device=dev)
ts = []
for i in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.adamw(master, m, v, g, p, 1e-6, 0.9, 0.999, 1e-8, 0.0, i + 1, gnorm_sq=gsq, max_norm=1.0); e1.record()
    torch.cuda.synchronize()
    if i >= 2: ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
print(json.dumps({"kernel": "adamw", "ctas_per_sm": os.environ.get("T1_ADAMW_CTAS_PER_SM", "default"), "ms": round(ms, 2), "gbs": round(n * 30 / ms / 1e6, 1)}))
