"""Time the device reward text path (t1_detokenize -> t1_parse_answers -> t1_iou_reward) on cfg2-shaped input:
2 prompt-groups x G=8 completions of C tokens over a Qwen-sized byte-level vocabulary, against the host route it
replaces (D2H of the ids + table walk + utf-8 decode + the reference's regexes + H2D of the rewards)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from time_r1_b200 import ops, rewards  # noqa: E402
from time_r1_b200.detok import DetokTable  # noqa: E402


def main():
    rng = np.random.default_rng(0)
    V = 151665
    frag = [b"<think>", b"</think>", b"<answer>", b"</answer>", b" to ", b" and ", b"\n", b" the", b" event", b".", b" "] + [str(d).encode() for d in range(10)]
    entries = [frag[i] if i < len(frag) else bytes(rng.integers(97, 123, int(rng.integers(1, 12))).astype(np.uint8)) for i in range(V)]
    special = list(range(151643, V))
    table = DetokTable.from_byte_strings(entries, special, "cuda")
    host_table = DetokTable.from_byte_strings(entries, special)
    out = {}
    for C in (200, 1024):
        n = 16
        ids = rng.integers(len(frag), 151643, (n, C)).astype(np.int32)
        for r in range(n):      # <think> + C - 21 random word tokens + </think><answer>12.5 to 17.25</answer><eos><pad...>
            tail = [1, 2, 12, 13, 9, 16, 4, 12, 18, 9, 13, 16, 3, 151645] + [151643] * 6
            ids[r, 0] = 0
            ids[r, C - len(tail):] = tail
        ids_d = torch.from_numpy(ids).cuda()
        gs, ge, du = (torch.full((n,), v, dtype=torch.float64, device="cuda") for v in (10.0, 20.0, 30.0))

        def device_route():
            text, tl = ops.detokenize(ids_d, table)
            s, e, v, f, f32 = ops.parse_answers(text, tl, 0)
            _, o32 = ops.iou_reward(1, s, e, v, gs, ge, du)
            return o32, f32

        o32, f32 = device_route()
        torch.cuda.synchronize()
        texts = [host_table.decode_bytes_host(r).decode("utf-8", "replace") for r in ids]
        want = rewards.iou_timestamp_reward_v2(texts, [(10.0, 20.0)] * n, durations=[30.0] * n)
        assert np.array_equal(o32.cpu().numpy(), np.asarray(want, np.float32)) and f32.min().item() == 1.0, "device route != host route"
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(5):
            device_route()
        e0.record()
        for _ in range(50):
            device_route()
        e1.record()
        torch.cuda.synchronize()
        dev_us = e0.elapsed_time(e1) * 1000 / 50
        t0 = time.perf_counter()
        for _ in range(5):
            h = ids_d.cpu().numpy()
            texts = [host_table.decode_bytes_host(r).decode("utf-8", "replace") for r in h]
            r1 = rewards.iou_timestamp_reward_v2(texts, [(10.0, 20.0)] * n, durations=[30.0] * n)
            r2 = rewards.format_reward(texts)
            torch.tensor(np.stack([r1, r2], 1), dtype=torch.float32).cuda()
            torch.cuda.synchronize()
        host_us = (time.perf_counter() - t0) * 1e6 / 5
        out[f"C={C}"] = {"rows": n, "device_route_us": round(dev_us, 1), "host_route_us": round(host_us, 1),
                         "text_bytes_per_row": int(np.mean([len(t.encode()) for t in texts]))}
    print(json.dumps({"scan_bench": out}))


if __name__ == "__main__":
    main()
