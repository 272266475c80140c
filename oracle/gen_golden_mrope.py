"""ORACLE tooling — M-RoPE position goldens from vLLM's `Qwen2_5_VLForConditionalGeneration.get_mrope_input_positions`
(vllm 0.22.0 in this container, model_executor/models/qwen2_5_vl.py:1243-1282), which keeps the transformers-4.51.1
index semantics the reference was written against.  The method is called unbound on a stub `self`.

    python oracle/gen_golden_mrope.py      ->  tests/golden/mrope_vllm.json
"""
import json
import os
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    from vllm.model_executor.models.qwen2_5_vl import Qwen2_5_VLForConditionalGeneration as M

    VID = 151656
    cases = []
    rng = np.random.default_rng(0)
    specs = [
        ([(2, 8, 12)], [1.0], 2, [3], 7),
        ([(16, 32, 32)], [1.0], 2, [4], 116),
        ([(4, 6, 10)], [0.5], 2, [5], 9),
        ([(8, 26, 46)], [1.0009], 2, [4], 30),
        ([(3, 4, 4), (2, 6, 8)], [2.0, 0.25], 2, [2, 6], 5),
        ([(147, 6, 12)], [1.0], 4, [4], 12),
    ]
    for grids, spgs, tps, text_before, tail in specs:
        ids, feats = [], []
        for g, spg, nb in zip(grids, spgs, text_before):
            ids += list(rng.integers(0, 1000, nb))
            n = g[0] * (g[1] // 2) * (g[2] // 2)
            feats.append(types.SimpleNamespace(
                modality="video", mm_position=types.SimpleNamespace(offset=len(ids)),
                data={"video_grid_thw": types.SimpleNamespace(data=torch.tensor(g)),
                      "second_per_grid_ts": types.SimpleNamespace(data=torch.tensor(spg))}))
            ids += [VID] * n
        ids += list(rng.integers(0, 1000, tail))
        stub = types.SimpleNamespace(config=types.SimpleNamespace(vision_config=types.SimpleNamespace(spatial_merge_size=2, tokens_per_second=tps)))
        stub.iter_mm_grid_thw = types.MethodType(M.iter_mm_grid_thw, stub)
        pos, delta = M.get_mrope_input_positions(stub, [int(t) for t in ids], feats)
        cases.append(dict(grids=grids, second_per_grid_ts=spgs, tokens_per_second=tps, ids=[int(t) for t in ids],
                          pos=pos.numpy().astype(int).tolist(), delta=int(delta)))
    with open(os.path.join(ROOT, "tests", "golden", "mrope_vllm.json"), "w") as f:
        json.dump(cases, f)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
