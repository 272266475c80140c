"""ORACLE tooling — generates tests/golden/hf_video_preproc.npz with the HF Qwen2-VL video processor installed in the
build container (transformers 5.5.0, `Qwen2VLVideoProcessor`, class defaults = the released preprocessor_config:
CLIP mean/std, rescale 1/255, patch 14, temporal patch 2, merge 2), fed with resized integer-valued frames exactly as the
reference hands them over (src/utils/vision_process.py:440-475 -> src/time_r1/rl/timer1_trainer.py:547-556,
do_resize=False).  Inputs are regenerated from the seed in the tests; only the outputs are stored.
    python oracle/gen_golden_preproc.py
"""
import os

import numpy as np
import torch
from transformers.models.qwen2_vl.video_processing_qwen2_vl import Qwen2VLVideoProcessor

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "hf_video_preproc.npz")
CASES = [(4, 56, 84, 0), (3, 28, 56, 1), (2, 84, 28, 2)]      # (frames, H, W, seed); 3 frames: odd count is padded


def frames_for(T, H, W, seed):
    return np.random.default_rng(seed).integers(0, 256, size=(T, 3, H, W), dtype=np.uint8)


def main():
    vp = Qwen2VLVideoProcessor()
    out = {}
    for T, H, W, seed in CASES:
        v = torch.from_numpy(frames_for(T, H, W, seed)).float()
        r = vp(videos=[v], do_resize=False, return_tensors="pt", do_sample_frames=False)
        out[f"pix_{T}_{H}_{W}_{seed}"] = r["pixel_values_videos"].numpy()
        out[f"grid_{T}_{H}_{W}_{seed}"] = r["video_grid_thw"].numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items()}, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
