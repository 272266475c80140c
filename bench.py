#!/usr/bin/env python
"""bench.py — GRPO rollout-samples/s of the hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the reference's CPU path (oracle port) on the host cores

A step = ONE complete GRPO update for one prompt-group per GPU: ViT, prompt prefill, C-token sampled
rollout of G completions, policy + reference scoring, rewards (device detokenise + parse + IoU), group advantage, clipped
loss, full backward and the sharded AdamW step.  Workload = cfg2: Qwen2.5-VL-7B dims (the class the reference
instantiates, timer1_trainer.py:244), 16 temporal slots of 448x448 -> grid (16,32,32), 4096 vision tokens, prompt
~4.2k tokens, G=8, C=200, synthetic uniform uint8 frames, random-init N(0,0.02) bf16 weights, synthetic template rewards.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant GEMM from the committed ncu --set full capture
# (profiles/r02_ncu_full_tensor_kernels.txt): the gate_up projection over the packed scoring rows in the form the step runs it
# (2-CTA kernel, SwiGLU epilogue, projection kept for the backward).  Algorithmic bytes of that launch: A 42.2 MB + B 271.6 MB +
# gate_up 446.2 MB + act 223.1 MB = 983 MB.
NCU_GEMM_TRAFFIC = {"bytes": 1399.2e6, "note": "ncu, one launch of gemm2_bf16_tn (cta_group::2) M=5888 N=2x18944 K=3584 with the SwiGLU epilogue: "
                                              "1.003 ms, tensor pipe 97.6% active, DRAM read 753.9 MB + write 645.3 MB vs 983 MB algorithmic (the 271.6 MB "
                                              "weight operand is read ~2.6x: it does not fit the L2 and each die caches its own copy); the plain form "
                                              "0.964 ms / 98.3%, dgrad 1.125 ms / 99.0% (2.6 GB read: dY re-read 3.6x), wgrad 1.019 ms / 98.2%: none is HBM-bound"}
METRIC = "grpo_rollout_samples_per_sec"
UNIT = "samples/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "hf_gpu"],
                    help="ours: this engine; reference: the reference's CPU path (oracle port) on the host cores; hf_gpu: the "
                         "reference's own GRPO step through the HF model classes on ONE GPU (oracle/hf_reference.py)")
    ap.add_argument("--no-gpu-reference", action="store_true",
                    help="skip the `gpu_reference` leg (HF classes, reference-style arithmetic, 1 warm-up + 2 steps in a subprocess)")
    ap.add_argument("--model", default="qwen2.5-vl-7b")
    ap.add_argument("--slots", type=int, default=16, help="temporal slots of 32x32 patches (16 -> cfg2, 32 -> cfg3, 64 -> cfg5)")
    ap.add_argument("--grid", default=None, help="t,h,w patch grid of the video instead of (slots,32,32), e.g. 147,6,12 = a 147 s "
                    "TimeRFT clip at 84x168 (cfg4, SURVEY 8d); frames are 2t x 3 x 14h x 14w; 'timerft' = a different TimeRFT-shaped "
                    "clip per prompt-group (cfg4: (147,6,12), (8,26,46), (60,16,28), (384,6,10) dealt over ranks and micro-batches)")
    ap.add_argument("--filter-pass", type=int, default=0, metavar="N",
                    help="cfg4: after the timed steps run ONE per-epoch sample-filtering pass (train_rl_SF.sh:71-111) over N synthetic "
                         "samples per rank — greedy rollout, difficulty = 100 x IoU, 0 < IoU <= 0.7 selection — and report its throughput")
    ap.add_argument("--group-size", type=int, default=8)
    ap.add_argument("--completion-len", type=int, default=200)
    ap.add_argument("--groups-per-step", type=int, default=2,
                    help="prompt-groups per GPU per optimizer step = gradient_accumulation_steps (2 in scripts/posttrain/train_rl.sh:27)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--reward-route", default="device", choices=["device", "host"],
                    help="device: template completions are detokenised, parsed and scored on the GPU; host: regex parse on the host")
    ap.add_argument("--no-vit-prefetch", dest="vit_prefetch", action="store_false",
                    help="do not overlap the next step's frozen ViT blocks with this step's optimizer phase")
    ap.add_argument("--train-vit", action="store_true",
                    help="fix_vit=False (the reference's default, main.py:46-49; every shipped script passes --fix_vit true): the ViT "
                         "patch-embed and blocks are trained too -- ViT-block backward, a separate reference-model tower, no ViT prefetch")
    ap.add_argument("--verify", default=None, metavar="OUT.pt",
                    help="numerical invariance check of the multi-GPU path instead of a benchmark: ONE optimizer step over a fixed "
                         "global batch of 8 teacher-forced prompt-groups dealt over the ranks; writes loss / grad-norm / strided "
                         "samples of the averaged gradient and of the post-step weights to OUT.pt (compare runs at different N with "
                         "tools/verify_compare.py)")
    ap.add_argument("--phase-times", action="store_true", help="print a per-phase breakdown to stderr")
    ap.add_argument("--kernel-times", action="store_true", help="print per-entry-point CUDA-event totals of one extra step to stderr")
    ap.add_argument("--ncu-range", action="store_true",
                    help="after the warm-up run ONE step inside cudaProfilerStart/Stop with only the first decode-graph replay "
                         "profiled (ncu --profile-from-start off), then exit: the launch list of a step = its eager launches + "
                         "one decode step (x C-1 replays)")
    return ap.parse_args()


# cfg4: TimeRFT clips at 360x640 through the reference's smart_resize — an 8.9 s, a 30 s, a 147 s (dataset mean)
# and a >= 384 s video
TIMERFT_GRIDS = [(147, 6, 12), (8, 26, 46), (60, 16, 28), (384, 6, 10)]


def grid_of(args, index: int = 0, world: int = 1):
    """Patch grid of prompt-group `index` = rank * groups_per_step + micro-batch.  cfg4: the world x groups_per_step clips of one
    optimizer step (the four TimeRFT shapes in rotation) are dealt to the ranks the way the trainer's dataloader deals a
    global batch — dist_utils.balanced_deal, even token counts per rank — because the step ends with the slowest rank."""
    if args.grid == "timerft":
        from time_r1_b200.dist_utils import balanced_deal

        ng = args.groups_per_step
        clips = [TIMERFT_GRIDS[i % len(TIMERFT_GRIDS)] for i in range(world * ng)]
        order = balanced_deal([t * h * w for t, h, w in clips], world) if not os.environ.get("T1_NO_BALANCE") else list(range(world * ng))
        rank, a = divmod(index, ng)
        return clips[order[a * world + rank]] if rank < world else clips[index % len(clips)]
    return tuple(int(v) for v in args.grid.split(",")) if args.grid else (args.slots, 32, 32)


def workload_name(args):
    if args.grid == "timerft":
        gs = ", ".join(f"({t},{h},{w})" for t, h, w in TIMERFT_GRIDS)
        return (f"cfg4: {args.model} dims, TimeRFT-shaped clips {gs} dealt over prompt-groups, G={args.group_size}, C={args.completion_len}, "
                f"{args.groups_per_step} prompt-group(s) per GPU per optimizer step")
    t, h, w = grid_of(args)
    return (f"cfg2-family: {args.model} dims, grid ({t},{h},{w}) = {t * h * w // 4} vision tokens, "
            f"G={args.group_size}, C={args.completion_len}, {args.groups_per_step} prompt-group(s) per GPU per optimizer step "
            f"(gradient_accumulation_steps of train_rl.sh:27)")


# ------------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi, during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.thread = index, [], None, None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline (oracle port on the host cores)
# ------------------------------------------------------------------------------------------------
def cpu_reference(args, steps: int, warmup: int) -> dict:
    """The reference's arithmetic on the host cores (oracle port, fp32, all cores): each step is a BOUNDED SAMPLE of the
    workload; `measured` is what the clock saw, `value` scales it to the metric's unit by reference-style FLOPs and is
    labelled an estimate."""
    import torch

    from oracle import cpu_baseline as B, qwen_vl_oracle as O
    from time_r1_b200 import config, flops

    full = config.PRESETS[args.model]()
    ocfg = O.default_cfg(**{k: v for k, v in full.to_dict().items() if k in O.default_cfg()})
    scfg = B.sample_cfg(ocfg)
    G_s, C_s, grid_s = 2, 4, [(2, 16, 16)]
    st = B.ReferenceStyleStep(scfg, grid_s, G=G_s, C=C_s)
    for _ in range(warmup):
        st.step()
    ts = [st.step() for _ in range(steps)]
    t_sample = sum(ts) / len(ts)
    # scale the bounded sample to the full workload by reference-style FLOPs
    import dataclasses
    s_model = dataclasses.replace(full, num_hidden_layers=1, vit_depth=1, fullatt_block_indexes=(0,), vocab_size=32768)
    f_sample = flops.reference_style_flops(s_model, grid_s, st.P, G_s, C_s)
    grid = [grid_of(args)]
    P_full = 4 + grid[0][0] * grid[0][1] * grid[0][2] // 4 + 1 + 110 + 5
    f_full = flops.reference_style_flops(full, grid, P_full, args.group_size, args.completion_len)
    t_full = t_sample * f_full / f_sample
    value = args.group_size / t_full
    return dict(value=value, unit=UNIT, value_kind="estimated: measured bounded sample scaled to the full workload by reference-style FLOPs",
                cores=st.threads, kind="port", sample=B.SAMPLE_DESC,
                measured={"sample_s_per_step": round(t_sample, 3), "sample_samples_per_s": round(G_s / t_sample, 4), "steps": len(ts),
                          "sample_tflop": round(f_sample / 1e12, 4), "cpu_tflops": round(f_sample / t_sample / 1e12, 3)},
                estimated={"full_tflop": round(f_full / 1e12, 1), "flop_scale": round(f_full / f_sample, 1), "full_step_s": round(t_full, 1)})


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_reference(args, max(1, args.steps), max(1, min(args.warmup, 2)))
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "value_kind": cb["value_kind"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        # what the clock measured: one bounded-sample step (NOT a full cfg2 step; the full-step estimate is in cpu_baseline.estimated)
        "ms_per_step": round(cb["measured"]["sample_s_per_step"] * 1e3, 1), "ms_per_step_is": "measured bounded-sample step",
        "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": workload_name(args), "timed": cb["sample"]},
        "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU reference: the reference's own step through the HF model classes on one B200 (the denominator of the >= 10x target)
# ------------------------------------------------------------------------------------------------
def run_hf_gpu(args):
    """HF Qwen2_5_VLForConditionalGeneration, random init, bf16, FA2 else SDPA, reference-style arithmetic exactly as
    timer1_trainer.py:568-737 (generate on G expanded copies, two full-length scoring passes with logits at every position)
    + backward through gradient-checkpointed layers + clip + AdamW, on the bench's own workload.  One prompt-group per
    micro-step, `groups_per_step` micro-steps per optimizer step, like train_rl.sh:26-27."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from oracle import hf_reference as H
    from time_r1_b200 import config

    if not torch.cuda.is_available():
        print(json.dumps({"impl": "hf_gpu", "unavailable": "no CUDA device"}), flush=True)
        return
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    cfg = config.PRESETS[args.model]()
    G, C, NG = args.group_size, args.completion_len, args.groups_per_step
    t_init = time.perf_counter()
    ref = H.ReferenceStyleGRPO(cfg, dev, torch.bfloat16, G=G, C=C, beta=0.04, temperature=1.0, top_k=50, grad_ckpt=True)
    inputs = [H.synthetic_prompt_inputs(cfg, grid_of(args), dev, torch.bfloat16, seed=42 + i) for i in range(NG)]
    torch.cuda.synchronize()
    t_init = time.perf_counter() - t_init
    gen = torch.Generator(device=dev).manual_seed(7)
    reward_fn = lambda comp: torch.rand((comp.shape[0], 2), device=dev, generator=gen)      # synthetic rewards (random-init model)

    def one_step():
        phases = {}
        for i in range(NG):
            out = ref.step(inputs[i], reward_fn, grad_accum=NG, do_optimizer=(i == NG - 1))
            for k, v in out["times"].items():
                phases[k] = phases.get(k, 0.0) + v
        return phases, float(out["loss"].item()), float(out["grad_norm"].item())

    for _ in range(args.warmup):
        one_step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    phases = {}
    for _ in range(args.steps):
        ph, loss, gn = one_step()
        for k, v in ph.items():
            phases[k] = phases.get(k, 0.0) + v
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    value = NG * G / (ms / 1e3)
    line = {
        "impl": "hf_gpu", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_name(args), "prompt_tokens": int(inputs[0]["input_ids"].shape[1])},
        "how": "HF %s (transformers %s) random-init bf16, attn_implementation=%s, reference-style arithmetic (oracle/hf_reference.py restates "
               "timer1_trainer.py:568-737; pinned to the unmodified compute_loss by tests/golden/compute_loss_hf_toy.json), gradient "
               "checkpointing, clip 1.0 + fused AdamW in HBM (no ZeRO-3 / CPU offload), generate in eval mode with the KV cache"
               % (type(ref.model).__name__, __import__("transformers").__version__, ref.attn_impl),
        "phase_s_per_step": {k: round(v / args.steps, 3) for k, v in phases.items()}, "init_s": round(t_init, 1),
        "loss": loss, "grad_norm": gn,
        "hbm_gb": {"max_allocated": round(torch.cuda.max_memory_allocated(dev) / 1e9, 1)},
    }
    print(json.dumps(line), flush=True)


def gpu_reference_leg(args) -> dict:
    """Runs `bench.py --impl hf_gpu` in a subprocess BEFORE the engine claims the HBM; returns its JSON line (or why not)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "hf_gpu", "--steps", "2", "--warmup", "1", "--model", args.model,
           "--slots", str(args.slots), "--group-size", str(args.group_size), "--completion-len", str(args.completion_len),
           "--groups-per-step", str(args.groups_per_step)] + (["--grid", args.grid] if args.grid else [])
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    except subprocess.TimeoutExpired:
        return {"unavailable": "hf_gpu leg exceeded 900 s"}
    for ln in reversed(r.stdout.strip().splitlines()):
        if ln.startswith("{"):
            try:
                return json.loads(ln)
            except Exception:
                pass
    return {"unavailable": "hf_gpu leg failed: " + (r.stderr.strip().splitlines() or ["no output"])[-1][:300]}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def synthetic_prompt(cfg, grid, rank: int, step: int, device):
    """SURVEY §8d synthetic inputs (seed 42 mirrors main.py:552)."""
    import numpy as np
    import torch

    rng = np.random.default_rng(42 + 1000 * rank + step)
    t, h, w = grid
    N = t * h * w
    nv = N // cfg.merge_unit
    text = rng.integers(0, 151643, 110)
    ids = np.concatenate([[151644, 872, 198, 151652], np.full(nv, cfg.video_token_id), [151653], text, [151645, 198, 151644, 77091, 198]]).astype(np.int64)
    g = torch.Generator().manual_seed(42 + 1000 * rank + step)
    # resized integer-valued frames as fetch_video_v3 hands them to the processor (vision_process.py:440-475):
    # 2 * slots frames of 448 x 448; the processor stage (rescale / normalise / patchify) runs on the device
    frames_host = torch.randint(0, 256, (2 * t, 3, 14 * h, 14 * w), generator=g, dtype=torch.uint8).pin_memory()
    return ids, frames_host, [(t, h, w)]


class TemplateRewards:
    """Completions are overwritten post hoc with template strings (random-init models never emit <answer> tags —
    SURVEY §0); the strings go through the product's reward path.  route "device" (default): the templates enter as token
    ids of a byte-level vocabulary and are detokenised, parsed and scored on the GPU (t1_detokenize -> t1_parse_answers ->
    t1_iou_reward; nothing is copied back before the loss) — the vocabulary is Qwen-sized (152064 entries, multi-byte tokens,
    specials at the Qwen ids), not a 258-entry byte table; route "host": regex parse on the host + t1_iou_reward."""

    def __init__(self, G: int, rank: int, route: str = "device"):
        import numpy as np

        self.rows = json.load(open(os.path.join(ROOT, "tests", "golden", "iou_train_2k5.json")))
        self.rng = np.random.default_rng(7 + rank)
        self.G, self.cursor = G, rank
        if route == "device":
            self.device = self._device_route      # the attribute Engine.grpo_multi_step looks for
            self._dr = None

    def _templates(self):
        row = self.rows[self.cursor % len(self.rows)]
        self.cursor += 1
        dur = row["duration"]
        texts = []
        for g in range(self.G):
            s, e = sorted(self.rng.uniform(0, dur, 2))
            if g % 4 == 3:
                texts.append(f"<think>unsure</think> {s:.2f} to {e:.2f}")      # malformed: no <answer> block
            else:
                texts.append(f"<think>the event spans these frames</think><answer>{s:.2f} to {e:.2f}</answer>")
        return row, dur, texts

    # ---- a Qwen-sized byte-level vocabulary (152064 entries, multi-byte tokens) for the device route -------------------
    # ids 0..255 = the single bytes; then every 2-character and the common 3/4-character strings of the completion alphabet
    # (what BPE merges look like: " to", "think", "</", "12", ".5" ...); specials at the Qwen ids (<|endoftext|> 151643,
    # <|im_end|> 151645); the rest of the table is filler words so its size and byte-offset range match the production table.
    VOCAB, EOT, IM_END = 152064, 151643, 151645

    @classmethod
    def _vocabulary(cls):
        if getattr(cls, "_vocab", None) is None:
            import itertools

            alphabet = "0123456789. <>/abcdefghijklmnopqrstuvwxyz"
            entries = [bytes([b]) for b in range(256)]
            multi = ["".join(p) for p in itertools.product(alphabet, repeat=2)]
            multi += ["<think>", "</think>", "<answer>", "</answer>", " to ", " to", "the ", "event", "spans", "these", "frames", "unsure",
                      "think", "answer", " the", "ans", "wer", "fra", "mes", "spa", "eve"]
            multi += [f"{a}{b}{c}" for a in "0123456789" for b in "0123456789." for c in "0123456789"]
            entries += [m.encode() for m in dict.fromkeys(multi)]
            n_real = len(entries)
            entries += [b"w%05d" % i for i in range(cls.VOCAB - n_real)]
            entries[cls.EOT], entries[cls.IM_END] = b"<|endoftext|>", b"<|im_end|>"
            lookup = {}
            for i, e in enumerate(entries[:n_real]):
                lookup.setdefault(e, i)
            cls._vocab = (entries, lookup, max(len(e) for e in entries[:n_real]))
        return cls._vocab

    @classmethod
    def _tokenize(cls, text: str):
        """Greedy longest-match over the vocabulary above (every byte is a token, so it always terminates)."""
        _, lookup, longest = cls._vocabulary()
        b = text.encode("utf-8")
        ids, i = [], 0
        while i < len(b):
            for n in range(min(longest, len(b) - i), 0, -1):
                t = lookup.get(b[i:i + n])
                if t is not None:
                    ids.append(t)
                    i += n
                    break
        return ids

    def _device_route(self, completion_ids):
        import numpy as np
        import torch

        from time_r1_b200 import rewards
        from time_r1_b200.detok import DetokTable

        if self._dr is None:
            entries, _, _ = self._vocabulary()
            table = DetokTable.from_byte_strings(entries, [self.EOT, self.IM_END], completion_ids.device)
            self._dr = rewards.DeviceRewards([rewards.iou_timestamp_reward_v2, rewards.format_reward], table)
        row, dur, texts = self._templates()
        toks = [self._tokenize(t) for t in texts]
        C = completion_ids.shape[1]
        ids = np.full((self.G, max(C, max(len(t) for t in toks) + 1)), self.EOT, np.int32)
        for g, t in enumerate(toks):
            ids[g, :len(t)] = t
            ids[g, len(t)] = self.IM_END                                       # <|im_end|>, then <|endoftext|> padding
        return self._dr(torch.from_numpy(ids).to(completion_ids.device, non_blocking=True), row["gt"], dur)

    def __call__(self, completion_ids):
        import numpy as np

        from time_r1_b200 import rewards

        row, dur, texts = self._templates()
        r1 = rewards.iou_timestamp_reward_v2(texts, [row["gt"]] * self.G, durations=[dur] * self.G)
        r2 = rewards.format_reward(texts)
        return np.stack([np.asarray(r1, np.float32), np.asarray(r2, np.float32)], 1)


def attention_microbench(eng, Ps, G: int, C: int):
    """The decode-attention launch of the captured decode step, alone, at mid-rollout (step C/2) for all prompt-groups
    of the step, L2 flushed between launches: unique K/V bytes / time (shared-prefix byte count)."""
    import torch

    from time_r1_b200.lib import call, ptr, stream_ptr

    cfg = eng.cfg
    st = eng._decode_state(tuple(Ps), G, C)
    step = C // 2
    st["step"].fill_(step)
    n = len(Ps)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.dev)
    Hq, Hkv, hd = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
    common = (ptr(st["qkv"]), st["qkv"].stride(0), ptr(st["group_table"][0]), st["kv_maps"][0].ctypes.data, n, G, Hq, Hkv, hd,
              st["cpad"], ptr(st["step"]), ptr(st["attn"]), st["attn"].stride(0), hd ** -0.5)
    wide = "attn_ws" in st

    def launch():
        if wide:
            call("t1_decode_attn_wide", *common, st["attn_splits"], ptr(st["attn_ws"]), st["attn_ws"].numel(), stream_ptr())
        else:
            call("t1_decode_attn_cluster", *common, 0, stream_ptr())

    ts = []
    for i in range(13):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launch()
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    # the same launch as it runs inside the decode step: the attention launches of ALL layers (each on its own K/V cache: 28 x
    # unique bytes >> L2) back to back in one CUDA graph -- launch latency overlapped by the queue, DRAM traffic as in the step
    L = cfg.num_hidden_layers
    side = torch.cuda.Stream(device=eng.dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        def launch_layer(l):
            cl = (ptr(st["qkv"]), st["qkv"].stride(0), ptr(st["group_table"][l]), st["kv_maps"][l].ctypes.data, n, G, Hq, Hkv, hd,
                  st["cpad"], ptr(st["step"]), ptr(st["attn"]), st["attn"].stride(0), hd ** -0.5)
            if wide:
                call("t1_decode_attn_wide", *cl, st["attn_splits"], ptr(st["attn_ws"]), st["attn_ws"].numel(), stream_ptr())
            else:
                call("t1_decode_attn_cluster", *cl, 0, stream_ptr())
        for l in range(L):
            launch_layer(l)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for l in range(L):
                launch_layer(l)
    torch.cuda.current_stream().wait_stream(side)
    tg = []
    for i in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        tg.append(e0.elapsed_time(e1) / (4 * L))
    tg.sort()
    ms_graph = tg[len(tg) // 2]
    del g
    kv_bytes = sum(2.0 * Hkv * hd * 2 * (P + G * (step + 1)) for P in Ps) + 2.0 * n * G * Hq * hd * 2
    ctas = (st["attn_splits"] if wide else 8) * Hkv * n
    return {"kernel": "decode_attn_tc (1 launch, all groups, %s)" % ("wide form: %d CTAs, L2 workspace merge" % ctas if wide
                                                                      else "cluster form: %d CTAs, DSMEM merge" % ctas),
            "groups": n, "ctas": ctas, "unique_kv_bytes": kv_bytes, "ms": round(ms, 5), "gbs": round(kv_bytes / ms / 1e6, 1),
            "in_graph": {"what": "the %d layers' launches back to back in one CUDA graph (distinct K/V per layer)" % L,
                         "ms": round(ms_graph, 5), "gbs": round(kv_bytes / ms_graph / 1e6, 1)}}


def decode_step_microbench(eng, Ps, G: int, C: int):
    """The captured decode step (one token for every sequence of the step's prompt-groups: 113 weight-streaming GEMMs,
    28 attention launches, norms, sampler) replayed at mid-rollout: HBM bytes that MUST move per step (every LLM
    weight once + the unique K/V of all layers) / time.  CUDA events around 16 graph replays, median of 9."""
    import torch

    cfg = eng.cfg
    st = eng._decode_state(tuple(Ps), G, C)
    if st.get("graph") is None:
        return None
    L, D, I, V = cfg.num_hidden_layers, cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size
    Hq, Hkv, hd = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
    step = C // 2
    w_bytes = 2.0 * (L * (cfg.qkv_dim * D + Hq * hd * D + 2 * I * D + D * I) + V * D)
    kv_bytes = L * sum(2.0 * Hkv * hd * 2 * (P + G * (step + 1)) for P in Ps)
    snap = (st["next_ids"].clone(), st["finished"].clone(), st["out_ids"].clone())
    ts = []
    for _ in range(9):
        st["step"].fill_(step)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(16):
            st["graph"].replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 16)
    st["next_ids"].copy_(snap[0]); st["finished"].copy_(snap[1]); st["out_ids"].copy_(snap[2])
    ts.sort()
    ms = ts[len(ts) // 2]
    paired = eng.run.decode_fold_norm and getattr(eng.run, "decode_pair", False) and st["launches_per_step"] < 4 * L
    gemms = ("stream_gemm_pair_kernel x%d (o_proj + gate_up, down_proj + next qkv: one launch each, RMSNorms folded, arrival counter "
             "between the two GEMMs) + stream_gemm_kernel x3" % (2 * L - 1)) if paired else (
        "stream_gemm_kernel x%d%s" % (4 * L + 1, " with the RMSNorms folded into their epilogues" if eng.run.decode_fold_norm else " + norm kernels"))
    return {"bound": "hbm", "kernel": "captured decode step (%s + decode_attn_tc x%d + sampler)" % (gemms, L),
            "launches": st["launches_per_step"], "weight_bytes": w_bytes, "kv_bytes": kv_bytes, "ms": round(ms, 4),
            "achieved": round((w_bytes + kv_bytes) / ms / 1e6, 1), "unit": "GB/s"}


def filter_pass(args, eng, cfg, rank, world, dev):
    """One per-epoch sample-filtering pass (scripts/posttrain/train_rl_SF.sh:71-111) over --filter-pass synthetic TimeRFT-shaped
    samples per rank: greedy rollout (T = 0, 64 new tokens: random-init models never stop) batched 4 prompts per decode
    batch, difficulty = 100 x IoU on the device (t1_iou_reward mode 2), 0 < IoU <= 0.7 + difficulty-sorted linspace
    selection over the rows of ALL ranks (process_data.py:61-80).  Device-timed, max over ranks."""
    import torch
    import torch.distributed as dist

    from time_r1_b200 import engine, filtering

    rows = json.load(open(os.path.join(ROOT, "tests", "golden", "iou_train_2k5.json")))
    n = args.filter_pass
    samples = []
    for i in range(n):
        gidx = rank * n + i
        ids, frames_host, grid = synthetic_prompt(cfg, TIMERFT_GRIDS[gidx % len(TIMERFT_GRIDS)], 10_000 + gidx, 0, dev)
        row = rows[gidx % len(rows)]
        samples.append(dict(qid=gidx, timestamp=row["gt"], duration=row["duration"],
                            _prompt=engine.Prompt(input_ids=ids, pixel_values=None, grid_thw=grid, second_per_grid_ts=[1.0],
                                                  frames=frames_host.to(dev, non_blocking=True))))
    decode = lambda ids: bytes(int(t) % 256 for t in ids).decode("latin-1")
    flt = filtering.GreedyFilter(eng, lambda s: s["_prompt"], decode, max_new_tokens=64, batch=4)
    flt.difficulties(samples[:4])                  # warm-up: captures the decode graph of this batch shape
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    diff = flt.difficulties(samples)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    merged = diff
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        parts = [None] * world
        dist.all_gather_object(parts, diff)
        merged = {k: v for p_ in parts for k, v in p_.items()}
    annotated = [dict(qid=q, difficulty=d["difficulty"], pred=d["pred"]) for q, d in sorted(merged.items())]
    selected = filtering.select_0070(annotated, k=max(1, len(annotated) // 2))
    return {"samples": n * world, "ms": round(float(ms.item()), 1), "samples_per_s": round(n * world / (float(ms.item()) / 1e3), 2),
            "max_new_tokens": 64, "decode_batch": 4, "annotated": len(annotated), "selected_0070": len(selected)}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from time_r1_b200 import config, engine, flops, lib, ops, weights

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    lib.load()
    gpu_ref = None
    if world == 1 and not args.no_gpu_reference and not args.ncu_range:
        gpu_ref = gpu_reference_leg(args)          # before this process allocates its 175 GB
    cfg = config.PRESETS[args.model]()
    run = config.RunConfig(num_generations=args.group_size, max_completion_length=args.completion_len, fix_vit=not args.train_vit)
    if os.environ.get("T1_SIDE_CTAS"):
        run.side_ctas = int(os.environ["T1_SIDE_CTAS"])
    if os.environ.get("T1_DECODE_FOLD"):
        run.decode_fold_norm = os.environ["T1_DECODE_FOLD"] == "1"
    if os.environ.get("T1_DECODE_CTAS"):
        run.decode_ctas = int(os.environ["T1_DECODE_CTAS"])
    G, C = run.num_generations, run.max_completion_length
    mw = weights.ModelWeights(cfg, dev, world_size=world, train_vit=args.train_vit)
    mw.init_random(seed=0)          # same weights on every rank (data parallel)
    eng = engine.Engine(cfg, mw, run, rank=rank, world_size=world, process_group=pg)
    NG = args.groups_per_step
    rewards = [TemplateRewards(G, rank * NG + i, args.reward_route) for i in range(NG)]
    inputs = [synthetic_prompt(cfg, grid_of(args, rank * NG + i, world), rank * NG + i, 0, dev) for i in range(NG)]
    ids, frames_host, grid = inputs[0]
    P = len(ids)
    # `value` runs with pixel_values_videos resident in HBM (the processor output, fp32 [N, 1176]); `e2e` starts from the
    # uint8 frames in pinned host memory
    pix_devs = [ops.video_patchify(inp[1].to(dev))[0] for inp in inputs]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass

    pending = {}      # ViT prefetch of the NEXT step's prompts, issued before this step's optimizer phase

    def make_prompts(e2e: bool):
        prompts = []
        for i, (ids_i, frames_host_i, grid_i) in enumerate(inputs):
            if e2e:      # H2D of the step's frames from pinned memory, device pre-processing inside the step
                prompts.append(engine.Prompt(input_ids=ids_i, pixel_values=None, grid_thw=grid_i, second_per_grid_ts=[1.0],
                                             frames=frames_host_i.to(dev, non_blocking=True)))
            else:
                prompts.append(engine.Prompt(input_ids=ids_i, pixel_values=pix_devs[i], grid_thw=grid_i, second_per_grid_ts=[1.0]))
        return prompts

    def one_step(step_idx: int, e2e: bool, timers=None):
        # the trainer's dataloader knows the next batch: its frozen ViT blocks run on a side stream under this step's
        # optimizer phase (--fix_vit: they depend on no trained weight); every step still runs its own ViT exactly once
        prompts, handle = pending.pop(e2e, (None, None))
        if prompts is None:
            prompts = make_prompts(e2e)
        outs = eng.grpo_multi_step(prompts, rewards, seed=run.seed + step_idx * world + rank, loss_scale=1.0 / NG, timers=timers,
                                   vit_prefetch=handle)
        if args.vit_prefetch:
            nxt = make_prompts(e2e)                  # e2e: the H2D copy of the next step's frames is issued here
            pending[e2e] = (nxt, eng.prefetch_vit(nxt))
        gn = eng.optimizer_step()
        if timers is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            timers["_events"].append(("optimizer(+vit prefetch issue)", ev))
        if e2e:
            return float(sum(o["loss"].item() for o in outs)), float(gn.item())   # D2H of the step's result
        return outs[0]["loss"], gn

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n: int, e2e: bool, base: int):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            one_step(base + i, e2e)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for i in range(args.warmup):
        one_step(i, False)
    if args.ncu_range:
        rt = torch.cuda.cudart()

        def hook(i, n):
            if i == 1:
                torch.cuda.synchronize()
                rt.cudaProfilerStop()        # replays 1 .. C-2 repeat replay 0: skipped
            elif i == n - 1 and n > 2:
                torch.cuda.synchronize()
                rt.cudaProfilerStart()

        eng.replay_hook = hook
        torch.cuda.synchronize()
        rt.cudaProfilerStart()
        one_step(args.warmup, False)
        torch.cuda.synchronize()
        rt.cudaProfilerStop()
        print(json.dumps({"ncu_range": "one step: eager launches + decode replays 0 and C-2", "decode_replays_per_step": C - 1,
                          "launches_per_decode_step": eng._last_decode_state["launches_per_step"]}), flush=True)
        return
    # the diagnostic steps run on EVERY rank (a step contains the optimizer's reduce-scatter / all-gather: a step taken by
    # rank 0 alone would leave the other ranks' collectives unmatched and hang the job); only rank 0 prints
    if args.phase_times:
        tm = {}
        one_step(args.warmup, False, timers=tm)
        phases = engine.finish_timers(tm)
        if rank == 0:
            print("[phase ms]", json.dumps({k: round(v, 1) for k, v in phases.items()}), file=sys.stderr)
    if args.kernel_times:
        lib.CALL_PROFILER.start()
        one_step(args.warmup + 1, False)
        lib.CALL_PROFILER.stop()
        summ = lib.CALL_PROFILER.summary()
        if rank == 0:
            print("[kernel ms, eager part of one step]", json.dumps({k: [v[0], round(v[1], 2)] for k, v in summ.items()}), file=sys.stderr)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.load().t1_launch_count(1)
    eng.metrics_launches = 0
    ops.PROFILER.start()
    ms_total = timed(args.steps, False, 100)
    ops.PROFILER.stop()
    launches = int(lib.load().t1_launch_count(0)) + eng.metrics_launches
    clocks = sampler.stop() if rank == 0 else {}
    prof = ops.PROFILER.summary()
    ms_e2e = timed(args.steps, True, 200)
    filt = filter_pass(args, eng, cfg, rank, world, dev) if args.filter_pass > 0 else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_per_step = ms_total / args.steps
    value = world * NG * G / (ms_per_step / 1e3)
    e2e_value = world * NG * G / (ms_e2e / args.steps / 1e3)
    gm = prof.get("gemm_bf16_tn", dict(launches=0, flops=0.0, ms=1.0))
    gemm_tflops = gm["flops"] / max(gm["ms"], 1e-9) / 1e9
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    algo = {}
    for ids_i, _, grid_i in inputs:                   # this rank's prompt-groups (they differ under --grid timerft)
        for k, v in flops.step_flops(cfg, grid_i, len(ids_i), G, C).items():
            algo[k] = algo.get(k, 0.0) + v
    roofline = {
        "bound": "tensor", "kernel": "gemm2_bf16_tn (tcgen05 cta_group::2, 256x256 tiles on SM pairs; entry point t1_gemm_bf16_tn / t1_gemm_swiglu)", "achieved": round(gemm_tflops, 1), "peak": peak_tf, "unit": "TFLOP/s",
        "frac": round(gemm_tflops / peak_tf, 4), "traffic": NCU_GEMM_TRAFFIC["bytes"], "traffic_note": NCU_GEMM_TRAFFIC["note"], "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback",
        "launches_per_step": gm["launches"] // max(1, args.steps), "gemm_ms_per_step": round(gm["ms"] / args.steps, 2),
        "gemm_tflop_per_step": round(gm["flops"] / args.steps / 1e12, 1),
        "step_algorithmic_tflop": round(algo["total"] / 1e12, 1),
        "step_frac_of_flop_roofline": round(algo["total"] / (ms_per_step / 1e3) / 1e12 / peak_tf, 4),
    }
    Ps_all = [len(inp[0]) for inp in inputs]
    attn = attention_microbench(eng, Ps_all, G, C)
    attn["peak_gbs"] = peaks.get("hbm_gbs", 6650.0)
    attn["frac"] = round(attn["gbs"] / attn["peak_gbs"], 4)
    attn["in_graph"]["frac"] = round(attn["in_graph"]["gbs"] / attn["peak_gbs"], 4)
    dec = decode_step_microbench(eng, Ps_all, G, C)
    if dec is not None:
        dec["peak"] = peaks.get("hbm_gbs", 6650.0)
        dec["frac"] = round(dec["achieved"] / dec["peak"], 4)
        dec["peak_source"] = "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback"
        dec["share_of_step"] = round(dec["ms"] * (C - 1) / ms_per_step, 4)
    line = {
        "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_per_step, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": workload_name(args), "prompt_tokens": P, "vision_tokens": (P - 120), "group_size": G, "groups_per_step": NG,
                   "completion_len": C, "fix_vit": not args.train_vit, "parallelism": f"dp{world}+zero2-sharded-adamw", "l2": "inputs larger than L2 (>=16 GB of weights streamed per pass)"},
        "clocks": clocks,
        "e2e": {"value": round(e2e_value, 4), "unit": UNIT, "h2d_bytes_per_step": int(sum(inp[1].numel() + inp[0].size * 4 for inp in inputs)), "input": "uint8 frames [2*slots, 3, 448, 448] from pinned host memory -> t1_video_patchify on the device", "d2h_bytes_per_step": int(NG * ((G * C * 4 if args.reward_route == "host" else 0) + 4) + 4),
                "d2h": "per-group loss + grad norm" + (" + completion ids for the host reward parse" if args.reward_route == "host" else "; completions are detokenised and parsed on the device")},
        "gpu_launches": launches,
        "reward_route": args.reward_route,
        "hbm_gb": {"max_allocated": round(torch.cuda.max_memory_allocated(dev) / 1e9, 1), "max_reserved": round(torch.cuda.max_memory_reserved(dev) / 1e9, 1),
                   "total": round(torch.cuda.get_device_properties(dev).total_memory / 1e9, 1)},
        "roofline": roofline,
        "roofline_decode": dec,
        "attn": attn,
    }
    if filt is not None:
        line["filter_pass"] = filt
    line["save_level"] = getattr(eng, "last_save_level", None)
    if gpu_ref is not None:
        keep = ("value", "unit", "ms_per_step", "steps", "warmup", "how", "phase_s_per_step", "hbm_gb", "unavailable")
        line["gpu_reference"] = {k: gpu_ref[k] for k in keep if k in gpu_ref}
        if gpu_ref.get("value"):
            line["gpu_reference"]["speedup_e2e_over_gpu_reference"] = round(e2e_value / gpu_ref["value"], 2)
    if world == 1 and not args.no_cpu_baseline:
        del eng
        torch.cuda.empty_cache()
        line["cpu_baseline"] = cpu_reference(args, steps=2, warmup=1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_verify(args):
    """Same global batch at any world size -> same loss, gradient and post-step weights up to bf16 rounding (SURVEY §4 iii).
    8 prompt-groups with FIXED completions and rewards (no sampling: the test is about the data split, the backward's
    accumulation and the NCCL reduce-scatter / all-gather of optimizer_step); rank r takes groups r, r + W, ..."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from time_r1_b200 import config, engine, lib, ops, weights

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    lib.load()
    cfg = config.PRESETS[args.model]()
    G, C, NGLOBAL = args.group_size, 16, 8
    assert NGLOBAL % world == 0
    run = config.RunConfig(num_generations=G, max_completion_length=C, learning_rate=1e-3, fix_vit=not args.train_vit)
    mw = weights.ModelWeights(cfg, dev, world_size=world, train_vit=args.train_vit)
    mw.init_random(seed=0)
    eng = engine.Engine(cfg, mw, run, rank=rank, world_size=world, process_group=pg)
    with torch.no_grad():      # policy != reference model, so the KL term and its gradient are live
        eng.w.train["model.layers.1.mlp.down_proj.weight"].mul_(1.25)
        if args.train_vit:     # ... inside the vision tower as well (the reference model keeps its own tower)
            eng.w.train["visual.blocks.3.attn.proj.weight"].mul_(1.25)
    grid = (4, 32, 32)
    stride = 16381
    # TWO optimizer steps: the second one starts while the first one's parameter all-gather may still be in flight (lazy
    # per-bucket waits in Engine.p) and its loss depends on the updated weights
    for it in range(2):
        losses = []
        for gi in range(rank, NGLOBAL, world):
            ids, frames_host, g = synthetic_prompt(cfg, grid, gi, 0, dev)
            rng = np.random.default_rng(1000 + gi + 100 * it)
            comp = torch.from_numpy(rng.integers(0, 151643, (G, C))).int()
            comp[gi % G, 5 + gi] = cfg.eos_token_id
            comp[gi % G, 6 + gi:] = cfg.pad_token_id
            rewards = rng.random((G, 2)).astype(np.float32)
            prompt = engine.Prompt(input_ids=ids, pixel_values=None, grid_thw=g, second_per_grid_ts=[1.0], frames=frames_host.to(dev))
            out = eng.grpo_step(prompt, lambda c, r=rewards: r, completion_override=comp, loss_scale=world / NGLOBAL,
                                last_before_optimizer=gi + world >= NGLOBAL)
            losses.append(out["loss"].double())
        gsample = eng.grads.buf[::stride].float().clone()
        if it == 0 and args.train_vit:     # the tower's share of the gradient: fp32 DP mean of the ranks' accumulated gradients
            n_vit = eng.grads.layout["visual.merger.ln_q.weight"][0]
            gv = eng.grads.buf[:n_vit].float()
            if world > 1:
                dist.all_reduce(gv, op=dist.ReduceOp.AVG)
            vit_norm = float(gv.norm().item())
            vit_blocks_norms = [float(gv[eng.grads.layout[f"visual.blocks.{b}.norm1.weight"][0]:
                                         (eng.grads.layout[f"visual.blocks.{b + 1}.norm1.weight"][0] if b + 1 < cfg.vit_depth else n_vit)].norm().item())
                                for b in range(cfg.vit_depth)]
            del gv
        loss = torch.stack(losses).sum().reshape(1) / world      # each rank holds (its groups) x world / 8: the DP mean equals the N=1 sum
        if world > 1:
            dist.all_reduce(gsample, op=dist.ReduceOp.AVG)
            dist.all_reduce(loss, op=dist.ReduceOp.SUM)
        if it == 0:
            eng.sync_params()
            w_before = eng.w.train.buf[::stride].float().clone()
            loss0, gsample0 = loss.clone(), gsample
        gnorm = eng.optimizer_step(lr=1e-3)
        if it == 0:
            gnorm1 = float(gnorm.item())
    eng.sync_params()
    torch.cuda.synchronize()
    w_after = eng.w.train.buf[::stride].float().clone()
    if rank == 0:
        rec = {"world": world, "loss": float(loss0.item()), "loss_step2": float(loss.item()), "grad_norm": float(gnorm.item()),
               "grad_norm_step1": gnorm1,
               "grad_sample": gsample0.cpu(), "grad_sample_step2": gsample.cpu(),
               "w_before": w_before.cpu(), "w_after": w_after.cpu(), "stride": stride, "model": args.model, "groups": NGLOBAL, "G": G, "C": C}
        if args.train_vit:
            rec["vit_grad_norm"], rec["vit_block_grad_norms"] = vit_norm, vit_blocks_norms
        torch.save(rec, args.verify)
        print(json.dumps({"verify": {"world": world, "loss": rec["loss"], "loss_step2": rec["loss_step2"], "grad_norm": rec["grad_norm"],
                                     "grad_sample_l2": float(gsample.norm().item()), "out": args.verify}}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.verify:
        return run_verify(args)
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "hf_gpu":
        run_hf_gpu(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
