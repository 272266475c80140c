"""ORACLE tooling — pins the GRPO numerics to the UNMODIFIED reference trainer, by execution.

Imports /root/reference/src/time_r1/rl/timer1_trainer.py as it is (its absent third-party imports — trl, deepspeed's
zero-3 probe, src.utils' video reader — are stubbed in sys.modules; transformers.Trainer is the installed one), builds
a `TimeR1_Trainer` with `object.__new__` (its constructor needs accelerate/DeepSpeed and a checkpoint on disk), gives
it stand-ins for the processor / accelerator / model, and runs ITS `compute_loss` (timer1_trainer.py:512-782):

 (1) tests/golden/compute_loss_golden.npz — a scripted model (fixed logits as a leaf tensor, scripted `generate`
     output, a `batch_decode` table of completion texts) through lines 575-777 for use_grpo in {False, True} x beta in
     {0, 0.04}, rows with / without EOS, with the reference's OWN reward functions (main.py:184-239): inputs (policy /
     reference logits at the completion positions, completion ids, texts, solution, duration) and outputs (loss, the
     gradient d loss / d logits, every `_metrics` key).  tests pin oracle/qwen_vl_oracle.grpo_loss, oracle/grpo_oracle.c,
     oracle/hf_reference.grpo_loss_of and the CUDA kernels (t1_logprob_entropy / t1_grpo_loss / t1_logits_backward /
     t1_group_advantage / t1_eos_mask) to it.
 (2) tests/golden/compute_loss_hf_toy.json — the unmodified `compute_loss` driving the installed HF
     Qwen2_5_VLForConditionalGeneration (toy config, fp32, CPU) next to oracle/hf_reference.ReferenceStyleGRPO.step on
     the same weights, inputs and RNG state: the restatement used for the GPU baseline must give the same completions
     and the same loss.

Build container only (the reference does not exist on the GPU box):   python oracle/gen_golden_compute_loss.py
"""
import contextlib
import importlib.util
import io
import json
import os
import sys
import types
from collections import defaultdict

import numpy as np
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


@contextlib.contextmanager
def _unwrap_model_for_generation(model, accelerator, **kw):
    yield model


def import_reference_trainer(video_stub):
    """The unmodified module object.  `video_stub` stands in for src.utils.process_vision_info_v3 (decord video decode)."""
    class _GRPOConfig:        # only named in type hints / the constructor we do not run
        pass

    _stub("trl")
    _stub("trl.data_utils", apply_chat_template=lambda *a, **k: (_ for _ in ()).throw(NotImplementedError),
          is_conversational=lambda ex: isinstance(ex.get("prompt"), list) and isinstance(ex["prompt"][0], dict) and "role" in ex["prompt"][0])
    _stub("trl.models", create_reference_model=None, prepare_deepspeed=None, unwrap_model_for_generation=_unwrap_model_for_generation)
    _stub("trl.trainer")
    _stub("trl.trainer.grpo_config", GRPOConfig=_GRPOConfig)
    _stub("trl.trainer.utils", generate_model_card=None, get_comet_experiment_url=None)
    if "src" not in sys.modules:
        _stub("src")
    _stub("src.utils", process_vision_info_v3=video_stub)
    spec = importlib.util.spec_from_file_location("ref_timer1_trainer", os.path.join(REF, "src/time_r1/rl/timer1_trainer.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def import_reference_rewards():
    from oracle.gen_golden_rewards import import_reference_main

    keep = {k: sys.modules.get(k) for k in ("src", "src.time_r1", "trl")}
    with contextlib.redirect_stdout(io.StringIO()):
        ref_main = import_reference_main()
    for k, v in keep.items():           # gen_golden_rewards stubs these; restore ours
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v
    return ref_main


class _Accelerator:
    def __init__(self, device):
        self.device = torch.device(device)

    def gather_for_metrics(self, x):
        return x

    def unwrap_model(self, m):
        return m


class _Args:
    def __init__(self, device):
        self.device = torch.device(device)
        self.past_index = -1


class _BatchEncoding(dict):
    pass


class ScriptedProcessor:
    """Stand-in for the HF processor: chat template -> one string; __call__ -> the prepared prompt tensors;
    batch_decode -> scripted completion texts (what the tokenizer would produce is irrelevant to lines 575-777)."""
    eos_token_id = None
    pad_token_id = None

    def __init__(self, prompt_inputs, texts, eos_token_id, pad_token_id):
        self.prompt_inputs, self.texts = prompt_inputs, texts
        self.eos_token_id, self.pad_token_id = eos_token_id, pad_token_id

    def apply_chat_template(self, prompt, tokenize=False, add_generation_prompt=True):
        return "<prompt>"

    def __call__(self, **kw):
        assert kw["padding_side"] == "left" and kw["add_special_tokens"] is False
        return _BatchEncoding(self.prompt_inputs)

    def batch_decode(self, ids, skip_special_tokens=True):
        return list(self.texts) if self.texts is not None else [" ".join(str(int(t)) for t in row) for row in ids]


class ScriptedModel(torch.nn.Module):
    """`generate` returns scripted ids; forward returns fixed logits (a leaf tensor: its .grad is d loss / d logits)."""

    def __init__(self, pc_ids, logits):
        super().__init__()
        self.pc_ids = pc_ids
        self.logits_leaf = torch.nn.Parameter(logits.clone())

    def generate(self, **kw):
        assert kw["use_model_defaults"] is False and kw["generation_config"] is not None
        return self.pc_ids.clone()

    def forward(self, input_ids, attention_mask=None, pixel_values_videos=None, video_grid_thw=None):
        assert input_ids.shape == self.pc_ids.shape and attention_mask.shape == input_ids.shape
        G = input_ids.shape[0]
        assert pixel_values_videos.shape[0] % G == 0 and video_grid_thw.shape[0] == G       # pixels repeated G times (:594-599)
        return types.SimpleNamespace(logits=self.logits_leaf)


def new_trainer(mod, *, processing_class, G, beta, use_grpo, ref_model, reward_funcs, device="cpu", generation_config=object()):
    t = object.__new__(mod.TimeR1_Trainer)
    t.processing_class = processing_class
    t.num_generations = G
    t.generation_config = generation_config
    t.accelerator = _Accelerator(device)
    t.args = _Args(device)
    t.is_deepspeed_enabled = False
    t.beta, t.use_grpo, t.ref_model = beta, use_grpo, ref_model
    t.reward_funcs = list(reward_funcs)
    t.reward_processing_classes = [None] * len(reward_funcs)
    t.epsilon_low, t.epsilon_high = 0.2, 0.2               # timer1_trainer.py:388-393
    t.prompt_type = "v1"
    t._metrics = defaultdict(list)
    return t


TEXTS = [
    "<think>the person sits down</think>\n<answer>12.54 to 17.83</answer>",
    "<think>x</think><answer>0 to 5</answer>",
    "no tags at all 3 to 4",
    "<think>t</think><answer>10.00 to 20.00</answer>",
    "<think>a</think><answer>from 5.0 to 7.25 seconds</answer> trailing",
    "<think>t</think>\n\n<answer>The event happens in 14.00 TO 19.50.</answer>",
    "<think>t</think><answer>1 to 2 and 30</answer>",
    "<think></think><answer></answer>",
]


def scripted_cases(mod, ref_main):
    rng = np.random.default_rng(20251017)
    out = {}
    meta = []
    ci = 0
    for (G, C, V, P) in ((4, 6, 32, 5), (8, 9, 48, 3)):
        for use_grpo in (False, True):
            for beta in (0.0, 0.04):
                EOS, PAD = V - 2, V - 1
                comp = rng.integers(0, V - 2, (G, C))
                comp[0, 2] = EOS; comp[0, 3:] = PAD               # stops early
                comp[1, 0] = EOS; comp[1, 1:] = PAD               # first token is EOS
                comp[2, C - 1] = EOS                              # EOS in the last slot
                if G > 4:
                    comp[5, 4] = EOS; comp[5, 5] = EOS            # two EOS: only the first counts
                prompt = rng.integers(0, V - 2, (1, P))
                pc = torch.from_numpy(np.concatenate([np.repeat(prompt, G, 0), comp], 1)).long()
                logits = torch.from_numpy(rng.standard_normal((G, P + C, V)).astype(np.float32) * 2.0)
                ref_logits = logits + torch.from_numpy(rng.standard_normal((G, P + C, V)).astype(np.float32) * 0.3)
                texts = [TEXTS[(g + ci) % len(TEXTS)] for g in range(G)]
                sol, dur = [10.0, 20.0], 30.0 + ci
                example = dict(problem="person sits", video_path="v.mp4", video_start=None, video_end=None, solution=sol, durations=dur)
                prompt_inputs = dict(input_ids=torch.from_numpy(prompt).long(), attention_mask=torch.ones(1, P, dtype=torch.long),
                                     pixel_values_videos=torch.zeros(8, 4), video_grid_thw=torch.tensor([[2, 2, 2]]))
                proc = ScriptedProcessor(prompt_inputs, texts, EOS, PAD)
                model = ScriptedModel(pc, logits)
                ref_model = ScriptedModel(pc, ref_logits) if beta != 0.0 else None
                rfs = [ref_main.iou_timestamp_reward_v2, ref_main.format_reward] if ci % 2 == 0 else [ref_main.iou_timestamp_reward, ref_main.format_reward]
                t = new_trainer(mod, processing_class=proc, G=G, beta=beta, use_grpo=use_grpo, ref_model=ref_model, reward_funcs=rfs)
                with contextlib.redirect_stdout(io.StringIO()):
                    loss = t.compute_loss(model, [example])                          # <- the unmodified reference
                loss.backward()
                k = f"c{ci}_"
                out[k + "completion_ids"] = comp.astype(np.int32)
                out[k + "logits"] = logits[:, P - 1:-1].numpy()                      # rows that predict the completion tokens
                out[k + "ref_logits"] = ref_logits[:, P - 1:-1].numpy()
                out[k + "loss"] = np.float64(loss.item())
                g_full = model.logits_leaf.grad
                assert float(g_full[:, :P - 1].abs().max()) == 0.0 and float(g_full[:, -1].abs().max()) == 0.0
                out[k + "dlogits"] = g_full[:, P - 1:-1].numpy()
                metrics = {name: float(v[0]) for name, v in t._metrics.items()}
                meta.append(dict(case=ci, G=G, C=C, V=V, P=P, use_grpo=use_grpo, beta=beta, eos_token_id=EOS, pad_token_id=PAD, texts=texts,
                                 solution=sol, durations=dur, reward_funcs=[f.__name__ for f in rfs], metrics=metrics))
                ci += 1
    out["meta"] = np.array(json.dumps(meta))
    return out, meta


def hf_toy_case(mod):
    """(2): the unmodified compute_loss vs oracle/hf_reference on the installed HF classes, same seed."""
    from oracle import hf_reference as H
    from time_r1_b200 import config
    from transformers import GenerationConfig

    cfg = config.toy("qwen2_5_vl")
    G, C, beta = 4, 8, 0.04
    dev, dt = "cpu", torch.float32
    ref = H.ReferenceStyleGRPO(cfg, dev, dt, G=G, C=C, beta=beta, grad_ckpt=False, attn_impl="eager", seed=11)
    inputs = H.synthetic_prompt_inputs(cfg, (2, 8, 12), dev, dt, seed=5, n_text=12)
    rewards = torch.tensor(np.random.default_rng(1).random((G, 2)).astype(np.float32))

    class _Gen:      # transformers 5.5 dropped generate(use_model_defaults=...) that the reference passes (4.51.1): drop it, keep the rest
        def __init__(self, m):
            self.m = m

        def __call__(self, *a, **k):
            return self.m(*a, **k)

        def generate(self, **kw):
            kw.pop("use_model_defaults")
            self.m.eval()                       # as hf_reference does (KV cache on); sampling is mode-independent here
            try:
                return self.m.generate(**kw, use_cache=True)
            finally:
                self.m.train()

    def reward_a(prompts, completions, **kw):
        return rewards[:, 0].tolist()

    def reward_b(prompts, completions, **kw):
        return rewards[:, 1].tolist()

    proc = ScriptedProcessor(inputs, None, cfg.eos_token_id, cfg.pad_token_id)
    t = new_trainer(mod, processing_class=proc, G=G, beta=beta, use_grpo=False, ref_model=_Gen(ref.ref_model), reward_funcs=[reward_a, reward_b],
                    generation_config=ref.generation_config)
    example = dict(problem="q", video_path="v.mp4", video_start=None, video_end=None)
    torch.manual_seed(1234)
    with contextlib.redirect_stdout(io.StringIO()):
        loss_ref = t.compute_loss(_Gen(ref.model), [example])
    ref.model.zero_grad(set_to_none=True)
    loss_ref.backward()
    gsq = sum(float((p.grad.double() ** 2).sum()) for p in ref.trainable)
    ref.model.zero_grad(set_to_none=True)
    torch.manual_seed(1234)
    out = ref.step(inputs, lambda c: rewards, do_optimizer=True)
    rec = dict(loss_unmodified_reference=float(loss_ref.item()), loss_restatement=float(out["loss"].item()),
               grad_norm_unmodified_reference=gsq ** 0.5, grad_norm_restatement=float(out["grad_norm"]),
               metrics={k: float(v[0]) for k, v in t._metrics.items()}, completion_ids=out["completion_ids"].tolist(),
               G=G, C=C, beta=beta, seed_model=11, seed_inputs=5, seed_sampling=1234, rewards=rewards.tolist(),
               torch=torch.__version__, transformers=__import__("transformers").__version__)
    assert abs(rec["loss_unmodified_reference"] - rec["loss_restatement"]) < 1e-6, rec
    assert abs(rec["grad_norm_unmodified_reference"] - rec["grad_norm_restatement"]) < 1e-4 * max(1.0, rec["grad_norm_restatement"]), rec
    return rec


def main():
    def video_stub(conversations, return_video_kwargs=False):
        # process_vision_info_v3 (src/utils/vision_process.py:547-578): (image_inputs, video_inputs, {"fps": [...]})
        return None, [torch.zeros(4, 3, 28, 28)], {"fps": [2.0]}

    ref_main = import_reference_rewards()
    mod = import_reference_trainer(video_stub)
    out, meta = scripted_cases(mod, ref_main)
    np.savez_compressed(os.path.join(OUT, "compute_loss_golden.npz"), **out)
    for m in meta:
        print(m["case"], "G", m["G"], "use_grpo", m["use_grpo"], "beta", m["beta"], "loss", float(out[f"c{m['case']}_loss"]), m["metrics"])
    rec = hf_toy_case(mod)
    with open(os.path.join(OUT, "compute_loss_hf_toy.json"), "w") as f:
        json.dump(rec, f, indent=1)
    print("hf toy:", {k: rec[k] for k in ("loss_unmodified_reference", "loss_restatement", "grad_norm_unmodified_reference", "grad_norm_restatement")})


if __name__ == "__main__":
    main()
