"""Parameter storage: flat bf16 buffers in HBM with named views.

Layout decisions (B200-first):
  * two flat buffers — `frozen` (ViT patch-embed + blocks under --fix_vit, timer1_trainer.py:272-280) and `train`
    (merger + LLM) — so the optimizer, the grad-norm reduction and the ZeRO reduce-scatter / all-gather each touch
    one contiguous range;
  * q/k/v projections and gate/up projections are stored fused ([q;k;v] rows, [gate;up] rows) so one tcgen05 GEMM
    produces them; HF checkpoint names map to row ranges of the fused tensors;
  * the Qwen2.5-VL ViT MLP width (3420) is zero-padded to 3456 so every TMA row stride is a multiple of 16 bytes;
  * every tensor starts on a 256-byte boundary.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from .config import ModelConfig

ALIGN = 128  # elements (256 B)


@dataclass
class Spec:
    name: str
    shape: tuple
    group: str  # "vit" | "merger" | "llm"
    kind: str = "matrix"  # "matrix" | "norm" | "bias"


def param_specs(cfg: ModelConfig) -> list[Spec]:
    s: list[Spec] = []
    vd, vi, ip = cfg.vit_hidden, cfg.vit_intermediate, cfg.vit_inter_padded
    q25 = cfg.is_q25
    s.append(Spec("visual.patch_embed.proj.weight", (vd, cfg.patch_dim), "vit"))
    for i in range(cfg.vit_depth):
        p = f"visual.blocks.{i}."
        s.append(Spec(p + "norm1.weight", (vd,), "vit", "norm"))
        s.append(Spec(p + "norm2.weight", (vd,), "vit", "norm"))
        if not q25:
            s.append(Spec(p + "norm1.bias", (vd,), "vit", "bias"))
            s.append(Spec(p + "norm2.bias", (vd,), "vit", "bias"))
        s.append(Spec(p + "attn.qkv.weight", (3 * vd, vd), "vit"))
        s.append(Spec(p + "attn.qkv.bias", (3 * vd,), "vit", "bias"))
        s.append(Spec(p + "attn.proj.weight", (vd, vd), "vit"))
        s.append(Spec(p + "attn.proj.bias", (vd,), "vit", "bias"))
        if q25:
            s.append(Spec(p + "mlp.gate_up.weight", (2 * ip, vd), "vit"))
            s.append(Spec(p + "mlp.gate_up.bias", (2 * ip,), "vit", "bias"))
            s.append(Spec(p + "mlp.down_proj.weight", (vd, ip), "vit"))
            s.append(Spec(p + "mlp.down_proj.bias", (vd,), "vit", "bias"))
        else:
            s.append(Spec(p + "mlp.fc1.weight", (vi, vd), "vit"))
            s.append(Spec(p + "mlp.fc1.bias", (vi,), "vit", "bias"))
            s.append(Spec(p + "mlp.fc2.weight", (vd, vi), "vit"))
            s.append(Spec(p + "mlp.fc2.bias", (vd,), "vit", "bias"))
    m = cfg.merge_unit
    s.append(Spec("visual.merger.ln_q.weight", (vd,), "merger", "norm"))
    if not q25:
        s.append(Spec("visual.merger.ln_q.bias", (vd,), "merger", "bias"))
    s.append(Spec("visual.merger.mlp.0.weight", (vd * m, vd * m), "merger"))
    s.append(Spec("visual.merger.mlp.0.bias", (vd * m,), "merger", "bias"))
    s.append(Spec("visual.merger.mlp.2.weight", (cfg.hidden_size, vd * m), "merger"))
    s.append(Spec("visual.merger.mlp.2.bias", (cfg.hidden_size,), "merger", "bias"))
    D, I = cfg.hidden_size, cfg.intermediate_size
    s.append(Spec("model.embed_tokens.weight", (cfg.vocab_size, D), "llm"))
    for i in range(cfg.num_hidden_layers):
        p = f"model.layers.{i}."
        s.append(Spec(p + "input_layernorm.weight", (D,), "llm", "norm"))
        s.append(Spec(p + "post_attention_layernorm.weight", (D,), "llm", "norm"))
        s.append(Spec(p + "self_attn.qkv.weight", (cfg.qkv_dim, D), "llm"))
        s.append(Spec(p + "self_attn.qkv.bias", (cfg.qkv_dim,), "llm", "bias"))
        s.append(Spec(p + "self_attn.o_proj.weight", (D, cfg.num_attention_heads * cfg.head_dim), "llm"))
        s.append(Spec(p + "mlp.gate_up.weight", (2 * I, D), "llm"))
        s.append(Spec(p + "mlp.down_proj.weight", (D, I), "llm"))
    s.append(Spec("model.norm.weight", (D,), "llm", "norm"))
    if not cfg.tie_word_embeddings:
        s.append(Spec("lm_head.weight", (cfg.vocab_size, D), "llm"))
    return s


def _numel(shape) -> int:
    n = 1
    for d in shape:
        n *= d
    return n


class FlatParams:
    """One flat bf16 buffer + named views. `layout` is shared between params / grads / reference copies."""

    def __init__(self, specs: list[Spec], device, pad_to: int = ALIGN, buf: torch.Tensor | None = None):
        self.specs = specs
        self.pad_to = pad_to
        self.layout: dict[str, tuple[int, tuple]] = {}
        off = 0
        for sp in specs:
            self.layout[sp.name] = (off, sp.shape)
            off += (_numel(sp.shape) + ALIGN - 1) // ALIGN * ALIGN
        self.numel = (off + pad_to - 1) // pad_to * pad_to
        self.buf = buf if buf is not None else torch.zeros(self.numel, device=device, dtype=torch.bfloat16)
        assert self.buf.numel() == self.numel
        self.views: dict[str, torch.Tensor] = {}
        for name, (o, shape) in self.layout.items():
            self.views[name] = self.buf[o:o + _numel(shape)].view(shape)

    def __getitem__(self, name: str) -> torch.Tensor:
        return self.views[name]

    def __contains__(self, name: str) -> bool:
        return name in self.views

    def like(self, zero: bool = True) -> "FlatParams":
        buf = torch.zeros_like(self.buf) if zero else torch.empty_like(self.buf)
        return FlatParams(self.specs, self.buf.device, buf=buf, pad_to=self.pad_to)

    def clone(self) -> "FlatParams":
        return FlatParams(self.specs, self.buf.device, buf=self.buf.clone(), pad_to=self.pad_to)


class ModelWeights:
    """frozen (ViT) + train (merger + LLM) flat parameter sets.  train_vit=True (the reference's default fix_vit=False,
    main.py:46-49 / timer1_trainer.py:272-280): the ViT patch-embed and blocks live at the START of the trainable buffer
    (their gradients are the last to become final in the backward) and `frozen` is empty."""

    def __init__(self, cfg: ModelConfig, device="cuda", world_size: int = 1, train_vit: bool = False):
        self.cfg = cfg
        self.train_vit = bool(train_vit)
        specs = param_specs(cfg)
        self.frozen = FlatParams([s for s in specs if s.group == "vit" and not train_vit], device)
        # pad so the flat trainable buffer splits evenly into world_size shards of whole 128-element blocks
        self.train = FlatParams([s for s in specs if s.group != "vit" or train_vit], device, pad_to=ALIGN * world_size * 4)
        self.kinds = {s.name: s.kind for s in specs}

    def get(self, name: str) -> torch.Tensor:
        return self.train[name] if name in self.train else self.frozen[name]

    # ---- initialisation ----
    @torch.no_grad()
    def init_random(self, seed: int = 0, std: float = 0.02):
        """HF-style init (initializer_range 0.02, norm weights 1, biases N(0, std)) generated on the device."""
        g = torch.Generator(device=self.train.buf.device).manual_seed(seed)
        for fp in (self.frozen, self.train):
            for sp in fp.specs:
                v = fp[sp.name]
                if sp.kind == "norm":
                    v.fill_(1.0)
                else:
                    # chunked to bound the fp32 temporary
                    flat = v.view(-1)
                    step = 1 << 26
                    for a in range(0, flat.numel(), step):
                        n = min(step, flat.numel() - a)
                        flat[a:a + n] = (torch.randn(n, device=flat.device, generator=g) * std).to(torch.bfloat16)
        self._zero_vit_padding()

    def _zero_vit_padding(self):
        cfg = self.cfg
        if not cfg.is_q25 or cfg.vit_inter_padded == cfg.vit_intermediate:
            return
        vi, ip = cfg.vit_intermediate, cfg.vit_inter_padded
        for i in range(cfg.vit_depth):
            p = f"visual.blocks.{i}.mlp."
            w = self.get(p + "gate_up.weight")
            w[vi:ip].zero_()
            w[ip + vi:].zero_()
            b = self.get(p + "gate_up.bias")
            b[vi:ip].zero_()
            b[ip + vi:].zero_()
            self.get(p + "down_proj.weight")[:, vi:].zero_()

    # ---- HF checkpoint-name mapping ----
    def _hf_slices(self):
        """yield (hf_name, physical_name, row_slice or None, col_slice or None)"""
        cfg = self.cfg
        vi, ip = cfg.vit_intermediate, cfg.vit_inter_padded
        out = []
        for sp in param_specs(cfg):
            n = sp.name
            if n.endswith("self_attn.qkv.weight") or n.endswith("self_attn.qkv.bias"):
                suffix = n.rsplit(".", 1)[1]
                base = n[: -len("qkv." + suffix)]
                q, kv = cfg.num_attention_heads * cfg.head_dim, cfg.num_key_value_heads * cfg.head_dim
                out.append((base + "q_proj." + suffix, n, slice(0, q), None))
                out.append((base + "k_proj." + suffix, n, slice(q, q + kv), None))
                out.append((base + "v_proj." + suffix, n, slice(q + kv, q + 2 * kv), None))
            elif n.startswith("model.layers") and n.endswith("mlp.gate_up.weight"):
                base = n[: -len("gate_up.weight")]
                I = cfg.intermediate_size
                out.append((base + "gate_proj.weight", n, slice(0, I), None))
                out.append((base + "up_proj.weight", n, slice(I, 2 * I), None))
            elif n.startswith("visual.blocks") and ".mlp.gate_up." in n:
                suffix = n.rsplit(".", 1)[1]
                base = n[: -len("gate_up." + suffix)]
                out.append((base + "gate_proj." + suffix, n, slice(0, vi), None))
                out.append((base + "up_proj." + suffix, n, slice(ip, ip + vi), None))
            elif n.startswith("visual.blocks") and n.endswith("mlp.down_proj.weight"):
                out.append((n, n, None, slice(0, vi)))
            else:
                out.append((n, n, None, None))
        return out

    @torch.no_grad()
    def load_state_dict(self, sd: dict):
        """Load HF-checkpoint-named tensors (any float dtype, any device)."""
        missing = []
        for hf, phys, rs, cs in self._hf_slices():
            if hf not in sd:
                missing.append(hf)
                continue
            dst = self.get(phys)
            if rs is not None:
                dst = dst[rs]
            if cs is not None:
                dst = dst[:, cs]
            src = sd[hf]
            if hf == "visual.patch_embed.proj.weight":
                src = src.reshape(src.shape[0], -1)
            dst.copy_(src.to(device=dst.device, dtype=torch.bfloat16))
        if missing:
            raise KeyError(f"state dict is missing {len(missing)} tensors, e.g. {missing[:3]}")
        self._zero_vit_padding()

    @torch.no_grad()
    def export_state_dict(self) -> dict:
        """HF-checkpoint-named bf16 tensors (views where possible) for save_model."""
        cfg = self.cfg
        out = {}
        for hf, phys, rs, cs in self._hf_slices():
            t = self.get(phys)
            if rs is not None:
                t = t[rs]
            if cs is not None:
                t = t[:, cs]
            if hf == "visual.patch_embed.proj.weight":
                t = t.reshape(cfg.vit_hidden, cfg.in_channels, cfg.temporal_patch_size, cfg.patch_size, cfg.patch_size)
            out[hf] = t
        return out
