"""GRPO rollout-and-update engine: host-side composition of the libtimer1 kernels.

One prompt-group (one video + question, G sampled completions) is the unit of work, exactly as in the reference
(`TimeR1_Trainer.compute_loss`, src/time_r1/rl/timer1_trainer.py:512-782, batch size forced to 1 :548-551), but the
arithmetic is restructured around what is shared inside a group (all exact-math savings):

  * the ViT runs ONCE per prompt (reference: 3 x G times, timer1_trainer.py:594-596); under --fix_vit the frozen
    blocks also serve the reference model, only the merger differs;
  * the prompt is prefilled ONCE; the G decode streams share one copy of the prompt K/V (split-KV decode attention
    over [prefix | per-sequence slots]); the decode step is a CUDA graph replayed C-1 times;
  * scoring packs [prompt | G completions] into ONE token axis with a prefix-sharing attention mask, so the prompt is
    scored once and its K/V gradient accumulates from all G completions inside the attention backward;
  * lm_head is applied only to the G*C rows whose logits are used (reference: all G*(P+C) rows, :452-460);
  * backward recomputes each layer from its checkpointed input (gradient_checkpointing=true in train_rl.sh:32).

torch is used for allocation, streams, CUDA graphs and torch.distributed only; every numerical step is a libtimer1
kernel.  There is no CPU / eager fallback: every op raises if the extension is missing.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch

from . import indexing, ops
from .config import ModelConfig, RunConfig
from .lib import call, ptr, stream_ptr
from .lib import load as load_lib
from .weights import FlatParams, ModelWeights

BF16, F32, I32 = torch.bfloat16, torch.float32, torch.int32


@dataclass
class Prompt:
    """One prompt-group's inputs, as produced by the HF processor in the reference (timer1_trainer.py:547-565)."""
    input_ids: np.ndarray                 # [P] int
    pixel_values: torch.Tensor | None     # [N, patch_dim] fp32 on the device (pixel_values_videos), or None with `frames`
    grid_thw: list                        # [(t, h, w)]
    second_per_grid_ts: list | None = None
    # alternative to pixel_values: the resized frames [T, 3, H, W] (uint8 or fp32, 0..255) of ONE video on the device;
    # the HF processor's rescale / normalise / patchify then runs in t1_video_patchify, fused with the window gather
    frames: torch.Tensor | None = None


@dataclass
class ScoreState:
    """What the policy scoring pass keeps for backward."""
    P: int
    G: int
    C: int
    Tp: int
    ckpt: list
    cos: torch.Tensor
    sin: torch.Tensor
    lo: torch.Tensor
    hi: torch.Tensor
    ids_packed: torch.Tensor
    text_scatter_idx: torch.Tensor
    vision_rows: torch.Tensor
    logit_rows: torch.Tensor
    targets: torch.Tensor
    hsel: torch.Tensor
    lse: torch.Tensor
    x_vit: torch.Tensor
    plan: indexing.VisionPlan
    gu: list | None = None      # per-layer gate_up projection [Tp, 2I] kept from the forward passes (RunConfig.save_gate_up)
    k_qkv: list | None = None   # per-layer post-rope q, k, v [Tp, qkv_dim], attention output [Tp, Hq*hd] and LSE [Hq, Tp]
    k_attn: list | None = None  #   kept from the forward passes: the backward then also skips the qkv GEMM, rope and the
    k_lse: list | None = None   #   attention forward of its recompute
    vit_keep: dict | None = None  # fix_vit=False: what Engine.vit_blocks(keep=...) recorded for the ViT-block backward


class Engine:
    # defaults for objects built without __init__ (decode-only tools): no parameter all-gather is ever pending
    _ag_pending: dict = {}
    _ag_capped = False

    def __init__(self, cfg: ModelConfig, weights: ModelWeights, run: RunConfig, with_ref: bool = True,
                 rank: int = 0, world_size: int = 1, process_group=None):
        if not torch.cuda.is_available():
            raise RuntimeError("time_r1_b200.Engine needs a CUDA device (sm_100a); there is no CPU fallback")
        self.cfg, self.w, self.run = cfg, weights, run
        if bool(getattr(weights, "train_vit", False)) != (not run.fix_vit):
            raise ValueError(f"RunConfig.fix_vit={run.fix_vit} needs ModelWeights(train_vit={not run.fix_vit}): the ViT blocks live in the "
                             "trainable flat buffer exactly when they receive gradients")
        self.dev = weights.train.buf.device
        self.rank, self.world, self.pg = rank, world_size, process_group
        self.ref: FlatParams | None = weights.train.clone() if (with_ref and run.beta != 0.0) else None
        self.grads = weights.train.like(zero=True)
        self.grad_fresh = True
        # ---- ZeRO-2 shard of the optimizer state over the flat trainable buffer, in buckets (dist_utils.BucketPlan) ----
        import os as _os

        from .dist_utils import BucketPlan

        n = weights.train.numel
        assert n % (world_size * 512) == 0
        self.plan = BucketPlan(n, world_size, n_buckets=int(_os.environ.get("T1_ZERO_BUCKETS", getattr(run, "zero_buckets", 32))))
        self.shard = self.plan.shard
        self.shard_lo = 0 if world_size == 1 else None          # one contiguous range only without sharding
        self.master = torch.empty(self.shard, device=self.dev, dtype=F32)
        self.reset_master_from_params()
        self.adam_m = torch.zeros(self.shard, device=self.dev, dtype=F32)
        self.adam_v = torch.zeros(self.shard, device=self.dev, dtype=F32)
        self.shard_g = torch.empty(self.shard, device=self.dev, dtype=BF16) if world_size > 1 else None
        self.opt_step = 0
        # gradient reduce-scatter overlapped with the LAST micro-batch's backward: bucket b is reduced on NCCL's stream as soon
        # as the backward has finished every gradient at or above its start; a communicator limited to a few CTAs leaves the
        # tensor-bound backward its SMs (the collective has the whole backward to hide in)
        self.zero_overlap = world_size > 1 and _os.environ.get("T1_ZERO_OVERLAP", "1") != "0" and getattr(run, "zero_overlap", True)
        self.nccl_ctas = int(_os.environ.get("T1_NCCL_MAX_CTAS", getattr(run, "zero_nccl_ctas", 4)))
        self.pg_rs = process_group
        if self.zero_overlap and self.nccl_ctas > 0:
            import torch.distributed as dist

            if dist.get_backend(process_group) == "nccl":
                opts = dist.ProcessGroupNCCL.Options()
                opts.config.max_ctas = self.nccl_ctas
                opts.config.min_ctas = 1
                self.pg_rs = dist.new_group(backend="nccl", pg_options=opts)      # collective call: every rank builds an Engine
        self._grad_version = 0
        self._rs = None          # {"version", "works"}: reduce-scatters issued during the last backward
        # parameter all-gather left in flight by optimizer_step: {bucket: work}; consumers wait per bucket (Engine.p), so the
        # next step's merger / prefill (which read the buckets in the order they are gathered) start under the gather
        self._ag_pending: dict = {}
        self._ag_capped = False
        # measured at N=2 (profiles/r02_lazy_gather_n2.log): optimizer phase 41.5 -> 27.8 ms but the next step's ViT / merger
        # phase 55.9 -> 86.1 ms (the CTA-limited gather is slow and the GEMMs beside it run capped): 1465 -> 1474 ms per step.
        # Correct (two-step N-invariance run) but a net loss: OFF by default, T1_LAZY_GATHER=1 enables it
        self.lazy_gather = world_size > 1 and _os.environ.get("T1_LAZY_GATHER", "0") == "1" and getattr(run, "zero_lazy_gather", True)
        self.pg_ag = process_group
        ag_ctas = int(_os.environ.get("T1_NCCL_AG_CTAS", getattr(run, "zero_ag_ctas", 8)))
        self.ag_ctas = ag_ctas
        if self.lazy_gather and ag_ctas > 0:
            import torch.distributed as dist

            if dist.get_backend(process_group) == "nccl":
                if ag_ctas == self.nccl_ctas and self.pg_rs is not process_group:
                    self.pg_ag = self.pg_rs
                else:
                    opts = dist.ProcessGroupNCCL.Options()
                    opts.config.max_ctas = ag_ctas
                    opts.config.min_ctas = 1
                    self.pg_ag = dist.new_group(backend="nccl", pg_options=opts)
        # ---- rotary constants (computed the way HF does, in fp32 on the host) ----
        hd = cfg.head_dim
        self.inv_freq_llm = (1.0 / (cfg.rope_theta ** (torch.arange(0, hd, 2, dtype=F32) / hd))).to(self.dev)
        self.axis_llm = torch.from_numpy(indexing.mrope_channel_axes(cfg)).to(self.dev)
        vh = cfg.vit_head_dim // 2
        inv_v = 1.0 / (10000.0 ** (torch.arange(0, vh, 2, dtype=F32) / vh))
        self.inv_freq_vit = torch.cat([inv_v, inv_v]).to(self.dev)
        self.axis_vit = torch.cat([torch.zeros(vh // 2, dtype=I32), torch.ones(vh // 2, dtype=I32)]).to(self.dev)
        self._decode_cache: dict = {}
        self.metrics_launches = 0
        self.replay_hook = None
        # keep the gate_up projection [Tp, 2I] of every layer from the forward passes (12.5 GB per prompt-group at cfg2) when
        # HBM allows: the backward's recompute then skips its largest GEMM.  "auto" decides per step from the free memory.
        self.save_gate_up = getattr(run, "save_gate_up", "auto")

    def _save_level(self, n_groups: int, Tp: int) -> int:
        """How much of the forward passes to keep for the backward: 0 = layer inputs only (recompute everything),
        1 = + gate_up projections, 2 = + q/k/v, attention output and LSE.  "auto" decides from the free HBM."""
        cfg = self.cfg
        top = 2 if cfg.head_dim == 128 else 1      # the kept LSE layout needs the tcgen05 attention forward
        if self.save_gate_up != "auto":
            return min(top, 2 if self.save_gate_up is True else int(self.save_gate_up))
        L = cfg.num_hidden_layers
        need_gu = 2.0 * n_groups * L * Tp * 2 * cfg.intermediate_size
        need_at = 2.0 * n_groups * L * Tp * (cfg.qkv_dim + cfg.num_attention_heads * cfg.head_dim)
        free, _ = torch.cuda.mem_get_info(self.dev)
        avail = free + torch.cuda.memory_reserved(self.dev) - torch.cuda.memory_allocated(self.dev)
        margin = 14e9                        # head-room for the backward's transients and the logits chunks
        if top == 2 and avail - need_gu - need_at > margin:
            return 2
        return 1 if avail - need_gu > margin else 0

    def _alloc_keep(self, level: int, Tp: int):
        cfg = self.cfg
        L, d = cfg.num_hidden_layers, self.dev
        keep = {}
        if level >= 1:
            keep["gu"] = [torch.empty((Tp, 2 * cfg.intermediate_size), device=d, dtype=BF16) for _ in range(L)]
        if level >= 2:
            keep["qkv"] = [torch.empty((Tp, cfg.qkv_dim), device=d, dtype=BF16) for _ in range(L)]
            keep["attn"] = [torch.empty((Tp, cfg.num_attention_heads * cfg.head_dim), device=d, dtype=BF16) for _ in range(L)]
            keep["lse"] = [torch.empty((cfg.num_attention_heads, Tp), device=d, dtype=F32) for _ in range(L)]
        return keep or None

    @staticmethod
    def _keep_rows(keep, lo, hi):
        """Row window [lo, hi) of every kept per-layer buffer (LSE keeps its absolute [Hq, Tp] layout)."""
        if keep is None:
            return None
        return {k: (v if k == "lse" else [t[lo:hi] for t in v]) for k, v in keep.items()}

    def reset_master_from_params(self):
        """fp32 master weights of this rank's shard <- the bf16 parameters (construction, checkpoints without optimizer state)."""
        for b in range(len(self.plan)):
            lo, hi = self.plan.slice_of(b, self.rank)
            a, z = self.plan.shard_range(b)
            ops.bf16_to_f32(self.w.train.buf[lo:hi], self.master[a:z])

    def _reduce_ready(self, done_from: int):
        """Issue the asynchronous reduce-scatter of every bucket whose gradients are final (all offsets >= done_from)."""
        rs = self._rs
        if rs is None:
            return
        for b in self.plan.ready_buckets(done_from):
            if not rs["works"]:
                sms = torch.cuda.get_device_properties(self.dev).multi_processor_count
                load_lib().t1_gemm_max_ctas(max(1, sms - self.nccl_ctas) if self.nccl_ctas > 0 else 0)
                rs["capped"] = True
            rs["works"].append(self.plan.reduce_scatter(b, self.grads.buf, self.shard_g, self.rank, group=self.pg_rs, async_op=True))

    # ------------------------------------------------------------------------------------------
    # parameter access
    # ------------------------------------------------------------------------------------------
    def p(self, name: str, wset: str = "policy") -> torch.Tensor:
        if wset == "ref" and self.ref is not None and name in self.ref:
            return self.ref[name]
        if self._ag_pending and name in self.w.train:
            self._wait_param(name)
        return self.w.get(name)

    def _wait_param(self, name: str):
        """The all-gather of the last optimizer step may still be in flight (optimizer_step(lazy_gather=True)): order the
        current stream after the gathers of the buckets that hold `name`, in bucket order (they complete in that order)."""
        lo, shape = self.w.train.layout[name]
        n = 1
        for d in shape:
            n *= d
        hi = lo + n
        pend = self._ag_pending
        last = -1
        for b in sorted(pend):
            blo, bhi = self.plan.bounds[b]
            if blo < hi and bhi > lo:
                last = b
        if last < 0:
            return
        for b in sorted(pend):
            if b > last:
                break
            w = pend.pop(b)
            if w is not None:
                w.wait()
        if not pend:
            self._end_lazy_gather()

    def _end_lazy_gather(self):
        if self._ag_capped:
            load_lib().t1_gemm_max_ctas(0)
            self._ag_capped = False

    def sync_params(self):
        """Order the current stream after every outstanding parameter all-gather (before anything that reads the flat
        parameter buffer outside Engine.p(): checkpoint export, graph replays, direct buffer access)."""
        for b in sorted(self._ag_pending):
            w = self._ag_pending.pop(b)
            if w is not None:
                w.wait()
        self._end_lazy_gather()

    def lm_head(self, wset: str = "policy") -> torch.Tensor:
        return self.p("model.embed_tokens.weight" if self.cfg.tie_word_embeddings else "lm_head.weight", wset)

    def _i32(self, a) -> torch.Tensor:
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(self.dev, non_blocking=True)

    # ------------------------------------------------------------------------------------------
    # ViT
    # ------------------------------------------------------------------------------------------
    def vit_blocks(self, pixel_values, grid_thw, frames=None, wset: str = "policy",
                   keep: dict | None = None) -> tuple[torch.Tensor, torch.Tensor, indexing.VisionPlan]:
        """Patch-embed + all vision blocks. Returns (x, pending_delta, plan): the merger's input is x + pending_delta
        (the last block's MLP output, whose residual add is fused into the merger norm).
        keep (fix_vit=False): receives what _vit_backward needs -- the window-ordered pixel rows, the rotary tables, the window /
        full-attention row bounds and every block's input (the checkpoints the block recompute starts from)."""
        cfg = self.cfg
        plan = indexing.vision_plan(cfg, grid_thw)
        N, vd, H, hd = plan.n_patches, cfg.vit_hidden, cfg.vit_heads, cfg.vit_head_dim
        perm = self._i32(plan.perm_rows)
        if frames is not None:
            assert len(grid_thw) == 1, "the frames path takes one video per prompt"
            _, pix, g = ops.video_patchify(frames, cfg.patch_size, cfg.temporal_patch_size, cfg.spatial_merge_size,
                                           perm_rows=perm, want_f32=False, want_bf16=True)
            assert tuple(g) == tuple(int(v) for v in grid_thw[0]), (g, grid_thw)
        else:
            assert pixel_values.shape == (N, cfg.patch_dim) and pixel_values.dtype == F32
            pix = ops.gather_rows_f32(pixel_values, perm)                   # cast + window-order permutation
        g = lambda name: self.p(name, wset)
        x = ops.gemm(pix, g("visual.patch_embed.proj.weight"))              # Conv3d == GEMM
        cos, sin = ops.rope_table(self._i32(plan.pos_hw), self.axis_vit, self.inv_freq_vit)
        win_lo, win_hi = self._i32(plan.win_lo), self._i32(plan.win_hi)
        full_lo, full_hi = self._i32(plan.full_lo), self._i32(plan.full_hi)
        scale = hd ** -0.5
        import os as _os

        # measured on B200, 16384 tokens x 16 heads x 80 (profiles/r02_vit_attn_bench.txt): per-slot FULL attention (1024 keys)
        # tcgen05 0.209 ms vs FA2-varlen 0.300 vs HMMA 0.423; 64-token WINDOWS tcgen05 0.107 vs FA2 0.088 vs HMMA 0.061 (one
        # half-masked 128 x 128 tile per CTA: fixed cost, not tensor time) -> "auto" takes the faster kernel per block kind
        mode = _os.environ.get("T1_VIT_ATTN", getattr(self.run, "vit_attn", "auto"))
        tc_ok = hd in (80, 128)
        delta = None
        h = torch.empty_like(x)
        qkv = torch.empty((N, 3 * vd), device=self.dev, dtype=BF16)
        attn = torch.empty((N, vd), device=self.dev, dtype=BF16)
        proj = torch.empty((N, vd), device=self.dev, dtype=BF16)
        down = torch.empty((N, vd), device=self.dev, dtype=BF16)
        if keep is not None:
            keep.update(pix=pix, cos=cos, sin=sin, win=(win_lo, win_hi), full=(full_lo, full_hi), ckpt=[])
        for i in range(cfg.vit_depth):
            pre = f"visual.blocks.{i}."
            if cfg.is_q25:
                ops.rmsnorm(x, g(pre + "norm1.weight"), 1e-6, delta=delta, out=h)
            else:
                ops.layernorm(x, g(pre + "norm1.weight"), g(pre + "norm1.bias"), 1e-6, delta=delta, out=h)
            if keep is not None:
                keep["ckpt"].append(x.clone())                              # x now holds this block's input (delta added in place)
            ops.gemm(h, g(pre + "attn.qkv.weight"), out=qkv, bias=g(pre + "attn.qkv.bias"))
            ops.rope_apply(qkv, 0, 2 * H, hd, cos, sin)                     # q and k heads are contiguous
            full = (not cfg.is_q25) or (i in cfg.fullatt_block_indexes)
            lo, hi = (full_lo, full_hi) if full else (win_lo, win_hi)
            # windowed (64-token) / per-slot full attention, non-causal: tcgen05 kernel for head_dim 80 / 128 (the released
            # ViTs are 80), HMMA kernel for anything else
            if not full and tc_ok and plan.win_tile_local and mode in ("auto", "tc"):
                # 64-token windows that never straddle a 128-row tile: the persistent block-local tcgen05 kernel
                ops.attn_fwd_tc_win(qkv[:, :vd], qkv[:, vd:2 * vd], qkv[:, 2 * vd:], lo, hi, H, hd, scale, out=attn)
            else:
                use_tc = tc_ok and (mode == "tc" or (mode == "auto" and full))
                (ops.attn_fwd_tc if use_tc else ops.attn_fwd)(qkv[:, :vd], qkv[:, vd:2 * vd], qkv[:, 2 * vd:], lo, hi, 0, H, H, hd, scale,
                                                              out=attn, need_lse=False)
            ops.gemm(attn, g(pre + "attn.proj.weight"), out=proj, bias=g(pre + "attn.proj.bias"))
            if cfg.is_q25:
                ops.rmsnorm(x, g(pre + "norm2.weight"), 1e-6, delta=proj, out=h)
                act = ops.gemm_swiglu(h, g(pre + "mlp.gate_up.weight"), bias=g(pre + "mlp.gate_up.bias"))   # SwiGLU in the epilogue
                ops.gemm(act, g(pre + "mlp.down_proj.weight"), out=down, bias=g(pre + "mlp.down_proj.bias"))
            else:
                ops.layernorm(x, g(pre + "norm2.weight"), g(pre + "norm2.bias"), 1e-6, delta=proj, out=h)
                f1 = ops.gemm(h, g(pre + "mlp.fc1.weight"), bias=g(pre + "mlp.fc1.bias"))
                act = ops.act(f1, 1)
                ops.gemm(act, g(pre + "mlp.fc2.weight"), out=down, bias=g(pre + "mlp.fc2.bias"))
            delta = down
        return x, delta, plan

    def merger(self, x_vit: torch.Tensor, delta, plan: indexing.VisionPlan, wset: str = "policy",
               keep: dict | None = None) -> torch.Tensor:
        """PatchMerger (+ inverse window permutation). If delta is given, x_vit += delta first (in place)."""
        cfg = self.cfg
        N, vd, unit = plan.n_patches, cfg.vit_hidden, cfg.merge_unit
        if cfg.is_q25:
            y = ops.rmsnorm(x_vit, self.p("visual.merger.ln_q.weight", wset), 1e-6, delta=delta)
        else:
            y = ops.layernorm(x_vit, self.p("visual.merger.ln_q.weight", wset), self.p("visual.merger.ln_q.bias", wset),
                              1e-6, delta=delta)
        y4 = y.view(N // unit, unit * vd)
        a1 = ops.gemm(y4, self.p("visual.merger.mlp.0.weight", wset), bias=self.p("visual.merger.mlp.0.bias", wset))
        g1 = ops.act(a1, 0)
        m = ops.gemm(g1, self.p("visual.merger.mlp.2.weight", wset), bias=self.p("visual.merger.mlp.2.bias", wset))
        if keep is not None:
            keep.update(y4=y4, a1=a1, g1=g1)
        if cfg.is_q25:
            out = torch.empty_like(m)
            ops.scatter_rows(m, self._i32(plan.window_index), out)          # == m[argsort(window_index)]
            return out
        return m

    def prefetch_vit(self, prompts: list):
        """Run the FROZEN part of the vision tower (patch-embed + blocks; --fix_vit, timer1_trainer.py:272-280) for the NEXT
        step's prompts on a side stream.  It depends on no trained weight, so it can overlap this step's optimizer phase
        (gradient reduce-scatter, the HBM-bound AdamW kernel, parameter all-gather).  Pass the returned handle to
        grpo_multi_step(vit_prefetch=...); the merger (trained) still runs inside the step."""
        if not self.run.fix_vit:
            return None
        if getattr(self, "_vit_stream", None) is None:
            self._vit_stream = torch.cuda.Stream(device=self.dev)
        main = torch.cuda.current_stream()
        side = self._vit_stream
        side.wait_stream(main)                      # the prompts' pixel tensors / H2D copies were issued on `main`
        with torch.cuda.stream(side):
            res = [self.vit_blocks(p.pixel_values, p.grid_thw, frames=p.frames) for p in prompts]
            ev = torch.cuda.Event()
            ev.record(side)
        for x, delta, _ in res:                      # consumed on the main stream later
            x.record_stream(main)
            if delta is not None:
                delta.record_stream(main)
        return dict(results=res, event=ev, n=len(prompts))

    def ref_prefill_async(self, prompt: Prompt, emb_ref: torch.Tensor, G: int, C: int):
        """Prompt rows of the REFERENCE-model scoring pass (timer1_trainer.py:613-632) on a side stream.  They depend on the
        prompt only, not on the sampled tokens, and they are tensor-bound while the decode is HBM-bound: the side stream's
        GEMMs keep to RunConfig.side_ctas SMs and the decode (on a high-priority stream) to the others.  Returns the handle
        score(..., wset="ref", prefill=handle) consumes: the reference K/V of the prompt rows and its last hidden state."""
        cfg, run = self.cfg, self.run
        ids = np.asarray(prompt.input_ids)
        P = len(ids)
        L, kd = cfg.num_hidden_layers, cfg.num_key_value_heads * cfg.head_dim
        spg = None if run.ref_quirk_mrope else prompt.second_per_grid_ts
        pos3, _ = indexing.mrope_positions(cfg, ids, prompt.grid_thw, spg)
        if getattr(self, "_side_stream", None) is None:
            self._side_stream = torch.cuda.Stream(device=self.dev)            # default (lowest) priority
        main, side = torch.cuda.current_stream(), self._side_stream
        side.wait_stream(main)                                                  # emb_ref is ready
        lib = load_lib()
        with torch.cuda.stream(side):
            lib.t1_gemm_max_ctas(int(run.side_ctas))
            try:
                cos, sin = self._rope_tables(pos3)
                vis_rows = self._i32(np.nonzero(ids == cfg.video_token_id)[0])
                x0 = self._embed(self._i32(ids), emb_ref, vis_rows, "ref")
                kv = [(torch.empty((P, kd), device=self.dev, dtype=BF16), torch.empty((P, kd), device=self.dev, dtype=BF16)) for _ in range(L)]
                lo = torch.zeros(P, device=self.dev, dtype=I32)
                hi = torch.arange(1, P + 1, device=self.dev, dtype=I32)
                hN = self.llm_forward(x0, cos, sin, lo, hi, 0, wset="ref", kv_out=kv)
                h_last = hN[P - 1:P].clone()
            finally:
                lib.t1_gemm_max_ctas(0)
            ev = torch.cuda.Event()
            ev.record(side)
        for k_, v_ in kv:
            k_.record_stream(main)
            v_.record_stream(main)
        h_last.record_stream(main)
        return dict(kv_prefix=kv, h_last=h_last, event=ev, P=P, ref=True)

    def vit_forward(self, prompt: Prompt, wset: str = "policy"):
        x, delta, plan = self.vit_blocks(prompt.pixel_values, prompt.grid_thw, frames=prompt.frames,
                                         wset=wset if not self.run.fix_vit else "policy")
        emb = self.merger(x, delta, plan, wset)
        return emb, x, plan

    def _ref_embeds(self, prompt: Prompt, x_vit: torch.Tensor, plan: indexing.VisionPlan) -> torch.Tensor:
        """Vision embeddings under the REFERENCE model (timer1_trainer.py:613-632).  Frozen ViT blocks (--fix_vit) are shared
        with the policy, so only the merger is run again; with fix_vit=False the reference keeps its own (initial) tower."""
        if self.run.fix_vit:
            return self.merger(x_vit, None, plan, "ref")
        x, delta, _ = self.vit_blocks(prompt.pixel_values, prompt.grid_thw, frames=prompt.frames, wset="ref")
        return self.merger(x, delta, plan, "ref")

    # ------------------------------------------------------------------------------------------
    # LLM forward over a packed token axis
    # ------------------------------------------------------------------------------------------
    def _rope_tables(self, pos3: np.ndarray):
        return ops.rope_table(self._i32(pos3), self.axis_llm, self.inv_freq_llm)

    def _embed(self, ids_dev: torch.Tensor, vision_embeds: torch.Tensor, vision_rows: torch.Tensor, wset: str = "policy") -> torch.Tensor:
        x = ops.gather_rows(self.p("model.embed_tokens.weight", wset), ids_dev)
        ops.scatter_rows(vision_embeds, vision_rows, x)                     # masked_scatter (HF :1309-1315)
        return x

    @staticmethod
    def _attn(q, k, v, lo, hi, shared, Hq, Hkv, hd, scale, out=None, lse=None, need_lse=True):
        """LLM attention: tcgen05/TMEM kernel for head_dim 128, HMMA kernel otherwise (same contract)."""
        fn = ops.attn_fwd_tc if hd == 128 else ops.attn_fwd
        return fn(q, k, v, lo, hi, shared, Hq, Hkv, hd, scale, out=out, lse=lse, need_lse=need_lse)

    def llm_forward(self, x0, cos, sin, lo, hi, shared, wset="policy", ckpt=None, kv_out=None, keep=None):
        """x0 [T, D] (consumed). Returns final-norm hidden states [T, D]. ckpt: list of L+1 [T, D] buffers receiving
        the residual stream at each layer input (+ final); kv_out: per-layer (K, V) caches receiving rows [0, T)."""
        cfg = self.cfg
        T, D = x0.shape
        L, Hq, Hkv, hd, I = cfg.num_hidden_layers, cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim, cfg.intermediate_size
        qd, kd = Hq * hd, Hkv * hd
        eps, scale = cfg.rms_norm_eps, hd ** -0.5
        h = torch.empty((T, D), device=self.dev, dtype=BF16)
        keep_attn = keep is not None and "qkv" in keep
        gu_out = keep["gu"] if keep is not None and "gu" in keep else None
        qkv = None if keep_attn else torch.empty((T, cfg.qkv_dim), device=self.dev, dtype=BF16)
        attn = None if keep_attn else torch.empty((T, qd), device=self.dev, dtype=BF16)
        o = torch.empty((T, D), device=self.dev, dtype=BF16)
        act = torch.empty((T, I), device=self.dev, dtype=BF16)
        down = torch.empty((T, D), device=self.dev, dtype=BF16)
        xm = torch.empty((T, D), device=self.dev, dtype=BF16) if ckpt is not None else None
        x = x0
        if ckpt is not None:
            assert ckpt[0] is None or ckpt[0].data_ptr() == x0.data_ptr()
            ckpt[0] = x0
        for l in range(L):
            pre = f"model.layers.{l}."
            if l == 0:
                ops.rmsnorm(x, self.p(pre + "input_layernorm.weight", wset), eps, out=h)
            elif ckpt is not None:
                if ckpt[l] is None:
                    ckpt[l] = torch.empty((T, D), device=self.dev, dtype=BF16)
                ops.rmsnorm(xm, self.p(pre + "input_layernorm.weight", wset), eps, delta=down, out=h, x_out=ckpt[l])
                x = ckpt[l]
            else:
                ops.rmsnorm(x, self.p(pre + "input_layernorm.weight", wset), eps, delta=down, out=h)
            if keep_attn:      # q, k, v (post-rope), the attention output and its LSE stay for the backward
                qkv, attn = keep["qkv"][l], keep["attn"][l]
            ops.gemm(h, self.p(pre + "self_attn.qkv.weight", wset), out=qkv, bias=self.p(pre + "self_attn.qkv.bias", wset))
            if kv_out is not None:
                Kc, Vc = kv_out[l]
                ops.rope_apply(qkv, 0, Hq, hd, cos, sin)
                ops.rope_apply(qkv, qd, Hkv, hd, cos, sin, out=Kc)
                ops.copy_cols(qkv[:, qd + kd:], Vc, kd)
            else:
                ops.rope_apply(qkv, 0, Hq + Hkv, hd, cos, sin)
            self._attn(qkv[:, :qd], qkv[:, qd:qd + kd], qkv[:, qd + kd:], lo, hi, shared, Hq, Hkv, hd, scale, out=attn,
                       lse=keep["lse"][l] if keep_attn else None, need_lse=keep_attn)
            ops.gemm(attn, self.p(pre + "self_attn.o_proj.weight", wset), out=o)
            if ckpt is not None:
                ops.rmsnorm(x, self.p(pre + "post_attention_layernorm.weight", wset), eps, delta=o, out=h, x_out=xm)
            else:
                ops.rmsnorm(x, self.p(pre + "post_attention_layernorm.weight", wset), eps, delta=o, out=h)
            ops.gemm_swiglu(h, self.p(pre + "mlp.gate_up.weight", wset), out=act,             # SwiGLU in the GEMM epilogue
                            gate_up_out=gu_out[l] if gu_out is not None else None)
            ops.gemm(act, self.p(pre + "mlp.down_proj.weight", wset), out=down)
        hN = torch.empty((T, D), device=self.dev, dtype=BF16)
        if ckpt is not None:
            if ckpt[L] is None:
                ckpt[L] = torch.empty((T, D), device=self.dev, dtype=BF16)
            ops.rmsnorm(xm, self.p("model.norm.weight", wset), eps, delta=down, out=hN, x_out=ckpt[L])
        else:
            ops.rmsnorm(x, self.p("model.norm.weight", wset), eps, delta=down, out=hN)
        return hN

    def llm_forward_suffix(self, ckpt_full, P, cos, sin, lo, hi, kv_prefix, keep=None, wset="policy"):
        """Policy scoring of the completion rows only: rows [P, Tp) of the packed layout are computed, the prompt
        rows' K/V (same weights, same positions) come from the rollout's prefill cache and their layer inputs are
        already in ckpt_full[l][:P].  Fills ckpt_full[l][P:] and returns the final-norm hidden states of rows
        [P, Tp).  Exactly the arithmetic of llm_forward(ckpt=...) restricted to the rows that are new."""
        cfg = self.cfg
        Tp, D = ckpt_full[0].shape
        Tc = Tp - P
        L, Hq, Hkv, hd, I = cfg.num_hidden_layers, cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim, cfg.intermediate_size
        assert hd == 128, "prefix reuse needs the tcgen05 attention kernel (head_dim 128)"
        qd, kd = Hq * hd, Hkv * hd
        eps, scale = cfg.rms_norm_eps, hd ** -0.5
        dev = self.dev
        h = torch.empty((Tc, D), device=dev, dtype=BF16)
        keep_attn = keep is not None and "qkv" in keep     # full [Tp, ...] buffers whose prompt rows the prefill filled
        gu_out = [g[P:] for g in keep["gu"]] if keep is not None and "gu" in keep else None
        qkv = None if keep_attn else torch.empty((Tp, cfg.qkv_dim), device=dev, dtype=BF16)   # prompt rows: only K/V columns filled
        attn = None if keep_attn else torch.empty((Tc, qd), device=dev, dtype=BF16)
        o = torch.empty((Tc, D), device=dev, dtype=BF16)
        act = torch.empty((Tc, I), device=dev, dtype=BF16)
        down = torch.empty((Tc, D), device=dev, dtype=BF16)
        xm = torch.empty((Tc, D), device=dev, dtype=BF16)
        cos_c, sin_c = cos[P:], sin[P:]
        for l in range(L):
            pre = f"model.layers.{l}."
            x = ckpt_full[l][P:]
            if l == 0:
                ops.rmsnorm(x, self.p(pre + "input_layernorm.weight", wset), eps, out=h)
            else:
                ops.rmsnorm(xm, self.p(pre + "input_layernorm.weight", wset), eps, delta=down, out=h, x_out=x)
            if keep_attn:
                qkv, attn = keep["qkv"][l], keep["attn"][l][P:]
            ops.gemm(h, self.p(pre + "self_attn.qkv.weight", wset), out=qkv[P:], bias=self.p(pre + "self_attn.qkv.bias", wset))
            ops.rope_apply(qkv[P:], 0, Hq + Hkv, hd, cos_c, sin_c)
            if not keep_attn:      # otherwise the prompt rows of qkv already hold the prefill's rotated k and v
                Kc, Vc = kv_prefix[l]
                ops.copy_cols(Kc, qkv[:P, qd:qd + kd], kd)
                ops.copy_cols(Vc, qkv[:P, qd + kd:], kd)
            ops.attn_fwd_tc(qkv[P:, :qd], qkv[:, qd:qd + kd], qkv[:, qd + kd:], lo, hi, P, Hq, Hkv, hd, scale, out=attn,
                            lse=keep["lse"][l] if keep_attn else None, need_lse=keep_attn, q_row0=P)
            ops.gemm(attn, self.p(pre + "self_attn.o_proj.weight", wset), out=o)
            ops.rmsnorm(x, self.p(pre + "post_attention_layernorm.weight", wset), eps, delta=o, out=h, x_out=xm)
            ops.gemm_swiglu(h, self.p(pre + "mlp.gate_up.weight", wset), out=act, gate_up_out=gu_out[l] if gu_out is not None else None)
            ops.gemm(act, self.p(pre + "mlp.down_proj.weight", wset), out=down)
        hN = torch.empty((Tc, D), device=dev, dtype=BF16)
        ops.rmsnorm(xm, self.p("model.norm.weight", wset), eps, delta=down, out=hN, x_out=ckpt_full[L][P:])
        return hN

    # ------------------------------------------------------------------------------------------
    # rollout: prefill once + graph-captured decode
    # ------------------------------------------------------------------------------------------
    def _decode_state(self, Ps: tuple, G: int, C: int):
        """Static buffers of the decode graph for n = len(Ps) prompt-groups decoded together (B = n*G sequences)."""
        key = (tuple(Ps), G, C)
        st = self._decode_cache.get(key)
        if st is not None:
            return st
        cfg = self.cfg
        if len(self._decode_cache) >= 2:       # bound the KV memory kept alive
            self._decode_cache.clear()
        kd = cfg.num_key_value_heads * cfg.head_dim
        cpad = (C + 127) // 128 * 128
        L = cfg.num_hidden_layers
        ratio = cfg.num_attention_heads // cfg.num_key_value_heads
        rows_pad = (G * ratio + 63) // 64 * 64
        d = self.dev
        n = len(Ps)
        B = n * G
        groups = []
        for P in Ps:
            S = P + G * cpad
            n_chunks = (P + 127) // 128 + G * (cpad // 128)
            groups.append(dict(
                P=P,
                kv=[(torch.zeros((S, kd), device=d, dtype=BF16), torch.zeros((S, kd), device=d, dtype=BF16)) for _ in range(L)],
                ws_o=torch.empty(n_chunks * cfg.num_key_value_heads * rows_pad * 128, device=d, dtype=F32),
                ws_ml=torch.empty(n_chunks * cfg.num_key_value_heads * rows_pad * 2, device=d, dtype=F32),
                cos=torch.empty((C, cfg.head_dim // 2), device=d, dtype=F32), sin=torch.empty((C, cfg.head_dim // 2), device=d, dtype=F32)))
        st = dict(
            cpad=cpad, groups=groups, B=B, streams=[torch.cuda.Stream(device=d) for _ in range(n - 1)],
            x=torch.empty((B, cfg.hidden_size), device=d, dtype=BF16), h=torch.empty((B, cfg.hidden_size), device=d, dtype=BF16),
            qkv=torch.empty((B, cfg.qkv_dim), device=d, dtype=BF16),
            attn=torch.empty((B, cfg.num_attention_heads * cfg.head_dim), device=d, dtype=BF16),
            o=torch.empty((B, cfg.hidden_size), device=d, dtype=BF16),
            gu=torch.empty((B, 2 * cfg.intermediate_size), device=d, dtype=BF16),
            act=torch.empty((B, cfg.intermediate_size), device=d, dtype=BF16),
            down=torch.empty((B, cfg.hidden_size), device=d, dtype=BF16),
            hN=torch.empty((B, cfg.hidden_size), device=d, dtype=BF16),
            fold_parts=torch.zeros((2, max(1, cfg.hidden_size // 128), 64), device=d, dtype=F32),
            logits=torch.empty((B, cfg.vocab_size), device=d, dtype=F32),
            step=torch.zeros(1, device=d, dtype=I32), next_ids=torch.zeros(B, device=d, dtype=I32),
            finished=torch.zeros(B, device=d, dtype=I32), out_ids=torch.zeros((B, C), device=d, dtype=I32),
            eos=self._i32(list(cfg.stop_token_ids)), seed_dev=torch.zeros(1, device=d, dtype=torch.int64), graph=None, launches_per_step=0,
        )
        st["sample_ws"] = torch.empty(int(load_lib().t1_sample_topk_workspace_bytes(B, cfg.vocab_size)), device=d, dtype=torch.uint8)
        st["impl"] = self._decode_impl()
        if st["impl"] == "stream":
            assert G * ratio <= 64, "grouped decode attention handles G * Hq/Hkv <= 64 query rows per KV head"
            st["gemm_ws"] = ops.stream_gemm_workspace(B, d)
            # device table of t1_decode_group records (include/timer1.h), one row per layer
            rec = np.zeros((L, n, 7), dtype=np.int64)
            for gi, grp in enumerate(groups):
                for l in range(L):
                    Kc, Vc = grp["kv"][l]
                    rec[l, gi, :6] = (Kc.data_ptr(), Vc.data_ptr(), grp["ws_o"].data_ptr(), grp["ws_ml"].data_ptr(),
                                      grp["cos"].data_ptr(), grp["sin"].data_ptr())
                    rec[l, gi, 6] = grp["P"]          # little-endian: low int32 = prefix_len, high = reserved 0
            st["group_table"] = torch.from_numpy(rec).to(d)
            st["max_P"] = max(Ps)
            # per-layer TMA tensor maps over every group's K and V cache (t1_decode_attn_cluster)
            flat = [t for l in range(L) for grp in groups for t in grp["kv"][l]]
            st["kv_maps"] = ops.kv_cache_tensor_maps(flat, cfg.num_key_value_heads, cfg.head_dim).reshape(L, n * 2 * 128)
            # "auto": the cluster form holds 8 CTAs per (group, KV head) -- P / 1024 128-key units per CTA; from ~4 units per
            # CTA on, the wide form's 144 streaming CTAs win (profiles/r02_decode_step_final_by_context.txt)
            form = self.run.decode_attn
            if form == "auto":
                form = "wide" if max(Ps) >= int(getattr(self.run, "decode_attn_wide_min_prompt", 4096)) else "cluster"
            st["attn_form"] = form
            if form == "wide":
                st["attn_splits"], st["attn_ws"] = ops.decode_attn_wide_workspace(n, cfg.num_key_value_heads, d)
        self._decode_cache[key] = st
        return st

    def _sample(self, st, logits, ld, B, temperature, top_k, seed, seed_ptr, step_ptr, step_offset, finished, out_ids, next_ids):
        """temperature/top-k sampling of B rows; the chunked two-kernel sampler when top_k fits, else the general one."""
        cfg = self.cfg
        common = (ptr(logits), ld, B, cfg.vocab_size, float(temperature), int(top_k), int(seed), seed_ptr, step_ptr, step_offset,
                  ptr(st["eos"]), len(cfg.stop_token_ids), cfg.pad_token_id, ptr(finished), ptr(out_ids), st["out_ids"].stride(0),
                  ptr(next_ids), 0)
        fast = (not temperature > 0.0 or 1 <= top_k <= 64) and cfg.vocab_size <= 262144
        if fast:
            call("t1_sample_topk", *common, ptr(st["sample_ws"]), st["sample_ws"].numel(), stream_ptr())
        else:
            call("t1_sample", *common, stream_ptr())

    def _decode_impl(self) -> str:
        cfg = self.cfg
        ok = cfg.head_dim == 128 and cfg.hidden_size % 8 == 0 and cfg.intermediate_size % 8 == 0
        return "stream" if (self.run.decode_impl == "stream" and ok) else "legacy"

    def _decode_step_stream(self, st, G: int, temperature: float, top_k: int):
        """One token for each of the B = n*G sequences as ONE chain of PDL-linked kernels: every linear streams its
        weights through the stream-K tcgen05 kernel (whose weight prefetch runs ahead of the dependency), rope/KV-append
        and the shared-prefix attention cover all groups in one launch each."""
        cfg = self.cfg
        L, Hq, Hkv, hd = cfg.num_hidden_layers, cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
        eps, scale = cfg.rms_norm_eps, hd ** -0.5
        B, n = st["B"], len(st["groups"])
        x, h, qkv, attn, o, gu, act, down = st["x"], st["h"], st["qkv"], st["attn"], st["o"], st["gu"], st["act"], st["down"]
        ws, tab = st["gemm_ws"], st["group_table"]
        from .lib import load
        load().t1_stream_gemm_max_ctas(int(st.get("decode_ctas", 0)))      # leave SMs to the overlapped reference prefill
        try:
            s = stream_ptr()
            # The step's FIRST kernel is launched in plain stream order (no programmatic edge): kernels further down the chain read
            # state of the previous step before their own griddepcontrol.wait (the attention prologue reads the step counter and
            # prefetches K/V), which is only safe if nothing of this step starts before the previous step has completed.  Inside a
            # captured graph this is the graph boundary anyway; it matters for eager multi-step runs (use_graph=False).
            load().t1_set_pdl(0)
            ops.gather_rows(self.p("model.embed_tokens.weight"), st["next_ids"], out=x)
            load().t1_set_pdl(1 if self.run.decode_pdl else 0)
            fuse_norm = self.run.decode_fuse_norm and cfg.hidden_size % 8 == 0
            # folded norms: the o / down GEMM epilogues update the residual stream and emit z = norm_w * resid plus per-slab
            # row sums of squares; the next GEMM streams against z and applies the row factor in its epilogue
            form = st.get("attn_form", self.run.decode_attn)
            wide = form == "wide"

            def attn_tc(l):
                """rope + KV append + shared-prefix attention of layer l, all groups, one tcgen05 launch"""
                common = (ptr(qkv), qkv.stride(0), ptr(tab[l]), st["kv_maps"][l].ctypes.data, n, G, Hq, Hkv, hd, st["cpad"],
                          ptr(st["step"]), ptr(attn), attn.stride(0), scale)
                if wide:      # every SM streams K/V; partials merged through the L2-resident workspace
                    call("t1_decode_attn_wide", *common, st["attn_splits"], ptr(st["attn_ws"]), st["attn_ws"].numel(), s)
                else:         # 8-CTA cluster per (group, KV head), partials merged through distributed shared memory
                    call("t1_decode_attn_cluster", *common, 0, s)

            fold = (self.run.decode_fold_norm and not fuse_norm and form in ("cluster", "wide") and cfg.hidden_size % 128 == 0
                    and cfg.intermediate_size % 64 == 0)
            D = cfg.hidden_size
            fp = st["fold_parts"]
            pf_bytes = int(self.run.decode_prefetch_mb) << 20

            def prefetch_gate_up(l):
                """hint for the NEXT stream GEMM launch: its idle warps pull layer l's first gate_up units into L2"""
                w = self.p(f"model.layers.{l}.mlp.gate_up.weight")
                call("t1_stream_gemm_prefetch_next", ptr(w), w.stride(0), B, w.shape[0], w.shape[1], 1, 0, pf_bytes)

            # paired launches (RunConfig.decode_pair): o_proj + gate_up and down_proj + the next layer's qkv run as ONE kernel
            # each (t1_stream_gemm_fold_pair) -- 3 kernels per layer instead of 5, the GEMM -> GEMM boundary is an arrival counter
            pair = fold and self.run.decode_pair and not pf_bytes and st.get("decode_ctas", 0) in (0, None)
            # CTAs of the o_proj GEMM (0 = all SMs).  The cluster-form attention launch holds 8 * Hkv * n SMs; o_proj CTAs that only
            # become resident when those exit start their weight load on the chain's critical path
            o_ctas = int(getattr(self.run, "decode_o_ctas", 0)) if (form == "cluster" and not st.get("decode_ctas", 0)) else 0
            qkv_ctas = int(getattr(self.run, "decode_qkv_ctas", 0)) if not st.get("decode_ctas", 0) else 0
            down_ctas = int(getattr(self.run, "decode_down_ctas", 0)) if not st.get("decode_ctas", 0) else 0
            if o_ctas < 0:
                o_ctas = max(16, torch.cuda.get_device_properties(self.dev).multi_processor_count - 8 * Hkv * n)
            for l in range(L):
                pre = f"model.layers.{l}."
                if pair:
                    last = l == L - 1
                    if l == 0:
                        ops.rmsnorm(x, self.p(pre + "input_layernorm.weight"), eps, out=h)
                        ops.stream_gemm(h, self.p(pre + "self_attn.qkv.weight"), ws, out=qkv, bias=self.p(pre + "self_attn.qkv.bias"))
                    attn_tc(l)
                    ops.stream_gemm_fold_pair(attn, self.p(pre + "self_attn.o_proj.weight"), ws, x, self.p(pre + "post_attention_layernorm.weight"),
                                              h, fp[0], self.p(pre + "mlp.gate_up.weight"), eps, out=act, swiglu=True)
                    if last:      # the final norm feeds the dynamically scheduled lm_head GEMM: kept as a kernel
                        ops.stream_gemm(act, self.p(pre + "mlp.down_proj.weight"), ws, out=down)
                    else:
                        nxt = f"model.layers.{l + 1}."
                        ops.stream_gemm_fold_pair(act, self.p(pre + "mlp.down_proj.weight"), ws, x, self.p(nxt + "input_layernorm.weight"),
                                                  h, fp[1], self.p(nxt + "self_attn.qkv.weight"), eps, out=qkv,
                                                  bias=self.p(nxt + "self_attn.qkv.bias"))
                    continue
                if fold:
                    last = l == L - 1
                    if pf_bytes and self.run.decode_prefetch_at == "qkv":
                        prefetch_gate_up(l)
                    if l == 0:
                        ops.rmsnorm(x, self.p(pre + "input_layernorm.weight"), eps, out=h)
                        ops.stream_gemm(h, self.p(pre + "self_attn.qkv.weight"), ws, out=qkv, bias=self.p(pre + "self_attn.qkv.bias"))
                    else:
                        if qkv_ctas:
                            load().t1_stream_gemm_max_ctas(qkv_ctas)
                        ops.stream_gemm_scaled(h, self.p(pre + "self_attn.qkv.weight"), ws, fp[1], D // 128, eps, out=qkv,
                                               bias=self.p(pre + "self_attn.qkv.bias"))
                        if qkv_ctas:
                            load().t1_stream_gemm_max_ctas(int(st.get("decode_ctas", 0)))
                    attn_tc(l)
                    if pf_bytes and self.run.decode_prefetch_at == "o":
                        prefetch_gate_up(l)
                    if o_ctas:     # o_proj only on SMs the attention launch leaves free: every CTA is resident (ring full) when attention ends
                        load().t1_stream_gemm_max_ctas(o_ctas)
                    ops.stream_gemm_fold(attn, self.p(pre + "self_attn.o_proj.weight"), ws, x,
                                         self.p(pre + "post_attention_layernorm.weight"), h, fp[0])
                    if o_ctas:
                        load().t1_stream_gemm_max_ctas(int(st.get("decode_ctas", 0)))
                    ops.stream_gemm_scaled(h, self.p(pre + "mlp.gate_up.weight"), ws, fp[0], D // 128, eps, out=act, swiglu=True)
                    if last:      # the final norm feeds the dynamically scheduled lm_head GEMM: kept as a kernel
                        ops.stream_gemm(act, self.p(pre + "mlp.down_proj.weight"), ws, out=down)
                    else:
                        if down_ctas:
                            load().t1_stream_gemm_max_ctas(down_ctas)
                        ops.stream_gemm_fold(act, self.p(pre + "mlp.down_proj.weight"), ws, x,
                                             self.p(f"model.layers.{l + 1}.input_layernorm.weight"), h, fp[1])
                        if down_ctas:
                            load().t1_stream_gemm_max_ctas(int(st.get("decode_ctas", 0)))
                    continue
                if l == 0 or not fuse_norm:
                    ops.rmsnorm(x, self.p(pre + "input_layernorm.weight"), eps, delta=down if (l > 0 and not fuse_norm) else None, out=h)
                ops.stream_gemm(h, self.p(pre + "self_attn.qkv.weight"), ws, out=qkv, bias=self.p(pre + "self_attn.qkv.bias"))
                # rope + KV append are fused into the attention kernel
                if form in ("cluster", "wide"):
                    attn_tc(l)
                else:
                    call("t1_decode_attn_groups", ptr(qkv), qkv.stride(0), ptr(tab[l]), n, G, Hq, Hkv, hd, st["max_P"], st["cpad"],
                         ptr(st["step"]), ptr(attn), attn.stride(0), scale, 1, s)
                if fuse_norm:      # o projection + residual add + post-attention RMSNorm in one launch
                    ops.stream_gemm_resnorm(attn, self.p(pre + "self_attn.o_proj.weight"), ws, x,
                                            self.p(pre + "post_attention_layernorm.weight"), eps, h, out=o)
                else:
                    ops.stream_gemm(attn, self.p(pre + "self_attn.o_proj.weight"), ws, out=o)
                    ops.rmsnorm(x, self.p(pre + "post_attention_layernorm.weight"), eps, delta=o, out=h)
                if cfg.intermediate_size % 64 == 0:
                    ops.stream_gemm_swiglu(h, self.p(pre + "mlp.gate_up.weight"), ws, out=act)    # SwiGLU in the epilogue
                else:
                    ops.stream_gemm(h, self.p(pre + "mlp.gate_up.weight"), ws, out=gu)
                    ops.swiglu(gu, out=act)
                if fuse_norm:      # down projection + residual add + the NEXT layer's input RMSNorm (or the final norm)
                    last = l == L - 1
                    ops.stream_gemm_resnorm(act, self.p(pre + "mlp.down_proj.weight"), ws, x,
                                            self.p("model.norm.weight" if last else f"model.layers.{l + 1}.input_layernorm.weight"),
                                            eps, st["hN"] if last else h, out=down)
                else:
                    ops.stream_gemm(act, self.p(pre + "mlp.down_proj.weight"), ws, out=down)
            if not fuse_norm:
                ops.rmsnorm(x, self.p("model.norm.weight"), eps, delta=down, out=st["hN"])
            ops.stream_gemm(st["hN"], self.lm_head(), ws, out=st["logits"])
            self._sample(st, st["logits"], st["logits"].stride(0), B, temperature, top_k, 0, ptr(st["seed_dev"]), ptr(st["step"]), 1,
                         st["finished"], st["out_ids"], st["next_ids"])
            call("t1_incr", ptr(st["step"]), 1, s)
        finally:
            load().t1_set_pdl(0)
            load().t1_stream_gemm_max_ctas(0)

    def _decode_step(self, st, G: int, temperature: float, top_k: int):
        """One token for each of the B = n*G sequences (captured into a CUDA graph by rollout_multi()). The linears
        run once over all B rows (the weight stream is amortised over the groups); rope/KV-append and the
        shared-prefix attention run per group because each group has its own prompt."""
        if st.get("impl") == "stream":
            return self._decode_step_stream(st, G, temperature, top_k)
        cfg = self.cfg
        L, Hq, Hkv, hd = cfg.num_hidden_layers, cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
        eps, scale = cfg.rms_norm_eps, hd ** -0.5
        s = stream_ptr()
        B = st["B"]
        x, h, qkv, attn, o, gu, act, down = st["x"], st["h"], st["qkv"], st["attn"], st["o"], st["gu"], st["act"], st["down"]
        ops.gather_rows(self.p("model.embed_tokens.weight"), st["next_ids"], out=x)
        for l in range(L):
            pre = f"model.layers.{l}."
            ops.rmsnorm(x, self.p(pre + "input_layernorm.weight"), eps, delta=down if l > 0 else None, out=h)
            ops.skinny_gemm(h, self.p(pre + "self_attn.qkv.weight"), out=qkv, bias=self.p(pre + "self_attn.qkv.bias"))
            # per-group rope/KV-append + shared-prefix attention are latency-bound and independent of each other:
            # fork them onto side streams (parallel branches of the captured graph), join before the o-projection
            main = torch.cuda.current_stream()
            for gi, grp in enumerate(st["groups"]):
                Kc, Vc = grp["kv"][l]
                q_g, a_g = qkv[gi * G:(gi + 1) * G], attn[gi * G:(gi + 1) * G]
                side = main if gi == 0 else st["streams"][gi - 1]
                if gi > 0:
                    side.wait_stream(main)
                with torch.cuda.stream(side):
                    sp = stream_ptr()
                    call("t1_decode_rope_append", ptr(q_g), qkv.stride(0), G, Hq, Hkv, hd, ptr(grp["cos"]), ptr(grp["sin"]),
                         ptr(st["step"]), ptr(Kc), ptr(Vc), grp["P"], st["cpad"], sp)
                    call("t1_decode_attn", ptr(q_g), qkv.stride(0), ptr(Kc), ptr(Vc), G, Hq, Hkv, hd, grp["P"], st["cpad"],
                         ptr(st["step"]), ptr(grp["ws_o"]), ptr(grp["ws_ml"]), ptr(a_g), attn.stride(0), scale, sp)
            for gi in range(1, len(st["groups"])):
                main.wait_stream(st["streams"][gi - 1])
            ops.skinny_gemm(attn, self.p(pre + "self_attn.o_proj.weight"), out=o)
            ops.rmsnorm(x, self.p(pre + "post_attention_layernorm.weight"), eps, delta=o, out=h)
            ops.skinny_gemm(h, self.p(pre + "mlp.gate_up.weight"), out=gu)
            ops.swiglu(gu, out=act)
            ops.skinny_gemm(act, self.p(pre + "mlp.down_proj.weight"), out=down)
        ops.rmsnorm(x, self.p("model.norm.weight"), eps, delta=down, out=st["hN"])
        ops.skinny_gemm(st["hN"], self.lm_head(), out=st["logits"])
        self._sample(st, st["logits"], st["logits"].stride(0), B, temperature, top_k, 0, ptr(st["seed_dev"]), ptr(st["step"]), 1,
                     st["finished"], st["out_ids"], st["next_ids"])
        call("t1_incr", ptr(st["step"]), 1, s)

    def rollout(self, prompt: Prompt, vision_embeds: torch.Tensor, G: int, C: int, temperature: float = 1.0,
                top_k: int = 50, seed: int = 0, use_graph: bool = True, check_every: int = 32) -> torch.Tensor:
        """Sample G completions of up to C tokens (temperature 0 => greedy). Returns ids [G, C] int32 on the device,
        pad_token_id after a stop token (HF generate semantics, timer1_trainer.py:568-577)."""
        return self.rollout_multi([prompt], [vision_embeds], G, C, temperature, top_k, seed, use_graph, check_every)[0]

    def rollout_multi(self, prompts: list, vision_embeds: list, G: int, C: int, temperature: float = 1.0,
                      top_k: int = 50, seed: int = 0, use_graph: bool = True, check_every: int = 32,
                      keep_prefill: bool = False, shared_sms: bool = False, before_decode=None) -> list:
        """Rollout for several prompt-groups at once (the micro-batches of one optimizer step see the same weights,
        so their decode streams can share every weight read). Returns one [G, C] int32 tensor per group."""
        cfg = self.cfg
        Ps = tuple(len(np.asarray(p.input_ids)) for p in prompts)
        st = self._decode_state(Ps, G, C)
        B = st["B"]
        # shared_sms: a side stream (the reference model's prompt prefill) runs during the decode: the weight-streaming
        # GEMMs keep to decode_ctas SMs and the captured step is replayed on a HIGH-priority stream, so that at every kernel
        # boundary the decode's CTAs are placed before the side stream's pending ones
        want_ctas = int(self.run.decode_ctas) if (shared_sms and st["impl"] == "stream") else 0
        if st.get("decode_ctas", 0) != want_ctas:
            st["decode_ctas"], st["graph"] = want_ctas, None
        st["step"].zero_()
        st["finished"].zero_()
        st["out_ids"].fill_(cfg.pad_token_id)
        st["seed_dev"].fill_(int(seed))
        save_level = self._save_level(len(prompts), max(indexing.packed_rows(P_, G, C)[0] for P_ in Ps)) if keep_prefill else 0
        if keep_prefill:
            self.last_save_level = save_level    # reported by bench.py: which fallback the free HBM allowed
        for gi, (prompt, emb, grp) in enumerate(zip(prompts, vision_embeds, st["groups"])):
            ids = np.asarray(prompt.input_ids)
            P = grp["P"]
            pos3, delta = indexing.mrope_positions(cfg, ids, prompt.grid_thw, prompt.second_per_grid_ts)
            cos, sin = self._rope_tables(pos3)
            ids_dev = self._i32(ids)
            vis_rows = self._i32(np.nonzero(ids == cfg.video_token_id)[0])
            lo = torch.zeros(P, device=self.dev, dtype=I32)
            hi = torch.arange(1, P + 1, device=self.dev, dtype=I32)
            kv_prefix = [(k[:P], v[:P]) for k, v in grp["kv"]]
            ckpt = None
            if keep_prefill:
                # layer-input checkpoints of the packed scoring layout; the prefill fills the prompt rows now, the
                # policy scoring pass fills the completion rows later (engine.score(reuse=...))
                Tp = indexing.packed_rows(P, G, C)[0]
                L = cfg.num_hidden_layers
                ckpt_full = [torch.empty((Tp, cfg.hidden_size), device=self.dev, dtype=BF16) for _ in range(L + 1)]
                ckpt = [c[:P] for c in ckpt_full]
                keep_full = self._alloc_keep(save_level, Tp)    # what the backward will not have to recompute
                x0 = ops.gather_rows(self.p("model.embed_tokens.weight"), ids_dev, out=ckpt[0])
                ops.scatter_rows(emb, vis_rows, x0)
            else:
                x0 = self._embed(ids_dev, emb, vis_rows)
            keep_full = keep_full if keep_prefill else None
            hN = self.llm_forward(x0, cos, sin, lo, hi, 0, ckpt=ckpt, kv_out=kv_prefix, keep=self._keep_rows(keep_full, 0, P))
            if keep_prefill:
                grp["prefill"] = dict(ckpt_full=ckpt_full, h_last=hN[P - 1:P].clone(), kv_prefix=kv_prefix, keep=keep_full,
                                      spg=None if prompt.second_per_grid_ts is None else list(prompt.second_per_grid_ts))
            logits0 = ops.skinny_gemm(hN[P - 1:P], self.lm_head(), out_dtype=F32)           # [1, V]
            # decode-position tables: completion token c sits at text position P + c + rope_delta on all three axes
            dpos = np.tile((P + delta + np.arange(C, dtype=np.int64))[None, :], (3, 1))
            dcos, dsin = self._rope_tables(dpos)
            grp["cos"].copy_(dcos)
            grp["sin"].copy_(dsin)
            # first token: G independent draws from the shared last-prompt-position distribution (row stride 0)
            sl = slice(gi * G, (gi + 1) * G)
            self._sample(st, logits0, 0, G, temperature, top_k, int(seed) + 7919 * (gi + 1), 0, 0, 0,
                         st["finished"][sl], st["out_ids"][sl], st["next_ids"][sl])
        n_steps = C - 1
        self.sync_params()           # the captured decode step reads every weight through raw pointers
        if before_decode is not None:
            before_decode()          # side-stream work that overlaps the decode loop is enqueued after the prompt prefills
        if n_steps > 0:
            gkey = (float(temperature), int(top_k))     # the seed lives in device memory: graphs are seed-independent
            if use_graph:
                if st["graph"] is None or st.get("gkey") != gkey:
                    # warm-up step outside capture (sets func attributes), then rewind the state it touched
                    snap = (st["next_ids"].clone(), st["finished"].clone(), st["out_ids"].clone())
                    self._decode_step(st, G, temperature, top_k)
                    torch.cuda.synchronize()
                    st["step"].zero_()
                    st["next_ids"].copy_(snap[0]); st["finished"].copy_(snap[1]); st["out_ids"].copy_(snap[2])
                    g = torch.cuda.CUDAGraph()
                    from .lib import load
                    before = load().t1_launch_count(0)
                    if getattr(self, "_hi_stream", None) is None:
                        import os as _os
                        self._hi_stream = torch.cuda.Stream(device=self.dev, priority=0 if _os.environ.get("T1_NO_HI_PRIO") else -1)
                    with torch.cuda.graph(g, stream=self._hi_stream):      # kernel nodes inherit the capture stream's priority
                        self._decode_step(st, G, temperature, top_k)
                    st["launches_per_step"] = int(load().t1_launch_count(0) - before)
                    st["graph"], st["gkey"] = g, gkey
                    st["step"].zero_()
                    st["next_ids"].copy_(snap[0]); st["finished"].copy_(snap[1]); st["out_ids"].copy_(snap[2])
                done = 0
                main, hi = torch.cuda.current_stream(), self._hi_stream
                while done < n_steps:
                    burst = min(check_every, n_steps - done)
                    hi.wait_stream(main)
                    with torch.cuda.stream(hi):
                        for i in range(burst):
                            if self.replay_hook is not None:
                                self.replay_hook(done + i, n_steps)     # profiling windows (bench.py --ncu-range)
                            st["graph"].replay()
                    main.wait_stream(hi)
                    self.metrics_launches += burst * st["launches_per_step"]
                    done += burst
                    if done < n_steps and bool(st["finished"].all().item()):
                        break
            else:
                for i in range(n_steps):
                    self._decode_step(st, G, temperature, top_k)
                    if (i + 1) % check_every == 0 and bool(st["finished"].all().item()):
                        break
        self._last_decode_state = st
        return [st["out_ids"][gi * G:(gi + 1) * G].clone() for gi in range(len(prompts))]

    # ------------------------------------------------------------------------------------------
    # scoring (teacher-forced log-probs of the completions) over the packed layout
    # ------------------------------------------------------------------------------------------
    def _score_inputs(self, prompt: Prompt, completion_ids: torch.Tensor):
        cfg = self.cfg
        ids = np.asarray(prompt.input_ids)
        P = len(ids)
        G, C = completion_ids.shape
        Tp, lo, hi, logit_rows = indexing.packed_rows(P, G, C)
        spg = None if self.run.ref_quirk_mrope else prompt.second_per_grid_ts
        pos3, delta = indexing.mrope_positions(cfg, ids, prompt.grid_thw, spg)
        pos = np.zeros((3, Tp), np.int64)
        pos[:, :P] = pos3
        pos[:, P:P + G * C] = np.tile(P + delta + np.arange(C, dtype=np.int64), G)[None, :]
        ids_packed = torch.zeros(Tp, device=self.dev, dtype=I32)
        ids_packed[:P] = self._i32(ids)
        ids_packed[P:P + G * C] = completion_ids.reshape(-1)
        is_vis = ids == cfg.video_token_id
        vision_rows = np.nonzero(is_vis)[0]
        # embedding-gradient scatter index: token id for text rows, -1 for vision rows and padding
        text_idx = torch.full((Tp,), -1, device=self.dev, dtype=I32)
        text_idx[:P + G * C] = ids_packed[:P + G * C]
        if len(vision_rows):
            text_idx[self._i32(vision_rows).long()] = -1
        R = G * C
        Rp = (R + 63) // 64 * 64
        rows = np.zeros(Rp, np.int32)
        rows[:R] = logit_rows
        targets = torch.zeros(Rp, device=self.dev, dtype=I32)
        targets[:R] = completion_ids.reshape(-1)
        return dict(P=P, G=G, C=C, Tp=Tp, lo=self._i32(lo), hi=self._i32(hi), pos=pos, ids_packed=ids_packed,
                    vision_rows=self._i32(vision_rows), text_idx=text_idx, logit_rows=self._i32(rows), targets=targets,
                    R=R, Rp=Rp)

    def _head_forward(self, hsel: torch.Tensor, targets: torch.Tensor, wset: str):
        if getattr(self.run, "fused_head", True):
            # lm_head + log-softmax + gather + entropy in the GEMM epilogue (timer1_trainer.py:452-481): no logits in HBM
            return ops.head_logprob_entropy(hsel, self.lm_head(wset), targets)
        Rp = hsel.shape[0]
        lp = torch.empty(Rp, device=self.dev, dtype=F32)
        ent = torch.empty(Rp, device=self.dev, dtype=F32)
        lse = torch.empty(Rp, device=self.dev, dtype=F32)
        W = self.lm_head(wset)
        step = self.run.logit_chunk_rows
        for a in range(0, Rp, step):
            b = min(Rp, a + step)
            logits = ops.gemm(hsel[a:b], W, out_dtype=F32)
            call("t1_logprob_entropy", ptr(logits), logits.stride(0), b - a, logits.shape[1], ptr(targets[a:b]),
                 ptr(lp[a:b]), ptr(ent[a:b]), ptr(lse[a:b]), stream_ptr())
        return lp, ent, lse

    def score(self, prompt: Prompt, completion_ids: torch.Tensor, vision_embeds: torch.Tensor, wset: str = "policy",
              save: bool = False, x_vit=None, plan=None, prefill=None, vit_keep=None):
        """Returns (logp [G, C], entropy [G, C], state-or-None)."""
        si = self._score_inputs(prompt, completion_ids)
        G, C, R = si["G"], si["C"], si["R"]
        cos, sin = self._rope_tables(si["pos"])
        P = si["P"]
        if prefill is not None and prefill.get("ref"):
            # reference model: the prompt rows were computed on the side stream while the rollout decoded
            assert wset == "ref" and not save and prefill["P"] == P
            torch.cuda.current_stream().wait_event(prefill["event"])
            xbuf = torch.empty((si["Tp"], self.cfg.hidden_size), device=self.dev, dtype=BF16)
            ckpt = [xbuf] * (self.cfg.num_hidden_layers + 1)          # no backward through the reference: one layer-input buffer
            ops.gather_rows(self.p("model.embed_tokens.weight", "ref"), si["ids_packed"][P:], out=xbuf[P:])
            hN_c = self.llm_forward_suffix(ckpt, P, cos, sin, si["lo"], si["hi"], prefill["kv_prefix"], wset="ref")
            hN = torch.empty((si["Tp"], self.cfg.hidden_size), device=self.dev, dtype=BF16)
            hN[P:] = hN_c
            hN[P - 1:P] = prefill["h_last"]
            ckpt = None
            keep_full = None
        elif prefill is not None:
            # policy == rollout weights: the prompt rows were already computed by the rollout's prefill
            assert wset == "policy" and save
            ckpt = prefill["ckpt_full"]
            assert ckpt[0].shape[0] == si["Tp"]
            ops.gather_rows(self.p("model.embed_tokens.weight"), si["ids_packed"][P:], out=ckpt[0][P:])
            keep_full = prefill.get("keep")
            hN_c = self.llm_forward_suffix(ckpt, P, cos, sin, si["lo"], si["hi"], prefill["kv_prefix"], keep=keep_full)
            hN = torch.empty((si["Tp"], self.cfg.hidden_size), device=self.dev, dtype=BF16)
            hN[P:] = hN_c
            hN[P - 1:P] = prefill["h_last"]
        else:
            x0 = self._embed(si["ids_packed"], vision_embeds, si["vision_rows"])
            ckpt = [None] * (self.cfg.num_hidden_layers + 1) if save else None
            keep_full = self._alloc_keep(self._save_level(1, si["Tp"]), si["Tp"]) if save else None
            hN = self.llm_forward(x0, cos, sin, si["lo"], si["hi"], P, wset=wset, ckpt=ckpt, keep=keep_full)
        hsel = ops.gather_rows(hN, si["logit_rows"])
        lp, ent, lse = self._head_forward(hsel, si["targets"], wset)
        state = None
        if save:
            state = ScoreState(P=si["P"], G=G, C=C, Tp=si["Tp"], ckpt=ckpt, cos=cos, sin=sin, lo=si["lo"], hi=si["hi"],
                               ids_packed=si["ids_packed"], text_scatter_idx=si["text_idx"], vision_rows=si["vision_rows"],
                               logit_rows=si["logit_rows"], targets=si["targets"], hsel=hsel, lse=lse, x_vit=x_vit, plan=plan, vit_keep=vit_keep,
                               **({("gu" if k == "gu" else "k_" + k): v for k, v in keep_full.items()} if keep_full else {}))
        return lp[:R].view(G, C), ent[:R].view(G, C), state

    # ------------------------------------------------------------------------------------------
    # backward
    # ------------------------------------------------------------------------------------------
    def _gview(self, name: str) -> torch.Tensor:
        return self.grads[name]

    def _wgrad(self, dy: torch.Tensor, x: torch.Tensor, name: str, acc: bool):
        """grad[name] (+)= dy^T x   (dy [T, N], x [T, K] -> [N, K]); both operands consumed in place (MN-major)."""
        ops.gemm(dy, x, out=self._gview(name), accumulate=acc, a_t=True, b_t=True)

    def _dgrad(self, dy: torch.Tensor, w: torch.Tensor, out=None) -> torch.Tensor:
        """dx = dy @ w   (dy [T, N], w [N, K] consumed in place as an MN-major B operand)."""
        return ops.gemm(dy, w, out=out, b_t=True)

    def _small_grad(self, f32: torch.Tensor, name: str, acc: bool):
        ops.f32_to_bf16(f32, self._gview(name).view(-1), accumulate=acc)

    def backward(self, st: ScoreState, dlogp: torch.Tensor, overlap_reduce: bool = False):
        """Accumulate d loss / d params into self.grads given d loss / d logp [G, C] (fp32).
        overlap_reduce: this is the LAST backward before optimizer_step — reduce-scatter each gradient bucket as soon as it
        is final (no-op at world size 1 or with T1_ZERO_OVERLAP=0)."""
        cfg = self.cfg
        self._grad_version += 1
        self._rs = None
        if overlap_reduce and self.zero_overlap:
            self.plan.reset()
            self._rs = dict(version=self._grad_version, works=[], capped=False)
        try:
            self._backward(st, dlogp)
        finally:
            if self._rs is not None and self._rs.get("capped"):
                load_lib().t1_gemm_max_ctas(0)

    def _backward(self, st: ScoreState, dlogp: torch.Tensor):
        cfg = self.cfg
        off = lambda name: self.grads.layout[name][0]
        train_vit = not self.run.fix_vit
        if train_vit and st.vit_keep is None and st.vision_rows.numel() > 0:
            raise RuntimeError("fix_vit=False: the scoring state carries no ViT checkpoints (vit_blocks(keep=...) -> score(vit_keep=...))")
        acc = not self.grad_fresh
        L, Hq, Hkv, hd, I, D = cfg.num_hidden_layers, cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim, cfg.intermediate_size, cfg.hidden_size
        qd, kd = Hq * hd, Hkv * hd
        eps, scale = cfg.rms_norm_eps, hd ** -0.5
        T = st.Tp
        R, Rp = st.G * st.C, st.hsel.shape[0]
        dev = self.dev
        embed_name = "model.embed_tokens.weight"
        head_name = embed_name if cfg.tie_word_embeddings else "lm_head.weight"
        if not acc and not cfg.tie_word_embeddings:
            self._gview(embed_name).zero_()     # only rows of seen tokens are written (scatter-add) below

        # ---- lm_head + log-softmax backward, chunked over logit rows ----
        dl = torch.zeros(Rp, device=dev, dtype=F32)
        dl[:R] = dlogp.reshape(-1)
        W = self.lm_head()
        dhsel = torch.empty((Rp, D), device=dev, dtype=BF16)
        step = self.run.logit_chunk_rows
        first = True
        for a in range(0, Rp, step):
            b = min(Rp, a + step)
            if getattr(self.run, "fused_head", True):   # recompute (1.7 TFLOP at cfg2) with the softmax backward in the epilogue
                dlog = ops.head_dlogits(st.hsel[a:b], W, st.targets[a:b], st.lse[a:b], dl[a:b])
            else:
                logits = ops.gemm(st.hsel[a:b], W, out_dtype=F32)
                dlog = ops.logits_backward(logits, st.targets[a:b], st.lse[a:b], dl[a:b])
                del logits
            ops.gemm(dlog, W, out=dhsel[a:b], b_t=True)
            ops.gemm(dlog, st.hsel[a:b], out=self._gview(head_name), accumulate=acc or not first, a_t=True, b_t=True)
            first = False
            del dlog
        if not cfg.tie_word_embeddings:
            self._reduce_ready(off(head_name))                              # lm_head is the last tensor of the flat buffer
        dh = torch.zeros((T, D), device=dev, dtype=BF16)
        ops.scatter_rows(dhsel, st.logit_rows, dh, accumulate=True)         # prompt row P-1 receives G contributions
        # ---- final norm ----
        f32s = torch.zeros(max(cfg.qkv_dim, 2 * D), device=dev, dtype=F32)
        dx = ops.rmsnorm_bwd(dh, st.ckpt[L], self.p("model.norm.weight"), eps, f32s[:D])
        self._small_grad(f32s[:D], "model.norm.weight", acc)
        if not cfg.tie_word_embeddings:
            self._reduce_ready(off("model.norm.weight"))
        del dh
        # ---- layers, last to first, recomputing each from its checkpointed input ----
        h = torch.empty((T, D), device=dev, dtype=BF16)
        qkv = torch.empty((T, cfg.qkv_dim), device=dev, dtype=BF16)
        attn = torch.empty((T, qd), device=dev, dtype=BF16)
        o = torch.empty((T, D), device=dev, dtype=BF16)
        xm = torch.empty((T, D), device=dev, dtype=BF16)
        h2 = torch.empty((T, D), device=dev, dtype=BF16)
        gu = torch.empty((T, 2 * I), device=dev, dtype=BF16) if st.gu is None else None
        act = torch.empty((T, I), device=dev, dtype=BF16)
        dqkv = torch.empty((T, cfg.qkv_dim), device=dev, dtype=BF16)
        lse2 = torch.empty((Hq, T), device=dev, dtype=F32)
        bwd_tc = hd == 128                    # tcgen05/TMEM attention backward; HMMA kernel for other head dims
        bwd_ws = torch.empty(int(load_lib().t1_attn_bwd_tc_workspace_bytes(T, T, Hq)), device=dev, dtype=torch.uint8) if bwd_tc else None
        for l in range(L - 1, -1, -1):
            pre = f"model.layers.{l}."
            x_in = st.ckpt[l]
            Wqkv, Wo = self.p(pre + "self_attn.qkv.weight"), self.p(pre + "self_attn.o_proj.weight")
            Wgu, Wd = self.p(pre + "mlp.gate_up.weight"), self.p(pre + "mlp.down_proj.weight")
            # recompute (whatever the forward passes did not keep)
            ops.rmsnorm(x_in, self.p(pre + "input_layernorm.weight"), eps, out=h)
            if st.k_qkv is not None:
                qkv, attn, lse2 = st.k_qkv[l], st.k_attn[l], st.k_lse[l]
                st.k_qkv[l] = st.k_attn[l] = st.k_lse[l] = None
                q, k, v = qkv[:, :qd], qkv[:, qd:qd + kd], qkv[:, qd + kd:]
            else:
                ops.gemm(h, Wqkv, out=qkv, bias=self.p(pre + "self_attn.qkv.bias"))
                ops.rope_apply(qkv, 0, Hq + Hkv, hd, st.cos, st.sin)
                q, k, v = qkv[:, :qd], qkv[:, qd:qd + kd], qkv[:, qd + kd:]
                self._attn(q, k, v, st.lo, st.hi, st.P, Hq, Hkv, hd, scale, out=attn, lse=lse2)
            ops.gemm(attn, Wo, out=o)
            ops.rmsnorm(x_in, self.p(pre + "post_attention_layernorm.weight"), eps, delta=o, out=h2, x_out=xm)
            if st.gu is not None:            # projection kept from the forward pass: only the activation is redone
                gu = st.gu[l]
                st.gu[l] = None
                ops.swiglu(gu, out=act)
            else:
                ops.gemm_swiglu(h2, Wgu, out=act, gate_up_out=gu)      # the backward needs the projection as well
            # MLP backward
            self._wgrad(dx, act, pre + "mlp.down_proj.weight", acc)
            dact = self._dgrad(dx, Wd, out=act)                             # act no longer needed: reuse its storage
            dgu = ops.swiglu_bwd(dact, gu, out=gu)                          # in place over gate_up
            self._wgrad(dgu, h2, pre + "mlp.gate_up.weight", acc)
            dh2 = self._dgrad(dgu, Wgu, out=h2)
            f32s.zero_()
            dxm = ops.rmsnorm_bwd(dh2, xm, self.p(pre + "post_attention_layernorm.weight"), eps, f32s[:D], dres=dx, out=o)
            self._small_grad(f32s[:D], pre + "post_attention_layernorm.weight", acc)
            # attention backward (dxm is also d o_proj output)
            self._wgrad(dxm, attn, pre + "self_attn.o_proj.weight", acc)
            dattn = self._dgrad(dxm, Wo)
            if bwd_tc:
                ops.attn_bwd_tc(q, k, v, attn, dattn, lse2, st.lo, st.hi, st.P, Hq, Hkv, hd, scale,
                                dqkv[:, :qd], dqkv[:, qd:qd + kd], dqkv[:, qd + kd:], ws=bwd_ws)
            else:
                ops.attn_bwd(q, k, v, attn, dattn, lse2, st.lo, st.hi, st.P, Hq, Hkv, hd, scale,
                             dqkv[:, :qd], dqkv[:, qd:qd + kd], dqkv[:, qd + kd:])
            del dattn
            ops.rope_apply(dqkv, 0, Hq + Hkv, hd, st.cos, st.sin, inverse=True)
            f32s.zero_()
            ops.colsum(dqkv, f32s[:cfg.qkv_dim])
            self._small_grad(f32s[:cfg.qkv_dim], pre + "self_attn.qkv.bias", acc)
            self._wgrad(dqkv, h, pre + "self_attn.qkv.weight", acc)
            dhh = self._dgrad(dqkv, Wqkv, out=h)
            f32s.zero_()
            dx = ops.rmsnorm_bwd(dhh, x_in, self.p(pre + "input_layernorm.weight"), eps, f32s[:D], dres=dxm, out=xm)
            self._small_grad(f32s[:D], pre + "input_layernorm.weight", acc)
            if not cfg.tie_word_embeddings:
                self._reduce_ready(off(pre + "input_layernorm.weight"))    # layer l is complete: everything above it is final
            dx = dx.clone() if l > 0 else dx                                 # xm is overwritten by the next recompute
        # ---- embeddings: text rows scatter-add into the table gradient; vision rows flow into the merger ----
        ops.scatter_rows(dx, st.text_scatter_idx, self._gview(embed_name), accumulate=True)
        if st.vision_rows.numel() > 0:
            dve = ops.gather_rows(dx, st.vision_rows)                       # [Nv, D] in original merged-token order
            dx_vit = self._merger_backward(st, dve, acc, need_dx=train_vit)
            if train_vit:
                if not cfg.tie_word_embeddings:
                    self._reduce_ready(off("visual.merger.ln_q.weight"))
                self._vit_backward(st, dx_vit, acc)
        elif not acc:
            self.grads.buf[:off(embed_name)].zero_()                        # no vision tokens: nothing wrote the tower's gradients
        # the buckets that only became final here (embedding table, merger; with fix_vit=False the first ViT blocks) are reduced
        # by optimizer_step on the full-speed communicator: nothing is left to hide them behind
        self.grad_fresh = False

    def _merger_backward(self, st: ScoreState, dve: torch.Tensor, acc: bool, need_dx: bool = False):
        """PatchMerger backward; returns d loss / d (final ViT stream) [N, vit_hidden] in window order when need_dx."""
        cfg = self.cfg
        plan, x_vit = st.plan, st.x_vit
        keep: dict = {}
        N, vd, unit = plan.n_patches, cfg.vit_hidden, cfg.merge_unit
        Nv = N // unit
        # recompute the merger forward from the saved ViT output
        self.merger(x_vit, None, plan, "policy", keep=keep)
        # Qwen2.5-VL: back to window order (Q25:512-513); Qwen2-VL has no post-merger permutation (Q2:771-790)
        dm = ops.gather_rows(dve, self._i32(plan.window_index)) if cfg.is_q25 else dve
        W0, W2 = self.p("visual.merger.mlp.0.weight"), self.p("visual.merger.mlp.2.weight")
        f32s = torch.zeros(max(unit * vd, cfg.hidden_size), device=self.dev, dtype=F32)
        self._wgrad(dm, keep["g1"], "visual.merger.mlp.2.weight", acc)
        ops.colsum(dm, f32s[:cfg.hidden_size])
        self._small_grad(f32s[:cfg.hidden_size], "visual.merger.mlp.2.bias", acc)
        dg1 = self._dgrad(dm, W2)
        da1 = ops.act_bwd(dg1, keep["a1"], 0)
        self._wgrad(da1, keep["y4"], "visual.merger.mlp.0.weight", acc)
        f32s.zero_()
        ops.colsum(da1, f32s[:unit * vd])
        self._small_grad(f32s[:unit * vd], "visual.merger.mlp.0.bias", acc)
        dy4 = self._dgrad(da1, W0)
        f32s.zero_()
        if cfg.is_q25:
            dxv = ops.rmsnorm_bwd(dy4.view(N, vd), x_vit, self.p("visual.merger.ln_q.weight"), 1e-6, f32s[:vd])
        else:      # Qwen2-VL: LayerNorm(eps 1e-6) with bias (HF modeling_qwen2_vl.py:313-326); frozen ViT blocks need no dx
            dxv = ops.layernorm_bwd(dy4.view(N, vd), x_vit, self.p("visual.merger.ln_q.weight"), 1e-6, f32s[:vd], f32s[vd:2 * vd],
                                    need_dx=need_dx)
            self._small_grad(f32s[vd:2 * vd], "visual.merger.ln_q.bias", acc)
        self._small_grad(f32s[:vd], "visual.merger.ln_q.weight", acc)
        return dxv if need_dx else None

    def _vit_backward(self, st: ScoreState, dx: torch.Tensor, acc: bool):
        """ViT blocks + patch-embed backward (fix_vit=False, the reference's default: main.py:46-49, timer1_trainer.py:272-280
        leaves model.visual trainable).  dx [N, vit_hidden]: gradient of the final ViT stream, window order.  Every block is
        recomputed from its checkpointed input (gradient checkpointing, train_rl.sh:32) and differentiated in place:
        HF Q25:281-300 (block), Q25:244-261 (attention, non-causal inside [lo, hi)), Q25:108-121 (SwiGLU MLP) /
        Q2:  QuickGELU MLP with LayerNorm."""
        cfg, kp, plan = self.cfg, st.vit_keep, st.plan
        q25 = cfg.is_q25
        N, vd, H, hd = plan.n_patches, cfg.vit_hidden, cfg.vit_heads, cfg.vit_head_dim
        iw = cfg.vit_inter_padded if q25 else cfg.vit_intermediate
        scale = hd ** -0.5
        dev = self.dev
        off = lambda name: self.grads.layout[name][0]
        cos, sin = kp["cos"], kp["sin"]
        e = lambda *shape: torch.empty(shape, device=dev, dtype=BF16)
        h, qkv, attn, proj, xm, h2 = e(N, vd), e(N, 3 * vd), e(N, vd), e(N, vd), e(N, vd), e(N, vd)
        pre_act = e(N, 2 * iw) if q25 else e(N, iw)          # gate_up projection (Q2.5) / fc1 output (Q2)
        act = e(N, iw)
        dqkv = e(N, 3 * vd)
        lse = torch.empty((H, N), device=dev, dtype=F32)
        f32s = torch.zeros(max(3 * vd, 2 * iw), device=dev, dtype=F32)
        f32b = torch.zeros(vd, device=dev, dtype=F32)

        def bias_grad(dy, name):
            f32s.zero_()
            n = dy.shape[1]
            ops.colsum(dy, f32s[:n])
            self._small_grad(f32s[:n], name, acc)

        def norm_fwd(x, name, delta=None, out=None, x_out=None):
            if q25:
                return ops.rmsnorm(x, self.p(name + ".weight"), 1e-6, delta=delta, out=out, x_out=x_out)
            return ops.layernorm(x, self.p(name + ".weight"), self.p(name + ".bias"), 1e-6, delta=delta, out=out, x_out=x_out)

        def norm_bwd(dy, x, name, dres, out):
            """out = dres + d norm(x) / dx ; accumulates the norm's weight (and bias) gradient."""
            f32s.zero_()
            if q25:
                r = ops.rmsnorm_bwd(dy, x, self.p(name + ".weight"), 1e-6, f32s[:vd], dres=dres, out=out)
            else:
                f32b.zero_()
                r = ops.layernorm_bwd(dy, x, self.p(name + ".weight"), 1e-6, f32s[:vd], f32b, out=out)
                ops.add_(r, dres)
                self._small_grad(f32b, name + ".bias", acc)
            self._small_grad(f32s[:vd], name + ".weight", acc)
            return r

        for i in range(cfg.vit_depth - 1, -1, -1):
            pre = f"visual.blocks.{i}."
            x_in = kp["ckpt"][i]
            kp["ckpt"][i] = None
            w_up, w_dn = ("mlp.gate_up", "mlp.down_proj") if q25 else ("mlp.fc1", "mlp.fc2")
            Wqkv, Wp = self.p(pre + "attn.qkv.weight"), self.p(pre + "attn.proj.weight")
            Wu, Wd = self.p(pre + w_up + ".weight"), self.p(pre + w_dn + ".weight")
            # ---- recompute the block from its input ----
            norm_fwd(x_in, pre + "norm1", out=h)
            ops.gemm(h, Wqkv, out=qkv, bias=self.p(pre + "attn.qkv.bias"))
            ops.rope_apply(qkv, 0, 2 * H, hd, cos, sin)
            full = (not q25) or (i in cfg.fullatt_block_indexes)
            lo, hi = kp["full"] if full else kp["win"]
            q, k, v = qkv[:, :vd], qkv[:, vd:2 * vd], qkv[:, 2 * vd:]
            ops.attn_fwd(q, k, v, lo, hi, 0, H, H, hd, scale, out=attn, lse=lse)
            ops.gemm(attn, Wp, out=proj, bias=self.p(pre + "attn.proj.bias"))
            norm_fwd(x_in, pre + "norm2", delta=proj, out=h2, x_out=xm)
            if q25:
                ops.gemm_swiglu(h2, Wu, bias=self.p(pre + w_up + ".bias"), out=act, gate_up_out=pre_act)
            else:
                ops.gemm(h2, Wu, out=pre_act, bias=self.p(pre + w_up + ".bias"))
                ops.act(pre_act, 1, out=act)
            # ---- MLP backward ----
            self._wgrad(dx, act, pre + w_dn + ".weight", acc)
            bias_grad(dx, pre + w_dn + ".bias")
            dact = self._dgrad(dx, Wd, out=act)
            dpre = ops.swiglu_bwd(dact, pre_act, out=pre_act) if q25 else ops.act_bwd(dact, pre_act, 1, out=pre_act)
            self._wgrad(dpre, h2, pre + w_up + ".weight", acc)
            bias_grad(dpre, pre + w_up + ".bias")
            dh2 = self._dgrad(dpre, Wu, out=h2)
            dxm = norm_bwd(dh2, xm, pre + "norm2", dres=dx, out=proj)
            # ---- attention backward (dxm is also the gradient of the projection output) ----
            self._wgrad(dxm, attn, pre + "attn.proj.weight", acc)
            bias_grad(dxm, pre + "attn.proj.bias")
            dattn = self._dgrad(dxm, Wp)
            ops.attn_bwd(q, k, v, attn, dattn, lse, lo, hi, 0, H, H, hd, scale, dqkv[:, :vd], dqkv[:, vd:2 * vd], dqkv[:, 2 * vd:])
            del dattn
            ops.rope_apply(dqkv, 0, 2 * H, hd, cos, sin, inverse=True)
            bias_grad(dqkv, pre + "attn.qkv.bias")
            self._wgrad(dqkv, h, pre + "attn.qkv.weight", acc)
            dh = self._dgrad(dqkv, Wqkv, out=h)
            dx = norm_bwd(dh, x_in, pre + "norm1", dres=dxm, out=xm).clone()     # xm is overwritten by the next recompute
            if not cfg.tie_word_embeddings:
                self._reduce_ready(off(pre + "norm1.weight"))                # block i is complete: everything above it is final
        # ---- patch-embed (Conv3d == GEMM, no bias): dW = dx^T pixels ----
        self._wgrad(dx, kp["pix"], "visual.patch_embed.proj.weight", acc)

    # ------------------------------------------------------------------------------------------
    # optimizer: global-norm clip + AdamW on this rank's shard of the flat buffer (ZeRO-2 semantics)
    # ------------------------------------------------------------------------------------------
    def optimizer_step(self, lr: float | None = None, grad_scale: float = 1.0, lazy_gather: bool | None = None) -> torch.Tensor:
        """Returns the (pre-clip) global gradient norm as a 1-element device tensor.
        lazy_gather (default: Engine.lazy_gather): do not wait for the all-gathers here; every later read of a trained
        parameter through Engine.p() waits for just the buckets it touches (sync_params() waits for all).
        world > 1: bucketed reduce-scatter (those issued during the last backward are only waited for) -> grad norm ->
        per bucket AdamW on this rank's slice + asynchronous IN-PLACE all-gather of the bucket (bucket b's gather runs under
        bucket b+1's AdamW); the main stream waits for the gathers before it returns to the caller's work."""
        run = self.run
        lr = run.learning_rate if lr is None else lr
        g = self.grads.buf
        plan = self.plan
        if self.world > 1:
            import torch.distributed as dist

            rs = self._rs
            if rs is None or rs["version"] != self._grad_version:      # nothing (valid) in flight: reduce everything now
                plan.reset()
                rs = dict(works=[])
            for b in plan.ready_buckets(0):
                rs["works"].append(plan.reduce_scatter(b, g, self.shard_g, self.rank, group=self.pg, async_op=True))
            for w in rs["works"]:
                if w is not None:
                    w.wait()                                            # stream-level: orders the main stream after the collective
            self._rs = None
            shard_g = self.shard_g
        else:
            shard_g = g
        gsq = torch.zeros(1, device=self.dev, dtype=F32)
        ops.sumsq(shard_g, gsq)
        if self.world > 1:
            dist.all_reduce(gsq, group=self.pg)
        self.opt_step += 1
        self.sync_params()                 # gathers of an earlier step (none in the normal flow: the next step consumed them)
        lazy = self.lazy_gather if lazy_gather is None else (lazy_gather and self.world > 1)
        works = []
        for b in range(len(plan)):
            lo, hi = plan.slice_of(b, self.rank)
            a, z = plan.shard_range(b)
            ops.adamw(self.master[a:z], self.adam_m[a:z], self.adam_v[a:z], shard_g[a:z], self.w.train.buf[lo:hi], lr, run.adam_beta1,
                      run.adam_beta2, run.adam_eps, run.weight_decay, self.opt_step, gnorm_sq=gsq, max_norm=run.max_grad_norm,
                      grad_scale=grad_scale)
            if self.world > 1:
                works.append(plan.all_gather(b, self.w.train.buf, self.rank, group=self.pg_ag if lazy else self.pg, async_op=True))
        if lazy:
            self._ag_pending = {b: w for b, w in enumerate(works)}
            if self.ag_ctas > 0:           # leave the gather its SMs while the next step's first GEMMs run beside it
                sms = torch.cuda.get_device_properties(self.dev).multi_processor_count
                load_lib().t1_gemm_max_ctas(max(1, sms - self.ag_ctas))
                self._ag_capped = True
        else:
            for w in works:
                if w is not None:
                    w.wait()
        self.grad_fresh = True
        self._invalidate_after_update()
        return gsq.sqrt() * grad_scale

    def _invalidate_after_update(self):
        pass  # decode graphs read weights through pointers into the flat buffer, which is updated in place

    # ------------------------------------------------------------------------------------------
    # one complete GRPO micro-step for one prompt group
    # ------------------------------------------------------------------------------------------
    def grpo_step(self, prompt: Prompt, reward_fn, seed: int | None = None, completion_override=None,
                  loss_scale: float = 1.0, timers: dict | None = None, vit_prefetch: dict | None = None,
                  last_before_optimizer: bool = True):
        """One prompt-group: see grpo_multi_step."""
        ov = None if completion_override is None else [completion_override]
        return self.grpo_multi_step([prompt], [reward_fn], seed=seed, completion_overrides=ov, loss_scale=loss_scale,
                                    timers=timers, vit_prefetch=vit_prefetch, last_before_optimizer=last_before_optimizer)[0]

    def grpo_multi_step(self, prompts: list, reward_fns: list, seed: int | None = None, completion_overrides=None,
                        loss_scale: float = 1.0, timers: dict | None = None, vit_prefetch: dict | None = None,
                        last_before_optimizer: bool = True) -> list:
        """The micro-batches of ONE optimizer step (gradient accumulation, train_rl.sh:27), each a prompt-group:
        ViT per group -> joint rollout (shared weight stream) -> per group: mask, policy/ref scoring, rewards (host
        callback), advantage, loss, backward (gradients accumulate).
        reward_fn(completion_ids_host [G, C] int64 numpy) -> rewards_per_func [G, n_funcs] float32 numpy; a reward_fn
        carrying a `.device` attribute (callable: completion ids [G, C] int32 device tensor -> [G, n_funcs] fp32 device
        tensor, e.g. rewards.DeviceRewards) is evaluated on the device instead and the completions are not copied out.
        last_before_optimizer: optimizer_step follows this call (the gradient buckets of the last group's backward are
        reduce-scattered while that backward still runs; a later backward simply invalidates them).
        Returns one dict of device tensors per group."""
        run, cfg = self.run, self.cfg
        G, C = run.num_generations, run.max_completion_length
        seed = run.seed if seed is None else seed

        def mark(name):
            if timers is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                timers.setdefault("_events", []).append((name, ev))

        mark("start")
        vit = []
        vkeeps = []
        if vit_prefetch is not None:
            assert vit_prefetch["n"] == len(prompts)
            torch.cuda.current_stream().wait_event(vit_prefetch["event"])
        for pi, prompt in enumerate(prompts):
            vkeep = None if run.fix_vit else {}
            if vit_prefetch is not None:
                assert run.fix_vit, "prefetch_vit runs FROZEN blocks only"
                x_vit, delta, plan = vit_prefetch["results"][pi]
            else:
                x_vit, delta, plan = self.vit_blocks(prompt.pixel_values, prompt.grid_thw, frames=prompt.frames, keep=vkeep)
            emb = self.merger(x_vit, delta, plan, "policy")                 # x_vit now holds the final ViT stream
            vit.append((emb, x_vit, plan))
            vkeeps.append(vkeep)
        mark("vit")
        prefills = [None] * len(prompts)
        ref_handles = [None] * len(prompts)
        overlap_ref = (completion_overrides is None and run.beta != 0.0 and self.ref is not None and run.side_ctas > 0
                       and cfg.head_dim == 128 and self._decode_impl() == "stream")
        def launch_ref_prefills():
            for gi, (prompt, (emb, x_vit, plan)) in enumerate(zip(prompts, vit)):
                ref_handles[gi] = self.ref_prefill_async(prompt, self._ref_embeds(prompt, x_vit, plan), G, C)

        if completion_overrides is None:
            reuse = run.reuse_prefill and cfg.head_dim == 128
            comps = self.rollout_multi(prompts, [v[0] for v in vit], G, C, run.temperature, run.top_k, seed, keep_prefill=reuse,
                                       shared_sms=overlap_ref, before_decode=launch_ref_prefills if overlap_ref else None)
            if reuse:
                for gi, prompt in enumerate(prompts):
                    pf = self._last_decode_state["groups"][gi].pop("prefill", None)
                    # the scoring pass may use different temporal M-RoPE positions than the rollout (reference quirk,
                    # timer1_trainer.py:601-607): reuse only when the prompt positions are identical
                    same_pos = (not run.ref_quirk_mrope) or pf is None or pf["spg"] is None or all(float(v) == 1.0 for v in pf["spg"]) or not cfg.is_q25
                    prefills[gi] = pf if same_pos else None
        else:
            comps = [c.to(device=self.dev, dtype=I32).contiguous() for c in completion_overrides]
        mark("rollout")
        outs = []
        for gi, (prompt, reward_fn, comp, (emb, x_vit, plan)) in enumerate(zip(prompts, reward_fns, comps, vit)):
            mask, lengths = ops.eos_mask(comp, cfg.eos_token_id)
            logp, ent, st = self.score(prompt, comp, emb, "policy", save=True, x_vit=x_vit, plan=plan, prefill=prefills[gi],
                                       vit_keep=vkeeps[gi])
            vkeeps[gi] = None
            prefills[gi] = None
            mark("score_policy")
            ref_logp = None
            if run.beta != 0.0:
                if self.ref is None:
                    raise RuntimeError("beta != 0 needs the reference parameter set (Engine(with_ref=True))")
                if ref_handles[gi] is not None:
                    ref_logp, _, _ = self.score(prompt, comp, None, "ref", prefill=ref_handles[gi])
                    ref_handles[gi] = None
                else:
                    emb_ref = self._ref_embeds(prompt, x_vit, plan)
                    ref_logp, _, _ = self.score(prompt, comp, emb_ref, "ref")
            mark("score_ref")
            on_device = getattr(reward_fn, "device", None)
            if on_device is not None:
                # registry rewards: detokenise + parse + IoU on the device, no host round trip before the loss
                rpf_dev = on_device(comp).to(F32).reshape(G, -1).contiguous()
            else:
                comp_host = comp.cpu().numpy().astype(np.int64)             # D2H: completions for the host reward fns
                rpf = np.ascontiguousarray(reward_fn(comp_host), dtype=np.float32).reshape(G, -1)
                rpf_dev = torch.from_numpy(rpf).to(self.dev)
            rewards, adv, mean, std = ops.group_advantage(rpf_dev, G)
            loss, dlogp, stats = ops.grpo_loss(logp, ref_logp, ent, adv, mask, run.beta, run.epsilon_low, run.epsilon_high,
                                               run.use_grpo, loss_scale=loss_scale)
            mark("reward_loss")
            self.backward(st, dlogp, overlap_reduce=last_before_optimizer and gi == len(prompts) - 1)
            mark("backward")
            del st
            outs.append(dict(loss=loss, stats=stats, rewards=rewards, rewards_per_func=rpf_dev, advantages=adv, reward_std=std,
                             completion_ids=comp, completion_mask=mask, logp=logp, ref_logp=ref_logp, entropy=ent, dlogp=dlogp))
        return outs


def finish_timers(timers: dict) -> dict:
    """Convert the CUDA events recorded by grpo_step(timers=...) into per-phase milliseconds."""
    ev = timers.pop("_events", [])
    torch.cuda.synchronize()
    out = {}
    for (n0, e0), (n1, e1) in zip(ev[:-1], ev[1:]):
        out[n1] = out.get(n1, 0.0) + e0.elapsed_time(e1)
    return out
