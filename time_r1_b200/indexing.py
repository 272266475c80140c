"""Host-side index construction (numpy): ViT window permutation / attention ranges / 2-D RoPE positions and the
LLM's M-RoPE position ids.  These are the integer-only parts of the forward pass; they run once per prompt.

Semantics: HF modeling_qwen2_5_vl.py rot_pos_emb :382-409, get_window_index :411-451, cu_seqlens :488-500, and the
transformers-4.51.1 get_rope_index behaviour the reference was written against.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .config import ModelConfig


@dataclass
class VisionPlan:
    n_patches: int
    perm_rows: np.ndarray      # [N] int32: physical row i of the ViT stream = original patch row perm_rows[i]
    window_index: np.ndarray   # [N/unit] int32 merged-token permutation (identity for Qwen2-VL)
    pos_hw: np.ndarray         # [2, N] int32 (h, w) position of each physical row
    win_lo: np.ndarray         # [N] int32 per-row key range for windowed blocks
    win_hi: np.ndarray
    full_lo: np.ndarray        # [N] int32 per-row key range for full (per temporal slot) blocks
    full_hi: np.ndarray
    win_tile_local: bool = False   # every row's window lies inside the row's own 128-row tile (-> t1_attn_fwd_tc_win)


def _ranges_from_cu(cu: np.ndarray, n: int):
    lo = np.zeros(n, np.int32)
    hi = np.zeros(n, np.int32)
    for a, b in zip(cu[:-1], cu[1:]):
        lo[a:b] = a
        hi[a:b] = b
    return lo, hi


def vision_plan(cfg: ModelConfig, grid_thw) -> VisionPlan:
    ms, unit = cfg.spatial_merge_size, cfg.merge_unit
    pos_list, widx_list, cu_win, cu_full = [], [], [0], [0]
    base = 0
    for t, h, w in grid_thw:
        t, h, w = int(t), int(h), int(w)
        assert h % ms == 0 and w % ms == 0, "grid h/w must be multiples of the spatial merge size"
        hp = np.broadcast_to(np.arange(h)[:, None], (h, w)).reshape(h // ms, ms, w // ms, ms).transpose(0, 2, 1, 3).reshape(-1)
        wp = np.broadcast_to(np.arange(w)[None, :], (h, w)).reshape(h // ms, ms, w // ms, ms).transpose(0, 2, 1, 3).reshape(-1)
        pos_list.append(np.tile(np.stack([hp, wp], 0), (1, t)))
        lh, lw = h // ms, w // ms
        if cfg.is_q25:
            vws = cfg.window_size // ms // cfg.patch_size
            index = np.arange(t * lh * lw).reshape(t, lh, lw)
            pad_h, pad_w = vws - lh % vws, vws - lw % vws   # pads a whole empty window when divisible (HF quirk)
            nh, nw = (lh + pad_h) // vws, (lw + pad_w) // vws
            ip = np.pad(index, ((0, 0), (0, pad_h), (0, pad_w)), constant_values=-100)
            ip = ip.reshape(t, nh, vws, nw, vws).transpose(0, 1, 3, 2, 4).reshape(t, nh * nw, vws, vws)
            seqlens = (ip != -100).sum((2, 3)).reshape(-1)
            flat = ip.reshape(-1)
            widx_list.append(flat[flat != -100] + base)
            for sl in seqlens:
                if sl > 0:  # torch.unique_consecutive drops the zero-length windows
                    cu_win.append(cu_win[-1] + int(sl) * unit)
        else:
            widx_list.append(np.arange(t * lh * lw) + base)
        for _ in range(t):
            cu_full.append(cu_full[-1] + h * w)
        base += t * lh * lw
    pos = np.concatenate(pos_list, 1).astype(np.int32)
    widx = np.concatenate(widx_list).astype(np.int32)
    n = pos.shape[1]
    perm = (widx[:, None].astype(np.int64) * unit + np.arange(unit)[None, :]).reshape(-1).astype(np.int32)
    pos = pos[:, perm]
    full_lo, full_hi = _ranges_from_cu(np.asarray(cu_full), n)
    if cfg.is_q25:
        win_lo, win_hi = _ranges_from_cu(np.asarray(cu_win), n)
    else:
        win_lo, win_hi = full_lo, full_hi
    rows = np.arange(n)
    local = bool(n > 0 and np.all(win_lo // 128 == rows // 128) and np.all((np.maximum(win_hi, 1) - 1) // 128 == rows // 128))
    return VisionPlan(n, perm, widx, np.ascontiguousarray(pos), win_lo, win_hi, full_lo, full_hi, local)


def mrope_positions(cfg: ModelConfig, input_ids, grid_thw, second_per_grid_ts=None):
    """3-axis position ids [3, L] (int32) + rope_delta for ONE un-padded sequence, transformers-4.51.1 semantics:
    text run -> all axes arange from the running start; vision block -> (t*spg*tokens_per_second, h, w) over the
    merged grid, offset by the running start; next run starts at max+1."""
    ids = np.asarray(input_ids)
    ms = cfg.spatial_merge_size
    is_vis = (ids == cfg.video_token_id) | (ids == cfg.image_token_id)
    L = len(ids)
    out = np.zeros((3, L), np.int64)
    i, vid, start = 0, 0, 0
    while i < L:
        if is_vis[i]:
            t, h, w = (int(v) for v in grid_thw[vid])
            lh, lw = h // ms, w // ms
            n = t * lh * lw
            tt = np.repeat(np.arange(t), lh * lw)
            if cfg.is_q25:
                spg = 1.0 if second_per_grid_ts is None else float(second_per_grid_ts[vid])
                tt = (tt.astype(np.float32) * np.float32(spg) * np.float32(cfg.tokens_per_second)).astype(np.int64)
            hh = np.tile(np.repeat(np.arange(lh), lw), t)
            ww = np.tile(np.arange(lw), t * lh)
            out[:, i:i + n] = np.stack([tt, hh, ww]) + start
            start = int(out[:, i:i + n].max()) + 1
            i += n
            vid += 1
        else:
            j = i
            while j < L and not is_vis[j]:
                j += 1
            out[:, i:j] = np.arange(j - i)[None, :] + start
            start = int(out[:, i:j].max()) + 1
            i = j
    delta = int(out.max()) + 1 - L
    return out.astype(np.int32), delta


def mrope_channel_axes(cfg: ModelConfig) -> np.ndarray:
    """axis (0=t, 1=h, 2=w) used by rotary channel i of the half head-dim (sections [16,24,24] -> t,h,w)."""
    ax = []
    for a, n in enumerate(cfg.mrope_section):
        ax += [a] * n
    assert len(ax) == cfg.head_dim // 2, "mrope_section must sum to head_dim / 2"
    return np.asarray(ax, np.int32)


def packed_rows(P: int, G: int, C: int, pad_to: int = 64):
    """Row layout of the packed scoring pass: [prompt P | G completions of C | pad]. Returns
    (T_padded, row_lo, row_hi, logit_rows) with logit_rows[g*C + c] = row whose hidden state predicts completion
    token c of sequence g (prompt row P-1 for c == 0)."""
    T = P + G * C
    Tp = (T + pad_to - 1) // pad_to * pad_to
    lo = np.zeros(Tp, np.int32)
    hi = np.zeros(Tp, np.int32)
    hi[:T] = np.arange(T) + 1
    for g in range(G):
        lo[P + g * C: P + (g + 1) * C] = P + g * C
    rows = np.empty(G * C, np.int32)
    for g in range(G):
        rows[g * C] = P - 1
        rows[g * C + 1:(g + 1) * C] = P + g * C + np.arange(C - 1)
    return Tp, lo, hi, rows
