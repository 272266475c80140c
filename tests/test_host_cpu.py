"""CPU suite, part 2: host logic of the product (indexing, layouts, config, C-ABI surface) and the world_size=2 gloo
path of the ZeRO-2 step."""
import os
import re
import socket

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    from time_r1_b200 import lib

    l = lib.load()
    header = open(os.path.join(ROOT, "include", "timer1.h")).read()
    declared = sorted(set(re.findall(r"\b(t1_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 38
    for name in declared:
        assert hasattr(l, name), f"{name} declared in include/timer1.h but not exported by libtimer1.so"
    for name in lib.exported_symbols():
        assert name in declared, f"{name} bound in lib.py but missing from include/timer1.h"
    assert l.t1_version() == 100
    assert lib.last_error() == "" or isinstance(lib.last_error(), str)


def test_product_has_no_oracle_dependency():
    """The product path must never import oracle/ (a CPU fallback would void every parity claim)."""
    for fn in os.listdir(os.path.join(ROOT, "time_r1_b200")):
        if fn.endswith(".py"):
            src = open(os.path.join(ROOT, "time_r1_b200", fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b|import_module\(.oracle|__import__\(.oracle", src, re.M), \
                f"time_r1_b200/{fn} imports the oracle"


def test_engine_fails_loudly_without_cuda():
    from time_r1_b200 import config, engine, weights

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = config.toy()
    mw = weights.ModelWeights(cfg, "cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        engine.Engine(cfg, mw, config.RunConfig())


@pytest.mark.parametrize("grid", [[(2, 8, 12)], [(1, 6, 10)], [(3, 4, 4), (2, 16, 16)], [(2, 32, 32)], [(4, 26, 46)]])
def test_vision_plan_matches_oracle_indexing(grid):
    from oracle import qwen_vl_oracle as O
    from time_r1_b200 import config, indexing

    cfg = config.ModelConfig()
    ocfg = O.default_cfg()
    plan = indexing.vision_plan(cfg, grid)
    widx, cu = O.vit_window_index(ocfg, grid)
    assert np.array_equal(plan.window_index, widx.numpy())
    # per-row window ranges reproduce unique_consecutive(cu_window_seqlens)
    cu = cu.numpy()
    for a, b in zip(cu[:-1], cu[1:]):
        assert (plan.win_lo[a:b] == a).all() and (plan.win_hi[a:b] == b).all()
    assert np.array_equal(np.sort(plan.perm_rows), np.arange(plan.n_patches))
    pos = O.vit_rot_pos(ocfg, grid).numpy()
    assert np.array_equal(plan.pos_hw.T, pos[plan.perm_rows])
    slot = 0
    for t, h, w in grid:
        for _ in range(t):
            assert (plan.full_lo[slot:slot + h * w] == slot).all() and (plan.full_hi[slot:slot + h * w] == slot + h * w).all()
            slot += h * w
    # Qwen2-VL: no windowing, identity permutation
    p2 = indexing.vision_plan(config.qwen2_vl_7b(), grid)
    assert np.array_equal(p2.perm_rows, np.arange(p2.n_patches)) and np.array_equal(p2.win_hi, p2.full_hi)


def test_packed_rows_layout_and_mask_semantics():
    from time_r1_b200 import indexing

    P, G, C = 37, 3, 5
    Tp, lo, hi, rows = indexing.packed_rows(P, G, C)
    assert Tp % 64 == 0 and Tp >= P + G * C
    # dense mask from (lo, hi, shared=P) == per-sequence causal masks of the reference's [prompt + completion] copies
    j = np.arange(Tp)[None, :]
    vis = (j < hi[:, None]) & ((j >= lo[:, None]) | (j < P))
    for g in range(G):
        seq_rows = list(range(P)) + list(range(P + g * C, P + (g + 1) * C))
        sub = vis[np.ix_(seq_rows, seq_rows)]
        assert np.array_equal(sub, np.tril(np.ones((P + C, P + C), bool)))
        other = [r for r in range(P, P + G * C) if r not in seq_rows]
        assert not vis[np.ix_(seq_rows[P:], other)].any(), "completions must not see each other"
    assert not vis[P + G * C:].any(), "padding rows see nothing"
    assert rows[0] == P - 1 and rows[1] == P and rows[C] == P - 1 and rows[C + 1] == P + C


def test_weights_layout_round_trip_and_fusion():
    from oracle import qwen_vl_oracle as O
    from time_r1_b200 import config, weights

    for fam in ("qwen2_5_vl", "qwen2_vl"):
        cfg = config.toy(fam)
        ocfg = O.default_cfg(**{k: v for k, v in cfg.to_dict().items() if k in O.default_cfg()})
        w = O.seeded_weights(ocfg, seed=7, std=0.05)
        mw = weights.ModelWeights(cfg, "cpu", world_size=2)
        mw.load_state_dict(w)
        sd = mw.export_state_dict()
        assert set(sd) == set(w)
        for k in w:
            assert torch.equal(sd[k].float().reshape(w[k].shape), w[k]), k
        assert mw.train.numel % (2 * 512) == 0
        qkv = mw.train["model.layers.0.self_attn.qkv.weight"]
        assert torch.equal(qkv[:256].float(), w["model.layers.0.self_attn.q_proj.weight"])
        for fp in (mw.train, mw.frozen):
            for t in fp.views.values():
                assert (t.data_ptr() - fp.buf.data_ptr()) % 256 == 0     # cudaMalloc bases are 256-byte aligned
    with pytest.raises(KeyError):
        weights.ModelWeights(config.toy(), "cpu").load_state_dict({})


def test_config_presets_and_flop_model():
    from time_r1_b200 import config, flops

    c = config.qwen2_5_vl_7b()
    assert c.qkv_dim == 4608 and c.vit_inter_padded == 3456 and c.patch_dim == 1176
    f = flops.step_flops(c, [(16, 32, 32)], 4236, 8, 200)
    assert 500e12 < f["total"] < 560e12                      # ~527 TFLOP per cfg2 step
    assert flops.reference_style_flops(c, [(16, 32, 32)], 4236, 8, 200) > 6 * f["total"]
    hf = {"model_type": "qwen2_5_vl", "hidden_size": 2048, "intermediate_size": 11008, "num_hidden_layers": 36,
          "num_attention_heads": 16, "num_key_value_heads": 2, "vocab_size": 151936, "tie_word_embeddings": True,
          "rope_scaling": {"mrope_section": [16, 24, 24]}, "vision_config": {"hidden_size": 1280, "depth": 32, "num_heads": 16,
          "intermediate_size": 3420, "window_size": 112, "fullatt_block_indexes": [7, 15, 23, 31], "tokens_per_second": 2}}
    c3 = config.from_hf_config(hf)
    assert c3.head_dim == 128 and c3.tie_word_embeddings and c3.fullatt_block_indexes == (7, 15, 23, 31)


def test_rank_indices_partition():
    from time_r1_b200 import dist_utils

    parts = [dist_utils.rank_indices(103, 4, r, seed=42) for r in range(4)]
    assert sorted(sum(parts, [])) == list(range(103))
    assert dist_utils.shard_bounds(4096, 4, 3) == (3072, 4096)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _zero2_worker(rank, world, port, q):
    import torch.distributed as dist

    from time_r1_b200 import dist_utils

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    n = 4096
    param = torch.randn(n, generator=g)
    local_grad = torch.randn(n, generator=torch.Generator().manual_seed(100 + rank))      # each rank: its own prompt-group
    lo, hi = dist_utils.shard_bounds(n, world, rank)
    m, v = torch.zeros(hi - lo), torch.zeros(hi - lo)

    def adamw(gs):
        nonlocal m, v
        m = 0.9 * m + 0.1 * gs
        v = 0.999 * v + 0.001 * gs * gs
        return param[lo:hi] - 1e-3 * (m / 0.1) / ((v / 0.001).sqrt() + 1e-8)

    dist_utils.zero2_step(param, local_grad, adamw, world, rank)
    metric = torch.tensor([float(rank), 1.0])
    dist.all_reduce(metric)
    q.put((rank, param.clone(), (metric / world).tolist()))
    dist.destroy_process_group()


def test_zero2_sharded_step_equals_unsharded_gloo_world2():
    import torch.multiprocessing as mp

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_zero2_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in ps:
        p.join(60)
    # unsharded reference: mean gradient over the two ranks, one AdamW step on the whole vector
    n = 4096
    param = torch.randn(n, generator=torch.Generator().manual_seed(0))
    grad = sum(torch.randn(n, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)) / world
    m, v = 0.1 * grad, 0.001 * grad * grad
    ref = param - 1e-3 * (m / 0.1) / ((v / 0.001).sqrt() + 1e-8)
    for rank, p, metric in res:
        assert torch.allclose(p, ref, atol=1e-6), f"rank {rank}: sharded step differs from the unsharded one"
        assert metric == [0.5, 1.0]


@pytest.mark.parametrize("grid", [[(147, 6, 12)], [(8, 26, 46)], [(60, 16, 28)], [(384, 6, 10)], [(2, 8, 12), (3, 6, 10)], [(16, 32, 32)]])
def test_window_plan_matches_hf_get_window_index_on_timerft_grids(grid):
    """cfg4: TimeRFT-shaped clips have many temporal slots with tiny, ragged spatial grids.  The host-side
    window plan (permutation + per-row window ranges) against HF's own Qwen2.5-VL `get_window_index` (Q25:411-451, the
    class timer1_trainer.py:244 instantiates) — only window_size / merge size / patch size matter, so the HF tower is built
    with one tiny block."""
    torch = pytest.importorskip("torch")
    pytest.importorskip("transformers")
    from transformers.models.qwen2_5_vl.configuration_qwen2_5_vl import Qwen2_5_VLVisionConfig
    from transformers.models.qwen2_5_vl.modeling_qwen2_5_vl import Qwen2_5_VisionTransformerPretrainedModel

    from time_r1_b200 import config, indexing

    tower = Qwen2_5_VisionTransformerPretrainedModel(
        Qwen2_5_VLVisionConfig(depth=1, hidden_size=32, num_heads=2, intermediate_size=64, out_hidden_size=32, fullatt_block_indexes=[0]))
    cfg = config.ModelConfig()
    assert (tower.window_size, tower.spatial_merge_size, tower.patch_size) == (cfg.window_size, cfg.spatial_merge_size, cfg.patch_size)
    widx, cu = tower.get_window_index(torch.tensor(grid))
    cu = torch.unique_consecutive(torch.tensor(cu)).numpy()          # as the forward does (Q25:476)
    plan = indexing.vision_plan(cfg, grid)
    assert np.array_equal(plan.window_index, widx.numpy())
    for a, b in zip(cu[:-1], cu[1:]):
        assert (plan.win_lo[a:b] == a).all() and (plan.win_hi[a:b] == b).all()
    assert cu[-1] == plan.n_patches


def _clip_metrics_worker(rank, world, port, q):
    import collections
    import types

    import torch.distributed as dist

    from time_r1_b200.trainer import TimeR1_Trainer_ft

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    local = torch.tensor([[0.125, 0.5], [0.375, 0.25]][rank])              # this rank's (low, high) clip fractions
    mean = local.clone()
    dist.all_reduce(mean)
    mean /= world                                                          # what _log_metrics hands over (AVG all-reduce)
    me = types.SimpleNamespace(_metrics=collections.defaultdict(list), _local_clip=local, world=world, metric_funcs=[], _last_completions=None)
    TimeR1_Trainer_ft._extra_metrics(me, float(mean[0]), float(mean[1]), [])
    q.put((rank, {k: v[0] for k, v in me._metrics.items()}))
    dist.destroy_process_group()


def test_ft_clip_ratio_metrics_across_two_ranks_gloo():
    """timer1_trainer_ft.py:820-842: gather the per-rank fractions, log nanmean / nanmin(low) / nanmax(high) — here one
    AVG all-reduce (in _log_metrics) plus one MIN all-reduce on (low, -high)."""
    import torch.multiprocessing as mp

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_clip_metrics_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in ps:
        p.join(60)
    for rank, m in res:
        assert m == {"clip_ratio/low_mean": 0.25, "clip_ratio/low_min": 0.125, "clip_ratio/high_mean": 0.375,
                     "clip_ratio/high_max": 0.5, "clip_ratio/region_mean": 0.625}, (rank, m)


def _bucket_worker(rank, world, port, q):
    import torch.distributed as dist

    from time_r1_b200.dist_utils import BucketPlan

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 4096
    plan = BucketPlan(n, world, n_buckets=5, align=16)
    param = torch.randn(n, generator=torch.Generator().manual_seed(0))
    grad = torch.randn(n, generator=torch.Generator().manual_seed(100 + rank))
    master = plan.gather_shard(param, rank, torch.empty(plan.shard))
    shard_g = torch.zeros(plan.shard)
    # the backward finishes the buffer from its end: hand the buckets out as the "final from" offset moves down
    order = []
    plan.reset()
    for done_from in (n, 3000, 2999, 1700, 900, 0):
        for b in plan.ready_buckets(done_from):
            order.append(b)
            plan.reduce_scatter(b, grad, shard_g, rank)
    assert plan.ready_buckets(0) == []
    m = 0.1 * shard_g
    v = 0.001 * shard_g * shard_g
    new = master - 1e-3 * (m / 0.1) / ((v / 0.001).sqrt() + 1e-8)
    for b in range(len(plan)):
        lo, hi = plan.slice_of(b, rank)
        a, z = plan.shard_range(b)
        param[lo:hi] = new[a:z]                      # AdamW writes the slice in place
        plan.all_gather(b, param, rank)
    q.put((rank, param.clone(), order, plan.bounds))
    dist.destroy_process_group()


def test_bucketed_zero2_exchange_equals_unsharded_step_gloo_world2():
    """dist_utils.BucketPlan (the engine's overlapped ZeRO-2 exchange): buckets handed out in backward order, reduce-scatter
    per bucket into the contiguous shard, AdamW per slice, in-place all-gather == one unsharded AdamW step on the mean gradient."""
    import torch.multiprocessing as mp

    from time_r1_b200.dist_utils import BucketPlan

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_bucket_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in ps:
        p.join(60)
    n = 4096
    param = torch.randn(n, generator=torch.Generator().manual_seed(0))
    grad = sum(torch.randn(n, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)) / world
    ref = param - 1e-3 * grad / (grad.abs() + 1e-8)
    for rank, p, order, bounds in res:
        assert torch.allclose(p, ref, atol=1e-6), f"rank {rank}"
        assert order == sorted(order, reverse=True) and len(order) == len(bounds)       # last bucket first, each exactly once
    # layout facts the engine relies on
    plan = BucketPlan(7_667_998_720, 8)
    assert len(plan) == 32 and plan.shard * 8 == plan.numel and all((hi - lo) % (8 * 512) == 0 for lo, hi in plan.bounds)
    assert [plan.shard_range(b)[1] - plan.shard_range(b)[0] for b in range(32)] == plan.slice_len
    one = BucketPlan(4096, 1)
    assert one.bounds == [(0, 4096)] and one.slice_of(0, 0) == (0, 4096)
