"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: launches, total device time, share.
With --decode-replays N the kernels of the captured decode step (two replays are in the list: the first and the
last) are scaled to N replays, which is what one bench step executes.
    python tools/ncu_summarize.py gpurun_out/launches.csv [--decode-replays 199 --decode-launches 202]"""
import collections
import csv
import re
import sys


def short(name: str) -> str:
    name = re.sub(r"<unnamed>::|\(anonymous namespace\)::|void ", "", name)
    m = re.match(r"([A-Za-z0-9_:]+(?:<[^(]*>)?)", name)
    return (m.group(1) if m else name)[:70]


def main():
    path = sys.argv[1]
    replays = int(sys.argv[sys.argv.index("--decode-replays") + 1]) if "--decode-replays" in sys.argv else 0
    per_step = int(sys.argv[sys.argv.index("--decode-launches") + 1]) if "--decode-launches" in sys.argv else 0
    rows = [r for r in csv.reader(open(path, newline="")) if len(r) > 8]
    hdr = rows[0]
    ki, mi, vi, ui, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
    launches = []
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        us = v / 1e3 if r[ui] in ("nsecond", "ns") else (v if r[ui] in ("usecond", "us") else v * 1e3)
        launches.append((int(r[ii]), short(r[ki]), us))
    launches.sort()
    # the decode-graph kernels are the last 2 * per_step launches of the rollout phase; identify them as the two
    # contiguous runs whose names start a decode step (gather_rows ... incr)
    decode_idx = set()
    if replays and per_step:
        names = [n for _, n, _ in launches]
        starts = [i for i in range(len(names) - per_step + 1) if names[i].startswith("gather_rows") and names[i + per_step - 1].startswith("incr")]
        for s in starts[:2]:
            decode_idx.update(range(s, s + per_step))
    agg = collections.defaultdict(lambda: [0.0, 0.0])
    for i, (_, n, us) in enumerate(launches):
        w = (replays / 2.0) if i in decode_idx else 1.0
        key = n + (" [decode graph]" if i in decode_idx else "")
        agg[key][0] += w
        agg[key][1] += us * w
    total = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(launches)} profiled launches; decode-graph launches found: {len(decode_idx)} (scaled to {replays} replays)")
    print(f"# total device time (serialised, cold-cache, per ncu): {total / 1e3:.1f} ms")
    print(f"{'kernel':86s} {'launches':>10s} {'ms':>10s} {'share':>7s} {'us/launch':>10s}")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:86s} {n:10.0f} {us / 1e3:10.2f} {100 * us / total:6.1f}% {us / max(n, 1):10.2f}")


if __name__ == "__main__":
    main()
