"""Per-step launch list from an ncu `--metrics gpu__time_duration.sum --csv` log of an eager bench.py run:
finds the period of the kernel-name sequence at the end of the log (= one step) and prints that step's launches
and the share of each kernel.  Times under ncu are cold-cache and serialised: compare SHARES, not absolutes.

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file l.csv python bench.py --steps 3 --warmup 3 --no-graph ...
    python tools/launch_list.py l.csv > profiles/rNN_launches_step.txt
"""
import csv
import sys
from collections import OrderedDict


def short(name):
    n = name.split("(")[0].strip()
    return n if len(n) <= 110 else n[:107] + "..."


def main(path):
    rows = []
    with open(path, newline="") as fh:
        lines = [l for l in fh if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            v_us = v / 1000.0 if unit.startswith("n") else (v if unit.startswith("u") else v * 1000.0)
            rows.append((short(r["Kernel Name"]), v_us))
    names = [n for n, _ in rows]
    N = len(names)
    # one step = the launches between two consecutive streaming-backward launches whose name sequences repeat
    # (the log also holds warm-up, the e2e leg and side measurements; the most common gap is the step length)
    anchors = [i for i, n in enumerate(names) if "k_raster_bwd2" in n]
    gaps = [b - a for a, b in zip(anchors, anchors[1:])]
    start = period = None
    if gaps:
        common = max(set(gaps), key=gaps.count)
        for a, b, c in zip(anchors, anchors[1:], anchors[2:]):
            if b - a == common and c - b == common and names[a:b] == names[b:c]:
                start, period = a + 1, common   # begin after the anchor: the step is printed in launch order up to the next one
                break
    if period is None:
        print("# no repeating step found in %d launches" % N)
        return
    step = rows[start:start + period]
    total = sum(t for _, t in step)
    print("# one eager step = %d launches (between two consecutive k_raster_bwd2 launches of the log), %.1f us" % (period, total))
    print("# (cold-cache, serialised under ncu: compare SHARES)")
    for n, t in step:
        print("%9.1f us  %s" % (t, n))
    agg = OrderedDict()
    for n, t in step:
        key = "torch glue (at::)" if ("at::" in n or "native::" in n) else n
        agg[key] = agg.get(key, 0.0) + t
    print("\n# share by kernel")
    for k, t in sorted(agg.items(), key=lambda kv: -kv[1]):
        print("%9.1f us %5.1f%%  %s" % (t, 100 * t / total, k))


if __name__ == "__main__":
    main(sys.argv[1])
