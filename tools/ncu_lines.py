"""Per-source-line hot spots of one kernel in an .ncu-rep (needs -lineinfo + --import-source on):
    python tools/ncu_lines.py <rep> <kernel regex> [top N]
Prints, per file:line, executed warp instructions, sampled stalls and the dominant stall reasons."""
import csv
import subprocess
import sys


def main(rep, kern, top=40):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur, hdr, agg = None, None, {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or r[2] != "-":  # keep the per-line summary rows (Address == "-")
            continue
        d = dict(zip(hdr[4:], r[4:]))
        key = (cur, int(r[0]))
        inst = int(d.get("Instructions Executed", 0) or 0)
        samp = int(d.get("# Samples", 0) or 0)
        stalls = {k[6:]: int(v or 0) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k}
        a = agg.setdefault(key, [0, 0, {}, r[1], 0])
        a[0] += inst
        a[1] += samp
        a[4] += int(d.get("Thread Instructions Executed", 0) or 0)
        for k, v in stalls.items():
            a[2][k] = a[2].get(k, 0) + v
    tot_i = sum(a[0] for a in agg.values()) or 1
    tot_s = sum(a[1] for a in agg.values()) or 1
    print("total warp instructions %d, samples %d" % (tot_i, tot_s))
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        st = sorted(a[2].items(), key=lambda kv: -kv[1])[:3]
        print("%-20s:%-5d inst %5.1f%%  samples %5.1f%%  lanes %4.1f  %-40s | %s" % (
            key[0][:20], key[1], 100.0 * a[0] / tot_i, 100.0 * a[1] / tot_s, a[4] / max(a[0], 1),
            " ".join("%s=%d" % kv for kv in st if kv[1]), a[3].strip()[:70]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
