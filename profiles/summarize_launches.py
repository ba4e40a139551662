"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:
    python profiles/summarize_launches.py profiles/<file>.csv
(per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes)."""
import collections
import csv
import re
import sys


def summarize(path):
    hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
    for r in csv.reader(open(path)):
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") == "gpu__time_duration.sum":
                name = re.sub(r"\(.*", "", d["Kernel Name"]).split("<")[0].replace("unnamed>::", "").replace("void ", "")
                v = float(d["Metric Value"].replace(",", ""))
                v = v / 1e3 if d["Metric Unit"] in ("ns", "nsecond") else v
                agg[name][0] += 1
                agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    lines = [f"{'kernel':40s} {'launches':>8s} {'total_us':>12s} {'share':>7s}"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{k:40s} {v[0]:8d} {v[1]:12.1f} {v[1] / tot:7.3f}")
    return "\n".join(lines)


if __name__ == "__main__":
    print(summarize(sys.argv[1]))
