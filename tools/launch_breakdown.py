"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py into the
per-kernel shares of ONE step (profiles/ evidence). Usage: launch_breakdown.py <csv> [step_index]"""
import collections
import csv
import json
import re
import sys


def main(path, which=3):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    data = [(r[kn], float(r[mv].replace(",", ""))) for r in rows[hi + 1:] if len(r) > mv and r[mv]]
    starts = [i for i, d in enumerate(data) if "lss_camera_mats" in d[0]]
    step = data[starts[which]:starts[which + 1]]
    tot = sum(t for _, t in step)
    agg = collections.OrderedDict()
    for n, t in step:
        n = re.sub(r"\(.*", "", n).replace("void ", "").replace("dbev::<unnamed>::", "dbev::")[:80]
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += t
    out = {"source": path, "note": "ncu per-launch times are cold-cache and serialised: compare SHARES",
           "kernels_in_step": len(step), "sum_us": round(tot / 1e3, 1),
           "ours_us": round(sum(t for n, (c, t) in agg.items() if n.startswith("dbev::")) / 1e3, 1),
           "kernels": [{"name": n, "launches": c, "us": round(t / 1e3, 1), "share": round(t / tot, 4)}
                       for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])]}
    return out


if __name__ == "__main__":
    res = main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3)
    print(json.dumps(res, indent=1))
