"""Per-launch table from an `ncu --metrics gpu__time_duration.sum --csv` log: stream, kernel, step index, grid, us."""
import csv, sys
from collections import OrderedDict


def load(path):
    rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 14 and r[0].isdigit()]
    L = OrderedDict()
    for r in rows:
        d = L.setdefault(int(r[0]), {"name": r[4].split("(")[0].replace("void ", "").replace("dbslmm::", ""),
                                     "grid": int(r[8].strip("()").split(",")[0]), "stream": r[6]})
        v = float(r[14].replace(",", ""))
        d[r[12]] = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(r[13], v)
    seq = list(L.values())
    kc = {}
    for s in seq:
        if "panel" in s["name"]:
            k = kc.get(s["stream"], 0); kc[s["stream"]] = k + 1; s["k"] = k
    return seq


if __name__ == "__main__":
    seq = load(sys.argv[1])
    tot = {}
    for s in seq:
        tot[s["name"]] = tot.get(s["name"], 0) + s["gpu__time_duration.sum"]
    for s in seq:
        if len(sys.argv) > 2:
            print(s["stream"], s["name"][:28], s.get("k", ""), s["grid"], round(s["gpu__time_duration.sum"], 1))
    print({k: round(v / 1e3, 3) for k, v in tot.items()})
