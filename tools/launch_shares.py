"""Per-kernel shares of ONE resident fit from an ncu launch list.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -k regex:'chol_|gram_|decode_rows|backsolve|fill_z|block_flags|rows_missing|block_missing|snp_stats' -c 700 --csv --log-file X.csv \
        python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity
    python tools/launch_shares.py X.csv profiles/rNN_launches.txt [profiles/chol_traffic.json]

A resident fit launches the row decoder exactly once: the launches from the 2nd decoder launch up to the 3rd are one fit.
Per-launch times are serialised and cold-cache: compare SHARES with bench.py's live CUDA-event numbers."""
import csv, json, sys
from collections import OrderedDict


def load(path):
    rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 14 and r[0].isdigit()]
    L = OrderedDict()
    for r in rows:
        d = L.setdefault(int(r[0]), {"name": r[4].split("(")[0].replace("void ", "").replace("dbslmm::", "")})
        v = float(r[14].replace(",", ""))
        unit = r[13]
        if r[12] == "gpu__time_duration.sum":
            v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(unit, v)          # -> us
        else:
            v = {"byte": v, "Kbyte": v * 1e3, "Mbyte": v * 1e6, "Gbyte": v * 1e9}.get(unit, v)   # -> bytes
        d[r[12]] = v
    return list(L.values())


def main():
    seq = load(sys.argv[1])
    starts = [i for i, s in enumerate(seq) if s["name"].startswith("decode_rows_kernel")]
    if len(starts) < 3:
        raise SystemExit(f"need at least three fits in the list, found {len(starts)} decoder launches")
    fit = seq[starts[1]:starts[2]]
    agg = OrderedDict()
    for s in fit:
        a = agg.setdefault(s["name"], [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += s.get("gpu__time_duration.sum", 0.0)
        a[2] += s.get("dram__bytes_read.sum", 0.0)
        a[3] += s.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    with open(sys.argv[2], "w") as f:
        f.write("# ONE genome-wide fit (C3, no missing calls, resident panel): the launches of the second fit of\n"
                "#   ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:<library kernels> "
                "python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity\n"
                "# per-launch times are serialised and cold-cache: compare SHARES with bench.py's live CUDA-event numbers\n")
        f.write(f"{'kernel':38s} {'launches':>8s} {'ms_per_fit':>10s} {'share':>7s} {'dram_read_GB':>13s} {'dram_write_GB':>14s}\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:38s} {a[0]:8d} {a[1]/1e3:10.3f} {a[1]/tot*100:6.1f}% {a[2]/1e9:13.3f} {a[3]/1e9:14.3f}\n")
        f.write(f"{'total':38s} {sum(a[0] for a in agg.values()):8d} {tot/1e3:10.3f}\n")
    print(open(sys.argv[2]).read())
    if len(sys.argv) > 3:
        ch = [a for k, a in agg.items() if k.startswith("chol_")]
        out = {"dram_bytes_per_fit": sum(a[2] + a[3] for a in ch), "dram_read_bytes": sum(a[2] for a in ch),
               "dram_write_bytes": sum(a[3] for a in ch), "launches": sum(a[0] for a in ch),
               "source": f"{sys.argv[2]} (ncu dram__bytes_read.sum + dram__bytes_write.sum over all chol_* launches of one fit)"}
        json.dump(out, open(sys.argv[3], "w"), indent=1)
        print(out)


if __name__ == "__main__":
    main()
