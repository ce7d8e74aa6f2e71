"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.
  python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r01_launches.txt
  python tools/summarize_ncu.py full gpurun_out/prof_chol_panel.ncu-rep profiles/r01_chol_panel_full.txt
"""
import collections, csv, io, subprocess, sys

KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "SM_A.TriageCompute.sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg.per_second"]


def launches(src, dst):
    txt = open(src).read()
    txt = txt[txt.index('"ID"'):]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(io.StringIO(txt)):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
        k = row["Kernel Name"].split("(")[0]
        a = agg.setdefault(k, [0, 0.0, 0.0])
        a[0] += 1; a[1] += v; a[2] = max(a[2], v)
        n += 1
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (serialised, cold cache: compare SHARES)\n# source: {src}; {n} launches, {tot/1e3:.3f} ms\n")
        f.write(f"{'kernel':42s} {'launches':>8s} {'total_ms':>10s} {'share':>7s} {'max_us':>9s}\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:42s} {a[0]:8d} {a[1]/1e3:10.3f} {a[1]/tot*100:6.1f}% {a[2]:9.1f}\n")
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    hdr, units, rows = r[0], r[1], r[2:]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none; source {src}; {len(rows)} launch(es)\n")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                f.write(f"{k} [{units[i]}]: " + " | ".join(row[i] for row in rows) + "\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
