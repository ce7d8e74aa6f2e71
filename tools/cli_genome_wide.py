#!/usr/bin/env python
"""The real drop-in, timed: a synthetic genome written as REAL files (22 x PLINK .bed/.bim/.fam reference panels, GEMMA
.assoc.txt small- and large-effect lists, EUR LD-block files) and fitted by ONE cold `build/dbslmm --manifest` process,
beside the unmodified reference binary (oracle/_ref/dbslmm_ref, when it travelled with the tree) on one chromosome.

  python tools/cli_genome_wide.py [--out DIR] [--gpus 1,8] [--json profiles/cli_genome_wide.json] [--config c3|c2]

Everything here is harness: the product under test is the `dbslmm` binary.  The reference reads the same files
(scr/dbslmm.cpp:232-364, scr/dtpr.cpp:83-123,178-220).
"""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def write_genome(out, cfg, seed, n_test=40):
    """Returns (manifest path, per-chromosome dicts).  SNP j of the synthetic genome is row j of its chromosome's panel."""
    import torch
    import bench
    from dbslmm_b200 import synth
    ns = argparse.Namespace(config=cfg, missing=0.0, seed=seed)
    dev = torch.device("cuda", 0)
    w = bench.build_workload(ns, torch, dev, seed)
    total, chroms, cap, n_ref, n_obs = bench.CONFIGS[cfg]
    lens = json.load(open(os.path.join(ROOT, "dbslmm_b200", "data", "eur_ld_block_lengths.json")))
    sizes = w["sizes"]
    large = np.zeros(w["n_snp"], bool)
    large[w["l_pos"]] = True
    rng = np.random.default_rng(seed + 5)
    # allele frequencies of the summary statistics = the panel's own (so -mafMax 0.2 keeps every SNP)
    from dbslmm_b200 import _abi
    eng = _abi.Engine(0)
    eng.load_bed(w["bed"], n_ref)
    af, _ = eng.snp_stats()
    eng.close()
    os.makedirs(out, exist_ok=True)
    jobs, man = [], []
    blk0 = snp0 = 0
    for c in chroms:
        L = np.asarray(lens[str(c)], np.int64)
        nb = L.size
        starts = np.concatenate([[1000], 1000 + np.cumsum(L)[:-1]])
        ends = starts + L
        m = sizes[blk0:blk0 + nb]
        n_c = int(m.sum())
        ps = np.concatenate([starts[b] + 1 + (np.arange(m[b], dtype=np.int64) * (L[b] - 2)) // max(int(m[b]), 1) for b in range(nb)]) if n_c else np.zeros(0, np.int64)
        idx = np.arange(snp0, snp0 + n_c)
        rs = np.char.add("rs", idx.astype(str))
        pre = os.path.join(out, f"chr{c}")
        with open(pre + ".bed", "wb") as f:
            f.write(bytes([0x6C, 0x1B, 0x01]))
            f.write(w["bed"][snp0:snp0 + n_c].tobytes())
        with open(pre + ".fam", "w") as f:
            f.write("".join(f"f{i} i{i} 0 0 0 -9\n" for i in range(n_ref)))
        a1 = np.where(idx % 2 == 0, "A", "C"); a2 = np.where(idx % 2 == 0, "G", "T")
        with open(pre + ".bim", "w") as f:
            f.write("\n".join(f"{c}\t{r}\t0\t{p}\t{x}\t{y}" for r, p, x, y in zip(rs, ps, a1, a2)) + "\n")
        with open(pre + "_blocks.bed", "w") as f:
            f.write("".join(f"chr{c}\t{s}\t{e}\n" for s, e in zip(starts, ends)))
        z = w["z"][snp0:snp0 + n_c]
        se = 0.01
        lines = [f"{c}\t{r}\t{p}\t0\t{n_obs}\t{x}\t{y}\t{f_:.6f}\t{zz * se:.10e}\t{se:.6e}\t{1.0:.3e}"
                 for r, p, x, y, f_, zz in zip(rs, ps, a1, a2, af[snp0:snp0 + n_c], z)]
        lg = large[snp0:snp0 + n_c]
        with open(pre + "_s.txt", "w") as f:
            f.write("\n".join(l for l, k in zip(lines, lg) if not k) + "\n")
        with open(pre + "_l.txt", "w") as f:
            f.write("".join(l + "\n" for l, k in zip(lines, lg) if k))
        man.append("\t".join([pre + "_s.txt", pre + "_l.txt" if lg.any() else "-", pre, pre + "_blocks.bed", pre + "_out"]))
        jobs.append({"chr": c, "pre": pre, "n_snp": n_c, "n_blocks": int(nb), "has_large": bool(lg.any()), "ps": ps, "n_ref": n_ref, "n_obs": n_obs})
        blk0 += nb
        snp0 += n_c
    mf = os.path.join(out, "manifest.txt")
    open(mf, "w").write("\n".join(man) + "\n")
    # a small test panel for the last chromosome (the fork's CLI cannot run without -dat_str / -test_indicator_file)
    j = jobs[-1]
    Gt = synth.make_genotypes(rng, [j["n_snp"]], n_test)
    with open(j["pre"] + "_test.bed", "wb") as f:
        f.write(bytes([0x6C, 0x1B, 0x01])); f.write(synth.pack_bed(Gt).tobytes())
    bim = open(j["pre"] + ".bim").read()
    open(j["pre"] + "_test.bim", "w").write(bim)
    open(j["pre"] + "_ind.txt", "w").write("1\n" * n_test)
    return mf, jobs, w


def run(cmd, cwd=None):
    t = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=cwd)
    return time.perf_counter() - t, r


def parse_timing(stdout):
    d = {}
    m = re.search(r"Fitting time: ([0-9.eE+-]+) seconds", stdout)
    if m:
        d["fitting_time_s"] = float(m.group(1))
    m = re.search(r"\[timing\] panel files ([0-9.eE+-]+) s, summary statistics \+ matching ([0-9.eE+-]+) s, plan \+ fit ([0-9.eE+-]+) s .*output ([0-9.eE+-]+) s; (\d+) host", stdout)
    if m:
        d.update(panel_files_s=float(m.group(1)), sumstats_matching_s=float(m.group(2)), plan_fit_s=float(m.group(3)), output_s=float(m.group(4)), host_threads=int(m.group(5)))
    return d


def read_eff(path):
    rows = [ln.split(" ") for ln in open(path).read().strip().split("\n") if ln]
    return [r[0] for r in rows], np.array([float(r[2]) for r in rows])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="/dev/shm/dbslmm_gw" if os.path.isdir("/dev/shm") else "/tmp/dbslmm_gw")
    ap.add_argument("--gpus", default="1")
    ap.add_argument("--json", default=os.path.join(ROOT, "gpurun_out", "cli_genome_wide.json"))
    ap.add_argument("--config", default="c3")
    ap.add_argument("--seed", type=int, default=20240003)
    args = ap.parse_args()
    cli = os.path.join(ROOT, "build", "dbslmm")
    t0 = time.perf_counter()
    mf, jobs, w = write_genome(args.out, args.config, args.seed)
    res = {"config": args.config, "chromosomes": len(jobs), "snps": int(sum(j["n_snp"] for j in jobs)), "blocks": int(sum(j["n_blocks"] for j in jobs)),
           "n_ref": jobs[0]["n_ref"], "files_written_s": time.perf_counter() - t0, "host_cores": os.cpu_count(),
           "bytes": {"bed": int(sum(os.path.getsize(j["pre"] + ".bed") for j in jobs)), "bim": int(sum(os.path.getsize(j["pre"] + ".bim") for j in jobs)),
                     "sumstats": int(sum(os.path.getsize(j["pre"] + "_s.txt") + os.path.getsize(j["pre"] + "_l.txt") for j in jobs))},
           "runs": []}
    nsnp, n_obs = res["snps"], jobs[0]["n_obs"]
    base = ["--manifest", mf, "-n", str(n_obs), "-nsnp", str(nsnp), "-h", "0.5", "-t", str(min(os.cpu_count() or 1, 100))]
    for g in [int(x) for x in args.gpus.split(",")]:
        for maf in ("1", "0.2"):
            for rep in range(2):           # every process is cold for CUDA (context, module load, workspace); rep 1 has the files in the page cache
                wall, r = run([cli] + base + ["-mafMax", maf, "--gpus", str(g), "--dump-beta-bin", os.path.join(args.out, f"beta_g{g}_m{maf}.bin")])
                d = {"gpus": g, "mafMax": float(maf), "rep": rep, "rc": r.returncode, "process_wall_s": wall}
                d.update(parse_timing(r.stdout))
                if r.returncode != 0:
                    d["stderr"] = r.stderr[-400:]
                res["runs"].append(d)
                print(d, flush=True)
    # ---- same betas whatever the GPU count / MAF path
    def load_bin(p):
        raw = open(p, "rb").read()
        nf, tl, ts = np.frombuffer(raw[:24], np.int64)
        v = np.frombuffer(raw[24:], np.float64)
        return v[: nf * tl], v[nf * tl:]
    ref_bin = None
    for g in [int(x) for x in args.gpus.split(",")]:
        for maf in ("1", "0.2"):
            p = os.path.join(args.out, f"beta_g{g}_m{maf}.bin")
            if os.path.exists(p):
                bl, bs = load_bin(p)
                if ref_bin is None:
                    ref_bin = (bl, bs)
                res.setdefault("beta_max_rel_vs_first_run", {})[f"gpus{g}_mafMax{maf}"] = float(
                    max(np.abs(bs - ref_bin[1]).max() / np.abs(ref_bin[1]).max(), np.abs(bl - ref_bin[0]).max() / max(np.abs(ref_bin[0]).max(), 1e-300)))
    # ---- the last chromosome alone: this CLI vs the unmodified reference CLI on the same files
    j = jobs[-1]
    one = ["-s", j["pre"] + "_s.txt", "-r", j["pre"], "-b", j["pre"] + "_blocks.bed", "-n", str(n_obs), "-nsnp", str(nsnp), "-h", "0.5", "-mafMax", "0.2"]
    if j["has_large"]:
        one += ["-l", j["pre"] + "_l.txt"]
    wall, r = run([cli] + one + ["-t", "1", "-eff", j["pre"] + "_one"])
    d = {"what": f"this CLI, chr{j['chr']} alone ({j['n_snp']} SNPs, {j['n_blocks']} blocks)", "rc": r.returncode, "process_wall_s": wall}
    d.update(parse_timing(r.stdout))
    res["one_chromosome"] = [d]
    print(d, flush=True)
    ref_cli = os.path.join(ROOT, "oracle", "_ref", "dbslmm_ref")
    if os.path.exists(ref_cli):
        thr = str(min(os.cpu_count() or 1, 100))
        wall, r = run([ref_cli] + one + ["-t", thr, "-eff", j["pre"] + "_ref", "-test_indicator_file", j["pre"] + "_ind.txt", "-dat_str", j["pre"] + "_test"], cwd=args.out)
        d = {"what": f"UNMODIFIED reference CLI (oracle/_ref/dbslmm_ref over oracle/shim), chr{j['chr']} alone, -t {thr} (incl. the fork's variance side channel on 40 test individuals)",
             "rc": r.returncode, "process_wall_s": wall}
        d.update(parse_timing(r.stdout))
        if r.returncode == 0 and os.path.exists(j["pre"] + "_ref.txt") and os.path.exists(j["pre"] + "_one.txt"):
            s1, b1 = read_eff(j["pre"] + "_one.txt")
            s2, b2 = read_eff(j["pre"] + "_ref.txt")
            d["same_rows"] = s1 == s2
            if s1 == s2:
                d["beta_max_rel_6_digits"] = float(np.abs(b1 - b2).max() / np.abs(b2).max())
        else:
            d["stderr"] = r.stderr[-400:]
        res["one_chromosome"].append(d)
        print(d, flush=True)
    os.makedirs(os.path.dirname(args.json), exist_ok=True)
    json.dump(res, open(args.json, "w"), indent=1)
    shutil.rmtree(args.out, ignore_errors=True)


if __name__ == "__main__":
    main()
