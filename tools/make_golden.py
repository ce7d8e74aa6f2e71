"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref, built by
oracle/build_ref.sh from /root/reference/scr over oracle/shim) in this container.

  c1_testdat.npz  BASELINE.json configs[0]: test_dat chr1 (summary_gemma_chr1.assoc.txt + ref_chr1),
                  h2 = 0.5, nsnp = 996, n = 2400, mafMax 0.2, -t 1.  Holds the input files as byte
                  strings (so the CLI test can re-create them on the GPU box), the matched CSR
                  problem, the reference's FP64 betas (harness: readSNPIm + nomalizeVec + estBlock)
                  for LMM and DBSLMM mode and the `dbslmm` CLI's text output.
  synth_ragged.npz  seeded synthetic panel with ragged blocks and 2 % missing calls + reference betas.
The large/small split restates `plink --clump` (r2 0.2, 1000 kb, p1 1e-6) as software/DBSLMM.R:140-145
calls it; plink itself is not in this image.
"""
import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dbslmm_b200 import hostio as H, synth
from oracle import oracle as O, refharness as R

REF = "/root/reference"
TD = REF + "/test_dat/"
GOLD = os.path.join(ROOT, "tests", "golden")


def clump(summ, lines, bim, bed, n_ref):
    p = np.array([float(l.split("\t")[10]) for l in lines])
    order = np.argsort(p, kind="stable")
    chosen, cache = [], {}
    def g(i):
        if i not in cache:
            x, _ = O.read_snp_im(bed, bim[summ.snp[i]][0], n_ref)
            cache[i] = O.normalize(x)
        return cache[i]
    for i in order:
        if p[i] >= 1e-6:
            break
        if summ.snp[i] not in bim:
            continue
        ok = True
        for c in chosen:
            if abs(summ.ps[i] - summ.ps[c]) <= 1_000_000:
                r = float(np.dot(g(i), g(c))) / (n_ref - 1)
                if r * r >= 0.2:
                    ok = False
                    break
        if ok:
            chosen.append(i)
    return sorted(chosen)


def c1():
    n_ref = H.read_fam_count(TD + "ref_chr1.fam")
    bim, nsnp = H.read_bim(TD + "ref_chr1.bim")
    bed = H.read_bed(TD + "ref_chr1.bed", nsnp, n_ref)
    lines = [l for l in open(TD + "summary_gemma_chr1.assoc.txt").read().split("\n") if l]
    summ = H.read_summ(TD + "summary_gemma_chr1.assoc.txt")
    large = clump(summ, lines, bim, bed, n_ref)
    ls = set(large)
    l_txt = "\n".join(lines[i] for i in large) + "\n"
    s_txt = "\n".join(lines[i] for i in range(len(lines)) if i not in ls) + "\n"
    blk_txt = open(REF + "/block_data/EUR/chr1.bed").read()
    bs, be = H.read_block(REF + "/block_data/EUR/chr1.bed")
    maf = O.snp_maf(bed, nsnp, n_ref)
    n_obs, sigma_s = 2400, 0.5 / 996
    out = {}
    with tempfile.TemporaryDirectory() as td:
        open(td + "/l.txt", "w").write(l_txt)
        open(td + "/s.txt", "w").write(s_txt)
        open(td + "/ind.txt", "w").write("\n".join(["1"] * 20 + ["0"] * 83) + "\n")
        # ---- the unmodified CLI (DBSLMM mode, then LMM mode = no -l)
        for mode, extra in (("dbslmm", ["-l", td + "/l.txt", "-s", td + "/s.txt"]), ("lmm", ["-s", TD + "summary_gemma_chr1.assoc.txt"])):
            cmd = [R.CLI] + extra + ["-r", TD + "ref_chr1", "-n", "2400", "-nsnp", "996", "-mafMax", "0.2", "-b",
                   REF + "/block_data/EUR/chr1.bed", "-h", "0.5", "-t", "1", "-eff", td + "/out_" + mode,
                   "-test_indicator_file", td + "/ind.txt", "-dat_str", TD + "test_chr1"]
            rc = subprocess.run(cmd, cwd=td, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL).returncode
            if rc != 0:
                # the fork's small-only calcBlock sizes `gg` with n_obs (dbslmmfit.cpp:585) and assigns it
                # into an n_test-row column (:593): it aborts unless n_obs == n_test, so the unmodified
                # CLI has no LMM-mode output on this fixture; LMM betas come from the harness below.
                print("reference CLI failed in mode", mode, "rc", rc)
                continue
            out["cli_" + mode + "_txt"] = open(td + "/out_" + mode + ".txt").read()
            out["cli_" + mode + "_badsnps"] = open(td + "/out_" + mode + ".badsnps").read()
            # the fork's variance.txt (n_test x num_block; the shim writes Armadillo's arma_ascii layout)
            vt = open(td + "/variance.txt").read().split("\n")
            nr, nc = (int(x) for x in vt[1].split())
            out["cli_" + mode + "_variance"] = np.array([[float(x) for x in ln.split()] for ln in vt[2:2 + nr]]).reshape(nr, nc)
    # ---- matched CSR problems + FP64 reference betas through the harness
    def problem(summ_subset_idx):
        sub = H.Summ([summ.snp[i] for i in summ_subset_idx], summ.ps[summ_subset_idx], [summ.a1[i] for i in summ_subset_idx],
                     [summ.a2[i] for i in summ_subset_idx], summ.maf[summ_subset_idx], summ.z[summ_subset_idx])
        keep, pos = H.match_ref(sub, bim, maf, 0.2)
        blk = H.add_block(sub.ps[keep], bs, be)
        ok = blk >= 0
        return H.to_csr(blk, len(bs)), pos[ok], sub.z[keep][ok]
    all_idx = np.arange(len(lines))
    small_idx = np.array([i for i in all_idx if i not in ls])
    large_idx = np.array(large)
    with R.BedFile(bed) as bf:
        off, pos, z = problem(all_idx)
        b_lmm, _ = R.est_path(bf.path, n_ref, n_obs, sigma_s, off, pos, z, threads=1)
        s_off, s_pos, s_z = problem(small_idx)
        l_off, l_pos, l_z = problem(large_idx)
        b_s, b_l = R.est_path(bf.path, n_ref, n_obs, sigma_s, s_off, s_pos, s_z, l_off, l_pos, l_z, threads=1)
    tn = H.read_fam_count(TD + "test_chr1.fam")
    _, tnsnp = H.read_bim(TD + "test_chr1.bim")
    out.update(dict(test_bed=H.read_bed(TD + "test_chr1.bed", tnsnp, tn), test_n_total=tn,
                    test_bim_txt=open(TD + "test_chr1.bim").read(), test_indicator=np.array([1] * 20 + [0] * 83, np.int32)))
    out.update(dict(bed=bed, n_ref=n_ref, n_obs=n_obs, sigma_s=sigma_s, ref_maf=maf,
                    bim_txt=open(TD + "ref_chr1.bim").read(), fam_lines=n_ref, summary_txt="\n".join(lines) + "\n",
                    l_txt=l_txt, s_txt=s_txt, block_txt=blk_txt,
                    lmm_off=off, lmm_pos=pos, lmm_z=z, lmm_beta=b_lmm,
                    s_off=s_off, s_pos=s_pos, s_z=s_z, l_off=l_off, l_pos=l_pos, l_z=l_z, beta_s=b_s, beta_l=b_l))
    np.savez_compressed(os.path.join(GOLD, "c1_testdat.npz"), **out)
    print("c1: LMM m=%d, DBSLMM m_s=%d m_l=%d" % (pos.size, s_pos.size, l_pos.size))


def ragged():
    sizes = [150, 0, 1, 9, 64, 70, 131]
    w = synth.make_workload(20240001, sizes, 403, missing_rate=0.02, frac_large=0.03)
    n_obs, sigma_s = 50_000, 0.4 / 20_000
    with R.BedFile(w["bed"]) as bf:
        b_s, b_l = R.est_path(bf.path, 403, n_obs, sigma_s, w["s_off"], w["s_pos"], w["s_z"], w["l_off"], w["l_pos"], w["l_z"], threads=4)
        off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
        zz = np.zeros(off[-1]); zz[w["s_pos"]] = w["s_z"]; zz[w["l_pos"]] = w["l_z"]
        b_lmm, _ = R.est_path(bf.path, 403, n_obs, sigma_s, off, np.arange(off[-1], dtype=np.int32), zz, threads=4)
        maf = np.array([R.read_snp_im(bf.path, i, 403)[1] for i in range(off[-1])])
    np.savez_compressed(os.path.join(GOLD, "synth_ragged.npz"), bed=w["bed"], G=w["G"], n_ref=403, n_obs=n_obs, sigma_s=sigma_s,
                        sizes=np.asarray(sizes), s_off=w["s_off"], s_pos=w["s_pos"], s_z=w["s_z"], l_off=w["l_off"],
                        l_pos=w["l_pos"], l_z=w["l_z"], beta_s=b_s, beta_l=b_l, lmm_off=off, lmm_z=zz, lmm_beta=b_lmm, ref_maf=maf)
    print("ragged: large per block", np.diff(w["l_off"]))


def synth_cli():
    """A small synthetic chromosome as real PLINK / GEMMA text files, run through the UNMODIFIED reference CLI
    including the fork's variance side channel (test_dat itself has SNPs that are monomorphic in the test
    panel, which turns its whole variance column into NaN in the reference and here alike)."""
    rng = np.random.default_rng(20240002)
    sizes = [70, 110, 45]
    n_ref, n_tt = 300, 90
    w = synth.make_workload(20240002, sizes, n_ref, missing_rate=0.0, frac_large=0.03)
    m = int(sum(sizes))
    Gt = synth.make_genotypes(rng, [m], n_tt, missing_rate=0.01)
    tbed = synth.pack_bed(Gt)
    ind = (rng.random(n_tt) < 0.6).astype(np.int32)
    starts = np.array([1000, 500000, 900000]); ends = np.array([500000, 900000, 2000000])
    ps = np.concatenate([np.sort(rng.choice(np.arange(starts[b] + 1, ends[b] - 1), size=sizes[b], replace=False)) for b in range(3)])
    z = np.zeros(m); z[w["s_pos"]] = w["s_z"]; z[w["l_pos"]] = w["l_z"]
    S = np.where(w["G"] < 0, 0, w["G"]).sum(1); af = S / (2.0 * n_ref)
    bim = "".join(f"1\trs{j}\t0\t{ps[j]}\tA\tG\n" for j in range(m))
    fam = "".join(f"f{i} i{i} 0 0 0 -9\n" for i in range(n_ref))
    tfam = "".join(f"t{i} i{i} 0 0 0 -9\n" for i in range(n_tt))
    line = lambda j: f"1\trs{j}\t{ps[j]}\t0\t5000\tA\tG\t{af[j]:.6f}\t{z[j] * 0.01:.10e}\t1.0000000000e-02\t1.0e-03"
    large = sorted(w["l_pos"].tolist()); ls = set(large)
    l_txt = "\n".join(line(j) for j in large) + "\n"
    s_txt = "\n".join(line(j) for j in range(m) if j not in ls) + "\n"
    blk = "".join(f"chr1\t{starts[b]}\t{ends[b]}\n" for b in range(3))
    out = dict(bed=w["bed"], n_ref=n_ref, bim_txt=bim, fam_txt=fam, test_bed=tbed, test_fam_txt=tfam, test_n_total=n_tt,
               test_indicator=ind, l_txt=l_txt, s_txt=s_txt, block_txt=blk, sizes=np.asarray(sizes))
    with tempfile.TemporaryDirectory() as td:
        for name, txt in (("ref.bim", bim), ("ref.fam", fam), ("test.bim", bim), ("test.fam", tfam), ("l.txt", l_txt), ("s.txt", s_txt), ("blocks.bed", blk)):
            open(os.path.join(td, name), "w").write(txt)
        R.write_bed(w["bed"], os.path.join(td, "ref.bed"))
        R.write_bed(tbed, os.path.join(td, "test.bed"))
        open(os.path.join(td, "ind.txt"), "w").write("\n".join(str(int(x)) for x in ind) + "\n")
        cmd = [R.CLI, "-s", td + "/s.txt", "-l", td + "/l.txt", "-r", td + "/ref", "-n", "5000", "-nsnp", "2000", "-mafMax", "0.2",
               "-b", td + "/blocks.bed", "-h", "0.4", "-t", "2", "-eff", td + "/out", "-test_indicator_file", td + "/ind.txt", "-dat_str", td + "/test"]
        subprocess.run(cmd, cwd=td, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        out["cli_txt"] = open(td + "/out.txt").read()
        out["cli_badsnps"] = open(td + "/out.badsnps").read()
        vt = open(td + "/variance.txt").read().split("\n")
        nr, nc = (int(x) for x in vt[1].split())
        out["cli_variance"] = np.array([[float(x) for x in ln.split()] for ln in vt[2:2 + nr]]).reshape(nr, nc)
    assert np.isfinite(out["cli_variance"]).all()
    np.savez_compressed(os.path.join(GOLD, "synth_cli.npz"), **out)
    print("synth_cli: variance", out["cli_variance"].shape, "lines", len(out["cli_txt"].strip().split("\n")))


def valid_synth():
    """The reference's `valid` binary (oracle/_ref/valid_ref = scr/main_valid.cpp + scr/validate.cpp, unmodified, over the
    shim) on the synthetic chromosome of synth_cli: -d is the reference dbslmm CLI's own output (<eff>.txt), -s an
    external summary file (snp a1 maf z) that drops some SNPs, flips some alleles (z2 changes sign, dtpr.cpp:424-428)
    and gives a few SNPs a MAF far from the panel's (filtered by -mafMax, dtpr.cpp:440)."""
    g = np.load(os.path.join(GOLD, "synth_cli.npz"))
    rng = np.random.default_rng(20240005)
    n_ref = int(g["n_ref"])
    bed = g["bed"]
    G = O_counts(bed, n_ref)
    af = G.sum(1) / (2.0 * n_ref)
    maf = np.minimum(af, 1 - af)
    rows = [ln.split(" ") for ln in str(g["cli_txt"]).strip().split("\n")]
    lines = []
    for r in rows:
        j = int(r[0][2:])
        u = rng.random()
        if u < 0.08:
            continue                                              # not in the external study
        a1 = "G" if u < 0.2 else r[1]                             # allele discrepancy -> sign flip
        mf = maf[j] + (0.35 if u > 0.95 else 0.0)                 # a few fail the MAF filter
        lines.append(f"{r[0]} {a1} {mf:.6f} {rng.standard_normal():.8f}")
    lines.append("rs_not_in_panel A 0.200000 1.50000000")
    ext_txt = "\n".join(lines) + "\n"
    out = dict(ext_txt=ext_txt)
    with tempfile.TemporaryDirectory() as td:
        for name, key in (("ref.bim", "bim_txt"), ("ref.fam", "fam_txt"), ("blocks.bed", "block_txt"), ("dbslmm.txt", "cli_txt")):
            open(os.path.join(td, name), "w").write(str(g[key]))
        open(os.path.join(td, "ext.txt"), "w").write(ext_txt)
        R.write_bed(bed, os.path.join(td, "ref.bed"))
        for tag, mm in (("c", "0.2"), ("u", "1")):
            cmd = [os.path.join(ROOT, "oracle", "_ref", "valid_ref"), "-d", td + "/dbslmm.txt", "-s", td + "/ext.txt", "-r", td + "/ref",
                   "-mafMax", mm, "-b", td + "/blocks.bed", "-r2", td + "/r2" + tag]
            subprocess.run(cmd, cwd=td, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            out["r2_" + tag] = open(td + "/r2" + tag + ".txt").read()
    np.savez_compressed(os.path.join(GOLD, "valid_synth.npz"), **out)
    print("valid_synth:", out["r2_c"].strip().replace("\n", " | "))


def O_counts(bed, n_ref):
    """allele counts 0/1/2 per SNP row (missing -> 0) from a packed .bed payload"""
    b = np.asarray(bed, np.uint8)
    codes = np.stack([(b >> s) & 3 for s in (0, 2, 4, 6)], axis=-1).reshape(b.shape[0], -1)[:, :n_ref]
    return np.where(codes == 0, 2, np.where(codes == 2, 1, 0))


if __name__ == "__main__":
    if "valid" in sys.argv[1:]:
        valid_synth()
        raise SystemExit(0)
    if not R.available():
        raise SystemExit("oracle/_ref missing: run oracle/build_ref.sh first")
    os.makedirs(GOLD, exist_ok=True)
    c1()
    ragged()
    synth_cli()
    valid_synth()
