"""The drop-in `dbslmm` command line: option handling on CPU, file-level parity against the
UNMODIFIED reference CLI's golden output on the GPU."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "build", "dbslmm")
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _write_fixture(tmp_path):
    d = np.load(os.path.join(GOLD, "c1_testdat.npz"))
    (tmp_path / "ref.bim").write_text(str(d["bim_txt"]))
    (tmp_path / "ref.fam").write_text("f i 0 0 0 -9\n" * int(d["fam_lines"]))
    with open(tmp_path / "ref.bed", "wb") as f:
        f.write(bytes([0x6C, 0x1B, 0x01]) + d["bed"].tobytes())
    (tmp_path / "summ.txt").write_text(str(d["summary_txt"]))
    (tmp_path / "l.txt").write_text(str(d["l_txt"]))
    (tmp_path / "s.txt").write_text(str(d["s_txt"]))
    (tmp_path / "blocks.bed").write_text(str(d["block_txt"]))
    return d


@pytest.mark.skipif(not os.path.exists(CLI), reason="CLI not built")
def test_banner_help_and_checks(tmp_path):
    out = subprocess.run([CLI], capture_output=True, text=True)
    assert out.returncode == 0 and "Deterministic Bayesian Sparse Linear Mixed Model" in out.stdout
    out = subprocess.run([CLI, "-h"], capture_output=True, text=True)
    assert out.returncode == 0 and "-mafMax" in out.stdout and "-eff" in out.stdout
    _write_fixture(tmp_path)
    base = ["-r", str(tmp_path / "ref"), "-b", str(tmp_path / "blocks.bed"), "-n", "2400", "-nsnp", "996", "-eff", str(tmp_path / "o")]
    r = subprocess.run([CLI, "-s", str(tmp_path / "nope.txt"), "-h", "0.5", "-t", "1"] + base, capture_output=True, text=True)
    assert r.returncode == 1 and "dose not exist" in r.stderr
    r = subprocess.run([CLI, "-s", str(tmp_path / "summ.txt"), "-h", "1.5", "-t", "1"] + base, capture_output=True, text=True)
    assert r.returncode == 1 and "-h is not correct" in r.stderr
    r = subprocess.run([CLI, "-s", str(tmp_path / "summ.txt"), "-h", "0.5", "-t", "101"] + base, capture_output=True, text=True)
    assert r.returncode == 1 and "-t is not correct" in r.stderr


def _parse(txt):
    rows = [ln.split(" ") for ln in txt.strip().split("\n")]
    return [(r[0], r[1], float(r[2]), float(r[3]), int(r[4])) for r in rows]


@pytest.mark.gpu
def test_cli_matches_reference_cli_output(tmp_path):
    d = _write_fixture(tmp_path)
    cmd = [CLI, "-s", str(tmp_path / "s.txt"), "-l", str(tmp_path / "l.txt"), "-r", str(tmp_path / "ref"), "-n", "2400",
           "-nsnp", "996", "-mafMax", "0.2", "-b", str(tmp_path / "blocks.bed"), "-h", "0.5", "-t", "1",
           "-eff", str(tmp_path / "out"), "--dump-beta-bin", str(tmp_path / "beta.bin")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Fitting time:" in r.stdout
    got = _parse((tmp_path / "out.txt").read_text())
    ref = _parse(str(d["cli_dbslmm_txt"]))
    assert [(g[0], g[1], g[4]) for g in got] == [(x[0], x[1], x[4]) for x in ref]       # same SNPs, alleles, flags, order
    gb, rb = np.array([g[2] for g in got]), np.array([x[2] for x in ref])
    gn, rn = np.array([g[3] for g in got]), np.array([x[3] for x in ref])
    # 6 significant digits in the text format; the reference's own PCG truncation is ~3e-8
    assert np.abs(gb - rb).max() / np.abs(rb).max() < 5e-6 and np.abs(gn - rn).max() / np.abs(rn).max() < 5e-6
    assert (tmp_path / "out.badsnps").read_text() == str(d["cli_dbslmm_badsnps"])
    raw = (tmp_path / "beta.bin").read_bytes()
    nf, nl, ns = struct.unpack("qqq", raw[:24])
    beta = np.frombuffer(raw[24:], dtype=np.float64)
    assert (nf, nl, ns) == (1, d["beta_l"].size, d["beta_s"].size)
    from parity_bars import bar
    assert np.abs(beta[:nl] - d["beta_l"]).max() / np.abs(d["beta_l"]).max() <= bar("c1_testdat", "beta_l")
    assert np.abs(beta[nl:] - d["beta_s"]).max() / np.abs(d["beta_s"]).max() <= bar("c1_testdat", "beta_s")


@pytest.mark.gpu
def test_cli_lmm_mode_and_folds(tmp_path):
    d = _write_fixture(tmp_path)
    cmd = [CLI, "-s", str(tmp_path / "summ.txt"), "-r", str(tmp_path / "ref"), "-n", "2400", "-nsnp", "996", "-mafMax", "0.2",
           "-b", str(tmp_path / "blocks.bed"), "-h", "0.5", "-t", "1", "-eff", str(tmp_path / "lmm"),
           "--dump-beta-bin", str(tmp_path / "b.bin"), "--h2-folds", "0.8,1.0,1.2"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = (tmp_path / "b.bin").read_bytes()
    nf, nl, ns = struct.unpack("qqq", raw[:24])
    assert (nf, nl, ns) == (3, 0, 716)
    beta = np.frombuffer(raw[24:], dtype=np.float64).reshape(3, 716)
    from parity_bars import bar
    assert np.abs(beta[1] - d["lmm_beta"]).max() / np.abs(d["lmm_beta"]).max() <= bar("c1_testdat", "lmm_beta")
    for f in range(3):
        assert len((tmp_path / f"lmm_f{f}.txt").read_text().strip().split("\n")) == 716


@pytest.mark.gpu
def test_cli_manifest_two_chromosomes(tmp_path):
    """several chromosomes in one process / one GPU plan give the same per-chromosome files as separate runs."""
    _write_fixture(tmp_path)
    common = ["-n", "2400", "-nsnp", "996", "-mafMax", "0.2", "-h", "0.5", "-t", "1"]
    ref, blk = str(tmp_path / "ref"), str(tmp_path / "blocks.bed")
    # separate runs
    r = subprocess.run([CLI, "-s", str(tmp_path / "s.txt"), "-l", str(tmp_path / "l.txt"), "-r", ref, "-b", blk, "-eff", str(tmp_path / "a")] + common,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([CLI, "-s", str(tmp_path / "summ.txt"), "-r", ref, "-b", blk, "-eff", str(tmp_path / "b")] + common,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # one manifest run
    (tmp_path / "manifest.tsv").write_text(
        f"{tmp_path / 's.txt'}\t{tmp_path / 'l.txt'}\t{ref}\t{blk}\t{tmp_path / 'ma'}\n"
        f"{tmp_path / 'summ.txt'}\t-\t{ref}\t{blk}\t{tmp_path / 'mb'}\n")
    r = subprocess.run([CLI, "--manifest", str(tmp_path / "manifest.tsv")] + common, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "2 chromosomes in one run" in r.stdout
    for single, multi in (("a", "ma"), ("b", "mb")):
        x, y = _parse((tmp_path / f"{single}.txt").read_text()), _parse((tmp_path / f"{multi}.txt").read_text())
        assert [(g[0], g[1], g[4]) for g in x] == [(g[0], g[1], g[4]) for g in y]
        vx, vy = np.array([g[2] for g in x]), np.array([g[2] for g in y])
        assert np.abs(vx - vy).max() <= 2e-6 * np.abs(vx).max()
        assert (tmp_path / f"{single}.badsnps").read_text() == (tmp_path / f"{multi}.badsnps").read_text()


def _read_variance(path):
    lines = path.read_text().split("\n")
    assert lines[0] == "ARMA_MAT_TXT_FN008"
    nr, nc = (int(x) for x in lines[1].split())
    return np.array([[float(x) for x in ln.split()] for ln in lines[2:2 + nr]]).reshape(nr, nc)


@pytest.mark.gpu
def test_cli_variance_txt_c1_nan_pattern(tmp_path):
    """test_dat has SNPs that are monomorphic in the test panel: the reference's variance column of block 0 is NaN
    (stddev 0 in nomalizeVec, dtpr.cpp:378) and so is ours; the 132 empty blocks are 0 in both."""
    d = _write_fixture(tmp_path)
    (tmp_path / "test.bim").write_text(str(d["test_bim_txt"]))
    with open(tmp_path / "test.bed", "wb") as f:
        f.write(bytes([0x6C, 0x1B, 0x01]) + d["test_bed"].tobytes())
    (tmp_path / "ind.txt").write_text("\n".join(str(int(x)) for x in d["test_indicator"]) + "\n")
    cmd = [CLI, "-s", str(tmp_path / "s.txt"), "-l", str(tmp_path / "l.txt"), "-r", str(tmp_path / "ref"), "-n", "2400",
           "-nsnp", "996", "-mafMax", "0.2", "-b", str(tmp_path / "blocks.bed"), "-h", "0.5", "-t", "1",
           "-eff", str(tmp_path / "out"), "-test_indicator_file", str(tmp_path / "ind.txt"), "-dat_str", str(tmp_path / "test")]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stderr
    got, ref = _read_variance(tmp_path / "variance.txt"), d["cli_dbslmm_variance"]
    assert got.shape == ref.shape == (20, 133)
    assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.isnan(got[:, 0]).all() and not got[:, 1:].any()
    got_txt, ref_txt = _parse((tmp_path / "out.txt").read_text()), _parse(str(d["cli_dbslmm_txt"]))
    assert [(g[0], g[1], g[4]) for g in got_txt] == [(x[0], x[1], x[4]) for x in ref_txt]     # betas unaffected


@pytest.mark.gpu
def test_cli_variance_txt_matches_reference_cli(tmp_path):
    """Synthetic chromosome as real PLINK/GEMMA files: <eff>.txt, <eff>.badsnps and variance.txt against what the
    UNMODIFIED reference CLI wrote for the same files (tools/make_golden.py::synth_cli)."""
    d = np.load(os.path.join(GOLD, "synth_cli.npz"))
    for name, key in (("ref.bim", "bim_txt"), ("ref.fam", "fam_txt"), ("test.bim", "bim_txt"), ("test.fam", "test_fam_txt"),
                      ("l.txt", "l_txt"), ("s.txt", "s_txt"), ("blocks.bed", "block_txt")):
        (tmp_path / name).write_text(str(d[key]))
    for name, key in (("ref.bed", "bed"), ("test.bed", "test_bed")):
        with open(tmp_path / name, "wb") as f:
            f.write(bytes([0x6C, 0x1B, 0x01]) + d[key].tobytes())
    (tmp_path / "ind.txt").write_text("\n".join(str(int(x)) for x in d["test_indicator"]) + "\n")
    cmd = [CLI, "-s", str(tmp_path / "s.txt"), "-l", str(tmp_path / "l.txt"), "-r", str(tmp_path / "ref"), "-n", "5000",
           "-nsnp", "2000", "-mafMax", "0.2", "-b", str(tmp_path / "blocks.bed"), "-h", "0.4", "-t", "2",
           "-eff", str(tmp_path / "out"), "-test_indicator_file", str(tmp_path / "ind.txt"), "-dat_str", str(tmp_path / "test")]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stderr
    got, ref = _read_variance(tmp_path / "variance.txt"), d["cli_variance"]
    assert got.shape == ref.shape and np.isfinite(got).all()
    assert np.abs(got - ref).max() <= 1e-9 * np.abs(ref).max()
    got_txt, ref_txt = _parse((tmp_path / "out.txt").read_text()), _parse(str(d["cli_txt"]))
    assert [(g[0], g[1], g[4]) for g in got_txt] == [(x[0], x[1], x[4]) for x in ref_txt]
    gb, rb = np.array([g[2] for g in got_txt]), np.array([x[2] for x in ref_txt])
    assert np.abs(gb - rb).max() <= 5e-6 * np.abs(rb).max()
    assert (tmp_path / "out.badsnps").read_text() == str(d["cli_badsnps"])


@pytest.mark.gpu
def test_cli_without_maf_constraint_streams_the_panel(tmp_path):
    """mafMax == 1 (the reference's 'do not consider the difference' branch, dbslmm.cpp:238-241): nothing needs the
    panel before the fit, so the CLI hands it to dbslmm_b200_fit (fit_args.bed, upload overlapped with the fit).
    The result must equal the upload-then-fit path bit for bit."""
    _write_fixture(tmp_path)
    outs = []
    for tag, env in (("stream", {}), ("plain", {"DBSLMM_B200_STREAM_BED": "0"})):
        cmd = [CLI, "-s", str(tmp_path / "s.txt"), "-l", str(tmp_path / "l.txt"), "-r", str(tmp_path / "ref"), "-n", "2400",
               "-nsnp", "996", "-mafMax", "1", "-b", str(tmp_path / "blocks.bed"), "-h", "0.5", "-t", "1",
               "-eff", str(tmp_path / tag), "--dump-beta-bin", str(tmp_path / (tag + ".bin"))]
        r = subprocess.run(cmd, capture_output=True, text=True, env={**os.environ, **env})
        assert r.returncode == 0, r.stderr
        outs.append(((tmp_path / (tag + ".txt")).read_text(), (tmp_path / (tag + ".bin")).read_bytes()))
    assert outs[0][0] == outs[1][0] and outs[0][1] == outs[1][1]
    assert len(outs[0][0].strip().split("\n")) > 700


@pytest.mark.gpu
def test_cli_two_gpus_match_one_gpu(tmp_path):
    """--gpus 2: blocks sharded by the library's cost model, every GPU gets only its shard of the panel (handed to
    dbslmm_b200_fit with the call, uploaded in batches), betas gathered on the host: same files as one GPU."""
    from dbslmm_b200 import _abi
    if _abi.load().dbslmm_b200_device_count() < 2:
        pytest.skip("needs two GPUs")
    _write_fixture(tmp_path)
    (tmp_path / "manifest.tsv").write_text(
        f"{tmp_path / 's.txt'}\t{tmp_path / 'l.txt'}\t{tmp_path / 'ref'}\t{tmp_path / 'blocks.bed'}\t{tmp_path / 'x'}\n"
        f"{tmp_path / 'summ.txt'}\t-\t{tmp_path / 'ref'}\t{tmp_path / 'blocks.bed'}\t{tmp_path / 'y'}\n")
    common = ["-n", "2400", "-nsnp", "996", "-mafMax", "0.2", "-h", "0.5", "-t", "1"]
    outs = {}
    for g in ("1", "2"):
        r = subprocess.run([CLI, "--manifest", str(tmp_path / "manifest.tsv"), "--gpus", g] + common, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        outs[g] = [_parse((tmp_path / f"{t}.txt").read_text()) for t in ("x", "y")]
    for a, b in zip(outs["1"], outs["2"]):
        assert [(g[0], g[1], g[4]) for g in a] == [(g[0], g[1], g[4]) for g in b]
        va, vb = np.array([g[2] for g in a]), np.array([g[2] for g in b])
        assert np.abs(va - vb).max() <= 2e-6 * np.abs(va).max()


@pytest.mark.gpu
def test_cli_manifest_three_real_sized_chromosomes(tmp_path):
    """Chromosomes 20-22 at the genome-wide SNP density (n_ref = 2,000, ~60k SNPs, 70 EUR LD blocks) as REAL files
    (tools/cli_genome_wide.py writes the PLINK / GEMMA / block files): one cold `dbslmm --manifest` process against the
    same problem fitted through the C ABI, with and without the MAF pre-pass, on one GPU and -- when the box has two --
    with the blocks sharded over two GPUs that each take the whole host panel (FLAG_PANEL_SUBSET)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import cli_genome_wide as G
    from dbslmm_b200 import _abi
    mf, jobs, w = G.write_genome(str(tmp_path / "gw"), "chr20_22", 20240003)
    nsnp, n_obs, n_ref = int(w["n_snp"]), int(w["n_obs"]), int(w["n_ref"])
    eng = _abi.Engine(0)
    try:
        eng.load_bed(w["bed"], n_ref)
        r = eng.fit(w["s_off"], w["s_pos"], w["z"][w["s_pos"]], w["l_off"], w["l_pos"], w["z"][w["l_pos"]],
                    sigma_s=[0.5 / nsnp], n_obs=n_obs)
        n_dev = _abi.load().dbslmm_b200_device_count()
    finally:
        eng.close()
    assert r["n_bad"] == 0
    runs = [("1", "1"), ("0.2", "1")] + ([("1", "2")] if n_dev >= 2 else [])
    for maf, gpus in runs:
        out = tmp_path / f"beta_{maf}_{gpus}.bin"
        cmd = [CLI, "--manifest", mf, "-n", str(n_obs), "-nsnp", str(nsnp), "-h", "0.5", "-t", "8", "-mafMax", maf, "--gpus", gpus,
               "--dump-beta-bin", str(out)]
        p = subprocess.run(cmd, capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        assert "Fitting time:" in p.stdout and "[timing] panel files" in p.stdout
        raw = out.read_bytes()
        nf, tl, ts = struct.unpack("<qqq", raw[:24])
        v = np.frombuffer(raw[24:], np.float64)
        assert (nf, tl, ts) == (1, w["l_pos"].size, w["s_pos"].size)
        bl, bs = v[:tl], v[tl:]
        # the text round trip of the z-scores (beta and se printed with 11 / 7 digits) bounds the agreement
        assert np.abs(bs - r["beta_s"][0]).max() <= 1e-9 * np.abs(r["beta_s"][0]).max()
        assert np.abs(bl - r["beta_l"][0]).max() <= 1e-9 * np.abs(r["beta_l"][0]).max()
        for j in jobs:
            rows = open(j["pre"] + "_out.txt").read().strip().split("\n")
            assert len(rows) == j["n_snp"]                       # every SNP matched, large effects first
            assert os.path.getsize(j["pre"] + "_out.badsnps") == 0
