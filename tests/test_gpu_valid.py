"""the reference's external-validation tool `valid` (scr/validate.cpp) on the GPU: the quadratic form
deno_b = z1' (X'X/n) z1 comes from the fit's decoder + exact integer Gram (tau = 1) and a quadratic-form kernel
(fit_args.quadform_out); build/valid is the drop-in command line."""
import os
import subprocess

import numpy as np
import pytest

from dbslmm_b200 import _abi, synth
from oracle import oracle as O
from oracle import refharness as R

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
VALID = os.path.join(ROOT, "build", "valid")


@pytest.mark.parametrize("miss", [0.0, 0.02])
@pytest.mark.parametrize("tau", [1.0, 0.8])
def test_quadform_matches_the_reference_float_path(engine, miss, tau):
    sizes = [300, 0, 1, 7, 64, 129, 700, 1500]
    w = synth.make_workload(77 + int(100 * miss), sizes, 403, missing_rate=miss, frac_large=0.0)
    engine.load_bed(w["bed"], 403)
    z = np.random.default_rng(9).standard_normal(w["s_pos"].size)
    got = engine.quadform(w["s_off"], w["s_pos"], z, tau=tau)
    for b in range(len(sizes)):
        lo, hi = w["s_off"][b], w["s_off"][b + 1]
        zb = z[lo:hi]
        want = float(zb @ O.sigma(w["bed"], 403, w["s_pos"][lo:hi], tau=tau) @ zb) if hi > lo else 0.0
        assert abs(got[b] - want) <= 1e-11 * max(abs(want), 1.0), (b, got[b], want)
    # the handle still fits afterwards (the two paths share the plan machinery)
    r = engine.fit(w["s_off"], w["s_pos"], z, sigma_s=[1e-4], n_obs=5000)
    assert r["n_bad"] == 0


def test_valid_cli_matches_the_reference_binary(tmp_path):
    g = np.load(os.path.join(GOLD, "synth_cli.npz"))
    v = np.load(os.path.join(GOLD, "valid_synth.npz"))
    for name, key in (("ref.bim", "bim_txt"), ("ref.fam", "fam_txt"), ("blocks.bed", "block_txt"), ("dbslmm.txt", "cli_txt")):
        (tmp_path / name).write_text(str(g[key]))
    (tmp_path / "ext.txt").write_text(str(v["ext_txt"]))
    R.write_bed(g["bed"], str(tmp_path / "ref.bed"))
    for tag, mm, key in (("c", "0.2", "r2_c"), ("u", "1", "r2_u")):
        cmd = [VALID, "-d", str(tmp_path / "dbslmm.txt"), "-s", str(tmp_path / "ext.txt"), "-r", str(tmp_path / "ref"), "-mafMax", mm,
               "-b", str(tmp_path / "blocks.bed"), "-r2", str(tmp_path / ("r2" + tag))]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert "SNPs intersect." in r.stdout and "blocks for the chromesome." in r.stdout
        got = np.array([[float(x) for x in ln.split()] for ln in (tmp_path / ("r2" + tag + ".txt")).read_text().strip().split("\n")])
        ref = np.array([[float(x) for x in ln.split()] for ln in str(v[key]).strip().split("\n")])
        assert got.shape == ref.shape
        assert np.abs(got - ref).max(axis=0)[0] <= 5e-6 * np.abs(ref[:, 0]).max()    # 6 printed digits
        assert np.abs(got - ref).max(axis=0)[1] <= 5e-6 * np.abs(ref[:, 1]).max()
    # banner / help / missing-file behaviour of main_valid.cpp:37-44 and validate.cpp:134-168
    assert subprocess.run([VALID], capture_output=True, text=True).returncode == 0
    assert "-r2" in subprocess.run([VALID, "-h"], capture_output=True, text=True).stdout
    r = subprocess.run([VALID, "-d", str(tmp_path / "nope.txt"), "-s", str(tmp_path / "ext.txt"), "-r", str(tmp_path / "ref"),
                        "-b", str(tmp_path / "blocks.bed"), "-r2", str(tmp_path / "x")], capture_output=True, text=True)
    assert r.returncode == 1 and "dose not exist" in r.stderr
