"""GPU: the fork's asymptotic-variance side channel against the oracle's dense restatement of
calc_nt_by_nt_matrix, which tests/test_oracle.py pins to the UNMODIFIED reference build (7e-16 on this input)."""
import numpy as np
import pytest

from dbslmm_b200 import _abi, synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _case(seed, sizes, n_ref, n_tt, frac_large, miss):
    w = synth.make_workload(seed, sizes, n_ref, missing_rate=miss, frac_large=frac_large)
    rng = np.random.default_rng(seed + 1)
    Gt = synth.make_genotypes(rng, [int(sum(sizes))], n_tt, missing_rate=miss)
    # the test panel lists the same SNPs in a different order: row of SNP p in the test .bed is perm[p]
    perm = rng.permutation(int(sum(sizes))).astype(np.int32)
    tbed = np.zeros_like(synth.pack_bed(Gt))
    tbed[perm] = synth.pack_bed(Gt)
    ind = (rng.random(n_tt) < 0.7).astype(np.int32)
    return w, tbed, perm, ind


@pytest.mark.parametrize("mode", ["dbslmm", "lmm"])
def test_variance_matches_oracle(engine, mode):
    sizes, n_ref, n_tt = [150, 0, 9, 70, 200], 400, 131
    w, tbed, perm, ind = _case(11, sizes, n_ref, n_tt, 0.03, 0.01)
    engine.load_bed(w["bed"], n_ref)
    sig, n_obs = 0.5 / 2000.0, 20_000
    if mode == "dbslmm":
        csr = (w["s_off"], w["s_pos"], w["s_z"], w["l_off"], w["l_pos"], w["l_z"])
        test = dict(bed=tbed, n_total=n_tt, indicator=ind, s_tpos=perm[w["s_pos"]], l_tpos=perm[w["l_pos"]])
    else:
        off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
        pos = np.arange(off[-1], dtype=np.int32)
        z = np.zeros(off[-1]); z[w["s_pos"]] = w["s_z"]; z[w["l_pos"]] = w["l_z"]
        csr = (off, pos, z)
        test = dict(bed=tbed, n_total=n_tt, indicator=ind, s_tpos=perm[pos], l_tpos=None)
    r = engine.fit(*csr, sigma_s=[sig], n_obs=n_obs, test=test)
    assert r["n_bad"] == 0
    var = r["variance"][0]
    assert var.shape == (len(sizes), int(ind.sum()))
    # betas are unaffected by the extra rows
    bs, bl, _, _ = O.est(w["bed"], n_ref, n_obs, sig, *csr, threads=4, mode=O.MODE_EXACT)
    assert np.abs(r["beta_s"][0] - bs).max() <= 1e-10 * np.abs(bs).max()
    for b, m in enumerate(sizes):
        if m == 0:
            assert not var[b].any()
            continue
        ps = csr[1][csr[0][b]:csr[0][b + 1]]
        pl = csr[4][csr[3][b]:csr[3][b + 1]] if mode == "dbslmm" else np.zeros(0, np.int32)
        exp = O.variance_block(w["bed"], n_ref, tbed, n_tt, ind, n_obs, sig, ps, perm[ps],
                               pl if pl.size else None, perm[pl] if pl.size else None)
        assert np.abs(var[b] - exp).max() <= 1e-10 * np.abs(exp).max(), (mode, b)


def test_variance_per_fold_and_errors(engine):
    w, tbed, perm, ind = _case(5, [90, 40], 300, 64, 0.0, 0.0)
    engine.load_bed(w["bed"], 300)
    csr = (w["s_off"], w["s_pos"], w["s_z"])
    test = dict(bed=tbed, n_total=64, indicator=ind, s_tpos=perm[w["s_pos"]], l_tpos=None)
    sig = np.array([1.0, 1.5]) * 1e-4
    r = engine.fit(*csr, sigma_s=sig, n_obs=9000, test=test)
    for f in range(2):
        exp = O.variance_block(w["bed"], 300, tbed, 64, ind, 9000, float(sig[f]), w["s_pos"][:90], perm[w["s_pos"][:90]])
        assert np.abs(r["variance"][f][0] - exp).max() <= 1e-10 * np.abs(exp).max()
    with pytest.raises(_abi.EngineError):
        engine.fit(*csr, sigma_s=[1e-4], n_obs=9000, test=test, solver=_abi.SOLVER_PCG)
