"""GPU: polygenic scoring over a validation panel (BASELINE config 4) against a numpy
restatement of `plink --score <eff>.txt 1 2 4 sum` (DBSLMM_script.sh:87): score_i = sum_j beta_j * dosage_ij,
dosage = copies of the scored allele, missing calls replaced by the SNP's mean dosage.  PLINK is external to
the reference, so this restatement -- not a reference run -- is the pin ("parity unpinned" for this row)."""
import numpy as np
import pytest

from dbslmm_b200 import synth

pytestmark = pytest.mark.gpu


def plink_score_sum(G, beta, flip=None):
    """G int8 [n_snp, n] copies of A1 (-1 missing); beta [n_folds, n_snp]."""
    D = G.astype(np.float64)
    miss = G < 0
    mu = np.where(miss, 0.0, D).sum(1) / (~miss).sum(1)
    if flip is not None:
        D = np.where(flip[:, None] != 0, 2.0 - D, D)
        mu = np.where(flip != 0, 2.0 - mu, mu)
    D = np.where(miss, mu[:, None], D)
    return beta @ D


@pytest.mark.parametrize("n_val,n_snp,n_folds,miss", [(1003, 700, 3, 0.01), (257, 130, 1, 0.0), (4096, 300, 5, 0.02)])
def test_prs_matches_plink_semantics(engine, n_val, n_snp, n_folds, miss):
    rng = np.random.default_rng(n_val)
    G = synth.make_genotypes(rng, [n_snp], n_val, missing_rate=miss)
    bed = synth.pack_bed(G)
    sel = np.sort(rng.choice(n_snp, size=n_snp - 17, replace=False)).astype(np.int32)
    beta = rng.standard_normal((n_folds, sel.size)) * 1e-2
    flip = (rng.random(sel.size) < 0.3).astype(np.uint8)
    got, ms = engine.score(bed, n_val, sel, beta, flip)
    exp = plink_score_sum(G[sel], beta, flip)
    assert got.shape == exp.shape
    assert np.abs(got - exp).max() <= 1e-12 * max(1.0, np.abs(exp).max())
    got2, _ = engine.score(bed, n_val, sel, beta, None)
    assert np.abs(got2 - plink_score_sum(G[sel], beta, None)).max() <= 1e-12 * max(1.0, np.abs(exp).max())
    assert ms >= 0.0


def test_prs_linearity_and_folds_after_fit(engine):
    """Config 4 end to end at test size: one Gram, three ridge folds, then all folds scored in one pass."""
    w = synth.make_workload(404, [150, 90], 500, frac_large=0.02)
    engine.load_bed(w["bed"], 500)
    sig = np.array([0.8, 1.0, 1.2]) * 0.5 / 240
    r = engine.fit(w["s_off"], w["s_pos"], w["s_z"], w["l_off"], w["l_pos"], w["l_z"], sigma_s=sig, n_obs=20000)
    beta = np.zeros((3, 240))
    beta[:, w["s_pos"]] = r["beta_s"]
    beta[:, w["l_pos"]] = r["beta_l"]
    rng = np.random.default_rng(5)
    Gv = synth.make_genotypes(rng, [240], 2000, missing_rate=0.005)
    bedv = synth.pack_bed(Gv)
    pos = np.arange(240, dtype=np.int32)
    s, _ = engine.score(bedv, 2000, pos, beta)
    assert np.abs(s - plink_score_sum(Gv, beta)).max() <= 1e-12 * np.abs(s).max()
    s2, _ = engine.score(bedv, 2000, pos, 2.0 * beta[:1] - beta[2:3])
    assert np.abs(s2[0] - (2.0 * s[0] - s[2])).max() <= 1e-12 * np.abs(s).max()      # linear in beta


def test_prs_prefetched_panel_overlaps_a_fit(engine):
    """score_prefetch announces the validation panel; the next fit issues its upload behind its own panel copies; score(None)
    then finds it on the device.  Same scores as the direct call; a second score(None) reuses the resident panel."""
    w = synth.make_workload(405, [120, 60], 400, frac_large=0.0)
    rng = np.random.default_rng(6)
    Gv = synth.make_genotypes(rng, [180], 1500, missing_rate=0.01)
    bedv = synth.pack_bed(Gv)
    pos = np.arange(180, dtype=np.int32)
    engine.score_prefetch(bedv, 1500)
    r = engine.fit(w["s_off"], w["s_pos"], w["s_z"], sigma_s=[1e-3, 2e-3], n_obs=20000, bed=w["bed"], n_ref=400)
    beta = np.zeros((2, 180)); beta[:, w["s_pos"]] = r["beta_s"]
    s1, _ = engine.score(None, 1500, pos, beta)
    exp = plink_score_sum(Gv, beta)
    assert np.abs(s1 - exp).max() <= 1e-12 * np.abs(exp).max()
    s2, _ = engine.score(None, 1500, pos, beta[:1])
    assert np.abs(s2[0] - exp[0]).max() <= 1e-12 * np.abs(exp).max()
    engine.score_prefetch(bedv, 1500)                           # announced, but no fit in between: score uploads it itself
    s3, _ = engine.score(None, 1500, pos, beta)
    assert np.array_equal(s3, s1)


def test_prs_argument_errors(engine):
    from dbslmm_b200 import _abi
    G = synth.make_genotypes(np.random.default_rng(1), [20], 100)
    with pytest.raises(_abi.EngineError):
        engine.score(synth.pack_bed(G), 100, np.array([25], np.int32), np.ones((1, 1)))
