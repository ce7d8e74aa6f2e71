"""GPU parity tests: every call goes through the C ABI (ctypes) of libdbslmm_b200.so.

Bars (DESIGN.md "Parity"):
  decoder codes, per-SNP counts, integer Gram planes ......... bit-exact
  maf ........................................................ <= 1e-15 abs (one FP64 division)
  Sigma ...................................................... <= 1e-13 abs vs the reference float path
  beta, Cholesky solver vs `exact` oracle .................... <= 1e-10 of max|beta| per call
  beta vs the UNMODIFIED reference's golden vectors .......... <= 2 x the gap measured per fixture (tests/parity_bars.py):
       the reference stops PCG at an absolute residual of 1e-7, so ITS answer carries up to a few 1e-8 of truncation error
"""
import os

import numpy as np
import pytest

from dbslmm_b200 import _abi, synth
from oracle import oracle as O
from parity_bars import bar, record

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def relmax(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def block_pos(w, b):
    return np.concatenate([w["s_pos"][w["s_off"][b]:w["s_off"][b + 1]], w["l_pos"][w["l_off"][b]:w["l_off"][b + 1]]]).astype(np.int32)


CASES = {
    # name: (block sizes, n_ref, missing rate)
    "ragged_plain": ([300, 0, 1, 7, 8, 63, 64, 65, 127, 128, 129, 200], 400, 0.0),
    "ragged_missing": ([150, 70, 5, 130, 0, 257], 403, 0.02),
    "odd_pitch": ([90, 33], 125 * 4 - 3, 0.0),        # pitch 125 B: unaligned rows, 3 padding samples
    "wide_n": ([260, 140], 2000, 0.005),
    "big_missing": ([1500, 40, 700], 403, 0.02),      # split-K steps + the four-plane Gram + the plan rebuilt with missing flags
}


@pytest.fixture(scope="module", params=list(CASES))
def case(request, engine):
    sizes, n_ref, miss = CASES[request.param]
    w = synth.make_workload(1234 + len(sizes), sizes, n_ref, missing_rate=miss, frac_large=0.02)
    engine.load_bed(w["bed"], n_ref)
    return request.param, w


def test_snp_stats(case, engine):
    _, w = case
    maf, nn = engine.snp_stats()
    assert np.array_equal(nn, (w["G"] >= 0).sum(axis=1))
    assert np.abs(maf - O.snp_maf(w["bed"], w["bed"].shape[0], w["n_ref"])).max() <= 1e-15


def test_decode_gram_sigma(case, engine):
    name, w = case
    n_ref = w["n_ref"]
    r = engine.fit(w["s_off"], w["s_pos"], w["s_z"], w["l_off"], w["l_pos"], w["l_z"], sigma_s=[1e-4], n_obs=10_000,
                   flags=_abi.FLAG_KEEP_INT_GRAM)
    assert r["n_bad"] == 0
    Gz = np.where(w["G"] < 0, 0, w["G"]).astype(np.int8)
    n_pad = (n_ref + 127) // 128 * 128
    for b, m in enumerate(w["block_sizes"]):
        if m == 0:
            continue
        pos = block_pos(w, b)
        miss = bool((w["G"][pos] < 0).any())
        for j in sorted({0, m // 2, m - 1}):
            codes = engine.row_codes(b, j, n_pad)
            assert np.array_equal(codes[:n_ref], Gz[pos[j]]) and not codes[n_ref:].any()
            if miss:
                mk = engine.row_codes(b, j, n_pad, plane=1)
                assert np.array_equal(mk[:n_ref], (w["G"][pos[j]] >= 0).astype(np.int8)) and not mk[n_ref:].any()
        Q, A, N = engine.block_gram(b, m)
        Qo, Ao, No = O.gram_int(w["bed"], n_ref, pos)
        assert np.array_equal(Q, Qo), (name, b)
        assert np.array_equal(A, Ao) and np.array_equal(N, No), (name, b)
        S = engine.block_sigma(b, m)
        assert np.abs(S - O.sigma(w["bed"], n_ref, pos)).max() <= 1e-13, (name, b)
        assert np.array_equal(S, S.T)


def test_fused_unpack_gram_is_bit_exact():
    """DBSLMM_B200_GRAM=packed (experimental): blocks without missing calls are built from 2-bit rows packed in plan order
    and expanded to int8 in shared memory; blocks with missing calls still take the decoder's int8 rows.  Q / A / N bit for
    bit and betas as in the default path, on a panel that mixes both kinds of blocks."""
    sizes, n_ref = [90, 260, 33, 140], 125 * 4 - 3
    w = synth.make_workload(99, sizes, n_ref, missing_rate=0.01, frac_large=0.02)
    G = w["G"].copy()
    G[90:350] = np.where(G[90:350] < 0, 1, G[90:350])            # block 1 has no missing call
    bed = synth.pack_bed(G)
    os.environ["DBSLMM_B200_GRAM"] = "packed"
    try:
        eng = _abi.Engine(0)
    finally:
        os.environ.pop("DBSLMM_B200_GRAM", None)
    try:
        csr = (w["s_off"], w["s_pos"], w["s_z"], w["l_off"], w["l_pos"], w["l_z"])
        for streaming in (False, True):
            if streaming:
                r = eng.fit(*csr, sigma_s=[1e-4], n_obs=10_000, flags=_abi.FLAG_KEEP_INT_GRAM, bed=bed, n_ref=n_ref)
            else:
                eng.load_bed(bed, n_ref)
                r = eng.fit(*csr, sigma_s=[1e-4], n_obs=10_000, flags=_abi.FLAG_KEEP_INT_GRAM)
            assert r["timing"]["n_blocks_missing"] == 3
            for b, m in enumerate(w["block_sizes"]):
                pos = block_pos(w, b)
                Q, A, N = eng.block_gram(b, m)
                Qo, Ao, No = O.gram_int(bed, n_ref, pos)
                assert np.array_equal(Q, Qo) and np.array_equal(A, Ao) and np.array_equal(N, No)
            bs, bl, _, _ = O.est(bed, n_ref, 10_000, 1e-4, *csr, threads=2, mode=O.MODE_EXACT)
            assert relmax(r["beta_s"][0], bs) <= 1e-10 and relmax(r["beta_l"][0], bl) <= 1e-10
    finally:
        eng.close()


@pytest.mark.parametrize("variant", ["single", "pair"])
def test_one_plane_gram_variants_are_bit_exact(variant):
    """Both one-plane correlation builders -- 128 x 128 tiles by one CTA (default) and 256 x 256 super tiles by a CTA pair
    (tcgen05 cta_group::2) -- give the integer Gram bit for bit and Sigma / betas as the oracle, on block sizes around the
    tile edges (1 .. 3 super tiles per side, last tile rows of 1, 127, 128 and 129 SNPs), resident and streaming."""
    sizes, n_ref = [129, 256, 1, 385, 640, 511, 257], 500
    w = synth.make_workload(321, sizes, n_ref, missing_rate=0.0, frac_large=0.02)
    os.environ["DBSLMM_B200_GRAM"] = variant
    try:
        eng = _abi.Engine(0)
    finally:
        os.environ.pop("DBSLMM_B200_GRAM", None)
    try:
        csr = (w["s_off"], w["s_pos"], w["s_z"], w["l_off"], w["l_pos"], w["l_z"])
        bs, bl, _, _ = O.est(w["bed"], n_ref, 10_000, 1e-4, *csr, threads=4, mode=O.MODE_EXACT)
        for streaming in (False, True):
            if streaming:
                r = eng.fit(*csr, sigma_s=[1e-4], n_obs=10_000, flags=_abi.FLAG_KEEP_INT_GRAM, bed=w["bed"], n_ref=n_ref)
            else:
                eng.load_bed(w["bed"], n_ref)
                r = eng.fit(*csr, sigma_s=[1e-4], n_obs=10_000, flags=_abi.FLAG_KEEP_INT_GRAM)
            assert r["n_bad"] == 0 and r["timing"]["n_blocks_missing"] == 0
            for b, m in enumerate(w["block_sizes"]):
                Q, A, N = eng.block_gram(b, m)
                Qo, Ao, No = O.gram_int(w["bed"], n_ref, block_pos(w, b))
                assert np.array_equal(Q, Qo) and np.array_equal(A, Ao) and np.array_equal(N, No)
            assert relmax(r["beta_s"][0], bs) <= 1e-10 and relmax(r["beta_l"][0], bl) <= 1e-10
        # the production path (no debug planes: the straight-line and the predicated epilogue paths instead of the
        # row-by-row one) must give the same Sigma and betas
        r2 = eng.fit(*csr, sigma_s=[1e-4], n_obs=10_000)
        assert r2["n_bad"] == 0
        assert relmax(r2["beta_s"][0], bs) <= 1e-10 and relmax(r2["beta_l"][0], bl) <= 1e-10
        for b, m in enumerate(w["block_sizes"]):
            S = eng.block_sigma(b, m)
            assert np.abs(S - O.sigma(w["bed"], n_ref, block_pos(w, b))).max() <= 1e-13, (variant, b)
    finally:
        eng.close()


@pytest.mark.parametrize("mode", ["dbslmm", "lmm"])
def test_beta_vs_exact_oracle(case, engine, mode):
    _, w = case
    n_ref = w["n_ref"]
    sig, n_obs = 0.5 / 3000.0, 30_000
    if mode == "dbslmm":
        csr = (w["s_off"], w["s_pos"], w["s_z"], w["l_off"], w["l_pos"], w["l_z"])
    else:
        sizes = w["block_sizes"]
        off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
        z = np.zeros(off[-1]); z[w["s_pos"]] = w["s_z"]; z[w["l_pos"]] = w["l_z"]
        csr = (off, np.arange(off[-1], dtype=np.int32), z)
    r = engine.fit(*csr, sigma_s=[sig], n_obs=n_obs)
    assert r["n_bad"] == 0 and not r["status"].any()
    bs, bl, sing, _ = O.est(w["bed"], n_ref, n_obs, sig, *csr, threads=4, mode=O.MODE_EXACT)
    assert sing == 0
    assert relmax(r["beta_s"][0], bs) <= 1e-10
    if mode == "dbslmm" and bl.size:
        assert relmax(r["beta_l"][0], bl) <= 1e-10
    assert r["timing"]["n_launches"] > 0


def test_h2_folds_share_one_gram(engine):
    w = synth.make_workload(77, [180, 90, 40], 500, frac_large=0.02)
    engine.load_bed(w["bed"], 500)
    csr = (w["s_off"], w["s_pos"], w["s_z"], w["l_off"], w["l_pos"], w["l_z"])
    sig = np.array([0.8, 1.0, 1.2]) * 0.4 / 2000.0
    r = engine.fit(*csr, sigma_s=sig, n_obs=20_000)
    assert r["beta_s"].shape == (3, w["s_pos"].size)
    for f in range(3):
        bs, bl, _, _ = O.est(w["bed"], 500, 20_000, float(sig[f]), *csr, threads=4, mode=O.MODE_EXACT)
        assert relmax(r["beta_s"][f], bs) <= 1e-10 and relmax(r["beta_l"][f], bl) <= 1e-10
    assert relmax(r["beta_s"][0], r["beta_s"][2]) > 1e-3          # the folds really differ


def test_plan_cache_reuses_layout_with_new_z(engine):
    w = synth.make_workload(5, [120, 300], 400, frac_large=0.0)
    engine.load_bed(w["bed"], 400)
    r1 = engine.fit(w["s_off"], w["s_pos"], w["s_z"], sigma_s=[1e-4], n_obs=5000)
    z2 = w["s_z"][::-1].copy()
    r2 = engine.fit(w["s_off"], w["s_pos"], z2, sigma_s=[1e-4], n_obs=5000, flags=_abi.FLAG_PLAN_CACHED)
    b2, _, _, _ = O.est(w["bed"], 400, 5000, 1e-4, w["s_off"], w["s_pos"], z2, threads=4, mode=O.MODE_EXACT)
    assert relmax(r2["beta_s"][0], b2) <= 1e-10 and relmax(r1["beta_s"][0], b2) > 1e-3


def test_golden_c1_testdat(engine):
    """BASELINE.json configs[0] against the unmodified reference's FP64 betas."""
    d = np.load(os.path.join(GOLD, "c1_testdat.npz"))
    n_ref, n_obs, sig = int(d["n_ref"]), int(d["n_obs"]), float(d["sigma_s"])
    engine.load_bed(d["bed"], n_ref)
    maf, _ = engine.snp_stats()
    assert np.abs(maf - d["ref_maf"]).max() <= 1e-15
    r = engine.fit(d["lmm_off"], d["lmm_pos"], d["lmm_z"], sigma_s=[sig], n_obs=n_obs)
    assert r["n_bad"] == 0
    g = relmax(r["beta_s"][0], d["lmm_beta"]); record("gpu_chol_vs_reference/c1_testdat/lmm_beta", g)
    assert g <= bar("c1_testdat", "lmm_beta")                       # the reference's own truncation gap
    r = engine.fit(d["s_off"], d["s_pos"], d["s_z"], d["l_off"], d["l_pos"], d["l_z"], sigma_s=[sig], n_obs=n_obs)
    gs, gl = relmax(r["beta_s"][0], d["beta_s"]), relmax(r["beta_l"][0], d["beta_l"])
    record("gpu_chol_vs_reference/c1_testdat/beta_s", gs); record("gpu_chol_vs_reference/c1_testdat/beta_l", gl)
    assert gs <= bar("c1_testdat", "beta_s") and gl <= bar("c1_testdat", "beta_l")
    bs, bl, _, _ = O.est(d["bed"], n_ref, n_obs, sig, d["s_off"], d["s_pos"], d["s_z"], d["l_off"], d["l_pos"], d["l_z"],
                         mode=O.MODE_EXACT)
    assert relmax(r["beta_s"][0], bs) <= 1e-10 and relmax(r["beta_l"][0], bl) <= 1e-10


def test_golden_synth_ragged(engine):
    d = np.load(os.path.join(GOLD, "synth_ragged.npz"))
    n_ref, n_obs, sig = int(d["n_ref"]), int(d["n_obs"]), float(d["sigma_s"])
    engine.load_bed(d["bed"], n_ref)
    r = engine.fit(d["s_off"], d["s_pos"], d["s_z"], d["l_off"], d["l_pos"], d["l_z"], sigma_s=[sig], n_obs=n_obs)
    assert r["n_bad"] == 0
    gs, gl = relmax(r["beta_s"][0], d["beta_s"]), relmax(r["beta_l"][0], d["beta_l"])
    record("gpu_chol_vs_reference/synth_ragged/beta_s", gs); record("gpu_chol_vs_reference/synth_ragged/beta_l", gl)
    assert gs <= bar("synth_ragged", "beta_s") and gl <= bar("synth_ragged", "beta_l")


def test_monomorphic_snp_poisons_only_its_block(engine):
    """stddev = 0 => NaN column in the reference (dtpr.cpp:378); we flag the block and leave the others intact."""
    w = synth.make_workload(9, [40, 50], 400, frac_large=0.0)
    G = w["G"].copy()
    G[3, :] = 1
    bed = synth.pack_bed(G)
    engine.load_bed(bed, 400)
    r = engine.fit(w["s_off"], w["s_pos"], w["s_z"], sigma_s=[1e-4], n_obs=5000)
    assert r["n_bad"] == 1 and r["status"][0] == 1 and r["status"][1] == 0
    assert not np.isfinite(r["beta_s"][0][:40]).all() and np.isfinite(r["beta_s"][0][40:]).all()


def test_argument_errors(engine):
    w = synth.make_workload(5, [20], 400, frac_large=0.0)
    engine.load_bed(w["bed"], 400)
    with pytest.raises(_abi.EngineError):
        engine.fit(w["s_off"], w["s_pos"] + 10_000, w["s_z"], sigma_s=[1e-4], n_obs=5000)
    with pytest.raises(_abi.EngineError):
        engine.fit(w["s_off"], w["s_pos"], w["s_z"], sigma_s=[-1.0], n_obs=5000)


def test_full_size_properties(engine):
    """At a BASELINE-sized block (m = 3000, n_ref = 2000) the oracle is too slow for a per-call check:
    use size-independent properties: K x = z residual from Sigma itself, symmetry, unit diagonal structure."""
    sizes = [3000, 646]
    rng = np.random.default_rng(3)
    G = synth.make_genotypes(rng, sizes, 2000)
    bed = synth.pack_bed(G)
    engine.load_bed(bed, 2000)
    off = np.array([0, 3000, 3646], np.int32)
    z = rng.standard_normal(3646)
    sig, n_obs = 0.5 / 1.1e6, 300_000
    r = engine.fit(off, np.arange(3646, dtype=np.int32), z, sigma_s=[sig], n_obs=n_obs, flags=_abi.FLAG_FULL_SIGMA)
    assert r["n_bad"] == 0
    for b, (lo, hi) in enumerate(((0, 3000), (3000, 3646))):
        S = engine.block_sigma(b, hi - lo)
        assert np.array_equal(S, S.T)
        assert np.abs(np.diag(S) - (0.8 * 1999 / 2000 + 0.2)).max() < 1e-12
        K = S + np.eye(hi - lo) / (sig * n_obs)
        x = r["beta_s"][0][lo:hi] * np.sqrt(n_obs)
        res = np.abs(K @ x - z[lo:hi]).max() / np.abs(z[lo:hi]).max()
        assert res < 1e-11, res


def test_baseline_size_block_vs_exact_oracle(engine):
    """A BASELINE-sized block (m = 3,000 SNPs, n_ref = 2,000: 47 panels, split-K steps, cluster back substitution) in DBSLMM
    mode, compared DIRECTLY with the exact oracle -- a wrong Sigma cannot hide here the way it could behind a residual
    computed from the library's own Sigma.  Genome-wide ridge (nsnp = 1.1 M, N = 300,000)."""
    w = synth.make_workload(3000, [3000], 2000, missing_rate=0.0, frac_large=0.003)
    assert w["l_pos"].size >= 3
    engine.load_bed(w["bed"], 2000)
    csr = (w["s_off"], w["s_pos"], w["s_z"], w["l_off"], w["l_pos"], w["l_z"])
    sig, n_obs = 0.5 / 1.1e6, 300_000
    r = engine.fit(*csr, sigma_s=[sig], n_obs=n_obs)
    assert r["n_bad"] == 0
    bs, bl, sing, _ = O.est(w["bed"], 2000, n_obs, sig, *csr, threads=1, mode=O.MODE_EXACT)
    assert sing == 0
    gs, gl = relmax(r["beta_s"][0], bs), relmax(r["beta_l"][0], bl)
    record("gpu_chol_vs_exact_oracle/m3000_n2000_dbslmm/beta_s", gs); record("gpu_chol_vs_exact_oracle/m3000_n2000_dbslmm/beta_l", gl)
    assert gs <= 1e-10 and gl <= 1e-10
    rs = engine.fit(*csr, sigma_s=[sig], n_obs=n_obs, bed=w["bed"], n_ref=2000)            # and through the streaming path
    assert relmax(rs["beta_s"][0], bs) <= 1e-10 and relmax(rs["beta_l"][0], bl) <= 1e-10


def test_config5_large_reference_panel(engine):
    """BASELINE config 5 (n_ref = 20,000): pitch 5,000 B rows, n_pad 20,096, Gram entries up to 4 n = 80,000."""
    n_ref = 20000
    w = synth.make_workload(55, [260, 70], n_ref, missing_rate=0.0, frac_large=0.0)
    engine.load_bed(w["bed"], n_ref)
    maf, nn = engine.snp_stats()
    assert np.array_equal(nn, np.full(330, n_ref))
    r = engine.fit(w["s_off"], w["s_pos"], w["s_z"], sigma_s=[0.5 / 1e5], n_obs=300_000, flags=_abi.FLAG_KEEP_INT_GRAM)
    assert r["n_bad"] == 0
    pos = np.arange(260, dtype=np.int32)
    Q, A, N = engine.block_gram(0, 260)
    Qo, Ao, No = O.gram_int(w["bed"], n_ref, pos)
    assert np.array_equal(Q, Qo) and np.array_equal(A, Ao) and np.array_equal(N, No)
    # the reference float path sums 20,000 rounded products per entry: ITS rounding error grows like sqrt(n) eps, so
    # the bar is 1e-12 here (the kernel's integer numerator is exact at any n)
    dS = np.abs(engine.block_sigma(0, 260) - O.sigma(w["bed"], n_ref, pos)).max()
    assert dS <= 1e-12, dS
    bs, _, _, _ = O.est(w["bed"], n_ref, 300_000, 0.5 / 1e5, w["s_off"], w["s_pos"], w["s_z"], threads=4, mode=O.MODE_EXACT)
    assert relmax(r["beta_s"][0], bs) <= 1e-10


def test_config5_large_block(engine):
    """A 5,000-SNP block (79 panels, split-K steps, 1024-thread back substitution): residual of K x = z."""
    m, n_ref = 5000, 800
    rng = np.random.default_rng(50)
    G = synth.make_genotypes(rng, [m], n_ref)
    engine.load_bed(synth.pack_bed(G), n_ref)
    z = rng.standard_normal(m)
    sig, n_obs = 0.5 / 1e5, 200_000
    r = engine.fit(np.array([0, m], np.int32), np.arange(m, dtype=np.int32), z, sigma_s=[sig], n_obs=n_obs, flags=_abi.FLAG_FULL_SIGMA)
    assert r["n_bad"] == 0
    S = engine.block_sigma(0, m)
    assert np.array_equal(S, S.T)
    K = S + np.eye(m) / (sig * n_obs)
    x = r["beta_s"][0] * np.sqrt(n_obs)
    assert np.abs(K @ x - z).max() / np.abs(z).max() < 1e-10


def test_config2_lmm_chr1(engine):
    """BASELINE config 2: LMM mode on chr1-sized data (133 EUR LD blocks, ~90k SNPs, n_ref = 500)."""
    sizes = synth.eur_block_sizes(90_000, 3000, chroms=[1])
    assert sizes.size == 133
    n_ref = 500
    rng = np.random.default_rng(2)
    G = synth.make_genotypes(rng, sizes, n_ref)
    bed = synth.pack_bed(G)
    engine.load_bed(bed, n_ref)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    pos = np.arange(off[-1], dtype=np.int32)
    z = rng.standard_normal(off[-1]) * 1.5
    sig, n_obs = 0.5 / off[-1], 300_000
    r = engine.fit(off, pos, z, sigma_s=[sig], n_obs=n_obs)
    assert r["n_bad"] == 0 and np.isfinite(r["beta_s"]).all()
    sample = list(range(0, 133, 12))
    s_off = np.concatenate([[0], np.cumsum(sizes[sample])]).astype(np.int32)
    s_pos = np.concatenate([pos[off[b]:off[b + 1]] for b in sample])
    bs, _, _, _ = O.est(bed, n_ref, n_obs, sig, s_off, s_pos, z[s_pos], threads=8, mode=O.MODE_EXACT)
    assert relmax(r["beta_s"][0][s_pos], bs) <= 1e-10


# ------------------------------------------------------------------------------------------
# reference-faithful PCG solver (csrc/pcg.cu): same stopping rule as DBSLMMFIT::PCGv, so it lands on the
# reference's truncated answers; agreement is limited by CG's sensitivity to summation order near the 1e-7
# threshold (tests/test_oracle.py), never by the integer Gram.
# ------------------------------------------------------------------------------------------
def test_pcg_solver_matches_ref_mode_oracle(engine):
    w = synth.make_workload(321, [210, 0, 1, 33, 140], 400, missing_rate=0.01, frac_large=0.03)
    engine.load_bed(w["bed"], 400)
    csr = (w["s_off"], w["s_pos"], w["s_z"], w["l_off"], w["l_pos"], w["l_z"])
    sig, n_obs = 0.5 / 3000.0, 30_000
    r = engine.fit(*csr, sigma_s=[sig], n_obs=n_obs, solver=_abi.SOLVER_PCG)
    assert r["n_bad"] == 0
    bs, bl, sing, it = O.est(w["bed"], 400, n_obs, sig, *csr, threads=4, mode=O.MODE_REF)
    assert sing == 0
    # two faithful PCG implementations (this kernel, the ref-mode oracle): same algorithm, other summation order; the
    # iteration count may flip by one at the 1e-7 threshold, which moves a solve by up to ~1e-7 / lambda_min(A) = 5e-7 of
    # |u|; measured here 1e-9 .. 3e-8
    gs, gl = relmax(r["beta_s"][0], bs), relmax(r["beta_l"][0], bl)
    record("gpu_pcg_vs_ref_oracle/synth321/beta_s", gs); record("gpu_pcg_vs_ref_oracle/synth321/beta_l", gl)
    assert gs <= 1e-7 and gl <= 1e-7
    its = [engine.block_iters(b) for b in range(5)]
    assert abs(max(its) - it) <= 1 and its[1] == 0
    # and it differs from the exact solve by the reference's own truncation error, not by more
    be, ble, _, _ = O.est(w["bed"], 400, n_obs, sig, *csr, threads=4, mode=O.MODE_EXACT)
    assert relmax(r["beta_s"][0], be) <= 5e-6


def test_pcg_solver_golden_c1(engine):
    d = np.load(os.path.join(GOLD, "c1_testdat.npz"))
    n_ref, n_obs, sig = int(d["n_ref"]), int(d["n_obs"]), float(d["sigma_s"])
    engine.load_bed(d["bed"], n_ref)
    r = engine.fit(d["lmm_off"], d["lmm_pos"], d["lmm_z"], sigma_s=[sig], n_obs=n_obs, solver=_abi.SOLVER_PCG)
    g = relmax(r["beta_s"][0], d["lmm_beta"]); record("gpu_pcg_vs_reference/c1_testdat/lmm_beta", g)
    assert g <= 5e-9                                                # same iteration count as the reference: ~1e-10 expected
    assert 40 <= engine.block_iters(0) <= 50                        # the reference takes 45-46 iterations here
    r = engine.fit(d["s_off"], d["s_pos"], d["s_z"], d["l_off"], d["l_pos"], d["l_z"], sigma_s=[sig], n_obs=n_obs,
                   solver=_abi.SOLVER_PCG)
    gs, gl = relmax(r["beta_s"][0], d["beta_s"]), relmax(r["beta_l"][0], d["beta_l"])
    record("gpu_pcg_vs_reference/c1_testdat/beta_s", gs); record("gpu_pcg_vs_reference/c1_testdat/beta_l", gl)
    assert gs <= 1e-7 and gl <= 1e-7                                # measured (ref-mode oracle vs golden): 1.8e-8 / 1.2e-8
