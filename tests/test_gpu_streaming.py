"""Streaming fit: the reference panel travels with dbslmm_b200_fit (fit_args.bed), as DBSLMMFIT::est receives its
bed_str (reference scr/dbslmmfit.hpp:38-67); the upload is cut into batches and overlaps the fit.  Results must be
IDENTICAL (same kernels, same arithmetic, only a different schedule and device layout) to load_bed + fit."""
import functools

import numpy as np
import pytest

from dbslmm_b200 import _abi, synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def relmax(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def csr_of(w):
    return (w["s_off"], w["s_pos"], w["s_z"], w["l_off"], w["l_pos"], w["l_z"])


@pytest.mark.parametrize("sizes,n_ref", [([300, 0, 1, 7, 8, 63, 64, 65, 127, 128, 129, 200], 400),
                                         ([90, 33], 125 * 4 - 3),
                                         ([520, 40, 1100, 70, 2100, 600, 9, 700], 500)])
def test_streaming_equals_resident_and_oracle(engine, sizes, n_ref):
    w = synth.make_workload(4321 + len(sizes), sizes, n_ref, missing_rate=0.0, frac_large=0.02)
    csr = csr_of(w)
    kw = dict(sigma_s=[2e-4], n_obs=20_000)
    rs = engine.fit(*csr, bed=w["bed"], n_ref=n_ref, **kw)                 # panel uploaded inside the call
    assert rs["n_bad"] == 0
    engine.load_bed(w["bed"], n_ref)
    rr = engine.fit(*csr, **kw)
    assert np.array_equal(rs["beta_s"], rr["beta_s"]) and np.array_equal(rs["beta_l"], rr["beta_l"])
    bs, bl, _, _ = O.est(w["bed"], n_ref, 20_000, 2e-4, *csr, threads=4, mode=O.MODE_EXACT)
    assert relmax(rs["beta_s"][0], bs) <= 1e-10
    if bl.size:
        assert relmax(rs["beta_l"][0], bl) <= 1e-10
    # the panel stayed resident and its statistics are computed on demand
    maf, nn = engine.snp_stats()
    assert np.array_equal(nn, (w["G"] >= 0).sum(axis=1))
    r3 = engine.fit(*csr, **kw)
    assert np.array_equal(r3["beta_s"], rr["beta_s"])


def test_streaming_many_blocks_sub_batches(engine):
    """A genome-like bulk (>= 256 small/medium blocks, >= 100k SNPs) is cut into four region batches (upload units)."""
    rng = np.random.default_rng(11)
    sizes = [int(x) for x in rng.integers(300, 440, size=290)] + [600, 1300, 2100]
    w = synth.make_workload(99, sizes, 400, missing_rate=0.0, frac_large=0.01)
    csr = csr_of(w)
    kw = dict(sigma_s=[1e-4, 3e-4], n_obs=50_000)
    rs = engine.fit(*csr, bed=w["bed"], n_ref=400, **kw)
    assert rs["n_bad"] == 0
    engine.load_bed(w["bed"], 400)
    rr = engine.fit(*csr, **kw)
    # (sub-batches may pick another split-K factor than the whole class: same arithmetic, different summation order)
    assert relmax(rs["beta_s"], rr["beta_s"]) <= 1e-12 and relmax(rs["beta_l"], rr["beta_l"]) <= 1e-12


@functools.lru_cache(maxsize=1)
def _order_case():
    """>= 256 bulk blocks and >= 100k SNPs (the bulk is cut into regions) next to three big blocks; betas of the exact oracle."""
    rng = np.random.default_rng(12)
    sizes = [int(x) for x in rng.integers(300, 440, size=285)] + [1300, 2100, 1100]
    w = synth.make_workload(98, sizes, 400, missing_rate=0.0, frac_large=0.01)
    bs, bl, _, _ = O.est(w["bed"], 400, 50_000, 2e-4, *csr_of(w), threads=4, mode=O.MODE_EXACT)
    return w, bs, bl


@pytest.mark.parametrize("order", ["0", "1", None])
def test_streaming_upload_order_does_not_change_the_result(engine, order, monkeypatch):
    """The big classes go out first when their dependency chains outlast the bulk (one rank's shard of a multi-GPU run, and
    this workload), a bulk region first when the fit is throughput-bound (the whole genome); DBSLMM_B200_UPLOAD_BULK_FIRST
    forces either.  Same batches, same kernels: the betas do not depend on it."""
    w, bs, bl = _order_case()
    csr = csr_of(w)
    kw = dict(sigma_s=[2e-4], n_obs=50_000)
    if order is None:
        monkeypatch.delenv("DBSLMM_B200_UPLOAD_BULK_FIRST", raising=False)
    else:
        monkeypatch.setenv("DBSLMM_B200_UPLOAD_BULK_FIRST", order)
    eng = _abi.Engine(0)                     # the switches are read when a handle is created
    try:
        rs = eng.fit(*csr, bed=w["bed"], n_ref=400, **kw)
        assert rs["n_bad"] == 0
        rs2 = eng.fit(*csr, bed=w["bed"], n_ref=400, **kw)
        assert np.array_equal(rs["beta_s"], rs2["beta_s"]) and np.array_equal(rs["beta_l"], rs2["beta_l"])
    finally:
        eng.close()
    engine.load_bed(w["bed"], 400)
    rr = engine.fit(*csr, **kw)
    assert relmax(rs["beta_s"], rr["beta_s"]) <= 1e-12 and relmax(rs["beta_l"], rr["beta_l"]) <= 1e-12
    assert relmax(rs["beta_s"][0], bs) <= 1e-10 and relmax(rs["beta_l"][0], bl) <= 1e-10


def test_streaming_subset_of_rows_and_unordered_blocks(engine):
    """Blocks use a subset of the panel rows, not in .bed order: unused rows are uploaded last, the panel ends up whole."""
    w = synth.make_workload(17, [150, 260, 90, 400], 400, missing_rate=0.0, frac_large=0.0)
    n_snp = w["bed"].shape[0]
    keep = np.ones(n_snp, bool)
    keep[::7] = False                                        # drop every 7th SNP from the fit
    s_off = [0]
    pos = []
    for b in (2, 0, 3, 1):                                   # block order != .bed order
        p = w["s_pos"][w["s_off"][b]:w["s_off"][b + 1]]
        p = p[keep[p]]
        pos.append(p)
        s_off.append(s_off[-1] + p.size)
    s_off = np.array(s_off, np.int32)
    s_pos = np.concatenate(pos).astype(np.int32)
    z = np.random.default_rng(5).standard_normal(s_pos.size)
    rs = engine.fit(s_off, s_pos, z, sigma_s=[1e-4], n_obs=9000, bed=w["bed"], n_ref=400)
    bs, _, _, _ = O.est(w["bed"], 400, 9000, 1e-4, s_off, s_pos, z, threads=4, mode=O.MODE_EXACT)
    assert rs["n_bad"] == 0 and relmax(rs["beta_s"][0], bs) <= 1e-10
    maf, _ = engine.snp_stats()
    assert np.abs(maf - O.snp_maf(w["bed"], n_snp, 400)).max() <= 1e-15


def test_streaming_with_missing_calls_does_not_fall_back(engine):
    """Which blocks have missing calls is decided on the device (decoder counts -> per-block flag -> one- or four-plane
    Gram), so a panel with missing calls streams like any other: no second pass on a resident copy."""
    w = synth.make_workload(23, [150, 70, 5, 130, 0, 257, 90], 403, missing_rate=0.02, frac_large=0.02)
    # one block without any missing call between blocks that have them (mixed one- / four-plane tiles in one launch)
    G = w["G"].copy()
    lo = int(np.sum(w["block_sizes"][:6])); hi = lo + 90
    G[lo:hi] = np.where(G[lo:hi] < 0, 1, G[lo:hi])
    bed = synth.pack_bed(G)
    csr = csr_of(w)
    bs, bl, _, _ = O.est(bed, 403, 10_000, 1e-4, *csr, threads=4, mode=O.MODE_EXACT)
    for rep in range(2):
        rs = engine.fit(*csr, sigma_s=[1e-4], n_obs=10_000, bed=bed, n_ref=403)
        assert rs["timing"]["streamed"] == 1 and rs["timing"]["n_blocks_missing"] == 5
        assert rs["n_bad"] == 0 and relmax(rs["beta_s"][0], bs) <= 1e-10 and relmax(rs["beta_l"][0], bl) <= 1e-10
    # the same layout again with a panel WITHOUT missing calls: the mask rows written above must not leak into it
    G0 = np.where(G < 0, 0, G)
    bed0 = synth.pack_bed(G0)
    b0, l0, _, _ = O.est(bed0, 403, 10_000, 1e-4, *csr, threads=4, mode=O.MODE_EXACT)
    r0 = engine.fit(*csr, sigma_s=[1e-4], n_obs=10_000, bed=bed0, n_ref=403)
    assert r0["timing"]["n_blocks_missing"] == 0 and relmax(r0["beta_s"][0], b0) <= 1e-10 and relmax(r0["beta_l"][0], l0) <= 1e-10
    # and back: SNPs whose mask row went back to the default pattern sit next to SNPs with missing calls again
    rs = engine.fit(*csr, sigma_s=[1e-4], n_obs=10_000, bed=bed, n_ref=403, flags=_abi.FLAG_KEEP_INT_GRAM)
    assert relmax(rs["beta_s"][0], bs) <= 1e-10
    pos0 = np.concatenate([w["s_pos"][w["s_off"][0]:w["s_off"][1]], w["l_pos"][w["l_off"][0]:w["l_off"][1]]]).astype(np.int32)
    Q, A, N = engine.block_gram(0, pos0.size)
    Qo, Ao, No = O.gram_int(bed, 403, pos0)
    assert np.array_equal(Q, Qo) and np.array_equal(A, Ao) and np.array_equal(N, No)


def test_streaming_failure_leaves_no_half_loaded_panel(engine):
    """A streaming fit that fails (here: a SNP row outside the panel) must not leave a panel a later call could use."""
    w = synth.make_workload(5, [60, 80], 400, frac_large=0.0)
    bad = w["s_pos"].copy(); bad[-1] = 10_000_000
    with pytest.raises(_abi.EngineError):
        engine.fit(w["s_off"], bad, w["s_z"], sigma_s=[1e-4], n_obs=5000, bed=w["bed"], n_ref=400)
    with pytest.raises(_abi.EngineError):
        engine.fit(w["s_off"], w["s_pos"], w["s_z"], sigma_s=[1e-4], n_obs=5000)          # no panel: ERR_STATE
    r = engine.fit(w["s_off"], w["s_pos"], w["s_z"], sigma_s=[1e-4], n_obs=5000, bed=w["bed"], n_ref=400)
    b, _, _, _ = O.est(w["bed"], 400, 5000, 1e-4, w["s_off"], w["s_pos"], w["s_z"], threads=2, mode=O.MODE_EXACT)
    assert relmax(r["beta_s"][0], b) <= 1e-10


def test_plan_cache_flag_rebuilds_when_the_block_lists_change(engine):
    """FLAG_PLAN_CACHED with the same totals but other SNP rows must not reuse the stale row map (fingerprint)."""
    w = synth.make_workload(6, [120, 300], 400, frac_large=0.0)
    engine.load_bed(w["bed"], 400)
    engine.fit(w["s_off"], w["s_pos"], w["s_z"], sigma_s=[1e-4], n_obs=5000)
    pos2 = w["s_pos"][::-1].copy()                                    # same counts, other rows per block
    r2 = engine.fit(w["s_off"], pos2, w["s_z"], sigma_s=[1e-4], n_obs=5000, flags=_abi.FLAG_PLAN_CACHED)
    b2, _, _, _ = O.est(w["bed"], 400, 5000, 1e-4, w["s_off"], pos2, w["s_z"], threads=2, mode=O.MODE_EXACT)
    assert relmax(r2["beta_s"][0], b2) <= 1e-10


def test_streaming_scattered_rows_use_plain_upload(engine):
    """Every block draws its SNPs from the whole panel: covering ranges would move ~n_blocks x the panel, so the call
    degrades to a plain upload followed by the resident fit."""
    w = synth.make_workload(31, [400, 400, 400], 400, missing_rate=0.0, frac_large=0.0)
    n_snp = w["bed"].shape[0]
    perm = np.random.default_rng(2).permutation(n_snp).astype(np.int32)
    s_off = np.array([0, 400, 800, 1200], np.int32)
    z = np.random.default_rng(3).standard_normal(n_snp)
    rs = engine.fit(s_off, perm, z, sigma_s=[1e-4], n_obs=9000, bed=w["bed"], n_ref=400)
    bs, _, _, _ = O.est(w["bed"], 400, 9000, 1e-4, s_off, perm, z, threads=4, mode=O.MODE_EXACT)
    assert relmax(rs["beta_s"][0], bs) <= 1e-10


def test_fit_multi_fans_out_over_handles(engine):
    """dbslmm_b200_fit_multi: blocks assigned by the scheduler, every handle uploads only its rows, betas gathered block-major.
    Two handles (on two GPUs when the box has them, else both on GPU 0) against the exact oracle and a single-handle fit."""
    n_dev = _abi.load().dbslmm_b200_device_count()
    w = synth.make_workload(77, [300, 0, 90, 1100, 40, 520, 260, 700], 400, missing_rate=0.003, frac_large=0.02)
    csr = csr_of(w)
    kw = dict(sigma_s=[1e-4, 2e-4], n_obs=20_000)
    second = _abi.Engine(1 if n_dev >= 2 else 0)
    try:
        rm = _abi.fit_multi([engine, second], *csr, bed=w["bed"], n_ref=400, **kw)
        r1 = engine.fit(*csr, bed=w["bed"], n_ref=400, **kw)
        assert rm["n_bad"] == 0 and rm["status"].shape == r1["status"].shape
        assert relmax(rm["beta_s"], r1["beta_s"]) <= 1e-12 and relmax(rm["beta_l"], r1["beta_l"]) <= 1e-12
        for f, sg in enumerate(kw["sigma_s"]):
            bs, bl, _, _ = O.est(w["bed"], 400, 20_000, sg, *csr, threads=4, mode=O.MODE_EXACT)
            assert relmax(rm["beta_s"][f], bs) <= 1e-10 and relmax(rm["beta_l"][f], bl) <= 1e-10
        # the handles keep no panel after a subset upload: a fit without fit_args.bed must say so
        with pytest.raises(_abi.EngineError):
            second.fit(*csr, **kw)
    finally:
        second.close()
