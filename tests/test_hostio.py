"""CPU: host ingest mirror (which SNPs / blocks reach the kernels) on the C1 fixture."""
import os

import numpy as np

from dbslmm_b200 import hostio as H

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _files(tmp_path):
    d = np.load(os.path.join(GOLD, "c1_testdat.npz"))
    (tmp_path / "ref.bim").write_text(str(d["bim_txt"]))
    (tmp_path / "ref.fam").write_text("x\n" * int(d["fam_lines"]))
    with open(tmp_path / "ref.bed", "wb") as f:
        f.write(bytes([0x6C, 0x1B, 0x01]) + d["bed"].tobytes())
    (tmp_path / "summ.txt").write_text(str(d["summary_txt"]))
    (tmp_path / "blocks.bed").write_text(str(d["block_txt"]))
    return d


def test_c1_matching_counts(tmp_path):
    d = _files(tmp_path)
    n_ref = H.read_fam_count(tmp_path / "ref.fam")
    bim, nsnp = H.read_bim(tmp_path / "ref.bim")
    assert (n_ref, nsnp) == (400, 723)
    bed = H.read_bed(tmp_path / "ref.bed", nsnp, n_ref)
    assert np.array_equal(bed, d["bed"])
    summ = H.read_summ(tmp_path / "summ.txt")
    assert len(summ.snp) == 996                                   # no header skipping
    keep_all, _ = H.match_ref(summ, bim, None, 1.0)               # mafMax == 1: filter vacuous
    assert keep_all.size == 717
    keep, pos = H.match_ref(summ, bim, d["ref_maf"], 0.2)
    assert keep.size == 716
    bs, be = H.read_block(tmp_path / "blocks.bed")
    assert bs.size == 133
    blk = H.add_block(summ.ps[keep], bs, be)
    assert (blk == 0).all()                                       # every SNP falls in EUR chr1 block 0
    off = H.to_csr(blk, bs.size)
    assert np.array_equal(off, d["lmm_off"]) and np.array_equal(pos, d["lmm_pos"])
    assert np.allclose(summ.z[keep], d["lmm_z"], rtol=0, atol=0)


def test_add_block_early_break_semantics():
    bs, be = np.array([10, 20, 30]), np.array([20, 30, 40])
    assert H.add_block(np.array([10, 19, 20, 35]), bs, be).tolist() == [0, 0, 1, 2]
    # a SNP before the first block stops the scan for good, exactly like the reference's sorted-input loop
    assert H.add_block(np.array([5, 12, 25]), bs, be).tolist() == [-1, -1, -1]
    assert H.to_csr(np.array([0, 0, 2], np.int32), 3).tolist() == [0, 2, 2, 3]
