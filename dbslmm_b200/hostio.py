"""Host-side ingest that defines WHICH SNPs/blocks reach the kernels.

Python mirror (used by tests, bench and the Python `est` wrapper) of the reference's
text readers and matchers; the C++ CLI (dbslmm_b200/host/) implements the same rules.
Reference: scr/dtpr.cpp:47-68 (readBlock), :83-123 (readBim), :178-220 (readSumm),
:383-408 (matchRef), :455-481 (addBlock); scr/dbslmm.cpp:232-317 (BatchRun ingest).
Quirks kept on purpose: no header skipping, alleles must match exactly
(no flips), first duplicate rs id wins, `start <= ps < end` with sorted-input early
break, mafMax == 1 disables the MAF pre-pass.
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class Summ:            # SUMM, dtpr.hpp:58-68
    snp: list
    ps: np.ndarray
    a1: list
    a2: list
    maf: np.ndarray
    z: np.ndarray


@dataclass
class Info:            # INFO as SoA, dtpr.hpp:115-125
    snp: list
    ps: np.ndarray
    pos: np.ndarray     # bim/bed row
    block: np.ndarray
    a1: list
    maf: np.ndarray     # summary-stat maf (what the output file uses)
    z: np.ndarray


def read_fam_count(path):                      # IO::getRow, dtpr.cpp:71-80
    with open(path, "rb") as f:
        return sum(1 for _ in f)


def read_bed(path, n_snp, n_ref):
    """Returns the .bed payload AFTER the 3 magic bytes as uint8[n_snp, ceil(n_ref/4)]."""
    pitch = (n_ref + 3) // 4
    raw = np.fromfile(path, dtype=np.uint8)
    if raw.size < 3 or raw[0] != 0x6C or raw[1] != 0x1B or raw[2] != 0x01:
        raise ValueError(f"{path}: not a SNP-major PLINK .bed")
    body = raw[3:]
    if body.size < n_snp * pitch:
        raise ValueError(f"{path}: truncated .bed")
    return np.ascontiguousarray(body[: n_snp * pitch].reshape(n_snp, pitch))


def read_block(path):                          # dtpr.cpp:47-68
    start, end = [], []
    with open(path) as f:
        for line in f:
            t = line.rstrip("\n").split("\t")
            start.append(int(t[1]))
            end.append(int(t[2]))
    return np.asarray(start, np.int64), np.asarray(end, np.int64)


def read_bim(path):                            # dtpr.cpp:107-121
    """dict rs -> (pos, a1, a2); first duplicate wins (map::insert)."""
    bim = {}
    n = 0
    with open(path) as f:
        for line in f:
            t = line.rstrip("\n").split("\t")
            if t[1] not in bim:
                bim[t[1]] = (n, t[4], t[5])
            n += 1
    return bim, n


def _atof(s):
    try:
        return float(s)
    except ValueError:
        return 0.0


def read_summ(path):                           # dtpr.cpp:178-220
    snp, ps, a1, a2, maf, z = [], [], [], [], [], []
    with open(path) as f:
        for line in f:
            t = line.rstrip("\n").split("\t")
            zz = 0.0
            if t[9][:1].isdigit():                         # :194
                se = _atof(t[9])
                if se - 0.0 > 1e-20:                       # :196
                    zz = _atof(t[8]) / se
            snp.append(t[1])
            ps.append(int(_atof(t[2])))
            a1.append(t[5])
            a2.append(t[6])
            af = _atof(t[7])
            maf.append(min(af, 1.0 - af))                  # :213
            z.append(zz)
    return Summ(snp, np.asarray(ps, np.int64), a1, a2, np.asarray(maf), np.asarray(z))


def match_ref(summ, bim, ref_maf, maf_max):    # dtpr.cpp:383-408
    """Returns (keep_idx into summ, pos).  ref_maf: per-bim-row maf or None (mafMax==1)."""
    keep, pos = [], []
    for i, rs in enumerate(summ.snp):
        ent = bim.get(rs)
        if ent is None:                        # default-constructed ALLELE: empty alleles never match
            continue
        p, b1, b2 = ent
        if b1 != summ.a1[i] or b2 != summ.a2[i]:
            continue
        rm = 0.0 if ref_maf is None else ref_maf[p]
        if not (abs(rm - summ.maf[i]) < maf_max):
            continue
        keep.append(i)
        pos.append(p)
    return np.asarray(keep, np.int64), np.asarray(pos, np.int32)


def add_block(ps, blk_start, blk_end):         # dtpr.cpp:455-481
    """Block id per SNP with the reference's sorted-input early-break scan.
    SNPs the scan never reaches keep block 0 of a value-initialised INFO -- the reference
    leaves them default-constructed (empty snp => dropped at output); we return -1."""
    n = len(ps)
    block = np.full(n, -1, np.int32)
    count = 0
    for i in range(len(blk_start)):
        j = count
        while j < n and blk_start[i] <= ps[j] < blk_end[i]:
            block[j] = i
            count += 1
            j += 1
    return block


def to_csr(block, n_blocks):
    """Offsets for block-sorted SNP arrays (count_snps_per_block, helpers.cpp:16-30)."""
    off = np.zeros(n_blocks + 1, np.int32)
    valid = block[block >= 0]
    np.add.at(off, valid + 1, 1)
    return np.cumsum(off).astype(np.int32)
