"""Synthetic workloads of the shapes BASELINE.json names.

Genotypes: per-SNP MAF ~ U(0.05, 0.5); within-block LD from an AR(1) latent Gaussian per
haplotype (rho = 0.9) thresholded to Bernoulli(p) alleles, independent across blocks;
optional missing calls; monomorphic columns are re-drawn.  Summary statistics: sparse large
effects + polygenic noise; the large-effect set is a greedy distance clump of |z| > 4.9
(mimicking `plink --clump` p1=1e-6 as used by software/DBSLMM.R:140-145).
numpy here (tests, small sizes); bench.py has a torch-CUDA twin of `make_bed` for the
genome-wide size.
"""
import json
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "eur_ld_block_lengths.json")


def eur_block_sizes(total_snps, cap, chroms=range(1, 23), min_size=4):
    """SNPs per EUR LD block for `total_snps` spread uniformly in bp over `chroms`,
    capped at `cap` per block with the excess redistributed."""
    lens = json.load(open(_DATA))
    L = np.concatenate([np.asarray(lens[str(c)], np.float64) for c in chroms])
    m = np.maximum(min_size, np.round(L / L.sum() * total_snps)).astype(np.int64)
    for _ in range(50):
        over = m > cap
        excess = int((m[over] - cap).sum())
        if excess <= 0:
            break
        m[over] = cap
        room = ~over
        m[room] += np.floor(excess * L[room] / L[room].sum()).astype(np.int64)
    return np.minimum(m, cap).astype(np.int32)


def pack_bed(G):
    """G: int8 [n_snp, n] with values 0,1,2 (allele count of A1) or -1 (missing) -> uint8 [n_snp, ceil(n/4)]
    PLINK codes: count 2 -> 00, 1 -> 10, 0 -> 11, missing -> 01 (dtpr.cpp:329-350)."""
    n_snp, n = G.shape
    pitch = (n + 3) // 4
    code = np.full((n_snp, pitch * 4), 0, np.uint8)       # padding bits 00, as PLINK writes
    lut = np.array([3, 2, 0, 1], np.uint8)                # index g in {0,1,2,-1 -> 3}
    idx = np.where(G < 0, 3, G).astype(np.int64)
    code[:, :n] = lut[idx]
    c = code.reshape(n_snp, pitch, 4)
    return (c[:, :, 0] | (c[:, :, 1] << 2) | (c[:, :, 2] << 4) | (c[:, :, 3] << 6)).astype(np.uint8)


def make_genotypes(rng, block_sizes, n, rho=0.9, missing_rate=0.0):
    """int8 [sum(block_sizes), n] allele counts (-1 = missing)."""
    from scipy.signal import lfilter
    from scipy.special import ndtri
    out = []
    s = np.sqrt(1.0 - rho * rho)
    for m in block_sizes:
        m = int(m)
        if m == 0:
            continue
        p = rng.uniform(0.05, 0.5, size=m)
        thr = ndtri(p)[:, None]
        E = rng.standard_normal((m, 2 * n))
        E[0] /= s
        Lat = lfilter([s], [1.0, -rho], E, axis=0)
        A = (Lat < thr)
        G = (A[:, :n].astype(np.int8) + A[:, n:].astype(np.int8))
        # re-draw monomorphic columns independently
        mono = (G.min(axis=1) == G.max(axis=1))
        for j in np.where(mono)[0]:
            while G[j].min() == G[j].max():
                G[j] = rng.binomial(2, max(p[j], 0.2), size=n).astype(np.int8)
        if missing_rate > 0:
            miss = rng.random((m, n)) < missing_rate
            G = np.where(miss, np.int8(-1), G)
        out.append(G)
    return np.concatenate(out, axis=0) if out else np.zeros((0, n), np.int8)


def make_sumstats(rng, block_sizes, frac_large=1e-3, clump_dist=10, z_thresh=4.9):
    """z-scores per SNP and a boolean large-effect mask (greedy per-block distance clump)."""
    tot = int(np.sum(block_sizes))
    z = rng.standard_normal(tot) * 1.3
    n_big = max(1, int(round(frac_large * tot))) if frac_large > 0 else 0
    if n_big:
        big = rng.choice(tot, size=n_big, replace=False)
        z[big] = rng.choice([-1.0, 1.0], size=n_big) * (6.0 + rng.exponential(2.0, size=n_big))
        # LD smears a big signal onto its neighbours
        for b in big:
            lo, hi = max(0, b - 3), min(tot, b + 4)
            z[lo:hi] += z[b] * 0.6 * (np.arange(lo, hi) != b)
    large = np.zeros(tot, bool)
    off = 0
    for m in block_sizes:
        m = int(m)
        idx = np.where(np.abs(z[off:off + m]) > z_thresh)[0]
        idx = idx[np.argsort(-np.abs(z[off + idx]))]
        chosen = []
        for i in idx:
            if all(abs(i - c) >= clump_dist for c in chosen):
                chosen.append(i)
        if chosen:
            large[off + np.asarray(chosen, np.int64)] = True
        off += m
    return z, large


def split_csr(block_sizes, large_mask):
    """CSR offsets + index arrays for small and large SNPs (block-major, small first)."""
    nb = len(block_sizes)
    s_off = np.zeros(nb + 1, np.int32)
    l_off = np.zeros(nb + 1, np.int32)
    s_idx, l_idx = [], []
    off = 0
    for b, m in enumerate(block_sizes):
        m = int(m)
        lm = large_mask[off:off + m]
        s_idx.append(off + np.where(~lm)[0])
        l_idx.append(off + np.where(lm)[0])
        s_off[b + 1] = s_off[b] + (m - int(lm.sum()))
        l_off[b + 1] = l_off[b] + int(lm.sum())
        off += m
    cat = lambda xs: np.concatenate(xs).astype(np.int32) if xs else np.zeros(0, np.int32)
    return s_off, cat(s_idx), l_off, cat(l_idx)


def make_workload(seed, block_sizes, n_ref, missing_rate=0.0, frac_large=1e-3):
    rng = np.random.default_rng(seed)
    G = make_genotypes(rng, block_sizes, n_ref, missing_rate=missing_rate)
    bed = pack_bed(G)
    z, large = make_sumstats(rng, block_sizes, frac_large=frac_large)
    s_off, s_pos, l_off, l_pos = split_csr(block_sizes, large)
    return {"bed": bed, "G": G, "n_ref": n_ref, "block_sizes": np.asarray(block_sizes, np.int32),
            "s_off": s_off, "s_pos": s_pos, "s_z": z[s_pos], "l_off": l_off, "l_pos": l_pos, "l_z": z[l_pos]}
