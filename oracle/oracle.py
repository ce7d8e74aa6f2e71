"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package (dbslmm_b200)
never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MODE_REF = 0    # PCG exactly as dbslmmfit.cpp:629-668 + the reference's beta_s form
MODE_EXACT = 1  # Cholesky + direct form


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "dbslmm_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.orc_read_snp_im.restype = C.c_int
        _LIB.orc_est_block.restype = C.c_int
        _LIB.orc_est.restype = C.c_int
        _LIB.orc_pcgv.restype = C.c_int
        _LIB.orc_num_threads.restype = C.c_int
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _bed(bed):
    bed = np.ascontiguousarray(bed, dtype=np.uint8)
    return bed


def read_snp_im(bed, pos, n_total, indicator=None):
    """IO::readSNPIm: returns (geno[kept], maf)."""
    bed = _bed(bed)
    ind = None if indicator is None else np.ascontiguousarray(indicator, dtype=np.int32)
    n_keep = n_total if ind is None else int(ind.sum())
    g = np.zeros(n_keep, dtype=np.float64)
    maf = C.c_double(0.0)
    lib().orc_read_snp_im(_p(bed), C.c_int64(pos), C.c_int(n_total), _p(ind), _p(g), C.byref(maf))
    return g, maf.value


def normalize(x):
    x = np.array(x, dtype=np.float64, copy=True)
    lib().orc_normalize(_p(x), C.c_int(x.size))
    return x


def gram_int(bed, n_ref, pos):
    bed = _bed(bed)
    pos = np.ascontiguousarray(pos, dtype=np.int32)
    m = pos.size
    Q = np.zeros((m, m), np.int32)
    A = np.zeros((m, m), np.int32)
    N = np.zeros((m, m), np.int32)
    lib().orc_gram_int(_p(bed), C.c_int(n_ref), _p(pos), C.c_int(m), _p(Q), _p(A), _p(N))
    return Q, A, N


def snp_maf(bed, n_snp, n_ref):
    bed = _bed(bed)
    maf = np.zeros(n_snp, np.float64)
    lib().orc_snp_maf(_p(bed), C.c_int64(n_snp), C.c_int(n_ref), _p(maf))
    return maf


def sigma(bed, n_ref, pos, tau=0.8):
    bed = _bed(bed)
    pos = np.ascontiguousarray(pos, dtype=np.int32)
    m = pos.size
    S = np.zeros((m, m), np.float64)
    lib().orc_sigma(_p(bed), C.c_int(n_ref), _p(pos), C.c_int(m), C.c_double(tau), _p(S))
    return S


def valid_block(bed, n_ref, pos, z1, z2):
    """One block of the external-validation tool `valid` (reference scr/validate.cpp:225-259): the SNPs are decoded and
    standardised like in the fit (readSNPIm + nomalizeVec, :249-251), Sigma = X'X / n WITHOUT the tau shrinkage
    (:255-256), nume = z1'z2 (:257), deno = z1' Sigma z1 (:258).  Returns (nume, deno)."""
    z1 = np.asarray(z1, np.float64)
    z2 = np.asarray(z2, np.float64)
    if z1.size == 0:
        return 0.0, 0.0
    S = sigma(bed, n_ref, pos, tau=1.0)
    return float(z1 @ z2), float(z1 @ S @ z1)


def valid_run(dbslmm_txt, ext_txt, bim_txt, block_txt, bed, n_ref, maf_max):
    """File-level restatement of the reference's `valid` tool (scr/validate.cpp:120-265 with the readers and matchers it
    calls: readDBSLMM dtpr.cpp:223-245, readExt :248-270, matchSumm :411-434, readBim :125-165, matchAll :436-453,
    the sequential block scan validate.cpp:225-259).  Returns (nume[num_block], deno[num_block])."""
    dbs = [ln.split(" ") for ln in dbslmm_txt.strip().split("\n") if ln]
    ext = {}
    for ln in ext_txt.strip().split("\n"):
        t = ln.split(" ")
        if len(t) >= 4 and t[0] not in ext:                         # map::insert keeps the first
            ext[t[0]] = (t[1], float(t[2]), float(t[3]))
    comb = []
    for t in dbs:
        if t[0] in ext:
            a1e, mafe, ze = ext[t[0]]
            comb.append((t[0], t[1], mafe, float(t[2]), ze if a1e == t[1] else -ze))
    n_snp = bed.shape[0]
    constr = not abs(maf_max - 1.0) < 1e-10
    maf = snp_maf(bed, n_snp, n_ref) if constr else np.zeros(n_snp)
    bim = {}
    for i, ln in enumerate(bim_txt.strip().split("\n")):
        t = ln.split("\t")
        if t[1] not in bim:
            bim[t[1]] = (i, int(t[3]), t[4], maf[i])
    rows = []
    for snp, a1, mafe, z1, z2 in comb:
        if snp in bim and bim[snp][2] == a1 and abs(bim[snp][3] - mafe) < maf_max:
            rows.append((bim[snp][1], bim[snp][0], z1, z2))
    rows.sort(key=lambda r: r[0])
    blocks = [tuple(int(x) for x in ln.split("\t")[1:3]) for ln in block_txt.strip().split("\n")]
    nume, deno = np.zeros(len(blocks)), np.zeros(len(blocks))
    cur = 0
    for b, (start, end) in enumerate(blocks):
        pos, z1, z2 = [], [], []
        while cur < len(rows) and start <= rows[cur][0] < end:     # stops at the first SNP outside the block, like the reference
            pos.append(rows[cur][1]); z1.append(rows[cur][2]); z2.append(rows[cur][3])
            cur += 1
        nume[b], deno[b] = valid_block(bed, n_ref, np.asarray(pos, np.int32), z1, z2)
    return nume, deno


def pcgv(A, b, maxiter=1000, tol=1e-7):
    A = np.asfortranarray(A, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros_like(b)
    it = lib().orc_pcgv(_p(A), _p(b), C.c_int(b.size), C.c_int(maxiter), C.c_double(tol), _p(x))
    return x, it


def est_block(bed, n_ref, n_obs, sigma_s, pos_s, z_s, pos_l=None, z_l=None, mode=MODE_REF, tau=0.8):
    bed = _bed(bed)
    pos_s = np.ascontiguousarray(pos_s, dtype=np.int32)
    z_s = np.ascontiguousarray(z_s, dtype=np.float64)
    ms = pos_s.size
    ml = 0 if pos_l is None else len(pos_l)
    if ml:
        pos_l = np.ascontiguousarray(pos_l, dtype=np.int32)
        z_l = np.ascontiguousarray(z_l, dtype=np.float64)
    else:
        pos_l = z_l = None
    bs = np.zeros(ms, np.float64)
    bl = np.zeros(max(ml, 1), np.float64)
    sing = C.c_int(0)
    it = lib().orc_est_block(_p(bed), C.c_int(n_ref), C.c_int(n_obs), C.c_double(sigma_s), C.c_double(tau),
                             _p(pos_s), _p(z_s), C.c_int(ms), _p(pos_l), _p(z_l), C.c_int(ml), C.c_int(mode),
                             _p(bs), _p(bl), C.byref(sing))
    return bs, bl[:ml], it, sing.value


def est(bed, n_ref, n_obs, sigma_s, s_off, s_pos, s_z, l_off=None, l_pos=None, l_z=None,
        threads=1, mode=MODE_REF, tau=0.8):
    """DBSLMMFIT::est over CSR blocks. Returns (beta_s, beta_l, n_singular, max_iters)."""
    bed = _bed(bed)
    s_off = np.ascontiguousarray(s_off, dtype=np.int32)
    s_pos = np.ascontiguousarray(s_pos, dtype=np.int32)
    s_z = np.ascontiguousarray(s_z, dtype=np.float64)
    nb = s_off.size - 1
    bs = np.zeros(s_pos.size, np.float64)
    if l_off is not None:
        l_off = np.ascontiguousarray(l_off, dtype=np.int32)
        l_pos = np.ascontiguousarray(l_pos, dtype=np.int32)
        l_z = np.ascontiguousarray(l_z, dtype=np.float64)
        bl = np.zeros(max(l_pos.size, 1), np.float64)
        nl = l_pos.size
    else:
        l_pos = l_z = None
        bl = np.zeros(1, np.float64)
        nl = 0
    mi = C.c_int(0)
    sing = lib().orc_est(_p(bed), C.c_int(n_ref), C.c_int(n_obs), C.c_double(sigma_s), C.c_double(tau),
                         C.c_int(nb), _p(s_off), _p(s_pos), _p(s_z), _p(l_off), _p(l_pos), _p(l_z),
                         C.c_int(threads), C.c_int(mode), _p(bs), _p(bl), C.byref(mi))
    return bs, bl[:nl], sing, mi.value


def num_threads():
    return lib().orc_num_threads()


def variance_block(bed, n_ref, tbed, n_test_total, indicator, n_obs, sigma_s, pos_s, tpos_s, pos_l=None, tpos_l=None, tau=0.8):
    """The fork's per-block prediction-variance diagonal (calc_nt_by_nt_matrix(...).diag())."""
    bed = _bed(bed); tbed = _bed(tbed)
    ind = np.ascontiguousarray(indicator, np.int32)
    pos_s = np.ascontiguousarray(pos_s, np.int32); tpos_s = np.ascontiguousarray(tpos_s, np.int32)
    ml = 0 if pos_l is None else len(pos_l)
    if ml:
        pos_l = np.ascontiguousarray(pos_l, np.int32); tpos_l = np.ascontiguousarray(tpos_l, np.int32)
    else:
        pos_l = tpos_l = None
    out = np.zeros(int(ind.sum()), np.float64)
    rc = lib().orc_variance_block(_p(bed), C.c_int(n_ref), _p(tbed), C.c_int(n_test_total), _p(ind), C.c_int(n_obs),
                                  C.c_double(sigma_s), C.c_double(tau), _p(pos_s), _p(tpos_s), C.c_int(pos_s.size),
                                  _p(pos_l), _p(tpos_l), C.c_int(ml), _p(out))
    if rc:
        raise RuntimeError(f"orc_variance_block: not positive definite ({rc})")
    return out
