"""ctypes binding of oracle/_ref/libref_harness.so -- the UNMODIFIED reference classes compiled
over oracle/shim (TEST INFRASTRUCTURE; exists only where oracle/build_ref.sh has run)."""
import ctypes as C
import os
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libref_harness.so")
CLI = os.path.join(_HERE, "_ref", "dbslmm_ref")
_lib = None


def available():
    return os.path.exists(SO)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(SO)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def write_bed(bed, path):
    with open(path, "wb") as f:
        f.write(bytes([0x6C, 0x1B, 0x01]))
        f.write(np.ascontiguousarray(bed, np.uint8).tobytes())


class BedFile:
    """The reference reads genotypes through an ifstream: give it a real file."""

    def __init__(self, bed):
        self.tmp = tempfile.NamedTemporaryFile(suffix=".bed", delete=False)
        self.tmp.close()
        write_bed(bed, self.tmp.name)
        self.path = self.tmp.name

    def close(self):
        try:
            os.unlink(self.path)
        except OSError:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def read_snp_im(bed_path, pos, n_total, indicator=None):
    ind = np.ones(n_total, np.int32) if indicator is None else np.ascontiguousarray(indicator, np.int32)
    g = np.zeros(int((ind != 0).sum()), np.float64)
    maf = C.c_double(0)
    lib().ref_read_snp_im(bed_path.encode(), C.c_int(pos), C.c_int(n_total), _p(ind), _p(g), C.byref(maf))
    return g, maf.value


def normalize(x):
    x = np.array(x, np.float64, copy=True)
    lib().ref_normalize(_p(x), C.c_int(x.size))
    return x


def pcgv(A, b, maxiter=1000, tol=1e-7):
    A = np.asfortranarray(A, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    x = np.zeros_like(b)
    lib().ref_pcgv(_p(A), _p(b), C.c_int(b.size), C.c_int(maxiter), C.c_double(tol), _p(x))
    return x


def est_block(bed_path, n_ref, n_obs, sigma_s, pos_s, z_s, pos_l=None, z_l=None):
    pos_s = np.ascontiguousarray(pos_s, np.int32)
    z_s = np.ascontiguousarray(z_s, np.float64)
    ml = 0 if pos_l is None else len(pos_l)
    if ml:
        pos_l = np.ascontiguousarray(pos_l, np.int32)
        z_l = np.ascontiguousarray(z_l, np.float64)
    else:
        pos_l = z_l = None
    bs = np.zeros(pos_s.size)
    bl = np.zeros(max(ml, 1))
    lib().ref_est_block(bed_path.encode(), C.c_int(n_ref), C.c_int(n_obs), C.c_double(sigma_s), _p(pos_s), _p(z_s),
                        C.c_int(pos_s.size), _p(pos_l), _p(z_l), C.c_int(ml), _p(bs), _p(bl))
    return bs, bl[:ml]


def est_path(bed_path, n_ref, n_obs, sigma_s, s_off, s_pos, s_z, l_off=None, l_pos=None, l_z=None, threads=1):
    s_off = np.ascontiguousarray(s_off, np.int32)
    s_pos = np.ascontiguousarray(s_pos, np.int32)
    s_z = np.ascontiguousarray(s_z, np.float64)
    bs = np.zeros(s_pos.size)
    if l_off is not None:
        l_off = np.ascontiguousarray(l_off, np.int32)
        l_pos = np.ascontiguousarray(l_pos, np.int32)
        l_z = np.ascontiguousarray(l_z, np.float64)
        bl = np.zeros(max(l_pos.size, 1))
        nl = l_pos.size
    else:
        bl = np.zeros(1)
        nl = 0
    lib().ref_est_path(bed_path.encode(), C.c_int(n_ref), C.c_int(n_obs), C.c_double(sigma_s), C.c_int(s_off.size - 1),
                       _p(s_off), _p(s_pos), _p(s_z), _p(l_off), _p(l_pos), _p(l_z), C.c_int(threads), _p(bs), _p(bl))
    return bs, bl[:nl]


def variance_block(bed_path, n_ref, tbed_path, n_test_total, indicator, n_obs, sigma_s, pos_s, tpos_s, pos_l=None, tpos_l=None):
    ind = np.ascontiguousarray(indicator, np.int32)
    pos_s = np.ascontiguousarray(pos_s, np.int32); tpos_s = np.ascontiguousarray(tpos_s, np.int32)
    ml = 0 if pos_l is None else len(pos_l)
    if ml:
        pos_l = np.ascontiguousarray(pos_l, np.int32); tpos_l = np.ascontiguousarray(tpos_l, np.int32)
    else:
        pos_l = tpos_l = None
    out = np.zeros(int(ind.sum()), np.float64)
    lib().ref_variance_block(bed_path.encode(), C.c_int(n_ref), tbed_path.encode(), C.c_int(n_test_total), _p(ind),
                             C.c_int(n_obs), C.c_double(sigma_s), _p(pos_s), _p(tpos_s), C.c_int(pos_s.size),
                             _p(pos_l), _p(tpos_l), C.c_int(ml), _p(out))
    return out
