"""Stage-by-stage GPU-vs-oracle report (bring-up aid; the assertions live in tests/)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dbslmm_b200 import _abi, synth
from oracle import oracle as O

def main():
    eng = _abi.Engine(0)
    for name, sizes, n_ref, miss in [("plain", [300, 0, 1, 7, 64, 65, 128, 129, 200], 400, 0.0),
                                     ("missing", [150, 70, 5, 130], 500, 0.01),
                                     ("n2000", [646, 1100], 2000, 0.0)]:
        w = synth.make_workload(11, sizes, n_ref, missing_rate=miss, frac_large=0.01)
        bed = w["bed"]
        eng.load_bed(bed, n_ref)
        maf, nn = eng.snp_stats()
        maf_o = O.snp_maf(bed, bed.shape[0], n_ref)
        print(f"[{name}] maf maxdiff {np.abs(maf - maf_o).max():.3e}  nonmiss ok {np.array_equal(nn, (w['G'] >= 0).sum(1))}")
        for mode, lo in (("LMM", False), ("DBSLMM", True)):
            if lo:
                args = (w["s_off"], w["s_pos"], w["s_z"], w["l_off"], w["l_pos"], w["l_z"])
            else:
                # all SNPs small
                off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
                pos = np.arange(off[-1], dtype=np.int32)
                zz = np.zeros(off[-1]); zz[w["s_pos"]] = w["s_z"]; zz[w["l_pos"]] = w["l_z"]
                args = (off, pos, zz)
            sig = 0.5 / 1000.0
            t = time.time()
            r = eng.fit(*args, sigma_s=[sig], n_obs=2400, flags=_abi.FLAG_KEEP_INT_GRAM)
            print(f"[{name}/{mode}] fit rc n_bad={r['n_bad']} wall {time.time()-t:.3f}s timing {r['timing']}")
            # decode check on first block rows
            Gz = np.where(w["G"] < 0, 0, w["G"])
            s_off = args[0]
            row = 0
            okc = True
            for b in range(len(sizes)):
                ps = args[1][s_off[b]:s_off[b + 1]]
                pl = args[4][args[3][b]:args[3][b + 1]] if lo else np.zeros(0, np.int32)
                pos_b = np.concatenate([ps, pl]).astype(np.int32)
                m = pos_b.size
                if m == 0:
                    continue
                has_miss = bool((w["G"][pos_b] < 0).any())
                for j in (0, m - 1):
                    c = eng.row_codes(b, j, n_ref)
                    okc &= np.array_equal(c, Gz[pos_b[j]])
                    if has_miss:
                        mk = eng.row_codes(b, j, n_ref, plane=1)
                        okc &= np.array_equal(mk, (w["G"][pos_b[j]] >= 0).astype(np.int8))
                Q, A, N = eng.block_gram(b, m)
                Qo, Ao, No = O.gram_int(bed, n_ref, pos_b)
                S = eng.block_sigma(b, m)
                So = O.sigma(bed, n_ref, pos_b)
                print(f"   block {b} m={m} miss={has_miss} Q {np.array_equal(Q, Qo)} A {np.array_equal(A, Ao)} N {np.array_equal(N, No)} sigma maxdiff {np.abs(S - So).max():.3e}")
                row += m * (2 if has_miss else 1)
            print(f"   codes ok {okc}")
            if lo:
                bs, bl, _, _ = O.est(bed, n_ref, 2400, sig, *args, threads=8, mode=1)
                ds = np.abs(r["beta_s"][0] - bs).max() / np.abs(bs).max()
                dl = np.abs(r["beta_l"][0] - bl).max() / max(np.abs(bl).max(), 1e-300) if bl.size else 0.0
                print(f"   beta_s rel {ds:.3e} beta_l rel {dl:.3e}")
            else:
                bs, _, _, _ = O.est(bed, n_ref, 2400, sig, *args, threads=8, mode=1)
                ds = np.abs(r["beta_s"][0] - bs).max() / np.abs(bs).max()
                print(f"   beta_s rel {ds:.3e}")
    eng.close()

if __name__ == "__main__":
    main()
