"""Resident fit vs streaming fit on the genome-wide workload: per-block differences (debugging aid)."""
import argparse, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dbslmm_b200 import _abi
from oracle import oracle as O

ns = argparse.Namespace(config="c3", missing=0.0, seed=20240003)
dev = torch.device("cuda", 0)
w = bench.build_workload(ns, torch, dev, ns.seed)
sz = w["z"][w["s_pos"]]; lz = w["z"][w["l_pos"]]
csr = (w["s_off"], w["s_pos"], sz, w["l_off"], w["l_pos"], lz)
sig, n_obs = 0.5 / w["nsnp_total"], w["n_obs"]
eng = _abi.Engine(0)
eng.load_bed(w["bed"], w["n_ref"])
r = eng.fit(*csr, sigma_s=[sig], n_obs=n_obs)
rs = eng.fit(*csr, sigma_s=[sig], n_obs=n_obs, bed=w["bed"], n_ref=w["n_ref"])
r2 = eng.fit(*csr, sigma_s=[sig], n_obs=n_obs)
for name, x, y in (("stream vs resident", rs, r), ("resident again vs resident", r2, r)):
    bad = []
    for b in range(len(w["sizes"])):
        s0, s1 = w["s_off"][b], w["s_off"][b + 1]
        if s1 == s0: continue
        d = np.abs(x["beta_s"][0][s0:s1] - y["beta_s"][0][s0:s1]).max() / max(np.abs(y["beta_s"][0][s0:s1]).max(), 1e-300)
        if d > 1e-11: bad.append((b, int(w["sizes"][b]), float(d)))
    print(name, "n_bad blocks", len(bad), bad[:12], flush=True)
# exact oracle on whichever blocks differ (first 3) -> which side is wrong
bad_b = [b for b, _, _ in bad[:0]]
