"""Host-side breakdown of the end-to-end step (load_bed + fit from host buffers)."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from dbslmm_b200 import _abi
sys.argv = ["bench.py"]
args = bench.parse()
dev = torch.device("cuda", 0)
w = bench.build_workload(args, torch, dev, args.seed)
owner = np.zeros(w["sizes"].size, np.int32)
sh = bench.shard_workload(w, owner, 0, torch)
eng = _abi.Engine(0)
csr = (sh["s_off"], sh["s_pos"], sh["s_z"], sh["l_off"], sh["l_pos"], sh["l_z"])
kw = dict(sigma_s=[0.5 / w["nsnp_total"]], n_obs=300000)
for i in range(4):
    t0 = time.perf_counter(); eng.load_bed(sh["bed"], 2000); t1 = time.perf_counter()
    r = eng.fit(*csr, **kw); t2 = time.perf_counter()
    r2 = eng.fit(*csr, flags=_abi.FLAG_PLAN_CACHED, **kw); t3 = time.perf_counter()
    t = r["timing"]
    print(f"load_bed {1e3*(t1-t0):.2f} ms | fit(new plan) {1e3*(t2-t1):.2f} ms [device total {t['total_ms']:.2f}: h2d {t['h2d_ms']:.2f} dec {t['decode_ms']:.2f} gram {t['gram_ms']:.2f} solve {t['solve_ms']:.2f} d2h {t['d2h_ms']:.2f}] | fit(cached plan) {1e3*(t3-t2):.2f} ms [device {r2['timing']['total_ms']:.2f}]")
