"""Host-side trace of the streaming fit (DBSLMM_B200_TRACE=1) on the bench workload."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DBSLMM_B200_TRACE"] = "1"
import bench
from dbslmm_b200 import _abi
shard = sys.argv[1] if len(sys.argv) > 1 else ""          # "R/N": trace rank R's shard of an N-GPU run
sys.argv = ["bench.py"]
args = bench.parse()
dev = torch.device("cuda", 0)
w = bench.build_workload(args, torch, dev, args.seed)
eng = _abi.Engine(0)
owner, my = np.zeros(w["sizes"].size, np.int32), 0
if shard:
    my, n = (int(x) for x in shard.split("/"))
    ms = (w["s_off"][1:] - w["s_off"][:-1]).astype(np.int32)
    ml = (w["l_off"][1:] - w["l_off"][:-1]).astype(np.int32)
    owner, _ = eng.plan_shards(ms, ml, 2000, n)
sh = bench.shard_workload(w, owner, my, torch)
csr = (sh["s_off"], sh["s_pos"], sh["s_z"], sh["l_off"], sh["l_pos"], sh["l_z"])
kw = dict(sigma_s=[0.5 / w["nsnp_total"]], n_obs=300000)
for i in range(3):
    t0 = time.perf_counter()
    r = eng.fit(*csr, bed=sh["bed"], n_ref=2000, **kw)
    print(f"--- streaming fit {i}: {1e3 * (time.perf_counter() - t0):.2f} ms", file=sys.stderr)
for i in range(2):
    t0 = time.perf_counter(); eng.load_bed(sh["bed"], 2000); t1 = time.perf_counter()
    r = eng.fit(*csr, **kw); t2 = time.perf_counter()
    print(f"--- load_bed {1e3*(t1-t0):.2f} + fit {1e3*(t2-t1):.2f} ms", file=sys.stderr)
