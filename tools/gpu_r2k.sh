#!/bin/bash
# round-2 session K: HEAD validation -- full GPU suite, headline bench, streaming trace (host plan now overlapped)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2k_pytest.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2k_bench.json").read().strip().splitlines()[-1]); o=d["rooflines_other"]
print("value",round(d["value"]),"ms",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["ms_per_step"],2),"chol",round(d["roofline"]["ms_per_step"],3),round(d["roofline"]["frac"],4),"dec",o["decode"]["ms"],"gram",o["gram"]["ms"],"cpu",d.get("cpu_baseline",{}).get("value"))
PY
DBSLMM_B200_TRACE=1 timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-parity > /dev/null 2> gpurun_out/r2k_trace.err; grep -n "trace" gpurun_out/r2k_trace.err | tail -12
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2k_ref.json 2> gpurun_out/r2k_ref.err; echo "ref rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2k_ref.json').read().strip().splitlines()[-1]);print('ref',d['value'],d['ms_per_step'],d['cpu_baseline']['cores'])"
