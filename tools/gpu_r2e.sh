#!/bin/bash
# round-2 session E: device-side missing-call flags (one decode pass, persistent four-plane Gram, no streaming fallback)
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
name,path=sys.argv[1],sys.argv[2]
try:
    d=json.loads(open(path).read().strip().splitlines()[-1])
    o=d["rooflines_other"]
    print(name, "ms/step", round(d["ms_per_step"],3), "chol_ms", round(d["roofline"]["ms_per_step"],3), "frac", round(d["roofline"]["frac"],4),
          "e2e ms", round(d["e2e"]["ms_per_step"],2), "2call", round(d["e2e"]["upload_then_fit_ms_per_step"],2), "class", [round(x,2) for x in o["chol_class_ms"]], "dec", round(o["decode"]["ms"],3), "gram", round(o["gram"]["ms"],3),
          "parity", d.get("parity",{}).get("max_rel_vs_exact_oracle"), d.get("parity",{}).get("gram_bit_exact"))
except Exception as e: print(name, "parse failed", e)
PY
}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_streaming.py -q -m gpu -x > gpurun_out/r2e_pytest1.log 2>&1; echo "pytest1 rc=$?"; tail -6 gpurun_out/r2e_pytest1.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"; show default gpurun_out/r2e_bench.json
timeout 600 python bench.py --steps 5 --warmup 3 --missing 0.005 --no-cpu-baseline > gpurun_out/r2e_missing.json 2> gpurun_out/r2e_missing.err; echo "missing rc=$?"; show missing gpurun_out/r2e_missing.json
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2e_pytest.log
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --emulate-shard 0/8"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'chol_|gram_|decode_rows|backsolve|fill_z|block_flags' -c 260 --csv --log-file gpurun_out/r2e_launches_shard.csv $CMD > gpurun_out/r2e_ncu1.log 2>&1
echo "ncu rc=$?"
