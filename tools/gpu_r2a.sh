#!/bin/bash
# round-2 session A: validate the TMA panel kernel, then A/B it against the cp.async kernel
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
build/dmma_probe > gpurun_out/r2a_dmma.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_modes.py -x -q -m gpu -k "solver_modes" > gpurun_out/r2a_modes.log 2>&1
echo "modes rc=$?" >> gpurun_out/r2a_modes.log
tail -3 gpurun_out/r2a_modes.log
for cfg in "default:" "legacy:DBSLMM_B200_PANEL=legacy" "tpc1:DBSLMM_B200_TPC=1" "tpc2:DBSLMM_B200_TPC=2,2" "tpc4w2:DBSLMM_B200_TPC=4,2" "tpc8w1:DBSLMM_B200_TPC=8,1" "perm0:DBSLMM_B200_TMAP_PERM=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_$name.json 2> gpurun_out/r2a_bench_$name.err
  echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2a_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", "ms/step", round(d["ms_per_step"],3), "chol_ms", round(d["roofline"]["ms_per_step"],3), "frac", round(d["roofline"]["frac"],4), "e2e ms", round(d["e2e"]["ms_per_step"],2), "class", [round(x,2) for x in d["rooflines_other"]["chol_class_ms"]])
except Exception as e: print("$name parse failed", e)
PY
done
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2a_pytest.log
