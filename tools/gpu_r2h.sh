#!/bin/bash
# round-2 session H: fit_multi, upload order A/B, traced streaming fit, c5 line
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
name,path=sys.argv[1],sys.argv[2]
try:
    d=json.loads(open(path).read().strip().splitlines()[-1])
    o=d["rooflines_other"]
    print(name, "ms/step", round(d["ms_per_step"],3), "chol_ms", round(d["roofline"]["ms_per_step"],3), "frac", round(d["roofline"]["frac"],4),
          "e2e ms", round(d["e2e"]["ms_per_step"],2), "2call", round(d["e2e"]["upload_then_fit_ms_per_step"],2), "dec", round(o["decode"]["ms"],3), "gram", round(o["gram"]["ms"],3),
          "parity", d.get("parity",{}).get("max_rel_vs_exact_oracle"), d.get("parity",{}).get("gram_bit_exact"))
except Exception as e: print(name, "parse failed", e)
PY
}
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2h_pytest.log
for cfg in "bulk1:DBSLMM_B200_UPLOAD_BULK_FIRST=1" "bulk0:DBSLMM_B200_UPLOAD_BULK_FIRST=0" "bulk2:DBSLMM_B200_UPLOAD_BULK_FIRST=2" "bulk4:DBSLMM_B200_UPLOAD_BULK_FIRST=4"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2h_$name.json 2> gpurun_out/r2h_$name.err
  echo "$name rc=$?"; show $name gpurun_out/r2h_$name.json
done
DBSLMM_B200_TRACE=1 timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-parity > /dev/null 2> gpurun_out/r2h_trace.err; grep -n "trace" gpurun_out/r2h_trace.err | tail -40
timeout 900 python bench.py --steps 3 --warmup 3 --config c5 --no-cpu-baseline > gpurun_out/r2h_c5.json 2> gpurun_out/r2h_c5.err; echo "c5 rc=$?"; show c5 gpurun_out/r2h_c5.json
