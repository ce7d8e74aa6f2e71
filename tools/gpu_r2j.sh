#!/bin/bash
# round-2 session J: fused Gram with 8 unpack warps vs int8-row Gram; host plan trace
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
name,path=sys.argv[1],sys.argv[2]
try:
    d=json.loads(open(path).read().strip().splitlines()[-1])
    o=d["rooflines_other"]
    print(name, "ms/step", round(d["ms_per_step"],3), "chol_ms", round(d["roofline"]["ms_per_step"],3), "e2e ms", round(d["e2e"]["ms_per_step"],2), "dec", round(o["decode"]["ms"],3), "gram", round(o["gram"]["ms"],3),
          "parity", d.get("parity",{}).get("max_rel_vs_exact_oracle"), d.get("parity",{}).get("gram_bit_exact"))
except Exception as e: print(name, "parse failed", e)
PY
}
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_modes.py -q -m gpu -x > gpurun_out/r2j_pytest1.log 2>&1; echo "pytest1 rc=$?"; tail -4 gpurun_out/r2j_pytest1.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_packed.json 2> gpurun_out/r2j_packed.err; echo "packed rc=$?"; show packed gpurun_out/r2j_packed.json
DBSLMM_B200_GRAM=codes timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2j_codes.json 2> gpurun_out/r2j_codes.err; echo "codes rc=$?"; show codes gpurun_out/r2j_codes.json
timeout 900 python bench.py --steps 3 --warmup 3 --config c5 --no-cpu-baseline --no-parity > gpurun_out/r2j_c5.json 2> gpurun_out/r2j_c5.err; echo "c5 rc=$?"; show c5 gpurun_out/r2j_c5.json
DBSLMM_B200_GRAM=codes DBSLMM_B200_TRACE=1 timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-parity > /dev/null 2> gpurun_out/r2j_trace.err; grep -n "trace" gpurun_out/r2j_trace.err | tail -24
