#!/bin/bash
# round-2 session I: fused unpack + Gram from packed 2-bit rows
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
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_modes.py tests/test_gpu_streaming.py -q -m gpu -x > gpurun_out/r2i_pytest1.log 2>&1; echo "pytest1 rc=$?"; tail -8 gpurun_out/r2i_pytest1.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_packed.json 2> gpurun_out/r2i_packed.err; echo "packed rc=$?"; show packed gpurun_out/r2i_packed.json
DBSLMM_B200_GRAM=codes timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2i_codes.json 2> gpurun_out/r2i_codes.err; echo "codes rc=$?"; show codes gpurun_out/r2i_codes.json
timeout 600 python bench.py --steps 5 --warmup 3 --missing 0.005 --no-cpu-baseline > gpurun_out/r2i_missing.json 2> gpurun_out/r2i_missing.err; echo "missing rc=$?"; show missing gpurun_out/r2i_missing.json
timeout 900 python bench.py --steps 3 --warmup 3 --config c5 --no-cpu-baseline --no-parity > gpurun_out/r2i_c5.json 2> gpurun_out/r2i_c5.err; echo "c5 rc=$?"; show c5 gpurun_out/r2i_c5.json
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2i_pytest.log
