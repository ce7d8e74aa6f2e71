#!/bin/bash
# round-2 session F: stacked four-plane Gram, PRS kernel + prefetched validation panel (c4), tightened bars, m=3000 oracle test
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
name,path=sys.argv[1],sys.argv[2]
try:
    d=json.loads(open(path).read().strip().splitlines()[-1])
    o=d["rooflines_other"]
    print(name, "ms/step", round(d["ms_per_step"],3), "chol_ms", round(d["roofline"]["ms_per_step"],3), "frac", round(d["roofline"]["frac"],4),
          "e2e ms", round(d["e2e"]["ms_per_step"],2), "2call", round(d["e2e"]["upload_then_fit_ms_per_step"],2), "dec", round(o["decode"]["ms"],3), "gram", round(o["gram"]["ms"],3), "prs", o.get("prs",{}).get("ms"),
          "parity", d.get("parity",{}).get("max_rel_vs_exact_oracle"), d.get("parity",{}).get("gram_bit_exact"))
except Exception as e: print(name, "parse failed", e)
PY
}
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2f_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --missing 0.005 --no-cpu-baseline > gpurun_out/r2f_missing.json 2> gpurun_out/r2f_missing.err; echo "missing rc=$?"; show missing gpurun_out/r2f_missing.json
timeout 900 python bench.py --steps 5 --warmup 3 --config c4 --no-cpu-baseline > gpurun_out/r2f_c4.json 2> gpurun_out/r2f_c4.err; echo "c4 rc=$?"; show c4 gpurun_out/r2f_c4.json
