#!/bin/bash
# round-2 session B: TMA panel kernel on the genome-wide workload (items-per-CTA sweep), full GPU test-suite, new bench lines
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
name,path=sys.argv[1],sys.argv[2]
try:
    d=json.loads(open(path).read().strip().splitlines()[-1])
    o=d["rooflines_other"]
    print(name, "ms/step", round(d["ms_per_step"],3), "chol_ms", round(d["roofline"]["ms_per_step"],3), "frac", round(d["roofline"]["frac"],4),
          "e2e ms", round(d["e2e"]["ms_per_step"],2), "class", [round(x,2) for x in o["chol_class_ms"]], "dec", round(o["decode"]["ms"],3), "gram", round(o["gram"]["ms"],3),
          "parity", d.get("parity",{}).get("max_rel_vs_exact_oracle"), d.get("parity",{}).get("max_rel_vs_reference"), d.get("parity",{}).get("gram_bit_exact"))
except Exception as e: print(name, "parse failed", e)
PY
}
for cfg in "default:" "tpc1:DBSLMM_B200_TPC=1" "tpc2w2:DBSLMM_B200_TPC=2,2" "tpc2w1:DBSLMM_B200_TPC=2,1" "tpc4w2:DBSLMM_B200_TPC=4,2" "tpc8w1:DBSLMM_B200_TPC=8,1" "perm0:DBSLMM_B200_TMAP_PERM=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2b_bench_$name.json 2> gpurun_out/r2b_bench_$name.err
  echo "$name rc=$?"; show $name gpurun_out/r2b_bench_$name.json
done
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2b_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_full.json 2> gpurun_out/r2b_full.err; echo "full rc=$?"; show full gpurun_out/r2b_full.json
timeout 600 python bench.py --steps 5 --warmup 3 --missing 0.005 --no-cpu-baseline > gpurun_out/r2b_missing.json 2> gpurun_out/r2b_missing.err; echo "missing rc=$?"; show missing gpurun_out/r2b_missing.json
timeout 900 python bench.py --steps 5 --warmup 3 --config c4 > gpurun_out/r2b_c4.json 2> gpurun_out/r2b_c4.err; echo "c4 rc=$?"; show c4 gpurun_out/r2b_c4.json
timeout 900 python bench.py --steps 3 --warmup 3 --config c5 --no-cpu-baseline > gpurun_out/r2b_c5.json 2> gpurun_out/r2b_c5.err; echo "c5 rc=$?"; show c5 gpurun_out/r2b_c5.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2b_ref.json 2> gpurun_out/r2b_ref.err; echo "ref rc=$?"; cat gpurun_out/r2b_ref.json | cut -c1-600
