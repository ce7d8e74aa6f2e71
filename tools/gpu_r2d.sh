#!/bin/bash
# round-2 session D: CLI genome-wide cold run, ingest changes, 8-GPU shard emulation, launch list of one fit
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
name,path=sys.argv[1],sys.argv[2]
try:
    d=json.loads(open(path).read().strip().splitlines()[-1])
    o=d["rooflines_other"]
    print(name, "ms/step", round(d["ms_per_step"],3), "chol_ms", round(d["roofline"]["ms_per_step"],3), "frac", round(d["roofline"]["frac"],4),
          "e2e ms", round(d["e2e"]["ms_per_step"],2), "class", [round(x,2) for x in o["chol_class_ms"]], "dec", round(o["decode"]["ms"],3), "gram", round(o["gram"]["ms"],3), "h2d", round(o["h2d_ms"],3), "d2h", round(o["d2h_ms"],3))
except Exception as e: print(name, "parse failed", e)
PY
}
timeout 900 python -m pytest tests/test_cli.py tests/test_gpu_streaming.py tests/test_gpu_modes.py -q -m gpu -x > gpurun_out/r2d_pytest1.log 2>&1; echo "pytest1 rc=$?"; tail -4 gpurun_out/r2d_pytest1.log
timeout 900 python tools/cli_genome_wide.py --gpus 1 --json gpurun_out/cli_genome_wide.json > gpurun_out/r2d_cli.log 2>&1; echo "cli rc=$?"; tail -12 gpurun_out/r2d_cli.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"; show default gpurun_out/r2d_bench.json
for cfg in "s8_default:" "s8_pdl05:DBSLMM_B200_PDL=0.5" "s8_pdl02:DBSLMM_B200_PDL=0.2" "s8_tpc1:DBSLMM_B200_TPC=1" "s8_legacy:DBSLMM_B200_PANEL=legacy"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --emulate-shard 0/8 > gpurun_out/r2d_$name.json 2> gpurun_out/r2d_$name.err
  echo "$name rc=$?"; show $name gpurun_out/r2d_$name.json
done
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'chol_|gram_|decode_rows|backsolve|fill_z|rows_missing|block_missing|snp_stats' -c 260 --csv --log-file gpurun_out/r2d_launches.csv $CMD > gpurun_out/r2d_ncu1.log 2>&1
echo "ncu launches rc=$?"
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2d_pytest.log
