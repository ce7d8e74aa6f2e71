#!/bin/bash
# session 4f (final): measurement set of the round -- headline bench, config lines, reference arm, launch list
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r4f_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r4f_tests.log
mkdir -p gpurun_out
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r4f_c3.json 2> gpurun_out/r4f_c3.err; echo "c3 rc=$?"; python tools/bench_brief.py gpurun_out/r4f_c3.json
timeout 400 python bench.py --steps 10 --warmup 3 --missing 0.005 --no-cpu-baseline > gpurun_out/r4f_missing.json 2> gpurun_out/r4f_missing.err; echo "missing rc=$?"; python tools/bench_brief.py gpurun_out/r4f_missing.json
timeout 600 python bench.py --steps 5 --warmup 3 --config c5 --no-cpu-baseline > gpurun_out/r4f_c5.json 2> gpurun_out/r4f_c5.err; echo "c5 rc=$?"; python tools/bench_brief.py gpurun_out/r4f_c5.json
timeout 600 python bench.py --steps 5 --warmup 3 --config c4 --no-cpu-baseline > gpurun_out/r4f_c4.json 2> gpurun_out/r4f_c4.err; echo "c4 rc=$?"; tail -c 700 gpurun_out/r4f_c4.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r4f_ref.json 2> gpurun_out/r4f_ref.err; echo "ref rc=$?"; tail -c 400 gpurun_out/r4f_ref.json
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'chol_|gram_|decode_rows|backsolve|fill_z|fill_rowmaps|block_flags|rows_missing|block_missing|snp_stats' -c 900 --csv --log-file gpurun_out/r4f_launches.csv $CMD > gpurun_out/r4f_ncu1.log 2>&1
echo "ncu launches rc=$?"; ls -la gpurun_out/r4f_launches.csv
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --emulate-shard 0/8 > gpurun_out/r4f_shard0of8.json 2> gpurun_out/r4f_shard0of8.err; echo "shard rc=$?"; python tools/bench_brief.py gpurun_out/r4f_shard0of8.json
