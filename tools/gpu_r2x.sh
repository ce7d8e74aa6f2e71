#!/bin/bash
# round-2 session X: L2 policy / tile hand-out experiments on the one-plane Gram kernels (Gram ms of bench.py, DRAM bytes by ncu)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2x_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2x_tests.log
DBSLMM_B200_GRAM=single DBSLMM_B200_GRAM_HINT=4 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2x_tests_dyn.log 2>&1; echo "tests dyn rc=$?"; tail -3 gpurun_out/r2x_tests_dyn.log
for cfg in single:0 single:1 single:2 single:4 single:5 single:7 pair:0 pair:1 pair:3; do
  k=${cfg%%:*}; hint=${cfg##*:}
  DBSLMM_B200_GRAM=$k DBSLMM_B200_GRAM_HINT=$hint timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2x_${k}_${hint}.json 2> gpurun_out/r2x_${k}_${hint}.err; echo "$cfg rc=$?"; python tools/bench_brief.py gpurun_out/r2x_${k}_${hint}.json
done
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity"
for cfg in single:4 single:7 pair:3; do
  k=${cfg%%:*}; hint=${cfg##*:}
  DBSLMM_B200_GRAM=$k DBSLMM_B200_GRAM_HINT=$hint timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:'gram_p' -c 2 --csv --log-file gpurun_out/r2x_ncu_${k}_${hint}.csv $CMD > /dev/null 2>&1; echo "ncu $cfg rc=$?"; grep -v "^==" gpurun_out/r2x_ncu_${k}_${hint}.csv | cut -d, -f5,13- | tail -8
done
