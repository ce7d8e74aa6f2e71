#!/bin/bash
# round-2 session O: ncu of the 64-row panel kernel (bulk class 1, steps 1-7) + launch list of one fit
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity"
$CMD > gpurun_out/r2o_plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/r2o_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2o_launches.csv $CMD > gpurun_out/r2o_ncu1.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:chol_panel_tma -s 80 -c 7 -o gpurun_out/prof_r2o_panel_c1 $CMD > gpurun_out/r2o_ncu2.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep
