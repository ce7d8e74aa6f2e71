#!/bin/bash
# round-2 session C: ncu launch list of one fit + full capture of the bulk class's panel steps (TMA kernel)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity"
$CMD > gpurun_out/r2c_plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/r2c_plain.log; exit 1; }
tail -c 600 gpurun_out/r2c_plain.log; echo
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_launches.csv $CMD > gpurun_out/r2c_ncu1.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:chol_panel_tma -s 79 -c 16 -o gpurun_out/prof_r2c_panel_c1 $CMD > gpurun_out/r2c_ncu2.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep
