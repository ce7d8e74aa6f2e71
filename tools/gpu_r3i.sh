#!/bin/bash
mkdir -p gpurun_out
for k in single pair single pair; do
  DBSLMM_B200_GRAM=$k timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r3i_$k.json 2>/dev/null; python tools/bench_brief.py gpurun_out/r3i_$k.json
done
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity"
for k in single pair; do
  DBSLMM_B200_GRAM=$k timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'gram_p' -c 1 --csv --log-file gpurun_out/r3i_ncu_$k.csv $CMD > /dev/null 2>&1; grep -v "^==" gpurun_out/r3i_ncu_$k.csv | cut -d, -f5,13- | tail -5
done
