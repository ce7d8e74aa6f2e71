#!/bin/bash
# round-2 last session: smoke of the final library + ncu --set full of the four-plane Gram (0.5 % missing calls, full C3 size)
mkdir -p gpurun_out
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r6a_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r6a_smoke.log
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --missing 0.005"
timeout 90 ncu --set full --clock-control none --import-source on -k regex:gram_missing -s 1 -c 1 -o gpurun_out/prof_r6a_gram_missing $CMD > gpurun_out/r6a_ncu.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/r6a_ncu.log | cut -c1-300
