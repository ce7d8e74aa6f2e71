#!/bin/bash
# round-2 session W: next-item chunk prefetch in the panel kernel (A/B), full ncu capture of both one-plane Gram kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_modes.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2w_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2w_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2w_c3_pf1.json 2> gpurun_out/r2w_c3_pf1.err; echo "c3 pf1 rc=$?"; python tools/bench_brief.py gpurun_out/r2w_c3_pf1.json
DBSLMM_B200_NEXT_PF=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2w_c3_pf0.json 2> gpurun_out/r2w_c3_pf0.err; echo "c3 pf0 rc=$?"; python tools/bench_brief.py gpurun_out/r2w_c3_pf0.json
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_pair -c 1 -o gpurun_out/r2w_gram_pair $CMD > gpurun_out/r2w_ncu1.log 2>&1; echo "ncu pair rc=$?"
DBSLMM_B200_GRAM=single timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_persistent -c 1 -o gpurun_out/r2w_gram_single $CMD > gpurun_out/r2w_ncu2.log 2>&1; echo "ncu single rc=$?"
ls -la gpurun_out/*.ncu-rep
