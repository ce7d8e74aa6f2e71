#!/bin/bash
# round-2 session L: 64-row panel items (128-thread CTAs, three per SM), W image, L2 prefetch -- correctness first, then A/B
mkdir -p gpurun_out
build/diag_probe > gpurun_out/r2l_diag_probe.txt 2>&1; tail -4 gpurun_out/r2l_diag_probe.txt
timeout 900 python -m pytest tests/test_gpu_modes.py tests/test_gpu_parity.py tests/test_gpu_streaming.py -x -q -m gpu > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2l_pytest.log
for cfg in "0 0" "0 4" "1 0" "1 4" "2 4" "1 8"; do
  set -- $cfg
  DBSLMM_B200_TILE64=$1 DBSLMM_B200_L2_PF=$2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2l_t$1_pf$2.json 2> gpurun_out/r2l_t$1_pf$2.err
  echo "tile64=$1 pf=$2 rc=$?"; python tools/bench_brief.py gpurun_out/r2l_t$1_pf$2.json
done
