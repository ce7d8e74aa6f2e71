#!/bin/bash
# session 3e: CTA-pair Gram with 16 epilogue warps (hand-over before the FP64 work)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_streaming.py -x -q -m gpu > gpurun_out/r3e_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r3e_tests.log
for cfg in pair:0 single:0; do
  k=${cfg%%:*}; hint=${cfg##*:}
  DBSLMM_B200_GRAM=$k DBSLMM_B200_GRAM_HINT=$hint timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r3e_${k}_${hint}.json 2> gpurun_out/r3e_${k}_${hint}.err; echo "$cfg rc=$?"; python tools/bench_brief.py gpurun_out/r3e_${k}_${hint}.json
done
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_pair -c 1 -o gpurun_out/r3e_gram_pair $CMD > gpurun_out/r3e_ncu.log 2>&1; echo "ncu rc=$?"
