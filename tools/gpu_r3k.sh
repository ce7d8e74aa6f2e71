#!/bin/bash
mkdir -p gpurun_out
DBSLMM_B200_GRAM=pair timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_streaming.py -x -q -m gpu 2>&1 | tail -2
for k in pair single pair; do
  DBSLMM_B200_GRAM=$k timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r3k_$k.json 2>/dev/null; python tools/bench_brief.py gpurun_out/r3k_$k.json
done
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity"
DBSLMM_B200_GRAM=pair timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_pair -c 1 -o gpurun_out/r3k_gram_pair $CMD > gpurun_out/r3k_ncu.log 2>&1; echo "ncu rc=$?"
