#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r3l_$name.json 2> gpurun_out/r3l_$name.err; python tools/bench_brief.py gpurun_out/r3l_$name.json; }
run base1 X=1
run f03r7a DBSLMM_B200_FIRST_REGION=0.3 DBSLMM_B200_REGIONS=7
run base2 X=1
run f03r7b DBSLMM_B200_FIRST_REGION=0.3 DBSLMM_B200_REGIONS=7
run f04r7 DBSLMM_B200_FIRST_REGION=0.4 DBSLMM_B200_REGIONS=7
run f03r8 DBSLMM_B200_FIRST_REGION=0.3 DBSLMM_B200_REGIONS=8
run pairh1 DBSLMM_B200_GRAM=pair DBSLMM_B200_GRAM_HINT=1
run pairh3 DBSLMM_B200_GRAM=pair DBSLMM_B200_GRAM_HINT=3
run singleh3 DBSLMM_B200_GRAM_HINT=3
