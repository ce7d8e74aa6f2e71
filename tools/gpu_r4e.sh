#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r4e_$name.json 2> gpurun_out/r4e_$name.err; echo -n "$name: "; python tools/bench_brief.py gpurun_out/r4e_$name.json | sed 's/.*e2e_ms=\([0-9.]*\).*/e2e \1/'; }
run base X=1
run f05 DBSLMM_B200_FIRST_REGION=0.5
run f03 DBSLMM_B200_FIRST_REGION=0.3
run f03r7 DBSLMM_B200_FIRST_REGION=0.3 DBSLMM_B200_REGIONS=7
run f05r7 DBSLMM_B200_FIRST_REGION=0.5 DBSLMM_B200_REGIONS=7
run base X=1
run f03r7p30 DBSLMM_B200_FIRST_REGION=0.3 DBSLMM_B200_REGIONS=7 DBSLMM_B200_PREPLAN_MB=30
run f02r7p20 DBSLMM_B200_FIRST_REGION=0.2 DBSLMM_B200_REGIONS=7 DBSLMM_B200_PREPLAN_MB=20
run f03r8 DBSLMM_B200_FIRST_REGION=0.3 DBSLMM_B200_REGIONS=8
run f04r7b2 DBSLMM_B200_FIRST_REGION=0.4 DBSLMM_B200_REGIONS=7 DBSLMM_B200_UPLOAD_BULK_FIRST=2
