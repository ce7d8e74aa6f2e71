#!/bin/bash
# session 3a: streaming fit lead-in -- size of the first bulk region, region count, pre-plan limit (e2e ms of bench.py + traces)
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r3a_$name.json 2> gpurun_out/r3a_$name.err; echo "$name rc=$?"; python tools/bench_brief.py gpurun_out/r3a_$name.json; }
run base X=1
run f05 DBSLMM_B200_FIRST_REGION=0.5
run f03 DBSLMM_B200_FIRST_REGION=0.3
run f03r7 DBSLMM_B200_FIRST_REGION=0.3 DBSLMM_B200_REGIONS=7
run f02r7 DBSLMM_B200_FIRST_REGION=0.2 DBSLMM_B200_REGIONS=7
run f03r7p30 DBSLMM_B200_FIRST_REGION=0.3 DBSLMM_B200_REGIONS=7 DBSLMM_B200_PREPLAN_MB=30
run f03r7b0 DBSLMM_B200_FIRST_REGION=0.3 DBSLMM_B200_REGIONS=7 DBSLMM_B200_UPLOAD_BULK_FIRST=0
DBSLMM_B200_FIRST_REGION=0.3 DBSLMM_B200_REGIONS=7 timeout 300 python tools/stream_trace.py 2> gpurun_out/r3a_trace.err; grep -A40 "streaming fit 1" gpurun_out/r3a_trace.err | head -45
