#!/bin/bash
# session 3r: upload order of the streaming fit (small first region, bulk regions ahead of the big classes)
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r3r_$name.json 2> gpurun_out/r3r_$name.err; echo -n "$name: "; python tools/bench_brief.py gpurun_out/r3r_$name.json | sed 's/.*e2e_ms=\([0-9.]*\).*/e2e \1/'; }
run base X=1
run f03r7b2 DBSLMM_B200_FIRST_REGION=0.3 DBSLMM_B200_REGIONS=7 DBSLMM_B200_UPLOAD_BULK_FIRST=2
run f03r7b3 DBSLMM_B200_FIRST_REGION=0.3 DBSLMM_B200_REGIONS=7 DBSLMM_B200_UPLOAD_BULK_FIRST=3
run base X=1
run f05r7b2 DBSLMM_B200_FIRST_REGION=0.5 DBSLMM_B200_REGIONS=7 DBSLMM_B200_UPLOAD_BULK_FIRST=2
run f03r8b2 DBSLMM_B200_FIRST_REGION=0.3 DBSLMM_B200_REGIONS=8 DBSLMM_B200_UPLOAD_BULK_FIRST=2
run r6b2 DBSLMM_B200_UPLOAD_BULK_FIRST=2
run f02r7b2 DBSLMM_B200_FIRST_REGION=0.2 DBSLMM_B200_REGIONS=7 DBSLMM_B200_UPLOAD_BULK_FIRST=2
run f03r7b2p40 DBSLMM_B200_FIRST_REGION=0.3 DBSLMM_B200_REGIONS=7 DBSLMM_B200_UPLOAD_BULK_FIRST=2 DBSLMM_B200_PREPLAN_MB=40
