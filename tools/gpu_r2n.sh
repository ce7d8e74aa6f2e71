#!/bin/bash
# round-2 session N: stage release behind the wait loop (race fix), 64-row panel items -- full suite + A/B
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python tools/stream_vs_resident.py 2>&1 | tail -2 | cut -c1-120; }
run DBSLMM_B200_TILE64=2
run DBSLMM_B200_TILE64=2 DBSLMM_B200_TPC=1 DBSLMM_B200_FUSE_DIAG=0 DBSLMM_B200_SPLITK=1,1
run DBSLMM_B200_TILE64=0
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2n_pytest.log
b() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2n_$tag.json 2> gpurun_out/r2n_$tag.err; python tools/bench_brief.py gpurun_out/r2n_$tag.json; }
b t0 DBSLMM_B200_TILE64=0 DBSLMM_B200_L2_PF=0
b t1 DBSLMM_B200_TILE64=1 DBSLMM_B200_L2_PF=0
b t2 DBSLMM_B200_TILE64=2 DBSLMM_B200_L2_PF=0
b t2b DBSLMM_B200_TILE64=2 DBSLMM_B200_L2_PF=0
b t2pf2 DBSLMM_B200_TILE64=2 DBSLMM_B200_L2_PF=2
