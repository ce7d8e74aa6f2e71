#!/bin/bash
# round-2 session P: mirrored warp slots (SMSP balance), vectorised diagonal-tile load / write-back
mkdir -p gpurun_out
build/diag_probe > gpurun_out/r2p_diag_probe.txt 2>&1; grep "rep 2\|max abs" gpurun_out/r2p_diag_probe.txt | cut -c1-330
timeout 900 python -m pytest tests/test_gpu_modes.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2p_pytest.log
b() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2p_$tag.json 2> gpurun_out/r2p_$tag.err; python tools/bench_brief.py gpurun_out/r2p_$tag.json; }
b d1 X=1
b d2 X=1
b tpc8 DBSLMM_B200_TPC=8,2
b tpc2 DBSLMM_B200_TPC=2,2
b cls DBSLMM_B200_CLASSES=6,10,16,32
