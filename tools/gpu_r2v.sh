#!/bin/bash
# round-2 session V: CTA-pair Gram kernel -- smoke, parity tests, C3 bench against the 128-tile kernel
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2v_smoke.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_streaming.py -x -q -m gpu > gpurun_out/r2v_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2v_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2v_c3_pair.json 2> gpurun_out/r2v_c3_pair.err; echo "c3 pair rc=$?"; python tools/bench_brief.py gpurun_out/r2v_c3_pair.json
DBSLMM_B200_GRAM=single timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2v_c3_single.json 2> gpurun_out/r2v_c3_single.err; echo "c3 single rc=$?"; python tools/bench_brief.py gpurun_out/r2v_c3_single.json
