#!/bin/bash
# session 3g: full GPU suite + C3 / missing-calls bench at the state of the Gram rework
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r3g_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r3g_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r3g_c3.json 2> gpurun_out/r3g_c3.err; echo "c3 rc=$?"; python tools/bench_brief.py gpurun_out/r3g_c3.json
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-parity --missing 0.005 > gpurun_out/r3g_missing.json 2> gpurun_out/r3g_missing.err; echo "missing rc=$?"; python tools/bench_brief.py gpurun_out/r3g_missing.json
