#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r4a_$name.json 2> gpurun_out/r4a_$name.err; echo -n "$name: "; python tools/bench_brief.py gpurun_out/r4a_$name.json | sed 's/.*e2e_ms=\([0-9.]*\).*/e2e \1/'; }
run nt1 DBSLMM_B200_NT=1
run nt0 DBSLMM_B200_NT=0
run nt1 DBSLMM_B200_NT=1
run nt0 DBSLMM_B200_NT=0
run nt1f12 DBSLMM_B200_NT=1 DBSLMM_B200_FILL_THREADS=12
run nt1f4 DBSLMM_B200_NT=1 DBSLMM_B200_FILL_THREADS=4
run nt1h0 DBSLMM_B200_NT=1 DBSLMM_B200_HOST_THREADS=6
for nt in 1 0; do
DBSLMM_B200_NT=$nt timeout 300 python tools/stream_trace.py 2> gpurun_out/r4a_trace_nt$nt.err; awk "/streaming fit 1/{f=1;next} /streaming fit 2/{f=0} f" gpurun_out/r4a_trace_nt$nt.err | grep "plan\|maps\|launched\|device done\|chain 0"
done
