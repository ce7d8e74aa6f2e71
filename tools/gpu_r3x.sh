#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r3x_$name.json 2> gpurun_out/r3x_$name.err; echo -n "$name: "; python tools/bench_brief.py gpurun_out/r3x_$name.json | sed 's/.*e2e_ms=\([0-9.]*\).*/e2e \1/'; }
for f in 4 6 8 12 4 6 8 12; do run f$f DBSLMM_B200_FILL_THREADS=$f; done
DBSLMM_B200_FILL_THREADS=6 timeout 300 python tools/stream_trace.py 2> gpurun_out/r3x_trace.err; awk "/streaming fit 2/{f=1;next} /load_bed/{f=0} f" gpurun_out/r3x_trace.err | head -9
