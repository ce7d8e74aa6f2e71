#!/bin/bash
mkdir -p gpurun_out
for wc in 0; do
DBSLMM_B200_BLOB_WC=$wc timeout 300 python tools/stream_trace.py 2> gpurun_out/r3t_trace_wc$wc.err; awk "/streaming fit 1/{f=1;next} /streaming fit 2/{f=0} f" gpurun_out/r3t_trace_wc$wc.err | grep "plan blob\|plan built\|chain 0\|device done\|per-SNP"
done
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r3t_$name.json 2> gpurun_out/r3t_$name.err; echo -n "$name: "; python tools/bench_brief.py gpurun_out/r3t_$name.json | sed 's/.*e2e_ms=\([0-9.]*\).*/e2e \1/'; }
run wc0 DBSLMM_B200_BLOB_WC=0

run wc0 DBSLMM_B200_BLOB_WC=0

