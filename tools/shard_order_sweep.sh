#!/bin/bash
# Tuning aid (round 2, last session): one rank's shard of an N-GPU run timed on one GPU (bench.py --emulate-shard):
# upload order of the streaming fit (DBSLMM_B200_UPLOAD_BULK_FIRST) and the CTA slot the bulk batches leave to the
# chain-bound classes (DBSLMM_B200_CHAIN_SLOT).  Writes one bench line per case into gpurun_out/ and prints step / e2e ms.
mkdir -p gpurun_out
TAG=${TAG:-r05u}
run() {   # name, bench args (quoted), env...
    local name=$1 bargs=$2; shift 2
    env "$@" timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $bargs \
        > "gpurun_out/${TAG}_${name}.json" 2> "gpurun_out/${TAG}_${name}.err"
    python - "$TAG" "$name" <<'P'
import json, sys
t, n = sys.argv[1:3]
try:
    d = json.loads(open(f"gpurun_out/{t}_{n}.json").read().strip().splitlines()[-1])
    par = d.get("parity") or {}
    print(f"{n:24s} step {d['ms_per_step']:.3f} ms  e2e {d['e2e']['ms_per_step']:.3f} ms  chol {d['roofline']['ms_per_step']:.3f} ms  class_ms {[round(x, 2) for x in d['rooflines_other']['chol_class_ms']]}  exact {par.get('max_rel_vs_exact_oracle')} stream==resident {par.get('streaming_vs_resident_max_rel')}")
except Exception as e:
    print(n, "FAILED", e)
P
}
run 0of8_default   "--emulate-shard 0/8"              DBSLMM_B200_X=0
run 0of8_noslot    "--emulate-shard 0/8 --no-parity"  DBSLMM_B200_CHAIN_SLOT=0
run 0of8_old       "--emulate-shard 0/8 --no-parity"  DBSLMM_B200_CHAIN_SLOT=0 DBSLMM_B200_UPLOAD_BULK_FIRST=1
run 0of4_default   "--emulate-shard 0/4 --no-parity"  DBSLMM_B200_X=0
run 0of4_noslot    "--emulate-shard 0/4 --no-parity"  DBSLMM_B200_CHAIN_SLOT=0
run c3_default     "--no-parity"                      DBSLMM_B200_X=0
