#!/bin/bash
# round-2 session G (2 GPUs): the N > 1 paths -- bench.py under torchrun (parity at every rank, gathered checksum), CLI --gpus 2
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2g_bench2.json 2> gpurun_out/r2g_bench2.err; echo "bench2 rc=$?"
tail -c 1500 gpurun_out/r2g_bench2.json; echo; tail -5 gpurun_out/r2g_bench2.err
timeout 600 python -m pytest tests/test_cli.py tests/test_gpu_score.py -q -m gpu > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2g_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/r2g_ref2.json 2> gpurun_out/r2g_ref2.err; echo "ref2 rc=$?"; cut -c1-300 gpurun_out/r2g_ref2.json
timeout 600 python bench.py --steps 5 --warmup 3 --config c4 --no-cpu-baseline --no-parity > gpurun_out/r2g_c4.json 2> gpurun_out/r2g_c4.err; echo "c4 rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2g_c4.json").read().strip().splitlines()[-1]); print("c4 ms/step",d["ms_per_step"],"e2e",d["e2e"]["ms_per_step"],"prs",d["rooflines_other"]["prs"])
PY
