timeout 600 python -m pytest tests/test_gpu_modes.py tests/test_gpu_parity.py tests/test_gpu_streaming.py -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2p_m$i.json 2> gpurun_out/r2p_m$i.err; python tools/bench_brief.py gpurun_out/r2p_m$i.json; done
python -c "
import json;d=json.loads(open('gpurun_out/r2p_m1.json').read().strip().splitlines()[-1]);print(d['parity'])"
