#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r3_final_8gpu.json 2> gpurun_out/bench_r3_final_8gpu.err; tail -c 400 gpurun_out/bench_r3_final_8gpu.json
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r3_final_8gpu.json') if l.startswith('{')][-1])
print('8 gpus: ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['ms_per_step'])
PY
