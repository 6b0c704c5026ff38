#!/bin/bash
# round 2 session 4, two B200s: the default bench at N = 2 with the driver's launch line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r4_final_2gpu.json 2> gpurun_out/bench_r4_final_2gpu.err; tail -c 200 gpurun_out/bench_r4_final_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r4_final_2gpu.json').read().strip().splitlines()[-1]); c=d['clocks']
print(f"2 GPUs: ms/step {d['ms_per_step']:.3f} value {d['value']:.0f} e2e {d['e2e']['ms_per_step']:.3f} frac {d['roofline']['frac']:.4f} n_gpus {d['n_gpus']}")
PY
