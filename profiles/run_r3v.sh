#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pipeline_gpu.py -m gpu -q -x > gpurun_out/r3v_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r3v_tests.log
tail -5 gpurun_out/r3v_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-gpu-baseline > gpurun_out/bench_r3v.json 2> gpurun_out/bench_r3v.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r3v.json').read().strip().splitlines()[-1]); c=d['clocks']
print(f"bench: ms/step {d['ms_per_step']:.3f} e2e {d['e2e']['ms_per_step']:.3f} frac {d['roofline']['frac']:.4f} launches {d['gpu_launches']}")
PY
