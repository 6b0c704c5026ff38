#!/bin/bash
# round 2, call c: fused GroupNorm (out_layers) + programmatic dependent launch
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_tests.log
tail -8 gpurun_out/r2c_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --profile-json gpurun_out/kb_r2c.json > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err
tail -c 2500 gpurun_out/bench_r2c.json; tail -5 gpurun_out/bench_r2c.err
VDM_PDL=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-gpu-baseline > gpurun_out/bench_r2c_nopdl.json 2> gpurun_out/bench_r2c_nopdl.err
python - <<'PY'
import json
for f in ('bench_r2c', 'bench_r2c_nopdl'):
    try:
        d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'], 3), 'e2e ms', round(d['e2e']['ms_per_step'], 3), 'frac', round(d['roofline']['frac'], 4))
    except Exception as e:
        print(f, 'failed', e)
PY
