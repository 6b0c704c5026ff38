#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "output_head or rowbias" > gpurun_out/r4a_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r4a_tests.log
tail -4 gpurun_out/r4a_tests.log
timeout 300 python profiles/head_conv_time.py > gpurun_out/head_conv_time_r4a.log 2>&1; tail -6 gpurun_out/head_conv_time_r4a.log
python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-stock-gpu-baseline > gpurun_out/bench_r4a.json 2> gpurun_out/bench_r4a.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r4a.json').read().strip().splitlines()[-1]); c=d['clocks']
print(f"ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  frac {d['roofline']['frac']:.4f}  clk {c['sm_mhz']} power {c.get('power_w')}")
PY
