#!/bin/bash
# host stepper with separate upload / download streams: test + bench (e2e leg)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pipeline_gpu.py -m gpu -q -x -k "pipelined_host_stepper" > gpurun_out/r4i_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r4i_tests.log
tail -2 gpurun_out/r4i_tests.log
for n in a b; do
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-stock-gpu-baseline > gpurun_out/bench_r4i_$n.json 2> gpurun_out/bench_r4i_$n.err
python - gpurun_out/bench_r4i_$n.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c=d['clocks']
print(f"ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f} (serial {d['e2e']['ms_per_step_serial']:.3f})  frac {d['roofline']['frac']:.4f}  clk {c['sm_mhz']}")
PY
done
