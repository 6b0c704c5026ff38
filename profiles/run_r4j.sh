#!/bin/bash
# host stepper: separate upload / download streams vs one copy stream, same box
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-stock-gpu-baseline > gpurun_out/bench_r4j_$name.json 2> gpurun_out/bench_r4j_$name.err
python - gpurun_out/bench_r4j_$name.json $name <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c=d['clocks']
print(f"{sys.argv[2]:12s} ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f} (serial {d['e2e']['ms_per_step_serial']:.3f})  e2e/value {d['e2e']['ms_per_step']/d['ms_per_step']:.4f}  clk {c['sm_mhz']}")
PY
}
run two X=1
run one VDM_STEPPER_ONE_STREAM=1
run two_again X=1
run one_again VDM_STEPPER_ONE_STREAM=1
