#!/bin/bash
# round 2, call d: micro-batches in parallel streams; PDL / fused-norm switches measured in isolation
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_tests.log
tail -8 gpurun_out/r2d_tests.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-gpu-baseline"
run() { name=$1; shift; env "$@" timeout 600 $B > gpurun_out/bench_r2d_$name.json 2> gpurun_out/bench_r2d_$name.err; }
run mb2 VDM_MICRO_BATCHES=2
run mb1 VDM_MICRO_BATCHES=1
run mb2_pdl VDM_MICRO_BATCHES=2 VDM_PDL=1
run mb1_pdl VDM_MICRO_BATCHES=1 VDM_PDL=1
run mb4 VDM_MICRO_BATCHES=4
run mb2_again VDM_MICRO_BATCHES=2
python - <<'PY'
import json
for f in ('mb2', 'mb1', 'mb2_pdl', 'mb1_pdl', 'mb4', 'mb2_again'):
    try:
        d = json.loads(open(f'gpurun_out/bench_r2d_{f}.json').read().strip().splitlines()[-1])
        print(f'{f:10s} ms/step {d["ms_per_step"]:.3f}  e2e ms {d["e2e"]["ms_per_step"]:.3f}  frac {d["roofline"]["frac"]:.4f}  clocks {d["clocks"]}')
    except Exception as e:
        print(f, 'failed', e, open(f'gpurun_out/bench_r2d_{f}.err').read()[-400:])
PY
