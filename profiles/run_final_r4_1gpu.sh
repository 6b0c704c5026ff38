#!/bin/bash
# round 2 session 4, final evidence on one B200: full GPU test suite, smoke, default bench (+ breakdown), reference arm,
# the ncu launch list of the bench command
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final4_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final4_tests.log
tail -3 gpurun_out/final4_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final4_smoke.log 2>&1; tail -3 gpurun_out/final4_smoke.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final4_reference_arm.json 2> gpurun_out/bench_final4_reference_arm.err; tail -c 300 gpurun_out/bench_final4_reference_arm.json
timeout 900 python bench.py --profile-json gpurun_out/kb_final4.json > gpurun_out/bench_final4.json 2> gpurun_out/bench_final4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final4.json').read().strip().splitlines()[-1]); c=d['clocks']
print(f"default bench: ms/step {d['ms_per_step']:.3f} value {d['value']:.0f} e2e {d['e2e']['ms_per_step']:.3f} frac {d['roofline']['frac']:.4f} clk {c['sm_mhz']} reasons {c['reasons']} launches {d['gpu_launches']} steps {d['steps']} warmup {d['warmup']}")
PY
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-stock-gpu-baseline"
$CMD > gpurun_out/final4_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r4.csv $CMD > gpurun_out/final4_ncu1.log 2>&1
tail -2 gpurun_out/final4_ncu1.log
du -sm gpurun_out
