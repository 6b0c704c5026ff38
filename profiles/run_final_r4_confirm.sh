#!/bin/bash
# last build of the round: full GPU suite, smoke, default bench (driver's command line)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/confirm4_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/confirm4_tests.log
tail -3 gpurun_out/confirm4_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/confirm4_smoke.log 2>&1; tail -3 gpurun_out/confirm4_smoke.log
timeout 900 python bench.py > gpurun_out/confirm4_bench.json 2> gpurun_out/confirm4_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/confirm4_bench.json').read().strip().splitlines()[-1]); c=d['clocks']
print(f"default bench: ms/step {d['ms_per_step']:.3f} value {d['value']:.0f} e2e {d['e2e']['ms_per_step']:.3f} ({d['e2e']['value']:.0f}) frac {d['roofline']['frac']:.4f} clk {c['sm_mhz']} reasons {c['reasons']} launches {d['gpu_launches']} steps {d['steps']} warmup {d['warmup']}")
PY
