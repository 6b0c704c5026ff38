#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-stock-gpu-baseline"
run() { name=$1; shift; env "$@" $B > gpurun_out/bench_r4f_$name.json 2> gpurun_out/bench_r4f_$name.err; python - gpurun_out/bench_r4f_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c=d['clocks']; print(f"{sys.argv[2]:24s} ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  frac {d['roofline']['frac']:.4f}  clk {c['sm_mhz']} power {c.get('power_w')}")
except Exception as e: print(sys.argv[2], 'failed', e)
PY
}
run prio X=1
run noprio VDM_SIDE_PRIORITY=0
run prio_again X=1
run noprio_again VDM_SIDE_PRIORITY=0
timeout 300 python profiles/step_timeline.py > gpurun_out/step_timeline_r4f_prio.json 2> gpurun_out/step_timeline_r4f_prio.err
VDM_SIDE_PRIORITY=0 timeout 300 python profiles/step_timeline.py > gpurun_out/step_timeline_r4f_noprio.json 2> gpurun_out/step_timeline_r4f_noprio.err
python - <<'PY'
import json
for n in ('prio','noprio'):
    try:
        d=json.load(open(f'gpurun_out/step_timeline_r4f_{n}.json'))
        ev=d.get('launches') or []
        side=[e for e in ev if e['s']==1][:12]
        print(n,'span',d['span_ms'],'side:',[(e['k'][:14],round(e['t0'],3),round(e['t1'],3)) for e in side])
    except Exception as e: print(n,'failed',e)
PY
