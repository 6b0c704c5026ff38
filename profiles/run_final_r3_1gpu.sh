#!/bin/bash
# round 2 session 3, final evidence on one B200: full GPU test suite, smoke, default bench (+ breakdown), sustained bench,
# reference arm, secondary workloads (c5, c4, short c3, c2chain), the ncu launch list of the bench command and
# --set full captures of the fused temporal kernel / attention / sampler / gn_apply probe
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/final3_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final3_tests.log
tail -3 gpurun_out/final3_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final3_smoke.log 2>&1; tail -3 gpurun_out/final3_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 --profile-json gpurun_out/kb_final3.json > gpurun_out/bench_final3.json 2> gpurun_out/bench_final3.err
tail -c 300 gpurun_out/bench_final3.json; tail -3 gpurun_out/bench_final3.err
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-stock-gpu-baseline > gpurun_out/bench_final3_sustained.json 2> gpurun_out/bench_final3_sustained.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final3_reference_arm.json 2> gpurun_out/bench_final3_reference_arm.err; tail -c 400 gpurun_out/bench_final3_reference_arm.json
timeout 600 python bench.py --workload c5 > gpurun_out/bench_final3_c5.json 2> gpurun_out/bench_final3_c5.err; tail -c 300 gpurun_out/bench_final3_c5.json
timeout 600 python bench.py --workload c3 --c3-frames 100 > gpurun_out/bench_final3_c3_T100.json 2> gpurun_out/bench_final3_c3.err; tail -c 300 gpurun_out/bench_final3_c3_T100.json
timeout 600 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline --no-stock-gpu-baseline --profile-json gpurun_out/kb_final3_c4.json > gpurun_out/bench_final3_c4.json 2> gpurun_out/bench_final3_c4.err; tail -c 300 gpurun_out/bench_final3_c4.json
timeout 600 python bench.py --workload c2chain > gpurun_out/bench_final3_c2chain.json 2> gpurun_out/bench_final3_c2chain.err; tail -c 300 gpurun_out/bench_final3_c2chain.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-stock-gpu-baseline"
$CMD > gpurun_out/final3_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r3.csv $CMD > gpurun_out/final3_ncu1.log 2>&1
python profiles/ncu_attn_sampler_probe.py > gpurun_out/final3_attn_probe_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'attn_temporal_fused|attn_spatial_sm100|sampler_step|gn_apply' -o gpurun_out/ncu_attn_sampler_r3 -f python profiles/ncu_attn_sampler_probe.py --once > gpurun_out/final3_ncu4.log 2>&1
cat gpurun_out/final3_attn_probe_plain.log
tail -2 gpurun_out/final3_ncu1.log gpurun_out/final3_ncu4.log
du -sm gpurun_out
while [ "$(du -sm gpurun_out | cut -f1)" -gt 58 ]; do f=$(ls -S gpurun_out | head -1); echo "dropping $f"; rm -f "gpurun_out/$f"; done
