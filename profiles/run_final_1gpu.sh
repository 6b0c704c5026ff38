#!/bin/bash
# round 2 final evidence on one B200: full GPU test suite, smoke, default bench (+ breakdown), reference arm, secondary
# workloads (c5, c4, short c3), then the ncu launch list of the bench command and one --set full capture of the GEMM probe
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/final_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_tests.log
tail -3 gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -3 gpurun_out/final_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 --profile-json gpurun_out/kb_final.json > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 400 gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err; tail -c 600 gpurun_out/bench_reference_arm.json
timeout 600 python bench.py --workload c5 > gpurun_out/bench_c5_1gpu.json 2> gpurun_out/bench_c5_1gpu.err; tail -c 500 gpurun_out/bench_c5_1gpu.json
timeout 600 python bench.py --workload c3 --c3-frames 100 > gpurun_out/bench_c3_1gpu_T100.json 2> gpurun_out/bench_c3_1gpu.err; tail -c 500 gpurun_out/bench_c3_1gpu_T100.json
timeout 600 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline --no-stock-gpu-baseline --profile-json gpurun_out/kb_final_c4.json > gpurun_out/bench_final_c4.json 2> gpurun_out/bench_final_c4.err; tail -c 300 gpurun_out/bench_final_c4.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-stock-gpu-baseline"
$CMD > gpurun_out/final_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r2.csv $CMD > gpurun_out/final_ncu1.log 2>&1
python profiles/gemm_ncu_probe.py > gpurun_out/final_probe_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gemm_tc' -o gpurun_out/ncu_gemm_r2 -f python profiles/gemm_ncu_probe.py > gpurun_out/final_ncu2.log 2>&1
tail -2 gpurun_out/final_ncu1.log gpurun_out/final_ncu2.log
