#!/bin/bash
# round 2 final evidence on one B200: full GPU test suite, smoke, default bench (+ breakdown), sustained bench, reference
# arm, secondary workloads (c5, c4, short c3), the ncu launch list of the bench command, an ncu metrics pass over the
# GEMM probe and --set full captures of the top conv launch and of the tcgen05 attention kernel (kept small: the merge
# back is limited to 64 MiB)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/final_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_tests.log
tail -3 gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -3 gpurun_out/final_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 --profile-json gpurun_out/kb_final.json > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 400 gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-stock-gpu-baseline > gpurun_out/bench_final_sustained.json 2> gpurun_out/bench_final_sustained.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err; tail -c 600 gpurun_out/bench_reference_arm.json
timeout 600 python bench.py --workload c5 > gpurun_out/bench_c5_1gpu.json 2> gpurun_out/bench_c5_1gpu.err; tail -c 500 gpurun_out/bench_c5_1gpu.json
timeout 600 python bench.py --workload c3 --c3-frames 100 > gpurun_out/bench_c3_1gpu_T100.json 2> gpurun_out/bench_c3_1gpu.err; tail -c 500 gpurun_out/bench_c3_1gpu_T100.json
timeout 600 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline --no-stock-gpu-baseline --profile-json gpurun_out/kb_final_c4.json > gpurun_out/bench_final_c4.json 2> gpurun_out/bench_final_c4.err; tail -c 300 gpurun_out/bench_final_c4.json
python profiles/energy_probe.py > gpurun_out/energy_probe_final.log 2>&1; tail -3 gpurun_out/energy_probe_final.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-stock-gpu-baseline"
$CMD > gpurun_out/final_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r2.csv $CMD > gpurun_out/final_ncu1.log 2>&1
python profiles/gemm_ncu_probe.py > gpurun_out/final_probe_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum --clock-control none -k regex:'gemm_tc' --csv --log-file gpurun_out/ncu_gemm_probe_metrics_r2.csv python profiles/gemm_ncu_probe.py > gpurun_out/final_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_halo_t' -c 2 -o gpurun_out/ncu_gemm_r2 -f python profiles/gemm_ncu_probe.py > gpurun_out/final_ncu3.log 2>&1
python profiles/ncu_attn_sampler_probe.py > gpurun_out/final_attn_probe_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'attn_spatial_sm100|attn_temporal_mma|sampler_step|gn_apply' -o gpurun_out/ncu_attn_sampler_r2 -f python profiles/ncu_attn_sampler_probe.py --once > gpurun_out/final_ncu4.log 2>&1
tail -2 gpurun_out/final_ncu1.log gpurun_out/final_ncu3.log gpurun_out/final_ncu4.log
du -sm gpurun_out; ls -la gpurun_out | sort -k5 -n | tail -4
# keep the merge under the 64 MiB limit: drop the largest report first if needed
while [ "$(du -sm gpurun_out | cut -f1)" -gt 58 ]; do f=$(ls -S gpurun_out | head -1); echo "dropping $f"; rm -f "gpurun_out/$f"; done
