#!/bin/bash
# round 2, call b: fused GroupNorm kernels -- kernel tests first (bounded), then model tests, then bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "fused_groupnorm or halo_kernel or large_mean" > gpurun_out/r2b_k.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_k.log
tail -15 gpurun_out/r2b_k.log
if grep -q "rc=0" gpurun_out/r2b_k.log; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_tests.log
  tail -8 gpurun_out/r2b_tests.log
  timeout 600 python bench.py --steps 20 --warmup 5 --profile-json gpurun_out/kb_r2b.json > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err
  tail -c 1500 gpurun_out/bench_r2b.json; tail -5 gpurun_out/bench_r2b.err
fi
