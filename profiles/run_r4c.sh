#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "output_head or rowbias" > gpurun_out/r4c_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r4c_tests.log
tail -3 gpurun_out/r4c_tests.log
VDM_HEAD_PREFETCH=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "output_head" > gpurun_out/r4c_tests_pf.log 2>&1; echo "rc=$?" >> gpurun_out/r4c_tests_pf.log
tail -2 gpurun_out/r4c_tests_pf.log
for v in "X=1" "VDM_HEAD_PREFETCH=1" "VDM_HEAD_CARVEOUT=1" "VDM_HEAD_PREFETCH=1 VDM_HEAD_CARVEOUT=1" "X=1"; do
  echo "== $v"; env $v timeout 300 python profiles/head_conv_time.py 2>&1 | tail -4
done > gpurun_out/head_conv_time_r4c.log 2>&1
cat gpurun_out/head_conv_time_r4c.log
