#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -x -k "micro_batches or session3" > gpurun_out/r3j_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r3j_tests.log
tail -15 gpurun_out/r3j_tests.log
