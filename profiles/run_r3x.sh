#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
VDM_LIB=$PWD/profiles/_diag/libvdm_trace.so TRACE_SHAPES=conv64_128_128_bf16,conv64_256_128,conv32_256_256,conv8_512_512 timeout 300 python profiles/gemm_trace.py > gpurun_out/gemm_trace_r3x.log 2>&1; grep -v "1cta\|2cta" gpurun_out/gemm_trace_r3x.log | tail -12
