#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export VDM_LIB=$PWD/video_diffusion_b200/libvdm_trace.so
TRACE_SHAPES=qkv16,proj16,qkv8,proj8 TRACE_EXPERIMENTS=1 timeout 600 python profiles/gemm_trace.py > gpurun_out/trace_linears_r2.log 2>&1
cat gpurun_out/trace_linears_r2.log
VDM_GEMM_TS=0 TRACE_SHAPES=qkv16,qkv8 timeout 600 python profiles/gemm_trace.py > gpurun_out/trace_linears_r2_nots.log 2>&1
cat gpurun_out/trace_linears_r2_nots.log
