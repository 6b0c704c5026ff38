#!/bin/bash
# round 2, two B200s: the rank-sharded workloads with their final NCCL gather (C3 long-video sampling, C5 ELBO sweep) and
# the default bench at N = 2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR bench.py --gpus 2 --workload c3 > gpurun_out/bench_r3_final_c3_2gpu.json 2> gpurun_out/bench_r3_final_c3_2gpu.err; tail -c 700 gpurun_out/bench_r3_final_c3_2gpu.json; tail -2 gpurun_out/bench_r3_final_c3_2gpu.err
timeout 600 $TR bench.py --gpus 2 --workload c5 > gpurun_out/bench_r3_final_c5_2gpu.json 2> gpurun_out/bench_r3_final_c5_2gpu.err; tail -c 600 gpurun_out/bench_r3_final_c5_2gpu.json
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r3_final_2gpu.json 2> gpurun_out/bench_r3_final_2gpu.err; tail -c 300 gpurun_out/bench_r3_final_2gpu.json
NCCL_DEBUG=INFO timeout 300 $TR bench.py --gpus 2 --workload c5 --c5-timesteps 20 2>&1 | grep -E "NCCL INFO (Channel|Connected|comm|ncclCommInit|Using network|NVLS)" | head -12 > gpurun_out/nccl_info_r3_2gpu.log; tail -5 gpurun_out/nccl_info_r3_2gpu.log
