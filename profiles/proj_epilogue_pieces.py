#!/usr/bin/env python
"""What does the proj_out epilogue pay for?  The attention output projection (short K, residual + fp16 stream output +
GroupNorm statistics) with pieces of its epilogue switched off, whole and half batch, CUDA-graph timing
(gemm_fixed_cost.bench); plus a standalone gn_stats pass over the same output for comparison."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_b200 import ops  # noqa: E402

dev = 'cuda'


def bench(fn, reps=20):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (3 * reps)


for name, M, N, K, HW in (('proj16', 40960, 384, 384, 256), ('proj16 half', 20480, 384, 384, 256),
                          ('proj8', 10240, 512, 512, 64), ('proj8 half', 5120, 512, 512, 64)):
    n, H = M // HW, int(HW ** 0.5)
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.zeros(N, device=dev)
    res = torch.randn(M, N, device=dev).half()
    out16 = torch.empty(M, N, device=dev, dtype=torch.float16)
    outb = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    st = torch.zeros(n, 2, N, device=dev, dtype=torch.int64)
    st64 = torch.zeros(n, 2, N, device=dev, dtype=torch.float64)
    geo = dict(n_img=n, H=H, W=H, taps=1, bias=bias)
    rows = [
        ('residual + fp16 out + stats (model)', lambda: ops.gemm(a, w, N, residual=res, out_f32=out16, stats_out=st, **geo)),
        ('residual + fp16 out', lambda: ops.gemm(a, w, N, residual=res, out_f32=out16, **geo)),
        ('fp16 out + stats', lambda: ops.gemm(a, w, N, out_f32=out16, stats_out=st, **geo)),
        ('fp16 out only', lambda: ops.gemm(a, w, N, out_f32=out16, **geo)),
        ('bf16 out only (TMA-store epilogue)', lambda: ops.gemm(a, w, N, out_bf16=outb, **geo)),
        ('standalone gn_stats of the output', lambda: ops.gn_stats(out16, n, HW, st64)),
    ]
    print(f'{name}: M={M} N={N} K={K}')
    for label, fn in rows:
        print(f'  {label:40s} {bench(fn):7.2f} us', flush=True)
