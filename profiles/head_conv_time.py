#!/usr/bin/env python
"""Time of the output-head conv (conv_small_n.cu) at the C2 / C4 shapes, raw fp16 stream with the fused
GroupNorm-apply + SiLU and the pre-activated bf16 operand; inputs (168 MB at C2) exceed L2 per launch pair."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_b200 import ops  # noqa: E402

dev = 'cuda'


def bench(fn, reps=10):
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


for n, H, W, C, N in ((160, 64, 64, 128, 3), (80, 64, 64, 128, 3), (32, 128, 128, 128, 3), (160, 64, 64, 128, 6)):
    M = n * H * W
    xs = [torch.randn(M, C, device=dev).half() for _ in range(2)]          # alternate two inputs: 2 x 168 MB > L2
    xb = [t.bfloat16() for t in xs]
    w = (torch.randn(N, 9 * C, device=dev) * 0.03).bfloat16()
    b = torch.zeros(N, device=dev)
    coef = torch.rand(n, C, 2, device=dev) + 0.5
    out = torch.empty(n, N, H, W, device=dev)
    k = [0]

    def fused():
        k[0] ^= 1
        ops.gemm(xs[k[0]], w, N, n_img=n, H=H, W=W, taps=9, bias=b, out_f32=out, out_nchw=True, a1_coef=coef, a1_act=True)

    def plain():
        k[0] ^= 1
        ops.gemm(xb[k[0]], w, N, n_img=n, H=H, W=W, taps=9, bias=b, out_f32=out, out_nchw=True)

    tf, tp = bench(fused), bench(plain)
    gb = M * C * 2 / 1e9
    print(f'n={n} {H}x{W} C={C} N={N}: fused-norm {tf:7.1f} us ({gb / tf * 1e6:6.0f} GB/s)   bf16 operand {tp:7.1f} us '
          f'({gb / tp * 1e6:6.0f} GB/s)', flush=True)
