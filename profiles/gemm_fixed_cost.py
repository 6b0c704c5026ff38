#!/usr/bin/env python
"""Fixed cost of one tcgen05 GEMM launch: back-to-back launches of shapes with ONE tile per CTA and growing K (the slope
is the main loop, the intercept is prologue + first loads + epilogue + drain), plus the attention linears of the step."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_b200 import ops  # noqa: E402

dev = 'cuda'


def bench(fn, reps=20):
    """GPU time per launch: `reps` launches captured into one CUDA graph (no host launch cost between them)."""
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


def linear(M, N, K, kind):
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.zeros(N, device=dev)
    if kind == 'bf16':
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        return lambda: ops.gemm(a, w, N, n_img=M, H=1, W=1, taps=1, bias=bias, out_bf16=out)
    HW = 256 if M % 256 == 0 else 64
    n = M // HW
    res = torch.randn(M, N, device=dev).half()
    out = torch.empty(M, N, device=dev, dtype=torch.float16)
    st = torch.zeros(n, 2, N, device=dev, dtype=torch.int64)
    H = int(HW ** 0.5)
    return lambda: ops.gemm(a, w, N, n_img=n, H=H, W=H, taps=1, bias=bias, residual=res, out_f32=out, stats_out=st)


print('one 128x128 tile per CTA (148 tiles), bf16 out, bias only')
for K in (64, 128, 256, 512, 1024, 2048, 4096):
    t = bench(linear(148 * 128, 128, K, 'bf16'))
    print(f'  M=18944 N=128 K={K:5d}: {t:7.2f} us', flush=True)
print('a single tile')
for K in (64, 512, 4096):
    t = bench(linear(128, 128, K, 'bf16'))
    print(f'  M=128 N=128 K={K:5d}: {t:7.2f} us', flush=True)
print('one tile per CTA, residual + fp16 out + statistics')
for K in (64, 512, 2048):
    t = bench(linear(148 * 128 // 256 * 256, 128, K, 'res'))
    print(f'  M={148 * 128 // 256 * 256} N=128 K={K:5d}: {t:7.2f} us', flush=True)
print('attention linears of the step')
for name, M, N, K, kind in [('qkv16', 40960, 1152, 384, 'bf16'), ('proj16', 40960, 384, 384, 'res'),
                            ('qkv8', 10240, 1536, 512, 'bf16'), ('proj8', 10240, 512, 512, 'res'),
                            ('qkv16 half batch', 20480, 1152, 384, 'bf16'), ('proj16 half batch', 20480, 384, 384, 'res'),
                            ('qkv8 half batch', 5120, 1536, 512, 'bf16'), ('proj8 half batch', 5120, 512, 512, 'res')]:
    t = bench(linear(M, N, K, kind))
    print(f'  {name:20s} M={M} N={N} K={K}: {t:7.2f} us  {2.0 * M * N * K / t / 1e6:7.0f} TFLOP/s', flush=True)
