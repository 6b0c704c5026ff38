#!/usr/bin/env python
"""GPU time of the step's small HBM / latency-bound launches, 20 of each captured into one CUDA graph (no host launch
cost): GroupNorm-apply at the 16x16 / 8x8 levels, temporal GroupNorm, the attention cores, against their byte counts."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_b200 import ops  # noqa: E402

dev = 'cuda'
B, T, heads = 8, 20, 4
n = B * T


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


def report(name, t, nbytes):
    print(f'{name:58s} {t:7.2f} us  {nbytes / t / 1e3:7.0f} GB/s', flush=True)


for HWs, C in ((64, 128), (32, 256), (16, 384), (8, 512)):
    M = n * HWs * HWs
    for in_dt, copy in ((torch.bfloat16, False), (torch.float16, False), (torch.float16, True)):
        x = torch.randn(M, C, device=dev).to(in_dt)
        out = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
        cp = torch.empty(M, C, device=dev, dtype=torch.float16) if copy else None
        st = torch.zeros(n, 2, C, device=dev, dtype=torch.int64)
        st[:, 1] = HWs * HWs * 2 ** 24
        g_, b_ = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        t = bench(lambda: ops.gn_apply(x, None, n, HWs, HWs, out, stats1=st, gamma=g_, beta=b_, silu=not copy, copy=cp))
        report(f'gn_apply {HWs}x{HWs} C={C} {str(in_dt)[6:]} copy={int(copy)}', t, M * C * (4 + (2 if copy else 0)))
for HW, C in ((256, 384), (64, 512)):
    M = n * HW
    x = torch.randn(M, C, device=dev).half()
    xn = torch.empty_like(x)
    xa = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    g_, b_ = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    t = bench(lambda: ops.gn_temporal(x, B, T, HW, C, g_, b_, xn, xa))
    report(f'gn_temporal HW={HW} C={C}', t, M * C * 6)
    hd = C // heads
    qkv = torch.randn(M, 3 * C, device=dev).bfloat16()
    att = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    t = bench(lambda: ops.attn_spatial(qkv, n, HW, heads, hd, att))
    report(f'attn_spatial L={HW} C={C}', t, M * C * 8)
    gpt = 1 if HW >= 128 else 128 // HW
    SW = 128 * gpt
    sk, sq = torch.randn(M, SW, device=dev), torch.randn(M, SW, device=dev)
    mask = torch.ones(B, T, device=dev)
    pm, pv = torch.zeros(M, SW, device=dev, dtype=torch.bfloat16), torch.empty(M, C, device=dev)
    t = bench(lambda: ops.attn_temporal_tc(qkv, sk, sq, mask, True, B, T, HW, heads, hd, gpt, pm, pv))
    report(f'attn_temporal_tc HW={HW} C={C}', t, M * 3 * C * 2 + 2 * M * SW * 4 + M * SW * 2 + M * C * 4)
