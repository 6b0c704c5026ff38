#!/usr/bin/env python
"""Back-to-back timing of the small per-attention-block kernels at C2 shapes (queue stays full, so the numbers
are GPU time, not Python launch latency)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_b200 import ops  # noqa: E402

dev = 'cuda'
B, T, heads = 8, 20, 4


def timeit(name, fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f'{name:40s} {e0.elapsed_time(e1) / n * 1e3:8.1f} us')


for C, HW in ((384, 256), (512, 64)):
    hd = C // heads
    rows = B * T * T
    e_t = torch.randn(B * T, 3 * C, device=dev)
    fi = torch.arange(T, device=dev).repeat(B, 1)
    wd, bd = torch.randn(3, C, 3, device=dev), torch.randn(3, C, device=dev)
    hid = torch.empty(3, rows, C, device=dev, dtype=torch.bfloat16)
    timeit(f'rpe_hidden C={C}', lambda: ops.rpe_hidden(e_t, fi, wd, bd, B, T, C, hid))
    R = torch.randn(3 * rows, C, device=dev)
    w = torch.randn(3 * C, C, device=dev).bfloat16()
    timeit(f'R gemm grouped C={C}', lambda: ops.gemm(hid.view(3 * rows, C), w, C, n_img=3 * rows, H=1, W=1, taps=1,
                                                     out_f32=R, w_group_tiles=rows // 128))
    gpt = 1 if HW >= 128 else 128 // HW
    SW, ntg = 128 * gpt, (B * T + gpt - 1) // gpt
    bq = torch.empty(ntg * SW, C, device=dev, dtype=torch.bfloat16)
    bk, bv = torch.empty_like(bq), torch.empty(ntg * C, SW, device=dev, dtype=torch.bfloat16)
    Rs = [R[i * rows:(i + 1) * rows] for i in range(3)]
    timeit(f'rpe_expand C={C} HW={HW}', lambda: ops.rpe_expand(Rs[0], Rs[1], Rs[2], B, T, heads, hd, gpt, bq, bk, bv))
    M = B * T * HW
    x = torch.randn(M, C, device=dev)
    g, b_ = torch.randn(C, device=dev), torch.randn(C, device=dev)
    xn, xa = torch.empty(M, C, device=dev), torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    timeit(f'gn_temporal C={C} HW={HW}', lambda: ops.gn_temporal(x, B, T, HW, C, g, b_, xn, xa))
    qkv = torch.randn(M, 3 * C, device=dev).bfloat16()
    sk, sq = torch.randn(M, SW, device=dev), torch.randn(M, SW, device=dev)
    mask = torch.ones(B, T, device=dev)
    pm, pv = torch.zeros(M, SW, device=dev, dtype=torch.bfloat16), torch.empty(M, C, device=dev)
    timeit(f'attn_temporal_tc C={C} HW={HW}', lambda: ops.attn_temporal_tc(qkv, sk, sq, mask, True, B, T, HW, heads, hd,
                                                                           gpt, pm, pv))
    att = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    timeit(f'attn_spatial C={C} L={HW}', lambda: ops.attn_spatial(qkv, B * T, HW, heads, hd, att))
    st = torch.zeros(B * T, 2, C, device=dev, dtype=torch.int64)
    timeit(f'gn_apply bf16 in C={C} HW={HW}', lambda: ops.gn_apply(xa, None, B * T, int(HW ** 0.5), int(HW ** 0.5), att,
                                                                   stats1=st, gamma=g, beta=b_, silu=True))
