#!/usr/bin/env python
"""One launch each of the kernels north_star wants ncu rows for, at C2 shapes: the spatial / temporal attention
cores (16x16 and 8x8 levels), gn_apply (fp32 -> bf16 at 64x64, bf16 -> bf16 at 32x32) and the sampler step at the
bench size (B = 8, L2-resident) and at a spill size (B = 256: 5 x 63 MB of operands, larger than the 126 MB L2).

    python profiles/ncu_attn_sampler_probe.py            # prints CUDA-event times (plain run)
    ncu --set full --clock-control none --import-source on -k regex:'attn_|sampler_step|gn_apply' \
        -o gpurun_out/ncu_attn_sampler python profiles/ncu_attn_sampler_probe.py --once
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_b200 import ops  # noqa: E402
from video_diffusion_b200 import create_gaussian_diffusion  # noqa: E402

once = '--once' in sys.argv
dev = 'cuda'
B, T, heads = 8, 20, 4


def run(name, fn, nbytes=0.0, flops=0.0):
    if once:
        fn()
        torch.cuda.synchronize()
        return
    for _ in range(3):
        fn()
    flush = torch.empty(64 << 20, device=dev, dtype=torch.float32)     # 256 MB: larger than L2
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    extra = (f' {nbytes / ms / 1e6:8.0f} GB/s' if nbytes else '') + (f' {flops / ms / 1e9:8.1f} TFLOP/s' if flops else '')
    print(f'{name:48s} {ms * 1e3:8.1f} us{extra}', flush=True)


for C, HW in ((384, 256), (512, 64)):
    hd = C // heads
    M = B * T * HW
    gpt = 1 if HW >= 128 else 128 // HW
    SW = 128 * gpt
    qkv = torch.randn(M, 3 * C, device=dev).bfloat16()
    att = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    run(f'attn_spatial C={C} L={HW}', lambda: ops.attn_spatial(qkv, B * T, HW, heads, hd, att),
        nbytes=qkv.numel() * 2 + att.numel() * 2, flops=4.0 * B * T * heads * HW * HW * hd)
    sk, sq = torch.randn(M, SW, device=dev), torch.randn(M, SW, device=dev)
    mask = torch.ones(B, T, device=dev)
    pm, pv = torch.zeros(M, SW, device=dev, dtype=torch.bfloat16), torch.empty(M, C, device=dev)
    run(f'attn_temporal_tc C={C} HW={HW}',
        lambda: ops.attn_temporal_tc(qkv, sk, sq, mask, True, B, T, HW, heads, hd, gpt, pm, pv),
        nbytes=qkv.numel() * 2 + 2 * sk.numel() * 4 + pm.numel() * 2 + pv.numel() * 4,
        flops=4.0 * B * HW * heads * T * T * hd)
    # the whole temporal block in one kernel (default path)
    R = [torch.randn(B * T * T, C, device=dev) * 0.5 for _ in range(3)]
    rq = torch.empty(B * T * heads * hd * 32, device=dev, dtype=torch.bfloat16)
    rk, rv = torch.empty_like(rq), torch.empty_like(rq)
    ops.rpe_pack(R[0], R[1], R[2], B, T, heads, hd, rq, rk, rv)
    pt = 16 if hd == 96 else 8
    run(f'attn_temporal_fused C={C} HW={HW} pt={pt}',
        lambda: ops.attn_temporal_fused(qkv, rq, rk, rv, mask, True, B, T, HW, heads, hd, 24, att, pixels_per_cta=pt),
        nbytes=qkv.numel() * 2 + att.numel() * 2, flops=10.0 * B * HW * heads * T * T * hd)

for (HWs, C, in_dtype) in ((64, 128, torch.float32), (32, 256, torch.bfloat16)):
    n = B * T
    x = torch.randn(n * HWs * HWs, C, device=dev).to(in_dtype)
    out = torch.empty(n * HWs * HWs, C, device=dev, dtype=torch.bfloat16)
    st = (torch.rand(n, 2, C, device=dev) * 2 ** 24 * HWs * HWs).long()
    st[:, 1] += st[:, 0].abs() * 4
    g, b_ = torch.randn(C, device=dev), torch.randn(C, device=dev)
    run(f'gn_apply {in_dtype} -> bf16 {HWs}x{HWs} C={C}',
        lambda: ops.gn_apply(x, None, n, HWs, HWs, out, stats1=st, gamma=g, beta=b_, silu=True),
        nbytes=x.numel() * x.element_size() + out.numel() * 2)

d = create_gaussian_diffusion(steps=1000, rescale_timesteps=True)
for Bs in (8, 256):
    x = torch.randn(Bs, T, 3, 64, 64, device=dev)
    eps, z = torch.randn_like(x), torch.randn_like(x)
    t = torch.full((Bs,), 500, device=dev, dtype=torch.long)
    sample, pred = torch.empty_like(x), torch.empty_like(x)
    tab = d.tables(x.device)
    run(f'sampler_step (ancestral) B={Bs} [{x.numel() * 4 / 1e6:.0f} MB per tensor]',
        lambda: ops.sampler_step(0, x, eps, z, t, tab, sample=sample, pred_xstart=pred), nbytes=5.0 * x.numel() * 4)
    run(f'sampler_step (ddim)      B={Bs}',
        lambda: ops.sampler_step(1, x, eps, z, t, tab, sample=sample, pred_xstart=pred), nbytes=5.0 * x.numel() * 4)
print('ok')
