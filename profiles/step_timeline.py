#!/usr/bin/env python
"""Timeline of one C2 step as it runs in the product configuration -- two micro-batch streams plus the embedding / RPE
side stream -- from CUDA events around every libvdm launch (eager launches, the host kept ahead of the GPU by a spin
kernel; not under a profiler).  What the CUDA graph replays is the same launches on the same streams.

Prints a JSON summary: span of the step, time with a tcgen05 GEMM running, time with work on >= 2 streams, idle time,
per kernel class the duration when it ran alone on the GPU vs beside another stream's kernel, and the longest gaps.
    python profiles/step_timeline.py > profiles/step_timeline_<tag>.json"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cases, synth  # noqa: E402
from video_diffusion_b200 import create_video_model_and_diffusion, ops, video_model_and_diffusion_defaults  # noqa: E402


def main():
    dev = 'cuda'
    B, F, S = 8, 20, 64
    kw = video_model_and_diffusion_defaults()
    kw.update(cases.ref_config('c2'))
    kw['timestep_respacing'] = ''
    model, diffusion = create_video_model_and_diffusion(compute_dtype=torch.bfloat16, **kw)
    spec = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'spec_c2.json')))
    model.load_state_dict(synth.make_state_dict(spec, seed=1))
    model = model.to(dev).eval()
    x0 = synth.make_video((B, F, 3, S, S), seed=41).to(dev)
    om = torch.zeros(B, F, 1, 1, 1, device=dev)
    om[:, :13] = 1
    mk = dict(x0=x0, obs_mask=om, latent_mask=1 - om, kinda_marg_mask=torch.zeros_like(om),
              frame_indices=torch.arange(F, device=dev).view(1, F).repeat(B, 1), x_t_minus_1=x0, observed_frames='x_0')
    x = torch.randn(B, F, 3, S, S, device=dev)
    t = torch.full((B,), 500, device=dev, dtype=torch.long)

    def step():
        return diffusion.p_sample(model, x, t, clip_denoised=True, model_kwargs=mk)['sample']

    with torch.no_grad():
        for _ in range(3):                       # graph path: allocates every buffer, warms the clocks
            step()
        model.use_cuda_graph = False
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        best = None
        for _ in range(3):
            ops.PROFILE, ops.PROFILE_STREAMS = [], []
            base = torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(40_000_000)        # ~20 ms head start for the host
            base.record()
            step()
            torch.cuda.synchronize()
            recs = [(name, base.elapsed_time(e0), base.elapsed_time(e1), s, meta)
                    for (name, e0, e1, _, _, meta), s in zip(ops.PROFILE, ops.PROFILE_STREAMS)]
            ops.PROFILE = ops.PROFILE_STREAMS = None
            span = max(r[2] for r in recs) - min(r[1] for r in recs)
            if best is None or span < best[0]:
                best = (span, recs)
    span, recs = best
    t_min = min(r[1] for r in recs)
    recs = sorted(((n, a - t_min, b - t_min, s, m) for n, a, b, s, m in recs), key=lambda r: (r[1], r[2]))
    streams = sorted({r[3] for r in recs})
    sid = {s: i for i, s in enumerate(streams)}
    # sweep: coverage by number of concurrently running kernels / by GEMM activity
    pts = []
    for n, a, b, s, m in recs:
        pts.append((a, 1, n.startswith('gemm_tc')))
        pts.append((b, -1, n.startswith('gemm_tc')))
    pts.sort()
    active = gemm = 0
    last = 0.0
    cover = {0: 0.0, 1: 0.0, 2: 0.0}
    gemm_time = gemm2_time = 0.0
    for tt, d, g in pts:
        dt = tt - last
        cover[min(active, 2)] += dt
        if gemm:
            gemm_time += dt
        if gemm >= 2:
            gemm2_time += dt
        last = tt
        active += d
        gemm += d if g else 0
    # per class: alone vs overlapped durations
    per = {}
    for i, (n, a, b, s, m) in enumerate(recs):
        ov = any(j != i and r[3] != s and r[1] < b - 1e-4 and r[2] > a + 1e-4 for j, r in enumerate(recs))
        d = per.setdefault(n, dict(alone_ms=0.0, alone_n=0, overlapped_ms=0.0, overlapped_n=0))
        key = 'overlapped' if ov else 'alone'
        d[key + '_ms'] += b - a
        d[key + '_n'] += 1
    # per stream busy time and the longest gaps on the union
    busy = {sid[s]: sum(b - a for n, a, b, ss, m in recs if ss == s) for s in streams}
    edges = sorted((a, b) for n, a, b, s, m in recs)
    gaps, cur_end = [], edges[0][1]
    for a, b in edges[1:]:
        if a > cur_end:
            gaps.append((a - cur_end, cur_end))
        cur_end = max(cur_end, b)
    gaps.sort(reverse=True)
    out = dict(
        what='one eager C2 step (B = 8, F = 20, 64x64, bf16 mode) on its product streams, CUDA events per launch',
        span_ms=span, launches=len(recs), streams=len(streams),
        busy_ms_per_stream=busy,
        ms_with_0_1_2plus_kernels_running=cover,
        ms_with_a_gemm_running=gemm_time, ms_with_two_gemms_running=gemm2_time,
        per_class={k: {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items()}
                   for k, v in sorted(per.items(), key=lambda kv: -(kv[1]['alone_ms'] + kv[1]['overlapped_ms']))},
        longest_gaps_ms=[dict(gap=round(g, 4), at=round(at, 3)) for g, at in gaps[:8]],
        timeline=[dict(k=n, s=sid[s], t0=round(a, 4), t1=round(b, 4), m=m[:60]) for n, a, b, s, m in recs],
    )
    print(json.dumps(out))


if __name__ == '__main__':
    main()
