"""Fused temporal attention kernel vs the three-launch path (grouped Sk/Sq GEMM + mma.sync core + grouped P.R_v GEMM)
at the two attention levels of the C2 model, CUDA events over graph replays, cold-ish L2 (the buffers of several
blocks are cycled).  python profiles/temporal_fused_microbench.py > profiles/temporal_fused_microbench_<tag>.json"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_b200 import ops as o  # noqa: E402


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device='cpu').manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


def time_graph(fn, n_rep=20, n_bufs=1):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(n_bufs):
            fn(i)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(n_bufs):
                fn(i)
        g.replay()
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(n_rep):
            g.replay()
        e1.record(s)
        s.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (n_rep * n_bufs)   # us per call


def main():
    out = []
    B, T, heads = 8, 20, 4
    for HW, hd in ((256, 96), (64, 128)):
        Cc, M, TP = heads * hd, B * T * HW, 24
        nb = 6                                        # distinct buffer sets cycled per replay (> L2 for 16x16)
        qkvs = [rnd(M, 3 * Cc, seed=10 + i).bfloat16() for i in range(nb)]
        R = [rnd(B * T * T, Cc, seed=2 + i, scale=0.5) for i in range(3)]
        mask = torch.ones(B, T).cuda()
        rq = torch.empty(B * T, heads, hd * 32, device='cuda', dtype=torch.bfloat16)
        rk, rv = torch.empty_like(rq), torch.empty_like(rq)
        o.rpe_pack(R[0], R[1], R[2], B, T, heads, hd, rq, rk, rv)
        atts = [torch.empty(M, Cc, device='cuda', dtype=torch.bfloat16) for _ in range(nb)]
        row = dict(HW=HW, hd=hd, M=M)
        for pt in ((8, 16) if hd == 96 else (8,)):
            row[f'fused_pt{pt}_us'] = time_graph(
                lambda i: o.attn_temporal_fused(qkvs[i], rq, rk, rv, mask, True, B, T, HW, heads, hd, TP, atts[i],
                                                pixels_per_cta=pt), n_bufs=nb)
        # per-phase cycles of thread 0, averaged over the CTAs of one launch (vdm_attn_temporal_fused_set_trace)
        from video_diffusion_b200 import _lib
        tr = torch.zeros(8, device='cuda', dtype=torch.int64)
        _lib.load().vdm_attn_temporal_fused_set_trace(tr.data_ptr())
        for pt in ((8, 16) if hd == 96 else (8,)):
            tr.zero_()
            o.attn_temporal_fused(qkvs[0], rq, rk, rv, mask, True, B, T, HW, heads, hd, TP, atts[0], pixels_per_cta=pt)
            torch.cuda.synchronize()
            c = tr.cpu().tolist()
            row[f'phase_cycles_pt{pt}'] = dict(zip(('stage_wait', 'p1a', 'p1b', 'p2a', 'p2b', 'p3'),
                                                   [round(v / max(c[7], 1)) for v in c[:6]]), ctas=c[7])
        _lib.load().vdm_attn_temporal_fused_set_trace(None)
        row['rpe_pack_us'] = time_graph(lambda i: o.rpe_pack(R[0], R[1], R[2], B, T, heads, hd, rq, rk, rv))
        # the three-launch path
        gpt, tpg = (1, HW // 128) if HW >= 128 else (128 // HW, 1)
        SW, ntg = 128 * gpt, (B * T + gpt - 1) // gpt
        bkq = torch.zeros(2, ntg * SW, Cc, device='cuda', dtype=torch.bfloat16)
        bv = torch.zeros(ntg * Cc, SW, device='cuda', dtype=torch.bfloat16)
        o.rpe_expand(R[0], R[1], R[2], B, T, heads, hd, gpt, bkq[1], bkq[0], bv)
        lin = dict(n_img=M, H=1, W=1, taps=1)
        sksq = [torch.empty(2, M, SW, device='cuda') for _ in range(nb)]
        pm = [torch.zeros(M, SW, device='cuda', dtype=torch.bfloat16) for _ in range(nb)]
        pv = [torch.empty(M, Cc, device='cuda') for _ in range(nb)]
        att_old = [torch.empty(M, Cc, device='cuda', dtype=torch.bfloat16) for _ in range(nb)]

        def old(i):
            o.gemm(qkvs[i][:, :Cc], bkq.view(2 * ntg * SW, Cc), SW, out_f32=sksq[i], w_group_tiles=tpg, C1=Cc, n_prob=2,
                   prob_a_cols=Cc, prob_w_rows=ntg * SW, prob_out_stride=M * SW, **lin)
            o.attn_temporal_tc(qkvs[i], sksq[i][0], sksq[i][1], mask, True, B, T, HW, heads, hd, gpt, pm[i], pv[i])
            o.gemm(pm[i], bv, Cc, residual=pv[i], out_bf16=att_old[i], w_group_tiles=tpg, **lin)
        row['three_launch_us'] = time_graph(old, n_bufs=nb)
        torch.cuda.synchronize()
        row['max_abs_diff_vs_three_launch'] = float((atts[0].float() - att_old[0].float()).abs().max())
        row['hbm_floor_us'] = (M * 3 * Cc * 2 + M * Cc * 2) / 6535.4e9 * 1e6
        out.append(row)
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
