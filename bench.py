#!/usr/bin/env python
"""Headline benchmark: denoised video frames/sec of the hot path (U-Net forward driven by the
ancestral sampler step) on BASELINE.json's config 2 -- MineRL-sized FDM U-Net, 64x64,
max_frames=20, batch 8 per GPU, bf16 tensor-core mode, synthetic video, de-zeroed random weights.

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]

A "step" is one `diffusion.p_sample` call: conditioning mix -> U-Net forward -> fused sampler
kernel, i.e. B*F = 160 frames denoised once.  Prints ONE JSON line (see the task contract):
  value  : frames/s with every input resident in HBM (CUDA-graph replay of the forward),
  e2e    : the same through the public API with pinned HOST buffers (H2D of x/x0/masks/indices,
           D2H of the sample inside the timed region),
  roofline: tensor-core roofline of the dominant kernel class (gemm_tc: every conv / linear),
           per-launch CUDA-event times gathered live in a separate eager profiling pass,
  cpu_baseline: the CPU oracle port of the reference path timed on this box's host cores.
  stock_gpu_baseline (only with --stock-gpu-baseline): the same forward through PyTorch eager cuDNN / cuBLAS kernels on
           this GPU (TF32 and bf16 autocast) -- the stock-library figure of SURVEY 8(d), reported next to the value.
`--impl reference` times that CPU path alone (the reference is pure PyTorch; it cannot travel to
the GPU box, so its restatement in oracle/ stands in -- kind "port").
Multi-GPU: one process per GPU (torchrun), the batch of videos is sharded, no collective on the
data path (weak scaling); timing is the max over ranks.
"""
import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = 'c2'
B_PER_GPU, FRAMES, SIZE = 8, 20, 64
N_OBS = 13                       # autoreg, step_size 7: 13 observed + 7 latent frames per window
FLOP_PER_FRAME = 63.40e9         # 10.144 TFLOP per (8 x 20)-frame forward
GEMM_FLOP_SHARE = (9238.5 + 234.9 + 555.7) / 10144.0   # conv + addmm + mm share of it (the gemm_tc launches)
METRIC = 'denoised video frames/sec (U-Net fwd, bf16)'
WORKLOAD = ('C2: MineRL-sized FDM U-Net 64x64 (ch 128, mult 1-2-3-4, 2 res blocks, 4 heads, '
            'attention at 16x16 and 8x8), max_frames=20, autoreg window 13 obs + 7 latent, '
            'one ancestral p_sample step per step')


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return dict(tflops=p['bf16_tflops_sustained'], hbm=p['hbm_gbs'], src='MEASURED_PEAKS.json (sustained bf16)')
    except Exception:
        return dict(tflops=1400.0, hbm=6650.0, src='fallback (B200_PROFILING.md)')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['unavailable'])
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == 'Active' for r in self.rows)]
        busy = [v for v in sm if v > 0.5 * sm[-1]] or sm
        return dict(sm_mhz=busy[len(busy) // 2], sm_max_mhz=int(self.rows[0][1]), reasons=reasons,
                    samples=len(sm))


def synth_state(cfg_name):
    from oracle import synth
    with open(os.path.join(ROOT, 'tests', 'golden', f'spec_{cfg_name}.json')) as f:
        spec = json.load(f)
    return synth.make_state_dict(spec, seed=1)


def window_inputs(B, seed):
    """One autoreg window of synthetic video: 13 observed + 7 latent frames (host tensors)."""
    from oracle import synth
    x0 = synth.make_video((B, FRAMES, 3, SIZE, SIZE), seed=seed)
    obs = torch.zeros(B, FRAMES, 1, 1, 1)
    obs[:, :N_OBS] = 1
    fi = torch.arange(23, 23 + FRAMES).view(1, FRAMES).repeat(B, 1)
    return dict(x0=x0, obs_mask=obs, latent_mask=1 - obs, kinda_marg_mask=torch.zeros_like(obs), frame_indices=fi)


def cpu_port_step(sd, batch, threads):
    """One step of the reference path on the CPU (oracle port): U-Net forward + ancestral sampler step."""
    from oracle import cases, diffusion_oracle as D, synth, unet_oracle as U
    cfg = U.model_config(**cases.ref_config(CFG))
    sched = D.Schedule(1000, 'linear', '')
    w = window_inputs(batch, seed=2)
    x = w['x0'].clone()
    t = torch.full((batch,), 500, dtype=torch.long)
    noise = synth.make_noise(tuple(x.shape), seed=4)
    torch.set_num_threads(threads)

    def step():
        with torch.no_grad():
            eps = U.cond_marg_forward(sd, cfg, x, w['x0'], w['obs_mask'], w['latent_mask'], w['kinda_marg_mask'],
                                      sched.model_time(t), w['frame_indices'])
            return D.p_sample(sched, eps, x, t, noise)['sample']
    return step


def stock_gpu_baseline(sd, batch, steps=5):
    """The same U-Net forward through stock PyTorch eager kernels (cuDNN / cuBLAS) on this GPU: the oracle's functional
    restatement moved to cuda:0, in the reference's own TF32 setting (scripts/video_sample.py:21-22) and under bf16
    autocast.  A reported baseline ("the stock-library kernel to beat"), opt-in, never on the product path."""
    from oracle import cases, diffusion_oracle as D, unet_oracle as U
    cfg = U.model_config(**cases.ref_config(CFG))
    sched = D.Schedule(1000, 'linear', '')
    dev = torch.device('cuda', torch.cuda.current_device())
    sdg = {k: v.to(dev) for k, v in sd.items()}
    w = {k: v.to(dev) for k, v in window_inputs(batch, seed=2).items()}
    x = w['x0'].clone()
    t = sched.model_time(torch.full((batch,), 500, dtype=torch.long)).to(dev)
    out = {}
    saved = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = True
    try:
        for name, ctx in (('tf32', contextlib.nullcontext()), ('bf16_autocast', torch.autocast('cuda', torch.bfloat16))):
            def fwd():
                with torch.no_grad(), ctx:
                    return U.cond_marg_forward(sdg, cfg, x, w['x0'], w['obs_mask'], w['latent_mask'],
                                               w['kinda_marg_mask'], t, w['frame_indices'])
            for _ in range(2):
                fwd()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fwd()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {'ms_per_forward': ms, 'frames_per_s': batch * FRAMES / ms * 1e3}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = saved
    out['what'] = (f'torch {torch.__version__} eager (cuDNN/cuBLAS) U-Net forward of the oracle restatement on the GPU, '
                   f'({batch},{FRAMES},3,{SIZE},{SIZE}) window, mean of {steps} after 2 warm-ups, CUDA events; forward only')
    return out


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sd = synth_state(CFG)
    step = cpu_port_step(sd, 1, threads)
    for _ in range(max(1, min(args.warmup, 1))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    fps = FRAMES / dt
    sample = f'oracle port of the reference CPU path, one (1,{FRAMES},3,{SIZE},{SIZE}) window per step (batch 1 of 8), fp32'
    print(json.dumps({
        'metric': METRIC, 'value': fps, 'unit': 'frames/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'impl': 'reference',
        'config': {'workload': 'C2: MineRL-sized FDM U-Net 64x64, max_frames=20, autoreg window 13 obs + 7 latent, '
                               'one ancestral p_sample step', 'frames_per_step': FRAMES},
        'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--stock-gpu-baseline', action='store_true',
                    help='also time the stock PyTorch eager (cuDNN/cuBLAS) forward on this GPU (reported baseline)')
    ap.add_argument('--profile-json', default=None, help='write the per-kernel-class breakdown here')
    ap.add_argument('--workload', default='c2', choices=['c2', 'c4'],
                    help="c2 (default) is the configuration BASELINE.json's metric is quoted on; c4 = the 128x128 model")
    args = ap.parse_args()
    if args.workload == 'c4':
        # BASELINE.json configs[3]: 128x128, channel_mult (1,1,2,3,4), 2 res blocks; SURVEY 8(d): 21.713 TFLOP per
        # (8 x 20)-frame forward, conv 95.8 %, the attention-level linears as in C2
        global CFG, SIZE, FLOP_PER_FRAME, GEMM_FLOP_SHARE, WORKLOAD
        CFG, SIZE, FLOP_PER_FRAME = 'c4', 128, 135.70e9
        GEMM_FLOP_SHARE = (0.958 * 21713.0 + 234.9 + 555.7) / 21713.0
        WORKLOAD = ('C4: 128x128 FDM U-Net (ch 128, mult 1-1-2-3-4, 2 res blocks, 4 heads, attention at 16x16 and 8x8), '
                    'max_frames=20, autoreg window 13 obs + 7 latent, one ancestral p_sample step per step')
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    if args.impl == 'reference':
        return run_reference(args)

    import torch.distributed as dist
    from video_diffusion_b200 import _lib, ops
    from video_diffusion_b200 import create_video_model_and_diffusion, video_model_and_diffusion_defaults
    from oracle import cases
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    _lib.load()

    kw = video_model_and_diffusion_defaults()
    kw.update(cases.ref_config(CFG))
    model, diffusion = create_video_model_and_diffusion(compute_dtype=torch.bfloat16, **kw)
    sd = synth_state(CFG)
    model.load_state_dict(sd)
    model = model.to(dev).eval()

    B = B_PER_GPU
    host = window_inputs(B, seed=2 + rank)                     # each rank denoises its own shard of videos
    pinned = {k: v.pin_memory() for k, v in host.items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    mk = dict(resident, x_t_minus_1=resident['x0'], observed_frames='x_0')
    x_dev = resident['x0'].clone()
    t_dev = torch.full((B,), 500, device=dev, dtype=torch.long)

    def step_resident():
        return diffusion.p_sample(model, x_dev, t_dev, clip_denoised=True, model_kwargs=mk)['sample']

    x_pin = host['x0'].clone().pin_memory()
    out_pin = torch.empty_like(x_pin).pin_memory()

    def step_e2e():
        d = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
        x = x_pin.to(dev, non_blocking=True)
        s = diffusion.p_sample(model, x, t_dev, clip_denoised=True,
                               model_kwargs=dict(d, x_t_minus_1=d['x0'], observed_frames='x_0'))['sample']
        out_pin.copy_(s, non_blocking=True)
        return s

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(ms.item())

    with torch.no_grad():
        # launches per step, counted once in eager mode
        model.use_cuda_graph = False
        n0 = _lib.launch_count()
        step_resident()
        launches_per_step = _lib.launch_count() - n0
        # per-kernel-class CUDA-event profile (eager), 3 passes, first discarded
        prof, prof_shapes = {}, {}
        for it in range(3):
            ops.PROFILE = []
            # head start for the host: the GPU spins ~8 ms while the whole step is enqueued, so each event pair
            # brackets a kernel that runs back to back with its neighbours instead of the host's launch latency
            torch.cuda._sleep(16_000_000)
            step_resident()
            torch.cuda.synchronize()
            if it:
                for name, e0, e1, fl, nb, meta in ops.PROFILE:
                    if meta and it == 2:
                        shapes = prof_shapes.setdefault((name, meta), dict(ms=0.0, flops=fl, n=0, bytes=nb))
                        shapes['ms'] += e0.elapsed_time(e1)
                        shapes['n'] += 1
                    d = prof.setdefault(name, dict(ms=0.0, flops=0.0, bytes=0.0, launches=0))
                    d['ms'] += e0.elapsed_time(e1) / 2
                    d['flops'] += fl / 2
                    d['bytes'] += nb / 2
                    d['launches'] += 0.5
            ops.PROFILE = None
        model.use_cuda_graph = True
        for _ in range(args.warmup):
            step_resident()
        with ClockSampler(local) as clk:
            ms = timed(step_resident, args.steps)
        for _ in range(args.warmup):
            step_e2e()
        ms_e2e = timed(step_e2e, args.steps)

    frames = world * B * FRAMES * args.steps
    value = frames / (ms / 1e3)
    e2e_value = frames / (ms_e2e / 1e3)
    pk = peaks()
    gemm = {k: v for k, v in prof.items() if k.startswith('gemm_tc')}
    g_ms = sum(v['ms'] for v in gemm.values())
    g_fl = sum(v['flops'] for v in gemm.values())
    g_n = sum(v['launches'] for v in gemm.values())
    # ALGORITHMIC FLOPs of the launches (the reference's conv + linear FLOPs for these frames; the kernel executes
    # fewer for the folded upsample convs and more for the block-diagonal RPE GEMMs) / summed launch time
    alg_fl = FLOP_PER_FRAME * GEMM_FLOP_SHARE * B * FRAMES
    achieved = alg_fl / (g_ms / 1e3) / 1e12 if g_ms else 0.0
    executed = g_fl / (g_ms / 1e3) / 1e12 if g_ms else 0.0
    total_prof_ms = sum(v['ms'] for v in prof.values())
    h2d = sum(v.numel() * v.element_size() for v in pinned.values()) + x_pin.numel() * 4
    result = {
        'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
        'config': {'workload': WORKLOAD,
                   'batch_per_gpu': B, 'frames_per_step_per_gpu': B * FRAMES, 'parallelism': f'dp{world} (videos sharded)',
                   'weights': 'random, de-zeroed (oracle/synth.py seed 1)', 'timestep_respacing': '',
                   'l2': 'no explicit flush: activations touched per step (several GB) exceed the 126 MB L2',
                   'model_tflop_per_step_per_gpu': FLOP_PER_FRAME * B * FRAMES / 1e12,
                   'model_tflops_achieved_per_gpu': FLOP_PER_FRAME * B * FRAMES / (ms / args.steps / 1e3) / 1e12},
        'clocks': clk.summary(),
        'e2e': {'value': e2e_value, 'unit': 'frames/s', 'h2d_bytes_per_step': int(h2d),
                'd2h_bytes_per_step': int(out_pin.numel() * 4), 'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': int(launches_per_step * args.steps),
        'roofline': {'bound': 'tensor', 'kernel': 'gemm_tc_halo_kernel + gemm_tc_kernel (all conv3x3 / 1x1 / linear launches of one step)',
                     'achieved': achieved, 'peak': pk['tflops'], 'unit': 'TFLOP/s',
                     'frac': achieved / pk['tflops'] if pk['tflops'] else None, 'traffic': None,
                     'peak_source': pk['src'], 'launches_per_step': g_n, 'kernel_ms_per_step': g_ms,
                     'share_of_step': g_ms / total_prof_ms if total_prof_ms else None,
                     'flops_per_step': alg_fl, 'executed_flops_per_step': g_fl, 'executed_tflops': executed,
                     'timing': 'CUDA events per launch, eager pass with the launch queue kept full (not under a profiler)'},
    }
    # the single heaviest launch shape of the step (3x3 conv 64x64, 256 -> 128, transposed-role halo kernel), per launch: live CUDA-event
    # time of this run; DRAM traffic from the committed `ncu --set full` capture of the same launch
    # (profiles/ncu_gemm_halo_r1h.md: 336.5 MB read + 141.1 MB written; algorithmic 335.5 + 167.8 MB)
    top = [(k, v) for k, v in prof_shapes.items() if k[0] == 'gemm_tc_conv3x3' and 'M=655360 N=128 K=2304' in k[1]]
    if top:
        (_, meta), v = top[0]
        t_launch = v['ms'] / v['n'] / 1e3
        tf = v['flops'] / t_launch / 1e12
        result['roofline']['traffic'] = 477.6e6
        result['roofline']['top_launch'] = {
            'kernel': 'gemm_tc_halo_t_kernel<3,4,6>: conv3x3 64x64 256->128, ' + meta, 'us': t_launch * 1e6,
            'achieved': tf, 'frac': tf / pk['tflops'] if pk['tflops'] else None, 'flops': v['flops'],
            'algorithmic_bytes': 655360 * 256 * 2 + 655360 * 128 * 2, 'traffic': 477.6e6,
            'traffic_source': 'ncu --set full, profiles/ncu_gemm_halo_r1h.md (dram__bytes_read.sum + dram__bytes_write.sum)'}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        step = cpu_port_step(sd, 1, threads)
        step()
        t0 = time.perf_counter()
        n = 3
        for _ in range(n):
            step()
        dt = (time.perf_counter() - t0) / n
        result['cpu_baseline'] = {'value': FRAMES / dt, 'unit': 'frames/s', 'cores': threads, 'kind': 'port',
                                  'sample': f'oracle port of the reference CPU path: U-Net forward + p_sample on one '
                                            f'(1,{FRAMES},3,{SIZE},{SIZE}) window (batch 1 of 8), mean of {n} after 1 warm-up'}
    if rank == 0 and world == 1 and args.stock_gpu_baseline:
        result['stock_gpu_baseline'] = stock_gpu_baseline(sd, B)
    if rank == 0:
        if args.profile_json:
            with open(args.profile_json, 'w') as f:
                json.dump({'per_kernel_class_per_step': prof, 'ms_per_step_graph': ms / args.steps,
                           'ms_per_step_eager_sum': total_prof_ms,
                           'gemm_shapes': [dict(kernel=k[0], shape=k[1], launches=v['n'], ms_total=v['ms'],
                                                tflops=v['flops'] * v['n'] / v['ms'] / 1e9 if v['ms'] else 0,
                                                gbs=v['bytes'] * v['n'] / v['ms'] / 1e6 if v['ms'] else 0)
                                           for k, v in sorted(prof_shapes.items(), key=lambda kv: -kv[1]['ms'])]},
                          f, indent=1)
        print(json.dumps(result))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
