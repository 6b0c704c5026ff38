#!/usr/bin/env python
"""Headline benchmark: denoised video frames/sec of the hot path (U-Net forward driven by the
ancestral sampler step) on BASELINE.json's config 2 -- MineRL-sized FDM U-Net, 64x64,
max_frames=20, batch 8 per GPU, bf16 tensor-core mode, synthetic video, de-zeroed random weights.

    python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--workload c2|c2chain|c4|c3|c5]

A "step" is one `diffusion.p_sample` call: conditioning mix -> U-Net forward -> fused sampler
kernel, i.e. B*F = 160 frames denoised once.  Prints ONE JSON line (see the task contract):
  value  : frames/s with every input resident in HBM (CUDA-graph replay of the forward),
  e2e    : the same through the public API with pinned HOST buffers (H2D of x/x0/masks/indices,
           D2H of the sample inside the timed region),
  roofline: `frac` = the north-star number, ALGORITHMIC FLOPs of one whole forward (63.40 GFLOP per
           frame) / ms_per_step / measured sustained bf16 peak; `frac_gemm` = the same over the summed CUDA-event
           times of the tcgen05 GEMM launches only (the dominant kernel class), gathered live in a separate eager pass,
  roofline_hbm: achieved GB/s of the bandwidth-bound kernels against the measured HBM peak -- every standalone
           GroupNorm-apply launch of a step (live events) and the sampler step at a size that spills L2 (B = 256),
  cpu_baseline / `--impl reference`: the REFERENCE's own CPU path (baseline/_ref, installed by baseline/install_ref.sh;
           kind "reference") on this box's host cores, same batch of 8 windows; the oracle port stands in (kind "port")
           only when baseline/_ref is absent,
  stock_gpu_baseline: the reference's eager CUDA forward (cuDNN / cuBLAS; TF32 as scripts/video_sample.py:21-22 sets it,
           and under bf16 autocast) on this GPU -- the stock-library figure of SURVEY 8(d).
Secondary workloads (own JSON line, not the driver's): `--workload c3` = BASELINE configs[2], a T=300 `exp-past`
DDIM-100 `infer_video` job over rank-sharded videos ending in the NCCL gather (sampled frames/s); `--workload c5` =
configs[4], `calc_bpd_loop_subsampled` over all 1000 timesteps (forward frames/s, videos/s); `--workload c4` = the
128x128 model.
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
import types

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
REF_DIR = os.path.join(ROOT, 'baseline', '_ref')

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
        return dict(tflops=p['bf16_tflops_sustained'], burst=p['bf16_tflops'], hbm=p['hbm_gbs'],
                    src='MEASURED_PEAKS.json (sustained bf16; hbm_gbs)')
    except Exception:
        return dict(tflops=1400.0, burst=1590.0, hbm=6650.0, src='fallback (B200_PROFILING.md)')


def shared_config(world):
    """`config` of BOTH arms (this one and --impl reference): same workload, same batch."""
    return {'workload': WORKLOAD, 'batch_per_gpu': B_PER_GPU, 'frames_per_step_per_gpu': B_PER_GPU * FRAMES,
            'parallelism': f'dp{world} (videos sharded)', 'weights': 'random, de-zeroed (oracle/synth.py seed 1)',
            'timestep_respacing': '',
            'l2': 'no explicit flush: activations touched per step (several GB) exceed the 126 MB L2',
            'model_tflop_per_step_per_gpu': FLOP_PER_FRAME * B_PER_GPU * FRAMES / 1e12}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw,power.limit')

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '50'],
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
        out = dict(sm_mhz=busy[len(busy) // 2], sm_max_mhz=int(self.rows[0][1]), reasons=reasons, samples=len(sm))
        try:        # board power under load against the enforced limit: the step runs into the power cap
            pw = sorted(float(r[6]) for r in self.rows if len(r) > 7 and int(r[0]) > 0.5 * sm[-1])
            out['power_w'], out['power_limit_w'] = pw[len(pw) // 2], float(self.rows[0][7])
        except Exception:
            pass
        return out


def synth_state(cfg_name):
    from oracle import synth
    with open(os.path.join(ROOT, 'tests', 'golden', f'spec_{cfg_name}.json')) as f:
        spec = json.load(f)
    return synth.make_state_dict(spec, seed=1)


def window_inputs(B, seed, frames=None):
    """One autoreg window of synthetic video: 13 observed + 7 latent frames (host tensors)."""
    from oracle import synth
    F = frames or FRAMES
    x0 = synth.make_video((B, F, 3, SIZE, SIZE), seed=seed)
    obs = torch.zeros(B, F, 1, 1, 1)
    obs[:, :N_OBS] = 1
    fi = torch.arange(23, 23 + F).view(1, F).repeat(B, 1)
    return dict(x0=x0, obs_mask=obs, latent_mask=1 - obs, kinda_marg_mask=torch.zeros_like(obs), frame_indices=fi)


# ----------------------------------------------------------------------------------------------- reference arms
def have_reference():
    return os.path.exists(os.path.join(REF_DIR, 'improved_diffusion', 'unet.py'))


def reference_model_and_diffusion(sd, device='cpu'):
    """The UNMODIFIED reference (baseline/_ref, see baseline/install_ref.sh) built by its own factory with this
    benchmark's config and the same synthetic weights."""
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    if 'lpips' not in sys.modules:             # inference_util imports it at module scope; not used here
        lp = types.ModuleType('lpips')
        lp.LPIPS = type('LPIPS', (), {'__init__': lambda s, *a, **k: None})
        sys.modules['lpips'] = lp
    from improved_diffusion.script_util import create_video_model_and_diffusion, video_model_and_diffusion_defaults
    from oracle import cases
    kw = video_model_and_diffusion_defaults()
    kw.update(cases.ref_config(CFG))
    model, diffusion = create_video_model_and_diffusion(**kw)
    model.load_state_dict(sd)
    return model.to(device).eval(), diffusion


def cpu_step(sd, batch, threads):
    """One step of the reference path on the CPU: U-Net forward + ancestral sampler step on a batch of windows.
    Returns (step function, kind, description)."""
    from oracle import synth
    w = window_inputs(batch, seed=2)
    x = w['x0'].clone()
    t = torch.full((batch,), 500, dtype=torch.long)
    noise = synth.make_noise(tuple(x.shape), seed=4)
    torch.set_num_threads(threads)
    if have_reference():
        model, diffusion = reference_model_and_diffusion(sd)      # puts baseline/_ref on sys.path
        import improved_diffusion.gaussian_diffusion as rgd
        kw = dict(w, x_t_minus_1=w['x0'], observed_frames='x_0')

        def step():
            saved = rgd.th.randn_like
            rgd.th.randn_like = lambda like: noise
            try:
                with torch.no_grad():
                    return diffusion.p_sample(model, x, t, clip_denoised=True, model_kwargs=kw)['sample']
            finally:
                rgd.th.randn_like = saved
        return step, 'reference', "the reference's own improved_diffusion package (baseline/_ref), PyTorch CPU fp32"
    from oracle import cases, diffusion_oracle as D, unet_oracle as U
    cfg = U.model_config(**cases.ref_config(CFG))
    sched = D.Schedule(1000, 'linear', '')

    def step():
        with torch.no_grad():
            eps = U.cond_marg_forward(sd, cfg, x, w['x0'], w['obs_mask'], w['latent_mask'], w['kinda_marg_mask'],
                                      sched.model_time(t), w['frame_indices'])
            return D.p_sample(sched, eps, x, t, noise)['sample']
    return step, 'port', 'oracle port of the reference CPU path (baseline/_ref absent), PyTorch CPU fp32'


def stock_gpu_baseline(sd, batch, steps=5):
    """The same U-Net forward through stock PyTorch eager kernels (cuDNN / cuBLAS) on this GPU -- the reference's own
    model when baseline/_ref is present (else the oracle's functional restatement) -- in the reference's TF32 setting
    (scripts/video_sample.py:21-22) and under bf16 autocast.  A reported baseline ("the stock-library kernel
    to beat"), never on the product path."""
    dev = torch.device('cuda', torch.cuda.current_device())
    w = {k: v.to(dev) for k, v in window_inputs(batch, seed=2).items()}
    x = w['x0'].clone()
    if have_reference():
        model, diffusion = reference_model_and_diffusion(sd, dev)
        t = diffusion._scale_timesteps(torch.full((batch,), 500, device=dev, dtype=torch.long))
        kw = dict(w, x_t_minus_1=w['x0'], observed_frames='x_0')
        call = lambda: model(x, timesteps=t, **kw)
        what = "the reference's CondMargVideoModel (baseline/_ref)"
    else:
        from oracle import cases, diffusion_oracle as D, unet_oracle as U
        cfg = U.model_config(**cases.ref_config(CFG))
        sdg = {k: v.to(dev) for k, v in sd.items()}
        t = D.Schedule(1000, 'linear', '').model_time(torch.full((batch,), 500, dtype=torch.long)).to(dev)
        call = lambda: U.cond_marg_forward(sdg, cfg, x, w['x0'], w['obs_mask'], w['latent_mask'], w['kinda_marg_mask'],
                                           t, w['frame_indices'])
        what = "the oracle's functional restatement (baseline/_ref absent)"
    out = {}
    saved = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = True
    try:
        for name, ctx in (('tf32', contextlib.nullcontext()), ('bf16_autocast', torch.autocast('cuda', torch.bfloat16))):
            def fwd():
                with torch.no_grad(), ctx:
                    return call()
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
    out['what'] = (f'torch {torch.__version__} eager (cuDNN/cuBLAS) forward of {what} on the GPU, '
                   f'({batch},{FRAMES},3,{SIZE},{SIZE}) window, mean of {steps} after 2 warm-ups, CUDA events; forward only')
    return out


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path on all host cores, same config (batch of 8
    windows per step).  Rank 0 only."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sd = synth_state(CFG)
    step, kind, what = cpu_step(sd, B_PER_GPU, threads)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    fps = B_PER_GPU * FRAMES / dt
    sample = f'{what}: U-Net forward + p_sample on one ({B_PER_GPU},{FRAMES},3,{SIZE},{SIZE}) batch of windows per step'
    print(json.dumps({
        'metric': METRIC, 'value': fps, 'unit': 'frames/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'impl': 'reference',
        'config': shared_config(1),
        'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': threads, 'kind': kind, 'sample': sample},
        'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }))


# ----------------------------------------------------------------------------------------------- HBM rooflines
def sampler_spill_roofline(diffusion, dev, pk):
    """The sampler step at a size that spills the 126 MB L2 (B = 256: five 63 MB operands), CUDA events, median of 10."""
    from video_diffusion_b200 import ops
    Bs = 256
    x = torch.randn(Bs, FRAMES, 3, 64, 64, device=dev)
    eps, z = torch.randn_like(x), torch.randn_like(x)
    t = torch.full((Bs,), 500, device=dev, dtype=torch.long)
    sample, pred = torch.empty_like(x), torch.empty_like(x)
    tab = diffusion.tables(dev)
    ts = []
    for i in range(13):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.sampler_step(0, x, eps, z, t, tab, sample=sample, pred_xstart=pred)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    nbytes = 5.0 * x.numel() * 4                       # x, eps, z in; sample, pred_xstart out (20 B / element)
    gbs = nbytes / ms / 1e6
    return {'kernel': 'sampler_step_kernel<4> (ancestral, + pred_xstart)', 'bound': 'hbm', 'achieved': gbs,
            'peak': pk['hbm'], 'unit': 'GB/s', 'frac': gbs / pk['hbm'], 'bytes_per_launch': nbytes, 'us': ms * 1e3,
            'size': f'B={Bs} x {FRAMES} x 3 x 64 x 64 fp32 (63 MB per tensor, 315 MB per launch: spills L2)'}


def gn_apply_roofline(prof_shapes, pk):
    rows = [(k, v) for k, v in prof_shapes.items() if k[0] == 'gn_apply' and 'HxW=1x1' not in k[1]]
    if not rows:
        return None
    ms = sum(v['ms'] for _, v in rows)
    nbytes = sum(v['bytes'] * v['n'] for _, v in rows)
    gbs = nbytes / ms / 1e6
    return {'kernel': 'gn_apply_kernel (every standalone GroupNorm-apply / cast launch of one step)', 'bound': 'hbm',
            'achieved': gbs, 'peak': pk['hbm'], 'unit': 'GB/s', 'frac': gbs / pk['hbm'], 'bytes_per_step': nbytes,
            'ms_per_step': ms, 'launches_per_step': sum(v['n'] for _, v in rows),
            'timing': 'CUDA events per launch, eager pass (not under a profiler)'}


# ----------------------------------------------------------------------------------------------- secondary workloads
def run_c3(args, model, diffusion_factory, dev, rank, world):
    """BASELINE configs[2]: GQN-Mazes-shaped long videos (64x64, T = 300, 36 observed), `exp-past` with step_size 8
    (33 windows of 20 frames), DDIM-100, test videos sharded across the ranks (8 per GPU, task t -> rank t % world),
    one final NCCL gather of the uint8 samples.  A step = one whole `infer_video` job of this rank's batch."""
    import torch.distributed as dist
    from oracle import synth
    from video_diffusion_b200 import dist as vdist
    from video_diffusion_b200.sampling import infer_video, to_uint8
    diffusion = diffusion_factory('ddim100')
    T, obs, B = args.c3_frames, 36, B_PER_GPU
    video = synth.make_video((B, T, 3, SIZE, SIZE), seed=100 + rank)
    ids = torch.tensor(vdist.task_video_indices(rank, B, world * B), device=dev)
    n_windows = sum(1 for _ in __import__('video_diffusion_b200').inference_util.inference_strategies['exp-past'](
        video_length=T, num_obs=obs, max_frames=FRAMES, step_size=8))
    infer_video('exp-past', model, diffusion, video[:, :obs + 16], FRAMES, obs, 8, return_tensor=True)   # warm-up: 2 windows
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    samples = infer_video('exp-past', model, diffusion, video, FRAMES, obs, 8, return_tensor=True)
    e1.record()
    rows, all_ids = vdist.gather_ragged(to_uint8(samples), ids)
    e2.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e2), e1.elapsed_time(e2)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms, gather_ms = float(ms[0]), float(ms[1])
    assert rows.shape[0] == world * B and all_ids.tolist() == list(range(world * B))
    sampled = world * B * (T - obs)
    if rank == 0:
        print(json.dumps({
            'metric': 'sampled video frames/sec (full DDIM-100 chain, exp-past long video)', 'value': sampled / total_ms * 1e3,
            'unit': 'frames/s', 'n_gpus': world, 'steps': 1, 'warmup': 1, 'ms_per_step': total_ms,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': {'workload': f'C3: 64x64 FDM U-Net, T={T} video, 36 observed, exp-past step_size 8 '
                                   f'({n_windows} windows of {FRAMES} frames), DDIM-100, {B} videos per GPU, final '
                                   f'gather of uint8 samples over {"NCCL" if world > 1 else "one rank (no collective)"}',
                       'windows': n_windows, 'network_calls': n_windows * 100, 'videos': world * B,
                       'denoised_frames_per_s': world * B * FRAMES * n_windows * 100 / total_ms * 1e3},
            'gather_ms': gather_ms, 'gathered_bytes': int(rows.numel())}))


def run_c2chain(args, model, diffusion_factory, dev, rank, world):
    """BASELINE configs[1] as a CHAIN: `autoreg` with step_size 7, ancestral 1000-step sampling, batch 8 -- one full window
    (13 observed + 7 latent frames, 1000 `p_sample` steps through `infer_video`, finished frames copied out by the
    AsyncSampleWriter) is timed; a T = 500 video with 36 observed frames is 67 such windows ("time >= 1
    full window + extrapolate")."""
    import torch.distributed as dist
    from oracle import synth
    from video_diffusion_b200.sampling import AsyncSampleWriter, infer_video
    diffusion = diffusion_factory('')
    obs, step, B = 36, 7, B_PER_GPU
    T = obs + step * args.c2_windows
    video = synth.make_video((B, T, 3, SIZE, SIZE), seed=300 + rank)
    short = diffusion_factory('ddim10')
    infer_video('autoreg', model, short, video[:, :obs + step], FRAMES, obs, step, return_tensor=True)    # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    writer = AsyncSampleWriter(tuple(video.shape), dev)
    e0.record()
    infer_video('autoreg', model, diffusion, video, FRAMES, obs, step, return_tensor=True, writer=writer)
    frames_u8 = writer.finish()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    n_calls = args.c2_windows * diffusion.num_timesteps
    if rank == 0:
        per_window = ms / args.c2_windows
        print(json.dumps({
            'metric': 'sampled video frames/sec (full ancestral 1000-step chain, autoreg)', 'unit': 'frames/s',
            'value': world * B * step * args.c2_windows / ms * 1e3, 'n_gpus': world, 'steps': 1, 'warmup': 1,
            'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16',
            'data': 'synthetic',
            'config': {'workload': f'C2 chain: 64x64 FDM U-Net, autoreg step_size 7, max_frames 20 (13 observed + 7 latent), '
                                   f'{diffusion.num_timesteps} ancestral steps per window, {args.c2_windows} window(s) timed, '
                                   f'{B} videos per GPU, finished frames leave on a copy stream (AsyncSampleWriter)',
                       'network_calls': n_calls, 'denoised_frames_per_s': world * B * FRAMES * n_calls / ms * 1e3,
                       'ms_per_window': per_window,
                       'extrapolated_T500_video_s': per_window * 67 / 1e3,
                       'host_uint8_bytes': int(frames_u8.size)}}))


def run_c5(args, model, diffusion_factory, dev, rank, world):
    """BASELINE configs[4]: video_nll ELBO over all 1000 timesteps (forward-only sweep), 64x64, 20 frames, 8 videos per
    GPU; one final gather of the per-video ELBO rows.  A step = one whole `run_bpd_evaluation` call."""
    import torch.distributed as dist
    from oracle import synth
    from video_diffusion_b200 import dist as vdist
    from video_diffusion_b200.sampling import run_bpd_evaluation
    diffusion = diffusion_factory('')
    B = B_PER_GPU
    video = synth.make_video((B, 60, 3, SIZE, SIZE), seed=200 + rank)
    obs_idx = [list(range(N_OBS))] * B
    lat_idx = [list(range(N_OBS, FRAMES))] * B
    n_t = args.c5_timesteps
    t_seq = list(range(n_t))[::-1]
    run_bpd_evaluation(model, diffusion, video, True, obs_idx, lat_idx, t_seq=t_seq[:5])      # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = run_bpd_evaluation(model, diffusion, video, True, obs_idx, lat_idx, t_seq=t_seq)
    ids = torch.tensor(vdist.task_video_indices(rank, B, world * B), device=dev)
    rows, _ = vdist.gather_ragged(torch.from_numpy(out['total_bpd']).to(dev), ids)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    if rank == 0:
        print(json.dumps({
            'metric': 'ELBO evaluation: denoised video frames/sec (forward-only sweep over all timesteps)',
            'value': world * B * FRAMES * n_t / ms * 1e3, 'unit': 'frames/s', 'n_gpus': world, 'steps': 1, 'warmup': 1,
            'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16',
            'data': 'synthetic',
            'config': {'workload': f'C5: video_nll ELBO, calc_bpd_loop_subsampled over {n_t} timesteps, 64x64, 13 observed + '
                                   f'7 latent frames, {B} videos per GPU, final gather of per-video ELBO rows',
                       'network_calls': n_t, 'videos': world * B, 'videos_per_s': world * B / ms * 1e3,
                       'total_bpd_mean': float(rows.mean())}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-stock-gpu-baseline', action='store_true')
    ap.add_argument('--stock-gpu-baseline', action='store_true', help='(default on; kept for old command lines)')
    ap.add_argument('--profile-json', default=None, help='write the per-kernel-class breakdown here')
    ap.add_argument('--workload', default='c2', choices=['c2', 'c2chain', 'c4', 'c3', 'c5'],
                    help="c2 (default) is the configuration BASELINE.json's metric is quoted on; c4 = the 128x128 model; "
                         'c3 / c5 = the long-video sampling job and the ELBO sweep (secondary workloads)')
    ap.add_argument('--c3-frames', type=int, default=300)
    ap.add_argument('--c5-timesteps', type=int, default=1000)
    ap.add_argument('--c2-windows', type=int, default=1, help='c2chain: autoreg windows (1000 ancestral steps each) to time')
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

    def build(respacing):
        kw = video_model_and_diffusion_defaults()
        kw.update(cases.ref_config(CFG))
        kw['timestep_respacing'] = respacing
        return create_video_model_and_diffusion(compute_dtype=torch.bfloat16, **kw)

    model, diffusion = build('')
    sd = synth_state(CFG)
    model.load_state_dict(sd)
    model = model.to(dev).eval()
    if args.workload in ('c3', 'c5', 'c2chain'):
        with torch.no_grad():
            {'c3': run_c3, 'c5': run_c5, 'c2chain': run_c2chain}[args.workload](args, model, lambda r: build(r)[1], dev, rank,
                                                                                world)
        if world > 1:
            dist.destroy_process_group()
        return

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

    def step_e2e_serial():
        d = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
        x = x_pin.to(dev, non_blocking=True)
        s = diffusion.p_sample(model, x, t_dev, clip_denoised=True,
                               model_kwargs=dict(d, x_t_minus_1=d['x0'], observed_frames='x_0'))['sample']
        out_pin.copy_(s, non_blocking=True)
        return s

    # the public host-buffer API: uploads of step i + 1 and the download of step i ride a copy stream under the compute
    from video_diffusion_b200.sampling import PipelinedHostStepper
    stepper = PipelinedHostStepper(model, diffusion, dev)

    def step_e2e():
        return stepper.step(x_pin, t_dev, pinned, out_pin, clip_denoised=True)

    def e2e_finish():      # inside the timed region: the last download has landed in host memory
        stepper.drain()
        torch.cuda.current_stream().wait_stream(stepper.copy)

    def timed(fn, steps, finish=None):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if finish is not None:
            finish()
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
        # (the pipelined stepper hands its results to a copy stream -- record_stream -- so the caching allocator needs a
        # few more steps than the resident loop before it stops growing: at least 8 untimed steps for this leg)
        for _ in range(max(args.warmup, 8)):
            step_e2e()
        e2e_finish()
        ms_e2e = timed(step_e2e, args.steps, e2e_finish)
        for _ in range(2):
            step_e2e_serial()
        ms_e2e_serial = timed(step_e2e_serial, args.steps)

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
    achieved_gemm = alg_fl / (g_ms / 1e3) / 1e12 if g_ms else 0.0
    executed = g_fl / (g_ms / 1e3) / 1e12 if g_ms else 0.0
    total_prof_ms = sum(v['ms'] for v in prof.values())
    # the north-star figure: the whole forward's algorithmic FLOPs over the whole (graph-replayed) step
    step_tflop = FLOP_PER_FRAME * B * FRAMES / 1e12
    achieved_step = step_tflop / (ms / args.steps / 1e3)
    h2d = sum(v.numel() * v.element_size() for v in pinned.values()) + x_pin.numel() * 4
    result = {
        'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
        'config': shared_config(world),
        'clocks': clk.summary(),
        'e2e': {'value': e2e_value, 'unit': 'frames/s', 'h2d_bytes_per_step': int(h2d),
                'd2h_bytes_per_step': int(out_pin.numel() * 4), 'ms_per_step': ms_e2e / args.steps,
                'api': 'sampling.PipelinedHostStepper.step (diffusion.p_sample on pinned host tensors; transfers on a '
                       'copy stream under the neighbouring steps, all inside the timed region, last download drained)',
                'ms_per_step_serial': ms_e2e_serial / args.steps, 'warmup_steps': max(args.warmup, 8)},
        'gpu_launches': int(launches_per_step * args.steps),
        'roofline': {'bound': 'tensor',
                     'kernel': 'whole U-Net forward (SURVEY 8d algorithmic FLOPs) over the graph-replayed step; dominant '
                               'kernel class = the tcgen05 GEMM family (frac_gemm)',
                     'achieved': achieved_step, 'peak': pk['tflops'], 'unit': 'TFLOP/s',
                     'frac': achieved_step / pk['tflops'], 'frac_of_burst_peak': achieved_step / pk['burst'],
                     'target_frac': 0.60, 'traffic': None, 'peak_source': pk['src'],
                     'flops_per_step': step_tflop * 1e12,
                     'frac_gemm': achieved_gemm / pk['tflops'] if pk['tflops'] else None,
                     'gemm': {'achieved': achieved_gemm, 'launches_per_step': g_n, 'kernel_ms_per_step': g_ms,
                              'share_of_step': g_ms / total_prof_ms if total_prof_ms else None,
                              'flops_per_step': alg_fl, 'executed_flops_per_step': g_fl, 'executed_tflops': executed,
                              'timing': 'CUDA events per launch, eager pass with the launch queue kept full '
                                        '(not under a profiler)'}},
    }
    # the single heaviest launch shape of the step (3x3 conv 64x64, 256 -> 128, transposed-role halo kernel), per
    # launch: live CUDA-event time of this run; DRAM traffic from the committed `ncu --set full` capture named in
    # profiles/ncu_top_launch.json (regenerated per build; the file names the build and the report it came from)
    top = [(k, v) for k, v in prof_shapes.items() if k[0] == 'gemm_tc_conv3x3' and 'M=655360 N=128 K=2304' in k[1]]
    if top:
        (_, meta), v = top[0]
        t_launch = v['ms'] / v['n'] / 1e3
        tf = v['flops'] / t_launch / 1e12
        tl = {'kernel': 'gemm_tc_halo_t_kernel: conv3x3 64x64 256->128, ' + meta, 'us': t_launch * 1e6,
              'achieved': tf, 'frac': tf / pk['tflops'] if pk['tflops'] else None, 'flops': v['flops'],
              'algorithmic_bytes': 655360 * 256 * 2 + 655360 * 128 * 2}
        try:
            with open(os.path.join(ROOT, 'profiles', 'ncu_top_launch.json')) as f:
                cap = json.load(f)
            tl['traffic'] = cap['dram_bytes_read'] + cap['dram_bytes_write']
            tl['traffic_source'] = f"profiles/ncu_top_launch.json (build {cap.get('build')}, {cap.get('source')})"
            result['roofline']['traffic'] = tl['traffic']
        except Exception:
            tl['traffic'] = None
        result['roofline']['top_launch'] = tl
    hbm = {}
    gr = gn_apply_roofline(prof_shapes, pk)
    if gr:
        hbm['gn_apply'] = gr
    if rank == 0 and world == 1 and args.workload == 'c2':
        with torch.no_grad():
            hbm['sampler_step'] = sampler_spill_roofline(diffusion, dev, pk)
    result['roofline_hbm'] = hbm
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        step, kind, what = cpu_step(sd, B, threads)
        step()
        t0 = time.perf_counter()
        n = 2
        for _ in range(n):
            step()
        dt = (time.perf_counter() - t0) / n
        result['cpu_baseline'] = {'value': B * FRAMES / dt, 'unit': 'frames/s', 'cores': threads, 'kind': kind,
                                  'sample': f'{what}: U-Net forward + p_sample on one ({B},{FRAMES},3,{SIZE},{SIZE}) '
                                            f'batch of windows (the same batch as the GPU step), mean of {n} after 1 warm-up'}
    if rank == 0 and world == 1 and not args.no_stock_gpu_baseline:
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
