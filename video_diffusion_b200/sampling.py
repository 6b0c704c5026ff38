"""The two callers of the hot path, with the reference scripts' signatures:

`infer_video` (scripts/video_sample.py:50-190) -- window assembly from an inference strategy,
ancestral chain started from `x0.clone()`, write-back of the latent frames -- and
`run_bpd_evaluation` (scripts/video_nll.py:142-188) -- ragged obs/latent packing + ELBO loop.

B200-first differences (results unchanged): the video being generated stays resident in HBM for
the whole job (the reference bounces every window through host memory, :143-146 / :185), frames
are gathered / scattered by index on the device, and there is exactly one device->host copy at
the end.  `to_uint8` / `save_samples` reproduce the on-disk format (:266-272).
"""
import os

import numpy as np
import torch

from .inference_util import inference_strategies


def get_masks(x0, num_obs):
    """Observation / latent / kinda-marginal masks of a window whose first num_obs frames are observed."""
    obs_mask = torch.zeros_like(x0[:, :, :1, :1, :1])
    obs_mask[:, :num_obs] = 1
    return obs_mask, 1 - obs_mask, torch.zeros_like(obs_mask)


@torch.no_grad()
def infer_video(mode, model, diffusion, batch, max_frames, obs_length, step_size=1, optimal_schedule_path=None, *,
                use_gradient_method=False, observed_frames='x_0', device=None, return_tensor=False,
                save_all_timesteps=False, writer=None):
    """batch: (B, T, C, H, W) in [-1, 1].  Returns (samples, all_timestep_samples) as numpy arrays like
    scripts/video_sample.py:50-190; `observed_frames` and `save_all_timesteps` are script globals (`args.*`) there and
    keyword arguments here.  all_timestep_samples is (B, num_timesteps, T, C, H, W) in chain order with
    save_all_timesteps, zeros([1]) otherwise.  Everything stays on the device until the final copy; with `writer` (an
    AsyncSampleWriter) the finished frames of each window are converted and copied out on a side stream while the next
    window runs."""
    if use_gradient_method:
        raise NotImplementedError('use_gradient_method needs autograd through the network')
    if 'adaptive' in mode or 'goal-directed' in mode:
        raise NotImplementedError(f'inference mode {mode!r}')
    device = device or next(model.parameters()).device
    B, T = batch.shape[:2]
    video = batch.to(device, non_blocking=True).float()
    samples = torch.zeros_like(video)
    samples[:, :obs_length] = video[:, :obs_length]
    schedule = iter(inference_strategies[mode](video_length=T, num_obs=obs_length, max_frames=max_frames,
                                               step_size=step_size, optimal_schedule_path=optimal_schedule_path))
    steps = list(range(diffusion.num_timesteps))[::-1]
    all_steps = None
    if save_all_timesteps:                                      # scripts/video_sample.py:84-89
        all_steps = torch.zeros(B, len(steps), *video.shape[1:], device=device)
        all_steps[:, :, :obs_length] = samples[:, :obs_length].unsqueeze(1)
    t_all = torch.arange(diffusion.num_timesteps, device=device).view(-1, 1).expand(-1, B).contiguous()
    if writer is not None:
        writer.push(samples, range(obs_length))
    for obs_idx, lat_idx in schedule:
        idx = torch.tensor(list(obs_idx) + list(lat_idx), device=device, dtype=torch.long)
        x0 = samples.index_select(1, idx)
        frame_indices = idx.view(1, -1).repeat(B, 1)
        obs_mask, latent_mask, kinda_marg_mask = get_masks(x0, len(obs_idx))
        kwargs = dict(frame_indices=frame_indices, x0=x0, obs_mask=obs_mask, latent_mask=latent_mask,
                      kinda_marg_mask=kinda_marg_mask, x_t_minus_1=x0, observed_frames=observed_frames)
        local = x0.clone()
        n_lat = len(lat_idx)
        for k, step in enumerate(steps):
            local = diffusion.p_sample(model, local, t=t_all[step], clip_denoised=True, model_kwargs=kwargs,
                                       return_attn_weights=False)['sample']
            if all_steps is not None:
                all_steps[:, k, idx[-n_lat:]] = local[:, -n_lat:]
        samples[:, idx[-n_lat:]] = local[:, -n_lat:]
        if writer is not None:
            writer.push(samples, lat_idx)
    if return_tensor:
        return samples if all_steps is None else (samples, all_steps)
    return samples.cpu().numpy(), (np.zeros([1], dtype=np.float32) if all_steps is None else all_steps.cpu().numpy())


@torch.no_grad()
def infer_video_full(mode, model, diffusion, batch, max_frames, obs_length, step_size=1, optimal_schedule_path=None, *,
                     vertical_steps=0, observed_frames='x_0', use_gradient_method=False, device=None,
                     return_tensor=False, save_all_timesteps=False):
    """The "vertical / horizontal" schedule of scripts/video_sample_full.py:50-323, an adjacent caller of the same
    p_sample: the first `vertical_steps` timesteps run window by window like infer_video (:88-203, each window a
    partial chain from x0.clone() with observed_frames='x_0'); every remaining timestep then sweeps ALL windows once
    (:205-313): a window's input is whatever the video holds at that moment -- its latent frames are the partly
    denoised state written back by the previous sweep -- one p_sample step is taken with `observed_frames` (the
    script's args.observed_frames) and the latent frames are written back.  With vertical_steps == 0 the latent frames
    start at zero, as in the script.  `vertical_steps`, `observed_frames`, `save_all_timesteps` are `args.*` globals
    there.  The video stays in HBM; one device->host copy at the end."""
    if use_gradient_method:
        raise NotImplementedError('use_gradient_method needs autograd through the network')
    if 'adaptive' in mode or 'goal-directed' in mode:
        raise NotImplementedError(f'inference mode {mode!r}')
    device = device or next(model.parameters()).device
    B, T = batch.shape[:2]
    video = batch.to(device, non_blocking=True).float()
    samples = torch.zeros_like(video)
    samples[:, :obs_length] = video[:, :obs_length]
    n_t = diffusion.num_timesteps
    steps = list(range(n_t))[::-1]
    all_steps = None
    if save_all_timesteps:                                      # scripts/video_sample_full.py:79-86
        all_steps = torch.zeros(B, n_t, *video.shape[1:], device=device)
        all_steps[:, :, :obs_length] = samples[:, :obs_length].unsqueeze(1)
    t_all = torch.arange(n_t, device=device).view(-1, 1).expand(-1, B).contiguous()

    def windows():
        return iter(inference_strategies[mode](video_length=T, num_obs=obs_length, max_frames=max_frames,
                                               step_size=step_size, optimal_schedule_path=optimal_schedule_path))

    def window(obs_idx, lat_idx, observed):
        idx = torch.tensor(list(obs_idx) + list(lat_idx), device=device, dtype=torch.long)
        x0 = samples.index_select(1, idx)
        obs_mask, latent_mask, kinda_marg_mask = get_masks(x0, len(obs_idx))
        return idx, x0, dict(frame_indices=idx.view(1, -1).repeat(B, 1), x0=x0, obs_mask=obs_mask,
                             latent_mask=latent_mask, kinda_marg_mask=kinda_marg_mask, x_t_minus_1=x0,
                             observed_frames=observed)

    if vertical_steps > 0:
        for obs_idx, lat_idx in windows():
            idx, x0, kwargs = window(obs_idx, lat_idx, 'x_0')
            local, n_lat = x0.clone(), len(lat_idx)
            for k, step in enumerate(steps[:vertical_steps]):
                local = diffusion.p_sample(model, local, t=t_all[step], clip_denoised=True, model_kwargs=kwargs,
                                           return_attn_weights=False)['sample']
                if all_steps is not None:
                    all_steps[:, k, idx[-n_lat:]] = local[:, -n_lat:]
            samples[:, idx[-n_lat:]] = local[:, -n_lat:]
    for k, step in enumerate(steps[vertical_steps:]):
        for obs_idx, lat_idx in windows():
            idx, x0, kwargs = window(obs_idx, lat_idx, observed_frames)
            n_lat = len(lat_idx)
            local = diffusion.p_sample(model, x0, t=t_all[step], clip_denoised=True, model_kwargs=kwargs,
                                       return_attn_weights=False)['sample']
            samples[:, idx[-n_lat:]] = local[:, -n_lat:]
        if all_steps is not None:
            all_steps[:, vertical_steps + k] = samples
    if return_tensor:
        return samples if all_steps is None else (samples, all_steps)
    return samples.cpu().numpy(), (np.zeros([1], dtype=np.float32) if all_steps is None else all_steps.cpu().numpy())


class AsyncSampleWriter:
    """Device->host copy and .npy write of finished frames, overlapped with the next window (the reference moves every
    window's result to the host synchronously, scripts/video_sample.py:179-189, and writes `sample_XXXX-k.npy` after
    the whole video, :266-272).  `push` is called by infer_video after each window: on a copy stream the frames are
    converted to uint8 exactly like `to_uint8` and copied into one pinned host buffer while the compute stream goes on
    with the next window; `finish` waits for the copies and writes the same files as `save_samples`."""

    def __init__(self, shape, device):
        B, T, C, H, W = shape
        self.host = torch.empty((B, T, C, H, W), dtype=torch.uint8).pin_memory()
        self.stream = torch.cuda.Stream(device=device)
        self.pushed = torch.zeros(T, dtype=torch.bool)

    def push(self, samples, frames):
        """samples: the (B, T, C, H, W) video on the device; frames: list of frame indices that are final now."""
        frames = [int(f) for f in frames]
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(samples.device))
        self.stream.wait_event(ready)
        with torch.cuda.stream(self.stream):
            idx = torch.tensor(frames, device=samples.device, dtype=torch.long)
            chunk = to_uint8(samples.index_select(1, idx))
            for j, f in enumerate(frames):
                self.host[:, f].copy_(chunk[:, j], non_blocking=True)
            chunk.record_stream(self.stream)
        # the compute stream must not overwrite these frames before the copy stream has read them: infer_video never
        # rewrites a finished frame, which is what makes the overlap safe
        self.pushed[frames] = True

    def finish(self, out_dir=None, dataset_indices=None, sample_idx=0):
        self.stream.synchronize()
        if not bool(self.pushed.all()):
            raise RuntimeError(f'frames never pushed: {(~self.pushed).nonzero().flatten().tolist()}')
        arr = self.host.numpy()
        if out_dir is None:
            return arr
        return save_samples(out_dir, arr, dataset_indices, sample_idx)


class PipelinedHostStepper:
    """`diffusion.p_sample` for callers whose inputs and results live in (pinned) HOST memory: the inputs of step i + 1 are
    uploaded on a copy stream while step i computes, and the sample of step i is downloaded while step i + 1 computes,
    so the PCIe transfers (15.7 MB up + 7.9 MB down per C2 step) disappear behind the 12-13 ms of compute instead of
    adding to them.  Uploads and downloads have a stream each (`up_stream`, `copy`): on ONE stream the upload of step
    i + 1 queues behind the download of step i, which waits for step i's compute -- the upload then starts only after
    the step it was meant to hide under.  Two device-side input slots; a slot is rewritten only after the step that
    read it has been enqueued past its input copy."""

    def __init__(self, model, diffusion, device):
        self.model, self.diffusion, self.device = model, diffusion, torch.device(device)
        self.copy = torch.cuda.Stream(device=self.device)           # downloads
        # uploads (VDM_STEPPER_ONE_STREAM=1: share the download stream, for A/B runs)
        self.up_stream = self.copy if os.environ.get('VDM_STEPPER_ONE_STREAM') == '1' else torch.cuda.Stream(device=self.device)
        self.slots = [None, None]
        self.up = [torch.cuda.Event(), torch.cuda.Event()]
        self.used = [None, None]
        self.i = 0

    def _upload(self, slot, x_pin, kw_pin):
        bufs = self.slots[slot]
        if bufs is None:
            bufs = self.slots[slot] = dict({k: torch.empty(v.shape, dtype=v.dtype, device=self.device)
                                            for k, v in kw_pin.items()}, __x=torch.empty(x_pin.shape, dtype=x_pin.dtype,
                                                                                        device=self.device))
        with torch.cuda.stream(self.up_stream):
            if self.used[slot] is not None:
                self.up_stream.wait_event(self.used[slot])
            bufs['__x'].copy_(x_pin, non_blocking=True)
            for k, v in kw_pin.items():
                bufs[k].copy_(v, non_blocking=True)
            self.up[slot].record(self.up_stream)
        return bufs

    def step(self, x_pin, t, kw_pin, out_pin, **p_sample_kwargs):
        """x_pin, kw_pin (x0, masks, frame_indices ...): pinned host tensors; t: device tensor; out_pin: pinned host tensor
        that receives the sample (valid after `drain()` or once a later step's download has completed)."""
        slot = self.i & 1
        self.i += 1
        bufs = self._upload(slot, x_pin, kw_pin)
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self.up[slot])
        kw = {k: v for k, v in bufs.items() if k != '__x'}
        s = self.diffusion.p_sample(self.model, bufs['__x'], t, model_kwargs=dict(kw, x_t_minus_1=kw['x0'],
                                                                                observed_frames='x_0'),
                                    **p_sample_kwargs)['sample']
        done = torch.cuda.Event()
        done.record(cur)
        self.used[slot] = done
        with torch.cuda.stream(self.copy):
            self.copy.wait_event(done)
            out_pin.copy_(s, non_blocking=True)
            s.record_stream(self.copy)
        return s

    def drain(self):
        self.up_stream.synchronize()
        self.copy.synchronize()


def to_uint8(samples):
    """(x+1)/2*255 clipped to uint8, as written to sample_XXXX-k.npy (scripts/video_sample.py:266-268)."""
    if torch.is_tensor(samples):
        return ((samples + 1) * 127.5).clamp(0, 255).to(torch.uint8)
    return ((np.asarray(samples) + 1) * 127.5).clip(0, 255).astype(np.uint8)


def save_samples(out_dir, samples_uint8, dataset_indices, sample_idx=0):
    """One `sample_{idx:04d}-{k}.npy` of shape (T, 3, H, W) per video (scripts/video_sample.py:211-272)."""
    os.makedirs(out_dir, exist_ok=True)
    paths = []
    for vid, idx in zip(samples_uint8, dataset_indices):
        path = os.path.join(out_dir, f'sample_{int(idx):04d}-{sample_idx}.npy')
        np.save(path, np.asarray(vid))
        paths.append(path)
    return paths


@torch.no_grad()
def run_bpd_evaluation(model, diffusion, batch, clip_denoised, obs_indices, lat_indices, t_seq=None, device=None):
    """ELBO of `lat_indices` frames given `obs_indices` frames, per video; index lists may be ragged
    across the batch (padded rows are masked out of attention and of the loss)."""
    device = device or next(model.parameters()).device
    batch = batch.to(device).float()
    B = batch.shape[0]
    max_frames = max(len(o) + len(l) for o, l in zip(obs_indices, lat_indices))
    x0 = torch.zeros_like(batch[:, :max_frames])
    obs_mask = torch.zeros_like(x0[:, :, :1, :1, :1])
    lat_mask = torch.zeros_like(obs_mask)
    frame_indices = torch.zeros(B, max_frames, device=device, dtype=torch.long)
    for i, (o, l) in enumerate(zip(obs_indices, lat_indices)):
        sel = torch.tensor(list(o) + list(l), device=device, dtype=torch.long)
        n = sel.numel()
        x0[i, :n] = batch[i].index_select(0, sel)
        obs_mask[i, :len(o)] = 1.0
        lat_mask[i, len(o):n] = 1.0
        frame_indices[i, :n] = sel
    # the reference script omits these two kwargs and crashes on them
    model_kwargs = dict(frame_indices=frame_indices, x0=x0, obs_mask=obs_mask, latent_mask=lat_mask,
                        kinda_marg_mask=torch.zeros_like(obs_mask), x_t_minus_1=x0, observed_frames='x_0')
    metrics = diffusion.calc_bpd_loop_subsampled(model, x0, clip_denoised=clip_denoised, model_kwargs=model_kwargs,
                                                 latent_mask=lat_mask, t_seq=t_seq)
    metrics = {k: (v.sum(dim=1) if v.ndim > 1 else v) * max_frames for k, v in metrics.items()}
    return {k: v.detach().cpu().numpy() for k, v in metrics.items()}
