"""CPU oracle for the two callers of the hot path.  TEST INFRASTRUCTURE ONLY.

Restates `infer_video` (scripts/video_sample.py:50-190: window assembly, chain started
from x0.clone() -- SURVEY Q5 --, ancestral p_sample per step, write-back of the latent
frames) and `run_bpd_evaluation` (scripts/video_nll.py:142-188) on top of the oracle
U-Net and the oracle sampler maths, plus the vertical / horizontal schedule of
scripts/video_sample_full.py:50-323.  Pinned by tests/golden/chain.npz and chain_full.npz.
"""
import torch

from . import diffusion_oracle as D
from . import strategies_oracle as S
from . import unet_oracle as U


def infer_video(sd, cfg, sched, video, mode, max_frames, obs_length, step_size, noise_fn):
    """video (B,T,3,H,W); noise_fn(shape) returns the next randn draw (one per p_sample)."""
    B, T = video.shape[:2]
    samples = torch.zeros_like(video)
    samples[:, :obs_length] = video[:, :obs_length]
    for obs, lat in S.schedule(mode, T, obs_length, max_frames, step_size):
        x0 = torch.cat([samples[:, obs], samples[:, lat]], dim=1).clone()
        fi, om, lm, km = (torch.from_numpy(a) for a in S.window_tensors(obs, lat, B))
        cur = x0.clone()
        for step in reversed(range(sched.num_timesteps)):
            t = torch.full((B,), step, dtype=torch.long)
            eps = U.cond_marg_forward(sd, cfg, cur, x0, om, lm, km, sched.model_time(t), fi)
            cur = D.p_sample(sched, eps, cur, t, noise_fn(tuple(cur.shape)))['sample']
        samples[:, lat] = cur[:, -len(lat):]
    return samples


def infer_video_full(sd, cfg, sched, video, mode, max_frames, obs_length, step_size, noise_fn, vertical_steps):
    """scripts/video_sample_full.py:50-323 with observed_frames='x_0': `vertical_steps` timesteps window by window
    (:88-203), then one sweep over all windows per remaining timestep (:205-313), each window restarting from the
    video's current contents."""
    B, T = video.shape[:2]
    samples = torch.zeros_like(video)
    samples[:, :obs_length] = video[:, :obs_length]
    steps = list(reversed(range(sched.num_timesteps)))

    def one(cur, x0, obs, lat, step):
        fi, om, lm, km = (torch.from_numpy(a) for a in S.window_tensors(obs, lat, B))
        t = torch.full((B,), step, dtype=torch.long)
        eps = U.cond_marg_forward(sd, cfg, cur, x0, om, lm, km, sched.model_time(t), fi)
        return D.p_sample(sched, eps, cur, t, noise_fn(tuple(cur.shape)))['sample']

    if vertical_steps > 0:
        for obs, lat in S.schedule(mode, T, obs_length, max_frames, step_size):
            x0 = torch.cat([samples[:, obs], samples[:, lat]], dim=1).clone()
            cur = x0.clone()
            for step in steps[:vertical_steps]:
                cur = one(cur, x0, obs, lat, step)
            samples[:, lat] = cur[:, -len(lat):]
    for step in steps[vertical_steps:]:
        for obs, lat in S.schedule(mode, T, obs_length, max_frames, step_size):
            x0 = torch.cat([samples[:, obs], samples[:, lat]], dim=1).clone()
            samples[:, lat] = one(x0, x0, obs, lat, step)[:, -len(lat):]
    return samples


def ddim_sample_loop(sd, cfg, sched, init, kw, noise_fn, eta=0.0):
    """gaussian_diffusion.py:670-748 for one window; kw holds x0/masks/frame_indices."""
    img = init
    B = init.shape[0]
    for step in reversed(range(sched.num_timesteps)):
        t = torch.full((B,), step, dtype=torch.long)
        eps = U.cond_marg_forward(sd, cfg, img, kw['x0'], kw['obs_mask'], kw['latent_mask'], kw['kinda_marg_mask'],
                                  sched.model_time(t), kw['frame_indices'])
        img = D.ddim_sample(sched, eps, img, t, noise_fn(tuple(img.shape)), eta=eta)['sample']
    return img


def run_bpd_evaluation(sd, cfg, sched, packed, noise_fn, t_seq=None):
    """packed: dict from oracle.cases.bpd_case_inputs.  Returns the raw calc_bpd_loop dict and the
    per-video metrics (sum over t, x max_frames) the script writes."""
    def eps_fn(x_t, t):
        return U.cond_marg_forward(sd, cfg, x_t, packed['x0'], packed['obs_mask'], packed['latent_mask'],
                                   packed['kinda_marg_mask'], sched.model_time(t), packed['frame_indices'])
    n = len(t_seq) if t_seq is not None else sched.num_timesteps
    noises = [noise_fn(tuple(packed['x0'].shape)) for _ in range(n)]
    raw = D.calc_bpd_loop(sched, eps_fn, packed['x0'], packed['latent_mask'], noises, t_seq)
    summed = {k: (v.sum(dim=1) if v.dim() > 1 else v) * packed['max_frames'] for k, v in raw.items()}
    return raw, {k: v.numpy() for k, v in summed.items()}
