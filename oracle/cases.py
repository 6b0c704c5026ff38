"""Named parity cases: configs + seeded inputs, shared by the golden generator and the
tests (so fixtures hold reference OUTPUTS only).  TEST INFRASTRUCTURE ONLY."""
import torch

from . import synth

# name -> kwargs overriding video_model_and_diffusion_defaults() (script_util.py:40-57).
# rp_alpha/beta/gamma must be set even with use_rpe_net=True.
_CFG = {
    'tiny':      dict(image_size=32, num_channels=64, num_res_blocks=1, T=30),
    'tiny_nrb2': dict(image_size=32, num_channels=64, num_res_blocks=2, T=30),
    'c2':        dict(image_size=64, num_channels=128, num_res_blocks=2, T=500),
    'c4':        dict(image_size=128, num_channels=128, num_res_blocks=2, T=1000),
    # use_rpe_net=False: bucketed lookup-table RPE; alpha < beta < gamma (gamma off the integer grid) exercises all
    # three pieces of the index function
    'tiny_lut':  dict(image_size=32, num_channels=64, num_res_blocks=1, T=30, use_rpe_net=False, rp_alpha=3, rp_beta=7,
                      rp_gamma=20.5),
    # model variants of SURVEY 8(f)-2: other conditioning embeddings, frame-index sinusoid
    'tiny_dup':  dict(image_size=32, num_channels=64, num_res_blocks=1, T=30, cond_emb_type='duplicate'),
    'tiny_t0':   dict(image_size=32, num_channels=64, num_res_blocks=1, T=30, cond_emb_type='t=0'),
    'tiny_fe':   dict(image_size=32, num_channels=64, num_res_blocks=1, T=30, use_frame_encoding=True),
    'tiny_fei':  dict(image_size=32, num_channels=64, num_res_blocks=1, T=30, use_frame_encoding=True,
                      enforce_position_invariance=True),
}


# the unconditioned entry point (UNetVideoModel.forward, unet.py:898-912): built directly, because the reference's
# create_video_model(do_cond_marg=False) passes cond_emb_type to a constructor that does not take it
_CFG['tiny_uncond'] = dict(image_size=32, num_channels=64, num_res_blocks=1, T=30, do_cond_marg=False)


def ref_config(name):
    c = dict(_CFG[name])
    for k in ('rp_alpha', 'rp_beta', 'rp_gamma'):
        c.setdefault(k, c['T'])
    return c


# (mode, video_length, obs_length, max_frames, step_size)
STRATEGY_GRID = [
    ('autoreg', 30, 5, 10, 5), ('autoreg', 500, 36, 20, 7), ('autoreg', 300, 36, 20, 10), ('autoreg', 33, 0, 10, 4),
    ('independent', 30, 5, 10, 5), ('independent', 300, 36, 20, 7), ('independent', 100, 10, 20, 10),
    ('exp-past', 300, 36, 20, 8), ('exp-past', 100, 4, 10, 4), ('exp-past', 500, 36, 20, 4),
    ('really-independent', 50, 5, 10, 3), ('mixed-autoreg-independent', 300, 36, 20, 7),
    ('hierarchy-2', 300, 36, 20, 10), ('hierarchy-2', 100, 10, 10, 5), ('hierarchy-3', 300, 36, 20, 10),
    ('hierarchy-2', 64, 0, 16, 8),
]

UNET_CASES = [
    # sampling-style window: observed prefix, everything else latent
    dict(name='tiny_window', cfg='tiny', B=1, F=10, n_obs=[5], n_lat=[5], t=[500],
         frame_indices=[[0, 1, 2, 3, 4, 5, 6, 7, 8, 9]]),
    # video_nll-style ragged rows with padding frames (neither observed nor latent)
    dict(name='tiny_ragged', cfg='tiny', B=2, F=10, n_obs=[3, 6], n_lat=[5, 4], t=[17, 803],
         frame_indices=[[0, 1, 2, 10, 11, 12, 13, 14, 0, 0], [3, 7, 9, 11, 12, 13, 20, 21, 22, 23]]),
    dict(name='tiny_nrb2_window', cfg='tiny_nrb2', B=1, F=7, n_obs=[2], n_lat=[5], t=[999],
         frame_indices=[[4, 29, 10, 11, 12, 13, 14]]),
]

# lookup-table RPE (stored in tests/golden/unet_lut.npz)
UNET_LUT_CASES = [
    dict(name='tiny_lut_ragged', cfg='tiny_lut', B=2, F=10, n_obs=[3, 6], n_lat=[5, 4], t=[17, 803],
         frame_indices=[[0, 1, 2, 10, 11, 12, 13, 14, 0, 0], [3, 7, 9, 11, 12, 13, 20, 25, 28, 29]]),
]


# cond_emb_type / use_frame_encoding / observed_frames variants (stored in tests/golden/unet_variants.npz); `observed`
# selects kwargs['observed_frames'], x_t_minus_1 and hybrid are seeded tensors (variant_kwargs)
_RAGGED = dict(B=2, F=10, n_obs=[3, 6], n_lat=[5, 4], t=[17, 803],
               frame_indices=[[0, 1, 2, 10, 11, 12, 13, 14, 0, 0], [3, 7, 9, 11, 12, 13, 20, 25, 28, 29]])
UNET_VARIANT_CASES = [
    dict(name='dup_ragged', cfg='tiny_dup', **_RAGGED),
    dict(name='t0_ragged', cfg='tiny_t0', **_RAGGED),
    dict(name='t0_no_obs', cfg='tiny_t0', B=2, F=6, n_obs=[0, 2], n_lat=[6, 4], t=[250, 600],
         frame_indices=[[0, 1, 2, 3, 4, 5], [9, 8, 7, 6, 5, 4]]),
    dict(name='fe_ragged', cfg='tiny_fe', **_RAGGED),
    dict(name='fei_ragged', cfg='tiny_fei', **_RAGGED),
    dict(name='obs_xt', cfg='tiny', observed='x_t', **_RAGGED),
    dict(name='obs_xtm1', cfg='tiny', observed='x_t_minus_1', **_RAGGED),
    dict(name='obs_hybrid500', cfg='tiny', observed='hybrid_500', **_RAGGED),
]


# the benchmarked architectures at their real sizes, one video each: the reference itself runs these on the CPU in
# seconds, so its eps and per-block fingerprints pin the full-size kernels (tests/golden/unet_full.npz)
FULL_CASES = [
    dict(name='c2_full', cfg='c2', B=1, F=20, n_obs=[7], n_lat=[12], t=[411.0]),     # one trailing padding frame
    dict(name='c4_full', cfg='c4', B=1, F=4, n_obs=[1], n_lat=[3], t=[37.0]),
]


def full_case_inputs(case):
    B, F = case['B'], case['F']
    size = _CFG[case['cfg']]['image_size']
    x0 = synth.make_video((B, F, 3, size, size), seed=41)
    x = synth.make_noise((B, F, 3, size, size), seed=42)
    obs = torch.zeros(B, F, 1, 1, 1)
    lat = torch.zeros(B, F, 1, 1, 1)
    for b in range(B):
        obs[b, :case['n_obs'][b]] = 1
        lat[b, case['n_obs'][b]:case['n_obs'][b] + case['n_lat'][b]] = 1
    fi = torch.stack([torch.randperm(200, generator=torch.Generator().manual_seed(7 + b))[:F] for b in range(B)])
    return dict(x=x, x0=x0, obs_mask=obs, latent_mask=lat, kinda_marg_mask=torch.zeros_like(obs), frame_indices=fi,
                t_model=torch.tensor(case['t'], dtype=torch.float32))


# UNetVideoModel.forward: per-frame timesteps (B, F), an attention mask with padding frames, no conditioning
UNCOND_CASE = dict(name='uncond_ragged', cfg='tiny_uncond', B=2, F=8, n_valid=[8, 6],
                   frame_indices=[[0, 1, 2, 3, 4, 5, 6, 7], [3, 7, 9, 11, 12, 13, 20, 25]])


def uncond_case_inputs(case):
    B, F = case['B'], case['F']
    size = _CFG[case['cfg']]['image_size']
    x = synth.make_noise((B, F, 3, size, size), seed=52)
    t = torch.from_numpy(__import__('numpy').random.RandomState(53).randint(0, 1000, size=(B, F))).float()
    mask = torch.zeros(B, F, 1, 1, 1)
    for b in range(B):
        mask[b, :case['n_valid'][b]] = 1
    return dict(x=x, timesteps=t, attn_mask=mask, frame_indices=torch.tensor(case['frame_indices'], dtype=torch.long))


# fixed-seed DDIM-50 chain (north_star's PSNR gate): tiny model, first window of CHAIN_CASE, replayed noise
DDIM50_CASE = dict(respacing='ddim50', noise_seed=7000)


def variant_kwargs(case, inp):
    """model kwargs of a variant case: the reference's names, seeded x_t_minus_1 / hybrid tensors."""
    kw = model_kwargs_for(inp)
    kw['observed_frames'] = case.get('observed', 'x_0')
    kw['x_t_minus_1'] = synth.make_noise(tuple(inp['x'].shape), seed=23)
    kw['hybrid'] = synth.make_video(tuple(inp['x'].shape), seed=24)
    return kw


def unet_case_inputs(case):
    B, F = case['B'], case['F']
    size = _CFG[case['cfg']]['image_size']
    x0 = synth.make_video((B, F, 3, size, size), seed=21)
    x = synth.make_noise((B, F, 3, size, size), seed=22)
    obs = torch.zeros(B, F, 1, 1, 1)
    lat = torch.zeros(B, F, 1, 1, 1)
    for b in range(B):
        obs[b, :case['n_obs'][b]] = 1
        lat[b, case['n_obs'][b]:case['n_obs'][b] + case['n_lat'][b]] = 1
    return dict(x=x, x0=x0, obs_mask=obs, latent_mask=lat, kinda_marg_mask=torch.zeros_like(obs),
                frame_indices=torch.tensor(case['frame_indices'], dtype=torch.long),
                t_model=torch.tensor(case['t'], dtype=torch.float32))


def model_kwargs_for(inp):
    return dict(x0=inp['x0'], obs_mask=inp['obs_mask'], latent_mask=inp['latent_mask'],
                kinda_marg_mask=inp['kinda_marg_mask'], frame_indices=inp['frame_indices'],
                x_t_minus_1=inp['x0'], observed_frames='x_0')


def fake_eps(x, t):
    """Stand-in network for the sampler-maths fixtures: depends on x and on the (rescaled) t."""
    return 0.5 * torch.tanh(x) + t.float().view(-1, *([1] * (x.dim() - 1))) / 2000.0


DIFFUSION_CASES = [
    dict(name='full1000', schedule='linear', respacing='', shape=(2, 4, 3, 8, 8),
         ts=dict(t0=[0, 0], tmid=[1, 517], tend=[999, 998]), t_seq=[999, 500, 1, 0]),
    dict(name='ddim10', schedule='linear', respacing='ddim10', shape=(2, 4, 3, 8, 8),
         ts=dict(t0=[0, 0], tmid=[3, 7], tend=[9, 9]), t_seq=None),
    dict(name='ddim50', schedule='linear', respacing='ddim50', shape=(3, 2, 3, 4, 4),
         ts=dict(tmid=[0, 25, 49]), t_seq=[49, 10, 0]),
    dict(name='sections', schedule='cosine', respacing='10,15,20', shape=(2, 2, 3, 4, 4),
         ts=dict(tmid=[0, 44]), t_seq=[44, 20, 0]),
]

# video_optimal_schedule.py:77-207 probes the ELBO terms at one random timestep PER VIDEO: a 2-D t_seq (rows = videos)
PROBE_T_SEQ = [[517, 3], [0, 999]]

CHAIN_CASE = dict(cfg='tiny', image_size=32, respacing='ddim10', bpd_respacing='ddim4', batch=1, video_length=30,
                  obs_length=5, max_frames=10, step_size=5, mode='independent', video_seed=31, noise_seed=5000,
                  bpd_obs=[[0, 1, 2], [3, 7, 9, 11, 12, 13]], bpd_lat=[[10, 11, 12, 13, 14], [20, 21, 22, 23]])


# scripts/video_sample_full.py schedule on the CHAIN_CASE video: 4 vertical timesteps, then 6 horizontal sweeps
FULL_SCHEDULE_CASE = dict(vertical_steps=4, noise_seed=8000)


def bpd_case_inputs(c):
    """run_bpd_evaluation's packing (scripts/video_nll.py:149-164) for ragged index lists."""
    obs_l, lat_l = c['bpd_obs'], c['bpd_lat']
    B = len(obs_l)
    video = synth.make_video((B, c['video_length'], 3, c['image_size'], c['image_size']), seed=c['video_seed'] + 1)
    mf = max(len(o) + len(l) for o, l in zip(obs_l, lat_l))
    x0 = torch.zeros(B, mf, 3, c['image_size'], c['image_size'])
    obs = torch.zeros(B, mf, 1, 1, 1)
    lat = torch.zeros(B, mf, 1, 1, 1)
    fi = torch.zeros(B, mf, dtype=torch.long)
    for i, (o, l) in enumerate(zip(obs_l, lat_l)):
        x0[i, :len(o)] = video[i, o]
        obs[i, :len(o)] = 1
        fi[i, :len(o)] = torch.tensor(o)
        x0[i, len(o):len(o) + len(l)] = video[i, l]
        lat[i, len(o):len(o) + len(l)] = 1
        fi[i, len(o):len(o) + len(l)] = torch.tensor(l)
    return dict(x0=x0, obs_mask=obs, latent_mask=lat, kinda_marg_mask=torch.zeros_like(obs), frame_indices=fi,
                max_frames=mf)
