"""Generate tests/golden/* by running the UNMODIFIED reference (imported from
/root/reference) on the CPU.  Run in the build container only:

    python oracle/make_golden.py

The reference cannot travel to the GPU box, so its outputs are committed as small
fixtures; every input is regenerated from seeds by oracle/synth.py, so fixtures
hold outputs only.  TEST INFRASTRUCTURE ONLY.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, 'tests', 'golden')
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')

# `inference_util` imports lpips at module scope (inference_util.py:3); stub it.
_lp = types.ModuleType('lpips')
_lp.LPIPS = type('LPIPS', (), {'__init__': lambda s, *a, **k: None})
_lp.normalize_tensor = lambda x: x
sys.modules['lpips'] = _lp

from improved_diffusion import gaussian_diffusion as gd                      # noqa: E402
from improved_diffusion import inference_util                                 # noqa: E402
from improved_diffusion.script_util import (create_gaussian_diffusion,       # noqa: E402
                                            create_video_model_and_diffusion,
                                            video_model_and_diffusion_defaults)

from oracle import synth                                                      # noqa: E402
from oracle.cases import (CHAIN_CASE, DDIM50_CASE, DIFFUSION_CASES, FULL_CASES, STRATEGY_GRID, UNCOND_CASE,  # noqa: E402
                          UNET_CASES, UNET_LUT_CASES, UNET_VARIANT_CASES, bpd_case_inputs, fake_eps, full_case_inputs,
                          model_kwargs_for, ref_config, uncond_case_inputs, unet_case_inputs, variant_kwargs)


class NoiseReplay:
    """Replace th.randn_like inside the reference by a seeded, regenerable stream."""

    def __init__(self, base_seed):
        self.i, self.base = 0, base_seed

    def __call__(self, like):
        z = synth.make_noise(tuple(like.shape), seed=self.base + self.i)
        self.i += 1
        return z


def build(cfg_name, respacing='', device='cpu'):
    kw = video_model_and_diffusion_defaults()
    kw.update(ref_config(cfg_name))
    kw['timestep_respacing'] = respacing
    with torch.device(device):
        model, diffusion = create_video_model_and_diffusion(**kw)
    return model.eval(), diffusion


def dump_specs(names=('tiny', 'tiny_nrb2', 'c2', 'c4')):
    for name in names:
        model, _ = build(name, device='meta')
        spec = {k: list(v.shape) for k, v in model.state_dict().items()}
        with open(os.path.join(GOLD, f'spec_{name}.json'), 'w') as f:
            json.dump(spec, f, indent=0, sort_keys=True)
        print('spec', name, len(spec), sum(int(np.prod(s)) for s in spec.values()))


def dump_strategies():
    out = []
    for mode, T, obs, mf, step in STRATEGY_GRID:
        it = inference_util.inference_strategies[mode](video_length=T, num_obs=obs, max_frames=mf, step_size=step)
        steps = [[[int(i) for i in o], [int(i) for i in l]] for o, l in it]
        out.append(dict(mode=mode, T=T, obs=obs, max_frames=mf, step_size=step, steps=steps))
    with open(os.path.join(GOLD, 'frame_indices.json'), 'w') as f:
        json.dump(out, f)
    print('strategies', len(out))


def load_ref_model(cfg_name, respacing=''):
    model, diffusion = build(cfg_name, respacing)
    spec = json.load(open(os.path.join(GOLD, f'spec_{cfg_name}.json')))
    model.load_state_dict(synth.make_state_dict(spec, seed=1))
    return model, diffusion


def dump_unet(cases_list=UNET_CASES, fname='unet.npz'):
    arrays = {}
    for case in cases_list:
        model, _ = load_ref_model(case['cfg'])
        inp = full_case_inputs(case) if fname == 'unet_full.npz' else unet_case_inputs(case)
        taps = {}
        hooks = []
        for grp in ('input_blocks', 'output_blocks'):
            for i, seq in enumerate(getattr(model, grp)):
                for j, mod in enumerate(seq):
                    hooks.append(mod.register_forward_hook(
                        lambda m, a, o, key=f'{grp}.{i}.{j}': taps.__setitem__(key, o)))
        for j, mod in enumerate(model.middle_block):
            hooks.append(mod.register_forward_hook(lambda m, a, o, key=f'middle_block.{j}': taps.__setitem__(key, o)))
        hooks.append(model.time_embed.register_forward_hook(lambda m, a, o: taps.__setitem__('emb', o)))
        kw = variant_kwargs(case, inp) if fname == 'unet_variants.npz' else model_kwargs_for(inp)
        with torch.no_grad():
            out, _ = model(inp['x'], timesteps=inp['t_model'].clone(), **kw)
        for h in hooks:
            h.remove()
        arrays[f"{case['name']}/eps"] = out.numpy()
        for k, v in taps.items():
            arrays[f"{case['name']}/tap/{k}"] = synth.fingerprint(v)
        print('unet', case['name'], float(out.abs().max()), float(out.std()))
    np.savez_compressed(os.path.join(GOLD, fname), **arrays)


def dump_diffusion():
    arrays = {}
    for case in DIFFUSION_CASES:
        name, resp = case['name'], case['respacing']
        d = create_gaussian_diffusion(steps=1000, noise_schedule=case['schedule'], timestep_respacing=resp,
                                      rescale_timesteps=True, rescale_learned_sigmas=True)
        for attr in ('betas', 'alphas_cumprod', 'alphas_cumprod_prev', 'sqrt_alphas_cumprod',
                     'sqrt_one_minus_alphas_cumprod', 'log_one_minus_alphas_cumprod', 'sqrt_recip_alphas_cumprod',
                     'sqrt_recipm1_alphas_cumprod', 'posterior_variance', 'posterior_log_variance_clipped',
                     'posterior_mean_coef1', 'posterior_mean_coef2'):
            arrays[f'{name}/{attr}'] = getattr(d, attr)
        arrays[f'{name}/timestep_map'] = np.array(d.timestep_map, dtype=np.int64)
        shape = case['shape']
        x = synth.make_noise(shape, seed=11)
        x0 = synth.make_video(shape, seed=12)
        noise = synth.make_noise(shape, seed=13)
        lat = torch.zeros(shape[0], shape[1], 1, 1, 1)
        lat[:, shape[1] // 2:] = 1
        model = lambda xx, timesteps, **kw: (fake_eps(xx, timesteps), None)
        for tag, t in case['ts'].items():
            t = torch.tensor(t)
            gd.th.randn_like = lambda like: noise
            ps = d.p_sample(model, x, t, clip_denoised=True, model_kwargs={})
            arrays[f'{name}/{tag}/p_sample'] = ps['sample'].numpy()
            arrays[f'{name}/{tag}/pred_xstart'] = ps['pred_xstart'].numpy()
            pm = d.p_mean_variance(model, x, t, clip_denoised=False, model_kwargs={})
            arrays[f'{name}/{tag}/mean_noclip'] = pm['mean'].numpy()
            arrays[f'{name}/{tag}/log_variance'] = pm['log_variance'].numpy()
            for eta in (0.0, 0.7):
                ds = d.ddim_sample(model, x, t, clip_denoised=True, model_kwargs={}, eta=eta)
                arrays[f'{name}/{tag}/ddim_eta{eta}'] = ds['sample'].numpy()
            arrays[f'{name}/{tag}/q_sample'] = d.q_sample(x0, t, noise=noise).numpy()
            xt = d.q_sample(x0, t, noise=noise)
            vb = d._vb_terms_bpd(model, x_start=x0, x_t=xt, t=t, clip_denoised=True, model_kwargs={}, latent_mask=lat)
            arrays[f'{name}/{tag}/vb'] = vb['output'].numpy()
        gd.th.randn_like = NoiseReplay(2000)
        bpd = d.calc_bpd_loop_subsampled(model, x0, clip_denoised=True, model_kwargs={}, latent_mask=lat,
                                         t_seq=case['t_seq'])
        for k, v in bpd.items():
            arrays[f'{name}/bpd/{k}'] = v.numpy()
        gd.th.randn_like = torch.randn_like
    np.savez_compressed(os.path.join(GOLD, 'diffusion.npz'), **arrays)
    print('diffusion', len(arrays))


def dump_chain():
    """Script-style sampling (scripts/video_sample.py:50-190), ddim_sample_loop and the
    ELBO loop (scripts/video_nll.py:142-188) for the tiny model, with replayed noise."""
    c = CHAIN_CASE
    arrays = {}
    model, diffusion = load_ref_model(c['cfg'], c['respacing'])
    B, T = c['batch'], c['video_length']
    video = synth.make_video((B, T, 3, c['image_size'], c['image_size']), seed=c['video_seed'])
    samples = torch.zeros_like(video)
    samples[:, :c['obs_length']] = video[:, :c['obs_length']]
    it = inference_util.inference_strategies[c['mode']](video_length=T, num_obs=c['obs_length'],
                                                         max_frames=c['max_frames'], step_size=c['step_size'])
    gd.th.randn_like = NoiseReplay(c['noise_seed'])
    with torch.no_grad():
        for obs, lat in it:
            x0 = torch.cat([samples[:, obs], samples[:, lat]], dim=1).clone()
            fi = torch.tensor(list(obs) + list(lat)).repeat((B, 1))
            om = torch.zeros_like(x0[:, :, :1, :1, :1])
            om[:, :len(obs)] = 1
            kw = dict(frame_indices=fi, x0=x0, obs_mask=om, latent_mask=1 - om,
                      kinda_marg_mask=torch.zeros_like(om), x_t_minus_1=x0, observed_frames='x_0')
            cur = x0.clone()
            for step in reversed(range(diffusion.num_timesteps)):
                cur = diffusion.p_sample(model, cur, t=torch.tensor([step] * B), clip_denoised=True,
                                         model_kwargs=kw)['sample']
            samples[:, lat] = cur[:, -len(lat):]
    arrays['chain/samples'] = samples.numpy()
    print('chain', float(samples.abs().max()), float(samples.std()))

    # ddim_sample_loop on the first window (gaussian_diffusion.py:670-748)
    it = inference_util.inference_strategies[c['mode']](video_length=T, num_obs=c['obs_length'],
                                                         max_frames=c['max_frames'], step_size=c['step_size'])
    obs, lat = next(iter(it))
    x0 = torch.cat([video[:, obs], torch.zeros_like(video[:, lat])], dim=1)
    fi = torch.tensor(list(obs) + list(lat)).repeat((B, 1))
    om = torch.zeros_like(x0[:, :, :1, :1, :1])
    om[:, :len(obs)] = 1
    kw = dict(frame_indices=fi, x0=x0, obs_mask=om, latent_mask=1 - om, kinda_marg_mask=torch.zeros_like(om),
              x_t_minus_1=x0, observed_frames='x_0')
    gd.th.randn_like = NoiseReplay(c['noise_seed'] + 500)
    init = synth.make_noise(tuple(x0.shape), seed=c['noise_seed'] + 499)
    with torch.no_grad():
        out = diffusion.ddim_sample_loop(model, tuple(x0.shape), noise=init, clip_denoised=True, model_kwargs=kw,
                                         device='cpu')
    arrays['ddim_loop/sample'] = out.numpy()

    # ELBO with ragged index lists (video_nll.py:142-188)
    model, diffusion = load_ref_model(c['cfg'], c['bpd_respacing'])
    inp = bpd_case_inputs(c)
    gd.th.randn_like = NoiseReplay(c['noise_seed'] + 900)
    kw = dict(frame_indices=inp['frame_indices'], x0=inp['x0'], obs_mask=inp['obs_mask'],
              latent_mask=inp['latent_mask'], kinda_marg_mask=inp['kinda_marg_mask'],
              x_t_minus_1=inp['x0'], observed_frames='x_0')
    metrics = diffusion.calc_bpd_loop_subsampled(model, inp['x0'], clip_denoised=True, model_kwargs=kw,
                                                 latent_mask=inp['latent_mask'], t_seq=None)
    for k, v in metrics.items():
        arrays[f'bpd/{k}'] = v.numpy()
    print('bpd total', metrics['total_bpd'])
    gd.th.randn_like = torch.randn_like
    np.savez_compressed(os.path.join(GOLD, 'chain.npz'), **arrays)


def dump_full_schedule():
    """scripts/video_sample_full.py:50-323 (vertical, then horizontal diffusion) on the CHAIN_CASE video with the
    reference's model and p_sample, replayed noise; the script itself reads `args` / `logger` globals, so its loop is
    followed statement by statement here (observed_frames='x_0', non-adaptive mode)."""
    from oracle.cases import FULL_SCHEDULE_CASE
    c, f = CHAIN_CASE, FULL_SCHEDULE_CASE
    model, diffusion = load_ref_model(c['cfg'], c['respacing'])
    B, T = c['batch'], c['video_length']
    video = synth.make_video((B, T, 3, c['image_size'], c['image_size']), seed=c['video_seed'])
    samples = torch.zeros_like(video)
    samples[:, :c['obs_length']] = video[:, :c['obs_length']]

    def windows():
        return inference_util.inference_strategies[c['mode']](video_length=T, num_obs=c['obs_length'],
                                                              max_frames=c['max_frames'], step_size=c['step_size'])

    def kwargs(x0, obs, lat):
        fi = torch.tensor(list(obs) + list(lat)).repeat((B, 1))
        om = torch.zeros_like(x0[:, :, :1, :1, :1])
        om[:, :len(obs)] = 1
        return dict(frame_indices=fi, x0=x0, obs_mask=om, latent_mask=1 - om, kinda_marg_mask=torch.zeros_like(om),
                    x_t_minus_1=x0, observed_frames='x_0')

    steps = list(range(diffusion.num_timesteps))[::-1]
    gd.th.randn_like = NoiseReplay(f['noise_seed'])
    with torch.no_grad():
        for obs, lat in windows():                                           # vertical diffusion (:88-203)
            x0 = torch.cat([samples[:, obs], samples[:, lat]], dim=1).clone()
            cur = x0.clone()
            for step in steps[:f['vertical_steps']]:
                cur = diffusion.p_sample(model, cur, t=torch.tensor([step] * B), clip_denoised=True,
                                         model_kwargs=kwargs(x0, obs, lat))['sample']
            samples[:, lat] = cur[:, -len(lat):]
        vertical = samples.clone()
        for step in steps[f['vertical_steps']:]:                             # horizontal diffusion (:205-313)
            for obs, lat in windows():
                x0 = torch.cat([samples[:, obs], samples[:, lat]], dim=1).clone()
                cur = diffusion.p_sample(model, x0, t=torch.tensor([step] * B), clip_denoised=True,
                                         model_kwargs=kwargs(x0, obs, lat))['sample']
                samples[:, lat] = cur[:, -len(lat):]
    gd.th.randn_like = torch.randn_like
    np.savez_compressed(os.path.join(GOLD, 'chain_full.npz'), **{'full/vertical': vertical.numpy(),
                                                                 'full/samples': samples.numpy()})
    print('full schedule', float(samples.abs().max()), float(samples.std()))


def dump_probe():
    """calc_bpd_loop_subsampled with a 2-D t_seq, one row of timesteps per video (gaussian_diffusion.py:960-969),
    the way scripts/video_optimal_schedule.py:97-105 calls it; stand-in network, replayed noise."""
    from oracle.cases import PROBE_T_SEQ
    case = DIFFUSION_CASES[0]
    d = create_gaussian_diffusion(steps=1000, noise_schedule=case['schedule'], timestep_respacing=case['respacing'],
                                  rescale_timesteps=True, rescale_learned_sigmas=True)
    shape = case['shape']
    x0 = synth.make_video(shape, seed=12)
    lat = torch.zeros(shape[0], shape[1], 1, 1, 1)
    lat[:, shape[1] // 2:] = 1
    model = lambda xx, timesteps, **kw: (fake_eps(xx, timesteps), None)
    gd.th.randn_like = NoiseReplay(2500)
    bpd = d.calc_bpd_loop_subsampled(model, x0, clip_denoised=True, model_kwargs={}, latent_mask=lat,
                                     t_seq=np.array(PROBE_T_SEQ))
    gd.th.randn_like = torch.randn_like
    np.savez_compressed(os.path.join(GOLD, 'probe.npz'), **{f'probe/{k}': v.numpy() for k, v in bpd.items()})
    print('probe', bpd['vb'])


def dump_attn():
    """return_attn_weights=True (unet.py:464-468, 797-801; gaussian_diffusion.py:496-524): per-layer head-averaged
    attention maps of one ragged forward, and the per-quartile averages p_sample_loop returns.  The maps are large
    (B*F x 256 x 256), so the fixtures keep synth.fingerprint signatures (moments + strided samples)."""
    arrays = {}
    case = UNET_CASES[1]
    model, _ = load_ref_model(case['cfg'])
    inp = unet_case_inputs(case)
    with torch.no_grad():
        out, attns = model(inp['x'], timesteps=inp['t_model'], return_attn_weights=True, **model_kwargs_for(inp))
    arrays['fwd/eps'] = synth.fingerprint(out, 64)
    for key, layers in attns.items():
        for i, a in enumerate(layers):
            arrays[f'fwd/{key}/{i}'] = synth.fingerprint(a, 256)
            arrays[f'fwd/{key}/{i}/shape'] = np.array(a.shape)
    print('attn fwd', {k: len(v) for k, v in attns.items()})

    c = CHAIN_CASE
    model, diffusion = load_ref_model(c['cfg'], c['respacing'])
    B, T = c['batch'], c['video_length']
    video = synth.make_video((B, T, 3, c['image_size'], c['image_size']), seed=c['video_seed'])
    it = inference_util.inference_strategies[c['mode']](video_length=T, num_obs=c['obs_length'],
                                                         max_frames=c['max_frames'], step_size=c['step_size'])
    obs, lat = next(iter(it))
    x0 = torch.cat([video[:, obs], torch.zeros_like(video[:, lat])], dim=1)
    fi = torch.tensor(list(obs) + list(lat)).repeat((B, 1))
    om = torch.zeros_like(x0[:, :, :1, :1, :1])
    om[:, :len(obs)] = 1
    real_cuda, real_to = torch.Tensor.cuda, torch.Tensor.to
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.Tensor.to = lambda self, *a, **k: self if a and a[0] == 'cuda' else real_to(self, *a, **k)
    try:
        kw = dict(frame_indices=fi, x0=x0, obs_mask=om, latent_mask=1 - om, kinda_marg_mask=torch.zeros_like(om),
                  observed_frames='x_0')
        gd.th.randn_like = NoiseReplay(c['noise_seed'] + 700)
        init = synth.make_noise(tuple(x0.shape), seed=c['noise_seed'] + 699)
        with torch.no_grad():
            out, attns = diffusion.p_sample_loop(model, tuple(x0.shape), noise=init, clip_denoised=True,
                                                 model_kwargs=kw, device='cpu', return_attn_weights=True)
    finally:
        torch.Tensor.cuda, torch.Tensor.to = real_cuda, real_to
        gd.th.randn_like = torch.randn_like
    for tag, a in attns.items():
        arrays['loop/' + tag] = synth.fingerprint(a, 512)
        arrays['loop/' + tag + '/shape'] = np.array(a.shape)
    print('attn loop', sorted(attns))
    np.savez_compressed(os.path.join(GOLD, 'attn.npz'), **arrays)


def first_window_kwargs(c):
    B, T = c['batch'], c['video_length']
    video = synth.make_video((B, T, 3, c['image_size'], c['image_size']), seed=c['video_seed'])
    it = inference_util.inference_strategies[c['mode']](video_length=T, num_obs=c['obs_length'],
                                                         max_frames=c['max_frames'], step_size=c['step_size'])
    obs, lat = next(iter(it))
    x0 = torch.cat([video[:, obs], torch.zeros_like(video[:, lat])], dim=1)
    fi = torch.tensor(list(obs) + list(lat)).repeat((B, 1))
    om = torch.zeros_like(x0[:, :, :1, :1, :1])
    om[:, :len(obs)] = 1
    return x0, dict(frame_indices=fi, x0=x0, obs_mask=om, latent_mask=1 - om, kinda_marg_mask=torch.zeros_like(om),
                    x_t_minus_1=x0, observed_frames='x_0')


def dump_ddim50():
    """north_star's sampled-frames gate: a fixed-seed DDIM-50 `ddim_sample_loop` (gaussian_diffusion.py:670-748) on the
    first window of the chain case, eta 0 and eta 1 (the stochastic branch draws replayed noise every step)."""
    c, d = CHAIN_CASE, DDIM50_CASE
    model, diffusion = load_ref_model(c['cfg'], d['respacing'])
    assert diffusion.num_timesteps == 50
    x0, kw = first_window_kwargs(c)
    arrays = {}
    for eta in (0.0, 1.0):
        gd.th.randn_like = NoiseReplay(d['noise_seed'] + 1)
        init = synth.make_noise(tuple(x0.shape), seed=d['noise_seed'])
        with torch.no_grad():
            out = diffusion.ddim_sample_loop(model, tuple(x0.shape), noise=init, clip_denoised=True, model_kwargs=kw,
                                             device='cpu', eta=eta)
        arrays[f'ddim50/eta{eta}'] = out.numpy()
        print('ddim50 eta', eta, float(out.abs().max()), float(out.std()))
    gd.th.randn_like = torch.randn_like
    np.savez_compressed(os.path.join(GOLD, 'ddim50.npz'), **arrays)


def dump_uncond():
    """UNetVideoModel.forward (unet.py:898-912): the class is built directly (script_util.py:287-300 passes
    cond_emb_type, which its constructor rejects), the state-dict spec goes to spec_tiny_uncond.json."""
    from improved_diffusion.unet import UNetVideoModel
    case = UNCOND_CASE
    cfg = ref_config(case['cfg'])
    model = UNetVideoModel(T=cfg['T'], use_frame_encoding=False, cross_frame_attention=True,
                           enforce_position_invariance=False, in_channels=3, model_channels=cfg['num_channels'],
                           out_channels=3, num_res_blocks=cfg['num_res_blocks'], attention_resolutions=(2, 4),
                           channel_mult=(1, 2, 2, 2), num_heads=4, use_scale_shift_norm=True, use_spatial_encoding=True,
                           image_size=cfg['image_size'], temporal_augment_type='add_manyhead_presoftmax_time',
                           use_rpe_net=True, bucket_params=dict(alpha=cfg['T'], beta=cfg['T'], gamma=cfg['T']),
                           allow_interactions_between_padding=True).eval()
    spec = {k: list(v.shape) for k, v in model.state_dict().items()}
    with open(os.path.join(GOLD, 'spec_tiny_uncond.json'), 'w') as f:
        json.dump(spec, f, indent=0, sort_keys=True)
    model.load_state_dict(synth.make_state_dict(spec, seed=1))
    inp = uncond_case_inputs(case)
    with torch.no_grad():
        out, attn = model(inp['x'], inp['timesteps'], frame_indices=inp['frame_indices'], attn_mask=inp['attn_mask'])
    assert attn is None
    print('uncond', float(out.abs().max()), float(out.std()))
    np.savez_compressed(os.path.join(GOLD, 'unet_uncond.npz'), **{f"{case['name']}/eps": out.numpy()})


def dump_diffusion_extra():
    """The posterior / prediction helpers and the DDIM reverse ODE (gaussian_diffusion.py:171-188, 208-227, 374-396,
    636-668) for the stand-in network: the closed-form methods the first fixture set did not cover."""
    arrays = {}
    for case in DIFFUSION_CASES:
        name = case['name']
        d = create_gaussian_diffusion(steps=1000, noise_schedule=case['schedule'], timestep_respacing=case['respacing'],
                                      rescale_timesteps=True, rescale_learned_sigmas=True)
        shape = case['shape']
        x, x0, noise = synth.make_noise(shape, seed=11), synth.make_video(shape, seed=12), synth.make_noise(shape, seed=13)
        model = lambda xx, timesteps, **kw: (fake_eps(xx, timesteps), None)
        for tag, t in case['ts'].items():
            t = torch.tensor(t)
            k = f'{name}/{tag}/'
            m, v, lv = d.q_mean_variance(x0, t)
            arrays[k + 'q_mean'], arrays[k + 'q_var'], arrays[k + 'q_logvar'] = m.numpy(), v.numpy(), lv.numpy()
            m, v, lv = d.q_posterior_mean_variance(x0, x, t)
            arrays[k + 'post_mean'], arrays[k + 'post_var'], arrays[k + 'post_logvar'] = m.numpy(), v.numpy(), lv.numpy()
            arrays[k + 'xstart_from_eps'] = d._predict_xstart_from_eps(x, t, noise).numpy()
            arrays[k + 'xstart_from_xprev'] = d._predict_xstart_from_xprev(x, t, noise).numpy()
            arrays[k + 'eps_from_xstart'] = d._predict_eps_from_xstart(x, t, x0).numpy()
            rv = d.ddim_reverse_sample(model, x, t, clip_denoised=True, model_kwargs={})
            arrays[k + 'ddim_reverse'] = rv['sample'].numpy()
    np.savez_compressed(os.path.join(GOLD, 'diffusion_extra.npz'), **arrays)
    print('diffusion_extra', len(arrays))


P_LOOP_MODES = ('x_0', 'x_t_minus_1', 'hybrid_5')


def dump_p_sample_loop():
    """p_sample_loop (gaussian_diffusion.py:450-595) on the first window of the chain case, one run per
    observed_frames mode.  The reference hard-codes `.cuda()` / `.to('cuda')` on two tiny tensors in this loop
    (:567-573); for this CPU run those two calls are mapped to no-ops, nothing else is touched."""
    c = CHAIN_CASE
    model, diffusion = load_ref_model(c['cfg'], c['respacing'])
    B, T = c['batch'], c['video_length']
    video = synth.make_video((B, T, 3, c['image_size'], c['image_size']), seed=c['video_seed'])
    it = inference_util.inference_strategies[c['mode']](video_length=T, num_obs=c['obs_length'],
                                                         max_frames=c['max_frames'], step_size=c['step_size'])
    obs, lat = next(iter(it))
    x0 = torch.cat([video[:, obs], torch.zeros_like(video[:, lat])], dim=1)
    fi = torch.tensor(list(obs) + list(lat)).repeat((B, 1))
    om = torch.zeros_like(x0[:, :, :1, :1, :1])
    om[:, :len(obs)] = 1
    real_cuda, real_to = torch.Tensor.cuda, torch.Tensor.to
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.Tensor.to = lambda self, *a, **k: self if a and a[0] == 'cuda' else real_to(self, *a, **k)
    arrays = {}
    try:
        for mode in P_LOOP_MODES:
            kw = dict(frame_indices=fi, x0=x0, obs_mask=om, latent_mask=1 - om, kinda_marg_mask=torch.zeros_like(om),
                      observed_frames=mode)
            gd.th.randn_like = NoiseReplay(c['noise_seed'] + 700)
            init = synth.make_noise(tuple(x0.shape), seed=c['noise_seed'] + 699)
            with torch.no_grad():
                out, _ = diffusion.p_sample_loop(model, tuple(x0.shape), noise=init, clip_denoised=True,
                                                 model_kwargs=kw, device='cpu')
            arrays[f'p_loop/{mode}'] = out.numpy()
            print('p_sample_loop', mode, float(out.abs().max()), float(out.std()))
    finally:
        torch.Tensor.cuda, torch.Tensor.to = real_cuda, real_to
        gd.th.randn_like = torch.randn_like
    np.savez_compressed(os.path.join(GOLD, 'p_loop.npz'), **arrays)


if __name__ == '__main__':
    torch.manual_seed(0)
    os.makedirs(GOLD, exist_ok=True)
    if sys.argv[1:] == ['lut']:       # only the lookup-table RPE fixtures (leaves the other files untouched)
        dump_specs(('tiny_lut',))
        dump_unet(UNET_LUT_CASES, 'unet_lut.npz')
        sys.exit(0)
    if sys.argv[1:] == ['round2']:    # full-size C2 / C4 forwards, unconditioned entry point, DDIM-50 chain, helpers
        dump_unet(FULL_CASES, 'unet_full.npz')
        dump_uncond()
        dump_ddim50()
        dump_diffusion_extra()
        sys.exit(0)
    if sys.argv[1:] == ['full_schedule']:
        dump_full_schedule()
        sys.exit(0)
    if sys.argv[1:] == ['attn']:
        dump_attn()
        sys.exit(0)
    if sys.argv[1:] == ['ploop']:
        dump_p_sample_loop()
        dump_probe()
        sys.exit(0)
    if sys.argv[1:] == ['variants']:  # cond_emb_type / frame-encoding / observed_frames variants only
        dump_specs(('tiny_dup', 'tiny_t0', 'tiny_fe', 'tiny_fei'))
        dump_unet(UNET_VARIANT_CASES, 'unet_variants.npz')
        sys.exit(0)
    dump_specs()
    dump_strategies()
    dump_diffusion()
    dump_unet()
    dump_chain()
