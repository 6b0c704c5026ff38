"""The CPU oracle (oracle/) against fixtures produced by the unmodified reference
(oracle/make_golden.py).  No GPU, no product code."""
import numpy as np
import pytest
import torch

from conftest import max_rel
from oracle import cases, diffusion_oracle as D, strategies_oracle as S, synth, unet_oracle as U


def test_frame_index_schedules_bit_exact(golden):
    for rec in golden.json('frame_indices'):
        got = [[o, l] for o, l in S.schedule(rec['mode'], rec['T'], rec['obs'], rec['max_frames'], rec['step_size'])]
        assert got == rec['steps'], rec['mode']


@pytest.mark.parametrize('case', cases.DIFFUSION_CASES, ids=lambda c: c['name'])
def test_schedule_tables_and_sampler_math(golden, case):
    g = golden.npz('diffusion')
    n = case['name']
    s = D.Schedule(1000, case['schedule'], case['respacing'])
    pairs = dict(betas=s.betas, alphas_cumprod=s.acp, alphas_cumprod_prev=s.acp_prev, sqrt_alphas_cumprod=s.sqrt_acp,
                 sqrt_one_minus_alphas_cumprod=s.sqrt_1m_acp, log_one_minus_alphas_cumprod=s.log_1m_acp,
                 sqrt_recip_alphas_cumprod=s.sqrt_recip_acp, sqrt_recipm1_alphas_cumprod=s.sqrt_recipm1_acp,
                 posterior_variance=s.post_var, posterior_log_variance_clipped=s.post_logvar,
                 posterior_mean_coef1=s.post_c1, posterior_mean_coef2=s.post_c2)
    for k, v in pairs.items():
        np.testing.assert_array_equal(v, g[f'{n}/{k}'], err_msg=k)          # float64, bit-exact
    assert s.timestep_map == g[f'{n}/timestep_map'].tolist()
    shape = case['shape']
    x, x0, noise = synth.make_noise(shape, 11), synth.make_video(shape, 12), synth.make_noise(shape, 13)
    lat = torch.zeros(shape[0], shape[1], 1, 1, 1)
    lat[:, shape[1] // 2:] = 1
    for tag, tl in case['ts'].items():
        t = torch.tensor(tl)
        eps = cases.fake_eps(x, s.model_time(t))
        ps = D.p_sample(s, eps, x, t, noise)
        np.testing.assert_allclose(ps['sample'].numpy(), g[f'{n}/{tag}/p_sample'], rtol=0, atol=1e-6)
        np.testing.assert_allclose(ps['pred_xstart'].numpy(), g[f'{n}/{tag}/pred_xstart'], rtol=0, atol=1e-6)
        pm = D.p_mean_variance(s, eps, x, t, clip_denoised=False)
        np.testing.assert_allclose(pm['mean'].numpy(), g[f'{n}/{tag}/mean_noclip'], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(pm['log_variance'].numpy(), g[f'{n}/{tag}/log_variance'], rtol=0, atol=0)
        for eta in (0.0, 0.7):
            ds = D.ddim_sample(s, eps, x, t, noise, eta=eta)
            np.testing.assert_allclose(ds['sample'].numpy(), g[f'{n}/{tag}/ddim_eta{eta}'], rtol=0, atol=2e-6)
        xt = D.q_sample(s, x0, t, noise)
        np.testing.assert_allclose(xt.numpy(), g[f'{n}/{tag}/q_sample'], rtol=0, atol=1e-6)
        vb = D.vb_terms(s, cases.fake_eps(xt, s.model_time(t)), x0, xt, t, lat)
        np.testing.assert_allclose(vb['output'].numpy(), g[f'{n}/{tag}/vb'], rtol=1e-5, atol=1e-6)
    t_seq = case['t_seq'] if case['t_seq'] is not None else list(range(s.num_timesteps))[::-1]
    noises = [synth.make_noise(shape, 2000 + i) for i in range(len(t_seq))]
    bpd = D.calc_bpd_loop(s, lambda xt, t: cases.fake_eps(xt, s.model_time(t)), x0, lat, noises, t_seq)
    for k, v in bpd.items():
        np.testing.assert_allclose(v.numpy(), g[f'{n}/bpd/{k}'], rtol=1e-5, atol=1e-6, err_msg=k)


def test_per_video_timestep_probe_matches_reference(golden):
    """2-D t_seq (one row of timesteps per video), the call scripts/video_optimal_schedule.py:97-105 makes."""
    case = cases.DIFFUSION_CASES[0]
    g = golden.npz('probe')
    s = D.Schedule(1000, case['schedule'], case['respacing'])
    shape = case['shape']
    x0 = synth.make_video(shape, 12)
    lat = torch.zeros(shape[0], shape[1], 1, 1, 1)
    lat[:, shape[1] // 2:] = 1
    t_seq = np.array(cases.PROBE_T_SEQ)
    noises = [synth.make_noise(shape, 2500 + i) for i in range(t_seq.shape[1])]
    bpd = D.calc_bpd_loop(s, lambda xt, t: cases.fake_eps(xt, s.model_time(t)), x0, lat, noises, t_seq)
    for k, v in bpd.items():
        np.testing.assert_allclose(v.numpy(), g[f'probe/{k}'], rtol=1e-5, atol=1e-6, err_msg=k)


def test_attention_map_logging_matches_reference(golden):
    """return_attn_weights=True: per-layer | head-mean | attention maps (unet.py:464-468)."""
    g = golden.npz('attn')
    case = cases.UNET_CASES[1]
    sd = synth.make_state_dict(golden.json('spec_' + case['cfg']), seed=1)
    cfg = U.model_config(**cases.ref_config(case['cfg']))
    inp = cases.unet_case_inputs(case)
    log = {'spatial': [], 'temporal': [], 'mixed': []}
    with torch.no_grad():
        U.cond_marg_forward(sd, cfg, inp['x'], inp['x0'], inp['obs_mask'], inp['latent_mask'], inp['kinda_marg_mask'],
                            inp['t_model'], inp['frame_indices'], attn_log=log)
    for key in ('spatial', 'temporal'):
        assert len(log[key]) == len([k for k in g.files if k.startswith(f'fwd/{key}/') and k.endswith('/shape')])
        for i, a in enumerate(log[key]):
            assert list(a.shape) == g[f'fwd/{key}/{i}/shape'].tolist()
            np.testing.assert_allclose(synth.fingerprint(a, 256), g[f'fwd/{key}/{i}'], rtol=1e-4, atol=1e-6)


def golden_file(case):
    return ('unet_lut' if case in cases.UNET_LUT_CASES else
            'unet_variants' if case in cases.UNET_VARIANT_CASES else 'unet')


def oracle_kwargs(case, inp):
    """Extra oracle arguments of a variant case (observed_frames / x_t_minus_1 / hybrid)."""
    if case not in cases.UNET_VARIANT_CASES:
        return {}
    kw = cases.variant_kwargs(case, inp)
    return dict(observed_frames=kw['observed_frames'], x_t_minus_1=kw['x_t_minus_1'], hybrid=kw['hybrid'])


@pytest.mark.parametrize('case', cases.UNET_CASES + cases.UNET_LUT_CASES + cases.UNET_VARIANT_CASES,
                         ids=lambda c: c['name'])
def test_unet_forward_matches_reference(golden, case):
    g = golden.npz(golden_file(case))
    spec = golden.json('spec_' + case['cfg'])
    sd = synth.make_state_dict(spec, seed=1)
    cfg = U.model_config(**cases.ref_config(case['cfg']))
    inp = cases.unet_case_inputs(case)
    taps = {}
    with torch.no_grad():
        out = U.cond_marg_forward(sd, cfg, inp['x'], inp['x0'], inp['obs_mask'], inp['latent_mask'],
                                  inp['kinda_marg_mask'], inp['t_model'], inp['frame_indices'], taps=taps,
                                  **oracle_kwargs(case, inp))
    assert max_rel(out.numpy(), g[f"{case['name']}/eps"]) < 2e-5
    checked = 0
    for key, val in taps.items():
        gk = f"{case['name']}/tap/" + ('emb' if key == 'emb' else key.rsplit('.', 1)[0])
        assert gk in g.files, gk
        np.testing.assert_allclose(synth.fingerprint(val), g[gk], rtol=2e-4, atol=2e-4, err_msg=key)
        checked += 1
    assert checked == len([k for k in g.files if k.startswith(case['name'] + '/tap/')])


@pytest.mark.parametrize('case', cases.FULL_CASES, ids=lambda c: c['name'])
def test_full_size_forward_matches_reference(golden, case):
    """The benchmarked 64x64 (C2) and 128x128 (C4) architectures at full size: oracle vs the reference's own eps and
    per-block fingerprints (tests/golden/unet_full.npz) -- pins the oracle where the GPU parity tests use it."""
    g = golden.npz('unet_full')
    sd = synth.make_state_dict(golden.json('spec_' + case['cfg']), seed=1)
    cfg = U.model_config(**cases.ref_config(case['cfg']))
    inp = cases.full_case_inputs(case)
    taps = {}
    with torch.no_grad():
        out = U.cond_marg_forward(sd, cfg, inp['x'], inp['x0'], inp['obs_mask'], inp['latent_mask'],
                                  inp['kinda_marg_mask'], inp['t_model'], inp['frame_indices'], taps=taps)
    assert max_rel(out.numpy(), g[f"{case['name']}/eps"]) < 2e-5
    for key, val in taps.items():
        gk = f"{case['name']}/tap/" + ('emb' if key == 'emb' else key.rsplit('.', 1)[0])
        np.testing.assert_allclose(synth.fingerprint(val), g[gk], rtol=2e-4, atol=2e-4, err_msg=key)


def test_unconditioned_video_forward_matches_reference(golden):
    """UNetVideoModel.forward (unet.py:898-912): per-frame timesteps, attention mask with padding frames."""
    case = cases.UNCOND_CASE
    g = golden.npz('unet_uncond')
    sd = synth.make_state_dict(golden.json('spec_' + case['cfg']), seed=1)
    cfg = U.model_config(**cases.ref_config(case['cfg']))
    inp = cases.uncond_case_inputs(case)
    with torch.no_grad():
        out = U.video_forward(sd, cfg, inp['x'], inp['timesteps'], inp['frame_indices'], inp['attn_mask'])
    assert max_rel(out.numpy(), g[f"{case['name']}/eps"]) < 2e-5


@pytest.mark.parametrize('case', cases.DIFFUSION_CASES, ids=lambda c: c['name'])
def test_posterior_and_prediction_helpers(golden, case):
    """q_mean_variance, q_posterior_mean_variance, _predict_*, ddim_reverse_sample (gaussian_diffusion.py:171-227,
    374-396, 636-668)."""
    g, n = golden.npz('diffusion_extra'), case['name']
    s = D.Schedule(1000, case['schedule'], case['respacing'])
    shape = case['shape']
    x, x0, noise = synth.make_noise(shape, 11), synth.make_video(shape, 12), synth.make_noise(shape, 13)
    for tag, tl in case['ts'].items():
        t = torch.tensor(tl)
        k = f'{n}/{tag}/'
        close = lambda a, key, tol=1e-6: np.testing.assert_allclose(a.numpy(), g[k + key], rtol=tol, atol=tol, err_msg=k + key)
        for got, key in zip(D.q_mean_variance(s, x0, t), ('q_mean', 'q_var', 'q_logvar')):
            close(got, key)
        for got, key in zip(D.q_posterior_mean_variance(s, x0, x, t), ('post_mean', 'post_var', 'post_logvar')):
            close(got, key)
        close(D.predict_xstart_from_eps(s, x, t, noise), 'xstart_from_eps')
        close(D.predict_xstart_from_xprev(s, x, t, noise), 'xstart_from_xprev', 2e-6)
        close(D.predict_eps_from_xstart(s, x, t, x0), 'eps_from_xstart')
        close(D.ddim_reverse_sample(s, cases.fake_eps(x, s.model_time(t)), x, t)['sample'], 'ddim_reverse', 2e-6)


def test_ddim50_chain_matches_reference(golden):
    """Fixed-seed DDIM-50 `ddim_sample_loop` (north_star's sampled-frames gate), eta 0 and the stochastic eta 1."""
    from oracle import pipeline_oracle as P
    g = golden.npz('ddim50')
    c, d = cases.CHAIN_CASE, cases.DDIM50_CASE
    sd = synth.make_state_dict(golden.json('spec_' + c['cfg']), seed=1)
    cfg = U.model_config(**cases.ref_config(c['cfg']))
    video = synth.make_video((c['batch'], c['video_length'], 3, c['image_size'], c['image_size']), seed=c['video_seed'])
    sched = D.Schedule(1000, 'linear', d['respacing'])
    obs, lat = next(S.schedule(c['mode'], c['video_length'], c['obs_length'], c['max_frames'], c['step_size']))
    x0 = torch.cat([video[:, obs], torch.zeros_like(video[:, lat])], dim=1)
    fi, om, lm, km = (torch.from_numpy(a) for a in S.window_tensors(obs, lat, c['batch']))
    kw = dict(x0=x0, obs_mask=om, latent_mask=lm, kinda_marg_mask=km, frame_indices=fi)
    for eta in (0.0, 1.0):
        init = synth.make_noise(tuple(x0.shape), seed=d['noise_seed'])
        with torch.no_grad():
            out = P.ddim_sample_loop(sd, cfg, sched, init, kw, _Replay(d['noise_seed'] + 1), eta=eta)
        assert np.abs(out.numpy() - g[f'ddim50/eta{eta}']).max() < 2e-3, eta


def test_spec_shapes_cover_oracle_plan(golden):
    """Every key the oracle reads exists in the reference's state_dict spec (all four configs)."""
    for name in ('tiny', 'tiny_nrb2', 'c2', 'c4'):
        spec = golden.json('spec_' + name)
        cfg = U.model_config(**cases.ref_config(name))
        inp, mid, out, _ = U.block_plan(cfg)
        n_res = sum(k == 'res' for m in inp + [mid] + out for k, _ in m)
        n_attn = sum(k == 'attn' for m in inp + [mid] + out for k, _ in m)
        assert n_res == len([k for k in spec if k.endswith('in_layers.2.weight')])
        assert n_attn == len([k for k in spec if k.endswith('temporal_attention.qkv.weight')])


class _Replay:
    def __init__(self, base):
        self.i, self.base = 0, base

    def __call__(self, shape):
        z = synth.make_noise(shape, seed=self.base + self.i)
        self.i += 1
        return z


def test_tiny_chain_ddim_loop_and_elbo_match_reference(golden):
    from oracle import pipeline_oracle as P
    g = golden.npz('chain')
    c = cases.CHAIN_CASE
    sd = synth.make_state_dict(golden.json('spec_' + c['cfg']), seed=1)
    cfg = U.model_config(**cases.ref_config(c['cfg']))
    video = synth.make_video((c['batch'], c['video_length'], 3, c['image_size'], c['image_size']), seed=c['video_seed'])
    with torch.no_grad():
        sched = D.Schedule(1000, 'linear', c['respacing'])
        samples = P.infer_video(sd, cfg, sched, video, c['mode'], c['max_frames'], c['obs_length'], c['step_size'],
                                _Replay(c['noise_seed']))
        assert np.abs(samples.numpy() - g['chain/samples']).max() < 5e-4
        obs, lat = next(S.schedule(c['mode'], c['video_length'], c['obs_length'], c['max_frames'], c['step_size']))
        x0 = torch.cat([video[:, obs], torch.zeros_like(video[:, lat])], dim=1)
        fi, om, lm, km = (torch.from_numpy(a) for a in S.window_tensors(obs, lat, c['batch']))
        kw = dict(x0=x0, obs_mask=om, latent_mask=lm, kinda_marg_mask=km, frame_indices=fi)
        init = synth.make_noise(tuple(x0.shape), seed=c['noise_seed'] + 499)
        out = P.ddim_sample_loop(sd, cfg, sched, init, kw, _Replay(c['noise_seed'] + 500))
        assert np.abs(out.numpy() - g['ddim_loop/sample']).max() < 5e-4
        sched4 = D.Schedule(1000, 'linear', c['bpd_respacing'])
        raw, _ = P.run_bpd_evaluation(sd, cfg, sched4, cases.bpd_case_inputs(c), _Replay(c['noise_seed'] + 900))
        for k, v in raw.items():
            np.testing.assert_allclose(v.numpy(), g[f'bpd/{k}'], rtol=2e-4, atol=1e-6, err_msg=k)


def test_vertical_horizontal_schedule_matches_reference(golden):
    """scripts/video_sample_full.py:50-323 (vertical, then horizontal diffusion) -- oracle restatement vs the
    reference's model / p_sample run through the script's loop (tests/golden/chain_full.npz)."""
    from oracle import pipeline_oracle as P
    g = golden.npz('chain_full')
    c, f = cases.CHAIN_CASE, cases.FULL_SCHEDULE_CASE
    sd = synth.make_state_dict(golden.json('spec_' + c['cfg']), seed=1)
    cfg = U.model_config(**cases.ref_config(c['cfg']))
    video = synth.make_video((c['batch'], c['video_length'], 3, c['image_size'], c['image_size']), seed=c['video_seed'])
    with torch.no_grad():
        sched = D.Schedule(1000, 'linear', c['respacing'])
        samples = P.infer_video_full(sd, cfg, sched, video, c['mode'], c['max_frames'], c['obs_length'], c['step_size'],
                                     _Replay(f['noise_seed']), f['vertical_steps'])
    assert np.abs(samples.numpy() - g['full/samples']).max() < 5e-4
    assert np.abs(g['full/samples'] - g['full/vertical']).max() > 1e-2      # the horizontal sweeps did change the video
