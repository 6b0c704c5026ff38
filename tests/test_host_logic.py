"""CPU tests of the product's host-side logic: frame-index schedules (bit-exact vs the reference),
schedule tables / respacing, state_dict compatibility, C-ABI surface, and the world_size-2 gather."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import cases, diffusion_oracle as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_frame_index_schedules_bit_exact(golden):
    from video_diffusion_b200.inference_util import inference_strategies
    for rec in golden.json('frame_indices'):
        it = inference_strategies[rec['mode']](video_length=rec['T'], num_obs=rec['obs'], max_frames=rec['max_frames'],
                                               step_size=rec['step_size'])
        got = [[[int(i) for i in o], [int(i) for i in l]] for o, l in it]
        assert got == rec['steps'], rec['mode']


def test_exp_past_overrun_asserts_like_reference():
    from video_diffusion_b200.inference_util import inference_strategies
    with pytest.raises(AssertionError):      # (T - obs) % step != 0 -> latent index >= T
        list(inference_strategies['exp-past'](video_length=300, num_obs=36, max_frames=20, step_size=7))
    with pytest.raises(NotImplementedError):
        inference_strategies['adaptive-autoreg'](video_length=30, num_obs=5, max_frames=10, step_size=5)


def test_masks_match_reference_script():
    from video_diffusion_b200.sampling import get_masks
    x0 = torch.zeros(2, 7, 3, 4, 4)
    obs, lat, kin = get_masks(x0, 3)
    assert obs.shape == (2, 7, 1, 1, 1)
    assert obs.flatten().tolist() == [1, 1, 1, 0, 0, 0, 0] * 2
    assert torch.equal(lat, 1 - obs) and float(kin.abs().sum()) == 0


@pytest.mark.parametrize('case', cases.DIFFUSION_CASES, ids=lambda c: c['name'])
def test_schedule_tables_bit_exact(golden, case):
    from video_diffusion_b200 import create_gaussian_diffusion
    from video_diffusion_b200.gaussian_diffusion import device_tables
    g = golden.npz('diffusion')
    d = create_gaussian_diffusion(steps=1000, noise_schedule=case['schedule'], timestep_respacing=case['respacing'],
                                  rescale_timesteps=True, rescale_learned_sigmas=True)
    for attr in ('betas', 'alphas_cumprod', 'alphas_cumprod_prev', 'sqrt_alphas_cumprod',
                 'sqrt_one_minus_alphas_cumprod', 'log_one_minus_alphas_cumprod', 'sqrt_recip_alphas_cumprod',
                 'sqrt_recipm1_alphas_cumprod', 'posterior_variance', 'posterior_log_variance_clipped',
                 'posterior_mean_coef1', 'posterior_mean_coef2'):
        np.testing.assert_array_equal(getattr(d, attr), g[f"{case['name']}/{attr}"], err_msg=attr)
    assert d.timestep_map == g[f"{case['name']}/timestep_map"].tolist()
    s = D.Schedule(1000, case['schedule'], case['respacing'])
    assert torch.equal(device_tables(d), device_tables(s_like=s))


def test_state_dict_keys_and_shapes_match_reference(golden):
    from video_diffusion_b200 import create_video_model_and_diffusion, video_model_and_diffusion_defaults
    # tiny_lut: use_rpe_net=False lookup tables; tiny_dup / tiny_t0: other cond_emb_type input widths
    for name in ('tiny', 'tiny_nrb2', 'c2', 'c4', 'tiny_lut', 'tiny_dup', 'tiny_t0', 'tiny_fe'):
        kw = video_model_and_diffusion_defaults()
        kw.update(cases.ref_config(name))
        with torch.device('meta'):
            model, _ = create_video_model_and_diffusion(**kw)
        assert {k: list(v.shape) for k, v in model.state_dict().items()} == golden.json('spec_' + name)


def test_zero_init_modules_like_reference():
    from video_diffusion_b200 import create_video_model_and_diffusion, video_model_and_diffusion_defaults
    kw = video_model_and_diffusion_defaults()
    kw.update(cases.ref_config('tiny'))
    model, _ = create_video_model_and_diffusion(**kw)
    sd = model.state_dict()
    zero = [k for k in sd if re.search(r'out_layers\.3|proj_out|rpe_net\.out|^out\.2', k)]
    assert zero and all(float(sd[k].abs().sum()) == 0 for k in zero)
    assert float(sd['input_blocks.1.0.in_layers.2.weight'].abs().sum()) > 0


def test_cond_emb_type_constructors_like_reference():
    """unet.py:932-947: input widths and the `-initzero` initialisations of the conditioning variants."""
    from video_diffusion_b200 import create_video_model_and_diffusion, video_model_and_diffusion_defaults
    kw = video_model_and_diffusion_defaults()
    kw.update(cases.ref_config('tiny'))
    want = {'channel': (5, 0), 'channel-initzero': (5, 0), 'duplicate': (6, 1), 'duplicate-initzero': (6, 1),
            'all': (6, 1), 'all-initzero': (6, 1), 't=0': (3, 2)}
    for name, (cin, mode) in want.items():
        model, _ = create_video_model_and_diffusion(**dict(kw, cond_emb_type=name))
        w = model.state_dict()['input_blocks.0.0.weight']
        assert w.shape[1] == cin and model._cond_mode == mode and model.cond_emb_type == name.replace('-initzero', '')
        if name == 'channel-initzero':
            assert float(w[:, 3].abs().sum()) == 0 and float(w[:, 4].abs().sum()) > 0
        if name in ('duplicate-initzero', 'all-initzero'):
            assert torch.equal(w[:, 3:], w[:, :3])
    with pytest.raises(NotImplementedError):
        create_video_model_and_diffusion(**dict(kw, cond_emb_type='concat'))


def test_c_abi_exports_every_declared_symbol():
    from video_diffusion_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'vdm.h')).read()
    declared = sorted(set(re.findall(r'\b(vdm_[a-z0-9_]+)\s*\(', header)))
    assert declared and set(declared) == set(_lib.EXPORTS)
    lib = ctypes.CDLL(_lib.LIB_PATH)       # loads without a GPU; no compute calls here
    for name in declared:
        assert hasattr(lib, name), name
    lib.vdm_version.restype = ctypes.c_int
    assert lib.vdm_version() >= 100


def _gather_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from video_diffusion_b200 import dist as vd
    vd.init('gloo')
    tasks = vd.shard_tasks(5, rank, world)                      # 5 tasks of 2 videos, 9 videos in total
    ids = [i for t in tasks for i in vd.task_video_indices(t, 2, 9)]
    rows = torch.stack([torch.full((3, 2), i, dtype=torch.uint8) for i in ids])
    out, out_ids = vd.gather_ragged(rows, torch.tensor(ids))
    q.put((rank, out_ids.tolist(), out[:, 0, 0].tolist()))
    dist.destroy_process_group()


def test_sharding_and_ragged_gather_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, ids, vals in res:
        assert ids == list(range(9)) and vals == list(range(9))


def test_result_paths_and_run_identifiers_match_reference(golden, tmp_path):
    """Naming produced by the reference's own test_util.py (oracle/make_golden_formats.py)."""
    import argparse
    from video_diffusion_b200 import test_util as T
    g = golden.json('formats')
    for item in g['paths']:
        c = item['case']
        ck = tmp_path / c['checkpoint']
        ck.parent.mkdir(parents=True, exist_ok=True)
        torch.save(dict(state_dict={}, config={}, step=c['step']), ck)
        args = argparse.Namespace(use_ddim=c['use_ddim'], timestep_respacing=c['timestep_respacing'],
                                  eval_dir=c['eval_dir'], checkpoint_path=str(ck))
        assert str(T.get_model_results_path(args, postfix=c['postfix'])) == item['expect']
    for item in g['ids']:
        assert T.get_eval_run_identifier(argparse.Namespace(**item['case']), postfix=item['postfix']) == item['expect']


def test_checkpoint_loader_and_elbo_pickles(golden, tmp_path):
    """Reference checkpoint layout {'state_dict','config','step'} (test_util.py:31-62) with an old config that lacks
    the newer keys; ELBO pickles laid out like scripts/video_nll.py:126-137."""
    import pickle
    from oracle import synth
    from video_diffusion_b200 import test_util as T, video_model_and_diffusion_defaults
    cfg = video_model_and_diffusion_defaults()
    cfg.update(cases.ref_config('tiny'))
    for k in ('enforce_position_invariance', 'cond_emb_type'):      # a checkpoint written before these existed
        cfg.pop(k, None)
    sd = synth.make_state_dict(golden.json('spec_tiny'), seed=1)
    path = tmp_path / 'checkpoints' / 'run' / 'ema_latest.pt'
    path.parent.mkdir(parents=True)
    torch.save(dict(state_dict=sd, config=cfg, step=42), path)
    (model, diffusion), margs = T.load_checkpoint(str(path), 'cpu', use_ddim=True, timestep_respacing='10')
    assert margs.cond_emb_type == 'channel' and margs.enforce_position_invariance is False
    assert diffusion.num_timesteps == 10 and not model.training
    got = model.state_dict()
    assert set(got) == set(sd) and all(torch.equal(got[k], sd[k]) for k in sd)

    rets = [dict(total_bpd=np.arange(3.0) + i, vb=np.ones((3, 5)) * i) for i in range(2)]   # two index types
    paths = T.save_elbos(tmp_path / 'eval', rets, dataset_indices=[7, 8, 11], postfix='_s')
    assert [p.name for p in paths] == ['elbo_7_s.pkl', 'elbo_8_s.pkl', 'elbo_11_s.pkl']
    d = pickle.load(open(paths[1], 'rb'))
    assert d['total_bpd'].shape == (2,) and d['vb'].shape == (2, 5) and d['total_bpd'][1] == 2.0


def test_micro_batch_join_range_of_the_launch_plan():
    """micro_batch_join_hw (opt-in): the plan pieces between which the two groups of videos meet -- from the downsample
    whose output has at most that many pixels to the upsample that leaves those levels.  Pure host logic."""
    from video_diffusion_b200 import create_video_model_and_diffusion, video_model_and_diffusion_defaults
    kw = video_model_and_diffusion_defaults()
    kw.update(cases.ref_config('c2'))
    model, _ = create_video_model_and_diffusion(**kw)
    want = {0: (None, None), 4096: (None, None),           # off / no downsample lands at 64x64
            64: ('input_blocks.9.0.op', 'output_blocks.2.2.conv'),        # the 8x8 level only
            256: ('input_blocks.6.0.op', 'output_blocks.5.2.conv'),       # 16x16 and 8x8
            1024: ('input_blocks.3.0.op', 'output_blocks.8.1.conv')}      # 32x32 and below
    for hw, (p_first, p_last) in want.items():
        model.micro_batch_join_hw = hw
        first, last = model._deep_range(64, 64)
        if p_first is None:
            assert first is None and last is None
            continue
        assert (model.plan[first]['kind'], model.plan[first]['p']) == ('down', p_first)
        assert (model.plan[last]['kind'], model.plan[last]['p']) == ('up', p_last)
        # between them the pushes and pops of the skip stack balance: one push per input block, one pop per res block
        # of the output path
        pushes = len({n['group'] for n in model.plan[first:last] if n['group'].startswith('input')})
        pops = sum(1 for n in model.plan[first + 1:last] if n['kind'] == 'res' and n.get('cat'))
        assert pushes == pops


def test_fused_temporal_kernel_is_chosen_where_it_fits():
    """Host-side dispatch of the temporal attention block: the fused kernel (16 pixels per CTA for 96-wide heads, else 8)
    wherever its shared memory fits, the three-launch path elsewhere.  vdm_attn_temporal_fused_smem is a pure host
    function of the C ABI: no GPU needed."""
    from video_diffusion_b200 import create_video_model_and_diffusion, ops, video_model_and_diffusion_defaults
    assert ops.attn_temporal_fused_smem(20, 96, 24, 8) <= 113 * 1024            # two CTAs per SM
    assert ops.attn_temporal_fused_smem(20, 96, 24, 16) <= 227 * 1024
    assert ops.attn_temporal_fused_smem(32, 128, 32, 8) > 227 * 1024            # too large: fallback
    assert ops.attn_temporal_fused_smem(20, 48, 24, 8) == -1                    # head width not instantiated
    kw = video_model_and_diffusion_defaults()
    kw.update(cases.ref_config('c2'))
    model, _ = create_video_model_and_diffusion(**kw)
    assert model._temporal_pt(20, 384, 256) == 16 and model._temporal_pt(20, 512, 64) == 8
    assert model._temporal_pt(20, 384, 8) == 8                                  # 8 pixels per image: one tile of 8
    assert model._temporal_pt(32, 512, 64) == 0 and model._temporal_pt(33, 384, 256) == 0
    assert model._temporal_pt(20, 64, 256) == 0                                 # 16-wide heads (the tiny model)
    model.fused_temporal = False
    assert model._temporal_pt(20, 384, 256) == 0
    import torch
    model2, _ = create_video_model_and_diffusion(compute_dtype=torch.float32, **kw)
    assert model2._temporal_pt(20, 384, 256) == 0                               # fp32 mode keeps the SIMT kernels
