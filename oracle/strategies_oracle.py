"""CPU oracle for the frame-index schedules.  TEST INFRASTRUCTURE ONLY.

Stateless generator restatement of the non-adaptive strategies in the reference's
`improved_diffusion/inference_util.py` (base iterator :80-123; Autoregressive
:232-246; Independent :249-260; ReallyIndependent :263-274; ExpPast :277-293;
MixedAutoregressiveIndependent :296-312; HierarchyNLevel :315-420).  Pure integer
work: the parity bar is bit-exact equality with tests/golden/frame_indices.json,
which oracle/make_golden.py dumps from the reference itself.
"""
import numpy as np


def _finish(step, video_length, done):
    obs, lat = step
    assert all(i in done for i in obs), (obs, lat, sorted(done))
    assert all(i < video_length for i in lat), (lat, video_length)
    done.update(lat)
    return [int(i) for i in obs], [int(i) for i in lat]


def schedule(mode, video_length, num_obs, max_frames, step_size):
    """Yield (obs_indices, latent_indices) until every frame is produced."""
    done = set(range(num_obs))
    observed = list(range(num_obs))
    state = {}
    first = True
    while len(done) < video_length:
        if num_obs == 0 and first:
            if mode.startswith('hierarchy'):
                state.update(level=1, last=video_length - 1)
                lat = [int(i) for i in np.linspace(0, video_length - 1, max_frames)]
            else:
                lat = list(range(max_frames))
            step = ([], lat)
            observed = lat
        else:
            step = _NEXT[mode.split('-')[0] if mode.startswith('hierarchy') else mode](
                mode, video_length, max_frames, step_size, done, observed, state)
        first = False
        yield _finish(step, video_length, done)


def _autoreg(mode, T, max_frames, step, done, observed, state):
    obs = sorted(done)[-(max_frames - step):]
    start = obs[-1] + 1
    return obs, list(range(start, min(start + step, T)))


def _independent(mode, T, max_frames, step, done, observed, state):
    obs = sorted(observed)[-(max_frames - step):]
    start = max(done) + 1
    return obs, list(range(start, min(start + step, T)))


def _really_independent(mode, T, max_frames, step, done, observed, state):
    start = max(done) + 1
    return [], list(range(start, min(start + max_frames, T)))


def _exp_past(mode, T, max_frames, step, done, observed, state):
    cur = max(done) + 1
    obs = [cur - int(2 ** e) for e in range(int(np.log2(cur)))]
    lat = list(range(cur, cur + min(step, T)))          # NOT clipped to T
    back = 1
    while back <= cur and len(obs) + len(lat) < max_frames:
        if cur - back not in obs:
            obs.append(cur - back)
        back += 1
    return obs, lat


def _mixed(mode, T, max_frames, step, done, observed, state):
    n_cond = max_frames - step
    cond = set(sorted(done)[-(n_cond // 2):])
    for i in sorted(observed, reverse=True):
        cond.add(i)
        if len(cond) == n_cond:
            break
    start = max(done) + 1
    return sorted(cond), list(range(start, min(start + step, T)))


def _hierarchy(mode, T, max_frames, step, done, observed, state):
    n_levels = int(mode.split('-')[1])
    if len(done) == len(observed):
        state.update(level=1, last=max(observed))

    def every():
        lvl1 = (T - len(observed)) / (step - 1)
        return int(lvl1 ** ((n_levels - state['level']) / (n_levels - 1)))

    n_cond, n_new = max_frames - step, step
    idx = state['last'] + every()
    if not [i for i in range(idx, T) if i not in done]:
        state['level'] += 1
        state['last'] = 0
        idx = min(i for i in range(T) if i not in done) - 1 + every()
    if state['level'] == 1:
        lat = [int(i) for i in np.linspace(max(observed) + 1, T - 0.001, n_new)]
    else:
        lat = []
        while len(lat) < n_new and idx < T:
            if idx not in done:
                lat.append(idx)
                idx += every()
            else:
                idx += 1
    obs = [i for i in range(min(lat), max(lat)) if i in done]
    around = n_cond - len(obs)
    if around < 2:
        if step == 1:
            raise Exception('Cannot condition before and after even with step size of 1')
        return _hierarchy(mode, T, max_frames, step - 1, done, observed, state)
    obs.extend([i for i in range(max(lat) + 1, T) if i in done][:around // 2])
    n_before = n_cond - len(obs)
    if state['level'] == 1:
        obs.extend(list(np.linspace(0, max(observed) + 0.999, n_before).astype(np.int32)))
    else:
        obs.extend([i for i in range(min(lat) - 1, -1, -1) if i in done][:n_before])
    state['last'] = max(lat)
    return obs, lat


_NEXT = {'autoreg': _autoreg, 'independent': _independent, 'really-independent': _really_independent,
         'exp-past': _exp_past, 'mixed-autoreg-independent': _mixed, 'hierarchy': _hierarchy}


def window_tensors(obs, lat, batch_size):
    """Frame indices and masks of one window, as scripts/video_sample.py:119-132 builds them."""
    frame_indices = np.tile(np.array(list(obs) + list(lat), dtype=np.int64), (batch_size, 1))
    obs_mask = np.zeros((batch_size, len(obs) + len(lat), 1, 1, 1), dtype=np.float32)
    obs_mask[:, :len(obs)] = 1
    return frame_indices, obs_mask, 1 - obs_mask, np.zeros_like(obs_mask)
