"""Frame-index schedules for long-video sampling, with the reference's interface
(`improved_diffusion/inference_util.py`): `inference_strategies[mode](video_length=, num_obs=,
max_frames=, step_size=, optimal_schedule_path=None)` is an iterator of
`(obs_frame_indices, latent_frame_indices)` lists.  Pure integer host logic; the parity bar is
bit-exact equality with the reference (tests/golden/frame_indices.json).

Covered: autoreg (:232-246), independent (:249-260), really-independent (:263-274), exp-past
(:277-293, including its unclipped latent range -- SURVEY Q7), mixed-autoreg-independent
(:296-312) and hierarchy-N (:315-420).  The adaptive / LPIPS-driven and visualisation-only
strategies need the `lpips` network weights and are out of scope: asking for them raises.
"""
import numpy as np


class FrameSchedule:
    """Base iterator: bookkeeping of generated frames + sanity checks (reference :80-123)."""
    name = 'base'

    def __init__(self, video_length, num_obs, max_frames, step_size, optimal_schedule_path=None):
        self._video_length = video_length
        self._max_frames = max_frames
        self._num_obs = num_obs
        self._step_size = step_size
        self._done_frames = set(range(num_obs))
        self._obs_frames = list(range(num_obs))
        self._current_step = 0
        self.optimal_schedule = None
        if optimal_schedule_path is not None:
            import torch
            self.optimal_schedule = torch.load(optimal_schedule_path)

    @property
    def typename(self):
        return type(self).__name__

    def __iter__(self):
        self.step = 0
        return self

    def is_done(self):
        return len(self._done_frames) >= self._video_length

    def first_unconditional_window(self):
        return list(range(self._max_frames))

    def next_indices(self):
        raise NotImplementedError

    def __next__(self):
        if self.is_done():
            raise StopIteration
        unconditional = self._num_obs == 0 and self._current_step == 0
        if unconditional:
            obs, lat = [], self.first_unconditional_window()
        else:
            obs, lat = self.next_indices()
            if self.optimal_schedule is not None:
                obs = self.optimal_schedule.get(self._current_step, [])
        assert isinstance(obs, list) and isinstance(lat, list)
        for i in obs:
            assert i in self._done_frames, (
                f'Attempting to condition on frame {i} while it is not generated yet.\n'
                f'Generated frames: {self._done_frames}\nObserving: {obs}\nGenerating: {lat}')
        assert np.all(np.array(lat) < self._video_length)
        self._done_frames.update(lat)
        if unconditional:
            self._obs_frames = lat
        self._current_step += 1
        return obs, lat

    # helpers shared by several strategies
    def _next_block(self, width):
        start = max(self._done_frames) + 1
        return list(range(start, min(start + width, self._video_length)))


class Autoregressive(FrameSchedule):
    def next_indices(self):
        if not self._done_frames:
            return [], list(range(self._max_frames))
        obs = sorted(self._done_frames)[-(self._max_frames - self._step_size):]
        start = obs[-1] + 1
        return obs, list(range(start, min(start + self._step_size, self._video_length)))


class Independent(FrameSchedule):
    def next_indices(self):
        obs = sorted(self._obs_frames)[-(self._max_frames - self._step_size):]
        return obs, self._next_block(self._step_size)


class ReallyIndependent(FrameSchedule):
    def next_indices(self):
        return [], self._next_block(self._max_frames)


class ExpPast(FrameSchedule):
    def next_indices(self):
        cur = max(self._done_frames) + 1
        obs = list(cur - 2 ** np.arange(int(np.log2(cur))))
        lat = list(range(cur, cur + min(self._step_size, self._video_length)))    # not clipped to T (Q7)
        for back in range(1, cur + 1):
            if len(obs) + len(lat) >= self._max_frames:
                break
            if cur - back not in obs:
                obs.append(cur - back)
        return obs, lat


class MixedAutoregressiveIndependent(FrameSchedule):
    def next_indices(self):
        n_cond = self._max_frames - self._step_size
        cond = set(sorted(self._done_frames)[-(n_cond // 2):])
        for i in sorted(self._obs_frames, reverse=True):
            cond.add(i)
            if len(cond) == n_cond:
                break
        return sorted(cond), self._next_block(self._step_size)


class HierarchyNLevel(FrameSchedule):
    N = None

    @property
    def typename(self):
        return f'{super().typename}-{self.N}'

    def _coarse_grid(self):
        self.current_level = 1
        self.last_sampled_idx = self._video_length - 1
        return [int(i) for i in np.linspace(0, self._video_length - 1, self._max_frames)]

    def first_unconditional_window(self):
        return self._coarse_grid()

    @property
    def sample_every(self):
        level1 = (self._video_length - len(self._obs_frames)) / (self._step_size - 1)
        return int(level1 ** ((self.N - self.current_level) / (self.N - 1)))

    def next_indices(self):
        T, done = self._video_length, self._done_frames
        if not done:
            return [], self._coarse_grid()
        if len(done) == len(self._obs_frames):
            self.current_level = 1
            self.last_sampled_idx = max(self._obs_frames)
        n_cond, n_new = self._max_frames - self._step_size, self._step_size
        idx = self.last_sampled_idx + self.sample_every
        if all(i in done for i in range(idx, T)):
            self.current_level += 1
            self.last_sampled_idx = 0
            idx = min(i for i in range(T) if i not in done) - 1 + self.sample_every
        if self.current_level == 1:
            lat = [int(i) for i in np.linspace(max(self._obs_frames) + 1, T - 0.001, n_new)]
        else:
            lat = []
            while len(lat) < n_new and idx < T:
                if idx in done:
                    idx += 1
                else:
                    lat.append(idx)
                    idx += self.sample_every
        obs = [i for i in range(min(lat), max(lat)) if i in done]
        around = n_cond - len(obs)
        if around < 2:      # shrink the step until frames before AND after the latents fit
            if self._step_size == 1:
                raise Exception('Cannot condition before and after even with step size of 1')
            self._step_size -= 1
            try:
                return self.next_indices()
            finally:
                self._step_size += 1
        obs.extend([i for i in range(max(lat) + 1, T) if i in done][:around // 2])
        n_before = n_cond - len(obs)
        if self.current_level == 1:
            obs.extend(list(np.linspace(0, max(self._obs_frames) + 0.999, n_before).astype(np.int32)))
        else:
            obs.extend([i for i in range(min(lat) - 1, -1, -1) if i in done][:n_before])
        self.last_sampled_idx = max(lat)
        return obs, lat


def get_hierarchy_n_level(n):
    return type('Hierarchy', (HierarchyNLevel,), {'N': n})


class _Unsupported:
    def __init__(self, mode):
        self.mode = mode

    def __call__(self, *a, **k):
        raise NotImplementedError(f"inference mode '{self.mode}' (adaptive / visualisation-only) needs the lpips "
                                  'embedder or is outside the accelerated path')


inference_strategies = {
    'autoreg': Autoregressive,
    'independent': Independent,
    'really-independent': ReallyIndependent,
    'exp-past': ExpPast,
    'mixed-autoreg-independent': MixedAutoregressiveIndependent,
    'hierarchy-2': get_hierarchy_n_level(2),
    'hierarchy-3': get_hierarchy_n_level(3),
    'hierarchy-4': get_hierarchy_n_level(4),
    'hierarchy-5': get_hierarchy_n_level(5),
}
for _m in ('adaptive-autoreg', 'adaptive-hierarchy-2', 'adaptive-hierarchy-3', 'goal-directed-autoreg',
           'goal-directed-mixed', 'goal-directed-hierarchy-2', 'ho-et-al-for-vis', 'baby-cond-ho-et-al-for-vis',
           'google', 'like-google'):
    inference_strategies[_m] = _Unsupported(_m)
