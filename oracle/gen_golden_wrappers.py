"""Golden vectors of the reference's two training wrappers — TEST INFRASTRUCTURE ONLY.

    python -m oracle.gen_golden_wrappers        (needs /root/reference)  ->  tests/golden/wrappers.npz

1. swarm_rl/env_wrappers/reward_shaping.py::QuadsRewardShapingWrapper (UNMODIFIED) is driven by a stand-in env that emits
   synthetic per-step `infos[i]['rewards']` dicts (the keys of quadrotor_single.py:68-85 / quadrotor_multi.py:533-540 built
   from raw terms and the coefficients in force), actions and dones.  Recorded: the per-step inputs and, per episode, every
   agent's `true_reward` and `episode_extra_stats`.  The wrapper kernel (csrc/qs_wrap.cuh) is fed the same per-step inputs
   through qs_wrap_apply on the GPU box and must reproduce those statistics.
2. gym_art/quadrotor_multi/quad_experience_replay.py::ExperienceReplayWrapper (UNMODIFIED) is driven by a stand-in env with
   a tick counter and scheduled collisions.  Recorded: which checkpoint tick each collision stored, which ticks replayed
   episodes started from.  tests/test_gpu_batched.py plants collisions at the same ticks in the real env.
"""
import contextlib
import io
import json
import os
import random

import numpy as np

from . import ref_harness as rh

TERM_KEYS = ('pos', 'action', 'crash', 'orient', 'spin', 'quadcol', 'proximity', 'quadcol_obstacle')


class _ShapingEnv:
    """QuadrotorEnvMulti's face towards the reward-shaping wrapper: rew_coeff, scenario.name(), is_multiagent, step()."""
    is_multiagent = True

    def __init__(self, N, terms, dones, coeff):
        self.num_agents = N
        self.terms, self.dones = terms, dones
        self.rew_coeff = dict(coeff)
        self.scenario = type('S', (), {'name': lambda self: 'Scenario_static_same_goal'})()
        self.t = 0

    @property
    def unwrapped(self):
        return self

    def reset(self):
        return np.zeros((self.num_agents, 3))

    def step(self, action):
        raw = self.terms[self.t]
        c = self.rew_coeff
        w = [c['pos'], c['effort'], c['crash'], c['orient'], c['spin'], c['quadcol_bin'], 1.0, c['quadcol_bin_obst']]
        infos = []
        for i in range(self.num_agents):
            r = {}
            for k, name in enumerate(TERM_KEYS):
                if name == 'proximity':
                    r['rew_proximity'] = raw[i, k]                       # delivered weighted (quadrotor_multi.py:507-513)
                    continue
                r[f'rewraw_{name}'] = raw[i, k]
                r[f'rew_{name}'] = raw[i, k] * w[k]
            r['rewraw_main'], r['rew_main'] = r['rewraw_pos'], r['rew_pos']
            infos.append({'rewards': r})
        d = bool(self.dones[self.t])
        self.t += 1
        return np.zeros((self.num_agents, 3)), [0.0] * self.num_agents, [d] * self.num_agents, infos


def shaping_case(seed, N=4, T=90, ep=30):
    rh._ensure_path()
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'stubs'))
    from swarm_rl.env_wrappers.reward_shaping import QuadsRewardShapingWrapper
    AnnealSchedule = type('AnnealSchedule', (), {'__init__': lambda self, n, f, s_: (setattr(self, 'coeff_name', n), setattr(self, 'final_value', f),
                                                                                 setattr(self, 'anneal_env_steps', s_)) and None})   # quad_utils.py:13-17
    rs = np.random.RandomState(seed)
    terms = -np.abs(rs.normal(0, 0.01, (T, N, 8))).astype(np.float32).astype(np.float64)
    terms[..., 5] = -(rs.uniform(size=(T, N)) < 0.05).astype(np.float64)         # rewraw_quadcol is 0 / -1
    terms[..., 7] = -(rs.uniform(size=(T, N)) < 0.03).astype(np.float64)
    terms[..., 2] = -0.005 * (rs.uniform(size=(T, N)) < 0.2)
    actions = rs.uniform(-1, 1, (T, N, 4)).astype(np.float32).astype(np.float64)
    dones = np.array([(t + 1) % ep == 0 for t in range(T)])
    coeff0 = dict(pos=1.0, effort=0.05, crash=1.0, orient=1.0, spin=0.1, quadcol_bin=5.0, quadcol_bin_smooth_max=4.0, quadcol_bin_obst=5.0,
                  vel=0.0, yaw=0.0)
    scheme = dict(quad_rewards=dict(pos=2.0, quadcol_bin=0.0, quadcol_bin_obst=1.5))
    env = _ShapingEnv(N, terms, dones, coeff0)
    w = QuadsRewardShapingWrapper(env, reward_shaping_scheme=scheme, annealing=[AnnealSchedule('quadcol_bin', 5.0, 1000.0)])
    w.training_info['approx_total_training_steps'] = 300
    w.reset()
    episodes, coeffs = [], []
    for t in range(T):
        coeffs.append([env.rew_coeff[k] if t > 0 or k not in scheme['quad_rewards'] else scheme['quad_rewards'][k]
                       for k in ('pos', 'effort', 'crash', 'orient', 'spin', 'quadcol_bin', 'quadcol_bin_smooth_max', 'quadcol_bin_obst')])
        obs, rew, d, infos = w.step(actions[t])
        if d[0]:
            episodes.append(dict(t=t, true_reward=[float(i['true_reward']) for i in infos],
                                 stats=[{k: float(v) for k, v in i['episode_extra_stats'].items()} for i in infos]))
    return dict(terms=terms, actions=actions, dones=dones, coeffs=np.array(coeffs), episodes=episodes)


class _ReplayEnv:
    """QuadrotorEnvMulti's face towards the replay wrapper (quad_experience_replay.py:71,140-151,183-188): a tick counter,
    scheduled collisions, the flags the wrapper reads and writes; deepcopy-able."""

    def __init__(self, ep_len, collisions):
        self.envs = [type('E', (), {'tick': 0, 'control_freq': 100.0})()]
        self.ep_len, self.collisions = ep_len, set(collisions)
        self.use_replay_buffer = True
        self.activate_replay_buffer = True
        self.saved_in_replay_buffer = False
        self.collisions_grace_period_seconds = 1.5
        self.use_obstacles = False
        self.last_step_unique_collisions = np.array([])
        self.curr_quad_col = []
        self.scenes = []
        self.obst_density = 0.2
        self.collisions_per_episode = self.collisions_after_settle = self.obst_quad_collisions_per_episode = self.obst_quad_collisions_after_settle = 0
        self.episode = 0
        self.started_at = []

    @property
    def unwrapped(self):
        return self

    def reset(self, obst_density=None, obst_size=None):
        self.envs[0].tick = 0
        self.episode += 1
        return np.zeros((2, 3))

    def step(self, action):
        e = self.envs[0]
        e.tick += 1
        hit = (self.episode, e.tick) in self.collisions and not self.saved_in_replay_buffer
        self.last_step_unique_collisions = np.array([0, 1]) if hit else np.array([])
        done = e.tick > self.ep_len
        infos = [{'rewards': {}}, {'rewards': {}}]
        if done:
            for i in infos:
                i['episode_extra_stats'] = {'num_collisions': 0.0}
        return np.full((2, 3), float(e.tick)), [0., 0.], [done, done], infos


def replay_case(seed):
    rh._ensure_path()
    from gym_art.quadrotor_multi.quad_experience_replay import ExperienceReplayWrapper
    random.seed(seed)
    np.random.seed(seed)
    # fresh episode 1: collisions at ticks 201 (stored), 260 (inside the 5 s cooldown: ignored), 720 (stored), 120 (grace)
    env = _ReplayEnv(ep_len=900, collisions=[(1, 120), (1, 201), (1, 260), (1, 720)])
    with contextlib.redirect_stdout(io.StringIO()):
        w = ExperienceReplayWrapper(env, 1.0, 0.2, 0.6)
        w.reset()
        stored, starts = [], []
        n0 = 0
        for t in range(3000):
            obs, rew, dones, infos = w.step(None)
            if len(w.replay_buffer) > n0:
                n0 = len(w.replay_buffer)
                stored.append(int(w.replay_buffer.buffer[-1].env.envs[0].tick))
            if dones[0]:
                starts.append(int(w.env.envs[0].tick))               # tick the next episode starts from (0 = fresh)
    return dict(collision_ticks=[120, 201, 260, 720], stored_checkpoint_ticks=stored, episode_start_ticks=starts,
                replayed_events=int(w.replayed_events), episodes=int(w.episode_counter))


def main():
    out = {}
    cases = [shaping_case(1), shaping_case(2, N=8, T=120, ep=40)]
    for k, c in enumerate(cases):
        for name in ('terms', 'actions', 'dones', 'coeffs'):
            out[f'shaping{k}_{name}'] = c[name]
        out[f'shaping{k}_episodes'] = np.array(json.dumps(c['episodes']))
    out['replay'] = np.array(json.dumps(replay_case(5)))
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'wrappers.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, json.loads(str(out['replay'])))


if __name__ == '__main__':
    main()
