"""The reference's training wrapper stack for E envs at once, on the device.

`make_quadrotor_env` of the reference returns
    QuadEnvCompatibility(QuadsRewardShapingWrapper(ExperienceReplayWrapper(QuadrotorEnvMulti)))
(swarm_rl/env_wrappers/quad_utils.py:20-110): three Python wrappers that walk lists of per-agent dicts and deep-copy the
env.  Here their per-step work is ONE kernel behind the step kernel (`qs_wrap_step`, csrc/qs_wrap.cuh): cumulative reward
terms, action statistics, true_reward, the episode_extra_stats sums, checkpoints every 0.5 s, collision events, replayed
episode starts — no PyTorch ops, no host synchronisation per step.  The host only

  * writes reward coefficients (shaping scheme, annealing: reward_shaping.py:55-61,110-118) — plain floats pushed with the
    next launch, and
  * fetches the statistics of the episodes finished so far whenever it wants to log them (`flush_stats`, one 600-byte copy).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L

TERM_NAMES = ('pos', 'action', 'crash', 'orient', 'spin', 'quadcol', 'proximity', 'quadcol_obstacle')      # QS_TERM_* order


def stats_dict(agg, use_obstacles, fallback_scenario):
    """Aggregate of finished episodes (QS_WA_* sums) -> the reference's `episode_extra_stats` keys, averaged over the
    agent-episodes / env-episodes it covers (Sample Factory averages these keys over episodes anyway)."""
    W = L.WA
    out = {}
    na, ne = float(agg[W['AGENT_EPISODES']]), float(agg[W['ENV_EPISODES']])
    if na > 0:
        out['rewraw_main'] = agg[W['TRUE_REWARD']] / na
        for k, name in enumerate(TERM_NAMES):
            if name == 'proximity':
                out['rew_proximity'] = agg[W['REW0'] + k] / na
                continue
            if name == 'quadcol_obstacle' and not use_obstacles:
                continue
            out[f'rewraw_{name}'] = agg[W['RAW0'] + k] / na
            out[f'rew_{name}'] = agg[W['REW0'] + k] / na
        out['rew_main'] = out['rew_pos']
        for k in range(4):
            out[f'z_action{k}_mean'] = agg[W['ACT_MEAN0'] + k] / na
            out[f'z_action{k}_std'] = agg[W['ACT_STD0'] + k] / na
        for k, name in enumerate(('1s', '3s', '5s')):
            out[f'distance_to_goal_{name}'] = agg[W['DIST0'] + k] / na
        out['metric/agent_success_rate'] = agg[W['SUCCESS']] / na
        out['metric/agent_deadlock_rate'] = agg[W['DEADLOCK']] / na
        out['metric/agent_col_rate'] = agg[W['COL']] / na
        out['metric/agent_neighbor_col_rate'] = agg[W['NEIGHBOR_COL']] / na
        out['metric/agent_obst_col_rate'] = agg[W['OBST_COL']] / na
    if ne > 0:
        for k in range(7):
            out[L.ENV_STAT_KEYS[k]] = agg[W['ENV_STAT0'] + k] / ne
        if use_obstacles:
            for k in range(7, 11):
                out[L.ENV_STAT_KEYS[k]] = agg[W['ENV_STAT0'] + k] / ne
    # per-scenario copies of the headline keys (reward_shaping.py:95-98, quadrotor_multi.py:680-718)
    for sid in range(16):
        row = agg[W['SCN0'] + 6 * sid: W['SCN0'] + 6 * sid + 6]
        if row[0] > 0:
            name = L.SCENARIO_NAMES.get(sid) or fallback_scenario
            out[f'Scenario_{name}/rew_pos'] = row[1] / row[0]
            out[f'Scenario_{name}/rew_crash'] = row[2] / row[0]
            out[f'{name}/distance_to_goal_1s'] = row[5] / row[0]
            if row[3] > 0:
                out[f'{name}/num_collisions'] = row[4] / row[3]
    nr = float(agg[W['REPLAY_ENV_EPISODES']])
    if nr > 0:                                              # quadrotor_multi.py:640-649
        out['num_collisions_replay'] = agg[W['REPLAY_COLLISIONS']] / nr
        if use_obstacles:
            out['num_collisions_obst_replay'] = agg[W['REPLAY_COLLISIONS_OBST']] / nr
    return {k: float(v) for k, v in out.items()}


class BatchedTrainingEnv:
    """E wrapped envs behind one object: `step(actions)` -> (obs, rewards, terminated, truncated, infos), CUDA tensors, the
    gymnasium 5-tuple of QuadEnvCompatibility (compatibility.py:33-50).  infos is empty on most steps; every `stats_every`
    steps (default: an episode length) it carries `episode_extra_stats` for the episodes finished since the last report."""

    def __init__(self, env, reward_shaping_scheme=None, annealing=None, replay_buffer_sample_prob=0.0, replay_buffer_size=20,
                 replay_always_active=False, stats_every=None):
        self.env, self.engine = env, env.engine
        self.reward_shaping_scheme = reward_shaping_scheme
        self.annealing = annealing
        self.training_info = {}                      # Sample Factory writes approx_total_training_steps here
        self.reward_shaping_updated = True
        self.replay_prob = float(replay_buffer_sample_prob)
        if self.replay_prob > 0.0 and env.device_scenario is None:
            raise ValueError("collision-event replay on the device needs a device-side scenario (host scenario objects are not "
                             "part of an env snapshot)")
        self.engine.wrap_enable(use_replay=self.replay_prob > 0.0, replay_buffer_size=replay_buffer_size,
                                replay_prob=self.replay_prob, replay_always_active=replay_always_active)
        self.stats_every = int(stats_every) if stats_every else self.engine.ep_len + 1
        self._since = 0
        self.num_agents = env.num_agents
        self.totals = dict(episodes=0.0, replayed_events=0.0, events_stored=0.0, checkpoints=0.0)

    def __getattr__(self, name):
        if name.startswith('_'):
            raise AttributeError(name)
        return getattr(self.env, name)

    # RewardShapingInterface (PBT), reward_shaping.py:33-44
    def get_default_reward_shaping(self):
        return dict(quad_rewards=dict())

    def get_current_reward_shaping(self, agent_idx):
        return dict(quad_rewards=dict())

    def set_reward_shaping(self, reward_shaping, unused_agent_idx):
        self.reward_shaping_scheme = dict(quad_rewards=dict())
        self.reward_shaping_updated = True

    def reset(self, **kw):
        return self.env.reset(**kw)

    def step(self, actions):
        env, eng = self.env, self.engine
        if self.reward_shaping_updated and self.reward_shaping_scheme:
            for key, weight in self.reward_shaping_scheme['quad_rewards'].items():
                eng.rew_coeff[key] = weight                       # pushed to the device with the next launch
            self.reward_shaping_updated = False
        obs, rew, term, trunc, infos = env.step(actions, wrapped=True)
        self._since += 1
        if self._since >= self.stats_every:
            infos = dict(infos)
            infos.update(self.flush_stats())
        return obs, rew, term, trunc, infos

    def flush_stats(self):
        """Statistics of the episodes finished since the last call (one small device -> host copy; synchronises)."""
        self._since = 0
        agg = self.engine.wrap_read(reset=True)
        W = L.WA
        n = float(agg[W['EPISODES_TOTAL']])
        self.totals['episodes'] += n
        self.totals['replayed_events'] += float(agg[W['REPLAYED_EVENTS']])
        self.totals['events_stored'] += float(agg[W['EVENTS_STORED']])
        self.totals['checkpoints'] += float(agg[W['CHECKPOINTS']])
        if n == 0:
            return {}
        stats = stats_dict(agg, self.env.use_obstacles, self.env.quads_mode)
        approx = self.training_info.get('approx_total_training_steps', 0)
        stats['z_approx_total_training_steps'] = approx
        if self.annealing:                                       # linear from 0 to the final value (reward_shaping.py:110-118)
            for sched in self.annealing:
                self.engine.rew_coeff[sched.coeff_name] = min(sched.final_value * approx / sched.anneal_env_steps, sched.final_value)
                stats[f'z_anneal_{sched.coeff_name}'] = self.engine.rew_coeff[sched.coeff_name]
        if self.replay_prob > 0.0:                               # quad_experience_replay.py:126-135
            ep = max(self.totals['episodes'], 1.0)
            stats['replay/replay_rate'] = self.totals['replayed_events'] / ep
            stats['replay/new_episode_rate'] = (ep - self.totals['replayed_events']) / ep
            stats['replay/events_stored'] = self.totals['events_stored']
        return {'episode_extra_stats': stats, 'true_reward': self.engine.wrap_true_reward(), 'episodes_finished': int(n)}
