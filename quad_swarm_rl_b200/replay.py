"""Collision-event replay for the B200 env.

Semantics of gym_art/quadrotor_multi/quad_experience_replay.py: while the drones can already fly, a snapshot of the env
is taken every 0.5 s; when a collision happens the snapshot from 1.5 s earlier goes into a 20-slot buffer; at every
episode start a buffered event is replayed with probability p instead of a fresh reset.  The reference deep-copies the
whole Python env; here a snapshot is the SoA device state (`qs_get_state`) plus the small host-side episode state, and
replaying is `qs_set_state` — no Python object graph is copied.
"""
import random
from collections import deque

import numpy as np

from .wrappers import Wrapper


class ReplayBufferEvent:
    def __init__(self, snapshot, obs):
        self.snapshot = snapshot
        self.obs = obs
        self.num_replayed = 0


class ReplayBuffer:
    """quad_experience_replay.py:16-63."""

    def __init__(self, control_frequency, cp_step_size=0.5, buffer_size=20):
        self.control_frequency = control_frequency
        self.cp_step_size_sec = cp_step_size
        self.cp_step_size_freq = self.cp_step_size_sec * self.control_frequency
        self.buffer_idx = 0
        self.buffer = deque([], maxlen=buffer_size)

    def write_cp_to_buffer(self, env, snapshot, obs):
        env.saved_in_replay_buffer = True
        evt = ReplayBufferEvent(snapshot, obs)
        if len(self.buffer) < self.buffer.maxlen:
            self.buffer.append(evt)
        else:
            self.buffer[self.buffer_idx] = evt
        self.buffer_idx = (self.buffer_idx + 1) % self.buffer.maxlen

    def sample_event(self):
        idx = random.randint(0, len(self.buffer) - 1)
        self.buffer[idx].num_replayed += 1
        return self.buffer[idx]

    def cleanup(self):
        self.buffer = deque([e for e in self.buffer if e.num_replayed < 10], maxlen=self.buffer.maxlen)

    def avg_num_replayed(self):
        stats = [e.num_replayed for e in self.buffer]
        return np.mean(stats) if stats else 0

    def __len__(self):
        return len(self.buffer)


class ExperienceReplayWrapper(Wrapper):
    """quad_experience_replay.py:66-209 on top of env.snapshot() / env.restore()."""

    def __init__(self, env, replay_buffer_sample_prob, default_obst_density, defulat_obst_size,
                 domain_random=False, obst_density_random=False, obst_size_random=False,
                 obst_density_min=0., obst_density_max=0., obst_size_min=0, obst_size_max=0.):
        super().__init__(env)
        self.replay_buffer = ReplayBuffer(env.envs[0].control_freq)
        self.replay_buffer_sample_prob = replay_buffer_sample_prob
        self.curr_obst_density = default_obst_density
        self.curr_obst_size = defulat_obst_size
        self.domain_random = domain_random
        self.obst_density_random = obst_density_random and domain_random
        self.obst_size_random = obst_size_random and domain_random
        if self.obst_density_random:
            self.obst_densities = np.arange(obst_density_min, obst_density_max, 0.05)
            self.curr_obst_density = 0.
        if self.obst_size_random:
            self.obst_sizes = np.arange(obst_size_min, obst_size_max, 0.1)
            self.curr_obst_size = 0.
        self.max_episode_checkpoints_to_keep = int(3.0 / self.replay_buffer.cp_step_size_sec)
        self.episode_checkpoints = deque([], maxlen=self.max_episode_checkpoints_to_keep)
        self.save_time_before_collision_sec = 1.5
        self.last_tick_added_to_buffer = -1e9
        self.replayed_events = 0
        self.episode_counter = 0

    def _randomised(self):
        obst_density = obst_size = None
        if self.obst_density_random:
            obst_density = np.random.choice(self.obst_densities)
            self.curr_obst_density = obst_density
        if self.obst_size_random:
            obst_size = np.random.choice(self.obst_sizes)
            self.curr_obst_size = obst_size
        return obst_density, obst_size

    def reset(self):
        return self.env.reset(*self._randomised())

    def step(self, action):
        env = self.env
        obs, rewards, dones, infos = env.step(action)
        if any(dones):
            obs = self.new_episode()
            for i in range(len(infos)):
                if not infos[i].get("episode_extra_stats"):
                    infos[i]["episode_extra_stats"] = dict()
                infos[i]["episode_extra_stats"].update({
                    "replay/replay_rate": self.replayed_events / self.episode_counter,
                    "replay/new_episode_rate": (self.episode_counter - self.replayed_events) / self.episode_counter,
                    "replay/replay_buffer_size": len(self.replay_buffer),
                    "replay/avg_replayed": self.replay_buffer.avg_num_replayed(),
                    "replay/obst_density": self.curr_obst_density,
                    "replay/obst_size": self.curr_obst_size,
                })
        else:
            tick, freq = env.envs[0].tick, env.envs[0].control_freq
            if env.use_replay_buffer and env.activate_replay_buffer and not env.saved_in_replay_buffer \
                    and tick % self.replay_buffer.cp_step_size_freq == 0:
                self.episode_checkpoints.append((env.snapshot(), np.array(obs, copy=True)))
            collision_flag = bool(np.asarray(env.last_step_unique_collisions).any())
            if env.use_obstacles:
                collision_flag = collision_flag or len(env.curr_quad_col) > 0
            if collision_flag and env.use_replay_buffer and env.activate_replay_buffer \
                    and tick > env.collisions_grace_period_seconds * freq and not env.saved_in_replay_buffer:
                if tick - self.last_tick_added_to_buffer > 5 * freq:
                    steps_ago = int(self.save_time_before_collision_sec / self.replay_buffer.cp_step_size_sec)
                    if steps_ago > len(self.episode_checkpoints):
                        raise IndexError(f"Tried to read past the boundary of checkpoint_history. Steps ago: {steps_ago}, "
                                         f"episode checkpoints: {len(self.episode_checkpoints)}, {tick}")
                    snap, snap_obs = self.episode_checkpoints[-steps_ago]
                    self.replay_buffer.write_cp_to_buffer(env, snap, snap_obs)
                    env.collision_occurred = False
                    self.last_tick_added_to_buffer = tick
        return obs, rewards, dones, infos

    def new_episode(self):
        env = self.env
        self.episode_counter += 1
        self.last_tick_added_to_buffer = -1e9
        self.episode_checkpoints = deque([], maxlen=self.max_episode_checkpoints_to_keep)
        if np.random.uniform(0, 1) < self.replay_buffer_sample_prob and env.activate_replay_buffer and len(self.replay_buffer) > 0:
            self.replayed_events += 1
            event = self.replay_buffer.sample_event()
            env.restore(event.snapshot, zero_collision_counters=True)      # the snapshot was taken with saved_in_replay_buffer = True
            self.curr_obst_density = env.obst_density
            self.replay_buffer.cleanup()
            return np.array(event.obs, copy=True)
        obs = env.reset(*self._randomised())
        env.saved_in_replay_buffer = False
        return obs
