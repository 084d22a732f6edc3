"""Wrapper semantics of swarm_rl/env_wrappers (reward shaping + annealing, 5-tuple compatibility) and the env factory.

Sample Factory and gymnasium are not installed in this image, so the wrappers are plain delegating objects that
follow the same protocol (`reset`, `step`, `unwrapped`, attribute forwarding).  When Sample Factory is present the
factory below can be registered as is: `register_env("quadrotor_multi", make_quadrotor_env)` (swarm_rl/train.py:18).
"""
import copy

import numpy as np

# swarm_rl/env_wrappers/reward_shaping.py:7-16
DEFAULT_QUAD_REWARD_SHAPING_SINGLE = dict(
    quad_rewards=dict(pos=1.0, effort=0.05, spin=0.1, vel=0.0, crash=1.0, orient=1.0, yaw=0.0),
)
DEFAULT_QUAD_REWARD_SHAPING = copy.deepcopy(DEFAULT_QUAD_REWARD_SHAPING_SINGLE)
DEFAULT_QUAD_REWARD_SHAPING['quad_rewards'].update(dict(quadcol_bin=0.0, quadcol_bin_smooth_max=0.0, quadcol_bin_obst=0.0))


class AnnealSchedule:                      # swarm_rl/env_wrappers/quad_utils.py:13-17
    def __init__(self, coeff_name, final_value, anneal_env_steps):
        self.coeff_name = coeff_name
        self.final_value = final_value
        self.anneal_env_steps = anneal_env_steps


class Wrapper:
    """Delegating wrapper with the gymnasium.Wrapper surface the reference relies on."""

    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name.startswith('_'):
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def close(self):
        return self.env.close()


class QuadsRewardShapingWrapper(Wrapper):
    """Reward-coefficient shaping / annealing and per-episode reward statistics (reward_shaping.py:19-123)."""

    def __init__(self, env, reward_shaping_scheme=None, annealing=None, with_pbt=False):
        super().__init__(env)
        self.reward_shaping_scheme = reward_shaping_scheme
        self.cumulative_rewards = None
        self.episode_actions = None
        self.num_agents = env.num_agents if hasattr(env, 'num_agents') else 1
        self.reward_shaping_updated = True
        self.annealing = annealing
        self.training_info = {}                 # TrainingInfoInterface: Sample Factory writes approx_total_training_steps here
        self.with_pbt = with_pbt

    # RewardShapingInterface (used with PBT)
    def get_default_reward_shaping(self):
        return dict(quad_rewards=dict())

    def get_current_reward_shaping(self, agent_idx):
        return dict(quad_rewards=dict())

    def set_reward_shaping(self, reward_shaping, unused_agent_idx):
        self.reward_shaping_scheme = dict(quad_rewards=dict())
        self.reward_shaping_updated = True

    def reset(self):
        obs = self.env.reset()
        self.cumulative_rewards = [dict() for _ in range(self.num_agents)]
        self.episode_actions = []
        return obs

    def step(self, action):
        self.episode_actions.append(action)
        if self.reward_shaping_updated:
            env_reward_shaping = self.env.unwrapped.rew_coeff
            for key, weight in self.reward_shaping_scheme['quad_rewards'].items():
                env_reward_shaping[key] = weight
            self.reward_shaping_updated = False

        obs, rewards, dones, infos = self.env.step(action)
        infos_multi, dones_multi = (infos, dones) if self.env.is_multiagent else ([infos], [dones])

        for i, info in enumerate(infos_multi):
            for key, value in info['rewards'].items():
                if key.startswith('rew'):
                    self.cumulative_rewards[i][key] = self.cumulative_rewards[i].get(key, 0) + value
            if dones_multi[i]:
                true_reward = self.cumulative_rewards[i]['rewraw_main']
                true_reward += 1000 * self.cumulative_rewards[i].get('rewraw_quadcol', 0)      # reward_shaping.py:80-84
                info['true_reward'] = true_reward
                self.cumulative_rewards[i]['rewraw_main'] = true_reward
                extra_stats = info.setdefault('episode_extra_stats', dict())
                extra_stats.update(self.cumulative_rewards[i])
                approx_total_training_steps = self.training_info.get('approx_total_training_steps', 0)
                extra_stats['z_approx_total_training_steps'] = approx_total_training_steps
                scenario = getattr(self.env.unwrapped, 'scenario', None)
                if scenario:
                    scenario_name = scenario.name()
                    for rew_key in ['rew_pos', 'rew_crash']:
                        extra_stats[f'{scenario_name}/{rew_key}'] = self.cumulative_rewards[i][rew_key]
                episode_actions = np.array(self.episode_actions).transpose()
                for action_idx in range(episode_actions.shape[0]):
                    extra_stats[f'z_action{action_idx}_mean'] = np.mean(episode_actions[action_idx])
                    extra_stats[f'z_action{action_idx}_std'] = np.std(episode_actions[action_idx])
                self.cumulative_rewards[i] = dict()
                if self.annealing:
                    env_reward_shaping = self.env.unwrapped.rew_coeff
                    for sched in self.annealing:       # linear from 0 to the final value (reward_shaping.py:110-118)
                        env_reward_shaping[sched.coeff_name] = min(
                            sched.final_value * approx_total_training_steps / sched.anneal_env_steps, sched.final_value)
                        extra_stats[f'z_anneal_{sched.coeff_name}'] = env_reward_shaping[sched.coeff_name]
        if any(dones_multi):
            self.episode_actions = []
        return obs, rewards, dones, infos


class QuadEnvCompatibility(Wrapper):
    """Old 4-tuple API -> gymnasium 5-tuple (compatibility.py:10-57): terminated = dones, truncated all False."""

    def reset(self, seed=None, options=None):
        return self.env.reset(), {}

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        if isinstance(info, dict) and isinstance(done, bool):
            done = [done]
        terminated = np.array(done, dtype=bool)
        truncated = np.zeros_like(terminated)
        return obs, reward, terminated, truncated, info

    def render(self):
        return self.env.render()


def make_quadrotor_env_multi(cfg, render_mode=None, **kwargs):
    """swarm_rl/env_wrappers/quad_utils.py:20-110 with the B200 env underneath.  `cfg` is any object carrying the
    `--quads_*` attributes of quadrotor_params.py:15-121 (argparse.Namespace, SF's AttrDict, ...)."""
    from .env import QuadrotorEnvMulti
    rew_coeff = DEFAULT_QUAD_REWARD_SHAPING['quad_rewards']
    use_replay_buffer = cfg.replay_buffer_sample_prob > 0.0
    if getattr(cfg, 'visualize_v_value', False):
        raise NotImplementedError("V-value visualisation is out of scope")
    env = QuadrotorEnvMulti(
        num_agents=cfg.quads_num_agents, ep_time=cfg.quads_episode_duration, rew_coeff=rew_coeff,
        obs_repr=cfg.quads_obs_repr,
        neighbor_visible_num=cfg.quads_neighbor_visible_num, neighbor_obs_type=cfg.quads_neighbor_obs_type,
        collision_hitbox_radius=cfg.quads_collision_hitbox_radius,
        collision_falloff_radius=cfg.quads_collision_falloff_radius,
        use_obstacles=cfg.quads_use_obstacles, obst_density=cfg.quads_obst_density, obst_size=cfg.quads_obst_size,
        obst_spawn_area=cfg.quads_obst_spawn_area, use_downwash=cfg.quads_use_downwash,
        use_numba=getattr(cfg, 'quads_use_numba', True), quads_mode=cfg.quads_mode, room_dims=cfg.quads_room_dims,
        use_replay_buffer=use_replay_buffer, quads_view_mode=getattr(cfg, 'quads_view_mode', ['topdown']),
        quads_render=getattr(cfg, 'quads_render', False),
        dynamics_params='Crazyflie', raw_control=True, raw_control_zero_middle=True,
        dynamics_randomize_every=None, dynamics_change=dict(noise=dict(thrust_noise_ratio=0.05), damp=dict(vel=0, omega_quadratic=0)),
        dyn_sampler_1=None, sense_noise='default', init_random_state=False, render_mode=render_mode,
        device=getattr(cfg, 'quads_device', 0), seed=getattr(cfg, 'seed', None),
    )
    if use_replay_buffer:                                  # quad_utils.py:67-70
        from .replay import ExperienceReplayWrapper
        env = ExperienceReplayWrapper(env, cfg.replay_buffer_sample_prob, cfg.quads_obst_density, cfg.quads_obst_size,
                                      getattr(cfg, 'quads_domain_random', False), getattr(cfg, 'quads_obst_density_random', False),
                                      getattr(cfg, 'quads_obst_size_random', False), getattr(cfg, 'quads_obst_density_min', 0.),
                                      getattr(cfg, 'quads_obst_density_max', 0.), getattr(cfg, 'quads_obst_size_min', 0.),
                                      getattr(cfg, 'quads_obst_size_max', 0.))
    reward_shaping = copy.deepcopy(DEFAULT_QUAD_REWARD_SHAPING)
    reward_shaping['quad_rewards']['quadcol_bin'] = cfg.quads_collision_reward
    reward_shaping['quad_rewards']['quadcol_bin_smooth_max'] = cfg.quads_collision_smooth_max_penalty
    reward_shaping['quad_rewards']['quadcol_bin_obst'] = cfg.quads_obst_collision_reward
    if cfg.anneal_collision_steps > 0:
        for k in ('quadcol_bin', 'quadcol_bin_smooth_max', 'quadcol_bin_obst'):
            reward_shaping['quad_rewards'][k] = 0.0
        annealing = [
            AnnealSchedule('quadcol_bin', cfg.quads_collision_reward, cfg.anneal_collision_steps),
            AnnealSchedule('quadcol_bin_smooth_max', cfg.quads_collision_smooth_max_penalty, cfg.anneal_collision_steps),
            AnnealSchedule('quadcol_bin_obst', cfg.quads_obst_collision_reward, cfg.anneal_collision_steps),
        ]
    else:
        annealing = None
    env = QuadsRewardShapingWrapper(env, reward_shaping_scheme=reward_shaping, annealing=annealing,
                                    with_pbt=getattr(cfg, 'with_pbt', False))
    return QuadEnvCompatibility(env)


def make_quadrotor_env_multi_batched(cfg, num_envs, env_id_offset=0):
    """The same factory for a batched sampler: `num_envs` envs behind one object, CUDA tensors in and out, the wrappers
    of quad_utils.py:67-110 in their batched form (batched.py).  Episodes, goal events, reward shaping statistics and the
    collision-event replay all stay on the device."""
    from .env import QuadrotorEnvMultiBatched
    from .batched import BatchedRewardShaping, BatchedExperienceReplay
    env = QuadrotorEnvMultiBatched(
        num_envs=num_envs, num_agents=cfg.quads_num_agents, ep_time=cfg.quads_episode_duration,
        rew_coeff=dict(DEFAULT_QUAD_REWARD_SHAPING['quad_rewards']), obs_repr=cfg.quads_obs_repr,
        neighbor_visible_num=cfg.quads_neighbor_visible_num, neighbor_obs_type=cfg.quads_neighbor_obs_type,
        collision_hitbox_radius=cfg.quads_collision_hitbox_radius, collision_falloff_radius=cfg.quads_collision_falloff_radius,
        use_obstacles=cfg.quads_use_obstacles, obst_density=cfg.quads_obst_density, obst_size=cfg.quads_obst_size,
        obst_spawn_area=cfg.quads_obst_spawn_area, use_downwash=cfg.quads_use_downwash, quads_mode=cfg.quads_mode,
        room_dims=cfg.quads_room_dims, device=getattr(cfg, 'quads_device', 0), seed=getattr(cfg, 'seed', None),
        env_id_offset=env_id_offset)
    if cfg.replay_buffer_sample_prob > 0.0:                # quad_utils.py:67-70
        env = BatchedExperienceReplay(env, cfg.replay_buffer_sample_prob, seed=getattr(cfg, 'seed', None) or 0)
    reward_shaping = copy.deepcopy(DEFAULT_QUAD_REWARD_SHAPING)
    reward_shaping['quad_rewards']['quadcol_bin'] = cfg.quads_collision_reward
    reward_shaping['quad_rewards']['quadcol_bin_smooth_max'] = cfg.quads_collision_smooth_max_penalty
    reward_shaping['quad_rewards']['quadcol_bin_obst'] = cfg.quads_obst_collision_reward
    annealing = None
    if cfg.anneal_collision_steps > 0:
        names = {'quadcol_bin': cfg.quads_collision_reward, 'quadcol_bin_smooth_max': cfg.quads_collision_smooth_max_penalty,
                 'quadcol_bin_obst': cfg.quads_obst_collision_reward}
        for k in names:
            reward_shaping['quad_rewards'][k] = 0.0
        annealing = [AnnealSchedule(k, v, cfg.anneal_collision_steps) for k, v in names.items()]
    return BatchedRewardShaping(env, reward_shaping_scheme=reward_shaping, annealing=annealing)


def make_quadrotor_env(env_name, cfg=None, _env_config=None, render_mode=None, **kwargs):
    """Sample Factory env factory, same signature as swarm_rl/env_wrappers/quad_utils.py:113-117."""
    if env_name == 'quadrotor_multi':
        return make_quadrotor_env_multi(cfg, render_mode, **kwargs)
    raise NotImplementedError
