"""Replay a golden reference trajectory through the CPU oracle — TEST INFRASTRUCTURE ONLY.

Used by tests/test_oracle_vs_reference.py: the oracle is driven with the reference's own random
streams (ReplayRng), the recorded actions and planted states, and must reproduce the reference's
observations / rewards / dones / reward terms / states.
"""
import json

import numpy as np

from .quadswarm_oracle import EnvConfig, OracleEnv, ReplayRng, ReferenceEpisodeSource, quad_params_from_constants
from .gen_golden import INFO_KEYS


def config_from_case(kw):
    cfg = EnvConfig(
        num_agents=kw['num_agents'], ep_time=kw.get('ep_time', 15.0), obs_repr=kw.get('obs_repr', 'xyz_vxyz_R_omega'),
        neighbor_visible_num=kw.get('neighbor_visible_num', 6), neighbor_obs_type=kw.get('neighbor_obs_type', 'pos_vel'),
        use_obstacles=kw.get('use_obstacles', False), obst_density=kw.get('obst_density', 0.2),
        obst_size=kw.get('obst_size', 0.6), obst_spawn_area=tuple(kw.get('obst_spawn_area', (8.0, 8.0))),
        use_downwash=kw.get('use_downwash', False), room_dims=tuple(kw.get('room_dims', (10., 10., 10.))))
    rc = kw.get('rew_coeff') or dict(pos=1.0, effort=0.05, spin=0.1, vel=0.0, crash=1.0, orient=1.0, yaw=0.0,
                                     quadcol_bin=5.0, quadcol_bin_smooth_max=10.0, quadcol_bin_obst=5.0)
    cfg.rew_coeff.update(rc)
    return cfg


def replay_golden(g, make_scenario):
    """g: dict-like loaded golden npz.  make_scenario(mode, cfg, rng) -> host scenario object under test."""
    case = json.loads(str(g['case_json']))
    kw, T, seed = case['kw'], case['T'], case['seed']
    cfg = config_from_case(kw)
    n = cfg.num_agents
    rng = ReplayRng(seed, seed + 1, [seed + 100 + i for i in range(n)])
    scenario = make_scenario(kw.get('quads_mode', 'static_same_goal'), cfg, rng.py)
    params, dyn_t = None, None
    if 'dyn_rows' in g and kw.get('dynamics_params', 'Crazyflie') != 'Crazyflie':
        from quad_swarm_rl_b200.quad_models import DYN_FIELDS
        dyn_t = [int(t) for t in g['dyn_t']]
        to_params = lambda rows: [quad_params_from_constants(dict(zip(DYN_FIELDS, r))) for r in rows]
        params = to_params(g['dyn_rows'][0])
    env = OracleEnv(cfg, rng, ReferenceEpisodeSource(scenario), params=params)
    if params is not None:
        env.env_arm = float(g['env_arm'])
        env.collision_threshold = cfg.collision_hitbox_radius * env.env_arm
        env.collision_falloff_threshold = cfg.collision_falloff_radius * env.env_arm
        if kw.get('dynamics_randomize_every'):
            # The reference resamples inside each drone's _reset when (traj_count + 1) % every == 0
            # (quadrotor_single.py:389-390), drawing from numpy's global stream.  The product's samplers
            # (quad_swarm_rl_b200/quad_models.py) run here on that same stream, in the same place: they must reproduce the
            # constants the reference derived (rows of the fixture) AND leave the stream where the reference left it.
            from quad_swarm_rl_b200.quad_models import DynamicsSource, derive_constants
            every = int(kw['dynamics_randomize_every'])
            change = kw.get('dynamics_change') or dict(noise=dict(thrust_noise_ratio=0.05), damp=dict(vel=0, omega_quadratic=0))
            srcs = [DynamicsSource(kw['dynamics_params'], change, kw.get('dyn_sampler_1'), rs=np.random.RandomState(0)) for _ in range(n)]
            for s_ in srcs:
                s_.rs = rng.py
            resets = [0] * n
            env.dyn_checked = 0

            def dyn_source(i):
                k = resets[i]
                resets[i] += 1
                if (k + 1) % every != 0:
                    return None
                c = derive_constants(srcs[i].sample())
                ref = dict(zip(DYN_FIELDS, g['dyn_rows'][k][i]))
                for key in DYN_FIELDS:
                    np.testing.assert_allclose(c[key], ref[key], rtol=1e-12, atol=1e-300, err_msg=f'resample {k} drone {i} {key}')
                env.dyn_checked += 1
                return quad_params_from_constants(c)
            env.dyn_source = dyn_source
    out = dict(obs0=env.reset())
    rewards = np.zeros((T, n))
    dones = np.zeros((T, n), dtype=bool)
    infos = np.full((T, n, len(INFO_KEYS)), np.nan)
    goals = np.zeros((T, n, 3))
    obs = {}
    states = {k: {} for k in ('pos', 'vel', 'rot', 'omega', 'thrust_rot_damp', 'thrust_cmds_damp', 'ou', 'on_floor')}
    ep_stats = []
    plant_t = g['plant_t']
    want_obs = set(int(t) for t in g['obs_t'])
    want_state = set(int(t) for t in g['state_t'])
    for t in range(T):
        for k in np.where(plant_t == t)[0]:
            d = env.drones[int(g['plant_i'][k])]
            d.pos, d.vel = g['plant_pos'][k].copy(), g['plant_vel'][k].copy()
            # set_state stores omega as float32 (quadrotor_dynamics.py:188)
            d.rot, d.omega = g['plant_rot'][k].copy(), g['plant_omega'][k].astype(np.float32).astype(np.float64)
            d.acc = np.zeros(3)                      # set_state, quadrotor_dynamics.py:185-186
            d.accelerometer = np.array([0., 0., 9.81])
        o, r, dn, inf = env.step(g['actions'][t])
        if t in want_obs:
            obs[t] = o
        rewards[t] = r
        dones[t] = dn
        for i in range(n):
            for k, key in enumerate(INFO_KEYS):
                if key in inf[i]['rewards']:
                    infos[t, i, k] = inf[i]['rewards'][key]
        goals[t] = np.array([d.goal for d in env.drones])
        if t in want_state:
            for k in states:
                states[k][t] = np.array([getattr(d, k) for d in env.drones])
        if dn[0]:
            ep_stats.append((t, {k: float(v) for k, v in inf[0]['episode_extra_stats'].items()}))
    out.update(rewards=rewards, dones=dones, infos=infos, goals=goals,
               obs=np.array([obs[int(t)] for t in g['obs_t']]), ep_stats=ep_stats,
               **{'state_' + k: np.array([states[k][int(t)] for t in g['state_t']]) for k in states})
    return out, env
