"""GPU tests of the batched wrappers (quad_swarm_rl_b200/batched.py): reward shaping / episode statistics and the
collision-event replay, both kept on the device."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _env(**kw):
    from quad_swarm_rl_b200.env import QuadrotorEnvMultiBatched
    base = dict(num_envs=24, num_agents=8, ep_time=0.4, neighbor_visible_num=6, quads_mode='static_same_goal', seed=4)
    base.update(kw)
    return QuadrotorEnvMultiBatched(**base)


def test_batched_reward_shaping_statistics_and_annealing():
    """Cumulative reward terms, true_reward, action statistics, the latched episode statistics and the annealed
    coefficient (reward_shaping.py:52-123) — checked against sums taken independently from the engine's outputs."""
    from quad_swarm_rl_b200.batched import BatchedRewardShaping
    from quad_swarm_rl_b200.wrappers import AnnealSchedule
    env = _env()
    w = BatchedRewardShaping(env, reward_shaping_scheme=dict(quad_rewards=dict(quadcol_bin=0.0, pos=1.0)),
                             annealing=[AnnealSchedule('quadcol_bin', 5.0, 1000.0)])
    w.training_info['approx_total_training_steps'] = 400
    w.reset()
    E, N = 24, 8
    raw_sum = torch.zeros((E, N, 8), device='cuda')
    rew_sum = torch.zeros((E, N), device='cuda')
    acts = []
    infos = {}
    g = torch.Generator(device='cuda'); g.manual_seed(1)
    for t in range(41):
        a = torch.rand((E * N, 4), device='cuda', generator=g) * 2 - 1
        obs, rew, term, trunc, infos = w.step(a)
        raw_sum += env.engine.rew_terms
        rew_sum += rew.view(E, N)
        acts.append(a.view(E, N, 4))
        if term.any():
            break
    assert t == 40 and term.all()
    st = infos['episode_extra_stats']
    assert env.engine.rew_coeff['quadcol_bin'] == pytest.approx(5.0 * 400 / 1000.0)          # annealed at the episode end
    assert st['z_anneal_quadcol_bin'] == pytest.approx(2.0) and st['z_approx_total_training_steps'] == 400
    np.testing.assert_allclose(st['rewraw_pos'], raw_sum[..., 0].mean().item(), rtol=1e-5)
    np.testing.assert_allclose(st['rew_crash'], raw_sum[..., 2].mean().item() * 1.0, rtol=1e-5)
    np.testing.assert_allclose(st['rew_quadcol'], 0.0, atol=1e-7)                           # coefficient 0 during this episode
    true_reward = raw_sum[..., 0] + 1000.0 * raw_sum[..., 5]
    assert torch.allclose(infos['true_reward'], true_reward, rtol=1e-5, atol=1e-4)
    # the env's reward is the weighted sum of the terms: pos + effort + crash + orient + spin + quadcol + proximity
    c = env.engine.rew_coeff
    A = torch.stack(acts)
    np.testing.assert_allclose(st['z_action2_mean'], A[..., 2].mean().item(), atol=1e-5)
    np.testing.assert_allclose(st['z_action1_std'], A[..., 1].std(dim=0, unbiased=False).mean().item(), rtol=1e-4)
    es, ags = env.engine.episode_stats()
    np.testing.assert_allclose(st['num_collisions'], es[:, 0].float().mean().item(), rtol=1e-6)
    np.testing.assert_allclose(st['distance_to_goal_1s'], ags[..., 0].mean().item(), rtol=1e-6)
    assert 'Scenario_static_same_goal/rew_pos' in st and 'static_same_goal/num_collisions' in st
    assert 0.0 <= st['metric/agent_col_rate'] <= 1.0
    env.close()


def test_batched_experience_replay_stores_and_replays_collision_events():
    """quad_experience_replay.py semantics per env, on the device: checkpoints every 0.5 s, the one from 1.5 s before a
    collision goes into the env's buffer, finished envs restart from a buffered event (p = 1 here), the returned
    observation and the restored state are the stored ones, and replayed episodes end early by the snapshot's tick."""
    from quad_swarm_rl_b200.batched import BatchedExperienceReplay
    env = _env(num_envs=32, ep_time=3.0, seed=7)
    rp = BatchedExperienceReplay(env, replay_buffer_sample_prob=1.0, always_active=True, seed=3)
    rp.reset()
    E, N = 32, 8
    g = torch.Generator(device='cuda'); g.manual_seed(2)
    hover = torch.zeros((E * N, 4), device='cuda') + 0.05

    def act():
        return hover + 0.3 * (torch.rand((E * N, 4), device='cuda', generator=g) * 2 - 1)

    first_done = None
    planted = torch.arange(E, device='cuda') % 2 == 0
    for t in range(301):
        if t == 200:                                         # plant a collision: drone 1 onto drone 0 in every other env
            st = env.engine.get_state()
            st['agent_f32'][planted, 1, 0:3] = st['agent_f32'][planted, 0, 0:3] + 0.01
            env.engine.set_state(st, env_mask=planted)
        obs, rew, term, trunc, infos = rp.step(act())
        if term.any():
            first_done = t
            break
    assert first_done == 300 and term.all()
    stored = rp.buf_valid.sum(dim=0)
    assert (stored[planted] == 1).all()                      # the planted collision was stored, once (one event per 5 s)
    assert stored.max().item() <= 1
    replayed = stored >= 1                                   # p = 1: every env with an event replays it
    # checkpoints existed for ticks 50, 100, 150, 200; the one 3 checkpoints back from the collision at tick 201 is tick 100
    assert (rp.buf['env_i32'][0, planted, 0] == 100).all()
    assert rp.replayed_events == int(replayed.sum())
    ridx = torch.nonzero(replayed).flatten()
    fresh = torch.nonzero(~replayed).flatten()
    ticks = rp.tick[ridx]
    assert ((ticks % 50) == 0).all() and (ticks >= 50).all() and (rp.tick[fresh] == 0).all()
    st = env.engine.get_state()
    assert torch.equal(st['env_i32'][ridx, 0].long(), ticks)
    assert torch.equal(st['agent_f32'][ridx], rp.buf['agent_f32'][0, ridx])          # slot 0 = first free slot
    assert torch.equal(obs.view(E, N, -1)[ridx], rp.buf_obs[0, ridx])
    assert (st['env_i32'][ridx, 4] == 0).all()                                       # collision counters zeroed for the replay
    assert rp.saved[ridx].all() and not rp.saved[fresh].any()
    # replayed envs finish ep_len + 1 - tick steps later, fresh ones after a full episode
    remaining = (300 + 1 - ticks).tolist()
    seen = {}
    for t in range(1, 302):
        obs, rew, term, trunc, infos = rp.step(act())
        d = term.view(E, N)[:, 0]
        for e in torch.nonzero(d).flatten().tolist():
            seen.setdefault(e, t)
        if len(seen) == E:
            break
    for e, r in zip(ridx.tolist(), remaining):
        assert seen[e] == r, (e, seen[e], r)
    for e in fresh.tolist():
        assert seen[e] == 301
    assert 'replay' in infos and infos['replay']['replay/replay_rate'] > 0
    env.close()


def test_batched_factory_with_the_obstacle_baseline_flags():
    """make_quadrotor_env_multi_batched with the flags of swarm_rl/runs/obstacles/quad_obstacle_baseline.py:12-21
    (mix with pillars, replay on, collision-reward annealing): everything stays on the device and statistics come out."""
    import types
    from quad_swarm_rl_b200.wrappers import make_quadrotor_env_multi_batched
    cfg = types.SimpleNamespace(
        quads_num_agents=8, quads_episode_duration=0.3, quads_obs_repr='xyz_vxyz_R_omega_floor', quads_neighbor_visible_num=2,
        quads_neighbor_obs_type='pos_vel', quads_collision_hitbox_radius=2.0, quads_collision_falloff_radius=4.0,
        quads_use_obstacles=True, quads_obst_density=0.2, quads_obst_size=0.6, quads_obst_spawn_area=[8.0, 8.0],
        quads_use_downwash=True, quads_mode='mix', quads_room_dims=[10., 10., 10.], replay_buffer_sample_prob=0.75,
        quads_collision_reward=5.0, quads_collision_smooth_max_penalty=4.0, quads_obst_collision_reward=5.0,
        anneal_collision_steps=300000000.0, seed=3)
    env = make_quadrotor_env_multi_batched(cfg, num_envs=64)
    assert env.device_scenario == 'mix' and env.engine.M == 12 and env.num_agents == 512
    env.training_info['approx_total_training_steps'] = 150000000
    obs, _ = env.reset()
    assert obs.shape == (512, 40)
    stats = None
    for t in range(31):
        obs, rew, term, trunc, infos = env.step(torch.rand((512, 4), device='cuda') * 2 - 1)
        if 'episode_extra_stats' in infos:
            stats = infos['episode_extra_stats']
    assert stats is not None and term.all()
    assert env.engine.rew_coeff['quadcol_bin'] == pytest.approx(2.5) and env.engine.rew_coeff['quadcol_bin_obst'] == pytest.approx(2.5)
    assert {'rew_pos', 'rewraw_quadcol_obstacle', 'num_collisions_obst_quad', 'metric/agent_success_rate',
            'z_anneal_quadcol_bin_smooth_max'} <= set(stats)
    assert any(k.startswith('Scenario_o_random/') for k in stats) and any(k.startswith('Scenario_o_static_same_goal/') for k in stats)
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
    env.close()
