"""GPU tests of the training wrappers as kernels (csrc/qs_wrap.cuh through training.BatchedTrainingEnv): reward shaping /
episode statistics and the collision-event replay, checked against sums and states taken
independently from a second, unwrapped engine stepping the same envs (same seeds -> same trajectories)."""
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _env(**kw):
    from quad_swarm_rl_b200.env import QuadrotorEnvMultiBatched
    base = dict(num_envs=24, num_agents=8, ep_time=0.4, neighbor_visible_num=6, quads_mode='static_same_goal', seed=4)
    base.update(kw)
    return QuadrotorEnvMultiBatched(**base)


def test_reward_shaping_statistics_and_annealing():
    """Cumulative reward terms, true_reward, action statistics, the latched episode statistics and the annealed
    coefficient (reward_shaping.py:52-123) out of the wrapper kernel — against sums taken from a twin engine that steps
    the same envs without the wrappers."""
    from quad_swarm_rl_b200.training import BatchedTrainingEnv
    from quad_swarm_rl_b200.wrappers import AnnealSchedule
    env, twin = _env(), _env()
    scheme = dict(quad_rewards=dict(quadcol_bin=0.0, pos=1.0))
    w = BatchedTrainingEnv(env, reward_shaping_scheme=scheme, annealing=[AnnealSchedule('quadcol_bin', 5.0, 1000.0)], stats_every=1 << 30)
    w.training_info['approx_total_training_steps'] = 400
    w.reset(); twin.reset()
    twin.engine.rew_coeff.update(scheme['quad_rewards'])
    E, N = 24, 8
    raw_sum = torch.zeros((E, N, 8), device='cuda')
    acts = []
    g = torch.Generator(device='cuda'); g.manual_seed(1)
    for t in range(41):
        a = torch.rand((E * N, 4), device='cuda', generator=g) * 2 - 1
        obs, rew, term, trunc, infos = w.step(a)
        o2, r2, t2, _, _ = twin.step(a, with_terms=True)
        assert torch.equal(obs, o2) and torch.equal(rew, r2) and torch.equal(term, t2)      # the wrappers do not perturb the env
        raw_sum += twin.engine.rew_terms
        acts.append(a.view(E, N, 4))
        assert infos == {}
        if term.any():
            break
    assert t == 40 and term.all()
    fin = w.flush_stats()
    st = fin['episode_extra_stats']
    assert fin['episodes_finished'] == E
    assert env.engine.rew_coeff['quadcol_bin'] == pytest.approx(5.0 * 400 / 1000.0)          # annealed when the stats are read
    assert st['z_anneal_quadcol_bin'] == pytest.approx(2.0) and st['z_approx_total_training_steps'] == 400
    np.testing.assert_allclose(st['rewraw_pos'], raw_sum[..., 0].mean().item(), rtol=1e-5)
    np.testing.assert_allclose(st['rew_crash'], raw_sum[..., 2].mean().item() * 1.0, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(st['rew_quadcol'], 0.0, atol=1e-7)                           # coefficient 0 during this episode
    np.testing.assert_allclose(st['rew_proximity'], raw_sum[..., 6].mean().item(), rtol=1e-4, atol=1e-7)
    true_reward = raw_sum[..., 0] + 1000.0 * raw_sum[..., 5]
    assert torch.allclose(fin['true_reward'], true_reward, rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(st['rewraw_main'], true_reward.mean().item(), rtol=1e-4)
    A = torch.stack(acts)
    np.testing.assert_allclose(st['z_action2_mean'], A[..., 2].mean().item(), atol=1e-5)
    # std over all agents and steps of an env's episode jointly (reward_shaping.py:100-106), then the mean over the envs
    joint = A[..., 1].permute(1, 0, 2).reshape(E, -1).std(dim=1, unbiased=False).mean().item()
    np.testing.assert_allclose(st['z_action1_std'], joint, rtol=1e-4)
    es, ags = twin.engine.episode_stats()
    np.testing.assert_allclose(st['num_collisions'], es[:, 0].float().mean().item(), rtol=1e-6)
    np.testing.assert_allclose(st['distance_to_goal_1s'], ags[..., 0].mean().item(), rtol=1e-5)
    np.testing.assert_allclose(st['static_same_goal/num_collisions'], es[:, 1].float().mean().item(), rtol=1e-6)
    assert 'Scenario_static_same_goal/rew_pos' in st and 0.0 <= st['metric/agent_col_rate'] <= 1.0
    assert w.flush_stats() == {}                                                            # nothing finished since
    env.close(); twin.close()


def test_experience_replay_stores_and_replays_collision_events():
    """quad_experience_replay.py semantics per env, in the wrapper kernel: checkpoints every 0.5 s, the one from 1.5 s before
    a collision goes into the env's buffer, finished envs restart from a buffered event (p = 1 here) with the stored
    observation and state, collision counters zeroed, and the replayed episode ends early by the snapshot's tick."""
    from quad_swarm_rl_b200.training import BatchedTrainingEnv
    env = _env(num_envs=32, ep_time=3.0, seed=7)
    w = BatchedTrainingEnv(env, replay_buffer_sample_prob=1.0, replay_always_active=True, stats_every=1 << 30)
    w.reset()
    E, N = 32, 8
    g = torch.Generator(device='cuda'); g.manual_seed(2)
    hover = torch.zeros((E * N, 4), device='cuda') + 0.05

    def act():
        return hover + 0.3 * (torch.rand((E * N, 4), device='cuda', generator=g) * 2 - 1)

    planted = torch.arange(E, device='cuda') % 2 == 0
    obs_at, state_at = {}, {}
    first_done = None
    for t in range(301):
        if t == 200:                                         # plant a collision: drone 1 onto drone 0 in every other env
            st = env.engine.get_state()
            st['agent_f32'][planted, 1, 0:3] = st['agent_f32'][planted, 0, 0:3] + 0.01
            env.engine.set_state(st, env_mask=planted)
        obs, rew, term, trunc, infos = w.step(act())
        if t + 1 in (50, 100, 150, 200, 250):                # tick after this step: the checkpoints the kernel takes
            obs_at[t + 1] = obs.view(E, N, -1).clone()
            state_at[t + 1] = {k: v.clone() for k, v in env.engine.get_state().items() if v is not None}
        if term.any():
            first_done = t
            break
    assert first_done == 300 and term.all()
    agg = env.engine.wrap_read(reset=False)
    from quad_swarm_rl_b200 import _lib as L
    n_planted = int(planted.sum())
    assert agg[L.WA['EVENTS_STORED']] >= n_planted             # the planted collision was stored (one event per 5 s per env)
    stc = env.engine.get_state()
    ticks = stc['env_i32'][:, 0]
    replayed = ticks > 0
    assert replayed[planted].all()                           # p = 1: every env with an event replays it
    assert agg[L.WA['REPLAYED_EVENTS']] == int(replayed.sum())
    # collision at tick 201: checkpoints 50, 100, 150, 200 exist; three back from the newest is tick 100
    assert (ticks[planted] == 100).all() and ((ticks[replayed] % 50) == 0).all()
    ridx = torch.nonzero(planted).flatten()
    assert torch.equal(stc['agent_f32'][ridx], state_at[100]['agent_f32'][ridx])
    assert torch.equal(obs.view(E, N, -1)[ridx], obs_at[100][ridx])
    assert (stc['env_i32'][ridx, 4] == 0).all()                                      # collision counters zeroed for the replay
    assert torch.equal(stc['env_i32'][ridx, 6:11], state_at[100]['env_i32'][ridx, 6:11])
    assert (stc['env_i32'][ridx, 1] > state_at[100]['env_i32'][ridx, 1]).all()       # the RNG step counter is NOT rewound
    # replayed envs finish ep_len + 1 - tick steps later, fresh ones after a full episode
    fresh = torch.nonzero(~replayed).flatten()
    remaining = (300 + 1 - ticks[replayed]).tolist()
    seen = {}
    for t in range(1, 302):
        obs, rew, term, trunc, infos = w.step(act())
        d = term.view(E, N)[:, 0]
        for e in torch.nonzero(d).flatten().tolist():
            seen.setdefault(e, t)
        if len(seen) == E:
            break
    for e, r in zip(torch.nonzero(replayed).flatten().tolist(), remaining):
        assert seen[e] == r, (e, seen[e], r)
    for e in fresh.tolist():
        assert seen[e] == 301
    fin = w.flush_stats()
    st = fin['episode_extra_stats']
    assert st['replay/replay_rate'] > 0 and 'num_collisions_replay' in st
    env.close()


def test_factory_with_the_obstacle_baseline_flags():
    """make_quadrotor_env_multi_batched with the flags of swarm_rl/runs/obstacles/quad_obstacle_baseline.py:12-21
    (mix with pillars, replay on, collision-reward annealing): everything stays on the device and statistics come out."""
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
    assert stats is not None and term.all()                    # stats_every = one episode: reported on the terminal step
    assert env.engine.rew_coeff['quadcol_bin'] == pytest.approx(2.5) and env.engine.rew_coeff['quadcol_bin_obst'] == pytest.approx(2.5)
    assert {'rew_pos', 'rewraw_quadcol_obstacle', 'num_collisions_obst_quad', 'metric/agent_success_rate',
            'z_anneal_quadcol_bin_smooth_max'} <= set(stats)
    assert any(k.startswith('Scenario_o_random/') for k in stats) and any(k.startswith('Scenario_o_static_same_goal/') for k in stats)
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
    env.close()


def test_single_env_factory_object_protocol():
    """make_quadrotor_env (the Sample Factory entry point, swarm_rl/env_wrappers/quad_utils.py:113-117) for one env: numpy in
    and out, lists per agent, 5-tuple; the terminal step carries true_reward and episode_extra_stats for every agent."""
    from quad_swarm_rl_b200.wrappers import make_quadrotor_env
    cfg = types.SimpleNamespace(
        quads_num_agents=4, quads_episode_duration=0.2, quads_obs_repr='xyz_vxyz_R_omega', quads_neighbor_visible_num=2,
        quads_neighbor_obs_type='pos_vel', quads_collision_hitbox_radius=2.0, quads_collision_falloff_radius=4.0,
        quads_use_obstacles=False, quads_obst_density=0.2, quads_obst_size=0.6, quads_obst_spawn_area=[8.0, 8.0],
        quads_use_downwash=False, quads_mode='mix', quads_room_dims=[10., 10., 10.], replay_buffer_sample_prob=0.75,
        quads_collision_reward=5.0, quads_collision_smooth_max_penalty=10.0, quads_obst_collision_reward=0.0,
        anneal_collision_steps=0.0, seed=5)
    env = make_quadrotor_env('quadrotor_multi', cfg=cfg)
    obs, info = env.reset()
    assert obs.shape == (4, 30) and info == {} and env.num_agents == 4 and env.is_multiagent
    rs = np.random.RandomState(0)
    for t in range(21):
        obs, rew, term, trunc, infos = env.step(rs.uniform(-1, 1, (4, 4)).astype(np.float32))
        assert isinstance(rew, list) and len(infos) == 4 and obs.dtype == np.float64
    assert term.all() and not trunc.any()
    assert all('true_reward' in i and 'rew_pos' in i['episode_extra_stats'] for i in infos)
    assert any(k.startswith('Scenario_') for k in infos[0]['episode_extra_stats'])
    env.close()


# ---- the wrapper kernel against the reference's own wrappers (fixtures: tests/golden/wrappers.npz, oracle/gen_golden_wrappers.py) ----
@pytest.mark.parametrize('case', [0, 1])
def test_reward_shaping_kernel_reproduces_the_reference_wrapper(case, golden_dir):
    """The UNMODIFIED QuadsRewardShapingWrapper was driven by a stand-in env with synthetic per-step reward dicts, actions
    and dones; the same per-step inputs go through qs_wrap_apply here.  Per episode: every agent's true_reward, and the
    episode_extra_stats keys as means over the agents (cumulative rew_* / rewraw_*, the joint action mean / std, the
    per-scenario copies), with the coefficients (shaping scheme at the first step, annealing at episode ends) in force."""
    import json, os
    from quad_swarm_rl_b200 import _lib as L
    from quad_swarm_rl_b200.engine import QuadSwarmEngine
    from quad_swarm_rl_b200.training import stats_dict
    g = np.load(os.path.join(golden_dir, 'wrappers.npz'))
    terms, actions, dones, coeffs = (g[f'shaping{case}_{k}'] for k in ('terms', 'actions', 'dones', 'coeffs'))
    episodes = json.loads(str(g[f'shaping{case}_episodes']))
    T, N = terms.shape[:2]
    eng = QuadSwarmEngine(num_envs=1, num_agents=N, neighbor_visible_num=min(2, N - 2), use_obstacles=True, obst_density=0.2, seed=1)
    eng.wrap_enable(use_replay=False)
    dev = eng.device
    ep_iter = iter(episodes)
    for t in range(T):
        for k, name in enumerate(L.REW_KEYS):
            eng.rew_coeff[name] = float(coeffs[t, k])
        eng.wrap_apply(torch.tensor(actions[t], dtype=torch.float32, device=dev).view(1, N, 4).contiguous(),
                       torch.tensor(terms[t], dtype=torch.float32, device=dev).view(1, N, 8).contiguous(),
                       torch.full((1, N), int(dones[t]), dtype=torch.uint8, device=dev))
        if dones[t]:
            ref = next(ep_iter)
            assert ref['t'] == t
            agg = eng.wrap_read(reset=True)
            st = stats_dict(agg, use_obstacles=True, fallback_scenario='static_same_goal')
            tr = eng.wrap_true_reward().cpu().numpy().reshape(-1)
            np.testing.assert_allclose(tr, ref['true_reward'], rtol=2e-5, atol=1e-5)
            mean = lambda key: float(np.mean([s[key] for s in ref['stats']]))
            for key in ('rewraw_main', 'rewraw_pos', 'rew_pos', 'rew_main', 'rewraw_action', 'rew_action', 'rewraw_crash', 'rew_crash',
                        'rewraw_orient', 'rew_orient', 'rewraw_spin', 'rew_spin', 'rewraw_quadcol', 'rew_quadcol', 'rew_proximity',
                        'rewraw_quadcol_obstacle', 'rew_quadcol_obstacle', 'z_action0_mean', 'z_action3_mean', 'z_action1_std', 'z_action2_std'):
                np.testing.assert_allclose(st[key], mean(key), rtol=3e-5, atol=2e-6, err_msg=f'{key} episode ending at {t}')
            np.testing.assert_allclose(st['Scenario_static_same_goal/rew_pos'], mean('Scenario_static_same_goal/rew_pos'), rtol=3e-5)
            np.testing.assert_allclose(st['Scenario_static_same_goal/rew_crash'], mean('Scenario_static_same_goal/rew_crash'), rtol=3e-5, atol=2e-6)
    eng.close()


def test_replay_kernel_follows_the_reference_wrapper_schedule(golden_dir):
    """The UNMODIFIED ExperienceReplayWrapper on a stand-in env with collisions at ticks 120 (grace period), 201, 260 (inside
    the 5 s cooldown) and 720 stored the checkpoints of ticks 100 and 600 and started every later episode from one of them
    (p = 1).  The same collisions are planted in a real env here: same stored checkpoints, same start ticks."""
    import json, os
    from quad_swarm_rl_b200 import _lib as L
    from quad_swarm_rl_b200.training import BatchedTrainingEnv
    ref = json.loads(str(np.load(os.path.join(golden_dir, 'wrappers.npz'))['replay']))
    assert ref['stored_checkpoint_ticks'] == [100, 600] and set(ref['episode_start_ticks']) == {100, 600}
    env = _env(num_envs=2, num_agents=2, ep_time=9.0, neighbor_visible_num=0, quads_mode='static_diff_goal', seed=11)
    w = BatchedTrainingEnv(env, replay_buffer_sample_prob=1.0, replay_always_active=True, stats_every=1 << 30)
    w.reset()
    g = torch.Generator(device='cuda'); g.manual_seed(5)
    hover = torch.zeros((4, 4), device='cuda') + 0.05
    starts = []
    for t in range(1, 4600):
        tick_next = int(env.engine.get_state()['env_i32'][0, 0]) + 1
        fresh = len(starts) == 0
        if fresh and tick_next in ref['collision_ticks']:          # plant: drone 1 of env 0 onto drone 0, as the stand-in env did
            st = env.engine.get_state()
            st['agent_f32'][0, 1, 0:3] = st['agent_f32'][0, 0, 0:3] + 0.01
            st['agent_f32'][0, :, 3:6] = 0
            env.engine.set_state(st, env_mask=torch.tensor([1, 0], dtype=torch.uint8, device='cuda'))
        elif fresh and tick_next - 1 in ref['collision_ticks']:    # and apart again, so that the next planted contact is a new one
            st = env.engine.get_state()
            st['agent_f32'][0, 1, 0:3] = st['agent_f32'][0, 0, 0:3] + torch.tensor([0.6, 0.0, 0.0], device='cuda')
            st['agent_f32'][0, :, 3:6] = 0
            env.engine.set_state(st, env_mask=torch.tensor([1, 0], dtype=torch.uint8, device='cuda'))
        obs, rew, term, trunc, infos = w.step(hover + 0.02 * (torch.rand((4, 4), device='cuda', generator=g) * 2 - 1))
        if bool(term[0]):
            starts.append(int(env.engine.get_state()['env_i32'][0, 0]))
            if len(starts) == 1:
                agg = env.engine.wrap_read(reset=False)
                assert agg[L.WA['EVENTS_STORED']] == 2, agg[L.WA['EVENTS_STORED']]        # 201 and 720; 120 and 260 ignored
            if len(starts) == len(ref['episode_start_ticks']):
                break
    assert len(starts) == len(ref["episode_start_ticks"]) and set(starts) <= {100, 600}, starts
    env.close()


def test_block_chained_wrapped_steps_equal_synchronised_ones():
    """qs_wrap_step on a chained handle with balanced step grids: no grid-wide barrier is left — the wrapper kernel's block b
    takes the `done` word of step block b and hands the block to the next step grid (csrc/qs_wrap.cuh, qs_wrap_kernel).  A
    CUDA graph of such control steps (replay on, clustered spawns so that collision events are stored and replayed) must
    leave what the same launches with a device synchronisation after every one leave: state, outputs, aggregate."""
    from quad_swarm_rl_b200.engine import QuadSwarmEngine
    kw = dict(num_agents=8, neighbor_visible_num=2, obs_repr='xyz_vxyz_R_omega_floor', use_obstacles=True, use_downwash=True,
              collision_hitbox_radius=30.0, collision_falloff_radius=32.0)      # neighbouring spawn cells touch: events for sure
    E, T = 600, 260
    mk = lambda: QuadSwarmEngine(num_envs=E, seed=12, ep_time=2.0, device_scenario='o_static_same_goal', **kw)
    e1, e2 = mk(), mk()
    g = torch.Generator(device='cuda'); g.manual_seed(3)
    a = (0.05 + 0.4 * (torch.rand((T, E, 8, 4), device='cuda', generator=g) * 2 - 1)).contiguous()
    for e in (e1, e2):
        e.wrap_enable(use_replay=True, replay_buffer_size=4, replay_prob=0.9, replay_always_active=True)
        e.set_chained(True)
        e.reset()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())     # resets / actions above were enqueued on the default stream
    with torch.cuda.stream(st):
        e1.wrap_step(a[0])
        st.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st):
            for t in range(1, T):
                e1.wrap_step(a[t])
        for _ in range(3):
            gr.replay()
        st.synchronize()
    e2.wrap_step(a[0])
    for _ in range(3):
        for t in range(1, T):
            e2.wrap_step(a[t])
            torch.cuda.synchronize()
    assert torch.equal(e1.obs, e2.obs) and torch.equal(e1.rewards, e2.rewards) and torch.equal(e1.dones, e2.dones)
    s1, s2 = e1.get_state(), e2.get_state()
    for k in ('agent_f32', 'agent_u32', 'env_i32'):
        assert torch.equal(s1[k], s2[k]), k
    g1, g2 = e1.wrap_read(), e2.wrap_read()
    from quad_swarm_rl_b200 import _lib as L
    assert g1[L.WA['EPISODES_TOTAL']] == g2[L.WA['EPISODES_TOTAL']] > 0
    assert g1[L.WA['EVENTS_STORED']] == g2[L.WA['EVENTS_STORED']] and g1[L.WA['REPLAYED_EVENTS']] == g2[L.WA['REPLAYED_EVENTS']] > 0
    np.testing.assert_allclose(g1, g2, rtol=2e-4, atol=1e-3)             # float atomics: the order of the additions differs
    assert e1.handover_timeouts == 0 and e2.handover_timeouts == 0
    e1.close(); e2.close()
