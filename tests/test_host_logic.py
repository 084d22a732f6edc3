"""Host-side logic that needs no GPU: episode generators, wrappers (reward shaping / annealing / 5-tuple), spaces,
the SVD period of the reference's float accumulator, and the multi-GPU sharding helpers (world_size-2 gloo)."""
import os
import subprocess
import sys
import types

import numpy as np
import pytest

from quad_swarm_rl_b200 import scenarios as sc
from quad_swarm_rl_b200 import wrappers as wr
from quad_swarm_rl_b200.spaces import make_observation_space, make_action_space
from quad_swarm_rl_b200.sharding import shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('formation', sc.FORMATIONS)
@pytest.mark.parametrize('n', [1, 8, 9, 32])
def test_formation_goals_shape_and_center(formation, n):
    g = sc.formation_goals(formation, n, 0.5, (1.0, -2.0, 3.0), 0.4, 50 if formation.startswith('grid') else 8)
    rows = max(n, 3) if formation == 'sphere' else n          # the reference's sphere generator never makes fewer than 3 points
    assert g.shape == (rows, 3) and np.all(np.isfinite(g))
    if formation.startswith(('grid', 'cube')):
        np.testing.assert_allclose(g.mean(axis=0), [1.0, -2.0, 3.0], atol=1e-9)


def test_o_random_uses_distinct_free_cells():
    rs = np.random.RandomState(3)
    s = sc.create_scenario('o_random', 8, rng=rs, use_obstacles=True)
    for _ in range(20):
        obst_map, pos, cells = sc.obstacle_map_given_density(rs, (8.0, 8.0), 0.2)
        assert obst_map.sum() == 12 and len(pos) == 12
        s.reset(obst_map=obst_map, cell_centers=cells)
        pillars = {tuple(p[:2]) for p in pos}
        for pts in (s.spawn_points, s.goals):
            xy = [tuple(p[:2]) for p in pts]
            assert len(set(xy)) == 8 and not (set(xy) & pillars)
            assert np.all(pts[:, 2] >= 1.0) and np.all(pts[:, 2] <= 3.0)
    assert s.name() == 'Scenario_o_random' and not s.dynamic


def test_swarm_vs_swarm_swaps_centres():
    s = sc.create_scenario('swarm_vs_swarm', 8, rng=np.random.RandomState(1))
    s.reset()
    c1, c2, period = s.c1.copy(), s.c2.copy(), s.period
    assert 400 <= period <= 600
    for t in range(1, period):
        s.step(t)
    assert np.array_equal(s.c1, c1)
    s.step(period)
    assert np.array_equal(s.c1, c2) and np.array_equal(s.c2, c1) and s.goals.shape == (8, 3)


def test_unknown_scenario_raises_like_the_reference():
    with pytest.raises(NameError):
        sc.create_scenario('o_diagonal', 8)


def test_spaces():
    obs = make_observation_space('xyz_vxyz_R_omega_floor', 2, True, (10., 10., 10.))
    assert obs.shape == (19 + 12 + 9,) and obs.dtype == np.float32
    assert make_action_space().shape == (4,) and float(make_action_space().low[0]) == -1.0


class _FakeEnv:
    """Stand-in with the QuadrotorEnvMulti protocol for wrapper tests."""
    is_multiagent = True
    num_agents = 2

    def __init__(self):
        self.rew_coeff = dict(pos=1.0, quadcol_bin=9.0, quadcol_bin_smooth_max=9.0, quadcol_bin_obst=9.0)
        self.scenario = types.SimpleNamespace(name=lambda: 'Scenario_static_same_goal')
        self.t = 0

    @property
    def unwrapped(self):
        return self

    def reset(self):
        self.t = 0
        return np.zeros((2, 3))

    def step(self, action):
        self.t += 1
        done = self.t >= 3
        infos = [{'rewards': {'rew_pos': -0.1, 'rew_crash': 0.0, 'rewraw_main': -0.2, 'rewraw_quadcol': -1.0 if i == 1 and self.t == 2 else 0.0}}
                 for i in range(2)]
        return np.ones((2, 3)) * self.t, [0.5, 0.5], [done, done], infos

    def close(self):
        pass


def test_reward_shaping_annealing_and_true_reward():
    env = _FakeEnv()
    scheme = dict(quad_rewards=dict(pos=1.0, quadcol_bin=0.0, quadcol_bin_smooth_max=0.0, quadcol_bin_obst=0.0))
    ann = [wr.AnnealSchedule('quadcol_bin', 5.0, 1000)]
    w = wr.QuadEnvCompatibility(wr.QuadsRewardShapingWrapper(env, reward_shaping_scheme=scheme, annealing=ann))
    obs, info = w.reset()
    assert info == {} and obs.shape == (2, 3)
    w.env.training_info['approx_total_training_steps'] = 500
    for t in range(3):
        obs, rew, term, trunc, infos = w.step(np.full((2, 4), 0.1 * t))
    assert env.rew_coeff['quadcol_bin'] == 2.5            # annealed: 5.0 * 500 / 1000
    assert term.all() and not trunc.any()
    assert infos[0]['true_reward'] == pytest.approx(-0.6)
    assert infos[1]['true_reward'] == pytest.approx(-0.6 - 1000.0)
    st = infos[1]['episode_extra_stats']
    assert st['Scenario_static_same_goal/rew_pos'] == pytest.approx(-0.3) and st['z_anneal_quadcol_bin'] == 2.5
    assert st['z_action0_mean'] == pytest.approx(0.1)


def test_factory_rejects_what_is_out_of_scope():
    cfg = types.SimpleNamespace(replay_buffer_sample_prob=0.0, visualize_v_value=True)
    with pytest.raises(NotImplementedError):
        wr.make_quadrotor_env('quadrotor_multi', cfg=cfg)
    with pytest.raises(NotImplementedError):
        wr.make_quadrotor_env('quadrotor_single', cfg=cfg)


class _FakeReplayEnv(_FakeEnv):
    """Adds what ExperienceReplayWrapper reads (quad_experience_replay.py:71,140-151,183-188)."""
    use_replay_buffer = True
    use_obstacles = False
    collisions_grace_period_seconds = 1.5
    obst_density = 0.2

    def __init__(self):
        super().__init__()
        self.activate_replay_buffer = True
        self.saved_in_replay_buffer = False
        self.tick = 0
        self.envs = [types.SimpleNamespace(control_freq=100.0)]
        self.last_step_unique_collisions = np.array([], dtype=int)
        self.curr_quad_col = []
        self.restored = []

    def reset(self, obst_density=None, obst_size=None):
        self.tick = 0
        return np.zeros((2, 3))

    def step(self, action):
        self.tick += 1
        self.envs[0].tick = self.tick
        done = self.tick >= 400
        self.last_step_unique_collisions = np.array([1, 2]) if self.tick == 320 else np.array([], dtype=int)
        infos = [{'rewards': {}} for _ in range(2)]
        if done:
            self.tick = 0
            self.envs[0].tick = 0
        return np.full((2, 3), float(self.tick)), [0., 0.], [done, done], infos

    def snapshot(self):
        return {'tick': self.tick}

    def restore(self, snap, zero_collision_counters=False):
        self.restored.append(snap['tick'])
        self.tick = snap['tick']
        self.saved_in_replay_buffer = True


def test_replay_wrapper_saves_the_checkpoint_from_1_5_s_before_a_collision():
    from quad_swarm_rl_b200.replay import ExperienceReplayWrapper
    env = _FakeReplayEnv()
    w = ExperienceReplayWrapper(env, 1.0, 0.2, 0.6)
    w.reset()
    for t in range(400):
        obs, rew, dones, infos = w.step(None)
    assert env.saved_in_replay_buffer and len(w.replay_buffer) == 1
    # collision at tick 320; checkpoints every 50 ticks -> [..., 200, 250, 300]; 1.5 s = 3 checkpoints back -> tick 200
    assert w.replay_buffer.buffer[0].snapshot == {'tick': 200}
    assert dones[0] and env.restored == [200] and w.replayed_events == 1
    assert obs[0, 0] == 200.0                                          # the observation stored with that checkpoint
    st = infos[0]['episode_extra_stats']
    assert st['replay/replay_rate'] == 1.0 and st['replay/replay_buffer_size'] == 1


def test_svd_period_is_100_substeps():
    """quadrotor_dynamics.py:547-551 accumulates 0.005 in float64 and re-orthogonalises when the sum exceeds 0.5;
    the kernels use an integer period of 100 sub-steps (qs_device.cuh SVD_PERIOD).  Check the float behaviour."""
    acc, fired = 0.0, []
    for k in range(1, 1001):
        acc += 0.005
        if acc > 0.5:
            fired.append(k)
            acc = 0
    assert fired == list(range(100, 1001, 100))


def test_shard_range_partitions():
    for total, world in ((32768, 8), (10, 3), (7, 8)):
        spans = [shard_range(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


_GLOO_CHILD = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from quad_swarm_rl_b200.sharding import shard_range, reduce_metrics
dist.init_process_group('gloo')
r, w = dist.get_rank(), dist.get_world_size()
lo, hi = shard_range(10, w, r)
m = torch.zeros(4); m[0] = hi - lo; m[1] = sum(range(lo, hi)); m[2] = 1
reduce_metrics(m)
assert m.tolist() == [10.0, 45.0, float(w), 0.0], m
mx = torch.tensor([float(r)]); reduce_metrics(mx, op='max'); assert mx.item() == w - 1
dist.destroy_process_group()
print('ok', r)
'''


def test_sharding_two_ranks_gloo(tmp_path):
    """N>1 host logic on CPU: two gloo ranks shard the env ids without overlap and reduce a metrics vector."""
    script = tmp_path / 'child.py'
    script.write_text(_GLOO_CHILD % ROOT)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29541', OMP_NUM_THREADS='1')
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                        '--master-addr', '127.0.0.1', '--master-port', '29541', str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count('ok') == 2


# ---- batched wrappers (batched.py) on CPU tensors with a stand-in engine ----
class _FakeBatchedEngine:
    """Counts ticks per env, auto-resets after ep_len steps, reports a collision when told to."""

    def __init__(self, E, N, D, ep_len):
        import torch
        self.device = torch.device('cpu')
        self.E, self.N, self.D, self.ep_len, self.M = E, N, D, ep_len, 0
        self.af = torch.zeros((E, N, 43)); self.au = torch.zeros((E, N, 4), dtype=torch.int32)
        self.ei = torch.zeros((E, 36), dtype=torch.int32)
        self.rew_terms = torch.zeros((E, N, 8))
        self.rew_coeff = dict(pos=1.0, effort=0.05, crash=1.0, orient=1.0, spin=0.1, quadcol_bin=5.0,
                              quadcol_bin_smooth_max=4.0, quadcol_bin_obst=5.0)
        self.collide = torch.zeros(E, dtype=torch.bool)
        self.stats_env = torch.zeros((E, 13), dtype=torch.int32)

    def get_state(self):
        import torch
        return dict(agent_f32=self.af.clone(), agent_u32=self.au.clone(), env_i32=self.ei.clone(),
                    obst_xy=torch.zeros((self.E, 0, 2)))

    def set_state(self, st, env_mask=None):
        m = env_mask.bool()
        self.af[m] = st['agent_f32'][m]; self.au[m] = st['agent_u32'][m]; self.ei[m] = st['env_i32'][m]

    def episode_stats(self):
        import torch
        return self.stats_env.clone(), torch.zeros((self.E, self.N, 4))


class _FakeBatchedEnv:
    device_scenario = 'static_same_goal'
    use_obstacles = False
    quads_mode = 'static_same_goal'

    def __init__(self, E=6, N=2, D=5, ep_len=300):
        self.engine = _FakeBatchedEngine(E, N, D, ep_len)
        self.num_envs, self.num_agents_per_env, self.num_agents = E, N, E * N

    def reset(self, **kw):
        e = self.engine
        e.af.zero_(); e.ei.zero_()
        return e.af[..., :e.D].reshape(self.num_agents, -1).clone(), {}

    def step(self, actions, with_terms=False):
        import torch
        e = self.engine
        e.ei[:, 0] += 1
        e.ei[:, 1] += 1                                   # RNG step counter: never rewinds
        e.af += 1.0                                       # the "physics": every state entry counts steps of the episode
        done = e.ei[:, 0] > e.ep_len
        e.rew_terms.zero_()
        e.rew_terms[..., 0] = -0.01
        e.rew_terms[e.collide, 0, 5] = -1.0
        e.collide.zero_()
        if done.any():                                    # auto-reset inside the "kernel"
            e.stats_env[done, 11] += 1
            e.stats_env[done, 12] = 2
            e.af[done] = 0.0
            e.ei[done, 0] = 0
            e.ei[done, 3] += 1
        obs = e.af[..., :e.D].reshape(self.num_agents, -1).clone()
        rew = e.rew_terms[..., 0].reshape(-1).clone()
        term = done.repeat_interleave(self.num_agents_per_env)
        return obs, rew, term, torch.zeros_like(term), {}


def test_batched_replay_logic_on_cpu_tensors():
    """BatchedExperienceReplay (quad_experience_replay.py:66-209 per env, masked): checkpoint every 50 ticks, the
    checkpoint from 3 checkpoints back is stored on a collision after the grace period, one event per 5 s, replay at the
    episode end restores that state (tick, state rows, observation) while the RNG step counter keeps running."""
    import torch
    from quad_swarm_rl_b200.batched import BatchedExperienceReplay
    env = _FakeBatchedEnv()
    rp = BatchedExperienceReplay(env, replay_buffer_sample_prob=1.0, always_active=True, seed=1)
    rp.reset()
    E, N = env.num_envs, env.num_agents_per_env
    a = torch.zeros((E * N, 4))
    for t in range(1, 302):
        if t == 120:
            env.engine.collide[0] = True                  # before the grace period (tick <= 150): ignored
        if t == 231:
            env.engine.collide[1] = True                  # checkpoints at 50..200 exist -> stores the one of tick 100
            env.engine.collide[2] = True
        if t == 260:
            env.engine.collide[1] = True                  # same episode again: already saved, ignored
        obs, rew, term, trunc, infos = rp.step(a)
    assert term.all() and rp.episode_counter == E
    assert rp.buf_valid.sum(dim=0).tolist() == [0, 1, 1, 0, 0, 0]
    assert rp.buf['env_i32'][0, 1, 0] == 100 and torch.all(rp.buf['agent_f32'][0, 1] == 100.0)
    # envs 1 and 2 were replayed (p = 1): tick 100, state rows of tick 100, the stored observation; others start fresh
    assert rp.tick.tolist() == [0, 100, 100, 0, 0, 0] and rp.replayed_events == 2
    assert torch.all(env.engine.af[1] == 100.0) and torch.all(env.engine.af[0] == 0.0)
    assert torch.all(obs.view(E, N, -1)[2] == 100.0) and torch.all(obs.view(E, N, -1)[3] == 0.0)
    assert env.engine.ei[1, 1] == 301 and env.engine.ei[1, 0] == 100          # step counter kept, tick restored
    assert rp.saved.tolist() == [False, True, True, False, False, False]
    # replayed envs end after ep_len + 1 - 100 more steps, and then (p = 1) replay the same event again
    for t in range(1, 202):
        obs, rew, term, trunc, infos = rp.step(a)
    d = term.view(E, N)[:, 0].tolist()
    assert d == [False, True, True, False, False, False]
    assert rp.buf_replayed[0, 1] == 2 and rp.tick.tolist() == [201, 100, 100, 201, 201, 201]
    assert infos['replay']['replay/replay_rate'] == pytest.approx(4 / 8)


def test_batched_reward_shaping_on_cpu_tensors():
    import torch
    from quad_swarm_rl_b200.batched import BatchedRewardShaping
    env = _FakeBatchedEnv(E=3, N=2, ep_len=4)
    w = BatchedRewardShaping(env, reward_shaping_scheme=dict(quad_rewards=dict(pos=2.0)),
                             annealing=[wr.AnnealSchedule('quadcol_bin', 5.0, 100)])
    w.training_info['approx_total_training_steps'] = 50
    w.reset()
    infos = {}
    for t in range(5):
        env.engine.collide[0] = (t == 2)
        obs, rew, term, trunc, infos = w.step(torch.full((6, 4), 0.5))
    st = infos['episode_extra_stats']
    assert term.all() and env.engine.rew_coeff['quadcol_bin'] == 2.5 and st['z_anneal_quadcol_bin'] == 2.5
    assert st['rewraw_pos'] == pytest.approx(-0.05) and st['rew_pos'] == pytest.approx(-0.10)       # coefficient 2.0
    assert st['rewraw_quadcol'] == pytest.approx(-1.0 / 6)                                          # one drone of six, once
    assert infos['true_reward'][0, 0].item() == pytest.approx(-0.05 - 1000.0)
    assert st['z_action0_mean'] == pytest.approx(0.5) and st['z_action0_std'] == pytest.approx(0.0, abs=1e-6)
    assert 'Scenario_static_same_goal/rew_pos' in st
