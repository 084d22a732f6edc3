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


class _FakeBatched:
    """Stand-in for training.BatchedTrainingEnv with one env of two drones (CPU tensors), for the single-env adapter."""

    def __init__(self):
        import torch
        self.torch = torch
        self.env = types.SimpleNamespace(num_agents_per_env=2)
        self.engine = types.SimpleNamespace(device='cpu')
        self.training_info = {}
        self.t = 0
        self.flushed = 0

    def reset(self):
        self.t = 0
        return self.torch.zeros((2, 3)), {}

    def step(self, a):
        self.t += 1
        done = self.t >= 3
        return (self.torch.ones((2, 3)) * self.t, self.torch.tensor([0.5, 0.25]), self.torch.tensor([done, done]),
                self.torch.tensor([False, False]), {})

    def flush_stats(self):
        self.flushed += 1
        return {'episode_extra_stats': {'rew_pos': -0.3, 'z_anneal_quadcol_bin': 2.5}, 'true_reward': self.torch.tensor([[-0.6, -1000.6]]),
                'episodes_finished': 1}

    def close(self):
        pass


def test_single_env_adapter_keeps_the_reference_protocol():
    """QuadEnvCompatibility(QuadsTrainingEnv(...)): numpy observations, lists of per-agent rewards / dones / infos, the
    gymnasium 5-tuple outside (compatibility.py:33-50); true_reward and episode_extra_stats appear on the terminal step."""
    b = _FakeBatched()
    w = wr.QuadEnvCompatibility(wr.QuadsTrainingEnv(b))
    obs, info = w.reset()
    assert info == {} and obs.shape == (2, 3) and obs.dtype == np.float64
    for t in range(3):
        obs, rew, term, trunc, infos = w.step(np.full((2, 4), 0.1 * t))
        assert isinstance(rew, list) and rew == [0.5, 0.25] and len(infos) == 2
        if t < 2:
            assert not term.any() and infos == [{}, {}] and b.flushed == 0
    assert term.all() and not trunc.any() and b.flushed == 1
    assert infos[0]['true_reward'] == pytest.approx(-0.6) and infos[1]['true_reward'] == pytest.approx(-1000.6)
    assert infos[1]['episode_extra_stats']['z_anneal_quadcol_bin'] == 2.5
    assert w.num_agents == 2 and w.is_multiagent


def test_factory_rejects_what_is_out_of_scope():
    cfg = types.SimpleNamespace(replay_buffer_sample_prob=0.0, visualize_v_value=True)
    with pytest.raises(NotImplementedError):
        wr.make_quadrotor_env('quadrotor_multi', cfg=cfg)
    with pytest.raises(NotImplementedError):
        wr.make_quadrotor_env('quadrotor_single', cfg=cfg)


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


# ---- training wrappers: aggregate of finished episodes -> the reference's episode_extra_stats keys ----
def test_stats_dict_names_and_means():
    """training.stats_dict turns the device-side sums (QS_WA_*, csrc/qs_wrap.cuh) into the keys reward_shaping.py:86-108
    and quadrotor_multi.py:626-718 produce, as means over the agent-episodes / env-episodes they cover."""
    from quad_swarm_rl_b200 import _lib as L
    from quad_swarm_rl_b200.training import stats_dict
    W = L.WA
    agg = np.zeros(L.QS_WRAP_AGG, np.float32)
    agg[W['AGENT_EPISODES']], agg[W['ENV_EPISODES']] = 16, 2          # two envs of 8 drones finished
    agg[W['TRUE_REWARD']] = -160.0
    agg[W['RAW0'] + 0], agg[W['REW0'] + 0] = -32.0, -64.0             # pos, coefficient 2
    agg[W['REW0'] + 6] = -1.6                                         # proximity
    agg[W['RAW0'] + 7], agg[W['REW0'] + 7] = -4.0, -20.0              # obstacle collisions
    agg[W['ACT_MEAN0'] + 2], agg[W['ACT_STD0'] + 1] = 1.6, 8.0
    agg[W['ENV_STAT0'] + 0], agg[W['ENV_STAT0'] + 7] = 6, 3
    agg[W['DIST0']] = 4.0
    agg[W['SUCCESS']], agg[W['COL']] = 4, 8
    agg[W['SCN0'] + 6 * 1: W['SCN0'] + 6 * 1 + 6] = [8, -40.0, -8.0, 1, 5, 2.0]       # o_random: one env
    agg[W['SCN0'] + 6 * 11: W['SCN0'] + 6 * 11 + 6] = [8, -24.0, 0.0, 1, 1, 2.0]      # o_static_same_goal
    agg[W['REPLAY_ENV_EPISODES']], agg[W['REPLAY_COLLISIONS']] = 2, 3
    st = stats_dict(agg, use_obstacles=True, fallback_scenario='mix')
    assert st['rewraw_main'] == -10.0 and st['rewraw_pos'] == -2.0 and st['rew_pos'] == -4.0 and st['rew_main'] == -4.0
    assert st['rew_proximity'] == pytest.approx(-0.1) and st['rewraw_quadcol_obstacle'] == -0.25 and st['rew_quadcol_obstacle'] == -1.25
    assert st['z_action2_mean'] == pytest.approx(0.1) and st['z_action1_std'] == 0.5
    assert st['num_collisions'] == 3.0 and st['num_collisions_obst_quad'] == 1.5 and st['distance_to_goal_1s'] == 0.25
    assert st['metric/agent_success_rate'] == 0.25 and st['metric/agent_col_rate'] == 0.5
    assert st['Scenario_o_random/rew_pos'] == -5.0 and st['o_random/num_collisions'] == 5.0
    assert st['Scenario_o_static_same_goal/rew_pos'] == -3.0 and st['num_collisions_replay'] == 1.5
    no_obst = stats_dict(agg, use_obstacles=False, fallback_scenario='mix')
    assert 'rew_quadcol_obstacle' not in no_obst and 'num_collisions_obst_quad' not in no_obst


