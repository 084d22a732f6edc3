"""CPU checks of the device-side scenario family's twin (oracle/scenario_gen.py) against golden vectors of the
reference (tests/golden/formations.npz, made by oracle/gen_golden_formations.py) and against the host generators
(quad_swarm_rl_b200/scenarios.py, themselves replayed against reference trajectories in test_oracle_vs_reference.py).
The GPU tests (test_gpu_parity.py) then compare the kernels with this twin."""
import os

import numpy as np
import pytest

from oracle import philox as px
from oracle import quadswarm_oracle as qo
from oracle import scenario_gen as sg
from quad_swarm_rl_b200 import scenarios as hs
from tests import parity_util as pc

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'formations.npz'))
SIZE, LAYER = float(GOLD['params'][0]), float(GOLD['params'][1])
CENTER = GOLD['params'][2:5]


class FixedDraws:
    """Stand-in for philox.KeyedDraws: every uniform is the same number."""

    def __init__(self, u):
        self.u = u

    def uniform(self, site, i, j, v):
        return self.u


def test_formation_names_and_order():
    assert tuple(GOLD['formations']) == sg.FORMATION_NAMES == hs.FORMATIONS


@pytest.mark.parametrize('f', range(8))
def test_formation_geometry_equals_reference(f):
    for n in range(1, 33):
        ref = GOLD[f'goals_{f}_{n}']
        pl = sg.per_layer_of(f)
        twin = np.array([sg.formation_point(f, n, k, SIZE, CENTER, LAYER, pl) for k in range(n)])
        np.testing.assert_allclose(twin, ref, rtol=0, atol=1e-12, err_msg=f'twin {sg.FORMATION_NAMES[f]} n={n}')
        host = hs.formation_goals(sg.FORMATION_NAMES[f], n, SIZE, CENTER, LAYER, pl)[:n]
        np.testing.assert_allclose(host, ref, rtol=0, atol=1e-12, err_msg=f'host {sg.FORMATION_NAMES[f]} n={n}')


def test_cube_side_table_of_the_kernel():
    # qs_scenario.cuh hard-codes int(np.power(n, 1/3)) for n <= 32 (27 -> 2: the float64 cube root is below 3)
    for n in range(1, 33):
        assert int(np.power(n, 1.0 / 3)) == (3 if n >= 28 else (2 if n >= 8 else 1))


def test_formation_size_ranges_equal_reference():
    modes = {'static_diff_goal': sg.STATIC_DIFF_GOAL, 'swap_goals': sg.SWAP_GOALS, 'dynamic_formations': sg.DYNAMIC_FORMATIONS}
    for name, mode in modes.items():
        for f in range(8):
            for n in range(1, 33):
                lo, hi = GOLD[f'range_{name}_{f}_{n}']
                # force the formation pick to f: u = (f + 0.5) / 8
                fm = sg.pick_formation(FixedDraws((f + 0.5) / 8.0), 1, mode, n)
                assert fm['f'] == f and fm['per_layer'] == (50 if 4 <= f <= 6 else 8)
                np.testing.assert_allclose([fm['lo'], fm['hi']], [lo, hi], rtol=1e-12, atol=0)
                assert fm['lo'] <= fm['size'] <= fm['hi'] and fm['lo'] <= fm['layer'] <= fm['hi']
    for mode in (sg.STATIC_SAME_GOAL, sg.DYNAMIC_SAME_GOAL, sg.EP_LISSAJOUS3D):     # one formation, zero size
        fm = sg.pick_formation(FixedDraws(0.7), 1, mode, 8)
        assert fm['f'] == 0 and fm['size'] == 0.0 and fm['layer'] == 0.0


def test_centre_height_equals_reference():
    for f in range(8):
        for n in range(1, 33):
            u, z = GOLD[f'z_{f}_{n}']                 # u ~ U(-1, 1) as drawn by the reference, z = get_z_value(...)
            got = sg.z_above_ground((u + 1.0) / 2.0, n, sg.per_layer_of(f), f, SIZE)
            assert abs(got - z) < 1e-12


def test_shuffle_rank_is_a_permutation_and_uniform_enough():
    counts = np.zeros((6, 6), int)
    for step in range(600):
        d = px.KeyedDraws(99, 3, step)
        perm = [sg.shuffle_rank(d, sg.STREAM_TICK, i, 0, 6) for i in range(6)]
        assert sorted(perm) == list(range(6))
        for i, k in enumerate(perm):
            counts[i, k] += 1
    assert counts.min() > 50 and counts.max() < 160          # 100 expected per cell
    d = px.KeyedDraws(99, 3, 7)                               # ranks inside a sub-range (swarm_vs_swarm halves)
    assert sorted(sg.shuffle_rank(d, sg.STREAM_TICK, i, 4, 8) for i in range(4, 8)) == [0, 1, 2, 3]


@pytest.mark.parametrize('mode', sorted(sg.MODE_IDS))
def test_twin_episode_semantics(mode):
    """Run the twin inside the oracle env: the reference's scenario semantics hold (scenarios/*.py)."""
    N = 8
    cfg = pc.cfg_to_oracle(dict(num_agents=N, obs_repr='xyz_vxyz_R_omega', neighbor_visible_num=2, ep_time=6.2))
    src = sg.DeviceScenarioSource(mode)
    env = qo.OracleEnv(cfg, qo.PhiloxRng(2024), src, env_id=5)
    env.reset()
    s0 = dict(src.s)
    g0 = np.array([d.goal for d in env.drones])
    rs = np.random.RandomState(0)
    assert src.name() == 'Scenario_' + sg.MODE_NAMES[s0['mode']]
    if s0['mode'] in (sg.STATIC_SAME_GOAL, sg.DYNAMIC_SAME_GOAL):
        assert np.all(g0 == np.array([0.0, 0.0, 2.0]))
    if s0['mode'] == sg.EP_LISSAJOUS3D:
        assert np.all(g0 == np.array([-2.0, 0.0, 2.0]))
    if s0['mode'] in (sg.STATIC_DIFF_GOAL, sg.DYNAMIC_DIFF_GOAL, sg.SWAP_GOALS, sg.DYNAMIC_FORMATIONS):
        # a shuffled formation of the picked type around (0, 0, 2): same point set as the host generator's
        want = hs.formation_goals(sg.FORMATION_NAMES[s0['f']], N, s0['size'], np.array([0., 0., 2.]), s0['layer'],
                                  sg.per_layer_of(s0['f']))[:N]
        assert np.allclose(sorted(map(tuple, np.round(g0, 9))), sorted(map(tuple, np.round(want, 9))))
    if s0['mode'] == sg.SWARM_VS_SWARM:
        h = N // 2
        for half, c in ((slice(0, h), s0['c1']), (slice(h, N), s0['c2'])):
            want = hs.formation_goals(sg.FORMATION_NAMES[s0['f']], h, s0['size'], c, s0['layer'], sg.per_layer_of(s0['f']))[:h]
            assert np.allclose(g0[half], want)                 # not shuffled at reset (swarm_vs_swarm.py:74-84)
        assert BOX_DIST_OK(s0['c1'], s0['c2'])
    periodic = s0['mode'] in (sg.DYNAMIC_SAME_GOAL, sg.DYNAMIC_DIFF_GOAL, sg.SWAP_GOALS, sg.SWARM_VS_SWARM)
    if periodic:
        assert 400 <= s0['period'] <= 599 and s0['next'] == s0['period']
    prev = g0
    for t in range(1, 611):
        env.step(rs.uniform(-1, 1, (N, 4)))
        g = np.array([d.goal for d in env.drones])
        if periodic:
            if t == s0['period']:
                assert src.events == 1
                if s0['mode'] == sg.SWAP_GOALS:
                    assert sorted(map(tuple, g)) == sorted(map(tuple, prev))
                if s0['mode'] == sg.SWARM_VS_SWARM:
                    assert np.allclose(src.s['c1'], s0['c2']) and np.allclose(src.s['c2'], s0['c1'])
                if s0['mode'] == sg.DYNAMIC_SAME_GOAL:
                    assert np.all(g == g[0]) and abs(g[0, 0]) <= 2 and abs(g[0, 1]) <= 2 and 0.25 <= g[0, 2] <= 3.0
            elif t < s0['period']:
                assert np.array_equal(g, prev)
        if s0['mode'] == sg.EP_LISSAJOUS3D:
            tt = t / 100.0
            step = np.array([0.03 * np.sin(tt), 0.01 * np.sin(2 * tt + 90), 0.01 * np.cos(2 * tt + 90)])
            assert np.allclose(g, prev[0] + step)
        if s0['mode'] == sg.DYNAMIC_FORMATIONS:
            assert abs(abs(src.s['size']) - 0) <= s0['hi'] + 0.004
        assert g[:, 2].min() > 0.0
        prev = g
    assert env.tick == 610 and (src.events >= 1 or s0['mode'] in (sg.STATIC_SAME_GOAL, sg.STATIC_DIFF_GOAL))


def BOX_DIST_OK(c1, c2):
    d = np.linalg.norm(np.asarray(c2) - np.asarray(c1))
    return 0.5 - 1e-9 <= d <= 2.0 + 1.5          # U(box/4, box) plus the push-apart along the formation normal


def test_mix_draws_every_available_scenario():
    seen = set()
    cfg = pc.cfg_to_oracle(dict(num_agents=4, obs_repr='xyz_vxyz_R_omega', neighbor_visible_num=0, ep_time=0.05))
    for env_id in range(150):
        src = sg.DeviceScenarioSource('mix')
        env = qo.OracleEnv(cfg, qo.PhiloxRng(7), src, env_id=env_id)
        env.reset()
        seen.add(src.s['mode'])
    assert seen == set(range(sg.STATIC_SAME_GOAL, sg.SWARM_VS_SWARM + 1)) | {sg.EP_RAND_BEZIER}      # the 9 modes of scenarios/utils.py:7-10
    single = set()
    cfg1 = pc.cfg_to_oracle(dict(num_agents=1, obs_repr='xyz_vxyz_R_omega', neighbor_visible_num=0, ep_time=0.05))
    for env_id in range(40):
        src = sg.DeviceScenarioSource('mix')
        qo.OracleEnv(cfg1, qo.PhiloxRng(7), src, env_id=env_id).reset()
        single.add(src.s['mode'])
    assert single == {sg.STATIC_SAME_GOAL, sg.STATIC_DIFF_GOAL, sg.EP_LISSAJOUS3D, sg.EP_RAND_BEZIER, sg.DYNAMIC_SAME_GOAL}


def test_largest_free_square_equals_host_generator():
    """oracle/scenario_gen.py:largest_free_square_cell (the kernels' twin) against scenarios.py's restatement of
    o_base.py:123-153, which runs on the reference's obst_map / cell_centers layout (replayed against the reference
    in test_oracle_vs_reference.py: c3 golden case uses the same layout)."""
    rs = np.random.RandomState(3)
    cells = hs.grid_cell_centers(8, 8)
    for trial in range(300):
        M = rs.randint(1, 40)
        occ = rs.choice(64, M, replace=False)
        obst_map = np.zeros((8, 8))
        mask = 0
        for c in occ:
            obst_map[c // 8, c % 8] = 1
            mask |= 1 << int(c)
        sc = hs.OStaticSameGoal(4, rng=np.random.RandomState(0), use_obstacles=True)
        sc.obstacle_map, sc.cell_centers = obst_map, cells
        want = sc._largest_free_square_center()[:2]
        got = sg._cell_center(sg.largest_free_square_cell(mask, 8, 8), 8, 8)
        assert np.allclose(got, want), (trial, got, want)


@pytest.mark.parametrize('scenario', ['o_static_same_goal', 'mix'])
def test_obstacle_twin_episode_semantics(scenario):
    kw = dict(num_agents=8, obs_repr='xyz_vxyz_R_omega_floor', neighbor_visible_num=2, use_obstacles=True, ep_time=0.1)
    cfg = pc.cfg_to_oracle(kw)
    modes = set()
    for env_id in range(12):
        src = sg.DeviceORandomSource(scenario=scenario)
        env = qo.OracleEnv(cfg, qo.PhiloxRng(11), src, env_id=env_id)
        env.reset()
        modes.add(src.mode)
        goals = np.array([d.goal for d in env.drones])
        pill = env.obst_xy
        assert len({tuple(p) for p in pill}) == cfg.num_obstacles                 # distinct pillar cells
        if src.mode == sg.O_STATIC_SAME_GOAL:
            assert np.all(goals == goals[0]) and 1.5 <= goals[0, 2] <= 3.0 and src.approch_goal_metric == 1.0
            assert src.name() == 'Scenario_o_static_same_goal'
            assert np.abs(pill - goals[0, :2]).max(axis=1).min() >= 1.0 - 1e-9    # the goal cell itself is free
        else:
            assert len({tuple(g[:2]) for g in goals}) == 8 and src.approch_goal_metric == 0.5
        for t in range(12):
            env.step(np.zeros((8, 4)))
    assert modes == ({sg.O_STATIC_SAME_GOAL} if scenario == 'o_static_same_goal' else {sg.O_RANDOM, sg.O_STATIC_SAME_GOAL})


def test_device_bezier_twin_follows_the_host_scenario_semantics():
    """oracle/scenario_gen.py's ep_rand_bezier (twin of the kernels') against the reference-pinned host class
    (scenarios.RandBezier, replayed against the reference in test_oracle_vs_reference.py[ep_rand_bezier_3]): same schedule
    (new segment at tick 1 and every 500 ticks, goal unchanged on those ticks), control points 5..10 m from the segment's
    start and 0.5 m inside the room, and the quadratic Bernstein interpolation at s = t / 499."""
    cfg = pc.cfg_to_oracle(dict(num_agents=3, obs_repr='xyz_vxyz_R_omega', neighbor_visible_num=2, ep_time=11.0))
    src = sg.DeviceScenarioSource('ep_rand_bezier')
    env = qo.OracleEnv(cfg, qo.PhiloxRng(11), src, env_id=3)
    env.reset()
    assert np.allclose(src.goals, [[0.0, 0.0, 2.0]] * 3)
    prev = src.goals.copy()
    for tick in range(1, 1010):
        g = src.step(env, tick)
        g = prev if g is None else g
        s = src.s
        p0, p1, p2 = np.array([s['size'], s['layer'], s['hi']]), np.asarray(s['c1']), np.asarray(s['c2'])
        if tick == 1 or tick % 500 == 0:
            assert np.allclose(g, prev) and np.allclose(p0, prev[0])                      # re-drawn, goal not moved
            for q in (p1, p2):
                d = np.linalg.norm(q - p0)
                assert abs(d - round(d)) < 1e-9 and 5 <= round(d) <= 10
                assert (q > [-4.5, -4.5, 0.5]).all() and (q < [4.5, 4.5, 9.5]).all()
        else:
            sp = (tick % 500) / 499.0
            want = hs.bezier_points(np.stack([p0, p1, p2], 1), np.array([sp]))[:, 0]
            assert np.allclose(g, np.tile(want, (3, 1)), atol=1e-12)
        prev = np.array(g)


@pytest.mark.parametrize('scenario', ['o_dynamic_same_goal', 'o_swap_goals', 'o_ep_rand_bezier'])
def test_ticked_obstacle_twin_follows_the_reference_semantics(scenario):
    """Twin of the kernels' ticked obstacle scenarios against what obstacles/o_dynamic_same_goal.py, o_swap_goals.py and
    o_ep_rand_bezier.py do (the host classes of scenarios.py replay those against the reference in
    test_oracle_vs_reference.py): event schedule, hop distance, free cells, permutation, control-point box, Bernstein form."""
    kw = dict(num_agents=5, obs_repr='xyz_vxyz_R_omega_floor', neighbor_visible_num=2, use_obstacles=True, ep_time=13.0)
    cfg = pc.cfg_to_oracle(kw)
    for env_id in range(3):
        src = sg.DeviceORandomSource(scenario=scenario)
        env = qo.OracleEnv(cfg, qo.PhiloxRng(5), src, env_id=env_id)
        env.reset()
        assert src.approch_goal_metric == 1.0 and src.name() == 'Scenario_' + scenario
        pill = {tuple(p) for p in env.obst_xy}
        goals = src.goals.copy()
        spawn = np.array([d.pos for d in env.drones])
        assert len({tuple(np.round(p[:2] * 2) / 2) for p in spawn}) <= 5                     # jittered around distinct free cells
        if scenario == 'o_swap_goals':
            assert 400 <= src.period <= 599 and src.next == src.period
            assert len({tuple(g) for g in goals}) == 5 and np.linalg.norm(goals - goals.mean(0), axis=1).max() < 1.5
        else:
            assert np.all(goals == goals[0]) and tuple(goals[0, :2]) not in pill
            assert src.next == 1
        events = []
        for tick in range(1, 1250):
            prev = src.goals.copy()
            g = src.step(env, tick)
            if g is None:
                continue
            events.append(tick)
            if scenario == 'o_dynamic_same_goal':
                assert np.all(g == g[0]) and np.linalg.norm(g[0] - prev[0]) <= 4.0 + 1e-9
                assert tuple(g[0, :2]) not in pill and 0.75 <= g[0, 2] <= 3.0
                assert abs(g[0, 0] % 1.0 - 0.5) < 1e-9 and abs(g[0, 1] % 1.0 - 0.5) < 1e-9     # a cell centre
            elif scenario == 'o_swap_goals':
                assert sorted(map(tuple, g)) == sorted(map(tuple, prev))
            else:
                if tick == 1 or tick % 600 == 0:
                    assert np.allclose(g, prev) and np.allclose(src.p0, prev[0])
                    for q in (src.p1, src.p2):
                        d = np.linalg.norm(q - src.p0)
                        assert min(abs(d - k) for k in (2, 3, 4, 5)) < 1e-9                   # randint(2.5, 6) -> 2..5
                        assert abs(q[0]) < 4.5 and abs(q[1]) < 4.5 and 2.0 < q[2] < 2.5      # o_ep_rand_bezier.py:31-44
                else:
                    t = tick % 600
                    want = hs.bezier_points(np.stack([src.p0, src.p1, src.p2], axis=1), np.linspace(0, 1, 600))[:, t]
                    assert np.allclose(g[0], want, atol=1e-12) and np.all(g == g[0])
        if scenario == 'o_dynamic_same_goal':
            assert events == [1] + list(range(src.period, 1250, src.period))
        elif scenario == 'o_swap_goals':
            assert events == list(range(src.period, 1250, src.period))
        else:
            assert events == list(range(1, 1250))


# ---- same DISTRIBUTIONS as the reference-pinned host generators (the kernels draw from keyed Philox streams, the reference
#      from numpy's global stream: the samples differ, the laws must not) ----
def _p_discrete(a, b, k):
    from scipy.stats import chi2_contingency
    t = np.array([np.bincount(a, minlength=k), np.bincount(b, minlength=k)])
    t = t[:, t.sum(axis=0) > 0]
    return chi2_contingency(t)[1]


def _p_cont(a, b):
    from scipy.stats import ks_2samp
    return ks_2samp(a, b)[1]


@pytest.mark.parametrize('mode', ['static_diff_goal', 'swap_goals', 'dynamic_formations', 'swarm_vs_swarm', 'o_swap_goals'])
def test_formation_pick_law_equals_host_generator(mode):
    """Formation type, size, layer distance and the 4-6 s period: two-sample tests (chi-square / Kolmogorov-Smirnov, n = 1500
    each) between the twin's keyed draws and the host class driven by a RandomState — the class whose trajectories replay
    the reference's (test_oracle_vs_reference.py)."""
    n, N = 1500, 8
    rs = np.random.RandomState(7)
    host = hs.create_scenario(mode, N, rng=rs, use_obstacles=mode.startswith('o_'))
    hf, hsz, hl, hp = [], [], [], []
    for _ in range(n):
        host.pick_formation()
        hf.append(hs.FORMATIONS.index(host.formation)); hsz.append(host.formation_size); hl.append(host.layer_dist)
        hp.append(int(rs.uniform(low=4.0, high=6.0) * 100.0))
    mid = sg.O_SWAP_GOALS if mode == 'o_swap_goals' else sg.MODE_IDS[mode]
    tf, tsz, tl, tp = [], [], [], []
    for e in range(n):
        d = px.KeyedDraws(31, e, px.EPISODE_KEY_BIT | 1)
        fm = sg.pick_formation(d, sg.STREAM_RESET, mid, N // 2 if mode == 'swarm_vs_swarm' else N)
        tf.append(fm['f']); tsz.append(fm['size']); tl.append(fm['layer'])
        tp.append(400 + sg._pk(d, sg.STREAM_RESET, sg.SV_PERIOD, 200))
    hf, tf = np.array(hf), np.array(tf)
    assert set(tf) == set(hf) and _p_discrete(hf, tf, 8) > 1e-3
    for f in set(hf):                                          # size law per formation type (the ranges differ by type)
        a, b = np.array(hsz)[hf == f], np.array(tsz)[tf == f]
        assert _p_cont(a, b) > 1e-3 and abs(a.min() - b.min()) < 0.05 * (a.max() - a.min() + 1e-9) + 1e-9
        assert _p_cont(np.array(hl)[hf == f], np.array(tl)[tf == f]) > 1e-3
    assert _p_cont(np.array(hp, float), np.array(tp, float)) > 1e-3 and min(tp) >= 400 and max(tp) <= 599


def test_obstacle_episode_law_equals_host_generator():
    """o_random: pillar cells, spawn cells and goal heights of the twin against obstacle_map_given_density + ORandom (host,
    reference-pinned): per-cell pillar frequency (chi-square over the 64 cells), spawn-on-free-cell, z ~ U(1, 3)."""
    n, N, L = 600, 8, 8
    rs = np.random.RandomState(3)
    host = hs.create_scenario('o_random', N, rng=rs, use_obstacles=True)
    hc, hz, hspawn = [], [], []
    for _ in range(n):
        obst_map, pos, cells = hs.obstacle_map_given_density(rs, (8.0, 8.0), 0.2)
        host.reset(obst_map=obst_map, cell_centers=cells)
        for p in pos:
            hc.append(int(np.floor(p[0] + 4)) * L + int(np.floor(p[1] + 4)))
        hz += list(host.goals[:, 2])
        hspawn += [int(np.floor(p[0] + 4)) * L + int(np.floor(p[1] + 4)) for p in host.spawn_points]
    tc, tz, tspawn = [], [], []
    for e in range(n):
        d = px.KeyedDraws(5, e, px.EPISODE_KEY_BIT | 2)
        goals, spawn, obst = sg.o_random_episode(d, N, 12, L, L)
        tc += [int(np.floor(p[0] + 4)) * L + int(np.floor(p[1] + 4)) for p in obst]
        tz += list(goals[:, 2])
        tspawn += [int(np.floor(p[0] + 4)) * L + int(np.floor(p[1] + 4)) for p in spawn]
        assert not ({tuple(p) for p in obst} & {tuple(p[:2]) for p in spawn})
    assert _p_discrete(np.array(hc), np.array(tc), 64) > 1e-3
    assert _p_discrete(np.array(hspawn), np.array(tspawn), 64) > 1e-3
    assert _p_cont(np.array(hz), np.array(tz)) > 1e-3


def test_run_away_twin_follows_the_reference_schedule_and_draw_range():
    """oracle/scenario_gen.py's run_away (twin of the kernels') against scenarios/run_away.py:14-25 as restated by the
    reference-pinned host class (scenarios.RunAway, replayed against the reference in
    test_oracle_vs_reference.py[run_away_5]): an event at every tick that is a positive multiple of 100 and at no other
    tick; drones 0 and 1 take goals that drones 1..N-1 held before the event, the rest keep theirs; the two draws are
    uniform over 1..N-1 and never 0 (counted at the first event of an episode, where all goals are still distinct)."""
    N = 5
    cfg = pc.cfg_to_oracle(dict(num_agents=N, obs_repr='xyz_vxyz_R_omega', neighbor_visible_num=2, ep_time=4.2))
    counts = np.zeros((2, N), dtype=int)
    for env_id in range(150):
        src = sg.DeviceScenarioSource('run_away')
        env = qo.OracleEnv(cfg, qo.PhiloxRng(23), src, env_id=env_id)
        env.reset()
        assert src.name() == 'Scenario_run_away' and src.s['mode'] == sg.RUN_AWAY
        g_prev = np.array([d.goal for d in env.drones])
        assert len({tuple(np.round(g, 9)) for g in g_prev}) == N          # a formation of N distinct points (size >= 5 arms)
        assert np.abs(g_prev[:, :2].mean(axis=0)).max() < 1.0             # around the centre (0, 0, 2), run_away.py:31
        T = 400 if env_id < 10 else 100
        for t in range(1, T + 1):
            env.step(np.zeros((N, 4)))
            g = np.array([d.goal for d in env.drones])
            if t % 100 == 0:
                assert np.array_equal(g[2:], g_prev[2:])
                for k in (0, 1):
                    hit = [j for j in range(1, N) if np.array_equal(g[k], g_prev[j])]
                    assert hit, (env_id, t, k)
                    if t == 100:
                        assert len(hit) == 1
                        counts[k, hit[0]] += 1
            else:
                assert np.array_equal(g, g_prev), (env_id, t)
            g_prev = g
        assert src.events == T // 100
    assert counts[:, 0].sum() == 0 and (counts.sum(axis=1) == 150).all()
    # 150 first events per drone over 4 sources: 37.5 expected each, sigma 5.3
    assert counts[:, 1:].min() > 12 and counts[:, 1:].max() < 65, counts


def test_run_away_host_class_draws_the_same_range():
    rs = np.random.RandomState(4)
    sc = hs.RunAway(5, rng=rs)
    sc.reset()
    seen = set()
    for k in range(1, 60):
        before = sc.goals.copy()
        sc.step(100 * k)
        for d in (0, 1):
            seen |= {j for j in range(5) if np.array_equal(sc.goals[d], before[j])}
        assert np.array_equal(sc.goals[2:], before[2:])
    assert 0 not in seen or len({tuple(g) for g in sc.goals}) < 5      # index 0 is never a source (equal goals aside)
