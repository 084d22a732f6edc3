"""GPU parity: the CUDA env step (through the C ABI) against the CPU oracle on identical seeds, actions and tables.

Tolerances (fp32 kernel vs fp64 oracle, teacher-forced every `resync` steps):
  observations / rewards / state: |err| <= 1e-4 + 1e-4 * |ref|   (BASELINE.json north_star: 1e-4 relative)
  done / collision / obstacle-collision / on-floor / kicked masks: bit-exact on every env-step whose decisions are
  further than 2e-5 from their thresholds in float64 (parity_util.MARGIN_EPS); at most 10 % of env-steps may be
  skipped for that reason (drones resting against each other / pillars sit near thresholds for many ticks)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

C2 = dict(num_agents=8, neighbor_visible_num=6, obs_repr='xyz_vxyz_R_omega', ep_time=1.0)
C3 = dict(num_agents=8, neighbor_visible_num=2, obs_repr='xyz_vxyz_R_omega_floor', use_obstacles=True,
          use_downwash=True, ep_time=1.0)
C4 = dict(num_agents=32, neighbor_visible_num=6, obs_repr='xyz_vxyz_R_omega', ep_time=0.6)
C1 = dict(num_agents=1, neighbor_visible_num=0, neighbor_obs_type='none', obs_repr='xyz_vxyz_R_omega', ep_time=0.8)
ALLN = dict(num_agents=5, neighbor_visible_num=-1, obs_repr='xyz_vxyz_R_omega_wall', ep_time=0.5, use_downwash=True)


def _run(kw, E, T, seed, **extra):
    from tests.parity_util import Pair, run_parity
    pair = Pair(E, kw, seed=seed, table_seed=seed + 1, rew_coeff=extra.pop('rew_coeff', None))
    rep = run_parity(pair, T, np.random.RandomState(seed + 2), **extra)
    frac = rep['skipped_env_steps'] / max(1, rep['skipped_env_steps'] + rep['compared_env_steps'])
    print(rep)
    assert frac < 0.10, rep
    pair.engine.close()
    return rep


def test_c1_single_drone():
    rep = _run(C1, E=24, T=200, seed=100)
    assert rep['dones'] >= 24 * 2


def test_c2_eight_drones_same_goal():
    rep = _run(C2, E=16, T=230, seed=200, rew_coeff=dict(quadcol_bin=5.0, quadcol_bin_smooth_max=10.0))
    assert rep['dones'] >= 16 * 2 and rep['floor'] > 0


def test_c3_obstacles_downwash_knn():
    rep = _run(C3, E=12, T=230, seed=300)
    assert rep['dones'] >= 12 * 2


def test_c4_thirtytwo_drones():
    _run(C4, E=3, T=70, seed=400)


def test_all_neighbors_wall_obs_non_pow2():
    _run(ALLN, E=9, T=120, seed=500, action_scale=1.3)


def _cluster_hook(center, spread, speed, every):
    """Every `every` steps plant the drones of each env in a tight cluster flying at each other (collisions,
    proximity penalties, downwash), by editing the ORACLE state and teacher-forcing the device from it."""
    def hook(pair, t):
        if t % every != 0:
            return
        rs = np.random.RandomState(1000 + t)
        for o in pair.oracles:
            c = np.array(center) + rs.uniform(-1, 1, 3) * [2, 2, 0.5]
            for d in o.drones:
                off = rs.uniform(-spread, spread, 3)
                off[2] = rs.uniform(-0.45, 0.45)
                d.pos = c + off
                d.vel = -speed * off / (np.linalg.norm(off) + 1e-9) + rs.uniform(-0.1, 0.1, 3)
                d.omega = rs.uniform(-1, 1, 3)
        pair.sync_device_from_oracle()
    return hook


def test_collisions_proximity_downwash_cluster():
    rep = _run(dict(num_agents=8, neighbor_visible_num=2, obs_repr='xyz_vxyz_R_omega', use_downwash=True, ep_time=2.0),
               E=10, T=150, seed=600, hook=_cluster_hook((0.0, 0.0, 3.0), 0.12, 0.6, 25),
               rew_coeff=dict(quadcol_bin=5.0, quadcol_bin_smooth_max=10.0))
    assert rep['quadcol'] > 10 and rep['kicked'] > 10, rep


def _room_hook(every):
    def hook(pair, t):
        if t % every != 0:
            return
        rs = np.random.RandomState(2000 + t)
        for o in pair.oracles:
            for i, d in enumerate(o.drones):
                kind = (i + t // every) % 5
                if kind == 0:
                    d.pos = np.array([4.99, rs.uniform(-3, 3), 3.0]); d.vel = np.array([3.0, 0.2, 0.1])
                elif kind == 1:
                    d.pos = np.array([rs.uniform(-3, 3), -4.99, 2.0]); d.vel = np.array([0.3, -3.0, 0.0])
                elif kind == 2:
                    d.pos = np.array([rs.uniform(-3, 3), rs.uniform(-3, 3), 9.99]); d.vel = np.array([0.1, 0.0, 4.0])
                elif kind == 3:
                    d.pos = np.array([rs.uniform(-3, 3), rs.uniform(-3, 3), 0.07]); d.vel = np.array([0.5, 0.3, -2.0])
                else:
                    d.pos = np.array([rs.uniform(-3, 3), rs.uniform(-3, 3), 0.07]); d.vel = np.array([-0.2, 0.4, -2.0])
                    c, s = np.cos(np.pi - 0.2), np.sin(np.pi - 0.2)
                    d.rot = d.rot @ np.array([[1., 0, 0], [0, c, -s], [0, s, c]])      # upside down -> random yaw branch
                d.on_floor = False
        pair.sync_device_from_oracle()
    return hook


def test_room_contacts_walls_ceiling_floor():
    rep = _run(dict(num_agents=5, neighbor_visible_num=2, obs_repr='xyz_vxyz_R_omega', ep_time=2.0),
               E=8, T=120, seed=700, hook=_room_hook(30))
    assert rep['kicked'] > 5 and rep['floor'] > 0, rep


def _obst_hook(every):
    def hook(pair, t):
        if t % every != 0 or t == 0:
            return
        rs = np.random.RandomState(3000 + t)
        for o in pair.oracles:
            for k, d in enumerate(o.drones[:4]):
                ob = o.obst_xy[k]
                ang = rs.uniform(-np.pi, np.pi)
                r = 0.3 + 0.046 + 0.02 if k else 0.15
                d.pos = np.array([ob[0] + r * np.cos(ang), ob[1] + r * np.sin(ang), rs.uniform(1.0, 3.0) if k else 5.0])
                d.vel = np.array([-1.5 * np.cos(ang), -1.5 * np.sin(ang), 0.1])
        pair.sync_device_from_oracle()
    return hook


def test_obstacle_collisions_and_sdf():
    rep = _run(dict(C3, ep_time=2.0), E=8, T=130, seed=800, hook=_obst_hook(20))
    assert rep['obstcol'] > 5, rep


def test_episode_stats_latch_matches_oracle():
    """The statistics latched at episode end equal the oracle's episode_extra_stats (quadrotor_multi.py:626-718)."""
    from tests.parity_util import Pair
    from quad_swarm_rl_b200 import _lib as L
    import torch
    kw = dict(C3, ep_time=0.4)
    pair = Pair(6, kw, seed=900, table_seed=901)
    pair.reset()
    rs = np.random.RandomState(902)
    stats = None
    for t in range(41):
        d, o = pair.step(rs.uniform(-1, 1, (6, 8, 4)).astype(np.float32))
        if (t + 1) % 10 == 0 and not o['dones'].any():
            pair.sync_device_from_oracle()
        if o['dones'].any():
            stats = o['infos']
    assert stats is not None
    es, ags = pair.engine.episode_stats()
    es, ags = es.cpu().numpy(), ags.cpu().numpy()
    for e in range(6):
        s0 = stats[e][0]['episode_extra_stats']
        for k, key in enumerate(L.ENV_STAT_KEYS[:11]):
            assert es[e, k] == s0[key], (e, key, es[e, k], s0[key])
        for i in range(8):
            si = stats[e][i]['episode_extra_stats']
            np.testing.assert_allclose(ags[e, i, 0], si['distance_to_goal_1s'], rtol=1e-4)
            np.testing.assert_allclose(ags[e, i, 1], si['distance_to_goal_3s'], rtol=1e-4)
            np.testing.assert_allclose(ags[e, i, 2], si['distance_to_goal_5s'], rtol=1e-4)
    pair.engine.close()


def test_device_side_o_random_generator_matches_twin():
    """QS_SCENARIO_O_RANDOM: episodes generated inside the reset path of the kernels equal oracle/scenario_gen.py
    (same keyed draws), and the whole trajectory stays in parity across auto-resets (no host tables involved)."""
    import torch
    from oracle import quadswarm_oracle as qo
    from oracle.scenario_gen import DeviceORandomSource
    from quad_swarm_rl_b200.engine import QuadSwarmEngine
    from tests import parity_util as pu
    kw = dict(C3, ep_time=0.5)
    E, seed = 10, 4321

    class DevPair(pu.Pair):
        def __init__(self):
            self.E, self.kw, self.N = E, dict(kw), 8
            self.engine = QuadSwarmEngine(num_envs=E, seed=seed, device_scenario='o_random', **kw)
            self.ocfg = pu.cfg_to_oracle(kw)
            self.oracles = [qo.OracleEnv(self.ocfg, qo.PhiloxRng(seed), DeviceORandomSource(), env_id=e) for e in range(E)]
            self.table_idx = 0

        def _push_table(self, k):
            pass

    pair = DevPair()
    rep = pu.run_parity(pair, 130, np.random.RandomState(5), resync=20)
    assert rep['dones'] >= 2 * E
    st = pair.engine.get_state()
    obst_dev = st['obst_xy'].cpu().numpy()
    goals_dev = st['agent_f32'][..., 30:33].cpu().numpy()
    for e, o in enumerate(pair.oracles):
        assert np.array_equal(obst_dev[e], o.obst_xy.astype(np.float32))
        np.testing.assert_allclose(goals_dev[e], np.array([d.goal for d in o.drones]), rtol=1e-6)
    pair.engine.close()


@pytest.mark.parametrize('scenario', ['o_static_same_goal', 'mix'])
def test_device_side_obstacle_scenarios_match_twin(scenario):
    """QS_SCENARIO_O_STATIC_SAME_GOAL and QS_SCENARIO_MIX with obstacles (o_random / o_static_same_goal drawn per episode,
    mix.py:45-57): pillars, spawn cells, the common goal above the largest free square and the per-scenario
    approch_goal_metric come from the kernels and equal the twin; reached-goal flags are compared bit for bit."""
    from oracle.scenario_gen import DeviceORandomSource
    from tests import parity_util as pu
    kw = dict(C3, ep_time=0.5)
    E = 10
    pair = pu.DevicePair(E, kw, 8642, scenario, lambda: DeviceORandomSource(scenario=scenario))
    seen, reached = set(), 0

    def hook(p, t):
        nonlocal reached
        if t > 0 and t % 17 == 0:
            fl = p.device_fields()['flags']
            for e, o in enumerate(p.oracles):
                if o.tick >= 6:
                    assert np.array_equal((fl[e] & pu.L.FLAG_REACHED_GOAL) != 0, np.asarray(o.reached_goal, bool)), (t, e)
                    reached += int(np.sum(o.reached_goal))
                seen.add(o.source.mode)

    rep = pu.run_parity(pair, 160, np.random.RandomState(6), resync=20, hook=hook)
    assert rep['dones'] >= 3 * E
    st = pair.engine.get_state()
    obst_dev = st['obst_xy'].cpu().numpy()
    goals_dev = st['agent_f32'][..., 30:33].cpu().numpy()
    for e, o in enumerate(pair.oracles):
        assert np.array_equal(obst_dev[e], o.obst_xy.astype(np.float32))
        np.testing.assert_allclose(goals_dev[e], np.array([d.goal for d in o.drones]), rtol=1e-6)
    es, _ = pair.engine.episode_stats()
    ids = set(es[:, 12].cpu().numpy().tolist())
    assert ids <= {1, 11} and (scenario == 'mix' or ids == {11})
    assert seen == ({11} if scenario == 'o_static_same_goal' else {1, 11})
    print('reached-goal flags compared while set:', reached)
    pair.engine.close()


@pytest.mark.parametrize('scenario,n', [('o_dynamic_same_goal', 8), ('o_swap_goals', 8), ('o_ep_rand_bezier', 8), ('o_ep_rand_bezier', 5),
                                        ('o_swap_goals', 3)])
def test_device_side_ticked_obstacle_scenarios_match_twin(scenario, n):
    """QS_SCENARIO_O_DYNAMIC_SAME_GOAL / O_SWAP_GOALS / O_EP_RAND_BEZIER (the evaluation scenarios of the obstacle family,
    scenarios/utils.py:18-20): pillars, spawn cells, goals and every goal event (hop to a free cell <= 4 m away, permutation,
    Bezier segments with the lane-parallel rejection search) happen inside the kernels and equal the twin, step by step,
    across the events at tick 1 / 4-6 s / 6 s and an auto-reset."""
    from oracle.scenario_gen import DeviceORandomSource
    from tests import parity_util as pu
    kw = dict(C3, ep_time=6.3, num_agents=n)
    E = 3
    pair = pu.DevicePair(E, kw, 8642 + n, scenario, lambda: DeviceORandomSource(scenario=scenario))
    rep = pu.run_parity(pair, 660, np.random.RandomState(6), resync=20)
    assert rep['dones'] >= E
    assert all(o.source.events >= 1 for o in pair.oracles)
    st = pair.engine.get_state()
    goals_dev = st['agent_f32'][..., 30:33].cpu().numpy()
    for e, o in enumerate(pair.oracles):
        np.testing.assert_allclose(goals_dev[e], np.array([d.goal for d in o.drones]), rtol=1e-5, atol=1e-5)
    es, _ = pair.engine.episode_stats()
    assert {int(x) for x in es[:, 12].cpu().numpy()} == {pu.L.DEVICE_SCENARIOS[scenario]}
    pair.engine.close()


DEVICE_FAMILY = ['static_same_goal', 'static_diff_goal', 'dynamic_same_goal', 'dynamic_diff_goal', 'swap_goals',
                 'dynamic_formations', 'ep_lissajous3D', 'swarm_vs_swarm', 'mix', 'ep_rand_bezier', 'run_away']


@pytest.mark.parametrize('mode', DEVICE_FAMILY)
def test_device_side_scenario_family_matches_twin(mode):
    """QS_SCENARIO_STATIC_SAME_GOAL .. QS_SCENARIO_MIX, EP_RAND_BEZIER, RUN_AWAY (an event every second): formation picks, goal formations and the timed / per-tick goal
    changes happen inside the kernels and equal oracle/scenario_gen.py (same keyed draws); the trajectory — goals
    included, compared every step — stays in parity across goal events and auto-resets."""
    from oracle.scenario_gen import DeviceScenarioSource
    from tests import parity_util as pu
    periodic = mode in ('dynamic_same_goal', 'dynamic_diff_goal', 'swap_goals', 'swarm_vs_swarm', 'ep_rand_bezier')   # events at 4-6 s / every 5 s
    kw = dict(num_agents=8, obs_repr='xyz_vxyz_R_omega', neighbor_visible_num=3, ep_time=6.3 if periodic else 1.2)
    E = 3 if periodic else 4
    pair = pu.DevicePair(E, kw, 97531, mode, lambda: DeviceScenarioSource(mode))
    T = 660 if periodic else (400 if mode == 'mix' else 260)
    rep = pu.run_parity(pair, T, np.random.RandomState(11), resync=20)
    assert rep['dones'] >= E
    if mode not in ('static_same_goal', 'static_diff_goal', 'mix'):     # every env saw goal events (period <= 599 ticks)
        assert all(o.source.events >= 1 for o in pair.oracles)
    es, _ = pair.engine.episode_stats()
    names = {int(x) for x in es[:, 12].cpu().numpy()}
    assert names <= (set(range(2, 10)) | {12, 16}) and (mode == 'mix' or names == {pu.L.DEVICE_SCENARIOS[mode]})
    assert mode != 'mix' or 16 not in names                                  # run_away is not one of mix's modes (utils.py:7-10)
    pair.engine.close()


def test_device_side_swarm_vs_swarm_32_drones_split_and_single(monkeypatch):
    """c4's scenario on the device at N = 32 (two 16-drone formations), in both kernel shapes; identical outputs."""
    import torch
    from oracle.scenario_gen import DeviceScenarioSource
    from tests import parity_util as pu
    kw = dict(num_agents=32, obs_repr='xyz_vxyz_R_omega', neighbor_visible_num=6, ep_time=0.3)
    outs = []
    for split in ('0', '1'):
        monkeypatch.setenv('QS_SPLIT', split)
        pair = pu.DevicePair(2, kw, 2468, 'swarm_vs_swarm', lambda: DeviceScenarioSource('swarm_vs_swarm'))
        rep = pu.run_parity(pair, 40, np.random.RandomState(3), resync=10)
        assert rep['dones'] >= 2
        outs.append(pair.engine.obs.clone())
        pair.engine.close()
    # same function, two instantiations: equal up to fp32 contraction differences over the <= 10 steps since the last
    # teacher-forcing point
    diff = (outs[0] - outs[1]).abs().max().item()
    print('split vs single max |diff|', diff)
    assert diff < 1e-4


@pytest.mark.parametrize('split', ['0', '1'])
def test_both_step_kernels_against_oracle(split, monkeypatch):
    """The single-warp kernel (QS_SPLIT=0) and the physics/observer split kernel (QS_SPLIT=1) are the same function:
    both must match the oracle on c3 with planted pillar contacts (kicks exercise the observer's re-build path), and
    their outputs agree with each other to fp32 rounding."""
    import torch
    from tests.parity_util import Pair
    monkeypatch.setenv('QS_SPLIT', split)
    rep = _run(dict(C3, ep_time=0.6), E=6, T=90, seed=1100, hook=_obst_hook(15))
    assert rep['obstcol'] > 3 and rep['dones'] >= 6
    outs = []
    for mode in ('0', '1'):
        monkeypatch.setenv('QS_SPLIT', mode)
        pair = Pair(5, dict(C3, ep_time=0.3), seed=77, table_seed=78)
        pair.engine.reset()
        g = torch.Generator(device='cuda'); g.manual_seed(3)
        a = (torch.rand((50, 5, 8, 4), device='cuda', generator=g) * 2 - 1).contiguous()
        outs.append([x.clone() for x in pair.engine.rollout(a)] + [pair.engine.get_state()['agent_f32'].clone()])
        pair.engine.close()
    # two instantiations of one function: equal up to fp32 contraction differences, which the dynamics amplify slowly
    obs0, obs1 = outs[0][0], outs[1][0]
    early = (obs0[:5] - obs1[:5]).abs().max().item()
    late = (obs0 - obs1).abs().max().item()
    print('single-warp vs split kernel: max |diff| first 5 steps', early, 'all 50 steps', late)
    assert early < 2e-5 and late < 5e-3
    assert torch.equal(outs[0][2], outs[1][2])                   # dones


def test_wide_observation_unstaged_path_32_drones_all_neighbours():
    """N = 32 with all 31 neighbours observed: D = 18 + 186 = 204 floats, too wide for the shared-memory staging tile,
    so rows are written straight to global memory (and the split kernel is not used)."""
    rep = _run(dict(num_agents=32, neighbor_visible_num=-1, obs_repr='xyz_vxyz_R_omega', ep_time=0.3), E=2, T=40, seed=1200)
    assert rep['dones'] >= 2


def test_dense_pillars_more_than_32_and_no_sensor_noise():
    """density 0.8 -> 51 pillars per env: exercises the tail scan of the SDF pass beyond the 32-bit candidate mask and
    many simultaneous pillar contacts; sense_noise=None bypasses the sensor noise (SensorNoise(bypass=True))."""
    kw = dict(num_agents=4, neighbor_visible_num=2, obs_repr='xyz_vxyz_R_omega_floor', use_obstacles=True, obst_density=0.8,
              use_downwash=False, ep_time=0.5, sense_noise=None)
    from tests import parity_util as pu


    def dense_tables(rs, E, N, M, use_obst, episodes=3, spread=1.5):
        cells = pu.qo.get_cell_centers(8, 8)
        eps = []
        for _ in range(episodes):
            goals = rs.uniform(-3, 3, (E, N, 3)).astype(np.float32); goals[..., 2] = rs.uniform(1, 3, (E, N))
            spawn = np.zeros((E, N, 3), np.float32); obst = np.zeros((E, M, 2), np.float32)
            for e in range(E):
                idx = rs.permutation(64)
                obst[e] = cells[idx[:M]]
                spawn[e, :, :2] = cells[idx[M:M + N]]
                spawn[e, :, 2] = rs.uniform(1.0, 3.0, N)
            eps.append(dict(goals=goals, spawn=spawn, obst=obst))
        return eps

    orig = pu.make_tables
    pu.make_tables = dense_tables
    try:
        rep = _run(kw, E=6, T=110, seed=1300)
    finally:
        pu.make_tables = orig
    assert rep['dones'] >= 6


def test_two_drones_one_env_minimal():
    """smallest multi-drone case: E = 1, N = 2, K = 1 (all neighbours)."""
    _run(dict(num_agents=2, neighbor_visible_num=-1, obs_repr='xyz_vxyz_R_omega', ep_time=0.4), E=1, T=60, seed=1400)


# ---- tight mask parity: teacher forcing after EVERY step, thresholds resolved to single-step fp32 rounding ----
@pytest.mark.parametrize('name', ['c2', 'c3', 'cluster', 'room'])
def test_masks_bit_exact_on_all_but_one_percent_of_env_steps(name):
    """Collision / pillar-collision / on-floor / kicked / done masks, bit for bit, with at most 1 % of the env-steps left
    out: the device is re-synchronised from the oracle after every step, so the two sides differ by ONE step of fp32
    arithmetic (position error <= ulp(5 m) = 4.8e-7) and only decisions closer than 3e-6 to their threshold are skipped
    (drones resting on each other or on a pillar sit exactly there).  The skipped fraction is printed."""
    from tests.parity_util import Pair, run_parity
    hook, kw, E, T, extra = None, None, 10, 110, {}
    if name == 'c2':
        kw, extra = C2, dict(rew_coeff=dict(quadcol_bin=5.0, quadcol_bin_smooth_max=10.0))
    elif name == 'c3':
        kw = C3
    elif name == 'cluster':
        kw = dict(num_agents=8, neighbor_visible_num=2, obs_repr='xyz_vxyz_R_omega', use_downwash=True, ep_time=2.0)
        hook, extra = _cluster_hook((0.0, 0.0, 3.0), 0.12, 0.6, 25), dict(rew_coeff=dict(quadcol_bin=5.0, quadcol_bin_smooth_max=10.0))
    else:
        kw = dict(num_agents=6, neighbor_visible_num=2, obs_repr='xyz_vxyz_R_omega', ep_time=1.0)
        hook = _room_hook(30)
    pair = Pair(E, kw, seed=4242, table_seed=4243, rew_coeff=extra.get('rew_coeff'))
    rep = run_parity(pair, T, np.random.RandomState(4244), resync=1, margin_eps=3e-6, gap_eps=5e-6, hook=hook)
    frac = rep['skipped_env_steps'] / max(1, rep['skipped_env_steps'] + rep['compared_env_steps'])
    print(f'{name}: skipped {rep["skipped_env_steps"]} of {rep["skipped_env_steps"] + rep["compared_env_steps"]} env-steps '
          f'({100 * frac:.2f} %), quadcol {rep["quadcol"]} obstcol {rep["obstcol"]} kicked {rep["kicked"]} floor {rep["floor"]}')
    # 'room' plants drones ON walls, ceiling and floor: a drone resting on a surface sits exactly on that threshold for as
    # long as it rests, which no finite precision separates — those env-steps are the ones left out (< 2 %)
    assert frac <= (0.02 if name == 'room' else 0.01), rep
    pair.engine.close()


# ---- full-size sampled parity: the BENCHMARKED grid shapes / kernel instantiations / launch chaining against the oracle ----
@pytest.mark.parametrize('name,chained', [('c3', False), ('c3', True), ('c2', True), ('c4', True), ('c4', False)])
def test_full_size_sampled_parity(name, chained):
    """BASELINE shapes (c2 8x1024, c3 8x4096, c4 32x2048), episodes generated on the device as in bench.py.  The keyed
    draws make env e a function of (seed, global env id, actions) only, so three oracle envs with ids {0, E/2, E-1} check
    the engine's rows 0, E/2, E-1 while all the other envs run beside them — i.e. the kernel instantiation, grid shape
    and (chained = True: prefetch across the dependency wait, per-block hand-over for c2 / c4) launch chaining that the
    benchmark times, not a 12-env stand-in."""
    from oracle.scenario_gen import DeviceORandomSource, DeviceScenarioSource
    from tests import parity_util as pu
    if name == 'c3':
        E, kw, scn, fac = 4096, dict(C3, ep_time=0.4), 'o_random', (lambda: DeviceORandomSource())
        T, rew = 100, dict(quadcol_bin=5.0, quadcol_bin_smooth_max=4.0, quadcol_bin_obst=5.0)
    elif name == 'c2':
        E, kw, scn, fac = 1024, dict(C2, ep_time=0.4), 'static_same_goal', (lambda: DeviceScenarioSource('static_same_goal'))
        T, rew = 100, dict(quadcol_bin=5.0, quadcol_bin_smooth_max=10.0)
    else:
        E, kw, scn, fac = 2048, dict(C4, ep_time=0.3), 'swarm_vs_swarm', (lambda: DeviceScenarioSource('swarm_vs_swarm'))
        T, rew = 45, dict(quadcol_bin=5.0, quadcol_bin_smooth_max=10.0)
    pair = pu.SampledPair(E, [0, E // 2, E - 1], kw, seed=24680, device_scenario=scn, source_factory=fac, chained=chained,
                          rew_coeff=rew)
    rep = pu.run_parity(pair, T, np.random.RandomState(12), resync=20)
    print(name, chained, rep)
    assert rep['dones'] >= 3 and rep['compared_env_steps'] >= 0.85 * 3 * T, rep
    assert pair.engine.handover_timeouts == 0
    pair.engine.close()


# ---- per-drone physical constants (other models, dynamics randomisation, rotor drag) ----
DYN_KINDS = ('defaultquad', 'mediumquad', 'randomquad', 'rotor_drag', 'relative', 'zoo')


def _dyn_rows(E, N, seed, kind='zoo'):
    """Airframes: the named models, RandomQuad draws, a Crazyflie with rotor drag / rolling moment, non-unit motor
    linearity and velocity / body-rate damping, perturbed Crazyflies — one row per drone (quad_models.constants_row);
    'zoo' mixes all of them inside every env."""
    from quad_swarm_rl_b200 import quad_models as qm
    rs = np.random.RandomState(seed)
    rows = np.zeros((E, N, qm.DYN_ROW), np.float32)
    for e in range(E):
        for i in range(N):
            k = DYN_KINDS[(e * N + i) % 5] if kind == 'zoo' else kind
            if k == 'defaultquad':
                p = qm.defaultquad_params()
            elif k == 'mediumquad':
                p = qm.mediumquad_params()
            elif k == 'randomquad':
                p = qm.RandomQuad().sample(rs=rs)
            elif k == 'rotor_drag':
                p = qm.crazyflie_params()
                p['motor'].update(C_drag=0.0028 * rs.uniform(0.5, 20), C_roll=0.003 * rs.uniform(0.5, 5), linearity=0.6)
                p['damp'].update(vel=0.001, omega_quadratic=0.01)
            else:
                p = qm.RelativeSampler(qm.crazyflie_params(), noise_ratio=0.2, sampler='uniform').sample(qm.crazyflie_params(), rs)
            qm.check_quad_param_limits(p)
            rows[e, i] = qm.constants_row(p)
    return rows


def _params_of(rows):
    from oracle import quadswarm_oracle as qo
    from quad_swarm_rl_b200.quad_models import DYN_FIELDS
    return [[qo.quad_params_from_constants(dict(zip(DYN_FIELDS, r.astype(np.float64)))) for r in env_rows] for env_rows in rows]


@pytest.mark.parametrize('kind', DYN_KINDS)
@pytest.mark.parametrize('kw_name', ['free', 'obstacles'])
def test_per_drone_dynamics_models_match_oracle(kw_name, kind):
    """qs_set_dynamics: every drone flies its own airframe (DefaultQuad, MediumQuad, RandomQuad draws, perturbed Crazyflies,
    rotor drag + rolling moment with linearity 0.6 and velocity / body-rate damping; 'zoo' = all of them in one env).
    Half-way the constants of every env are replaced with at_next_reset = True: the kernel latches them at the env's
    auto-reset (OU state and SVD counter restart), the oracle does the same through its dyn_source hook
    (quadrotor_single.py:387-390)."""
    from tests import parity_util as pu
    if kw_name == 'free':
        kw = dict(num_agents=5, neighbor_visible_num=2, obs_repr='xyz_vxyz_R_omega_wall', ep_time=0.6, use_downwash=True)
    else:
        kw = dict(C3, ep_time=0.5, num_agents=4)
    E, N = 6, kw['num_agents']
    rows0, rows1 = _dyn_rows(E, N, 1, kind), _dyn_rows(E, N, 2, kind)
    pair = pu.Pair(E, kw, seed=777, table_seed=778)
    pair.engine.set_dynamics(rows0)
    P0, P1 = _params_of(rows0), _params_of(rows1)
    pending = [[False] * N for _ in range(E)]
    for e, o in enumerate(pair.oracles):
        o.Ps = list(P0[e])
        o.P = o.Ps[0]

        def src(i, e=e):
            if pending[e][i]:
                pending[e][i] = False              # latched once, like the kernel's dyn_pending word
                return P1[e][i]
            return None
        o.dyn_source = src

    def hook(p, t):
        if t == 40:                                 # uploaded mid-episode: nothing changes before the next reset
            p.engine.set_dynamics(rows1, at_next_reset=True)
            for e in range(E):
                pending[e] = [True] * N
    rep = pu.run_parity(pair, 150, np.random.RandomState(3), resync=20, hook=hook)
    print(kind, rep)
    assert rep['dones'] >= 2 * E
    frac = rep['skipped_env_steps'] / max(1, rep['skipped_env_steps'] + rep['compared_env_steps'])
    assert frac < 0.10, rep
    pair.engine.close()


def test_per_episode_obstacle_density_and_size_randomisation():
    """qs_set_obstacle_randomization (ExperienceReplayWrapper's domain randomisation -> reset(obst_density, obst_size),
    quad_experience_replay.py:108-118, quadrotor_multi.py:339-351): every episode of every env draws its pillar count and
    pillar size on the device; collision test, contact response and the 3x3 distance patch use the episode's radius.  The
    twin (oracle/scenario_gen.py) makes the same draws; trajectories stay in parity across auto-resets."""
    from oracle.scenario_gen import DeviceORandomSource
    from tests import parity_util as pu
    dens = [0.05, 0.1, 0.15000000000000002, 0.2]           # np.arange(0.05, 0.25, 0.05)
    sizes = [0.3, 0.4, 0.5, 0.6000000000000001, 0.7000000000000002]
    kw = dict(C3, ep_time=0.4)
    E = 10
    pair = pu.DevicePair(E, kw, 1357, 'o_random', lambda: DeviceORandomSource(densities=dens, sizes=sizes))
    pair.engine.set_obstacle_randomization(dens, sizes)
    seen_m, seen_r = set(), set()

    def hook(p, t):
        if t % 41 == 5:
            st = p.engine.get_state()
            base = 4 + pu.L.QS_NUM_ENV_STATS
            rad_m = st['env_i32'][:, base + 4:base + 6].cpu().numpy().view(np.float32)
            for e, o in enumerate(p.oracles):
                assert rad_m[e, 0] == np.float32(o.obst_size / 2) and int(rad_m[e, 1]) == len(o.obst_xy) == o.source.num_pillars
                ob = st['obst_xy'][e].cpu().numpy()
                assert np.array_equal(ob[:len(o.obst_xy)], o.obst_xy.astype(np.float32)) and (ob[len(o.obst_xy):] == 1.0e4).all()
                seen_m.add(len(o.obst_xy)); seen_r.add(float(rad_m[e, 0]))
    rep = pu.run_parity(pair, 170, np.random.RandomState(8), resync=20, hook=hook)
    print(rep, sorted(seen_m), sorted(seen_r))
    assert rep['dones'] >= 4 * E and len(seen_m) >= 3 and len(seen_r) >= 3
    frac = rep['skipped_env_steps'] / max(1, rep['skipped_env_steps'] + rep['compared_env_steps'])
    assert frac < 0.10, rep
    pair.engine.close()
