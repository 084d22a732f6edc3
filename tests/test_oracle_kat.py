"""Known-answer tests for the CPU oracle: the reference's own unit-test vectors and the
vectors extracted from the reference in SURVEY.md Appendix E, plus Philox4x32-10 test vectors (Random123)."""
import json
import os

import numpy as np

from oracle import quadswarm_oracle as qo
from oracle import philox as px

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def test_constants_match_reference_dump():
    c = json.load(open(os.path.join(GOLDEN, 'crazyflie_constants.json')))
    P = qo.QuadParams()
    for k in ('mass', 'thrust_max', 'torque_max', 'motor_linearity', 'arm', 'motor_tau_up', 'motor_tau_down',
              'omega_max', 'damp_omega_quadratic', 'vel_damp', 'since_last_svd_limit', 'mu', 'dt', 'ou_theta',
              'ou_sigma', 'ou_mu', 'pos_norm_std', 'vel_norm_std', 'gyro_noise_density'):
        assert getattr(P, k) == c[k], k
    assert list(P.inertia) == c['inertia']
    assert [list(r) for r in P.prop_crossproducts] == c['prop_crossproducts']
    assert list(P.prop_ccw) == c['prop_ccw']
    assert c['C_rot_drag'] == 0 and c['C_rot_roll'] == 0
    cfg = qo.EnvConfig()
    env = qo.OracleEnv(cfg, qo.PhiloxRng(0), None)
    assert env.ep_len == c['ep_len_15s'] == 1500
    assert env.collision_threshold == c['collision_threshold']
    assert env.collision_falloff_threshold == c['collision_falloff_threshold']


def test_e1_dynamics_two_substeps_zero_noise():
    """SURVEY Appendix E1."""
    P = qo.QuadParams()
    d = qo.Drone()
    d.pos = np.array([0.5, -0.25, 2.0])
    d.vel = np.array([0.1, 0.2, -0.3])
    cz, sz, cx, sx = np.cos(0.3), np.sin(0.3), np.cos(0.1), np.sin(0.1)
    d.rot = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1.]]) @ np.array([[1., 0, 0], [0, cx, -sx], [0, sx, cx]])
    # the survey probe went through set_state, which stores omega as float32 (quadrotor_dynamics.py:188)
    d.omega = np.array([0.5, -0.4, 0.3], dtype=np.float32).astype(np.float64)
    cmd = np.array([0.6, 0.55, 0.5, 0.65])
    box = np.array([[-5., -5, 0], [5, 5, 10]])
    qo.dynamics_substep(d, P, cmd, np.zeros(4), box, None, 0, 0)
    np.testing.assert_allclose(d.pos, [0.5005, -0.249, 1.9985], rtol=1e-11)
    np.testing.assert_allclose(d.vel, [0.100026987843, 0.199906314443, -0.348102354736], rtol=1e-9)
    np.testing.assert_allclose(d.omega, [0.5005357222, -0.40457724056, 0.300287870033], rtol=1e-9)
    np.testing.assert_allclose(d.thrust_rot_damp, [0.103278867373, 0.098881987281, 0.094280275623, 0.107496053337], rtol=1e-9)
    np.testing.assert_allclose(d.thrust_cmds_damp, [0.010666524446, 0.009777647409, 0.008888770372, 0.011555401483], rtol=1e-9)
    qo.dynamics_substep(d, P, cmd, np.zeros(4), box, None, 0, 1)
    np.testing.assert_allclose(d.pos, [0.501000134939, -0.248000468428, 1.996759488226], rtol=1e-11)
    np.testing.assert_allclose(d.vel, [0.100117074233, 0.199569987014, -0.393851216307], rtol=1e-9)
    np.testing.assert_allclose(d.omega, [0.501078094637, -0.422198283249, 0.301226915247], rtol=1e-9)
    np.testing.assert_allclose(d.rot.flatten(), [0.954563781677, -0.296768367495, 0.027138952908, 0.297975536653,
                                                 0.949181115241, -0.101320235033, 0.004308859157, 0.104803370771,
                                                 0.994483628426], rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(d.thrust_cmds_damp, [0.037166947244, 0.03406970164, 0.030972456036, 0.040264192847], rtol=1e-9)
    np.testing.assert_allclose(d.acc, [0.018017278069, -0.067265485696, -9.149772314371], rtol=1e-8)


def test_e2_collision_matrix_and_proximity():
    """SURVEY Appendix E2; same construction as the reference's collisions/test/unit_test/quadrotor.py:6-51."""
    arm = qo.QuadParams().arm
    pos = np.array([[0, 0, 2.], [0.08, 0, 2.], [0.15, 0, 2.], [3, 3, 3.]])
    col, pairs, rows = qo.calculate_collision_matrix(pos, 2 * arm)
    assert list(col) == [1, 1, 1, -1000]
    assert pairs == [(0, 1), (1, 2)]
    np.testing.assert_allclose(rows[:, 2], [0.08, 0.15, 4.358898943541, 0.07, 4.3042304771, 4.257052971247], rtol=1e-10)
    near = rows[rows[:, 2] <= 4 * arm]
    pen = qo.calculate_drone_proximity_penalties(near, 4 * arm, 0.01, 10.0, 4)
    np.testing.assert_allclose(pen, [0.074896492559, 0.118410756017, 0.080335775492, 0], rtol=1e-9)
    # the reference's own unit test: 7 coincident drones + 1 far away, threshold 0.2
    pos = np.zeros((8, 3))
    pos[7] = [5., 5., 5.]
    col, pairs, rows = qo.calculate_collision_matrix(pos, 0.2)
    assert list(col) == [1] * 7 + [-1000]
    assert len(pairs) == 21 and all(j < 7 for _, j in pairs)


def test_e3_obstacle_sdf_and_detection():
    """SURVEY Appendix E3 (obstacles/test/unit_test.py:6-47 geometry)."""
    sdf = qo.get_surround_sdfs(np.array([[0., 0.]]), np.array([[0.5, 0.5], [-1.5, 2.5]]), 0.3)
    np.testing.assert_allclose(sdf[0], [0.548528137424, 0.481024967591, 0.421110255093, 0.481024967591,
                                        0.407106781187, 0.340312423743, 0.421110255093, 0.340312423743,
                                        0.265685424949], rtol=1e-10)
    det = qo.obstacle_collision_detection(np.array([[0.2, 0.5], [3., 3.]]), np.array([[0.5, 0.5], [-1.5, 2.5]]),
                                          0.3, qo.QuadParams().arm)
    assert list(det) == [0, -1]
    cc = qo.get_cell_centers(8, 8)
    assert cc.shape == (64, 2) and tuple(cc[0]) == (-3.5, 3.5) and tuple(cc[-1]) == (3.5, -3.5)


def test_e4_reference_obstacle_normal_kat():
    """collisions/test/unit_test/obstacles.py:6-18 — the one true KAT in the reference's tests."""
    vnew, normal = qo.compute_col_norm_and_new_vel_obst(np.zeros(3), np.array([1., 0., 0.]), np.array([0.5, 0.5, 5.]))
    np.testing.assert_allclose(vnew, -np.sqrt(2) / 2, rtol=1e-12)
    np.testing.assert_allclose(normal, [-np.sqrt(2) / 2, -np.sqrt(2) / 2, 0], rtol=1e-12)


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    assert px.philox4x32_10(0, 0, 0, 0, 0, 0) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    f = 0xffffffff
    assert px.philox4x32_10(f, f, f, f, f, f) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert px.philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_keyed_draw_statistics():
    d = px.KeyedDraws(1234, 5, 77)
    z = np.array([d.normal(px.SITE_SENSOR0, i, 0, v) for i in range(400) for v in range(9)])
    u = np.array([d.uniform(px.SITE_WALL_U, i, 0, v) for i in range(400) for v in range(11)])
    assert abs(z.mean()) < 0.06 and abs(z.std() - 1) < 0.05
    assert 0 <= u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.02


def test_device_o_random_twin_properties():
    """oracle/scenario_gen.py (twin of the device generator): distinct pillar cells, spawn / goal cells distinct and
    free, heights in [1, 3), every cell equally likely to hold a pillar (12/64)."""
    from oracle import scenario_gen as sg
    counts = np.zeros((8, 8))
    for env in range(400):
        d = px.KeyedDraws(99, env, 7)
        goals, spawn, obst = sg.o_random_episode(d, 8, 12, 8, 8)
        pil = {tuple(o) for o in obst}
        assert len(pil) == 12
        for pts in (goals, spawn):
            xy = [tuple(p[:2]) for p in pts]
            assert len(set(xy)) == 8 and not (set(xy) & pil)
            assert np.all(pts[:, 2] >= 1.0) and np.all(pts[:, 2] < 3.0)
        for o in obst:
            counts[int(o[0] + 3.5), int(o[1] + 3.5)] += 1
    assert abs(counts.mean() - 400 * 12 / 64) < 1e-9 and counts.min() > 40 and counts.max() < 115
    cells = qo.get_cell_centers(8, 8)
    assert {tuple(c) for c in cells} == {tuple(sg._cell_center(k, 8, 8)) for k in range(64)}
