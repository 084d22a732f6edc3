"""The product's physical models and dynamics-randomisation samplers (quad_swarm_rl_b200/quad_models.py) against the
reference's (fixtures: tests/golden/dyn_models.json, generator oracle/gen_golden_dyn.py): same numpy draws in the same
order under the same seed -> the same derived constants (mass, inertia, thrust / torque limits, propeller positions, motor
time constants, OU sigma) to float64 rounding.  The end-to-end pin (samplers running on the reference's own random stream
inside a replayed trajectory) is tests/test_oracle_vs_reference.py::...[randomquad_relsampler_5]."""
import json
import os

import numpy as np
import pytest

from quad_swarm_rl_b200 import quad_models as qm

G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'dyn_models.json')))
RTOL = 1e-12


def _check(c, d):
    np.testing.assert_allclose(c['mass'], d['mass'], rtol=RTOL)
    np.testing.assert_allclose([c['ixx'], c['iyy'], c['izz']], d['inertia'], rtol=RTOL)
    np.testing.assert_allclose([c[f'thrust_max{m}'] for m in range(4)], d['thrust_max'], rtol=RTOL)
    np.testing.assert_allclose([c[f'torque_max{m}'] for m in range(4)], d['torque_max'], rtol=RTOL)
    pp = np.array([[c[f'px{m}'], c[f'py{m}'], c[f'pz{m}']] for m in range(4)])
    np.testing.assert_allclose(pp, np.array(d['prop_pos']), rtol=RTOL, atol=1e-18)
    np.testing.assert_allclose(np.stack([pp[:, 1], -pp[:, 0], 0 * pp[:, 0]], 1), np.array(d['prop_crossproducts']), rtol=RTOL, atol=1e-18)
    for a, b in (('arm', 'arm'), ('tau_up', 'motor_tau_up'), ('tau_down', 'motor_tau_down'), ('linearity', 'linearity'),
                 ('ou_sigma', 'ou_sigma'), ('c_drag', 'c_drag'), ('c_roll', 'c_roll'), ('vel_damp', 'vel_damp'),
                 ('omega_quadratic', 'omega_quadratic')):
        np.testing.assert_allclose(c[a], d[b], rtol=RTOL, err_msg=a)


@pytest.mark.parametrize('name', sorted(G['named']))
def test_named_models(name):
    _check(qm.derive_constants(getattr(qm, name)()), G['named'][name])


def test_crazyflie_row_matches_the_kernel_constants():
    """The default model's row must be the constants the kernels use when no dynamics table is uploaded."""
    c = qm.derive_constants(qm.crazyflie_params())
    ref = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'crazyflie_constants.json')))
    assert abs(c['mass'] - ref['mass']) < 1e-15 and abs(c['arm'] - ref['arm']) < 1e-15
    np.testing.assert_allclose([c['ixx'], c['iyy'], c['izz']], ref['inertia'], rtol=1e-13)
    assert abs(c['thrust_max0'] - ref['thrust_max']) < 1e-15 and abs(c['torque_max0'] - ref['torque_max']) < 1e-17
    row = qm.constants_row(qm.crazyflie_params())
    assert row.shape == (qm.DYN_ROW,) and row.dtype == np.float32 and qm.DYN_ROW % 4 == 0


@pytest.mark.parametrize('entry', G['randomquad'], ids=lambda e: f"seed{e['seed']}")
def test_randomquad(entry):
    rs = np.random.RandomState(entry['seed'])
    p = qm.RandomQuad().sample(rs=rs)
    qm.check_quad_param_limits(p)
    _check(qm.derive_constants(p), entry['derived'])


@pytest.mark.parametrize('entry', G['relative'], ids=lambda e: f"{e['sampler']}{e['ratio']}")
def test_relative_sampler(entry):
    base = qm.crazyflie_params()
    s = qm.RelativeSampler(base, noise_ratio=entry['ratio'], sampler=entry['sampler'])
    rs = np.random.RandomState(entry['seed'])
    for d in entry['draws']:
        p = s.sample(base, rs)
        qm.check_quad_param_limits(p)
        _check(qm.derive_constants(p), d)


def test_dynamics_source_pipeline():
    src = qm.DynamicsSource('RandomQuad', dict(noise=dict(thrust_noise_ratio=0.05), damp=dict(vel=0, omega_quadratic=0)),
                            {'class': 'RelativeSampler', 'noise_ratio': 0.05, 'sampler': 'normal'}, seed=3)
    rows = np.stack([src.sample_row() for _ in range(5)])
    assert rows.shape == (5, qm.DYN_ROW) and np.isfinite(rows).all()
    assert (rows[:, 0] > 0.001).all() and len(np.unique(rows[:, 0])) == 5          # masses differ per draw
    k = qm.DYN_FIELDS.index('ou_sigma')
    np.testing.assert_allclose(rows[:, k], np.float32(0.2 * 0.05) * (1 + 0 * rows[:, k]), rtol=0.2)   # dynamics_change then 5 % noise


@pytest.mark.parametrize('entry', G['rotor_drag'], ids=lambda e: f"seed{e['seed']}")
def test_oracle_rotor_drag_matches_reference_numpy_path(entry):
    """Rotor drag / rolling moment exist only in the reference's numpy path (step1, quadrotor_dynamics.py:256-289); the
    oracle's sub-step carries the term and reproduces 12 airborne reference sub-steps to float64 rounding."""
    from oracle import quadswarm_oracle as qo
    P = qo.quad_params_from_constants(dict(zip(qm.DYN_FIELDS, entry['constants'])))
    assert P.c_drag != 0 or P.c_roll != 0
    d = qo.Drone()
    for k, v in entry['init'].items():
        setattr(d, k, np.array(v))
    room = np.array([[-5, -5, 0], [5, 5, 10.]])
    for st in entry['steps']:
        qo.dynamics_substep(d, P, np.array(st['cmd']), np.array(st['noise']), room, None, 0, 0)
        for k in ('pos', 'vel', 'rot', 'omega', 'thrust_cmds_damp', 'thrust_rot_damp'):
            np.testing.assert_allclose(getattr(d, k), np.array(st[k]), rtol=1e-10, atol=1e-12, err_msg=k)
