"""Golden vectors of the reference's physical models and dynamics-randomisation samplers — TEST INFRASTRUCTURE ONLY.

Run in a container where /root/reference is mounted:  python -m oracle.gen_golden_dyn
Writes tests/golden/dyn_models.json: for the named parameter sets and for seeded RandomQuad / RelativeSampler draws, the
constants `QuadrotorDynamics.update_model` derives (quadrotor_dynamics.py:104-166) — what quad_swarm_rl_b200/quad_models.py
must reproduce (same numpy draws in the same order under the same seed).
"""
import contextlib
import copy
import io
import json
import os

import numpy as np

from . import ref_harness as rh


def derived(params, use_numba=True):
    from gym_art.quadrotor_multi.quadrotor_dynamics import QuadrotorDynamics
    with contextlib.redirect_stdout(io.StringIO()):
        d = QuadrotorDynamics(model_params=copy.deepcopy(params), dynamics_steps_num=2, room_box=np.array([[-5, -5, 0], [5, 5, 10.]]),
                              dim_mode='3D', use_numba=use_numba, dt=0.005)
    return dict(mass=float(d.mass), inertia=[float(x) for x in d.inertia], thrust_max=[float(x) for x in d.thrust_max],
                torque_max=[float(x) for x in d.torque_max], prop_pos=np.asarray(d.model.prop_pos).tolist(),
                prop_crossproducts=np.asarray(d.prop_crossproducts).tolist(), arm=float(d.arm),
                motor_tau_up=float(d.motor_tau_up), motor_tau_down=float(d.motor_tau_down), linearity=float(d.motor_linearity),
                ou_sigma=float(d.thrust_noise.sigma), c_drag=float(d.C_rot_drag), c_roll=float(d.C_rot_roll),
                vel_damp=float(d.vel_damp), omega_quadratic=float(d.damp_omega_quadratic))


def main():
    rh._ensure_path()
    from gym_art.quadrotor_multi import quad_models as qm
    from gym_art.quadrotor_multi import quadrotor_randomization as qr
    out = {'named': {}, 'randomquad': [], 'relative': []}
    for name in ('crazyflie_params', 'defaultquad_params', 'mediumquad_params', 'crazyflie_lowinertia_params'):
        out['named'][name] = derived(getattr(qm, name)())
    for seed in (1, 2, 3, 4, 5, 6):
        np.random.seed(seed)
        p = qr.RandomQuad().sample()
        qr.check_quad_param_limits(p)
        out['randomquad'].append(dict(seed=seed, derived=derived(p)))
    for seed, sampler, ratio in ((11, 'normal', 0.1), (12, 'uniform', 0.2), (13, 'normal', 0.3)):
        base = qm.crazyflie_params()
        rs_ = qr.RelativeSampler(base, noise_ratio=ratio, sampler=sampler)
        np.random.seed(seed)
        draws = []
        for _ in range(3):
            p = rs_.sample(base)
            qr.check_quad_param_limits(p)
            draws.append(derived(p))
        out['relative'].append(dict(seed=seed, sampler=sampler, ratio=ratio, draws=draws))
    # rotor drag / rolling moment: only the reference's numpy path has the term (step1, quadrotor_dynamics.py:256-289);
    # airborne states, so that the two paths' different floor code is not involved
    from gym_art.quadrotor_multi.quadrotor_dynamics import QuadrotorDynamics
    from quad_swarm_rl_b200.quad_models import DYN_FIELDS
    out['rotor_drag'] = []
    for seed, cd, cr in ((21, 0.0028, 0.003), (22, 0.05, 0.0), (23, 0.0, 0.02)):
        params = qm.crazyflie_params()
        params['motor']['C_drag'], params['motor']['C_roll'] = cd, cr
        with contextlib.redirect_stdout(io.StringIO()):
            d = QuadrotorDynamics(model_params=copy.deepcopy(params), dynamics_steps_num=2,
                                  room_box=np.array([[-5, -5, 0], [5, 5, 10.]]), dim_mode='3D', use_numba=False, dt=0.005)
        rs = np.random.RandomState(seed)
        from gym_art.quadrotor_multi.quad_utils import rand_uniform_rot3d
        np.random.seed(seed)
        rot = rand_uniform_rot3d()
        d.set_state(rs.uniform(-2, 2, 3) + np.array([0, 0, 4.]), rs.uniform(-2.5, 2.5, 3), rot, rs.uniform(-6, 6, 3))
        d.thrust_rot_damp, d.thrust_cmds_damp = rs.uniform(0.2, 0.9, 4), rs.uniform(0.2, 0.8, 4)
        d.omega = np.float64(d.omega)     # set_state stores float32: the numpy path would then do its first update partly in float32
        rows = rh.dynamics_rows(type('E', (), {'envs': [type('S', (), {'dynamics': d})()]})())[0]
        init = dict(pos=d.pos.tolist(), vel=d.vel.tolist(), rot=d.rot.tolist(), omega=np.float64(d.omega).tolist(),
                    thrust_rot_damp=d.thrust_rot_damp.tolist(), thrust_cmds_damp=d.thrust_cmds_damp.tolist())
        steps = []
        for _ in range(12):
            cmd, noise = rs.uniform(0, 1, 4), 0.01 * rs.standard_normal(4)
            d.step1(cmd, 0.005, noise)
            steps.append(dict(cmd=cmd.tolist(), noise=noise.tolist(), pos=d.pos.tolist(), vel=d.vel.tolist(), rot=d.rot.tolist(),
                              omega=np.float64(d.omega).tolist(), thrust_cmds_damp=d.thrust_cmds_damp.tolist(),
                              thrust_rot_damp=d.thrust_rot_damp.tolist()))
        out['rotor_drag'].append(dict(seed=seed, constants=[rows[k] for k in DYN_FIELDS], init=init, steps=steps))
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'dyn_models.json')
    json.dump(out, open(path, 'w'), indent=1)
    print('wrote', path)


if __name__ == '__main__':
    main()
