"""Golden goal formations from the UNMODIFIED reference — TEST INFRASTRUCTURE ONLY.

Calls `QuadrotorScenario.generate_goals` (gym_art/quadrotor_multi/scenarios/base.py:39-113) and the formation-range
helpers (scenarios/utils.py:92-146,163-175) of the reference for every formation and swarm size the device-side
generators support (1..32 drones), and stores the results in tests/golden/formations.npz.  The CPU twin of the
kernels' formation code (oracle/scenario_gen.py) and the host generators (quad_swarm_rl_b200/scenarios.py) are checked
against this file, so the fixture travels to the GPU box while the reference does not.

Run here (needs /root/reference):  python -m oracle.gen_golden_formations
"""
import os

import numpy as np

from . import ref_harness


def main():
    ref_harness._ensure_path()
    from gym_art.quadrotor_multi.scenarios.base import QuadrotorScenario
    from gym_art.quadrotor_multi.scenarios.utils import QUADS_FORMATION_LIST, get_formation_range, get_z_value, \
        QUADS_PARAMS_DICT
    out = {}
    size, layer, center = 0.37, 0.41, np.array([0.1, -0.2, 2.0])
    for f, name in enumerate(QUADS_FORMATION_LIST):
        for n in range(1, 33):
            sc = QuadrotorScenario('static_diff_goal', [], n, [10., 10., 10.])
            sc.formation = name
            sc.num_agents_per_layer = 50 if name.startswith('grid') else 8
            sc.formation_size = size
            goals = np.array(sc.generate_goals(n, formation_center=center, layer_dist=layer), dtype=np.float64)
            out[f'goals_{f}_{n}'] = goals[:n]
            # lowest / highest formation size for the inter-goal distances of every mode that uses a formation
            for mode in ('static_diff_goal', 'swap_goals', 'dynamic_formations'):
                low, high = QUADS_PARAMS_DICT[mode][1]
                lo, hi = get_formation_range(mode=mode, formation=name, num_agents=n, low=low, high=high,
                                             num_agents_per_layer=sc.num_agents_per_layer)
                out[f'range_{mode}_{f}_{n}'] = np.array([lo, hi], dtype=np.float64)
            # get_z_value draws one uniform from numpy's global stream: seed it, record the draw and the result
            np.random.seed(1000 * f + n)
            u = np.random.RandomState(1000 * f + n).uniform(low=-0.5 * 2.0, high=0.5 * 2.0)
            z = get_z_value(num_agents=n, num_agents_per_layer=sc.num_agents_per_layer, box_size=2.0, formation=name,
                            formation_size=size)
            out[f'z_{f}_{n}'] = np.array([u, z], dtype=np.float64)
    out['formations'] = np.array(list(QUADS_FORMATION_LIST))
    out['params'] = np.array([size, layer, *center])
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'formations.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, len(out), 'arrays', os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
