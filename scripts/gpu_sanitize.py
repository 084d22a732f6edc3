"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): both step kernels, resets, contacts."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from quad_swarm_rl_b200.engine import QuadSwarmEngine
from tests.parity_util import make_tables
for split in ('0', '1'):
    os.environ['QS_SPLIT'] = split
    for kw, E in ((dict(num_agents=8, neighbor_visible_num=2, obs_repr='xyz_vxyz_R_omega_floor', use_obstacles=True, use_downwash=True), 13),
                  (dict(num_agents=5, neighbor_visible_num=-1, obs_repr='xyz_vxyz_R_omega_wall'), 7),
                  (dict(num_agents=32, neighbor_visible_num=6), 3)):
        for scn in ((None, 'o_random') if kw.get('use_obstacles') else (None,)):
            eng = QuadSwarmEngine(num_envs=E, seed=1, ep_time=0.12, device_scenario=scn, **kw)
            if scn is None:
                t = make_tables(np.random.RandomState(2), E, kw['num_agents'], eng.M, kw.get('use_obstacles', False), episodes=1, spread=0.05)[0]
                t['spawn'] = t['goals'].copy() if not kw.get('use_obstacles') else t['spawn']     # tight clusters -> contacts
                eng.set_next_episode(t['goals'], t['spawn'], t['obst'])
            eng.reset()
            a = torch.rand((30, E, kw['num_agents'], 4), device='cuda') * 2 - 1
            for k in range(14):
                eng.step(a[k].contiguous(), with_terms=True)
            eng.rollout(a[14:].contiguous())
            eng.get_state(); eng.episode_stats()
            torch.cuda.synchronize()
            eng.close()
print('sanitize workload done')
