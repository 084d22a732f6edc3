"""Quick on-GPU sanity probe (not a test): create / reset / step a few configs with timing prints."""
import os
import sys
import time
import faulthandler

faulthandler.dump_traceback_later(100, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from quad_swarm_rl_b200.engine import QuadSwarmEngine

print('cpus', os.cpu_count(), 'torch', torch.__version__, torch.cuda.get_device_name(0), flush=True)
for name, kw, E in [('n1', dict(num_agents=1, neighbor_visible_num=0), 64),
                    ('n8', dict(num_agents=8, neighbor_visible_num=6), 64),
                    ('n8obst', dict(num_agents=8, neighbor_visible_num=2, use_obstacles=True, use_downwash=True,
                                    obs_repr='xyz_vxyz_R_omega_floor'), 4096),
                    ('n5all', dict(num_agents=5, neighbor_visible_num=-1, obs_repr='xyz_vxyz_R_omega_wall'), 33),
                    ('n32', dict(num_agents=32, neighbor_visible_num=6), 128)]:
    t0 = time.time()
    eng = QuadSwarmEngine(num_envs=E, ep_time=0.2, **kw)
    if eng.M:
        rs = np.random.RandomState(0)
        ob = rs.uniform(-4, 4, (E, eng.M, 2)).astype(np.float32)
        g = rs.uniform(-3, 3, (E, eng.N, 3)).astype(np.float32) + np.array([0, 0, 4], np.float32)
        eng.set_next_episode(g, g, ob)
    obs = eng.reset()
    torch.cuda.synchronize()
    print(name, 'reset ok', time.time() - t0, float(obs.abs().max()), flush=True)
    a = torch.rand((E, eng.N, 4), device='cuda') * 2 - 1
    for t in range(50):
        o, r, d = eng.step(a)
    torch.cuda.synchronize()
    print(name, 'step ok', time.time() - t0, 'obs finite', bool(torch.isfinite(o).all()), 'rew mean', float(r.mean()),
          'dones', int(d.sum()), flush=True)
    t1 = time.time()
    for t in range(1000):
        eng.step(a)
    torch.cuda.synchronize()
    dt = time.time() - t1
    print(name, f'1000 steps {dt*1e3:.1f} ms -> {E*eng.N*1000/dt/1e6:.1f} M agent-steps/s (python launch loop)', flush=True)
    eng.close()
print('sanity done', flush=True)
