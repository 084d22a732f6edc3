"""Dev probe: decompose the per-step time (graph replay vs rollout-in-one-launch vs tiny-E launch floor)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from quad_swarm_rl_b200.engine import QuadSwarmEngine
import bench

def timed(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n

for cfgname, E in (('c3', 4096), ('c3', 32), ('c2', 1024), ('c2', 4096), ('c4', 2048), ('c3', 16384)):
    cfg = bench.CONFIGS[cfgname]
    eng = QuadSwarmEngine(num_envs=E, seed=0, rew_coeff=cfg['rew'], **cfg['kw'])
    g, s, o = bench.make_episode_tables(cfg, E, 1)
    eng.set_next_episode(g, s, o); eng.reset()
    N = eng.N
    T = 64
    acts = (torch.rand((T, E, N, 4), device='cuda') * 2 - 1).contiguous()
    obs = torch.empty((T, E, N, eng.D), device='cuda'); rew = torch.empty((T, E, N), device='cuda'); dn = torch.empty((T, E, N), dtype=torch.uint8, device='cuda')
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for t in range(3): eng.step(acts[t], obs_out=obs[t], rewards_out=rew[t], dones_out=dn[t])
        st.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st):
            for t in range(T): eng.step(acts[t], obs_out=obs[t], rewards_out=rew[t], dones_out=dn[t])
        gr.replay(); st.synchronize()
        us_graph = timed(lambda: [gr.replay() for _ in range(20)], 20 * T)
        eng.rollout(acts, obs_out=obs, rewards_out=rew, dones_out=dn)
        us_roll = timed(lambda: [eng.rollout(acts, obs_out=obs, rewards_out=rew, dones_out=dn) for _ in range(5)], 5 * T)
        o1 = obs[:1]
        us_roll_last = timed(lambda: [eng.rollout(acts, obs_out=o1, rewards_out=rew, dones_out=dn, last_obs_only=True) for _ in range(5)], 5 * T)
    A = E * N
    print(f'{cfgname} E={E} A={A}: graph {us_graph:.2f} us/step ({A/us_graph/1e3:.2f} G/s) | rollout {us_roll:.2f} us/step ({A/us_roll/1e3:.2f} G/s) | rollout last-obs {us_roll_last:.2f} us/step', flush=True)
    eng.close()
