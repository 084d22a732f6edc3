"""Dev probe: per-launch floor of dependent kernel nodes in a CUDA graph on this GPU (tiny kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from quad_swarm_rl_b200.engine import QuadSwarmEngine
eng = QuadSwarmEngine(num_envs=1, num_agents=8, neighbor_visible_num=6)
g = torch.zeros((1, 8, 3), device='cuda')
x = torch.zeros(32, device='cuda')
a = torch.zeros((1, 8, 4), device='cuda')
eng.reset()
def timed(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n
for name, body in (('tiny <<<>>> kernel (qs_set_goals, 8 threads)', lambda: eng.set_goals(g)),
                   ('torch add_ on 32 floats', lambda: x.add_(1.0)),
                   ('qs_step, 1 env x 8 drones', lambda: eng.step(a))):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3): body()
        st.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st):
            for _ in range(256): body()
        gr.replay(); st.synchronize()
        us = timed(lambda: [gr.replay() for _ in range(20)], 20 * 256)
    print(f'{name}: {us:.2f} us per dependent launch in a graph', flush=True)
