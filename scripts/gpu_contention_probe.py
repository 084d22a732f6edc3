"""Probe: does a chained graph of step grids still equal one rollout when several processes share the GPU by time-slicing?
(The xdist run of the GPU suite on 6 workers saw one mismatch in test_back_to_back_step_grids_equal_one_rollout[c3_small];
serially the test passes.)  Usage: gpu_contention_probe.py <tag> <seconds> [E] [background]
Environment (QS_PDL, QS_SPLIT) selects the launch shape, as in the test."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from tests.test_gpu_api import C3, _actions, _engine

tag, secs = sys.argv[1], float(sys.argv[2])
E = int(sys.argv[3]) if len(sys.argv) > 3 else 37
background = len(sys.argv) > 4 and sys.argv[4] == 'bg'
T, N = 96, 8
t_end = time.time() + secs
if background:          # a heavy neighbour: 4096 envs stepping without pause
    e, _ = _engine(4096, C3, ep_time=1.0)
    a = _actions(8, 4096, N)
    e.reset()
    n = 0
    while time.time() < t_end:
        for t in range(8):
            e.step(a[t])
        torch.cuda.synchronize(); n += 8
    print(f'[{tag}] background steps {n}', flush=True)
    sys.exit(0)
trials = bad = tmo = 0
first = []
while time.time() < t_end:
    e1, _ = _engine(E, C3, ep_time=1.0); e2, _ = _engine(E, C3, ep_time=1.0)
    e1.set_chained(True); e2.set_chained(True)
    a = _actions(T, E, N)
    st = torch.cuda.Stream()
    if '--no-wait-stream' not in sys.argv:
        st.wait_stream(torch.cuda.current_stream())     # tables / actions were enqueued on the default stream
    obs = torch.empty((T, E, N, e1.D), device='cuda'); rew = torch.empty((T, E, N), device='cuda')
    dn = torch.empty((T, E, N), dtype=torch.uint8, device='cuda')
    with torch.cuda.stream(st):
        e1.reset()
        for t in range(3):
            e1.step(a[t], obs_out=obs[t], rewards_out=rew[t], dones_out=dn[t])
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for t in range(T):
                e1.step(a[t], obs_out=obs[t], rewards_out=rew[t], dones_out=dn[t])
    e2.reset()
    for t in range(3):
        e2.step(a[t])
    for r in range(3):
        g.replay(); torch.cuda.synchronize()
        o2, r2, d2 = e2.rollout(a); torch.cuda.synchronize()
        trials += 1
        if not (torch.equal(obs, o2) and torch.equal(rew, r2) and torch.equal(dn, d2)):
            bad += 1
            neq = (obs != o2).flatten(2).any(dim=2)              # [T, E]
            ts = torch.nonzero(neq.any(dim=1)).flatten()
            t0 = int(ts[0]) if len(ts) else -1
            envs = torch.nonzero(neq[t0]).flatten().tolist() if t0 >= 0 else []
            if len(first) < 4:
                first.append((r, t0, envs[:8], float((obs[t0] - o2[t0]).abs().max()) if t0 >= 0 else None))
            break
    tmo += e1.handover_timeouts + e2.handover_timeouts
    try:
        e1.close(); e2.close()
    except Exception as ex:          # a latched hand-over error makes later calls fail
        print(f'[{tag}] close: {ex}', flush=True)
print(f'[{tag}] replays {trials} mismatches {bad} handover_timeouts {tmo} first (replay, step, envs, max|d|): {first}', flush=True)
