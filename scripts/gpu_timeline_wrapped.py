"""Timeline of WRAPPED control steps (debug build, -DQS_TIMELINE): step-kernel stamps as scripts/gpu_timeline.py plus four
stamps of the wrapper kernel's block (entry, `done` taken, body finished, block handed on).
    QS_LIB=$PWD/tune/libquadswarm_tl.so python scripts/gpu_timeline_wrapped.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from quad_swarm_rl_b200 import _lib as L
from quad_swarm_rl_b200.engine import QuadSwarmEngine

cfg = bench.CONFIGS['c3']
E, kw = cfg['E'], cfg['kw']
N = kw['num_agents']
eng = QuadSwarmEngine(num_envs=E, seed=0, rew_coeff=cfg['rew'], device_scenario=cfg['mode'], **kw)
eng.reset()
st = eng.get_state()
st['env_i32'][:, 0] = torch.randint(0, eng.ep_len + 1, (E,), device='cuda', dtype=torch.int32)
eng.set_state(st)
eng.wrap_enable(use_replay=os.environ.get('QS_WRAP_REPLAY', '1') != '0', replay_buffer_size=20, replay_prob=0.75, replay_always_active=True)
eng.set_chained(True)
K = 20
act = (torch.rand((K, E, N, 4), device='cuda') * 2 - 1).contiguous()
obs = torch.empty((K, E, N, eng.D), device='cuda'); rew = torch.empty((K, E, N), device='cuda')
dn = torch.empty((K, E, N), dtype=torch.uint8, device='cuda')
s = torch.cuda.Stream()
lib = L.load()
lib.qs_debug_timeline.argtypes = [C.c_void_p, C.c_void_p]
lib.qs_debug_timeline_rewind.argtypes = [C.c_void_p]
with torch.cuda.stream(s):
    for t in range(3):
        eng.wrap_step(act[t], obs_out=obs[t], rewards_out=rew[t], dones_out=dn[t])
    s.synchronize()
    lib.qs_debug_timeline_rewind(eng.h)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for t in range(K):
            eng.wrap_step(act[t], obs_out=obs[t], rewards_out=rew[t], dones_out=dn[t])
    for _ in range(30):
        g.replay()
    s.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(50):
        g.replay()
    e1.record(s)
    s.synchronize()
print(f'c3 wrapped: {e0.elapsed_time(e1) / 50 / K * 1e3:.2f} us per control step (instrumented build)')
buf = np.zeros((64, 4096, 16), np.uint64)
L.check(lib.qs_debug_timeline(eng.h, buf.ctypes.data_as(C.c_void_p)))
tl = buf[:K].astype(np.int64)
nb = int((tl[0, :, 0] > 0).sum())
t0 = tl[:, :nb, 0].min()
T = (tl[:, :nb, :] - t0) / 1e3
names = ['s.entry', 's.waited', 's.loaded', 's.dyn', 's.pairs', 's.obs', 's.emit', 's.exit', None, None, None, None, 'w.entry', 'w.taken', 'w.body', 'w.exit']
cols = [0, 1, 2, 3, 4, 5, 6, 7, 12, 13, 14, 15]
print(f'blocks {nb}; median over blocks of (stamp - first step entry of the control step)')
for k in range(2, K):
    ent = T[k, :, 0].min()
    print(f'step {k:2d}: first entry {ent:8.2f} (+{ent - T[k - 1, :, 0].min():5.2f}) | ' + ' '.join(f'{names[j]}={np.median(T[k, :, j]) - ent:6.2f}' for j in cols[1:]))
d = T[2:]
print('median per-block durations (us): ' + ' '.join(f'{names[cols[j]]}-{names[cols[j - 1]]}={np.median(d[:, :, cols[j]] - d[:, :, cols[j - 1]]):.2f}' for j in range(1, len(cols))))
print('p95: ' + ' '.join(f'{np.percentile(d[:, :, cols[j]] - d[:, :, cols[j - 1]], 95):.2f}' for j in range(1, len(cols))))
print('same block, next control step: s.waited(t+1) - w.exit(t), median / p95:', np.round(np.median(d[1:, :, 1] - d[:-1, :, 15]), 2), np.round(np.percentile(d[1:, :, 1] - d[:-1, :, 15], 95), 2))
print('per-block period s.waited(t+1) - s.waited(t): median', np.round(np.median(d[1:, :, 1] - d[:-1, :, 1]), 2), ' p95', np.round(np.percentile(d[1:, :, 1] - d[:-1, :, 1], 95), 2))
eng.close()
