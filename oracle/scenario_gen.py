"""CPU twin of the device-side episode generators (quad_swarm_rl_b200/csrc/qs_device.cuh: o_random_episode) —
TEST INFRASTRUCTURE ONLY.

o_random (reference: scenarios/obstacles/o_random.py:27-52, o_base.py:71-83, quadrotor_multi.py:304-325): M distinct
pillar cells on the L x W grid, N distinct free spawn cells, N distinct free goal cells, heights ~ U(1, 3).  The
reference samples with numpy's permutation-based `choice`; the kernels sample sequentially without replacement with
keyed Philox draws (same distribution: uniform over distinct subsets, exchangeable order).  This file restates the
kernels' procedure exactly so that the oracle and the device produce the SAME episode from the same seed.
"""
import numpy as np

from . import philox as px


def _nth_free(mask, r, cells):
    k = 0
    while k < cells:
        if not (mask >> k) & 1:
            if r == 0:
                break
            r -= 1
        k += 1
    return k


def _pick(draws, v, n):
    k = int(draws.uniform(px.SITE_SCENARIO_U, 0, 0, v) * 16777216.0)
    return (k * n) >> 24


def _cell_center(cell, L, W):
    rid, cid = divmod(cell, W)
    return np.array([cid + 0.5 - (L // 2), (W - 1 - rid) + 0.5 - (W // 2)])


def o_random_episode(draws, n_agents, M, L, W):
    """draws: philox.KeyedDraws of (seed, env, step_count).  Returns goals[N,3], spawn[N,3], obst_xy[M,2]."""
    cells = L * W
    mask = 0
    obst = np.zeros((M, 2))
    for m in range(M):
        c = _nth_free(mask, _pick(draws, m, cells - m), cells)
        mask |= 1 << c
        obst[m] = _cell_center(c, L, W)
    free = cells - M
    ms = mg = mask
    spawn = np.zeros((n_agents, 3))
    goals = np.zeros((n_agents, 3))
    for k in range(n_agents):
        cs = _nth_free(ms, _pick(draws, 64 + k, free - k), cells)
        ms |= 1 << cs
        cg = _nth_free(mg, _pick(draws, 192 + k, free - k), cells)
        mg |= 1 << cg
        spawn[k, :2] = _cell_center(cs, L, W)
        goals[k, :2] = _cell_center(cg, L, W)
        spawn[k, 2] = 1.0 + (3.0 - 1.0) * draws.uniform(px.SITE_SCENARIO_U, 0, 0, 128 + k)
        goals[k, 2] = 1.0 + (3.0 - 1.0) * draws.uniform(px.SITE_SCENARIO_U, 0, 0, 256 + k)
    return goals, spawn, obst


class DeviceORandomSource:
    """Episode source for OracleEnv that mirrors QS_SCENARIO_O_RANDOM (needs a PhiloxRng)."""

    approch_goal_metric = 0.5

    def __init__(self, L=8, W=8):
        self.L, self.W = L, W

    def name(self):
        return 'Scenario_o_random'

    def reset(self, env):
        return o_random_episode(env.rng.draws, env.num_agents, env.cfg.num_obstacles, self.L, self.W)

    def step(self, env, tick):
        return None
