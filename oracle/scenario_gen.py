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


def largest_free_square_cell(mask, L, W):
    """o_base.py:123-153 on the occupancy bit mask (bit row * W + col); quirks as in the kernels' twin function."""
    dp = [[0] * W for _ in range(L)]
    for j in range(W):
        dp[0][j] = (mask >> j) & 1
    best = cx = cy = 0
    for i in range(1, L):
        dp[i][0] = (mask >> (i * W)) & 1
        for j in range(1, W):
            if not (mask >> (i * W + j)) & 1:
                dp[i][j] = min(dp[i - 1][j], dp[i][j - 1], dp[i - 1][j - 1]) + 1
                if dp[i][j] > best:
                    best = dp[i][j]
                    cx, cy = i - (best - 1) // 2, j - (best - 1) // 2
    return cx * W + cy


O_RANDOM, O_MIX, O_STATIC_SAME_GOAL = 1, 10, 11
O_DYNAMIC_SAME_GOAL, O_SWAP_GOALS, O_EP_RAND_BEZIER = 13, 14, 15
O_NAMES = {O_RANDOM: 'o_random', O_MIX: 'mix', O_STATIC_SAME_GOAL: 'o_static_same_goal', O_DYNAMIC_SAME_GOAL: 'o_dynamic_same_goal',
           O_SWAP_GOALS: 'o_swap_goals', O_EP_RAND_BEZIER: 'o_ep_rand_bezier'}
O_BEZIER_STEPS, O_BEZIER_MAX_TRIES, O_DYN_MAX_TRIES = 600, 4096, 256


class DeviceORandomSource:
    """Episode source for OracleEnv that mirrors the device-side obstacle scenarios (needs a PhiloxRng):
    QS_SCENARIO_O_RANDOM (default), QS_SCENARIO_O_STATIC_SAME_GOAL, QS_SCENARIO_MIX over the two, or one of the ticked
    scenarios o_dynamic_same_goal / o_swap_goals / o_ep_rand_bezier (qs_scenario.cuh: o_episode_extras,
    obstacle_scenario_tick; reference: scenarios/obstacles/*.py)."""

    def __init__(self, L=8, W=8, scenario=O_RANDOM, densities=None, sizes=None):
        # densities / sizes: the choice lists of the per-episode randomisation (qs_set_obstacle_randomization), or None
        self.densities, self.sizes = densities, sizes
        self.L, self.W = L, W
        self.scenario = {v: k for k, v in O_NAMES.items()}.get(scenario, scenario)
        self.mode = O_RANDOM if self.scenario == O_MIX else self.scenario
        self.period, self.next, self.events = 0, NEVER, 0

    @property
    def approch_goal_metric(self):
        return 0.5 if self.mode == O_RANDOM else 1.0                    # o_base.py:16, o_random.py:10

    def name(self):
        return 'Scenario_' + O_NAMES[self.mode]

    def _mask(self, obst):
        mask = 0
        for xy in obst:                                    # the occupancy mask from the pillar cells
            if xy[0] > 1.0e3:
                continue
            cid = int(np.floor(xy[0] + self.L // 2))
            rid = self.W - 1 - int(np.floor(xy[1] + self.W // 2))
            mask |= 1 << (rid * self.W + cid)
        return mask

    def reset(self, env):
        d = env.rng.episode_draws
        M = env.cfg.num_obstacles
        N = env.num_agents
        if self.densities is not None:
            # int(density * area) pillars (quadrotor_multi.py:128), the float32 list entry as the kernels hold it
            M = int(float(np.float32(self.densities[_pick(d, 322, len(self.densities))])) * self.L * self.W)
            env.obst_size = 2.0 * float(np.float32(0.5 * np.float32(self.sizes[_pick(d, 323, len(self.sizes))])))
        self.num_pillars = M
        goals, spawn, obst = o_random_episode(d, N, M, self.L, self.W)
        if self.scenario == O_MIX:
            self.mode = O_RANDOM if _pick(d, 321, 2) == 0 else O_STATIC_SAME_GOAL
        self.mask = mask = self._mask(obst)
        self.period, self.next = 0, NEVER
        cells = self.L * self.W
        if self.mode in (O_STATIC_SAME_GOAL, O_DYNAMIC_SAME_GOAL, O_SWAP_GOALS):
            c = _cell_center(largest_free_square_cell(mask, self.L, self.W), self.L, self.W)
            z = 1.5 + (3.0 - 1.5) * d.uniform(px.SITE_SCENARIO_U, 0, 0, 320)
            center = np.array([c[0], c[1], z])
            goals = np.tile(center, (N, 1))
            if self.mode != O_STATIC_SAME_GOAL:
                self.period = 400 + _pk(d, STREAM_RESET, SV_PERIOD, 200)
                self.next = 1 if self.mode == O_DYNAMIC_SAME_GOAL else self.period
            if self.mode == O_SWAP_GOALS:
                fm = pick_formation(d, STREAM_RESET, O_SWAP_GOALS, N)
                goals = np.array([formation_point(fm['f'], N, shuffle_rank(d, STREAM_RESET, i, 0, N), fm['size'], center, fm['layer'],
                                                  fm['per_layer']) for i in range(N)])
        elif self.mode == O_EP_RAND_BEZIER:
            c = _cell_center(_nth_free(mask, _pk(d, STREAM_RESET, SV_CX, cells - M), cells), self.L, self.W)
            goals = np.tile(np.array([c[0], c[1], 0.75 + (3.0 - 0.75) * _u(d, STREAM_RESET, SV_CZ)]), (N, 1))
            self.period, self.next = 1, 1
            self.p0 = self.p1 = self.p2 = goals[0].copy()
        self.goals = goals
        return goals, spawn, obst

    def step(self, env, tick):
        if tick != self.next:
            return None
        d, N = env.rng.draws, env.num_agents
        g = self.goals
        nxt = tick + self.period if self.period > 0 else NEVER
        cells = self.L * self.W
        if self.mode == O_SWAP_GOALS:
            g = np.array([g[shuffle_rank(d, STREAM_TICK, i, 0, N)] for i in range(N)])
        elif self.mode == O_DYNAMIC_SAME_GOAL:
            g0 = np.array(g[0], dtype=np.float64)
            free = cells - bin(self.mask).count('1')
            for k in range(O_DYN_MAX_TRIES):
                v0 = SV_BEZIER + 8 * k
                c = _cell_center(_nth_free(self.mask, _pk(d, STREAM_TICK, v0, free), cells), self.L, self.W)
                cand = np.array([c[0], c[1], 0.75 + (3.0 - 0.75) * _u(d, STREAM_TICK, v0 + 1)])
                if np.sqrt(np.sum((g0 - cand) ** 2)) <= 4.0:
                    g = np.tile(cand, (N, 1))
                    break
            nxt = (tick // self.period + 1) * self.period
        elif self.mode == O_EP_RAND_BEZIER:
            t = tick % O_BEZIER_STEPS
            g0 = np.array(g[0], dtype=np.float64)
            if t == 0 or tick == 1:
                hx, hy, hz = 5.0, 5.0, 3.0
                p1 = p2 = g0
                lo, hi = np.array([-hx + 0.5, -hy + 0.5, 1.5 + 0.5]), np.array([hx - 0.5, hy - 0.5, hz - 0.5])
                for k in range(O_BEZIER_MAX_TRIES):
                    v0 = SV_BEZIER + 8 * k
                    u = [-h + 2.0 * h * _u(d, STREAM_TICK, v0 + j) for j, h in enumerate((hx, hy, hz, hx, hy, hz))]
                    dist = float(2 + _pk(d, STREAM_TICK, v0 + 6, 4))
                    a, b = np.array([u[0], u[2], u[4]]), np.array([u[1], u[3], u[5]])
                    q1 = g0 + a * (dist / np.sqrt(np.sum(a * a)))
                    q2 = g0 + b * (dist / np.sqrt(np.sum(b * b)))
                    if (q1 > lo).all() and (q1 < hi).all() and (q2 > lo).all() and (q2 < hi).all():
                        p1, p2 = q1, q2
                        break
                self.p0, self.p1, self.p2 = g0, p1, p2
            if t != 0 and tick > 1:
                sp = t / (O_BEZIER_STEPS - 1)
                g = np.tile((1 - sp) ** 2 * self.p0 + 2 * (1 - sp) * sp * self.p1 + sp ** 2 * self.p2, (N, 1))
        self.next = nxt
        self.events += 1
        self.goals = g
        return g


# ------------------------------------------------------------------------------------------------------------------
# Obstacle-free scenario family (twin of quad_swarm_rl_b200/csrc/qs_scenario.cuh; reference: scenarios/base.py:39-150,
# utils.py:24-175, dynamic_same_goal.py, dynamic_diff_goal.py, swap_goals.py, dynamic_formations.py,
# ep_lissajous3D.py, swarm_vs_swarm.py, mix.py:37-93).  float64; the draws and every integer decision are the kernels'.
# ------------------------------------------------------------------------------------------------------------------
STATIC_SAME_GOAL, STATIC_DIFF_GOAL, DYNAMIC_SAME_GOAL, DYNAMIC_DIFF_GOAL, SWAP_GOALS, DYNAMIC_FORMATIONS, \
    EP_LISSAJOUS3D, SWARM_VS_SWARM, MIX = range(2, 11)
EP_RAND_BEZIER = 12
RUN_AWAY = 16
SV_BEZIER, BEZIER_STEPS, BEZIER_MAX_TRIES = 64, 500, 512
MODE_NAMES = {STATIC_SAME_GOAL: 'static_same_goal', STATIC_DIFF_GOAL: 'static_diff_goal',
              DYNAMIC_SAME_GOAL: 'dynamic_same_goal', DYNAMIC_DIFF_GOAL: 'dynamic_diff_goal', SWAP_GOALS: 'swap_goals',
              DYNAMIC_FORMATIONS: 'dynamic_formations', EP_LISSAJOUS3D: 'ep_lissajous3D',
              SWARM_VS_SWARM: 'swarm_vs_swarm', MIX: 'mix', EP_RAND_BEZIER: 'ep_rand_bezier', RUN_AWAY: 'run_away'}
MODE_IDS = {v: k for k, v in MODE_NAMES.items()}
FORMATION_NAMES = ('circle_horizontal', 'circle_vertical_xz', 'circle_vertical_yz', 'sphere',
                   'grid_horizontal', 'grid_vertical_xz', 'grid_vertical_yz', 'cube')          # utils.py:24-25

SV_MIX, SV_PERIOD, SV_FORMATION, SV_SIZE, SV_LAYER, SV_CX, SV_CY, SV_CZ, SV_DIST, SV_PHI, SV_THETA, SV_GROW, SV_SPEED = range(13)
SV_RUN0, SV_RUN1 = 13, 14
SV_SHUFFLE = 16
STREAM_RESET, STREAM_TICK = 1, 2
NEVER = 0x7fffffff
BOX = 2.0
CONTROL_FREQ = 100.0


def _u24(draws, stream, v):
    return int(draws.uniform(px.SITE_SCENARIO_U, 0, stream, v) * 16777216.0)


def _u(draws, stream, v):
    return draws.uniform(px.SITE_SCENARIO_U, 0, stream, v)


def _pk(draws, stream, v, n):
    return (_u24(draws, stream, v) * n) >> 24


def grid_dims(num):
    d1 = int(np.floor(np.sqrt(num)))
    while d1 > 1 and num % d1 != 0:
        d1 -= 1
    d1 = max(d1, 1)
    return d1, num // d1


def _place(plane, a, b, l):
    return np.array([[a, b, l], [a, l, b], [l, a, b]][plane], dtype=np.float64)


def per_layer_of(f):
    return 50 if 4 <= f <= 6 else 8


def formation_raw(f, n, k, size, layer_dist, per_layer):
    if f <= 2:
        layer = k // per_layer
        m = (per_layer if layer < n // per_layer else n % per_layer) if n > per_layer else n
        ang = 2.0 * np.pi * (k % m) / m
        return _place(f, size * np.cos(ang), size * np.sin(ang), layer * layer_dist)
    if f == 3:
        nn = float(max(n, 3))
        x = 0.1 + 1.2 * nn
        start = -1.0 + 1.0 / (nn - 1.0)
        inc = (2.0 - 2.0 / (nn - 1.0)) / (nn - 1.0)
        s = start + inc * k
        lon = s * x
        lat = 0.5 * np.pi * np.sign(s) * (1.0 - np.sqrt(1.0 - abs(s)))
        return size * np.array([np.cos(lon) * np.cos(lat), np.sin(lon) * np.cos(lat), np.sin(lat)])
    if f <= 6:
        layer = k // per_layer
        nl = n if n <= per_layer else (per_layer if layer < n // per_layer else n % per_layer)
        d1, d2 = grid_dims(nl)
        return _place(f - 4, size * (k % d2), size * ((k // d2) % d1), layer * layer_dist)
    side = int(np.power(n, 1.0 / 3))                      # base.py:98-99 (27 -> 2: the cube root rounds below 3)
    return np.array([size * (k // (side * side)), size * ((k // side) % side), size * (k % side)], dtype=np.float64)


def formation_point(f, n, k, size, c, layer_dist, per_layer):
    r = formation_raw(f, n, k, size, layer_dist, per_layer)
    if f >= 4:
        r = r - np.mean([formation_raw(f, n, q, size, layer_dist, per_layer) for q in range(n)], axis=0)
    return r + np.asarray(c, dtype=np.float64)


def pick_formation(draws, stream, mode, n):
    count, low, high = 8, 0.25, 0.5
    if mode in (STATIC_SAME_GOAL, DYNAMIC_SAME_GOAL, EP_LISSAJOUS3D, EP_RAND_BEZIER):
        count, low, high = 1, 0.0, 0.0
    elif mode == SWAP_GOALS:
        low, high = 0.4, 0.8
    elif mode == O_SWAP_GOALS:
        count, low, high = 7, 0.4, 0.8
    elif mode == DYNAMIC_FORMATIONS:
        low, high = 0.0, 1.0
    f = _pk(draws, stream, SV_FORMATION, count) if count > 1 else 0
    per_layer = per_layer_of(f)
    if f <= 2:
        inv = 0.5 / np.sin(np.pi / per_layer)
        lo, hi = low * inv, high * inv
    elif f == 3:
        A, B, C, D = 1.75388487222762, 0.860487305801679, 10.3632729642351, 0.0920858134405214
        inv = 1.0 / ((A - D) / (1.0 + (n / C) ** B) + D)
        lo, hi = low * inv, high * inv
    else:
        lo, hi = low, high
    size = lo + (hi - lo) * _u(draws, stream, SV_SIZE)
    layer = lo + (hi - lo) * _u(draws, stream, SV_LAYER)
    return dict(f=f, per_layer=per_layer, lo=lo, hi=hi, size=size, layer=layer)


def z_above_ground(u, num_agents, per_layer, f, size):
    z = (-0.5 * BOX + BOX * u) + 2.0
    lower = 0.25
    if 1 <= f <= 3:
        lower = size + 0.25
    elif f in (5, 6):
        lower = grid_dims(min(num_agents, per_layer))[0] * size + 0.25
    return max(lower, z)


def shuffle_rank(draws, stream, i, g0, g1):
    ui = _u24(draws, stream, SV_SHUFFLE + i)
    r = 0
    for j in range(g0, g1):
        uj = _u24(draws, stream, SV_SHUFFLE + j)
        r += 1 if (uj < ui or (uj == ui and j < i)) else 0
    return r


class DeviceScenarioSource:
    """Episode source for OracleEnv that mirrors the device-side scenario family (needs a PhiloxRng).  `mode` is a
    QS_SCENARIO_* id (2..10) or its reference name."""

    approch_goal_metric = 0.5

    def __init__(self, mode):
        self.cfg_mode = MODE_IDS[mode] if isinstance(mode, str) else int(mode)
        self.s = None
        self.events = 0                 # goal events so far (tests)

    def name(self):
        return 'Scenario_' + MODE_NAMES[self.s['mode'] if self.s else self.cfg_mode]

    def _svs_goal(self, N, k):
        s = self.s
        h = N // 2
        if k < h:
            return formation_point(s['f'], h, k, s['size'], s['c1'], s['layer'], per_layer_of(s['f']))
        return formation_point(s['f'], N - h, k - h, s['size'], s['c2'], s['layer'], per_layer_of(s['f']))

    def reset(self, env):
        d, N = env.rng.episode_draws, env.num_agents
        mode = self.cfg_mode
        if mode == MIX:
            if N == 1:
                mode = (STATIC_SAME_GOAL, STATIC_DIFF_GOAL, EP_LISSAJOUS3D, EP_RAND_BEZIER, DYNAMIC_SAME_GOAL)[_pk(d, STREAM_RESET, SV_MIX, 5)]
            else:
                k = _pk(d, STREAM_RESET, SV_MIX, 9)
                mode = STATIC_SAME_GOAL + k if k < 8 else EP_RAND_BEZIER
        svs = mode == SWARM_VS_SWARM
        fm = pick_formation(d, STREAM_RESET, mode, N // 2 if svs else N)
        s = dict(mode=mode, period=0, next=NEVER, growing=0, speed=0.0, f=fm['f'], size=fm['size'], layer=fm['layer'],
                 hi=fm['hi'], c1=np.array([0.0, 0.0, 2.0]), c2=np.array([0.0, 0.0, 2.0]))
        self.s = s
        if mode in (DYNAMIC_SAME_GOAL, DYNAMIC_DIFF_GOAL, SWAP_GOALS, SWARM_VS_SWARM):
            s['period'] = 400 + _pk(d, STREAM_RESET, SV_PERIOD, 200)
            s['next'] = s['period']
        elif mode == RUN_AWAY:
            s['period'], s['next'] = 100, 100           # run_away.py:15-18: every second, never at tick 0
        if svs:
            c1 = np.array([-BOX + 2.0 * BOX * _u(d, STREAM_RESET, SV_CX), -BOX + 2.0 * BOX * _u(d, STREAM_RESET, SV_CY),
                           z_above_ground(_u(d, STREAM_RESET, SV_CZ), N, fm['per_layer'], fm['f'], fm['size'])])
            dist = 0.25 * BOX + (BOX - 0.25 * BOX) * _u(d, STREAM_RESET, SV_DIST)
            phi = -np.pi + 2.0 * np.pi * _u(d, STREAM_RESET, SV_PHI)
            theta = -0.5 * np.pi + np.pi * _u(d, STREAM_RESET, SV_THETA)
            c2 = c1 + dist * np.array([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)])
            plane = fm['f'] if fm['f'] <= 2 else (fm['f'] - 4 if 4 <= fm['f'] <= 6 else -1)
            if plane >= 0:
                ax = (2, 1, 0)[plane]
                diff = c2[ax] - c1[ax]
                if abs(diff) < fm['lo']:
                    c2[ax] = np.sign(diff) * fm['lo'] + c1[ax]
            s['c1'], s['c2'] = c1, c2
            goals = np.array([self._svs_goal(N, i) for i in range(N)])
        elif mode == EP_LISSAJOUS3D:
            s['c1'] = np.array([-2.0, 0.0, 2.0])
            s['period'], s['next'] = 1, 1
            goals = np.tile(s['c1'], (N, 1))
        elif mode == EP_RAND_BEZIER:
            s['period'], s['next'] = 1, 1
            s['size'], s['layer'], s['hi'] = 0.0, 0.0, 2.0          # P0 of the running segment lives in these three slots
            goals = np.tile(s['c1'], (N, 1))
        else:
            if mode == DYNAMIC_FORMATIONS:
                s['growing'] = 1 if _u24(d, STREAM_RESET, SV_GROW) < (1 << 23) else 0
                s['speed'] = 1.0 + 2.0 * _u(d, STREAM_RESET, SV_SPEED)
                s['period'], s['next'] = 1, 1
            goals = np.array([formation_point(s['f'], N, shuffle_rank(d, STREAM_RESET, i, 0, N), s['size'], s['c1'],
                                              s['layer'], fm['per_layer']) for i in range(N)])
        self.goals = goals
        return goals, None, None

    def step(self, env, tick):
        s = self.s
        if s is None or tick != s['next']:
            return None
        d, N = env.rng.draws, env.num_agents
        mode = s['mode']
        g = self.goals
        if mode == SWAP_GOALS:
            g = np.array([g[shuffle_rank(d, STREAM_TICK, i, 0, N)] for i in range(N)])
        elif mode == RUN_AWAY:
            # run_away.py:19-24: goals[0] = goals[g0]; goals[1] = goals[g1] with g0, g1 ~ randint(1, N)
            g = np.array(g, dtype=np.float64)
            g0, g1 = 1 + _pk(d, STREAM_TICK, SV_RUN0, N - 1), 1 + _pk(d, STREAM_TICK, SV_RUN1, N - 1)
            g[0] = g[g0]
            g[1] = g[g1]
        elif mode == DYNAMIC_SAME_GOAL:
            s['c1'] = np.array([-BOX + 2.0 * BOX * _u(d, STREAM_TICK, SV_CX), -BOX + 2.0 * BOX * _u(d, STREAM_TICK, SV_CY),
                                max(0.25, (-0.5 * BOX + BOX * _u(d, STREAM_TICK, SV_CZ)) + 2.0)])
            g = np.array([formation_point(s['f'], N, i, s['size'], s['c1'], 0.0, per_layer_of(s['f'])) for i in range(N)])
        elif mode == DYNAMIC_DIFF_GOAL:
            s['c1'] = np.array([-BOX + 2.0 * BOX * _u(d, STREAM_TICK, SV_CX), -BOX + 2.0 * BOX * _u(d, STREAM_TICK, SV_CY),
                                z_above_ground(_u(d, STREAM_TICK, SV_CZ), N, per_layer_of(s['f']), s['f'], s['size'])])
            fm = pick_formation(d, STREAM_TICK, mode, N)
            s.update(f=fm['f'], size=fm['size'], layer=fm['layer'], hi=fm['hi'])
            g = np.array([formation_point(s['f'], N, shuffle_rank(d, STREAM_TICK, i, 0, N), s['size'], s['c1'], s['layer'],
                                          fm['per_layer']) for i in range(N)])
        elif mode == DYNAMIC_FORMATIONS:
            if s['size'] <= -s['hi']:
                s['growing'], s['speed'] = 1, 1.0 + 2.0 * _u(d, STREAM_TICK, SV_SPEED)
            elif s['size'] >= s['hi']:
                s['growing'], s['speed'] = 0, 1.0 + 2.0 * _u(d, STREAM_TICK, SV_SPEED)
            s['size'] += (0.001 if s['growing'] else -0.001) * s['speed']
            g = np.array([formation_point(s['f'], N, i, s['size'], s['c1'], s['layer'], per_layer_of(s['f'])) for i in range(N)])
        elif mode == EP_LISSAJOUS3D:
            t = tick / CONTROL_FREQ
            g = np.tile(g[0] + np.array([0.03 * np.sin(t), 0.01 * np.sin(2 * t + 90), 0.01 * np.cos(2 * t + 90)]), (N, 1))
        elif mode == EP_RAND_BEZIER:
            t = tick % BEZIER_STEPS
            g0 = np.array(g[0], dtype=np.float64)
            if t == 0 or tick == 1:
                hx, hy, hz = 5.0, 5.0, 10.0
                p1 = p2 = g0
                for k in range(BEZIER_MAX_TRIES):
                    v0 = SV_BEZIER + 8 * k
                    u = [-h + 2.0 * h * _u(d, STREAM_TICK, v0 + j) for j, h in enumerate((hx, hy, hz, hx, hy, hz))]
                    dist = float(5 + _pk(d, STREAM_TICK, v0 + 6, 6))
                    a, b = np.array([u[0], u[2], u[4]]), np.array([u[1], u[3], u[5]])
                    p1 = g0 + a * (dist / np.sqrt(np.sum(a * a)))
                    p2 = g0 + b * (dist / np.sqrt(np.sum(b * b)))
                    lo, hi = np.array([-hx + 0.5, -hy + 0.5, 0.5]), np.array([hx - 0.5, hy - 0.5, hz - 0.5])
                    if (p1 > lo).all() and (p1 < hi).all() and (p2 > lo).all() and (p2 < hi).all():
                        break
                s['size'], s['layer'], s['hi'] = float(g0[0]), float(g0[1]), float(g0[2])
                s['c1'], s['c2'] = p1, p2
            if t != 0 and tick > 1:
                sp = t / (BEZIER_STEPS - 1)
                p0 = np.array([s['size'], s['layer'], s['hi']])
                g = np.tile((1 - sp) ** 2 * p0 + 2 * (1 - sp) * sp * s['c1'] + sp ** 2 * s['c2'], (N, 1))
        elif mode == SWARM_VS_SWARM:
            s['c1'], s['c2'] = s['c2'], s['c1']
            fm = pick_formation(d, STREAM_TICK, mode, N // 2)
            s.update(f=fm['f'], size=fm['size'], layer=fm['layer'], hi=fm['hi'])
            h = N // 2
            g = np.array([self._svs_goal(N, shuffle_rank(d, STREAM_TICK, i, 0, h) if i < h
                                         else h + shuffle_rank(d, STREAM_TICK, i, h, N)) for i in range(N)])
        s['next'] = tick + s['period'] if s['period'] > 0 else NEVER
        self.events += 1
        self.goals = g
        return g
