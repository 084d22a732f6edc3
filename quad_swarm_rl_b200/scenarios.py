"""Host-side episode generators (goal formations, spawn points, timed goal switches).

These sit NEXT TO the hot path: they hand `goals[N,3]` / `spawn[N,3]` to the
device at reset and on the (rare) ticks where goals move.  The semantics follow the reference's
`gym_art/quadrotor_multi/scenarios/*` so that episode statistics keep their meaning
(`Scenario_<mode>` names feed the `episode_extra_stats` keys, reward_shaping.py:95-98), but the
implementation is table-driven and draws from an explicit `numpy.random.RandomState` instead of the
global generator.  Drawing from `RandomState(seed)` in the reference's call order reproduces the
reference's goal sequences bit-for-bit, which `tests/test_oracle_vs_reference.py` relies on.

Reference behaviour cited per function (paths under gym_art/quadrotor_multi/scenarios/).
"""
import math

import numpy as np

QUAD_ARM_NOMINAL = 0.05                       # utils.py:31
FORMATIONS = ('circle_horizontal', 'circle_vertical_xz', 'circle_vertical_yz', 'sphere',
              'grid_horizontal', 'grid_vertical_xz', 'grid_vertical_yz', 'cube')      # utils.py:24-25
FORMATIONS_OBST = FORMATIONS[1:]                                                        # utils.py:27-28

# mode -> (number of formations to pick from, (low, high) inter-goal distance), utils.py:32-51
_SAME = (1, (0.0, 0.0))
MODE_TABLE = {
    'static_same_goal': _SAME, 'dynamic_same_goal': _SAME, 'ep_lissajous3D': _SAME, 'ep_rand_bezier': _SAME,
    'static_diff_goal': (8, (5 * QUAD_ARM_NOMINAL, 10 * QUAD_ARM_NOMINAL)),
    'dynamic_diff_goal': (8, (5 * QUAD_ARM_NOMINAL, 10 * QUAD_ARM_NOMINAL)),
    'swarm_vs_swarm': (8, (5 * QUAD_ARM_NOMINAL, 10 * QUAD_ARM_NOMINAL)),
    'swap_goals': (8, (8 * QUAD_ARM_NOMINAL, 16 * QUAD_ARM_NOMINAL)),
    'dynamic_formations': (8, (0.0, 20 * QUAD_ARM_NOMINAL)),
    'run_away': (8, (5 * QUAD_ARM_NOMINAL, 10 * QUAD_ARM_NOMINAL)),
    'o_random': _SAME, 'o_static_same_goal': _SAME, 'o_dynamic_same_goal': _SAME,
    'o_swap_goals': (7, (8 * QUAD_ARM_NOMINAL, 16 * QUAD_ARM_NOMINAL)), 'o_ep_rand_bezier': _SAME,
}

MIX_MODES = ('static_same_goal', 'static_diff_goal', 'ep_lissajous3D', 'ep_rand_bezier', 'dynamic_same_goal',
             'dynamic_diff_goal', 'dynamic_formations', 'swap_goals', 'swarm_vs_swarm')   # utils.py:7-10
MIX_MODES_SINGLE = ('static_same_goal', 'static_diff_goal', 'ep_lissajous3D', 'ep_rand_bezier', 'dynamic_same_goal')
MIX_MODES_OBST = ('o_random', 'o_static_same_goal')                                       # utils.py:17
MIX_MODES_OBST_SINGLE = ('o_random',)


# ------------------------------------------------------------------------------------------
# formation geometry
# ------------------------------------------------------------------------------------------
def grid_dims(num):
    """Largest divisor of num not above sqrt(num), and its cofactor (utils.py:109-121)."""
    d1 = int(math.floor(math.sqrt(num)))
    while d1 > 1 and num % d1 != 0:
        d1 -= 1
    d1 = max(d1, 1)
    return d1, num // d1


def sphere_points(n):
    """Spiral points on the unit sphere (utils.py:74-90); at least 3 points are generated."""
    n = max(int(n), 3)
    x = 0.1 + 1.2 * n
    start = -1. + 1. / (n - 1.)
    inc = (2. - 2. / (n - 1.)) / (n - 1.)
    s = start + inc * np.arange(n)
    lon = s * x
    lat = np.pi / 2. * np.sign(s) * (1. - np.sqrt(1. - np.abs(s)))
    return np.stack([np.cos(lon) * np.cos(lat), np.sin(lon) * np.cos(lat), np.sin(lat)], axis=1)


def circle_radius(num, dist):
    return (0.5 * dist) / math.sin((2 * math.pi / num) / 2)              # utils.py:102-106


def sphere_radius(num, dist):
    A, B, C, D = 1.75388487222762, 0.860487305801679, 10.3632729642351, 0.0920858134405214
    return dist / ((A - D) / (1 + (num / C) ** B) + D)                   # utils.py:92-99


def _place(formation, a, b, layer):
    """Embed planar coordinates (a, b) + layer offset according to the formation's plane (utils.py:149-160)."""
    if formation.endswith('horizontal'):
        return np.array([a, b, layer])
    if formation.endswith('vertical_xz'):
        return np.array([a, layer, b])
    if formation.endswith('vertical_yz'):
        return np.array([layer, a, b])
    raise NotImplementedError("Unknown formation")


def formation_goals(formation, num_agents, size, center, layer_dist, per_layer):
    """Goal points of a formation (base.py:39-113)."""
    center = np.asarray(center, dtype=np.float64)
    n = num_agents
    if formation.startswith('circle'):
        counts = [per_layer] * (n // per_layer) + ([n % per_layer] if n % per_layer else []) if n > per_layer else [n]
        pts = []
        for i in range(n):
            m = counts[i // per_layer]
            ang = 2 * np.pi * (i % m) / m
            pts.append(_place(formation, size * np.cos(ang), size * np.sin(ang), (i // per_layer) * layer_dist))
        return np.array(pts) + center
    if formation == 'sphere':
        return size * sphere_points(n) + center
    if formation.startswith('grid'):
        if n <= per_layer:
            dims = [grid_dims(n)]
        else:
            dims = [grid_dims(per_layer)] * (n // per_layer)
            if n % per_layer:
                dims.append(grid_dims(n % per_layer))
        pts = []
        for i in range(n):
            d1, d2 = dims[i // per_layer]
            pts.append(_place(formation, size * (i % d2), size * (int(i / d2) % d1), (i // per_layer) * layer_dist))
        pts = np.array(pts)
        return pts - np.mean(pts, axis=0) + center
    if formation.startswith('cube'):
        side = int(np.power(n, 1.0 / 3))
        pts = np.array([[center[2] + size * (i // np.square(side)), size * (int(i / side) % side), size * (i % side)]
                        for i in range(n)])
        return pts - np.mean(pts, axis=0) + center
    raise NotImplementedError("Unknown formation")


def z_above_ground(rng, num_agents, per_layer, box, formation, size):
    """Random formation-centre height that keeps every goal above the floor (utils.py:163-175)."""
    z = rng.uniform(low=-0.5 * box, high=0.5 * box) + 2.0
    lower = 0.25
    if formation == 'sphere' or formation.startswith('circle_vertical'):
        lower = size + 0.25
    elif formation.startswith('grid_vertical'):
        d1, _ = grid_dims(min(num_agents, per_layer))
        lower = d1 * size + 0.25
    return max(lower, z)


def grid_cell_centers(area_length, area_width, grid_size=1.0):
    """Centres of the pillar grid cells, column-major from the top-left (obstacles/utils.py:47-58)."""
    L_, W_ = int(area_length / grid_size), int(area_width / grid_size)
    xs = np.arange(0, area_length, grid_size) + grid_size / 2 - area_length // 2
    ys = np.arange(area_width - grid_size, -grid_size, -grid_size) + grid_size / 2 - area_width // 2
    out = np.zeros((L_ * W_, 2))
    out[:, 0] = np.repeat(xs, len(ys))
    out[:, 1] = np.tile(ys, len(xs))
    return out


def obstacle_map_given_density(rng, obst_spawn_area, obst_density, room_height=10.0, grid_size=1.0):
    """Random pillar placement on the grid without replacement (quadrotor_multi.py:304-325).
    Returns (obst_map[L,W] of 0/1, pillar positions [M,3] with z = room_height / 2, cell centres)."""
    L_, W_ = int(obst_spawn_area[0]), int(obst_spawn_area[1])
    n_cells = L_ * W_
    cells = grid_cell_centers(L_, W_, grid_size)
    picks = rng.choice(a=list(range(n_cells)), size=int(n_cells * obst_density), replace=False)
    obst_map = np.zeros([L_, W_])
    pos = []
    for k in picks:
        rid, cid = k // W_, k - (k // W_) * W_
        obst_map[rid, cid] = 1
        c = cells[rid + int(L_ / grid_size) * cid]
        pos.append([c[0], c[1], room_height / 2.])
    return obst_map, pos, cells


# ------------------------------------------------------------------------------------------
# scenario objects
# ------------------------------------------------------------------------------------------
class Scenario:
    """Common state + the formation bookkeeping of base.py:8-150."""
    mode = None
    dynamic = False          # True when step() can move goals (the env then calls it every tick)

    def __init__(self, num_agents, room_dims=(10., 10., 10.), rng=None, control_freq=100.0, ep_time=15.0, box=2.0,
                 use_obstacles=False):
        self.num_agents = num_agents
        self.room_dims = room_dims
        self.rng = rng if rng is not None else np.random.RandomState()
        self.control_freq = control_freq
        self.ep_time = ep_time
        self.box = box
        self.use_obstacles = use_obstacles
        self.goals = None
        self.spawn_points = None
        self.formation = None
        self.formation_center = None
        self.formation_size = 1.0
        self.lowest_formation_size, self.highest_formation_size = 1.0, 2.0
        self.num_agents_per_layer = 8
        self.layer_dist = self.lowest_formation_size
        self.approch_goal_metric = 0.5

    def name(self):
        return 'Scenario_' + self.mode

    # base.py:123-136 + utils.py:54-65,124-146
    def pick_formation(self):
        count, (low, high) = MODE_TABLE[self.mode]
        # the reference indexes the GLOBAL formation list with an index drawn over the mode's list (Appendix D-14)
        self.formation = FORMATIONS[self.rng.randint(low=0, high=count)]
        self.num_agents_per_layer = 50 if self.formation.startswith('grid') else 8
        n = self.num_agents // 2 if self.mode == 'swarm_vs_swarm' else self.num_agents
        if self.formation.startswith('circle'):
            lo, hi = circle_radius(self.num_agents_per_layer, low), circle_radius(self.num_agents_per_layer, high)
        elif self.formation.startswith('sphere'):
            lo, hi = sphere_radius(n, low), sphere_radius(n, high)
        else:
            lo, hi = low, high
        self.lowest_formation_size, self.highest_formation_size = lo, hi
        self.formation_size = self.rng.uniform(low=lo, high=hi)
        self.layer_dist = self.rng.uniform(low=lo, high=hi)

    def make_goals(self, num_agents=None, center=None, layer_dist=None):
        return formation_goals(self.formation, self.num_agents if num_agents is None else num_agents,
                               self.formation_size, self.formation_center if center is None else center,
                               self.layer_dist if layer_dist is None else layer_dist, self.num_agents_per_layer)

    def standard_reset(self, formation_center=None):
        self.pick_formation()
        self.formation_center = np.array([0.0, 0.0, 2.0]) if formation_center is None else formation_center
        self.goals = self.make_goals()
        self.rng.shuffle(self.goals)

    def reset(self, obst_map=None, cell_centers=None):
        self.standard_reset()

    def step(self, tick):
        return

    def _draw_period(self):
        self.period = int(self.rng.uniform(low=4.0, high=6.0) * self.control_freq)


class StaticSameGoal(Scenario):
    mode = 'static_same_goal'


class StaticDiffGoal(Scenario):
    mode = 'static_diff_goal'


class DynamicSameGoal(Scenario):
    """dynamic_same_goal.py: the common goal teleports every 4-6 s."""
    mode = 'dynamic_same_goal'
    dynamic = True

    def reset(self, obst_map=None, cell_centers=None):
        self._draw_period()
        self.standard_reset()

    def step(self, tick):
        if tick % self.period == 0 and tick > 0:
            x, y = self.rng.uniform(low=-self.box, high=self.box, size=(2,))
            z = max(0.25, self.rng.uniform(low=-0.5 * self.box, high=0.5 * self.box) + 2.0)
            self.formation_center = np.array([x, y, z])
            self.goals = self.make_goals(layer_dist=0.0)


class DynamicDiffGoal(Scenario):
    """dynamic_diff_goal.py: a new formation at a new centre every 4-6 s."""
    mode = 'dynamic_diff_goal'
    dynamic = True

    def reset(self, obst_map=None, cell_centers=None):
        self._draw_period()
        self.standard_reset()

    def step(self, tick):
        if tick % self.period == 0 and tick > 0:
            x, y = self.rng.uniform(low=-self.box, high=self.box, size=(2,))
            z = z_above_ground(self.rng, self.num_agents, self.num_agents_per_layer, self.box, self.formation,
                               self.formation_size)
            self.formation_center = np.array([x, y, z])
            self.pick_formation()
            self.goals = self.make_goals()
            self.rng.shuffle(self.goals)


class SwapGoals(Scenario):
    """swap_goals.py: goals are permuted among the drones every 4-6 s."""
    mode = 'swap_goals'
    dynamic = True

    def reset(self, obst_map=None, cell_centers=None):
        self._draw_period()
        self.standard_reset()

    def step(self, tick):
        if tick % self.period == 0 and tick > 0:
            self.rng.shuffle(self.goals)


class DynamicFormations(Scenario):
    """dynamic_formations.py: the formation breathes, goals move every tick."""
    mode = 'dynamic_formations'
    dynamic = True

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.growing = True
        self.speed = self.rng.uniform(low=1.0, high=3.0)

    def reset(self, obst_map=None, cell_centers=None):
        self.growing = bool(self.rng.uniform(low=0.0, high=1.0) < 0.5)
        self.speed = self.rng.uniform(low=1.0, high=3.0)
        self.standard_reset()

    def step(self, tick):
        if self.formation_size <= -self.highest_formation_size:
            self.growing = True
            self.speed = self.rng.uniform(low=1.0, high=3.0)
        elif self.formation_size >= self.highest_formation_size:
            self.growing = False
            self.speed = self.rng.uniform(low=1.0, high=3.0)
        self.formation_size += (0.001 if self.growing else -0.001) * self.speed
        self.goals = self.make_goals()


class Lissajous3D(Scenario):
    """ep_lissajous3D.py: all drones chase one goal that random-walks along Lissajous increments."""
    mode = 'ep_lissajous3D'
    dynamic = True

    def reset(self, obst_map=None, cell_centers=None):
        self.pick_formation()
        self.formation_center = np.array([-2.0, 0.0, 2.0])
        self.goals = self.make_goals(layer_dist=0.0)

    def step(self, tick):
        t = tick / self.control_freq
        d = np.array([0.03 * np.sin(t), 0.01 * np.sin(2 * t + 90), 0.01 * np.cos(2 * t + 90)])
        self.goals = np.array([self.goals[0] + d for _ in range(self.num_agents)])


def bezier_points(nodes, s):
    """Points of the Bezier curve with control points nodes[:, k] at parameters s (Bernstein form); the reference gets them
    from the third-party `bezier` package (ep_rand_bezier.py:40-42)."""
    from math import comb
    nodes = np.asarray(nodes, dtype=np.float64)
    n = nodes.shape[1] - 1
    out = np.zeros((nodes.shape[0], len(s)))
    for k in range(n + 1):
        out += np.outer(nodes[:, k], comb(n, k) * (1.0 - s) ** (n - k) * s ** k)
    return out


class RandBezier(Scenario):
    """ep_rand_bezier.py: all drones chase one goal that follows quadratic Bezier segments: every `num_secs` seconds (and at
    tick 1) two control points are drawn 5..10 m away in random directions, re-drawn until both lie inside the room."""
    mode = 'ep_rand_bezier'
    dynamic = True
    num_secs = 5
    z_low, z_high, cap = 0.0, None, 30

    def _bounds(self):
        room = np.array(self.room_dims) - self.formation_size
        zh = room[2] if self.z_high is None else self.z_high
        return room, np.array([-room[0] / 2, -room[1] / 2, self.z_low]), np.array([room[0] / 2, room[1] / 2, zh])

    def step(self, tick):
        control_steps = int(self.num_secs * self.control_freq)
        t = tick % control_steps
        room, low, high = self._bounds()
        max_dist = min(self.cap, max(room))
        min_dist = max_dist / 2
        if t == 0 or tick == 1:
            while True:
                new_pos = self.rng.uniform(low=-high, high=high, size=(2, 3)).reshape(3, 2)
                new_pos = new_pos * self.rng.randint(min_dist, max_dist + 1) / np.linalg.norm(new_pos, axis=0)
                new_pos = self.goals[0].reshape(3, 1) + new_pos
                if (new_pos > low[:, None] + 0.5).all() and (new_pos < high[:, None] - 0.5).all():
                    break
            nodes = np.concatenate((self.goals[0].reshape(3, 1), new_pos), axis=1)
            self.interp = bezier_points(nodes, np.linspace(0, 1, control_steps))
        if t != 0 and tick > 1:
            self.goals = np.array([self.interp[:, t] for _ in range(self.num_agents)])


class RunAway(Scenario):
    """run_away.py: every second the goals of drones 0 and 1 jump onto the goals of two random other drones."""
    mode = 'run_away'
    dynamic = True

    def step(self, tick):
        if tick % int(1.0 * self.control_freq) == 0 and tick > 0:
            g = self.rng.randint(low=1, high=self.num_agents, size=2)      # needs >= 2 drones, as the reference
            self.goals = self.goals.copy()
            self.goals[0] = self.goals[g[0]]
            self.goals[1] = self.goals[g[1]]


class SwarmVsSwarm(Scenario):
    """swarm_vs_swarm.py: two half-swarms whose formation centres swap every 4-6 s."""
    mode = 'swarm_vs_swarm'
    dynamic = True

    def _centers(self):
        box = self.box
        x, y = self.rng.uniform(low=-box, high=box, size=(2,))
        z = z_above_ground(self.rng, self.num_agents, self.num_agents_per_layer, box, self.formation,
                           self.formation_size)
        c1 = np.array([x, y, z])
        dist = self.rng.uniform(low=box / 4, high=box)
        phi = self.rng.uniform(low=-np.pi, high=np.pi)
        theta = self.rng.uniform(low=-0.5 * np.pi, high=0.5 * np.pi)
        c2 = c1 + dist * np.array([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)])
        axis = {'horizontal': 2, 'vertical_xz': 1, 'vertical_yz': 0}
        for suffix, ax in axis.items():
            if self.formation.endswith(suffix):
                diff = c2[ax] - c1[ax]
                if abs(diff) < self.lowest_formation_size:
                    c2[ax] = np.sign(diff) * self.lowest_formation_size + c1[ax]
        return c1, c2

    def _formations(self):
        h = self.num_agents // 2
        self.goals_1 = self.make_goals(num_agents=h, center=self.c1)
        self.goals_2 = self.make_goals(num_agents=self.num_agents - h, center=self.c2)
        self.goals = np.concatenate([self.goals_1, self.goals_2])

    def reset(self, obst_map=None, cell_centers=None):
        self._draw_period()
        self.pick_formation()
        self.c1, self.c2 = self._centers()
        self._formations()
        self.formation_center = (self.c1 + self.c2) / 2

    def step(self, tick):
        if tick % self.period == 0 and tick > 0:
            self.c1, self.c2 = self.c2.copy(), self.c1.copy()
            self.pick_formation()
            self._formations()
            self.rng.shuffle(self.goals_1)
            self.rng.shuffle(self.goals_2)
            self.goals = np.concatenate([self.goals_1, self.goals_2])


class _ObstacleScenario(Scenario):
    """Shared pieces of obstacles/o_base.py."""

    def _free_cells(self, obst_map, cell_centers):
        self.obstacle_map = obst_map
        self.cell_centers = cell_centers
        if obst_map is None or cell_centers is None:
            raise NotImplementedError
        self.free_space = list(zip(*np.where(obst_map == 0)))

    def _sample_free_points(self, n, z_low=1.0, z_high=3.0):
        """o_base.py:71-83: n distinct free cells, z uniform."""
        ids = self.rng.choice(range(len(self.free_space)), n, replace=False)
        width = self.obstacle_map.shape[0]
        pts = []
        for idx in ids:
            x, y = self.free_space[idx][0], self.free_space[idx][1]
            px_, py_ = self.cell_centers[x + width * y]
            pts.append(np.array([px_, py_, self.rng.uniform(low=z_low, high=z_high)]))
        return np.array(pts)

    def _sample_free_point(self):
        """o_base.py:54-69 (check_surroundings=False)."""
        idx = self.rng.choice(a=len(self.free_space), replace=False)
        x, y = self.free_space[idx][0], self.free_space[idx][1]
        px_, py_ = self.cell_centers[x + self.obstacle_map.shape[0] * y]
        return np.array([px_, py_, self.rng.uniform(low=0.75, high=3.0)])


class ORandom(_ObstacleScenario):
    """obstacles/o_random.py: every drone flies from its own free cell to its own free cell."""
    mode = 'o_random'

    def reset(self, obst_map=None, cell_centers=None):
        self._free_cells(obst_map, cell_centers)
        for _ in range(self.num_agents):              # the reference draws these and discards them (o_random.py:38-40)
            self._sample_free_point()
            self._sample_free_point()
        start = self._sample_free_points(self.num_agents)
        end = self._sample_free_points(self.num_agents)
        self.duration_step = int(self.rng.uniform(low=2.0, high=4.0) * self.control_freq)
        self.pick_formation()
        self.formation_center = np.array((0., 0., 2.))
        self.spawn_points = start.copy()
        self.goals = end.copy()
        self.approch_goal_metric = 0.5


class OStaticSameGoal(_ObstacleScenario):
    """obstacles/o_static_same_goal.py: spawn on free cells, common goal at the centre of the largest free square."""
    mode = 'o_static_same_goal'

    def _largest_free_square_center(self):
        """o_base.py:123-153."""
        m_ = self.obstacle_map
        n, m = m_.shape
        dp = np.zeros((n, m), dtype=int)
        dp[0] = m_[0]
        dp[:, 0] = m_[:, 0]
        best, cx, cy = 0, 0, 0
        for i in range(1, n):
            for j in range(1, m):
                if m_[i][j] == 0:
                    dp[i][j] = min(dp[i - 1][j], dp[i][j - 1], dp[i - 1][j - 1]) + 1
                    if dp[i][j] > best:
                        best = dp[i][j]
                        cx = i - (best - 1) // 2
                        cy = j - (best - 1) // 2
        px_, py_ = self.cell_centers[cx + m * cy]
        return np.array([px_, py_, self.rng.uniform(low=1.5, high=3.0)])

    def reset(self, obst_map=None, cell_centers=None):
        self.duration_time = self.rng.uniform(low=4.0, high=6.0)
        self._free_cells(obst_map, cell_centers)
        start = self._sample_free_points(self.num_agents)
        end = self._largest_free_square_center()
        self.pick_formation()
        self.spawn_points = start.copy()
        self.goals = np.array([end for _ in range(self.num_agents)])
        self.approch_goal_metric = 1.0


class ODynamicSameGoal(OStaticSameGoal):
    """obstacles/o_dynamic_same_goal.py: the common goal hops to a random free cell at most 4 m away, on the first tick
    and then every 4-6 s."""
    mode = 'o_dynamic_same_goal'
    dynamic = True
    max_dist = 4.0

    def reset(self, obst_map=None, cell_centers=None):
        self.period = int(self.rng.uniform(low=4.0, high=6.0) * self.control_freq)
        self._free_cells(obst_map, cell_centers)
        start = self._sample_free_points(self.num_agents)
        self.end_point = self._largest_free_square_center()
        self.pick_formation()
        self.spawn_points = start.copy()
        self.goals = np.array([self.end_point for _ in range(self.num_agents)])
        self.approch_goal_metric = 1.0

    def step(self, tick):
        if tick % self.period == 0 or tick == 1:
            new_goal = self._sample_free_point()
            while np.linalg.norm(self.end_point - new_goal) > self.max_dist:
                new_goal = self._sample_free_point()
            self.end_point = new_goal
            self.goals = np.array([new_goal for _ in range(self.num_agents)])


class OSwapGoals(_ObstacleScenario):
    """obstacles/o_swap_goals.py: a formation around the centre of the largest free square; the goals are permuted among
    the drones every 4-6 s."""
    mode = 'o_swap_goals'
    dynamic = True

    _largest_free_square_center = OStaticSameGoal._largest_free_square_center

    def reset(self, obst_map=None, cell_centers=None):
        self._free_cells(obst_map, cell_centers)
        self.period = int(self.rng.uniform(low=4.0, high=6.0) * self.control_freq)
        self.pick_formation()
        start = self._sample_free_points(self.num_agents)
        self.spawn_points = start.copy()
        self.formation_center = self._largest_free_square_center()
        self.goals = self.make_goals()
        self.rng.shuffle(self.goals)
        self.approch_goal_metric = 1.0                # o_base.py:16

    def step(self, tick):
        if tick % self.period == 0 and tick > 0:
            self.goals = self.goals.copy()
            self.rng.shuffle(self.goals)


class OEpRandBezier(_ObstacleScenario):
    """obstacles/o_ep_rand_bezier.py: spawn on free cells, one common goal above a random free cell that then follows
    quadratic Bezier segments of 6 s; control points 2..5 m away (randint(2.5, 6) truncates its bounds), accepted when both
    lie in x, y within +-4.5 m and z in (2, 2.5)."""
    mode = 'o_ep_rand_bezier'
    dynamic = True
    num_secs = 6
    z_low, z_high, cap = 1.5, 3.0, 5
    _bounds = RandBezier._bounds
    step = RandBezier.step

    def reset(self, obst_map=None, cell_centers=None):
        self._free_cells(obst_map, cell_centers)
        start = self._sample_free_points(self.num_agents)
        end = self._sample_free_point()
        # o_ep_rand_bezier.py:74-93: ten "trajectory points" are drawn with distance rejections and never used afterwards;
        # the draws are repeated here (quirks included: the index into the shrinking free-cell list is used as an index into
        # the cell-centre table) so that the stream stays aligned with the reference
        picked = []
        free = list(self.free_space)
        while len(picked) < 10:
            idx = self.rng.choice(len(free))
            if picked and np.any(np.array([np.linalg.norm(cell_centers[q] - cell_centers[idx]) for q in picked]) > 4.0):
                continue
            picked.append(idx)
            free.pop(idx)
        self.pick_formation()
        self.spawn_points = start.copy()
        self.goals = np.array([end for _ in range(self.num_agents)])
        self.approch_goal_metric = 1.0


SCENARIOS = {c.mode: c for c in (StaticSameGoal, StaticDiffGoal, DynamicSameGoal, DynamicDiffGoal, SwapGoals,
                                  DynamicFormations, Lissajous3D, RandBezier, RunAway, SwarmVsSwarm, ORandom, OStaticSameGoal,
                                  ODynamicSameGoal, OSwapGoals, OEpRandBezier)}


class Mix(Scenario):
    """mix.py:37-93: a fresh scenario drawn uniformly per episode from the reference's mode lists (scenarios/utils.py:7-28)."""
    mode = 'mix'
    dynamic = True

    def __init__(self, num_agents, **kw):
        super().__init__(num_agents, **kw)
        self.kw = kw
        if num_agents == 1:
            self.modes = MIX_MODES_OBST_SINGLE if self.use_obstacles else MIX_MODES_SINGLE
        else:
            self.modes = MIX_MODES_OBST if self.use_obstacles else MIX_MODES
        self.scenario = None

    def name(self):
        return self.scenario.name()

    def reset(self, obst_map=None, cell_centers=None):
        while True:
            mode = self.modes[self.rng.randint(low=0, high=len(self.modes))]
            if mode in SCENARIOS:
                break
        kw = dict(self.kw)
        kw['rng'] = self.rng
        self.scenario = SCENARIOS[mode](self.num_agents, **kw)
        self.scenario.reset(obst_map, cell_centers)
        self._sync()

    def step(self, tick):
        self.scenario.step(tick)
        self._sync()

    def _sync(self):
        self.goals = self.scenario.goals
        self.spawn_points = self.scenario.spawn_points
        self.formation_size = self.scenario.formation_size
        self.approch_goal_metric = self.scenario.approch_goal_metric


def create_scenario(quads_mode, num_agents, room_dims=(10., 10., 10.), rng=None, control_freq=100.0, ep_time=15.0,
                    box=None, use_obstacles=False):
    """Factory with the reference's mode names (mix.py:31-34).  Unknown names raise, as the reference does
    for CLI choices without a class."""
    if box is None:
        box = 0.1 if use_obstacles else 2.0
    kw = dict(room_dims=room_dims, rng=rng, control_freq=control_freq, ep_time=ep_time, box=box,
              use_obstacles=use_obstacles)
    if quads_mode == 'mix':
        return Mix(num_agents, **kw)
    if quads_mode not in SCENARIOS:
        raise NameError(f"Scenario_{quads_mode} is not defined")
    return SCENARIOS[quads_mode](num_agents, **kw)
