"""Philox4x32-10 counter RNG + the draw-site table shared with the CUDA kernels.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  This is the numpy/python twin of
`quad_swarm_rl_b200/csrc/qs_rng.cuh`; the two files define the same function
    (seed, env, step_count, site, i, j, value_index) -> random value
so that the CPU oracle and the CUDA kernels consume *identical* random numbers
("identical seeds" in BASELINE.json:north_star) and parity tests need no noise tensors.

The reference draws from two order-dependent Mersenne-Twister streams (numba's and
numpy's global one); a GPU cannot replay those, so the product
uses the keyed draws below and the oracle can run on either source (see
quadswarm_oracle.ReplayRng / PhiloxRng).

Conversions (exact in fp32 and fp64, so both sides start from the same number):
    uniform01(x) = (x >> 8) * 2**-24                        in [0, 1)
    normal pair (xa, xb):  u1 = ((xa >> 9) + 0.5) * 2**-23  in (0, 1)
                           u2 = (xb >> 8) * 2**-24
                           r = sqrt(-2 ln u1);  n0 = r cos(2 pi u2);  n1 = r sin(2 pi u2)
A 4-word Philox block yields 4 uniforms, or 4 normals (words 0,1 -> n0,n1; words 2,3 -> n2,n3).
Value index v lives in block v // 4, word v % 4.
Counter = (env_id, step_count, site | i << 8 | j << 16, block); key = (seed_lo, seed_hi).
"""
import math

M0 = 0xD2511F53
M1 = 0xCD9E8D57
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = 0xFFFFFFFF

# ---- draw sites (must match qs_rng.cuh) -------------------------------------------------
SITE_OU = 0            # (i)    normals v0..3                        numba_utils.py:103
SITE_FLOOR_YAW = 1     # (i)    uniforms v0 (sub-step 0), v1 (sub-step 1)   quadrotor_dynamics.py:617
SITE_SENSOR0 = 2       # (i)    normals v0..2 pos, v3..5 vel, v6..8 gyro    sensor_noise.py:241-251
SITE_SENSOR1 = 3       # (i)    same layout; the re-draw after a contact response  quadrotor_multi.py:598-599
SITE_SENSOR_RESET = 4  # (i)    same layout; the observation returned by an (auto-)reset
SITE_DW_I = 5          # (i)    uniforms v0 acc noise, v1 omega noise        downwash.py:30,35
SITE_DW_IJ = 6         # (i,j)  uniforms v0..2 z-axis noise, v3..5 omega dir downwash.py:56,62
SITE_PAIR_N = 7        # (i<j)  normals, try t: v[12t+0..2] shared, [12t+3..5] n1, [12t+6..8] n2  collisions/quadrotors.py:36-38
SITE_PAIR_U = 8        # (i<j)  uniforms v0 decay1, v1 decay2, v2..4 omega dir, v5 omega mag     collisions/utils.py:9,26,30
SITE_OBST_N = 9        # (i)    normals, try t: v[8t+0..2] shared, v[8t+3..5] own                collisions/obstacles.py:33-34
SITE_OBST_U = 10       # (i)    uniforms v0 decay, v1..3 omega dir, v4 omega mag
SITE_WALL_U = 11       # (i)    uniforms v0 speed, v1..3 dir, v4 x, v5 y, v6 z, v7..9 omega dir, v10 omega mag  collisions/room.py:10-40
SITE_CEIL_U = 12       # (i)    uniforms v0 speed, v1..3 dir, v4 z, v5..7 omega dir, v8 omega mag               collisions/room.py:94-110
SITE_SPAWN_U = 13      # (i)    uniforms v0..2                         quadrotor_single.py:394
SITE_RESET_YAW_U = 14  # (i)    uniforms v[k], k = rejection try       quadrotor_single.py:432-434
SITE_SCENARIO_U = 15   # (slot) uniforms, env-level scenario generators (see scenario_gen.py)
SITE_HOT = 16          # (i)    compact layout of the two draws every drone makes every step: SITE_OU v0..3 and
                       #        SITE_SENSOR0 v0..8 are served from TWO blocks of this site, one word per normal pair
                       #        (hot_normal below); every other site keeps the full-precision layout above

RESET_YAW_MAX_TRIES = 64
EPISODE_KEY_BIT = 0x80000000     # counter word 1 of the episode-defining draws = this bit | episode number (qs_rng.cuh)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """One Philox4x32-10 block (Salmon et al., SC'11). All arguments/results are uint32 python ints."""
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> 32, p0 & MASK
        hi1, lo1 = p1 >> 32, p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


def u01(x):
    return (x >> 8) * (2.0 ** -24)


def normal_pair(xa, xb):
    u1 = ((xa >> 9) + 0.5) * (2.0 ** -23)
    u2 = (xb >> 8) * (2.0 ** -24)
    r = math.sqrt(-2.0 * math.log(u1))
    a = 2.0 * math.pi * u2
    return r * math.cos(a), r * math.sin(a)


def normal_pair16(x):
    """Compact pair of one word: u1 = ((x >> 16) + 0.5) 2^-16, u2 = (x & 0xffff) 2^-16 (|n| <= 4.85)."""
    u1 = ((x >> 16) + 0.5) * (2.0 ** -16)
    u2 = (x & 0xFFFF) * (2.0 ** -16)
    r = math.sqrt(-2.0 * math.log(u1))
    a = 2.0 * math.pi * u2
    return r * math.cos(a), r * math.sin(a)


# (site, v) -> index of the normal inside the SITE_HOT stream: OU 0..3 -> 0..3, SENSOR0 0..8 -> 4..12;
# stream index n lives in word n // 2 of the two blocks (words 0..3 = block 0, 4..7 = block 1), half n % 2
def hot_index(site, v):
    if site == SITE_OU and 0 <= v < 4:
        return v
    if site == SITE_SENSOR0 and 0 <= v < 9:
        return 4 + v
    return None


class KeyedDraws:
    """(env, step_count)-scoped view of the keyed generator with a small block cache."""

    def __init__(self, seed, env_id, step_count):
        self.k0 = seed & MASK
        self.k1 = (seed >> 32) & MASK
        self.env_id = env_id & MASK
        self.step_count = step_count & MASK
        self._cache = {}

    def _block(self, site, i, j, b):
        key = (site, i, j, b)
        blk = self._cache.get(key)
        if blk is None:
            c2 = (site | (i << 8) | (j << 16)) & MASK
            blk = philox4x32_10(self.env_id, self.step_count, c2, b & MASK, self.k0, self.k1)
            self._cache[key] = blk
        return blk

    def uniform(self, site, i, j, v):
        return u01(self._block(site, i, j, v // 4)[v % 4])

    def normal(self, site, i, j, v):
        if site in (SITE_SENSOR1, SITE_SENSOR_RESET) and 0 <= v < 9:
            # re-drawn sensor noise: the same compact layout inside the site's own blocks 0 / 1 (qs_device.cuh, sensor_noise)
            word = v // 2
            return normal_pair16(self._block(site, i, 0, word // 4)[word % 4])[v % 2]
        n = hot_index(site, v)
        if n is not None:
            word = n // 2
            blk = self._block(SITE_HOT, i, 0, word // 4)
            return normal_pair16(blk[word % 4])[n % 2]
        blk = self._block(site, i, j, v // 4)
        w = v % 4
        pair = normal_pair(blk[0], blk[1]) if w < 2 else normal_pair(blk[2], blk[3])
        return pair[w & 1]
