// Device-side episode generators of the obstacle-free scenario family: goal formations, the per-episode
// scenario draw of `mix`, and the timed goal switches, all inside the step / reset kernels — no host work per episode or
// per tick.  Reference behaviour (paths under gym_art/quadrotor_multi/scenarios/):
//   formations           base.py:39-113, utils.py:74-121,149-175
//   pick of a formation  base.py:123-136, utils.py:32-65,124-146
//   static_same_goal / static_diff_goal      base.py:138-150 (standard_reset)
//   dynamic_same_goal    dynamic_same_goal.py      dynamic_diff_goal   dynamic_diff_goal.py
//   swap_goals           swap_goals.py             dynamic_formations  dynamic_formations.py:21-35
//   ep_lissajous3D       ep_lissajous3D.py         swarm_vs_swarm      swarm_vs_swarm.py:8-101
//   mix                  mix.py:37-93 (the two bezier modes need a third-party package and are left out, as in scenarios.py)
// The reference draws from numpy's global stream; here every draw is a keyed Philox value (SITE_SCENARIO_U, stream 1 = at
// reset, stream 2 = at a tick event, slot v below), so an episode is a pure function of (seed, env id, step counter).
// oracle/scenario_gen.py restates this file draw for draw; the formation geometry of both is pinned to the reference by
// tests/golden/formations.npz.
//
// Everything here is lane-local except the goal permutation of swap_goals (one warp shuffle): each lane evaluates the
// formation point of the index it was dealt, so no shared memory and no exchange is needed.
#pragma once
#include "qs_device.cuh"

namespace qs {

// draw slots (v) inside a stream
enum ScnSlot {
    SV_MIX = 0, SV_PERIOD = 1, SV_FORMATION = 2, SV_SIZE = 3, SV_LAYER = 4, SV_CX = 5, SV_CY = 6, SV_CZ = 7,
    SV_DIST = 8, SV_PHI = 9, SV_THETA = 10, SV_GROW = 11, SV_SPEED = 12, SV_RUN0 = 13, SV_RUN1 = 14,
    SV_SHUFFLE = 16,        // SV_SHUFFLE + drone index
    SV_BEZIER = 64          // + 8 * try + j: six direction uniforms (j = 0..5) and the distance draw (j = 6) of a rejection try
};
constexpr int BEZIER_STEPS = 500;      // int(num_secs * control_freq), ep_rand_bezier.py:13-14
constexpr int BEZIER_MAX_TRIES = 512;  // the reference re-draws until both control points lie in the room (acceptance ~5 %)
constexpr int SCN_STREAM_RESET = 1, SCN_STREAM_TICK = 2;
constexpr int SCN_NEVER = 0x7fffffff;
constexpr float SCN_BOX = 2.0f;                    // scenario box of the obstacle-free family (base.py:18, quadrotor_multi.py:118)
constexpr float SCN_CONTROL_FREQ = 100.0f;

struct Formation { int f, per_layer; float lo, hi, size, layer; };
struct ScnState {
    int mode, period, next, f, growing;
    float size, layer, hi, speed;
    V3 c1, c2;
};
struct ScnOut { V3 goal; int next; };

__device__ __forceinline__ uint32_t scn_u24(const RngKey& key, int stream, int v) {
    const uint4 b = rng_block(key, SITE_SCENARIO_U, 0, stream, v >> 2);
    const int w = v & 3;
    const uint32_t x = w == 0 ? b.x : (w == 1 ? b.y : (w == 2 ? b.z : b.w));
    return x >> 8;
}
__device__ __forceinline__ float scn_u(const RngKey& key, int stream, int v) { return (float)scn_u24(key, stream, v) * (1.0f / 16777216.0f); }
// floor(u * n) in integer arithmetic (identical in the fp64 twin)
__device__ __forceinline__ int scn_pick(const RngKey& key, int stream, int v, int n) {
    return (int)(((unsigned long long)scn_u24(key, stream, v) * (unsigned long long)n) >> 24);
}

// utils.py:109-121: largest divisor of num not above sqrt(num), and its cofactor
__device__ __forceinline__ void grid_dims(int num, int& d1, int& d2) {
    d1 = (int)floorf(sqrtf((float)num));
    while (d1 > 1 && (num % d1) != 0) --d1;
    d1 = max(d1, 1);
    d2 = num / d1;
}

__device__ __forceinline__ V3 place(int plane, float a, float b, float l) {       // utils.py:149-160
    V3 r;
    if (plane == 0) { r.x = a; r.y = b; r.z = l; }             // horizontal
    else if (plane == 1) { r.x = a; r.y = l; r.z = b; }        // vertical_xz
    else { r.x = l; r.y = a; r.z = b; }                        // vertical_yz
    return r;
}

// point k of an n-point formation before centring (base.py:39-113).  f: 0-2 circle_{horizontal,vertical_xz,vertical_yz},
// 3 sphere, 4-6 grid_*, 7 cube.
__device__ V3 formation_raw(int f, int n, int k, float size, float layer_dist, int per_layer) {
    if (f <= 2) {
        const int layer = k / per_layer;
        const int m = (n > per_layer) ? ((layer < n / per_layer) ? per_layer : (n % per_layer)) : n;
        const float ang = 2.0f * PI_F * (float)(k % m) / (float)m;
        float sn, cs;
        sincosf(ang, &sn, &cs);
        return place(f, size * cs, size * sn, (float)layer * layer_dist);
    }
    if (f == 3) {                                               // utils.py:74-90 (at least 3 points are generated)
        const float nn = (float)max(n, 3);
        const float x = 0.1f + 1.2f * nn;
        const float start = -1.0f + 1.0f / (nn - 1.0f);
        const float inc = (2.0f - 2.0f / (nn - 1.0f)) / (nn - 1.0f);
        const float s = start + inc * (float)k;
        const float lon = s * x;
        const float sg = s > 0.f ? 1.f : (s < 0.f ? -1.f : 0.f);
        const float lat = 0.5f * PI_F * sg * (1.0f - sqrtf(1.0f - fabsf(s)));
        float sl, cl, sa, ca;
        sincosf(lon, &sl, &cl);
        sincosf(lat, &sa, &ca);
        V3 r = {size * cl * ca, size * sl * ca, size * sa};
        return r;
    }
    if (f <= 6) {
        const int layer = k / per_layer;
        const int nl = (n <= per_layer) ? n : ((layer < n / per_layer) ? per_layer : (n % per_layer));
        int d1, d2;
        grid_dims(nl, d1, d2);
        return place(f - 4, size * (float)(k % d2), size * (float)((k / d2) % d1), (float)layer * layer_dist);
    }
    // cube, base.py:97-109: side = int(np.power(n, 1/3)) — float64 gives 2.9999999999999996 for n = 27, so the side is
    // 2 for 8 <= n <= 27 and 3 for 28 <= n <= 32 (tests/test_host_logic.py checks this table against numpy)
    const int side = n >= 28 ? 3 : (n >= 8 ? 2 : 1);
    V3 r = {size * (float)(k / (side * side)), size * (float)((k / side) % side), size * (float)(k % side)};
    return r;
}

__device__ V3 formation_point(int f, int n, int k, float size, V3 c, float layer_dist, int per_layer) {
    V3 r = formation_raw(f, n, k, size, layer_dist, per_layer);
    if (f >= 4) {                                               // grids and the cube are centred on their mean
        float mx = 0.f, my = 0.f, mz = 0.f;
#pragma unroll 1
        for (int q = 0; q < n; ++q) {
            const V3 t = formation_raw(f, n, q, size, layer_dist, per_layer);
            mx += t.x; my += t.y; mz += t.z;
        }
        const float inv = 1.0f / (float)n;
        r.x -= mx * inv; r.y -= my * inv; r.z -= mz * inv;
    }
    r.x += c.x; r.y += c.y; r.z += c.z;
    return r;
}

// base.py:123-136 + utils.py:32-65,124-146.  `n` = drones per formation (N/2 for swarm_vs_swarm).
__device__ Formation pick_formation(const RngKey& key, int stream, int mode, int n) {
    int count = 8;
    float low = 0.25f, high = 0.5f;                              // 5 / 10 nominal arm lengths (utils.py:31-51)
    if (mode == QS_SCENARIO_STATIC_SAME_GOAL || mode == QS_SCENARIO_DYNAMIC_SAME_GOAL || mode == QS_SCENARIO_EP_LISSAJOUS3D ||
        mode == QS_SCENARIO_EP_RAND_BEZIER) {
        count = 1; low = 0.f; high = 0.f;
    } else if (mode == QS_SCENARIO_SWAP_GOALS) {
        low = 0.4f; high = 0.8f;
    } else if (mode == QS_SCENARIO_O_SWAP_GOALS) {
        // utils.py:55-57 indexes QUADS_FORMATION_LIST with a draw below len(QUADS_FORMATION_LIST_OBSTACLES) = 7: every
        // formation except the cube, circle_horizontal included
        count = 7; low = 0.4f; high = 0.8f;
    } else if (mode == QS_SCENARIO_DYNAMIC_FORMATIONS) {
        low = 0.f; high = 1.0f;
    }
    Formation fm;
    fm.f = count > 1 ? scn_pick(key, stream, SV_FORMATION, count) : 0;
    fm.per_layer = (fm.f >= 4 && fm.f <= 6) ? 50 : 8;
    if (fm.f <= 2) {                                             // utils.py:102-106 with num = drones per layer
        const float inv = 0.5f / sinf(PI_F / (float)fm.per_layer);
        fm.lo = low * inv; fm.hi = high * inv;
    } else if (fm.f == 3) {                                      // utils.py:92-99
        const float A = 1.75388487222762f, B = 0.860487305801679f, C = 10.3632729642351f, D = 0.0920858134405214f;
        const float inv = 1.0f / ((A - D) / (1.0f + powf((float)n / C, B)) + D);
        fm.lo = low * inv; fm.hi = high * inv;
    } else {
        fm.lo = low; fm.hi = high;
    }
    fm.size = fm.lo + (fm.hi - fm.lo) * scn_u(key, stream, SV_SIZE);
    fm.layer = fm.lo + (fm.hi - fm.lo) * scn_u(key, stream, SV_LAYER);
    return fm;
}

// utils.py:163-175
__device__ float z_above_ground(float u, int num_agents, int per_layer, int f, float size) {
    const float z = (-0.5f * SCN_BOX + SCN_BOX * u) + 2.0f;
    float lower = 0.25f;
    if (f >= 1 && f <= 3) lower = size + 0.25f;
    else if (f == 5 || f == 6) {
        int d1, d2;
        grid_dims(min(num_agents, per_layer), d1, d2);
        lower = (float)d1 * size + 0.25f;
    }
    return fmaxf(lower, z);
}

// rank of drone i among drones [g0, g1) by their shuffle keys = the index it is dealt by a uniform random permutation
__device__ int shuffle_rank(const RngKey& key, int stream, int i, int g0, int g1) {
    const uint32_t ui = scn_u24(key, stream, SV_SHUFFLE + i);
    int r = 0;
#pragma unroll 1
    for (int j = g0; j < g1; ++j) {
        const uint32_t uj = scn_u24(key, stream, SV_SHUFFLE + j);
        r += (uj < ui || (uj == ui && j < i)) ? 1 : 0;
    }
    return r;
}

__device__ __forceinline__ void scn_store_to(int4* si, float4* sf, const ScnState& s) {
    si[0] = make_int4(s.mode, s.period, s.next, s.f | (s.growing << 8));
    sf[0] = make_float4(s.size, s.layer, s.hi, s.speed);
    sf[1] = make_float4(s.c1.x, s.c1.y, s.c1.z, 0.f);
    sf[2] = make_float4(s.c2.x, s.c2.y, s.c2.z, 0.f);
}
__device__ __forceinline__ void scn_store(const DevState& st, int env, const ScnState& s) {
    scn_store_to(st.scn_i + env, st.scn_f + 3 * (long long)env, s);
}
__device__ __forceinline__ ScnState scn_load(const DevState& st, int env) {
    ScnState s;
    const int4 a = QS_LD(st.scn_i + env);
    const float4 b = QS_LD(st.scn_f + 3 * (long long)env + 0), c = QS_LD(st.scn_f + 3 * (long long)env + 1),
                 d = QS_LD(st.scn_f + 3 * (long long)env + 2);
    s.mode = a.x; s.period = a.y; s.next = a.z; s.f = a.w & 0xff; s.growing = (a.w >> 8) & 1;
    s.size = b.x; s.layer = b.y; s.hi = b.z; s.speed = b.w;
    s.c1.x = c.x; s.c1.y = c.y; s.c1.z = c.z; s.c2.x = d.x; s.c2.y = d.y; s.c2.z = d.z;
    return s;
}

__device__ __forceinline__ int per_layer_of(int f) { return (f >= 4 && f <= 6) ? 50 : 8; }

// swarm_vs_swarm.py:31-72: the two formation centres
__device__ void svs_centers(const RngKey& key, int N, const Formation& fm, V3& c1, V3& c2) {
    const float box = SCN_BOX;
    c1.x = -box + 2.0f * box * scn_u(key, SCN_STREAM_RESET, SV_CX);
    c1.y = -box + 2.0f * box * scn_u(key, SCN_STREAM_RESET, SV_CY);
    c1.z = z_above_ground(scn_u(key, SCN_STREAM_RESET, SV_CZ), N, fm.per_layer, fm.f, fm.size);
    const float dist = 0.25f * box + (box - 0.25f * box) * scn_u(key, SCN_STREAM_RESET, SV_DIST);
    const float phi = -PI_F + 2.0f * PI_F * scn_u(key, SCN_STREAM_RESET, SV_PHI);
    const float theta = -0.5f * PI_F + PI_F * scn_u(key, SCN_STREAM_RESET, SV_THETA);
    float sp, cp, st, ct;
    sincosf(phi, &sp, &cp);
    sincosf(theta, &st, &ct);
    c2.x = c1.x + dist * st * cp; c2.y = c1.y + dist * st * sp; c2.z = c1.z + dist * ct;
    // formations that lie in a plane are kept apart along the plane's normal (swarm_vs_swarm.py:52-70)
    const int plane = fm.f <= 2 ? fm.f : (fm.f >= 4 && fm.f <= 6 ? fm.f - 4 : -1);
    if (plane >= 0) {
        float* a2 = plane == 0 ? &c2.z : (plane == 1 ? &c2.y : &c2.x);
        const float a1 = plane == 0 ? c1.z : (plane == 1 ? c1.y : c1.x);
        const float diff = *a2 - a1;
        if (fabsf(diff) < fm.lo) *a2 = (diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f)) * fm.lo + a1;
    }
}

__device__ V3 svs_goal(const ScnState& s, int N, int k) {
    const int h = N / 2;
    return k < h ? formation_point(s.f, h, k, s.size, s.c1, s.layer, per_layer_of(s.f))
                 : formation_point(s.f, N - h, k - h, s.size, s.c2, s.layer, per_layer_of(s.f));
}

// scenario.reset() of the env's mode (for `mix`: of the scenario drawn for this episode).  Every lane of the env computes
// the same scenario state; lane 0 stores it to (si, sf): the env's live scenario state, or its next-episode record.
__device__ __noinline__ ScnOut scenario_reset(RngKey key, int cfg_mode, int N, int i, int4* si, float4* sf) {
    ScnState s;
    s.mode = cfg_mode;
    if (cfg_mode == QS_SCENARIO_MIX) {
        // mix.py:59-63 draws uniformly from the mode list of scenarios/utils.py:7-16 (9 modes, 5 for a single drone)
        if (N == 1) {
            const int m5[5] = {QS_SCENARIO_STATIC_SAME_GOAL, QS_SCENARIO_STATIC_DIFF_GOAL, QS_SCENARIO_EP_LISSAJOUS3D,
                               QS_SCENARIO_EP_RAND_BEZIER, QS_SCENARIO_DYNAMIC_SAME_GOAL};
            s.mode = m5[scn_pick(key, SCN_STREAM_RESET, SV_MIX, 5)];
        } else {
            const int k = scn_pick(key, SCN_STREAM_RESET, SV_MIX, 9);
            s.mode = k < 8 ? QS_SCENARIO_STATIC_SAME_GOAL + k : QS_SCENARIO_EP_RAND_BEZIER;
        }
    }
    s.period = 0; s.next = SCN_NEVER; s.growing = 0; s.speed = 0.f;
    const bool svs = s.mode == QS_SCENARIO_SWARM_VS_SWARM;
    const Formation fm = pick_formation(key, SCN_STREAM_RESET, s.mode, svs ? N / 2 : N);
    s.f = fm.f; s.size = fm.size; s.layer = fm.layer; s.hi = fm.hi;
    s.c1.x = 0.f; s.c1.y = 0.f; s.c1.z = 2.0f;
    s.c2 = s.c1;
    if (s.mode == QS_SCENARIO_DYNAMIC_SAME_GOAL || s.mode == QS_SCENARIO_DYNAMIC_DIFF_GOAL || s.mode == QS_SCENARIO_SWAP_GOALS || svs) {
        s.period = 400 + scn_pick(key, SCN_STREAM_RESET, SV_PERIOD, 200);       // int(U(4, 6) s * 100 Hz)
        s.next = s.period;
    } else if (s.mode == QS_SCENARIO_RUN_AWAY) {
        s.period = 100; s.next = 100;                                           // run_away.py:15-18: every second, never at tick 0
    }
    ScnOut o;
    if (svs) {
        svs_centers(key, N, fm, s.c1, s.c2);
        o.goal = svs_goal(s, N, i);                                               // not shuffled at reset
    } else if (s.mode == QS_SCENARIO_EP_LISSAJOUS3D) {
        s.c1.x = -2.0f; s.c1.y = 0.f; s.c1.z = 2.0f;
        s.period = 1; s.next = 1;
        o.goal = s.c1;                                                            // size 0, layer distance 0
    } else if (s.mode == QS_SCENARIO_EP_RAND_BEZIER) {
        // standard_reset with formation size 0: every goal sits on the centre (0, 0, 2); the curve is drawn at tick 1.
        // (size, layer, hi) hold P0 and (c1, c2) the control points P1, P2 of the running segment.
        s.period = 1; s.next = 1;
        s.size = 0.f; s.layer = 0.f; s.hi = 2.0f;
        o.goal = s.c1;
    } else {
        if (s.mode == QS_SCENARIO_DYNAMIC_FORMATIONS) {
            s.growing = scn_u24(key, SCN_STREAM_RESET, SV_GROW) < (1u << 23) ? 1 : 0;
            s.speed = 1.0f + 2.0f * scn_u(key, SCN_STREAM_RESET, SV_SPEED);
            s.period = 1; s.next = 1;
        }
        const int k = shuffle_rank(key, SCN_STREAM_RESET, i, 0, N);               // standard_reset shuffles the goals
        o.goal = formation_point(s.f, N, k, s.size, s.c1, s.layer, fm.per_layer);
    }
    o.next = s.next;
    if (i == 0) scn_store_to(si, sf, s);
    return o;
}

// scenario.step() on a tick where something happens (`active` lanes; the others pass through).  Called from a
// warp-uniform branch: swap_goals permutes the goals with a shuffle.
template <int NP>
__device__ __noinline__ ScnOut scenario_tick(RngKey key, int N, int i, int tick, V3 goal, bool active, DevState st, int env) {
    ScnOut o;
    o.goal = goal; o.next = SCN_NEVER;
    ScnState s;
    s.mode = -1;
    if (active) s = scn_load(st, env);
    // swap_goals.py: goals are permuted among the drones
    int src = (active && s.mode == QS_SCENARIO_SWAP_GOALS) ? shuffle_rank(key, SCN_STREAM_TICK, i, 0, N) : i;
    // run_away.py:19-24: goals[0] = goals[g0], goals[1] = goals[g1] with g0, g1 ~ randint(1, N): neither reads goals[0], so
    // the two sequential assignments equal one gather of the old goals
    if (active && s.mode == QS_SCENARIO_RUN_AWAY && i < 2 && N >= 2) src = 1 + scn_pick(key, SCN_STREAM_TICK, i == 0 ? SV_RUN0 : SV_RUN1, N - 1);
    const float gx = shfl<NP>(goal.x, src), gy = shfl<NP>(goal.y, src), gz = shfl<NP>(goal.z, src);
    if (active) {
        if (s.mode == QS_SCENARIO_SWAP_GOALS || s.mode == QS_SCENARIO_RUN_AWAY) {
            o.goal.x = gx; o.goal.y = gy; o.goal.z = gz;
        } else if (s.mode == QS_SCENARIO_DYNAMIC_SAME_GOAL) {
            s.c1.x = -SCN_BOX + 2.0f * SCN_BOX * scn_u(key, SCN_STREAM_TICK, SV_CX);
            s.c1.y = -SCN_BOX + 2.0f * SCN_BOX * scn_u(key, SCN_STREAM_TICK, SV_CY);
            s.c1.z = fmaxf(0.25f, (-0.5f * SCN_BOX + SCN_BOX * scn_u(key, SCN_STREAM_TICK, SV_CZ)) + 2.0f);
            o.goal = formation_point(s.f, N, i, s.size, s.c1, 0.0f, per_layer_of(s.f));
        } else if (s.mode == QS_SCENARIO_DYNAMIC_DIFF_GOAL) {
            s.c1.x = -SCN_BOX + 2.0f * SCN_BOX * scn_u(key, SCN_STREAM_TICK, SV_CX);
            s.c1.y = -SCN_BOX + 2.0f * SCN_BOX * scn_u(key, SCN_STREAM_TICK, SV_CY);
            s.c1.z = z_above_ground(scn_u(key, SCN_STREAM_TICK, SV_CZ), N, per_layer_of(s.f), s.f, s.size);   // old formation
            const Formation fm = pick_formation(key, SCN_STREAM_TICK, s.mode, N);
            s.f = fm.f; s.size = fm.size; s.layer = fm.layer; s.hi = fm.hi;
            o.goal = formation_point(s.f, N, shuffle_rank(key, SCN_STREAM_TICK, i, 0, N), s.size, s.c1, s.layer, fm.per_layer);
        } else if (s.mode == QS_SCENARIO_DYNAMIC_FORMATIONS) {
            if (s.size <= -s.hi) {
                s.growing = 1; s.speed = 1.0f + 2.0f * scn_u(key, SCN_STREAM_TICK, SV_SPEED);
            } else if (s.size >= s.hi) {
                s.growing = 0; s.speed = 1.0f + 2.0f * scn_u(key, SCN_STREAM_TICK, SV_SPEED);
            }
            s.size += (s.growing ? 0.001f : -0.001f) * s.speed;
            o.goal = formation_point(s.f, N, i, s.size, s.c1, s.layer, per_layer_of(s.f));      // unshuffled
        } else if (s.mode == QS_SCENARIO_EP_LISSAJOUS3D) {
            const float t = (float)tick / SCN_CONTROL_FREQ;
            // every drone follows drone 0's goal (ep_lissajous3D.py:20-27) — all goals are equal from the reset on, so each
            // lane advances its own copy; the "+ 90" is in radians, as in the reference
            o.goal.x = goal.x + 0.03f * sinf(t);
            o.goal.y = goal.y + 0.01f * sinf(2.0f * t + 90.0f);
            o.goal.z = goal.z + 0.01f * cosf(2.0f * t + 90.0f);
        } else if (s.mode == QS_SCENARIO_EP_RAND_BEZIER) {
            // ep_rand_bezier.py:7-50 with room_dims - formation_size = the room (formation size 0): two control points at
            // distance d = randint(5, 11) in independent random directions from the current goal, re-drawn until both lie
            // 0.5 m inside the box [-5, 5] x [-5, 5] x [0, 10] (the reference's loop has no bound; 512 tries fail with
            // probability < 1e-10); then the goal follows B(s) = (1-s)^2 P0 + 2 (1-s) s P1 + s^2 P2, s = t / 499
            const int t = tick % BEZIER_STEPS;
            if (t == 0 || tick == 1) {
                const float hx = 5.0f, hy = 5.0f, hz = 10.0f;
                V3 p1 = goal, p2 = goal;
#pragma unroll 1
                for (int k = 0; k < BEZIER_MAX_TRIES; ++k) {
                    const int v0 = SV_BEZIER + 8 * k;
                    // uniform(low=-high, high=high, size=(2, 3)).reshape(3, 2): draws u0..u5 in order, column c takes
                    // (u[c], u[2 + c], u[4 + c]) with ranges (x, y | z, x | y, z): the reference's reshape quirk is kept
                    const float u0 = -hx + 2.f * hx * scn_u(key, SCN_STREAM_TICK, v0 + 0), u1 = -hy + 2.f * hy * scn_u(key, SCN_STREAM_TICK, v0 + 1),
                                u2 = -hz + 2.f * hz * scn_u(key, SCN_STREAM_TICK, v0 + 2), u3 = -hx + 2.f * hx * scn_u(key, SCN_STREAM_TICK, v0 + 3),
                                u4 = -hy + 2.f * hy * scn_u(key, SCN_STREAM_TICK, v0 + 4), u5 = -hz + 2.f * hz * scn_u(key, SCN_STREAM_TICK, v0 + 5);
                    const float d = (float)(5 + scn_pick(key, SCN_STREAM_TICK, v0 + 6, 6));
                    const float n1 = d / sqrtf(u0 * u0 + u2 * u2 + u4 * u4), n2 = d / sqrtf(u1 * u1 + u3 * u3 + u5 * u5);
                    p1.x = goal.x + u0 * n1; p1.y = goal.y + u2 * n1; p1.z = goal.z + u4 * n1;
                    p2.x = goal.x + u1 * n2; p2.y = goal.y + u3 * n2; p2.z = goal.z + u5 * n2;
                    const bool ok = p1.x > -hx + 0.5f && p1.x < hx - 0.5f && p1.y > -hy + 0.5f && p1.y < hy - 0.5f && p1.z > 0.5f && p1.z < hz - 0.5f &&
                                    p2.x > -hx + 0.5f && p2.x < hx - 0.5f && p2.y > -hy + 0.5f && p2.y < hy - 0.5f && p2.z > 0.5f && p2.z < hz - 0.5f;
                    if (ok) break;
                }
                s.size = goal.x; s.layer = goal.y; s.hi = goal.z;
                s.c1 = p1; s.c2 = p2;
            }
            if (t != 0 && tick > 1) {
                const float sp = (float)t / (float)(BEZIER_STEPS - 1), a = (1.f - sp) * (1.f - sp), b = 2.f * (1.f - sp) * sp, c = sp * sp;
                o.goal.x = a * s.size + b * s.c1.x + c * s.c2.x;
                o.goal.y = a * s.layer + b * s.c1.y + c * s.c2.y;
                o.goal.z = a * s.hi + b * s.c1.z + c * s.c2.z;
            }
        } else if (s.mode == QS_SCENARIO_SWARM_VS_SWARM) {
            const V3 t = s.c1; s.c1 = s.c2; s.c2 = t;
            const Formation fm = pick_formation(key, SCN_STREAM_TICK, s.mode, N / 2);
            s.f = fm.f; s.size = fm.size; s.layer = fm.layer; s.hi = fm.hi;
            const int h = N / 2;
            const int k = i < h ? shuffle_rank(key, SCN_STREAM_TICK, i, 0, h) : h + shuffle_rank(key, SCN_STREAM_TICK, i, h, N);
            o.goal = svs_goal(s, N, k);
        }
        s.next = s.period > 0 ? tick + s.period : SCN_NEVER;
        o.next = s.next;
        if (i == 0) scn_store(st, env, s);
    }
    __syncwarp();
    return o;
}

// ---------------------------------------------------------------------------------------------------------------------
// Ticked obstacle scenarios (obstacles/o_dynamic_same_goal.py, o_swap_goals.py, o_ep_rand_bezier.py).  Pillars, spawn cells
// and the centre of the largest free square come from o_random_episode (qs_device.cuh); this part finishes the goals and
// writes the env's scenario words.  State layout of an obstacle env (differs from ScnState: the first two floats belong to
// the per-episode pillar size / count):
//   scn_i = (mode, period, next event tick, 0)
//   scn_f[0] = (pillar radius, pillar count, P0.x, P0.y)   scn_f[1] = (P1, approch_goal_metric)   scn_f[2] = (P2, P0.z)
// with P0..P2 the running Bezier segment of o_ep_rand_bezier (zero otherwise).  Twin: oracle/scenario_gen.py.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int O_BEZIER_STEPS = 600;        // int(6 s * control_freq), o_ep_rand_bezier.py:18-19
constexpr int O_BEZIER_MAX_TRIES = 4096;   // acceptance ~0.5 %: the reference loops until it succeeds (P(fail) ~ 1e-10 here)
constexpr int O_DYN_MAX_TRIES = 256;

__device__ __forceinline__ bool ticked_obstacle_scenario(int mode) {
    return mode == QS_SCENARIO_O_DYNAMIC_SAME_GOAL || mode == QS_SCENARIO_O_SWAP_GOALS || mode == QS_SCENARIO_O_EP_RAND_BEZIER;
}

// `goal`: what o_random_episode returned (the largest-square centre for the first two modes).  Every lane of the env
// computes the same env-level values; lane 0 stores them.
__device__ __noinline__ ScnOut o_episode_extras(RngKey key, int mode, int N, int i, unsigned long long mask, int L, int W, V3 goal,
                                                float obst_r, int M_e, int4* si, float4* sf) {
    ScnOut o;
    o.goal = goal;
    int period = 0;
    V3 p0 = {0.f, 0.f, 0.f};
    if (mode == QS_SCENARIO_O_DYNAMIC_SAME_GOAL) {
        period = 400 + scn_pick(key, SCN_STREAM_RESET, SV_PERIOD, 200);           // int(U(4, 6) s * 100 Hz), :33-34
        o.next = 1;                                                                // first hop at tick 1 (:20)
    } else if (mode == QS_SCENARIO_O_SWAP_GOALS) {
        period = 400 + scn_pick(key, SCN_STREAM_RESET, SV_PERIOD, 200);
        o.next = period;
        const Formation fm = pick_formation(key, SCN_STREAM_RESET, mode, N);
        o.goal = formation_point(fm.f, N, shuffle_rank(key, SCN_STREAM_RESET, i, 0, N), fm.size, goal, fm.layer, fm.per_layer);
    } else {
        // o_ep_rand_bezier.py:71-72: the common goal starts above a random free cell, z ~ U(0.75, 3) (o_base.py:54-69)
        const int cells = L * W;
        const int c = nth_free_cell(mask, scn_pick(key, SCN_STREAM_RESET, SV_CX, cells - M_e), cells);
        const float2 xy = cell_center(c, L, W);
        o.goal.x = xy.x; o.goal.y = xy.y; o.goal.z = 0.75f + (3.0f - 0.75f) * scn_u(key, SCN_STREAM_RESET, SV_CZ);
        period = 1;
        o.next = 1;
        p0 = o.goal;
    }
    if (i == 0) {
        si[0] = make_int4(mode, period, o.next, 0);
        sf[0] = make_float4(obst_r, (float)M_e, p0.x, p0.y);
        sf[1] = make_float4(p0.x, p0.y, p0.z, 1.0f);
        sf[2] = make_float4(p0.x, p0.y, p0.z, p0.z);
    }
    return o;
}

// occupancy mask of the env's pillar table (slots beyond this episode's pillar count stand far outside the grid)
__device__ __forceinline__ unsigned long long pillar_mask(const float2* obst, int M_table, int L, int W) {
    unsigned long long mask = 0ull;
#pragma unroll 1
    for (int m = 0; m < M_table; ++m) {
        const float2 xy = QS_LD(obst + m);
        if (xy.x > 1.0e3f) continue;
        const int cid = (int)floorf(xy.x + (float)(L / 2)), rid = W - 1 - (int)floorf(xy.y + (float)(W / 2));
        mask |= 1ull << (rid * W + cid);
    }
    return mask;
}

// scenario.step() of the three scenarios on an event tick.  Called by every lane of the warp from a warp-uniform branch
// (`active` is uniform over the NP lanes of an env): the goal permutation and the parallel rejection search use shuffles.
template <int NP>
__device__ __noinline__ ScnOut obstacle_scenario_tick(RngKey key, int N, int i, int tick, V3 goal, bool active, DevState st, int env,
                                                      int L, int W, int M_table) {
    ScnOut o;
    o.goal = goal; o.next = SCN_NEVER;
    int4 a = make_int4(-1, 0, SCN_NEVER, 0);
    float4 f0 = make_float4(0.f, 0.f, 0.f, 0.f), f1 = f0, f2 = f0;
    const long long e3 = 3 * (long long)env;
    if (active) {
        a = QS_LD(st.scn_i + env);
        f0 = QS_LD(st.scn_f + e3); f1 = QS_LD(st.scn_f + e3 + 1); f2 = QS_LD(st.scn_f + e3 + 2);
    }
    const int mode = a.x, period = a.y;
    // o_swap_goals.py:14-17: the goals are permuted among the drones
    const int src = (active && mode == QS_SCENARIO_O_SWAP_GOALS) ? shuffle_rank(key, SCN_STREAM_TICK, i, 0, N) : i;
    const float gx = shfl<NP>(goal.x, src), gy = shfl<NP>(goal.y, src), gz = shfl<NP>(goal.z, src);
    int next = period > 0 ? tick + period : SCN_NEVER;

    // o_ep_rand_bezier.py:27-45: a new segment at tick 1 and at every multiple of 600 ticks.  The NP lanes of the env test
    // NP consecutive tries per round; the lowest accepted try wins, as in the reference's sequential loop.
    const int t = tick % O_BEZIER_STEPS;
    bool searching = active && mode == QS_SCENARIO_O_EP_RAND_BEZIER && (t == 0 || tick == 1);
    const bool new_segment = searching;
    V3 p1 = goal, p2 = goal;
    // every drone carries the same goal; the padding lanes of the group take part in the search with drone 0's copy
    const V3 g0 = {shfl<NP>(goal.x, 0), shfl<NP>(goal.y, 0), shfl<NP>(goal.z, 0)};
#pragma unroll 1
    for (int base = 0; base < O_BEZIER_MAX_TRIES && __any_sync(0xffffffffu, searching); base += NP) {
        bool ok = false;
        V3 q1 = g0, q2 = g0;
        if (searching) {
            const float hx = 5.0f, hy = 5.0f, hz = 3.0f;
            const int v0 = SV_BEZIER + 8 * (base + i);
            // uniform(low=-high, high=high, size=(2, 3)).reshape(3, 2): u0..u5 in draw order; column c takes (u[c], u[2 + c],
            // u[4 + c]) — the ranges follow the draw order (x, y, z, x, y, z), the reference's reshape quirk is kept
            const float u0 = -hx + 2.f * hx * scn_u(key, SCN_STREAM_TICK, v0 + 0), u1 = -hy + 2.f * hy * scn_u(key, SCN_STREAM_TICK, v0 + 1),
                        u2 = -hz + 2.f * hz * scn_u(key, SCN_STREAM_TICK, v0 + 2), u3 = -hx + 2.f * hx * scn_u(key, SCN_STREAM_TICK, v0 + 3),
                        u4 = -hy + 2.f * hy * scn_u(key, SCN_STREAM_TICK, v0 + 4), u5 = -hz + 2.f * hz * scn_u(key, SCN_STREAM_TICK, v0 + 5);
            const float d = (float)(2 + scn_pick(key, SCN_STREAM_TICK, v0 + 6, 4));     // randint(2.5, 6) truncates its bounds: 2..5
            const float n1 = d / sqrtf(u0 * u0 + u2 * u2 + u4 * u4), n2 = d / sqrtf(u1 * u1 + u3 * u3 + u5 * u5);
            q1.x = g0.x + u0 * n1; q1.y = g0.y + u2 * n1; q1.z = g0.z + u4 * n1;
            q2.x = g0.x + u1 * n2; q2.y = g0.y + u3 * n2; q2.z = g0.z + u5 * n2;
            ok = q1.x > -hx + 0.5f && q1.x < hx - 0.5f && q1.y > -hy + 0.5f && q1.y < hy - 0.5f && q1.z > 1.5f + 0.5f && q1.z < hz - 0.5f &&
                 q2.x > -hx + 0.5f && q2.x < hx - 0.5f && q2.y > -hy + 0.5f && q2.y < hy - 0.5f && q2.z > 1.5f + 0.5f && q2.z < hz - 0.5f;
        }
        const uint32_t b = group_ballot<NP>(ok);
        const int w = b ? (__ffs(b) - 1) : 0;
        const float a1 = shfl<NP>(q1.x, w), a2 = shfl<NP>(q1.y, w), a3 = shfl<NP>(q1.z, w);
        const float b1 = shfl<NP>(q2.x, w), b2 = shfl<NP>(q2.y, w), b3 = shfl<NP>(q2.z, w);
        if (searching && b) {
            p1.x = a1; p1.y = a2; p1.z = a3; p2.x = b1; p2.y = b2; p2.z = b3;
            searching = false;
        }
    }

    if (active) {
        if (mode == QS_SCENARIO_O_SWAP_GOALS) {
            o.goal.x = gx; o.goal.y = gy; o.goal.z = gz;
        } else if (mode == QS_SCENARIO_O_DYNAMIC_SAME_GOAL) {
            // o_dynamic_same_goal.py:20-28: a random free cell, z ~ U(0.75, 3), re-drawn while it is more than 4 m from the
            // current goal (the goal's own cell always qualifies; after O_DYN_MAX_TRIES failures the goal stays)
            const unsigned long long mask = pillar_mask(st.obst + (long long)env * M_table, M_table, L, W);
            const int cells = L * W, free_cells = cells - __popcll(mask);
#pragma unroll 1
            for (int k = 0; k < O_DYN_MAX_TRIES; ++k) {
                const int v0 = SV_BEZIER + 8 * k;
                const float2 xy = cell_center(nth_free_cell(mask, scn_pick(key, SCN_STREAM_TICK, v0, free_cells), cells), L, W);
                const float z = 0.75f + (3.0f - 0.75f) * scn_u(key, SCN_STREAM_TICK, v0 + 1);
                const float dx = goal.x - xy.x, dy = goal.y - xy.y, dz = goal.z - z;
                if (sqrtf(dx * dx + dy * dy + dz * dz) <= 4.0f) { o.goal.x = xy.x; o.goal.y = xy.y; o.goal.z = z; break; }
            }
            next = (tick / period + 1) * period;
        } else if (mode == QS_SCENARIO_O_EP_RAND_BEZIER) {
            V3 p0 = {f0.z, f0.w, f2.w};
            if (new_segment) {
                p0 = goal;
                f0.z = p0.x; f0.w = p0.y; f2.w = p0.z;
                f1.x = p1.x; f1.y = p1.y; f1.z = p1.z;
                f2.x = p2.x; f2.y = p2.y; f2.z = p2.z;
            }
            if (t != 0 && tick > 1) {
                const float sp = (float)t / (float)(O_BEZIER_STEPS - 1), ca = (1.f - sp) * (1.f - sp), cb = 2.f * (1.f - sp) * sp, cc = sp * sp;
                o.goal.x = ca * p0.x + cb * f1.x + cc * f2.x;
                o.goal.y = ca * p0.y + cb * f1.y + cc * f2.y;
                o.goal.z = ca * p0.z + cb * f1.z + cc * f2.z;
            }
        }
        o.next = next;
        if (i == 0) {
            st.scn_i[env] = make_int4(mode, period, next, 0);
            if (mode == QS_SCENARIO_O_EP_RAND_BEZIER && new_segment) {
                st.scn_f[e3] = f0; st.scn_f[e3 + 1] = f1; st.scn_f[e3 + 2] = f2;
            }
        }
    }
    __syncwarp();
    return o;
}

}  // namespace qs
