// The fused env-step kernel: per-drone dynamics -> all-pairs collision / proximity / downwash ->
// contact responses -> observation assembly -> episode bookkeeping and auto-reset, ONE launch per
// control step (or T control steps per launch for qs_rollout, state kept in registers).
//
// Thread mapping: NP = next power of two >= N lanes per env, 32/NP envs per warp; lane i of a group owns
// drone i.  All cross-drone traffic is __shfl_sync within the group; the env's pillar table is staged in
// shared memory.  Follows QuadrotorEnvMulti.step, quadrotor_multi.py:413-724 (see DESIGN.md for the map).
#pragma once
#include "qs_device.cuh"

namespace qs {

struct EnvCtr {
    int tick, step_count, svd_count, episode_idx;
};

// Observation row of one drone: get_state.py:6-72 (self part), quadrotor_multi.py:233-274 (neighbours),
// obstacles/utils.py:5-27 (3x3 SDF).  `nvel` is the velocity the neighbour block sees (stale after a reset,
// SURVEY Appendix D-6); `site` picks the sensor-noise draw.
template <int NP>
__device__ __forceinline__ void write_observation(const StepParams& p, const RngKey& key, const Agent& s, const float nvel[3],
                                                  int i, bool valid, uint32_t site, const float2* s_obst_env,
                                                  float* __restrict__ row) {
    // ---- self observation
    float o[24];
    {
        float np_[3] = {0.f, 0.f, 0.f}, nv[3] = {0.f, 0.f, 0.f}, nw[3] = {0.f, 0.f, 0.f};
        if (p.sense_noise) {
            const float4 a = rng_normal4(key, site, i, 0, 0), b = rng_normal4(key, site, i, 0, 1),
                         c = rng_normal4(key, site, i, 0, 2);
            np_[0] = POS_NOISE_STD * a.x; np_[1] = POS_NOISE_STD * a.y; np_[2] = POS_NOISE_STD * a.z;
            nv[0] = VEL_NOISE_STD * a.w; nv[1] = VEL_NOISE_STD * b.x; nv[2] = VEL_NOISE_STD * b.y;
            nw[0] = GYRO_NOISE_STD * b.z; nw[1] = GYRO_NOISE_STD * b.w; nw[2] = GYRO_NOISE_STD * c.x;
        }
        const float px = s.pos[0] + np_[0], py = s.pos[1] + np_[1], pz = s.pos[2] + np_[2];
        o[0] = px - s.goal[0]; o[1] = py - s.goal[1]; o[2] = pz - s.goal[2];
        o[3] = s.vel[0] + nv[0]; o[4] = s.vel[1] + nv[1]; o[5] = s.vel[2] + nv[2];
        if (p.sense_noise) {
            observed_rotation(s.R, o + 6);
        } else {
#pragma unroll
            for (int k = 0; k < 9; ++k) o[6 + k] = s.R[k];
        }
        o[15] = s.om[0] + nw[0]; o[16] = s.om[1] + nw[1]; o[17] = s.om[2] + nw[2];
        if (p.obs_repr == QS_OBS_XYZ_VXYZ_R_OMEGA_FLOOR) {
            o[18] = pz;
        } else if (p.obs_repr == QS_OBS_XYZ_VXYZ_R_OMEGA_WALL) {
            o[18] = clampf(px - p.room_lo[0], 0.f, 5.f); o[19] = clampf(py - p.room_lo[1], 0.f, 5.f);
            o[20] = clampf(pz - p.room_lo[2], 0.f, 5.f);
            o[21] = clampf(p.room_hi[0] - px, 0.f, 5.f); o[22] = clampf(p.room_hi[1] - py, 0.f, 5.f);
            o[23] = clampf(p.room_hi[2] - pz, 0.f, 5.f);
        }
    }
    if (valid) {
#pragma unroll
        for (int k = 0; k < 24; ++k)
            if (k < p.S) row[k] = o[k];
    }

    // ---- neighbour block: K nearest by distance + closing speed, or all others in index order
    if (NP > 1 && p.K > 0) {
        const float rx = p.room_hi[0] - p.room_lo[0], ry = p.room_hi[1] - p.room_lo[1], rz = p.room_hi[2] - p.room_lo[2];
        const float rv = 2.0f * VXYZ_MAX;
        float* nrow = row + p.S;
        if (p.K == p.N - 1) {
            int slot = 0;
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                const float qx = shfl<NP>(s.pos[0], j), qy = shfl<NP>(s.pos[1], j), qz = shfl<NP>(s.pos[2], j);
                const float wx = shfl<NP>(nvel[0], j), wy = shfl<NP>(nvel[1], j), wz = shfl<NP>(nvel[2], j);
                if (j < p.N && j != i && valid) {
                    float* d = nrow + 6 * slot;
                    d[0] = clampf(qx - s.pos[0], -rx, rx); d[1] = clampf(qy - s.pos[1], -ry, ry);
                    d[2] = clampf(qz - s.pos[2], -rz, rz);
                    d[3] = clampf(wx - nvel[0], -rv, rv); d[4] = clampf(wy - nvel[1], -rv, rv);
                    d[5] = clampf(wz - nvel[2], -rv, rv);
                    ++slot;
                }
            }
        } else {
            float score[NP];
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                const float dx = shfl<NP>(s.pos[0], j) - s.pos[0], dy = shfl<NP>(s.pos[1], j) - s.pos[1],
                            dz = shfl<NP>(s.pos[2], j) - s.pos[2];
                const float ux = shfl<NP>(nvel[0], j) - nvel[0], uy = shfl<NP>(nvel[1], j) - nvel[1],
                            uz = shfl<NP>(nvel[2], j) - nvel[2];
                const float dist = fmaxf(norm3(dx, dy, dz), 0.01f);
                const float sc = dist + ((dx / dist) * ux + (dy / dist) * uy + (dz / dist) * uz);
                score[j] = (j < p.N && j != i) ? sc : __int_as_float(0x7f800000);   // +inf: never selected
            }
            uint32_t taken = 0u;
            for (int k = 0; k < p.K; ++k) {
                float best = __int_as_float(0x7f800000);
                int bj = -1;
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    const bool ok = !((taken >> j) & 1u) && j < p.N && j != i;
                    // stable argsort: strictly smaller score wins, ties keep the lower index (Appendix D-11)
                    if (ok && (bj < 0 || score[j] < best)) { best = score[j]; bj = j; }
                }
                const int src = bj < 0 ? i : bj;
                taken |= 1u << src;
                const float qx = shfl<NP>(s.pos[0], src), qy = shfl<NP>(s.pos[1], src), qz = shfl<NP>(s.pos[2], src);
                const float wx = shfl<NP>(nvel[0], src), wy = shfl<NP>(nvel[1], src), wz = shfl<NP>(nvel[2], src);
                if (valid) {
                    float* d = nrow + 6 * k;
                    d[0] = clampf(qx - s.pos[0], -rx, rx); d[1] = clampf(qy - s.pos[1], -ry, ry);
                    d[2] = clampf(qz - s.pos[2], -rz, rz);
                    d[3] = clampf(wx - nvel[0], -rv, rv); d[4] = clampf(wy - nvel[1], -rv, rv);
                    d[5] = clampf(wz - nvel[2], -rv, rv);
                }
            }
        }
    }

    // ---- 3x3 signed-distance patch around the drone (resolution 0.1 m)
    if (p.use_obst) {
        const float res = 0.1f;
        const float gx[3] = {s.pos[0] - res, s.pos[0], s.pos[0] + res};
        const float gy[3] = {s.pos[1] - res, s.pos[1], s.pos[1] + res};
        float best[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) best[c] = 100.0f * 100.0f;
        for (int m = 0; m < p.M; ++m) {
            const float2 ob = s_obst_env[m];
            float ex[3], ey[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float dx = gx[c] - ob.x, dy = gy[c] - ob.y;
                ex[c] = dx * dx; ey[c] = dy * dy;
            }
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) best[3 * a + b] = fminf(best[3 * a + b], ex[a] + ey[b]);
        }
        if (valid) {
            float* srow = row + p.S + 6 * p.K;
#pragma unroll
            for (int c = 0; c < 9; ++c) srow[c] = sqrtf(best[c]) - p.obst_radius;
        }
    }
}

// Copy the next-episode tables into the live episode of one env and respawn its drones
// (QuadrotorEnvMulti.reset, quadrotor_multi.py:339-411).  Returns the velocity the neighbour block must see.
template <int NP>
__device__ __forceinline__ void reset_env(const StepParams& p, const RngKey& key, Agent& s, long long a, int env, int i,
                                          bool do_reset, bool valid, int tick_before_reset, float2* s_obst_env,
                                          float nvel[3]) {
    // Called by ALL lanes of a warp (the branch around it is warp-uniform); `do_reset` is per env.
    const DevState& st = p.st;
    if (do_reset && valid) {
        // stale velocity (Appendix D-6): the multi-env's self.vel is only refreshed by step()
        if (tick_before_reset > 0) {
            nvel[0] = s.vel[0]; nvel[1] = s.vel[1]; nvel[2] = s.vel[2];
        } else {
            const float4 sv = st.slots[SL_STALE_VEL * st.a_pad + a];
            nvel[0] = sv.x; nvel[1] = sv.y; nvel[2] = sv.z;
        }
        st.slots[SL_STALE_VEL * st.a_pad + a] = make_float4(nvel[0], nvel[1], nvel[2], 0.f);
        const float4 g = st.next_goal[a], sp = st.next_spawn[a];
        s.goal[0] = g.x; s.goal[1] = g.y; s.goal[2] = g.z;
        const float spawn[3] = {sp.w != 0.f ? sp.x : g.x, sp.w != 0.f ? sp.y : g.y, sp.w != 0.f ? sp.z : g.z};
        reset_agent(s, key, i, spawn, p.use_obst ? 0.1f : 2.0f);      // box: quadrotor_single.py:215-218
        st.slots[SL_DIST_SUMS * st.a_pad + a] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (p.use_obst) {
        if (do_reset) {
            for (int m = i; m < p.M; m += NP) {
                const float2 ob = st.next_obst[(long long)env * p.M + m];
                st.obst[(long long)env * p.M + m] = ob;
                s_obst_env[m] = ob;
            }
        }
        __syncwarp();
    }
}

template <int NP>
__global__ void __launch_bounds__(128) qs_step_kernel(const __grid_constant__ StepParams p) {
    extern __shared__ float2 s_obst[];
    const DevState& st = p.st;
    const int lane = threadIdx.x & 31;
    const int i = lane & (NP - 1);
    const int envs_per_block = blockDim.x / NP;
    const int env_local = threadIdx.x / NP;
    const int env = blockIdx.x * envs_per_block + env_local;
    const bool env_ok = env < p.E;
    const bool valid = env_ok && i < p.N;
    const long long a = (long long)env * p.N + i;
    const long long A = (long long)p.E * p.N;

    // stage this block's pillar tables (contiguous [envs_per_block][M] float2) in shared memory
    if (p.use_obst) {
        const long long base = (long long)blockIdx.x * envs_per_block * p.M;
        const long long total = (long long)p.E * p.M;
        for (int k = threadIdx.x; k < envs_per_block * p.M; k += blockDim.x)
            if (base + k < total) s_obst[k] = st.obst[base + k];
        __syncthreads();
    }
    float2* s_obst_env = s_obst + env_local * p.M;

    Agent s;
    EnvCtr ctr = {0, 0, 0, 0};
    if (valid) load_agent(st, a, s);
    else {
#pragma unroll
        for (int k = 0; k < 3; ++k) { s.pos[k] = 1e9f + 1e6f * i; s.vel[k] = 0.f; s.om[k] = 0.f; s.goal[k] = 0.f; }
#pragma unroll
        for (int k = 0; k < 9; ++k) s.R[k] = (k % 4 == 0) ? 1.f : 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) { s.rd[k] = 0.f; s.cd[k] = 0.f; s.ou[k] = 0.f; s.ring[k] = 0.f; }
        s.flags = 0u; s.prev_col = 0u;
    }
    if (env_ok) {
        const int4 c = st.env_ctr[env];
        ctr.tick = c.x; ctr.step_count = c.y; ctr.svd_count = c.z; ctr.episode_idx = c.w;
    }
    bool goal_dirty = false;

    for (int t = 0; t < p.T; ++t) {
        RngKey key;
        key.k0 = p.seed_lo; key.k1 = p.seed_hi;
        key.env = (uint32_t)(p.env_id_offset + env);
        key.step = (uint32_t)ctr.step_count;

        // ================= per-drone part: QuadrotorSingle._step, quadrotor_single.py:341-357 =================
        float act[4] = {0.f, 0.f, 0.f, 0.f};
        if (valid) {
            const float4 av = p.actions[(long long)t * A + a];
            act[0] = av.x; act[1] = av.y; act[2] = av.z; act[3] = av.w;
        }
        float cmd[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) cmd[m] = 0.5f * (clampf(act[m], -1.f, 1.f) + 1.f);    // RawControl.step, quadrotor_control.py:53-57
        {   // OU thrust noise, once per control step (numba_utils.py:101-105, quadrotor_dynamics.py:209)
            const float4 z = rng_normal4(key, SITE_OU, i, 0, 0);
            s.ou[0] += OU_THETA * (0.f - s.ou[0]) + OU_SIGMA * z.x;
            s.ou[1] += OU_THETA * (0.f - s.ou[1]) + OU_SIGMA * z.y;
            s.ou[2] += OU_THETA * (0.f - s.ou[2]) + OU_SIGMA * z.z;
            s.ou[3] += OU_THETA * (0.f - s.ou[3]) + OU_SIGMA * z.w;
        }
#pragma unroll
        for (int sub = 0; sub < SIM_STEPS; ++sub) {
            ctr.svd_count += 1;
            const bool do_svd = ctr.svd_count >= SVD_PERIOD;
            if (do_svd) ctr.svd_count = 0;
            dynamics_substep(s, cmd, do_svd, p, key, i, sub);
        }
        // compute_reward_weighted, quadrotor_single.py:34-92 (dt = SIM dt, raw unclipped action)
        const bool on_floor = (s.flags & QS_FLAG_ON_FLOOR) != 0u;
        const float dist = norm3(s.goal[0] - s.pos[0], s.goal[1] - s.pos[1], s.goal[2] - s.pos[2]);
        const float raw_effort = sqrtf(act[0] * act[0] + act[1] * act[1] + act[2] * act[2] + act[3] * act[3]);
        const float raw_orient = on_floor ? 1.0f : -s.R[8];
        const float raw_spin = sqrtf(s.om[0] * s.om[0] + s.om[1] * s.om[1] + s.om[2] * s.om[2]);
        const float raw_crash = on_floor ? 1.0f : 0.0f;
        float reward = -SIM_DT * (p.rew[QS_REW_POS] * dist + p.rew[QS_REW_EFFORT] * raw_effort + p.rew[QS_REW_CRASH] * raw_crash +
                                  p.rew[QS_REW_ORIENT] * raw_orient + p.rew[QS_REW_SPIN] * raw_spin);
        const int tick_before = ctr.tick;
        const int time_remain = p.ep_len - tick_before;
        ctr.tick = tick_before + 1;
        const bool done = ctr.tick > p.ep_len;

        // ================= env part: all-pairs pass (positions only) =================
        // calculate_collision_matrix (collisions/quadrotors.py:63-91), proximity penalties (:95-103), downwash
        // detection (aerodynamics/downwash.py:4-51) — drone i scans every other drone j of its env.
        uint32_t cur_col = 0u;
        float prox = 0.f;
        bool dw_applied = false;
        float dw_dv[3] = {0.f, 0.f, 0.f}, dw_dw[3] = {0.f, 0.f, 0.f};
        if (NP > 1) {
            const float pen_ratio = -p.rew[QS_REW_QUADCOL_BIN_SMOOTH_MAX] / p.falloff_thr;
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                const float qx = shfl<NP>(s.pos[0], j), qy = shfl<NP>(s.pos[1], j), qz = shfl<NP>(s.pos[2], j);
                float zx = 0.f, zy = 0.f, zz = 0.f;
                if (p.use_downwash) { zx = shfl<NP>(s.R[2], j); zy = shfl<NP>(s.R[5], j); zz = shfl<NP>(s.R[8], j); }
                if (j < p.N && j != i && valid) {
                    // same expression order as the reference for i<j: (p_lo - p_hi)^2 is symmetric
                    const float dx = s.pos[0] - qx, dy = s.pos[1] - qy, dz = s.pos[2] - qz;
                    const float d = sqrtf(dx * dx + dy * dy + dz * dz);
                    if (d <= p.col_thr) cur_col |= 1u << j;
                    if (d <= p.falloff_thr) prox += pen_ratio * d + p.rew[QS_REW_QUADCOL_BIN_SMOOTH_MAX];
                    if (p.use_downwash) {
                        // is drone i (me) inside the downwash cylinder below drone j?
                        const float rel_z = dx * zx + dy * zy + dz * zz;
                        const float rel_xy = sqrtf(d * d - rel_z * rel_z);         // NaN -> comparison false, as in numpy
                        if (-0.7f < rel_z && rel_z < 0.f && rel_xy < 0.1f) {
                            const float4 ui = rng_uniform4(key, SITE_DW_I, j, 0, 0);
                            const float4 u0 = rng_uniform4(key, SITE_DW_IJ, j, i, 0), u1 = rng_uniform4(key, SITE_DW_IJ, j, i, 1);
                            const float acc = fmaxf(1e-6f, (6.f / 17.f) * (-10.f * d + 7.f) + (-0.1f + 0.2f * ui.x));
                            const float omd = fmaxf(1e-6f, 0.3f * (d - 1.f) * (d - 1.f) + (-0.01f + 0.02f * ui.y));
                            float ax = zx + (-0.1f + 0.2f * u0.x), ay = zy + (-0.1f + 0.2f * u0.y), az = zz + (-0.1f + 0.2f * u0.z);
                            float mag = norm3(ax, ay, az);
                            float den = (mag == 0.f) ? mag + 1e-6f : mag;
                            ax = -(ax / den); ay = -(ay / den); az = -(az / den);
                            float bx = -1.f + 2.f * u0.w, by = -1.f + 2.f * u1.x, bz = -1.f + 2.f * u1.y;
                            mag = norm3(bx, by, bz);
                            den = (mag == 0.f) ? mag + 1e-6f : mag;
                            dw_dv[0] += acc * ax * CONTROL_DT; dw_dv[1] += acc * ay * CONTROL_DT; dw_dv[2] += acc * az * CONTROL_DT;
                            dw_dw[0] += omd * (bx / den) * CONTROL_DT; dw_dw[1] += omd * (by / den) * CONTROL_DT;
                            dw_dw[2] += omd * (bz / den) * CONTROL_DT;
                            dw_applied = true;
                        }
                    }
                }
            }
        }
        // collision bookkeeping, quadrotor_multi.py:433-459 (quirks D-2..D-4 reproduced)
        const bool in_u = (cur_col != 0u) && (s.prev_col == 0u);                     // flattened-id set difference
        const uint32_t u_mask = group_ballot<NP>(in_u && valid);
        const int col_curr_tick = __popc(u_mask) / 2;
        const bool u_any = (u_mask & ~1u) != 0u;                                      // ids.any(): id 0 alone is falsy
        const float raw_quadcol = (u_any && in_u) ? -1.0f : 0.0f;
        uint32_t new_pairs = cur_col & ~s.prev_col;                                   // pair-level novelty (:437-438)
        const bool settled = (float)ctr.tick >= p.grace_steps;
        if (col_curr_tick > 0 && settled && in_u) s.flags &= ~QS_FLAG_NO_COL_AGENT;
        s.prev_col = cur_col;

        // obstacles: first pillar in index order within arm + radius (obstacles/utils.py:31-43), :462-488
        int hit = -1;
        if (p.use_obst) {
            for (int m = p.M - 1; m >= 0; --m) {
                const float2 ob = s_obst_env[m];
                const float dx = s.pos[0] - ob.x, dy = s.pos[1] - ob.y;
                if (sqrtf(dx * dx + dy * dy) <= p.obst_col_thr) hit = m;
            }
        }
        const bool new_obst = (hit >= 0) && !(s.flags & QS_FLAG_PREV_OBST) && valid;
        const uint32_t obst_mask = p.use_obst ? group_ballot<NP>(new_obst) : 0u;
        const float raw_obst = new_obst ? -1.0f : 0.0f;
        s.flags = (hit >= 0) ? (s.flags | QS_FLAG_PREV_OBST) : (s.flags & ~QS_FLAG_PREV_OBST);
        int far35 = 0, far5 = 0;
        if (new_obst && settled) {
            s.flags &= ~QS_FLAG_NO_COL_OBST;
            // distance to goal of the FIRST-draw noisy position (quadrotor_multi.py:474)
            float nx = 0.f, ny = 0.f, nz = 0.f;
            if (p.sense_noise) {
                const float4 n = rng_normal4(key, SITE_SENSOR0, i, 0, 0);
                nx = POS_NOISE_STD * n.x; ny = POS_NOISE_STD * n.y; nz = POS_NOISE_STD * n.z;
            }
            const float q = norm3((s.pos[0] + nx) - s.goal[0], (s.pos[1] + ny) - s.goal[1], (s.pos[2] + nz) - s.goal[2]);
            far35 = q > 3.5f; far5 = q > 5.0f;
        }

        // room, quadrotor_multi.py:289-302,491-497 (quirk D-5: novelty against the previously RETURNED lists)
        const bool floor_c = (s.flags & QS_FLAG_CRASHED_FLOOR) != 0u && valid;
        const bool wall_c = (s.flags & QS_FLAG_CRASHED_WALL) && !(s.flags & QS_FLAG_PREV_WALL) && valid;
        const bool ceil_c = (s.flags & QS_FLAG_CRASHED_CEILING) && !(s.flags & QS_FLAG_PREV_CEILING) && valid;
        const bool room_c = (floor_c || wall_c || ceil_c) && !(s.flags & QS_FLAG_PREV_ROOM);
        s.flags &= ~(QS_FLAG_PREV_WALL | QS_FLAG_PREV_CEILING | QS_FLAG_PREV_ROOM | QS_FLAG_KICKED | QS_FLAG_NEW_QUADCOL | QS_FLAG_NEW_OBSTCOL);
        if (wall_c) s.flags |= QS_FLAG_PREV_WALL;
        if (ceil_c) s.flags |= QS_FLAG_PREV_CEILING;
        if (room_c) s.flags |= QS_FLAG_PREV_ROOM;
        if (u_any && in_u) s.flags |= QS_FLAG_NEW_QUADCOL;
        if (new_obst) s.flags |= QS_FLAG_NEW_OBSTCOL;

        // rewards, quadrotor_multi.py:499-540
        const float rew_prox = -1.0f * (CONTROL_DT * prox);
        reward += p.rew[QS_REW_QUADCOL_BIN] * raw_quadcol;
        reward += rew_prox;
        if (p.use_obst) reward += p.rew[QS_REW_QUADCOL_BIN_OBST] * raw_obst;

        // goal-distance log and reached_goal, quadrotor_multi.py:542-546
        {
            const float m5 = (dist + s.ring[0] + s.ring[1] + s.ring[2] + s.ring[3]) * 0.2f;
            if (ctr.tick >= 5 && m5 < p.approach_metric) s.flags |= QS_FLAG_REACHED_GOAL;
            s.ring[3] = s.ring[2]; s.ring[2] = s.ring[1]; s.ring[1] = s.ring[0]; s.ring[0] = dist;
            const int len = p.ep_len + 1;
            const int w5 = min(len, 500);
            if (valid && ctr.tick > len - w5) {
                float4 sums = st.slots[SL_DIST_SUMS * st.a_pad + a];
                if (ctr.tick > len - min(len, 100)) sums.x += dist;
                if (ctr.tick > len - min(len, 300)) sums.y += dist;
                sums.z += dist;
                st.slots[SL_DIST_SUMS * st.a_pad + a] = sums;
            }
        }

        // episode counters (lane 0 of the env), quadrotor_multi.py:448-456,468-478,522-526
        {
            const uint32_t floor_m = group_ballot<NP>(floor_c), wall_m = group_ballot<NP>(wall_c),
                           ceil_m = group_ballot<NP>(ceil_c), room_m = group_ballot<NP>(room_c);
            const uint32_t f35 = group_ballot<NP>(far35 != 0), f5 = group_ballot<NP>(far5 != 0);
            const int n_obst = __popc(obst_mask);
            const bool any_event = col_curr_tick > 0 || n_obst > 0 || ((floor_m | wall_m | ceil_m | room_m) != 0u && settled);
            if (any_event && i == 0 && env_ok) {
                int32_t* c = st.env_cnt + (long long)env * QS_NUM_ENV_STATS;
                c[QS_STAT_NUM_COLLISIONS] += col_curr_tick;
                if (col_curr_tick > 0 && settled) c[QS_STAT_NUM_COLLISIONS_AFTER_SETTLE] += col_curr_tick;
                if (col_curr_tick > 0 && (float)time_remain <= p.final_steps) c[QS_STAT_NUM_COLLISIONS_FINAL_5S] += col_curr_tick;
                c[QS_STAT_NUM_COLLISIONS_OBST] += n_obst;
                if (settled) {
                    c[QS_STAT_NUM_COLLISIONS_OBST_AFTER_SETTLE] += n_obst;
                    c[QS_STAT_NUM_COLLISIONS_OBST_3_5] += __popc(f35);
                    c[QS_STAT_NUM_COLLISIONS_OBST_5] += __popc(f5);
                    c[QS_STAT_NUM_COLLISIONS_ROOM] += __popc(room_m);
                    c[QS_STAT_NUM_COLLISIONS_FLOOR] += __popc(floor_m);
                    c[QS_STAT_NUM_COLLISIONS_WALL] += __popc(wall_m);
                    c[QS_STAT_NUM_COLLISIONS_CEILING] += __popc(ceil_m);
                }
            }
        }

        // ================= contact responses, quadrotor_multi.py:548-587 =================
        bool kicked = false;
        if (p.use_downwash) {
            if (dw_applied) {
#pragma unroll
                for (int k = 0; k < 3; ++k) { s.vel[k] += dw_dv[k]; s.om[k] += dw_dw[k]; }
            }
            kicked = group_ballot<NP>(dw_applied) != 0u;
        }
        if (NP > 1) {
            // new colliding pairs, lexicographic order, one at a time (the second response of a drone sees the first)
            uint32_t pending = new_pairs & ~((2u << i) - 1u);      // partners j > i: lane i owns pair (i, j)
            if (!valid) pending = 0u;
            while (__any_sync(0xffffffffu, pending != 0u)) {
                // every lane of the warp runs the same shuffles; groups without a pending pair just idle
                const uint32_t owners = group_ballot<NP>(pending != 0u);
                const bool act = owners != 0u;
                const int pa = act ? __ffs(owners) - 1 : 0;
                const uint32_t pend_a = shfl_u<NP>(pending, pa);
                const int pb = act ? __ffs(pend_a) - 1 : 0;
                float p1[3], v1[3], p2[3], v2[3], dwv[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    p1[k] = shfl<NP>(s.pos[k], pa); v1[k] = shfl<NP>(s.vel[k], pa);
                    p2[k] = shfl<NP>(s.pos[k], pb); v2[k] = shfl<NP>(s.vel[k], pb);
                }
                if (act) {
                    pair_response(key, pa, pb, p1, v1, p2, v2, dwv);
                    if (i == pa) {
#pragma unroll
                        for (int k = 0; k < 3; ++k) { s.vel[k] = v1[k]; s.om[k] += dwv[k]; }
                        pending &= ~(1u << pb);
                    } else if (i == pb) {
#pragma unroll
                        for (int k = 0; k < 3; ++k) { s.vel[k] = v2[k]; s.om[k] -= dwv[k]; }
                    }
                    kicked = true;
                }
            }
        }
        if (p.use_obst) {
            if (new_obst) {
                const float2 ob = s_obst_env[hit];
                obstacle_response(key, i, s, ob.x, ob.y, 0.5f * (p.room_hi[2] - p.room_lo[2]) + p.room_lo[2], p.obst_half_size);
            }
            kicked = kicked || obst_mask != 0u;
        }
        {
            if (wall_c) wall_response(key, i, s, p);
            if (ceil_c) ceiling_response(key, i, s);
            // NB: the ballot must not sit behind a short-circuit `||` — `kicked` differs between the envs of a warp
            const bool room_any = group_ballot<NP>(wall_c || ceil_c) != 0u;
            kicked = kicked || room_any;
        }
        if (kicked) s.flags |= QS_FLAG_KICKED;

        // ================= outputs of this step =================
        const long long ta = (long long)t * A + a;
        if (valid) {
            p.rewards[ta] = reward;
            p.dones[ta] = done ? 1 : 0;
            if (p.rew_terms) {
                float* tr = p.rew_terms + ta * QS_NUM_TERMS;
                tr[QS_TERM_RAW_POS] = SIM_DT * -dist;
                tr[QS_TERM_RAW_ACTION] = SIM_DT * -raw_effort;
                tr[QS_TERM_RAW_CRASH] = SIM_DT * -raw_crash;
                tr[QS_TERM_RAW_ORIENT] = SIM_DT * -raw_orient;
                tr[QS_TERM_RAW_SPIN] = SIM_DT * -raw_spin;
                tr[QS_TERM_RAW_QUADCOL] = raw_quadcol;
                tr[QS_TERM_PROXIMITY] = rew_prox;
                tr[QS_TERM_RAW_QUADCOL_OBST] = raw_obst;
            }
        }

        // ================= episode end: latch statistics, auto-reset (quadrotor_multi.py:626-722) =================
        float nvel[3] = {s.vel[0], s.vel[1], s.vel[2]};
        uint32_t site = kicked ? SITE_SENSOR1 : SITE_SENSOR0;
        const bool do_reset = done && env_ok;
        if (__any_sync(0xffffffffu, do_reset)) {          // warp-uniform branch
            if (do_reset && valid) {
                const float4 sums = st.slots[SL_DIST_SUMS * st.a_pad + a];
                const int len = p.ep_len + 1;
                const uint32_t fbits = ((s.flags & QS_FLAG_NO_COL_AGENT) ? 1u : 0u) | ((s.flags & QS_FLAG_NO_COL_OBST) ? 2u : 0u) |
                                       ((s.flags & QS_FLAG_REACHED_GOAL) ? 4u : 0u);
                st.stats_agent[a] = make_float4(sums.x / (float)min(len, 100), sums.y / (float)min(len, 300),
                                                sums.z / (float)min(len, 500), __uint_as_float(fbits));
            }
            if (do_reset && i == 0) {
                int32_t* c = st.env_cnt + (long long)env * QS_NUM_ENV_STATS;
                int32_t* o = st.stats_env + (long long)env * QS_NUM_ENV_STATS;
                for (int k = 0; k < QS_NUM_ENV_STATS; ++k) { o[k] = c[k]; c[k] = 0; }
                o[QS_STAT_EPISODES_DONE] = ctr.episode_idx + 1;
            }
            reset_env<NP>(p, key, s, a, env, i, do_reset, valid, ctr.tick, s_obst_env, nvel);
            if (do_reset) {
                ctr.tick = 0;
                ctr.episode_idx += 1;
                goal_dirty = true;
                site = SITE_SENSOR_RESET;
            }
        }

        // ================= observation (of the post-response, or freshly reset, state) =================
        if (!p.last_obs_only || t == p.T - 1) {
            float* row = p.obs + ((p.last_obs_only ? 0 : (long long)t * A) + a) * p.D;
            write_observation<NP>(p, key, s, nvel, i, valid, site, s_obst_env, row);
        }
        ctr.step_count += 1;
    }

    if (valid) store_agent(st, a, s, goal_dirty);
    if (env_ok && i == 0) st.env_ctr[env] = make_int4(ctr.tick, ctr.step_count, ctr.svd_count, ctr.episode_idx);
}

// Explicit reset of the masked envs: QuadrotorEnvMulti.reset, quadrotor_multi.py:339-411.
template <int NP>
__global__ void __launch_bounds__(128) qs_reset_kernel(const __grid_constant__ StepParams p) {
    extern __shared__ float2 s_obst[];
    const DevState& st = p.st;
    const int lane = threadIdx.x & 31;
    const int i = lane & (NP - 1);
    const int envs_per_block = blockDim.x / NP;
    const int env_local = threadIdx.x / NP;
    const int env = blockIdx.x * envs_per_block + env_local;
    const bool env_ok = env < p.E && (p.env_mask == nullptr || p.env_mask[env] != 0);
    const bool valid = env_ok && i < p.N;
    const long long a = (long long)env * p.N + i;
    float2* s_obst_env = s_obst + env_local * p.M;

    Agent s;
    if (valid) load_agent(st, a, s);
    else {
#pragma unroll
        for (int k = 0; k < 3; ++k) { s.pos[k] = 1e9f + 1e6f * i; s.vel[k] = 0.f; s.om[k] = 0.f; s.goal[k] = 0.f; }
#pragma unroll
        for (int k = 0; k < 9; ++k) s.R[k] = (k % 4 == 0) ? 1.f : 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) { s.rd[k] = 0.f; s.cd[k] = 0.f; s.ou[k] = 0.f; s.ring[k] = 0.f; }
        s.flags = 0u; s.prev_col = 0u;
    }
    EnvCtr ctr = {0, 0, 0, 0};
    if (env_ok) {
        const int4 c = st.env_ctr[env];
        ctr.tick = c.x; ctr.step_count = c.y; ctr.svd_count = c.z; ctr.episode_idx = c.w;
    }
    RngKey key;
    key.k0 = p.seed_lo; key.k1 = p.seed_hi;
    key.env = (uint32_t)(p.env_id_offset + env);
    key.step = (uint32_t)ctr.step_count;
    float nvel[3] = {0.f, 0.f, 0.f};
    // every lane of the warp takes part in the shuffles below; lanes of unmasked envs write nothing
    reset_env<NP>(p, key, s, a, env, i, env_ok, valid, ctr.tick, s_obst_env, nvel);
    if (env_ok && i == 0) {
        int32_t* c = st.env_cnt + (long long)env * QS_NUM_ENV_STATS;
        for (int k = 0; k < QS_NUM_ENV_STATS; ++k) c[k] = 0;
        st.env_ctr[env] = make_int4(0, ctr.step_count + 1, ctr.svd_count, ctr.episode_idx);
    }
    write_observation<NP>(p, key, s, nvel, i, valid, SITE_SENSOR_RESET, s_obst_env, p.obs + a * p.D);
    if (valid) store_agent(st, a, s, true);
}

}  // namespace qs
