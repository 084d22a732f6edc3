// Device-side building blocks of the QuadSwarm env step (sm_100a, fp32).
//
// Each function states the reference behaviour it reproduces (file:line under
// gym_art/quadrotor_multi/ of Zhehui-Huang/quad-swarm-rl); the float64 restatement the parity tests
// compare against is oracle/quadswarm_oracle.py.  This is a from-scratch data-parallel design:
// one thread owns one drone, the drones of an env sit in adjacent lanes of one warp, and every
// all-pairs quantity is exchanged with warp shuffles (N <= 32), never through global memory.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/quadswarm.h"
#include "qs_rng.cuh"

namespace qs {

// ---- Crazyflie constants (oracle.QuadParams; tests/golden/crazyflie_constants.json) ----
constexpr float GRAV = 9.81f;
constexpr float MASS = 0.028000000000000008f;
constexpr float INV_MASS = (float)(1.0 / 0.028000000000000008);
constexpr float IXX = 1.3669232142857143e-05f, IYY = 1.4356732142857143e-05f, IZZ = 2.656158333333334e-05f;
constexpr float INV_IXX = (float)(1.0 / 1.3669232142857143e-05), INV_IYY = (float)(1.0 / 1.4356732142857143e-05),
                INV_IZZ = (float)(1.0 / 2.656158333333334e-05);
constexpr float THRUST_MAX = 0.13047300000000003f;
constexpr float TORQUE_MAX = 0.0007828380000000002f;
constexpr float PROP_ARM_XY = 0.0325f;            // |prop_crossproducts| components, quadrotor_dynamics.py:141
constexpr float ARM = 0.04596194077712559f;       // quadrotor_dynamics.py:158 (also the njit floor threshold, :378)
constexpr float MOTOR_TAU_UP = 0.13333244445037035f, MOTOR_TAU_DOWN = 0.13333244445037035f;
constexpr float OMEGA_MAX = 40.0f;
constexpr float FLOOR_MU = 0.6f;
constexpr float SIM_DT = 0.005f;                  // quadrotor_single.py:157
constexpr float CONTROL_DT = 0.01f;               // quadrotor_multi.py:83
constexpr int SIM_STEPS = 2;                      // quadrotor_single.py:102
constexpr int SVD_PERIOD = 100;                   // sub-steps between re-orthogonalisations: the float64 accumulator
                                                  // of quadrotor_dynamics.py:547-551 first exceeds 0.5 on its 100th += 0.005
constexpr float OU_THETA = 0.15f, OU_SIGMA = 0.01f;   // float32 in the reference too (numba_utils.py:69-71)
constexpr float POS_NOISE_STD = 0.005f, VEL_NOISE_STD = 0.01f, GYRO_NOISE_STD = 0.000175f;   // sensor_noise.py:70-76
constexpr float EPS_DYN = 1e-6f;                  // quadrotor_dynamics.py:13
constexpr float EPS_COL = 1e-5f;                  // quad_utils.py:10
constexpr float PI_F = 3.14159265358979323846f;
constexpr float VXYZ_MAX = 3.0f;                  // quadrotor_dynamics.py:50 (neighbour rel-vel clip = 2x, quadrotor_single.py:295)

// ---- state slots: struct-of-arrays of float4, slot-major [NUM_SLOTS][A_pad] -> one LDG.128 per slot per thread ----
enum Slot {
    SL_POS_VX = 0,   // pos.xyz, vel.x
    SL_V_OM,         // vel.y, vel.z, omega.x, omega.y
    SL_OM_R0,        // omega.z, R00, R01, R02
    SL_R1_R20,       // R10, R11, R12, R20
    SL_R2_FLAGS,     // R21, R22, flags(u32), prev_collision_row(u32)
    SL_ROT_DAMP,     // thrust_rot_damp[4]
    SL_CMDS_DAMP,    // thrust_cmds_damp[4]
    SL_OU,           // OU noise state[4]
    SL_DIST_RING,    // previous 4 goal distances, newest first
    NUM_RW_SLOTS,    // ---- slots above are read+written every step ----
    SL_GOAL = NUM_RW_SLOTS,   // goal.xyz, 0   (read every step, written by reset / set_goals)
    SL_DIST_SUMS,    // running sums of the goal distance over the last 1 s / 3 s / 5 s of the episode
    SL_STALE_VEL,    // QuadrotorEnvMulti.self.vel as a reset following a reset would see it (Appendix D-6)
    NUM_SLOTS
};

struct DevState {
    float4* slots;          // [NUM_SLOTS][A_pad]
    long long a_pad;
    int4* env_ctr;          // [E] tick, step_count, svd_count, episode_idx
    int32_t* env_cnt;       // [E][QS_NUM_ENV_STATS] running episode counters
    float2* obst;           // [E][M]
    float4* next_goal;      // [A]
    float4* next_spawn;     // [A]  (w = 1: use it, w = 0: spawn at the goal)
    float2* next_obst;      // [E][M]
    int32_t* stats_env;     // [E][QS_NUM_ENV_STATS]  latched at episode end
    float4* stats_agent;    // [A]                     latched at episode end
};

struct StepParams {
    DevState st;
    const float4* actions;  // [T][A]
    float* obs;             // [T][A][D] or [A][D]
    float* rewards;         // [T][A]
    uint8_t* dones;         // [T][A]
    float* rew_terms;       // [T][A][QS_NUM_TERMS] or null
    const uint8_t* env_mask;    // reset / set_state kernels only
    int E, N, K, D, S, M;
    int obs_repr, use_obst, use_downwash, sense_noise;
    int ep_len, T, last_obs_only;
    float room_lo[3], room_hi[3];
    float col_thr, falloff_thr, obst_radius, obst_col_thr, obst_half_size;
    float grace_steps, final_steps, approach_metric;
    float rew[QS_NUM_REW_COEFF];
    uint32_t seed_lo, seed_hi;
    int env_id_offset;
};

struct Agent {
    float pos[3], vel[3], R[9], om[3];
    float rd[4], cd[4], ou[4];
    float goal[3];
    float ring[4];
    uint32_t flags, prev_col;
};

// ---- small helpers ----
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ float norm3(float x, float y, float z) { return sqrtf(x * x + y * y + z * z); }

template <int NP>
__device__ __forceinline__ float shfl(float v, int src) { return __shfl_sync(0xffffffffu, v, src, NP); }
template <int NP>
__device__ __forceinline__ uint32_t shfl_u(uint32_t v, int src) { return __shfl_sync(0xffffffffu, v, src, NP); }
template <int NP>
__device__ __forceinline__ uint32_t group_ballot(bool pred) {
    const uint32_t b = __ballot_sync(0xffffffffu, pred);
    if (NP == 32) return b;
    const int base = (threadIdx.x & 31) & ~(NP - 1);
    return (b >> base) & ((1u << (NP & 31)) - 1u);
}

__device__ __forceinline__ void load_agent(const DevState& st, long long a, Agent& s) {
    const float4* p = st.slots + a;
    const float4 q0 = p[SL_POS_VX * st.a_pad], q1 = p[SL_V_OM * st.a_pad], q2 = p[SL_OM_R0 * st.a_pad],
                 q3 = p[SL_R1_R20 * st.a_pad], q4 = p[SL_R2_FLAGS * st.a_pad], q5 = p[SL_ROT_DAMP * st.a_pad],
                 q6 = p[SL_CMDS_DAMP * st.a_pad], q7 = p[SL_OU * st.a_pad], q8 = p[SL_DIST_RING * st.a_pad],
                 q9 = p[SL_GOAL * st.a_pad];
    s.pos[0] = q0.x; s.pos[1] = q0.y; s.pos[2] = q0.z; s.vel[0] = q0.w;
    s.vel[1] = q1.x; s.vel[2] = q1.y; s.om[0] = q1.z; s.om[1] = q1.w;
    s.om[2] = q2.x; s.R[0] = q2.y; s.R[1] = q2.z; s.R[2] = q2.w;
    s.R[3] = q3.x; s.R[4] = q3.y; s.R[5] = q3.z; s.R[6] = q3.w;
    s.R[7] = q4.x; s.R[8] = q4.y; s.flags = __float_as_uint(q4.z); s.prev_col = __float_as_uint(q4.w);
    s.rd[0] = q5.x; s.rd[1] = q5.y; s.rd[2] = q5.z; s.rd[3] = q5.w;
    s.cd[0] = q6.x; s.cd[1] = q6.y; s.cd[2] = q6.z; s.cd[3] = q6.w;
    s.ou[0] = q7.x; s.ou[1] = q7.y; s.ou[2] = q7.z; s.ou[3] = q7.w;
    s.ring[0] = q8.x; s.ring[1] = q8.y; s.ring[2] = q8.z; s.ring[3] = q8.w;
    s.goal[0] = q9.x; s.goal[1] = q9.y; s.goal[2] = q9.z;
}

__device__ __forceinline__ void store_agent(const DevState& st, long long a, const Agent& s, bool store_goal) {
    float4* p = st.slots + a;
    p[SL_POS_VX * st.a_pad] = make_float4(s.pos[0], s.pos[1], s.pos[2], s.vel[0]);
    p[SL_V_OM * st.a_pad] = make_float4(s.vel[1], s.vel[2], s.om[0], s.om[1]);
    p[SL_OM_R0 * st.a_pad] = make_float4(s.om[2], s.R[0], s.R[1], s.R[2]);
    p[SL_R1_R20 * st.a_pad] = make_float4(s.R[3], s.R[4], s.R[5], s.R[6]);
    p[SL_R2_FLAGS * st.a_pad] = make_float4(s.R[7], s.R[8], __uint_as_float(s.flags), __uint_as_float(s.prev_col));
    p[SL_ROT_DAMP * st.a_pad] = make_float4(s.rd[0], s.rd[1], s.rd[2], s.rd[3]);
    p[SL_CMDS_DAMP * st.a_pad] = make_float4(s.cd[0], s.cd[1], s.cd[2], s.cd[3]);
    p[SL_OU * st.a_pad] = make_float4(s.ou[0], s.ou[1], s.ou[2], s.ou[3]);
    p[SL_DIST_RING * st.a_pad] = make_float4(s.ring[0], s.ring[1], s.ring[2], s.ring[3]);
    if (store_goal) p[SL_GOAL * st.a_pad] = make_float4(s.goal[0], s.goal[1], s.goal[2], 0.f);
}

// R -> pure yaw (quadrotor_dynamics.py:579-581, :614-621)
__device__ __forceinline__ void yaw_only(float R[9]) {
    const float theta = atan2f(R[3], R[0] + EPS_DYN);
    float s, c;
    sincosf(theta, &s, &c);
    R[0] = c; R[1] = -s; R[2] = 0.f; R[3] = s; R[4] = c; R[5] = 0.f; R[6] = 0.f; R[7] = 0.f; R[8] = 1.f;
}

__device__ __forceinline__ void set_yaw(float R[9], float theta) {
    float s, c;
    sincosf(theta, &s, &c);
    R[0] = c; R[1] = -s; R[2] = 0.f; R[3] = s; R[4] = c; R[5] = 0.f; R[6] = 0.f; R[7] = 0.f; R[8] = 1.f;
}

// Nearest orthogonal matrix (polar factor) = U V^T of the SVD the reference takes every 0.5 s
// (quadrotor_dynamics.py:547-551).  R is orthogonal to rounding on entry, so two Newton steps
// X <- (X + X^-T) / 2 reach fp32 precision; no LAPACK-style SVD is needed on the device.
__device__ __forceinline__ void orthonormalize(float R[9]) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const float c00 = R[4] * R[8] - R[5] * R[7], c01 = R[5] * R[6] - R[3] * R[8], c02 = R[3] * R[7] - R[4] * R[6];
        const float c10 = R[2] * R[7] - R[1] * R[8], c11 = R[0] * R[8] - R[2] * R[6], c12 = R[1] * R[6] - R[0] * R[7];
        const float c20 = R[1] * R[5] - R[2] * R[4], c21 = R[2] * R[3] - R[0] * R[5], c22 = R[0] * R[4] - R[1] * R[3];
        const float det = R[0] * c00 + R[1] * c01 + R[2] * c02;
        const float h = 0.5f / det;     // X^-T = cofactor / det
        R[0] = 0.5f * R[0] + h * c00; R[1] = 0.5f * R[1] + h * c01; R[2] = 0.5f * R[2] + h * c02;
        R[3] = 0.5f * R[3] + h * c10; R[4] = 0.5f * R[4] + h * c11; R[5] = 0.5f * R[5] + h * c12;
        R[6] = 0.5f * R[6] + h * c20; R[7] = 0.5f * R[7] + h * c21; R[8] = 0.5f * R[8] + h * c22;
    }
}

// One 5 ms physics sub-step of the njit path: step1_numba, quadrotor_dynamics.py:348-383
// (calculate_torque_integrate_rotations_and_update_omega :497-566, room clip :360-367,
//  floor_interaction_numba :569-639, compute_velocity_and_acceleration :642-649).
// `cmd` is the RawControl output in [0,1]; the OU state s.ou is the thrust noise of this control step.
__device__ __forceinline__ void dynamics_substep(Agent& s, const float cmd[4], bool do_svd, const StepParams& p,
                                                 const RngKey& key, int i, int sub) {
    // motor lag on sqrt(thrust) + OU thrust noise (:504-517); linearity = 1 so thrust = thrust_max * cmds_damp
    float thr[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const float c = clampf(cmd[m], 0.f, 1.f);
        float tau = (c < s.cd[m]) ? MOTOR_TAU_DOWN : MOTOR_TAU_UP;
        tau = fminf(tau, 1.f);
        s.rd[m] = tau * (sqrtf(c) - s.rd[m]) + s.rd[m];
        s.cd[m] = clampf(s.rd[m] * s.rd[m] + c * s.ou[m], 0.f, 1.f);
        thr[m] = THRUST_MAX * s.cd[m];
    }
    // torques: prop_crossproducts x thrust, plus rotor reaction torque about z (:520-526); rotor drag is zero (:529)
    const float tq0 = PROP_ARM_XY * ((thr[2] + thr[3]) - (thr[0] + thr[1]));
    const float tq1 = PROP_ARM_XY * ((thr[1] + thr[2]) - (thr[0] + thr[3]));
    const float tq2 = TORQUE_MAX * ((s.cd[1] + s.cd[3]) - (s.cd[0] + s.cd[2]));
    const float thrust_z = (thr[0] + thr[1]) + (thr[2] + thr[3]);

    // Rodrigues rotation by the world-frame angular velocity (:537-544)
    {
        const float wx = s.R[0] * s.om[0] + s.R[1] * s.om[1] + s.R[2] * s.om[2];
        const float wy = s.R[3] * s.om[0] + s.R[4] * s.om[1] + s.R[5] * s.om[2];
        const float wz = s.R[6] * s.om[0] + s.R[7] * s.om[1] + s.R[8] * s.om[2];
        const float wn = norm3(wx, wy, wz);
        if (wn != 0.f) {
            const float inv = 1.f / wn;
            const float kx = wx * inv, ky = wy * inv, kz = wz * inv;
            float sn, cs, sh, ch;
            const float ang = wn * SIM_DT;
            sincosf(ang, &sn, &cs);
            sincosf(0.5f * ang, &sh, &ch);
            const float omc = 2.f * sh * sh;          // 1 - cos(ang) without cancellation
            (void)cs; (void)ch;
#pragma unroll
            for (int c = 0; c < 3; ++c) {             // column c of R: v + sin (k x v) + (1-cos) (k (k.v) - v)
                const float vx = s.R[c], vy = s.R[3 + c], vz = s.R[6 + c];
                const float kd = kx * vx + ky * vy + kz * vz;
                const float cx = ky * vz - kz * vy, cy = kz * vx - kx * vz, cz = kx * vy - ky * vx;
                s.R[c] = vx + sn * cx + omc * (kx * kd - vx);
                s.R[3 + c] = vy + sn * cy + omc * (ky * kd - vy);
                s.R[6 + c] = vz + sn * cz + omc * (kz * kd - vz);
            }
        }
    }
    if (do_svd) orthonormalize(s.R);

    // Euler step of the body rates with gyroscopic term (:555-560); quadratic damping is zero
    {
        const float ox = s.om[0], oy = s.om[1], oz = s.om[2];
        const float ix = IXX * ox, iy = IYY * oy, iz = IZZ * oz;
        const float cxx = (-oy) * iz - (-oz) * iy;
        const float cyy = (-oz) * ix - (-ox) * iz;
        const float czz = (-ox) * iy - (-oy) * ix;
        s.om[0] = clampf(ox + SIM_DT * (INV_IXX * (cxx + tq0)), -OMEGA_MAX, OMEGA_MAX);
        s.om[1] = clampf(oy + SIM_DT * (INV_IYY * (cyy + tq1)), -OMEGA_MAX, OMEGA_MAX);
        s.om[2] = clampf(oz + SIM_DT * (INV_IZZ * (czz + tq2)), -OMEGA_MAX, OMEGA_MAX);
    }
    // position with the OLD velocity (:563), room clip and wall / ceiling flags (:360-367)
    const float px = s.pos[0] + SIM_DT * s.vel[0], py = s.pos[1] + SIM_DT * s.vel[1], pz = s.pos[2] + SIM_DT * s.vel[2];
    s.pos[0] = clampf(px, p.room_lo[0], p.room_hi[0]);
    s.pos[1] = clampf(py, p.room_lo[1], p.room_hi[1]);
    s.pos[2] = clampf(pz, p.room_lo[2], p.room_hi[2]);
    uint32_t fl = s.flags & ~(QS_FLAG_CRASHED_WALL | QS_FLAG_CRASHED_CEILING | QS_FLAG_CRASHED_FLOOR);
    if (px != s.pos[0] || py != s.pos[1]) fl |= QS_FLAG_CRASHED_WALL;
    if (pz > s.pos[2]) fl |= QS_FLAG_CRASHED_CEILING;

    // floor contact / friction, threshold = arm (:569-639)
    float acc[3];
    if (s.pos[2] <= ARM) {
        s.pos[2] = ARM;
        float fx = s.R[2] * thrust_z, fy = s.R[5] * thrust_z;
        const float fz = s.R[8] * thrust_z;
        if (fl & QS_FLAG_ON_FLOOR) {
            yaw_only(s.R);
            const float fric = FLOOR_MU * (MASS * GRAV - fz);
            if (norm3(s.vel[0], s.vel[1], s.vel[2]) < EPS_DYN) {
                const float mag = fmaxf(sqrtf(fx * fx + fy * fy) - fric, 0.f);
                if (mag == 0.f) {
                    fx = 0.f; fy = 0.f;
                } else {
                    float sa, ca;
                    sincosf(atan2f(fy, fx), &sa, &ca);
                    fx = mag * ca; fy = mag * sa;
                }
            } else {
                float sa, ca;
                sincosf(atan2f(s.vel[1], s.vel[0]), &sa, &ca);
                fx -= ca * fric; fy -= sa * fric;
            }
        } else {
            fl |= QS_FLAG_ON_FLOOR | QS_FLAG_CRASHED_FLOOR;
            s.vel[0] = s.vel[1] = s.vel[2] = 0.f;
            s.om[0] = s.om[1] = s.om[2] = 0.f;
            if (s.R[8] < 0.f) {                         // upside down: random yaw (:616-619)
                const float4 u = rng_uniform4(key, SITE_FLOOR_YAW, i, 0, 0);
                const float uu = sub == 0 ? u.x : u.y;
                set_yaw(s.R, -PI_F + (PI_F - (-PI_F)) * uu);
            } else {
                yaw_only(s.R);
            }
#pragma unroll
            for (int m = 0; m < 4; ++m) { s.rd[m] = 0.f; s.cd[m] = 0.f; }
        }
        acc[0] = INV_MASS * fx; acc[1] = INV_MASS * fy; acc[2] = fmaxf(0.f, -GRAV + INV_MASS * fz);
    } else {
        fl &= ~QS_FLAG_ON_FLOOR;
        acc[0] = INV_MASS * (s.R[2] * thrust_z);
        acc[1] = INV_MASS * (s.R[5] * thrust_z);
        acc[2] = -GRAV + INV_MASS * (s.R[8] * thrust_z);
    }
    s.flags = fl;
    // velocity with the NEW acceleration (:645); vel_damp = 0.  The accelerometer reading (:648) is never observed.
    s.vel[0] += SIM_DT * acc[0]; s.vel[1] += SIM_DT * acc[1]; s.vel[2] += SIM_DT * acc[2];
}

// rot2quat -> quat2R round trip of the observed rotation (sensor_noise.py:34-63,205-210; quad_utils.py:133-138);
// the rotation-noise quaternion is the identity for the default noise set.
__device__ __forceinline__ void observed_rotation(const float R[9], float out[9]) {
    const float trace = R[0] + R[4] + R[8];
    float qw, qx, qy, qz;
    if (trace > 0.f) {
        const float S = sqrtf(trace + 1.0f) * 2.f;
        qw = 0.25f * S; qx = (R[7] - R[5]) / S; qy = (R[2] - R[6]) / S; qz = (R[3] - R[1]) / S;
    } else if (R[0] > R[4] && R[0] > R[8]) {
        const float S = sqrtf(1.0f + R[0] - R[4] - R[8]) * 2.f;
        qw = (R[7] - R[5]) / S; qx = 0.25f * S; qy = (R[1] + R[3]) / S; qz = (R[2] + R[6]) / S;
    } else if (R[4] > R[8]) {
        const float S = sqrtf(1.0f + R[4] - R[0] - R[8]) * 2.f;
        qw = (R[2] - R[6]) / S; qx = (R[1] + R[3]) / S; qy = 0.25f * S; qz = (R[5] + R[7]) / S;
    } else {
        const float S = sqrtf(1.0f + R[8] - R[0] - R[4]) * 2.f;
        qw = (R[3] - R[1]) / S; qx = (R[2] + R[6]) / S; qy = (R[5] + R[7]) / S; qz = 0.25f * S;
    }
    out[0] = 1.0f - 2.f * qy * qy - 2.f * qz * qz; out[1] = 2.f * qx * qy - 2.f * qz * qw; out[2] = 2.f * qx * qz + 2.f * qy * qw;
    out[3] = 2.f * qx * qy + 2.f * qz * qw; out[4] = 1.0f - 2.f * qx * qx - 2.f * qz * qz; out[5] = 2.f * qy * qz - 2.f * qx * qw;
    out[6] = 2.f * qx * qz - 2.f * qy * qw; out[7] = 2.f * qy * qz + 2.f * qx * qw; out[8] = 1.0f - 2.f * qx * qx - 2.f * qy * qy;
}

// compute_new_vel, collisions/utils.py:8-18
__device__ __forceinline__ void compute_new_vel(float u, float max_vel_magn, float vel[3], const float shift[3],
                                                float low, float high) {
    const float decay = low + (high - low) * u;
    const float nx = vel[0] + shift[0], ny = vel[1] + shift[1], nz = vel[2] + shift[2];
    float mag = norm3(nx, ny, nz);
    const float den = (mag == 0.f) ? mag + EPS_COL : mag;
    const float dx = nx / den, dy = ny / den, dz = nz / den;
    mag = fminf(mag * decay, max_vel_magn);
    const float vx = dx * mag, vy = dy * mag, vz = dz * mag;
    vel[0] += vx - vel[0]; vel[1] += vy - vel[1]; vel[2] += vz - vel[2];
}

// compute_new_omega, collisions/utils.py:22-33 (u3 = direction uniforms, um = magnitude uniform)
__device__ __forceinline__ void compute_new_omega(float u0, float u1, float u2, float um, float magn_scale, float out[3]) {
    const float omega_max = magn_scale * PI_F;
    const float x = -1.f + 2.f * u0, y = -1.f + 2.f * u1, z = -1.f + 2.f * u2;
    const float mag = norm3(x, y, z);
    const float den = (mag == 0.f) ? mag + EPS_COL : mag;
    const float lo = omega_max * 0.5f;
    const float m = lo + (omega_max - lo) * um;
    out[0] = x / den * m; out[1] = y / den * m; out[2] = z / den * m;
}

// perform_collision_between_drones, collisions/quadrotors.py:24-59.  Every lane of the env evaluates the pair
// (a < b) from shuffled copies of both drones and keyed draws, lanes a and b keep their half of the result.
__device__ __forceinline__ void pair_response(const RngKey& key, int a, int b, const float p1[3], float v1[3],
                                              const float p2[3], float v2[3], float domega[3]) {
    float nx = p1[0] - p2[0], ny = p1[1] - p2[1], nz = p1[2] - p2[2];
    const float nm = norm3(nx, ny, nz);
    const float den = (nm == 0.f) ? nm + EPS_COL : nm;
    nx /= den; ny /= den; nz /= den;
    const float v1n = v1[0] * nx + v1[1] * ny + v1[2] * nz;
    const float v2n = v2[0] * nx + v2[1] * ny + v2[2] * nz;
    const float ch[3] = {(v2n - v1n) * nx, (v2n - v1n) * ny, (v2n - v1n) * nz};
    float s1[3] = {ch[0], ch[1], ch[2]}, s2[3] = {-ch[0], -ch[1], -ch[2]};
    for (int t = 0; t < 3; ++t) {
        const float4 n0 = rng_normal4(key, SITE_PAIR_N, a, b, 3 * t), n1 = rng_normal4(key, SITE_PAIR_N, a, b, 3 * t + 1),
                     n2 = rng_normal4(key, SITE_PAIR_N, a, b, 3 * t + 2);
        const float cons[3] = {0.8f * n0.x, 0.8f * n0.y, 0.8f * n0.z};
        const float e1[3] = {n0.w, n1.x, n1.y}, e2[3] = {n1.z, n1.w, n2.x};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            s1[k] = ch[k] + (cons[k] + 0.15f * e1[k]);
            s2[k] = -ch[k] + (-cons[k] + 0.15f * e2[k]);
        }
        const float d1 = (v1[0] + s1[0]) * nx + (v1[1] + s1[1]) * ny + (v1[2] + s1[2]) * nz;
        const float d2 = (v2[0] + s2[0]) * nx + (v2[1] + s2[1]) * ny + (v2[2] + s2[2]) * nz;
        if (d1 > 0.f && 0.f > d2) break;
    }
    const float maxv = fmaxf(norm3(v1[0], v1[1], v1[2]), norm3(v2[0], v2[1], v2[2]));
    const float4 u0 = rng_uniform4(key, SITE_PAIR_U, a, b, 0), u1 = rng_uniform4(key, SITE_PAIR_U, a, b, 1);
    compute_new_vel(u0.x, maxv, v1, s1, 0.2f, 0.8f);
    compute_new_vel(u0.y, maxv, v2, s2, 0.2f, 0.8f);
    compute_new_omega(u0.z, u0.w, u1.x, u1.y, 20.0f, domega);
}

// perform_collision_with_obstacle, collisions/obstacles.py:23-50 (obstacle z = room_height / 2, quadrotor_multi.py:322)
__device__ __forceinline__ void obstacle_response(const RngKey& key, int i, Agent& s, float ox, float oy, float oz,
                                                  float obst_half_size) {
    float nx = s.pos[0] - ox, ny = s.pos[1] - oy;
    const float nm = sqrtf(nx * nx + ny * ny);
    const float den = (nm == 0.f) ? nm + EPS_COL : nm;
    nx /= den; ny /= den;
    const float vmag = norm3(s.vel[0], s.vel[1], s.vel[2]);
    const float nv[3] = {vmag * nx, vmag * ny, 0.f};
    float noise[3] = {0.f, 0.f, 0.f};
    for (int t = 0; t < 3; ++t) {
        const float4 n0 = rng_normal4(key, SITE_OBST_N, i, 0, 2 * t), n1 = rng_normal4(key, SITE_OBST_N, i, 0, 2 * t + 1);
        const float tmp[3] = {0.1f * n0.x + 0.05f * n0.w, 0.1f * n0.y + 0.05f * n1.x, 0.1f * n0.z + 0.05f * n1.y};
        if ((nv[0] + tmp[0]) * nx + (nv[1] + tmp[1]) * ny > 0.f) {
            noise[0] = tmp[0]; noise[1] = tmp[1]; noise[2] = tmp[2];
            break;
        }
    }
    const float4 u0 = rng_uniform4(key, SITE_OBST_U, i, 0, 0), u1 = rng_uniform4(key, SITE_OBST_U, i, 0, 1);
    const float shift[3] = {nv[0] - s.vel[0] + noise[0], nv[1] - s.vel[1] + noise[1], nv[2] - s.vel[2] + noise[2]};
    const bool inside = norm3(s.pos[0] - ox, s.pos[1] - oy, s.pos[2] - oz) < obst_half_size;
    compute_new_vel(u0.x, vmag, s.vel, shift, inside ? 1.0f : 0.2f, inside ? 1.0f : 0.8f);
    float dw[3];
    compute_new_omega(u0.y, u0.z, u0.w, u1.x, 1.0f, dw);
    s.om[0] += dw[0]; s.om[1] += dw[1]; s.om[2] += dw[2];
}

__device__ __forceinline__ void room_omega_kick(float u0, float u1, float u2, float um, Agent& s) {
    const float omega_max = 20.f * PI_F;
    float x = -1.f + 2.f * u0, y = -1.f + 2.f * u1, z = -1.f + 2.f * u2;
    const float inv = 1.f / (norm3(x, y, z) + 1e-5f);
    const float lo = omega_max * 0.5f;
    const float m = lo + (omega_max - lo) * um;
    s.om[0] += x * inv * m; s.om[1] += y * inv * m; s.om[2] += z * inv * m;
}

// perform_collision_with_wall, collisions/room.py:6-44
__device__ __forceinline__ void wall_response(const RngKey& key, int i, Agent& s, const StepParams& p) {
    const float4 u0 = rng_uniform4(key, SITE_WALL_U, i, 0, 0), u1 = rng_uniform4(key, SITE_WALL_U, i, 0, 1),
                 u2 = rng_uniform4(key, SITE_WALL_U, i, 0, 2);
    const float speed = norm3(s.vel[0], s.vel[1], s.vel[2]);
    const float lo = 0.2f * speed, hi = 0.8f * speed;
    const float real_speed = clampf(lo + (hi - lo) * u0.x, 0.1f, 6.0f);
    float dx = -1.f + 2.f * u0.y, dy = -1.f + 2.f * u0.z;
    if (s.pos[0] == p.room_lo[0]) dx = 0.1f + (1.0f - 0.1f) * u1.x;
    else if (s.pos[0] == p.room_hi[0]) dx = -1.0f + (-0.1f - -1.0f) * u1.x;
    if (s.pos[1] == p.room_lo[1]) dy = 0.1f + (1.0f - 0.1f) * u1.y;
    else if (s.pos[1] == p.room_hi[1]) dy = -1.0f + (-0.1f - -1.0f) * u1.y;
    const float dz = -1.0f + (-0.5f - -1.0f) * u1.z;
    const float inv = 1.f / (norm3(dx, dy, dz) + 1e-5f);
    s.vel[0] = real_speed * (dx * inv); s.vel[1] = real_speed * (dy * inv); s.vel[2] = real_speed * (dz * inv);
    room_omega_kick(u1.w, u2.x, u2.y, u2.z, s);
}

// perform_collision_with_ceiling, collisions/room.py:91-113
__device__ __forceinline__ void ceiling_response(const RngKey& key, int i, Agent& s) {
    const float4 u0 = rng_uniform4(key, SITE_CEIL_U, i, 0, 0), u1 = rng_uniform4(key, SITE_CEIL_U, i, 0, 1),
                 u2 = rng_uniform4(key, SITE_CEIL_U, i, 0, 2);
    const float speed = norm3(s.vel[0], s.vel[1], s.vel[2]);
    const float lo = 0.2f * speed, hi = 0.8f * speed;
    const float real_speed = clampf(lo + (hi - lo) * u0.x, 0.1f, 6.0f);
    const float dx = -1.f + 2.f * u0.y, dy = -1.f + 2.f * u0.z;
    const float dz = -1.0f + (-0.5f - -1.0f) * u1.x;
    const float inv = 1.f / (norm3(dx, dy, dz) + 1e-5f);
    s.vel[0] = real_speed * (dx * inv); s.vel[1] = real_speed * (dy * inv); s.vel[2] = real_speed * (dz * inv);
    room_omega_kick(u1.y, u1.z, u1.w, u2.x, s);
}

// QuadrotorSingle._reset, quadrotor_single.py:387-447: spawn jitter, z >= 0.75, random yaw facing the origin,
// zero rates and motor state, cleared contact flags.  OU state and the SVD counter are NOT reset (Appendix D-9).
__device__ __forceinline__ void reset_agent(Agent& s, const RngKey& key, int i, const float spawn[3], float box) {
    const float4 u = rng_uniform4(key, SITE_SPAWN_U, i, 0, 0);
    s.pos[0] = (-box + (box - (-box)) * u.x) + spawn[0];
    s.pos[1] = (-box + (box - (-box)) * u.y) + spawn[1];
    s.pos[2] = fmaxf((-box + (box - (-box)) * u.z) + spawn[2], 0.75f);
    // to_xyhat(-pos), quad_utils.py:75-82,112-116
    float hx = -s.pos[0], hy = -s.pos[1];
    const float n = sqrtf(hx * hx + hy * hy);
    if (!(n < 0.00001f)) { hx /= n; hy /= n; }
    float sn = 0.f, cs = 1.f;
    for (int k = 0; k < RESET_YAW_MAX_TRIES; ++k) {
        const float4 uy = rng_uniform4(key, SITE_RESET_YAW_U, i, 0, k >> 2);
        const float uu = (k & 3) == 0 ? uy.x : (k & 3) == 1 ? uy.y : (k & 3) == 2 ? uy.z : uy.w;
        sincosf(-PI_F + (PI_F - (-PI_F)) * uu, &sn, &cs);
        if (cs * hx + sn * hy >= 0.5f) break;          // rotation[:, 0] = (cos, sin, 0)
    }
    s.R[0] = cs; s.R[1] = -sn; s.R[2] = 0.f; s.R[3] = sn; s.R[4] = cs; s.R[5] = 0.f; s.R[6] = 0.f; s.R[7] = 0.f; s.R[8] = 1.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) { s.vel[k] = 0.f; s.om[k] = 0.f; }
#pragma unroll
    for (int m = 0; m < 4; ++m) { s.rd[m] = 0.f; s.cd[m] = 0.f; s.ring[m] = 0.f; }
    s.flags = QS_FLAG_NO_COL_AGENT | QS_FLAG_NO_COL_OBST;
    s.prev_col = 0u;
}

}  // namespace qs
