// C ABI of the B200-native QuadSwarm env step (see include/quadswarm.h for the contract and the
// reference interfaces each entry point replaces).  Build: nvcc -gencode arch=compute_100a,code=sm_100a.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <cuda.h>          // CUtensorMap + enums only; the encoder is fetched with cudaGetDriverEntryPoint (no -lcuda)

#include "qs_step.cuh"
#include "qs_wrap.cuh"

using namespace qs;

// ------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------
struct QsHandle {
    QsConfig cfg;
    int device;
    int NP;               // lanes per env (next pow2 >= N)
    int K, S, D, M, ep_len;
    long long A, a_pad;
    DevState st;
    float rew[QS_NUM_REW_COEFF];
    int64_t launches;
    int split_mode;       // -1 auto, 0 single-warp kernel, 1 split kernel (QS_SPLIT, read at qs_create)
    int pdl_env;          // QS_PDL at the first step launch (-2 = not read yet, -1 = unset)
    int handover;         // -1 not decided yet, 0 grid-wide wait between step grids, 1 per-block hand-over (launch_step)
    int obst_random, n_obst_counts, n_obst_radii;
    int obst_counts[QS_MAX_OBST_CHOICES];
    float obst_radii[QS_MAX_OBST_CHOICES];
    bool wrap_on;
    WrapState wrap;
    float* wrap_agg_host; // pinned
    int pregen_every;     // step launches between two launches of the next-episode generator (0 = never), QS_PREGEN overrides
    int since_pregen;
    unsigned long long capture_id;   // id of the stream capture that saw the last generator launch at its head
    int chained;          // qs_set_chained: consecutive qs_step / qs_rollout launches follow each other directly on the stream
    int last_was_step;    // the last launch this handle enqueued was a step / rollout grid
    int last_was_wrap;    // ... was the wrapper kernel of a wrapped control step launched block-chained (qs_wrap_step)
    int in_wrap_step;     // launch_step is called from qs_wrap_step
    int step_wrap_chain;  // the step grid just launched leaves its blocks to the wrapper kernel (StepParams.wrap_chain)
    int wrap_block;       // worker threads per block of that step grid (the wrapper kernel uses the same env -> block mapping)
    int bulk_mode;        // QS_OBS_BULK: -1 auto, 0 never use the bulk-copy engine for the observation write-out, 2 linear copies only
    int* err_host;        // mapped page-locked word the step kernels set when a hand-over wait timed out (sticky)
    cudaEvent_t ev_sync;  // the *_host entry points (own stream) order themselves after the caller-stream work below
#ifdef QS_TIMELINE
    unsigned long long* tl;
    int tl_next;
#endif
    cudaStream_t last_stream;   // stream of the most recent asynchronous call of this handle
    bool async_pending;
    // staging for the *_host entry points (pinned host + device mirrors)
    float *d_actions, *d_obs, *d_rewards, *d_terms;
    uint8_t *d_dones, *d_mask;
    float *h_actions, *h_obs, *h_rewards, *h_terms;
    uint8_t *h_dones, *h_mask;
    cudaStream_t own_stream;
};

static thread_local std::string g_err;

static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define QS_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return fail(QS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));          \
    } while (0)

extern "C" const char* qs_last_error(void) { return g_err.c_str(); }

static inline int blocks_for(long long n) { return (int)((n + 255) / 256); }

static int next_pow2(int n) {
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}

static void fill_params(const QsHandle* h, StepParams& p) {
    memset(&p, 0, sizeof(p));
    const QsConfig& c = h->cfg;
    p.st = h->st;
    p.E = c.num_envs; p.N = c.num_agents; p.K = h->K; p.D = h->D; p.S = h->S; p.M = h->M;
    p.obs_repr = c.obs_repr; p.use_obst = c.use_obstacles ? 1 : 0; p.use_downwash = c.use_downwash ? 1 : 0;
    p.sense_noise = c.sense_noise ? 1 : 0;
    p.ep_len = h->ep_len; p.T = 1; p.last_obs_only = 0;
    // room_box, quadrotor_single.py:146-147
    p.room_lo[0] = -c.room_dims[0] / 2.f; p.room_lo[1] = -c.room_dims[1] / 2.f; p.room_lo[2] = 0.f;
    p.room_hi[0] = c.room_dims[0] / 2.f; p.room_hi[1] = c.room_dims[1] / 2.f; p.room_hi[2] = c.room_dims[2];
    const double arm = c.quad_arm > 0.f ? (double)c.quad_arm : 0.04596194077712559;      // quadrotor_multi.py:81
    p.col_thr = (float)(c.collision_hitbox_radius * arm);        // quadrotor_multi.py:154
    p.falloff_thr = (float)(c.collision_falloff_radius * arm);   // quadrotor_multi.py:155
    p.obst_radius = (float)(c.obst_size / 2.0);
    p.obst_col_thr = (float)(arm + c.obst_size / 2.0);           // obstacles/utils.py:33
    p.obst_half_size = (float)(c.obst_size / 2.0);
    p.grace_steps = 150.f;                                       // 1.5 * control_freq, quadrotor_multi.py:146
    p.final_steps = 500.f;                                       // 5.0 * control_freq, quadrotor_multi.py:150
    p.approach_metric = c.approch_goal_metric;
    for (int k = 0; k < QS_NUM_REW_COEFF; ++k) p.rew[k] = h->rew[k];
    p.seed_lo = (uint32_t)(c.seed & 0xffffffffull);
    p.seed_hi = (uint32_t)(c.seed >> 32);
    p.env_id_offset = c.env_id_offset;
    // observation staging tile: only when a warp's tile fits comfortably in shared memory
    const int D = h->D;
    const int V = (D % 4 == 0) ? 4 : ((D % 2 == 0) ? 2 : 1);
    const int Q = D / V;
    p.obs_v = V; p.obs_q = Q;
    p.obs_dp = (Q % 2 == 0) ? D + V : D;                 // odd number of V-wide words per row: fewer bank conflicts
    p.obs_magic = ((1 << 20) + Q - 1) / Q;
    p.obs_stage = (D <= 72) ? 1 : 0;
    p.obs_bulk = 0;
    p.chained = 0;
    p.obst_random = h->obst_random; p.n_obst_counts = h->n_obst_counts; p.n_obst_radii = h->n_obst_radii;
    for (int k = 0; k < QS_MAX_OBST_CHOICES; ++k) { p.obst_counts[k] = h->obst_counts[k]; p.obst_radii[k] = h->obst_radii[k]; }
    p.scenario = c.scenario; p.grid_l = c.obst_grid[0]; p.grid_w = c.obst_grid[1];
}

// Observation write-out mode of a step launch (qs_step.cuh, emit_observation_tile): the bulk-copy engine needs a 16-byte
// aligned destination in device memory; host-mapped (zero-copy) destinations keep the vector-store loop.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
        else
            cudaGetLastError();
    }
    return fn;
}

static void choose_obs_writeout(const QsHandle* h, StepParams& p, bool dst_is_device_memory) {
    static_assert(sizeof(CUtensorMap) == sizeof(p.obs_map), "tensor map size");
    if (!p.obs_stage || !dst_is_device_memory || h->bulk_mode == 0) return;
    if (((uintptr_t)p.obs & 15u) != 0) return;
    if (p.D % 4 == 0 && h->bulk_mode != 2) {
        // tensor map of the caller's observation array: [T][A][D] floats, box [1][rows of a warp tile][Dp] — the box is as
        // wide as the padded shared-memory row; the columns >= D (and rows >= A of a ragged last tile) are clipped
        EncodeTiledFn enc = tensor_map_encoder();
        if (!enc) return;
        const cuuint64_t T = p.last_obs_only ? 1 : (cuuint64_t)p.T;
        const cuuint64_t A = (cuuint64_t)p.E * p.N;
        const cuuint64_t gdim[3] = {(cuuint64_t)p.D, A, T};
        const cuuint64_t gstr[2] = {(cuuint64_t)p.D * 4, A * (cuuint64_t)p.D * 4};
        const cuuint32_t box[3] = {(cuuint32_t)p.obs_dp, (cuuint32_t)((32 / h->NP) * p.N), 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        if (enc((CUtensorMap*)p.obs_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)p.obs, gdim, gstr, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return;
        p.obs_bulk = 1;
    } else {
        p.obs_bulk = 2;                   // one linear copy per warp tile: rows staged back to back
        p.obs_dp = p.D;
    }
}

// Every asynchronous call remembers its stream; the host-buffer entry points, which run on the handle's own
// non-blocking stream, order themselves after it (join_caller_stream) so that e.g. a qs_set_goals / qs_set_state issued on
// the caller's stream is complete before qs_step_host reads the state.  Nothing is recorded on the hot path.
static void note_async(QsHandle* h, cudaStream_t s, bool is_step) {
    h->last_was_step = is_step ? 1 : 0;
    h->last_was_wrap = 0;
    if (s != h->own_stream) { h->last_stream = s; h->async_pending = true; }
}
static void join_caller_stream(QsHandle* h) {
    if (!h->async_pending) return;
    h->async_pending = false;
    if (cudaEventRecord(h->ev_sync, h->last_stream) == cudaSuccess && cudaStreamWaitEvent(h->own_stream, h->ev_sync, 0) == cudaSuccess) return;
    cudaGetLastError();                  // stream gone or being captured: fall back to a full device synchronisation
    cudaDeviceSynchronize();
}

// ------------------------------------------------------------------------------------------
// small kernels: tables, goals, state import / export
// ------------------------------------------------------------------------------------------
__global__ void k_set_next_episode(DevState st, int E, int N, int M, const uint8_t* mask, const float* goals,
                                   const float* spawn, const float* obst) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long A = (long long)E * N;
    if (t < A) {
        const int env = (int)(t / N);
        if (mask == nullptr || mask[env]) {
            st.next_goal[t] = make_float4(goals[3 * t], goals[3 * t + 1], goals[3 * t + 2], 0.f);
            st.next_spawn[t] = spawn ? make_float4(spawn[3 * t], spawn[3 * t + 1], spawn[3 * t + 2], 1.f)
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    if (obst != nullptr && t < (long long)E * M) {
        const int env = (int)(t / M);
        if (mask == nullptr || mask[env]) st.next_obst[t] = make_float2(obst[2 * t], obst[2 * t + 1]);
    }
}

__global__ void k_set_goals(DevState st, int E, int N, const uint8_t* mask, const float* goals) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)E * N) return;
    const int env = (int)(t / N);
    if (mask == nullptr || mask[env])
        st.slots[SL_GOAL * st.a_pad + t] = make_float4(goals[3 * t], goals[3 * t + 1], goals[3 * t + 2], 0.f);
}

static_assert(QS_STATE_ENV_I32 >= 4 + QS_NUM_ENV_STATS + 17, "env state row too short");
__global__ void k_get_state(DevState st, int E, int N, int M, float* af, uint32_t* au, int32_t* ei, float* obst) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long A = (long long)E * N;
    if (t < A) {
        Agent s;
        load_agent(st, t, s);
        float* o = af + t * QS_STATE_F32;
        int k = 0;
        for (int c = 0; c < 3; ++c) o[k++] = s.pos[c];
        for (int c = 0; c < 3; ++c) o[k++] = s.vel[c];
        for (int c = 0; c < 9; ++c) o[k++] = s.R[c];
        for (int c = 0; c < 3; ++c) o[k++] = s.om[c];
        for (int c = 0; c < 4; ++c) o[k++] = s.rd[c];
        for (int c = 0; c < 4; ++c) o[k++] = s.cd[c];
        for (int c = 0; c < 4; ++c) o[k++] = s.ou[c];
        for (int c = 0; c < 3; ++c) o[k++] = s.goal[c];
        for (int c = 0; c < 4; ++c) o[k++] = s.ring[c];
        const float4 sums = st.slots[SL_DIST_SUMS * st.a_pad + t], sv = st.slots[SL_STALE_VEL * st.a_pad + t];
        o[k++] = sums.x; o[k++] = sums.y; o[k++] = sums.z;
        o[k++] = sv.x; o[k++] = sv.y; o[k++] = sv.z;
        uint32_t* u = au + t * QS_STATE_U32;
        u[0] = s.flags; u[1] = s.prev_col; u[2] = 0u; u[3] = 0u;
    }
    if (t < E) {
        const int4 c = st.env_ctr[t];
        int32_t* e = ei + t * QS_STATE_ENV_I32;
        e[0] = c.x; e[1] = c.y; e[2] = c.z; e[3] = c.w;
        for (int k = 0; k < QS_NUM_ENV_STATS; ++k) e[4 + k] = st.env_cnt[t * QS_NUM_ENV_STATS + k];
        int32_t* sc = e + 4 + QS_NUM_ENV_STATS;
        const int4 si = st.scn_i[t];
        sc[0] = si.x; sc[1] = si.y; sc[2] = si.z; sc[3] = si.w;
        for (int q = 0; q < 3; ++q) {
            const float4 f = st.scn_f[3 * t + q];
            sc[4 + 4 * q] = __float_as_int(f.x); sc[5 + 4 * q] = __float_as_int(f.y);
            sc[6 + 4 * q] = __float_as_int(f.z); sc[7 + 4 * q] = __float_as_int(f.w);
        }
        e[4 + QS_NUM_ENV_STATS + 16] = st.epi[t].x;                                        // episode number (keys the episode draws)
        for (int k = 4 + QS_NUM_ENV_STATS + 17; k < QS_STATE_ENV_I32; ++k) e[k] = 0;      // reserved
    }
    if (obst != nullptr && t < (long long)E * M) {
        const float2 ob = st.obst[t];
        obst[2 * t] = ob.x; obst[2 * t + 1] = ob.y;
    }
}

__global__ void k_set_state(DevState st, int E, int N, int M, const uint8_t* mask, const float* af, const uint32_t* au,
                            const int32_t* ei, const float* obst) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long A = (long long)E * N;
    if (t < A && (mask == nullptr || mask[t / N])) {
        Agent s;
        const float* o = af + t * QS_STATE_F32;
        int k = 0;
        for (int c = 0; c < 3; ++c) s.pos[c] = o[k++];
        for (int c = 0; c < 3; ++c) s.vel[c] = o[k++];
        for (int c = 0; c < 9; ++c) s.R[c] = o[k++];
        for (int c = 0; c < 3; ++c) s.om[c] = o[k++];
        for (int c = 0; c < 4; ++c) s.rd[c] = o[k++];
        for (int c = 0; c < 4; ++c) s.cd[c] = o[k++];
        for (int c = 0; c < 4; ++c) s.ou[c] = o[k++];
        for (int c = 0; c < 3; ++c) s.goal[c] = o[k++];
        for (int c = 0; c < 4; ++c) s.ring[c] = o[k++];
        st.slots[SL_DIST_SUMS * st.a_pad + t] = make_float4(o[k], o[k + 1], o[k + 2], 0.f);
        k += 3;
        st.slots[SL_STALE_VEL * st.a_pad + t] = make_float4(o[k], o[k + 1], o[k + 2], 0.f);
        const uint32_t* u = au + t * QS_STATE_U32;
        s.flags = u[0]; s.prev_col = u[1];
        store_agent(st, t, s, true);
    }
    if (t < E && (mask == nullptr || mask[t])) {
        const int32_t* e = ei + t * QS_STATE_ENV_I32;
        st.env_ctr[t] = make_int4(e[0], e[1], e[2], e[3]);
        for (int k = 0; k < QS_NUM_ENV_STATS; ++k) st.env_cnt[t * QS_NUM_ENV_STATS + k] = e[4 + k];
        const int32_t* sc = e + 4 + QS_NUM_ENV_STATS;
        st.scn_i[t] = make_int4(sc[0], sc[1], sc[2], sc[3]);
        for (int q = 0; q < 3; ++q)
            st.scn_f[3 * t + q] = make_float4(__int_as_float(sc[4 + 4 * q]), __int_as_float(sc[5 + 4 * q]),
                                              __int_as_float(sc[6 + 4 * q]), __int_as_float(sc[7 + 4 * q]));
        // the pre-generated next-episode record stays: it is a function of (seed, env, episode number) only and is used
        // only if its number still matches (reset_env)
        st.epi[t] = make_int2(e[4 + QS_NUM_ENV_STATS + 16], st.epi[t].y);
    }
    if (obst != nullptr && t < (long long)E * M && (mask == nullptr || mask[t / M]))
        st.obst[t] = make_float2(obst[2 * t], obst[2 * t + 1]);
}

// qs_set_dynamics: rows [A][QS_DYN_ROW] -> the live table (now) or the table latched at the env's next reset
__global__ void k_set_dynamics(DevState st, int E, int N, const uint8_t* mask, const float4* rows, int at_next_reset) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)E * N * (QS_DYN_ROW / 4);
    if (t < total) {
        const int env = (int)(t / ((long long)N * (QS_DYN_ROW / 4)));
        if (mask == nullptr || mask[env]) (at_next_reset ? st.next_dyn : st.dyn)[t] = rows[t];
    }
    if (at_next_reset && t < E && (mask == nullptr || mask[t])) st.dyn_pending[t] = 1;
}

__global__ void k_read_stats(DevState st, int E, int N, int32_t* env_stats, float* agent_stats) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < (long long)E * N && agent_stats) {
        const float4 v = st.stats_agent[t];
        agent_stats[4 * t] = v.x; agent_stats[4 * t + 1] = v.y; agent_stats[4 * t + 2] = v.z;
        agent_stats[4 * t + 3] = (float)__float_as_uint(v.w);
    }
    if (t < (long long)E * QS_NUM_ENV_STATS && env_stats) env_stats[t] = st.stats_env[t];
}

// ------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------
static int block_size() {
    // threads per CTA (multiple of 32, <= 128); QS_BLOCK overrides for tuning experiments
    static int b = 0;
    if (b == 0) {
        const char* e = getenv("QS_BLOCK");
        b = e ? atoi(e) : 64;
        if (b < 32 || b > QS_LB || (b % 32) != 0) b = 64;
    }
    return b;
}

template <typename F>
static int dispatch_np(int NP, F&& f) {
    switch (NP) {
        case 1: return f(std::integral_constant<int, 1>());
        case 2: return f(std::integral_constant<int, 2>());
        case 4: return f(std::integral_constant<int, 4>());
        case 8: return f(std::integral_constant<int, 8>());
        case 16: return f(std::integral_constant<int, 16>());
        case 32: return f(std::integral_constant<int, 32>());
    }
    return fail(QS_ERR_UNSUPPORTED, "num_agents > 32 is not supported by this build");
}

static int launch_pregen(QsHandle* h, cudaStream_t s);

static int launch_step(QsHandle* h, const StepParams& p_in, cudaStream_t s, bool obs_in_device_memory = true) {
    if (h->err_host && *(volatile int*)h->err_host != 0)
        return fail(QS_ERR_CUDA, "a per-block hand-over between step grids timed out earlier: the env state of this handle is "
                                 "not trustworthy any more (qs_handover_timeouts); destroy the handle");
    // split kernel: physics warp + observer warp per 32 drones (QS_SPLIT=0/1 at qs_create overrides the heuristic)
    // Measured (profiles/r01_notes.md): splitting shortens one warp's dependency chain (32 envs: 6.9 -> 5.8 us per
    // launch, 8 x 1024 envs: 8.05 -> 7.17 us) but adds work, so it only pays while the GPU has idle issue slots, i.e.
    // up to about one physics warp per SM sub-partition (4 x 148 on B200).
    if (h->pregen_every > 0) {
        // next-episode records for the envs that consumed theirs (qs_pregen_kernel): every pregen_every step launches.  A
        // captured graph repeats exactly the launches of its capture: a short graph captured between two generator launches
        // and replayed forever never refills a record, and every auto-reset then generates its episode inside the step (same
        // results; ~12 us per reset on c3, which the per-block hand-over mostly hides: 10.0 -> 10.4 us per step).  With
        // QS_PREGEN_HEAD=1 every captured graph starts with a generator launch instead (costs ~10 us per replay).
        bool due = (h->since_pregen += p_in.T) >= h->pregen_every;
        static int head = -1;
        if (head < 0) { const char* e = getenv("QS_PREGEN_HEAD"); head = e ? atoi(e) : 0; }
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        unsigned long long cid = 0;
        if (head && cudaStreamGetCaptureInfo(s, &cs, &cid) == cudaSuccess && cs == cudaStreamCaptureStatusActive && cid != h->capture_id) {
            h->capture_id = cid;
            due = true;
        }
        if (due) {
            int rcp = launch_pregen(h, s);
            if (rcp != QS_OK) return rcp;
        }
    }
    StepParams p = p_in;
    choose_obs_writeout(h, p, obs_in_device_memory);
    const long long phys_warps = ((long long)h->cfg.num_envs * h->NP + 31) / 32;
    // A chained handle whose batch gives every SM at least two warps steps faster in the balanced shape with a courier warp
    // (below) than in the split shape: c2 (1024 envs x 8 drones) 7.03 -> 6.42 us per step.
    static int courier_env = -1;
    if (courier_env < 0) { const char* e = getenv("QS_COURIER"); courier_env = e ? atoi(e) : 1; }
    int sms = 0;
    QS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
    const int wpc_all = (int)((phys_warps + sms - 1) / sms);
    const bool courier_shape = h->chained && courier_env && h->NP < 16 && h->st.dyn == nullptr && wpc_all >= 2 &&
                               (wpc_all + 1) * 32 <= QS_LB && ((wpc_all * 32) % h->NP) == 0;
    const bool want_split = h->split_mode == 1 || (h->split_mode == -1 && phys_warps <= 4 * 148 && !courier_shape);
    const bool split = want_split && p.obs_stage && h->NP > 1 && h->st.dyn == nullptr && !h->obst_random;
    // QS_BALANCE=1 (experiment): one CTA per SM, ceil(warps / SMs) warps each — every SM then holds the same number of warps
    // whatever the CTA scheduler does while two step grids overlap (the timeline of the debug build showed SMs with 6 CTAs
    // of 2 warps next to SMs with 2, and the step ends with the slowest block)
    static int balance = -1;
    if (balance < 0) { const char* e = getenv("QS_BALANCE"); balance = e ? atoi(e) : 1; }
    int kBlock = split ? 64 : block_size();
    if (h->NP >= 16 && kBlock > 128) kBlock = 128;          // launch bounds of the NP >= 16 instantiations
    bool balanced = false;
    if (balance && !split && h->NP < 16) {
        const int wpc = wpc_all;
        // measured: c5 (8 x 4096, K = 6, staggered resets) 13.7 -> 10.0 us per step, c3 unchanged
        if (wpc >= 2 && wpc * 32 <= QS_LB && ((wpc * 32) % h->NP) == 0) { kBlock = wpc * 32; balanced = true; }
    }
    if (h->pdl_env == -2) {          // read once per handle
        const char* e = getenv("QS_PDL");
        const int m = e ? atoi(e) : -1;
        h->pdl_env = (m < -1 || m > 4) ? -1 : m;
    }
    const int pdl_env = h->pdl_env;
    // Balanced single-wave grids use the per-block hand-over with a COURIER warp (qs_step.cuh): one more warp per CTA that
    // carries no envs and does the hand-over's flag traffic — acquire of the predecessor's state word, early release of
    // this block's state (before the observation is built), the `done` word that orders the observation rows of consecutive
    // steps.  Measured (profiles/r02_notes.md): c3 10.0 -> 8.5 us per step, c5 9.3 -> 8.6.  QS_COURIER=0 switches it off.
    if (h->handover < 0 && balanced)          // decided before the first launch so that every grid of a chain has the same shape
        h->handover = pdl_env >= 0 ? (pdl_env == 3) : ((courier_env && kBlock + 32 <= QS_LB) || h->cfg.use_obstacles != 0);
    const bool courier = balanced && courier_env && h->handover == 1 && h->chained && h->st.dyn == nullptr && kBlock + 32 <= QS_LB;
    if (courier) kBlock += 32;
    p.courier = courier ? 1 : 0;
    const int work_warps = kBlock / 32 - (courier ? 1 : 0);
    const int envs_per_block = (split ? 32 : work_warps * 32) / h->NP;
    const int grid = (h->cfg.num_envs + envs_per_block - 1) / envs_per_block;
    size_t smem = h->cfg.use_obstacles ? (size_t)envs_per_block * h->M * sizeof(float2) : 0;
    smem = (smem + 127) / 128 * 128;              // TMA sources are 128-byte aligned
    p.smem_tile_off = (int)(smem / sizeof(float));
    if (p.obs_stage) smem += (size_t)(split ? 1 : work_warps) * 32 * p.obs_dp * sizeof(float);
    if (split) smem += (size_t)HAND_FLOATS * sizeof(float);
    // shared-memory footprint of a balanced CTA.  Without a courier warp: 120 KB, i.e. one CTA per SM and never two CTAs of
    // the same grid on one SM.  With it: 64 KB, so that the successor's CTA (whose block was released early) already runs on
    // the SM while this one writes its observation rows; the register file limits an SM to two such CTAs anyway.
    // Measured (c3 / c5, us per step): 120 KB 9.7 / 9.3, 100 KB 9.1 / 9.3, 70 KB 8.8 / 8.8, 48 KB 8.8 / 8.7.  QS_BALANCE_KB overrides.
    static int balance_kb = -1;
    if (balance_kb < 0) { const char* e = getenv("QS_BALANCE_KB"); balance_kb = e ? atoi(e) : 0; }
    const int pad_kb = balance_kb > 0 ? balance_kb : (courier ? 64 : 120);
    if (balanced && smem < (size_t)pad_kb * 1024) smem = (size_t)pad_kb * 1024;
    // Programmatic dependent launch between consecutive step grids (QS_PDL overrides; default -1 = choose per handle):
    //   0 off; 1 grid-wide wait, trigger at kernel start (measured 2 us slower); 2 grid-wide wait, trigger before the final
    //   stores (0.2-0.4 us faster per step than 0); 3 per-block hand-over, no grid-wide wait (qs_step.cuh).
    // Measured (profiles/r01_notes.md): 3 wins when a step grid needs more than one wave of CTAs (c4: 29.3 -> 20.1 us,
    // 16384 x 8 drones: 22.4 -> 19.8 us) and for the split shape (c2: 7.18 -> 6.99 us); for a single-wave grid whose
    // warps run in lock-step anyway (c3) its acquire / release costs what the hidden launch latency saves, so 2 stays.
    using KernelFn = void (*)(StepParams);
    const bool ticked_obst = p.scenario >= QS_SCENARIO_O_DYNAMIC_SAME_GOAL && p.scenario <= QS_SCENARIO_O_EP_RAND_BEZIER;
    const bool scn = p.use_obst ? ticked_obst
                                : ((p.scenario >= QS_SCENARIO_DEVICE_FAMILY_FIRST && p.scenario <= QS_SCENARIO_MIX) ||
                                   p.scenario == QS_SCENARIO_EP_RAND_BEZIER || p.scenario == QS_SCENARIO_RUN_AWAY);
    KernelFn fn_wait = nullptr, fn_ho = nullptr;
    int rc = dispatch_np(h->NP, [&](auto np) {
        constexpr int NPv = decltype(np)::value;
        if (split) {
            fn_wait = scn ? (KernelFn)qs_step_kernel<NPv, true, true, false> : (KernelFn)qs_step_kernel<NPv, true, false, false>;
            fn_ho = scn ? (KernelFn)qs_step_kernel<NPv, true, true, true> : (KernelFn)qs_step_kernel<NPv, true, false, true>;
        } else {
            fn_wait = scn ? (KernelFn)qs_step_kernel<NPv, false, true, false> : (KernelFn)qs_step_kernel<NPv, false, false, false>;
            fn_ho = scn ? (KernelFn)qs_step_kernel<NPv, false, true, true> : (KernelFn)qs_step_kernel<NPv, false, false, true>;
        }
        return QS_OK;
    });
    if (rc != QS_OK) return rc;
    KernelFn fn_dyn = nullptr;
    if (h->st.dyn != nullptr) {           // per-drone physical constants: single-warp shape, grid-wide wait
        dispatch_np(h->NP, [&](auto np) {
            constexpr int NPv = decltype(np)::value;
            fn_dyn = scn ? (KernelFn)qs_step_kernel<NPv, false, true, false, true> : (KernelFn)qs_step_kernel<NPv, false, false, false, true>;
            return QS_OK;
        });
    }
    if (h->handover < 0) {
        if (pdl_env >= 0) h->handover = pdl_env == 3;
        else {
            int per_sm = 0, sms = 0;
            QS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn_ho, kBlock, smem));
            QS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
            // measured on c3 with envs resetting in different steps (round 2, profiles/r02_*): the reset of an env with a
            // pillar table makes its block ~2 us late; with the grid-wide wait every step pays that (11.2 us), with the
            // hand-over only the block's own chain does (10.2 us).  Lock-step envs: 9.2 vs 10.0 us (QS_PDL=2 selects it).
            h->handover = split || (long long)grid > (long long)per_sm * sms || h->cfg.use_obstacles != 0;
        }
    }
    // The hand-over kernels pay off only between step grids that follow each other directly; an unchained handle uses
    // the grid-wide wait (formally safe after any predecessor) and never pre-fetches across the dependency wait.
    // inside qs_wrap_step the stream predecessor that counts is the block-chained wrapper kernel of the previous control step
    p.chained = (h->chained && (h->in_wrap_step ? h->last_was_wrap : h->last_was_step)) ? 1 : 0;
    p.wrap_chain = (h->in_wrap_step && courier) ? 1 : 0;
    static int poll_mode = -1;
    if (poll_mode < 0) { const char* e = getenv("QS_POLL"); poll_mode = e ? atoi(e) : 40; }
    p.poll_mode = poll_mode;
    h->step_wrap_chain = p.wrap_chain;
    h->wrap_block = work_warps * 32;
#ifdef QS_TIMELINE
    if (!h->tl) { QS_CUDA(cudaMalloc((void**)&h->tl, sizeof(unsigned long long) * 64 * 4096 * 16)); QS_CUDA(cudaMemset(h->tl, 0, sizeof(unsigned long long) * 64 * 4096 * 16)); }
    p.tl = h->tl; p.tl_slot = h->tl_next; h->tl_next = (h->tl_next + 1) % 64;
#endif
    const bool use_ho = h->handover && h->chained && fn_dyn == nullptr;
    const int pdl_mode = use_ho ? 3 : ((pdl_env >= 0 && pdl_env != 3) ? pdl_env : 2);
    const bool use_pdl = pdl_mode != 0;
    p.pdl_mode = pdl_mode;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(grid); lc.blockDim = dim3(kBlock); lc.dynamicSmemBytes = smem; lc.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = attr; lc.numAttrs = use_pdl ? 1 : 0;
    KernelFn fn = fn_dyn ? fn_dyn : (use_ho ? fn_ho : fn_wait);
    if (smem + 1024 > 48 * 1024) QS_CUDA(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      // + the kernel's static words
    static int carve = -2;          // QS_CARVEOUT (experiment): preferred shared-memory carve-out in % for the step and wrapper kernels
    if (carve == -2) { const char* e = getenv("QS_CARVEOUT"); carve = e ? atoi(e) : -1; }
    if (carve >= 0) QS_CUDA(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
    const cudaError_t lerr = cudaLaunchKernelEx(&lc, fn, p);
    if (lerr != cudaSuccess) return fail(QS_ERR_CUDA, std::string("cudaLaunchKernelEx: ") + cudaGetErrorString(lerr));
    QS_CUDA(cudaGetLastError());
    h->launches += 1;
    note_async(h, s, true);
    return QS_OK;
}

static int launch_pregen(QsHandle* h, cudaStream_t s) {
    StepParams p;
    fill_params(h, p);
    const int kBlock = 128;
    const int envs_per_block = kBlock / h->NP;
    const int grid = (h->cfg.num_envs + envs_per_block - 1) / envs_per_block;
    int rc = dispatch_np(h->NP, [&](auto np) {
        qs_pregen_kernel<decltype(np)::value><<<grid, kBlock, 0, s>>>(p);
        return QS_OK;
    });
    if (rc != QS_OK) return rc;
    QS_CUDA(cudaGetLastError());
    h->launches += 1;
    h->since_pregen = 0;
    note_async(h, s, false);
    return QS_OK;
}

static int launch_reset(QsHandle* h, const StepParams& p, cudaStream_t s) {
    const int kBlock = block_size() > 128 ? 128 : block_size();
    const int envs_per_block = kBlock / h->NP;
    const int grid = (h->cfg.num_envs + envs_per_block - 1) / envs_per_block;
    const size_t smem = h->cfg.use_obstacles ? (size_t)envs_per_block * h->M * sizeof(float2) : 0;
    int rc = dispatch_np(h->NP, [&](auto np) {
        qs_reset_kernel<decltype(np)::value><<<grid, kBlock, smem, s>>>(p);
        return QS_OK;
    });
    if (rc != QS_OK) return rc;
    QS_CUDA(cudaGetLastError());
    h->launches += 1;
    note_async(h, s, false);
    return QS_OK;
}

// ------------------------------------------------------------------------------------------
// API
// ------------------------------------------------------------------------------------------
extern "C" int qs_create(const QsConfig* cfg, int device, QsHandle** out) {
    if (!cfg || !out) return fail(QS_ERR_INVALID_ARG, "null argument");
    if (cfg->num_envs < 1) return fail(QS_ERR_INVALID_ARG, "num_envs must be >= 1");
    if (cfg->num_agents < 1) return fail(QS_ERR_INVALID_ARG, "num_agents must be >= 1");
    if (cfg->num_agents > QS_MAX_AGENTS) return fail(QS_ERR_UNSUPPORTED, "num_agents > 32 is not supported by this build");
    if (cfg->obs_repr < 0 || cfg->obs_repr > 2) return fail(QS_ERR_INVALID_ARG, "unknown obs_repr");
    int K = cfg->neighbor_visible_num == -1 ? cfg->num_agents - 1 : cfg->neighbor_visible_num;
    // quadrotor_multi.py:253-274: K must be 0, N-1, or in [1, N-2]
    if (K < 0 || K > cfg->num_agents - 1) return fail(QS_ERR_INVALID_ARG, "Incorrect number of neigbors");
    if (cfg->use_obstacles && cfg->num_obstacles < 1) return fail(QS_ERR_INVALID_ARG, "use_obstacles needs num_obstacles >= 1");
    if (cfg->ep_time <= 0.f) return fail(QS_ERR_INVALID_ARG, "ep_time must be positive");
    if (cfg->scenario < QS_SCENARIO_HOST_TABLES || cfg->scenario > QS_SCENARIO_LAST)
        return fail(QS_ERR_INVALID_ARG, "unknown scenario");
    if (((cfg->scenario >= QS_SCENARIO_DEVICE_FAMILY_FIRST && cfg->scenario < QS_SCENARIO_MIX) || cfg->scenario == QS_SCENARIO_EP_RAND_BEZIER ||
         cfg->scenario == QS_SCENARIO_RUN_AWAY) && cfg->use_obstacles)
        return fail(QS_ERR_INVALID_ARG, "the device-side goal-formation scenarios are obstacle-free (use_obstacles must be 0)");
    if ((cfg->scenario == QS_SCENARIO_O_STATIC_SAME_GOAL || cfg->scenario == QS_SCENARIO_O_RANDOM ||
         (cfg->scenario >= QS_SCENARIO_O_DYNAMIC_SAME_GOAL && cfg->scenario <= QS_SCENARIO_O_EP_RAND_BEZIER)) && !cfg->use_obstacles)
        return fail(QS_ERR_INVALID_ARG, "the o_* scenarios need use_obstacles");
    if (cfg->scenario == QS_SCENARIO_RUN_AWAY && cfg->num_agents < 2)
        return fail(QS_ERR_INVALID_ARG, "run_away needs at least two drones (run_away.py:20 draws randint(1, num_agents))");
    if (cfg->use_obstacles && cfg->scenario != QS_SCENARIO_HOST_TABLES) {
        const int cells = cfg->obst_grid[0] * cfg->obst_grid[1];
        if (cfg->obst_grid[0] < 1 || cfg->obst_grid[1] < 1 || cells > 64)
            return fail(QS_ERR_UNSUPPORTED, "the device-side obstacle scenarios support pillar grids of at most 64 cells");
        if (cells - cfg->num_obstacles < cfg->num_agents)
            return fail(QS_ERR_INVALID_ARG, "obstacle scenario: fewer free grid cells than drones");
    }
    QS_CUDA(cudaSetDevice(device));
    QsHandle* h = new (std::nothrow) QsHandle();
    if (!h) return fail(QS_ERR_INVALID_ARG, "out of host memory");
    memset(h, 0, sizeof(*h));
    h->cfg = *cfg;
    h->device = device;
    h->NP = next_pow2(cfg->num_agents);
    h->K = K;
    h->S = cfg->obs_repr == 0 ? 18 : (cfg->obs_repr == 1 ? 19 : 24);
    h->M = cfg->use_obstacles ? cfg->num_obstacles : 0;
    h->D = h->S + 6 * K + (cfg->use_obstacles ? 9 : 0);
    h->ep_len = (int)((double)cfg->ep_time / (0.005 * 2));      // quadrotor_single.py:158
    h->A = (long long)cfg->num_envs * cfg->num_agents;
    {
        const char* e = getenv("QS_SPLIT");
        h->split_mode = e ? (atoi(e) != 0 ? 1 : 0) : -1;
        h->handover = -1;
        h->pdl_env = -2;
        // next-episode generator cadence: an env consumes its record once per episode, so a quarter of an episode is ample
        const char* pg = getenv("QS_PREGEN");
        const bool dev_gen = cfg->scenario != QS_SCENARIO_HOST_TABLES;
        h->pregen_every = pg ? atoi(pg) : (dev_gen ? (h->ep_len / 4 < 16 ? 16 : (h->ep_len / 4 > 256 ? 256 : h->ep_len / 4)) : 0);
        if (!dev_gen) h->pregen_every = 0;
        h->since_pregen = 0;
        const char* c = getenv("QS_CHAINED");
        h->chained = c ? (atoi(c) != 0) : 0;
        const char* b = getenv("QS_OBS_BULK");
        h->bulk_mode = b ? atoi(b) : -1;          // 0 vector stores only, 2 linear bulk copies only, else automatic
    }
    h->a_pad = (h->A + 31) / 32 * 32;
    // QuadrotorEnvMulti defaults, quadrotor_multi.py:91-94
    const float def[QS_NUM_REW_COEFF] = {1.f, 0.05f, 1.f, 1.f, 0.1f, 5.f, 4.f, 5.f};
    memcpy(h->rew, def, sizeof(def));
    DevState& st = h->st;
    st.a_pad = h->a_pad;
    const long long E = cfg->num_envs, A = h->A, M = h->M > 0 ? h->M : 1;
#define QS_ALLOC0(ptr, bytes)                                     \
    do {                                                          \
        QS_CUDA(cudaMalloc((void**)&(ptr), (size_t)(bytes)));     \
        QS_CUDA(cudaMemset((ptr), 0, (size_t)(bytes)));           \
    } while (0)
    QS_ALLOC0(st.slots, sizeof(float4) * NUM_SLOTS * h->a_pad);
    QS_ALLOC0(st.env_ctr, sizeof(int4) * E);
    QS_ALLOC0(st.env_cnt, sizeof(int32_t) * E * QS_NUM_ENV_STATS);
    QS_ALLOC0(st.obst, sizeof(float2) * E * M);
    QS_ALLOC0(st.next_goal, sizeof(float4) * A);
    QS_ALLOC0(st.next_spawn, sizeof(float4) * A);
    QS_ALLOC0(st.next_obst, sizeof(float2) * E * M);
    QS_ALLOC0(st.stats_env, sizeof(int32_t) * E * QS_NUM_ENV_STATS);
    QS_ALLOC0(st.stats_agent, sizeof(float4) * A);
    QS_ALLOC0(st.scn_i, sizeof(int4) * E);
    QS_ALLOC0(st.scn_f, sizeof(float4) * 3 * E);
    QS_ALLOC0(st.next_scn_i, sizeof(int4) * E);
    QS_ALLOC0(st.next_scn_f, sizeof(float4) * 3 * E);
    QS_ALLOC0(st.epi, sizeof(int2) * E);
    {   // per-block hand-over words (at most one block per env), all "ready"
        // hand-over words per block, [HW_ROWS][E + 1] (rows HW_*, qs_step.cuh): the `ready` flag of the thread-0 hand-over (1) and
        // the courier warps' counters T, S, D, Tw, Dw, Rw (0); [0][E] is the time-out counter
        QS_CUDA(cudaMalloc((void**)&st.ready, sizeof(int) * HW_ROWS * (E + 1)));
        std::vector<int> init((size_t)HW_ROWS * (E + 1), 0);
        for (long long k = 0; k < E; ++k) init[(size_t)k] = 1;
        QS_CUDA(cudaMemcpy(st.ready, init.data(), sizeof(int) * HW_ROWS * (E + 1), cudaMemcpyHostToDevice));
    }
    // rotation = identity so that a never-reset env still holds a valid state
    {
        std::string tmp;
        float4* init = (float4*)malloc(sizeof(float4) * h->a_pad);
        for (long long a = 0; a < h->a_pad; ++a) init[a] = make_float4(0.f, 1.f, 0.f, 0.f);   // omega.z, R00, R01, R02
        QS_CUDA(cudaMemcpy(st.slots + SL_OM_R0 * h->a_pad, init, sizeof(float4) * h->a_pad, cudaMemcpyHostToDevice));
        for (long long a = 0; a < h->a_pad; ++a) init[a] = make_float4(0.f, 1.f, 0.f, 0.f);   // R10, R11, R12, R20
        QS_CUDA(cudaMemcpy(st.slots + SL_R1_R20 * h->a_pad, init, sizeof(float4) * h->a_pad, cudaMemcpyHostToDevice));
        for (long long a = 0; a < h->a_pad; ++a) init[a] = make_float4(0.f, 1.f, 0.f, 0.f);   // R21, R22, flags, prev
        QS_CUDA(cudaMemcpy(st.slots + SL_R2_FLAGS * h->a_pad, init, sizeof(float4) * h->a_pad, cudaMemcpyHostToDevice));
        free(init);
    }
    // default episode table: every goal at (0, 0, 2), spawn at the goal (scenarios/base.py:137-139, static_same_goal)
    {
        float4* g = (float4*)malloc(sizeof(float4) * A);
        for (long long a = 0; a < A; ++a) g[a] = make_float4(0.f, 0.f, 2.f, 0.f);
        QS_CUDA(cudaMemcpy(st.next_goal, g, sizeof(float4) * A, cudaMemcpyHostToDevice));
        QS_CUDA(cudaMemcpy(st.slots + SL_GOAL * h->a_pad, g, sizeof(float4) * A, cudaMemcpyHostToDevice));
        free(g);
    }
    // staging buffers for the host entry points
    QS_CUDA(cudaMalloc((void**)&h->d_actions, sizeof(float) * 4 * A));
    QS_CUDA(cudaMalloc((void**)&h->d_obs, sizeof(float) * h->D * A));
    QS_CUDA(cudaMalloc((void**)&h->d_rewards, sizeof(float) * A));
    QS_CUDA(cudaMalloc((void**)&h->d_terms, sizeof(float) * QS_NUM_TERMS * A));
    QS_CUDA(cudaMalloc((void**)&h->d_dones, A));
    QS_CUDA(cudaMalloc((void**)&h->d_mask, E));
    QS_CUDA(cudaMallocHost((void**)&h->h_actions, sizeof(float) * 4 * A));
    QS_CUDA(cudaMallocHost((void**)&h->h_obs, sizeof(float) * h->D * A));
    QS_CUDA(cudaMallocHost((void**)&h->h_rewards, sizeof(float) * A));
    QS_CUDA(cudaMallocHost((void**)&h->h_terms, sizeof(float) * QS_NUM_TERMS * A));
    QS_CUDA(cudaMallocHost((void**)&h->h_dones, A));
    QS_CUDA(cudaMallocHost((void**)&h->h_mask, E));
    QS_CUDA(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    QS_CUDA(cudaEventCreateWithFlags(&h->ev_sync, cudaEventDisableTiming));
    QS_CUDA(cudaHostAlloc((void**)&h->err_host, sizeof(int), cudaHostAllocMapped));
    *h->err_host = 0;
    QS_CUDA(cudaHostGetDevicePointer((void**)&st.err_flag, h->err_host, 0));
    *out = h;
    return QS_OK;
}

extern "C" int qs_destroy(QsHandle* h) {
    if (!h) return QS_OK;
    cudaSetDevice(h->device);
    DevState& st = h->st;
    cudaFree(st.slots); cudaFree(st.env_ctr); cudaFree(st.env_cnt); cudaFree(st.obst); cudaFree(st.next_goal);
    cudaFree(st.next_spawn); cudaFree(st.next_obst); cudaFree(st.stats_env); cudaFree(st.stats_agent);
    cudaFree(st.dyn); cudaFree(st.next_dyn); cudaFree(st.dyn_pending);
    cudaFree(st.scn_i); cudaFree(st.scn_f); cudaFree(st.ready); cudaFree(st.next_scn_i); cudaFree(st.next_scn_f); cudaFree(st.epi);
    cudaFree(h->d_actions); cudaFree(h->d_obs); cudaFree(h->d_rewards); cudaFree(h->d_terms); cudaFree(h->d_dones);
    cudaFree(h->d_mask);
    cudaFreeHost(h->h_actions); cudaFreeHost(h->h_obs); cudaFreeHost(h->h_rewards); cudaFreeHost(h->h_terms);
    cudaFreeHost(h->h_dones); cudaFreeHost(h->h_mask);
    if (h->wrap_on) {
        WrapState& w = h->wrap;
        cudaFree(w.acc); cudaFree(w.ep_steps); cudaFree(w.true_reward); cudaFree(w.agg); cudaFree(w.snap_slots); cudaFree(w.snap_obs);
        cudaFree(w.snap_env); cudaFree(w.snap_obst); cudaFree(w.rp); cudaFree(w.rq); cudaFree(w.crash_now); cudaFree(w.crash_hist);
        cudaFree(w.ev_state); cudaFreeHost(h->wrap_agg_host);
    }
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->ev_sync) cudaEventDestroy(h->ev_sync);
    if (h->err_host) cudaFreeHost(h->err_host);
    delete h;
    return QS_OK;
}

extern "C" int qs_obs_dim(const QsHandle* h) { return h ? h->D : 0; }
extern "C" int qs_num_envs(const QsHandle* h) { return h ? h->cfg.num_envs : 0; }
extern "C" int qs_num_agents(const QsHandle* h) { return h ? h->cfg.num_agents : 0; }
extern "C" int qs_num_obstacles(const QsHandle* h) { return h ? h->M : 0; }
extern "C" int qs_ep_len(const QsHandle* h) { return h ? h->ep_len : 0; }
extern "C" int64_t qs_launch_count(const QsHandle* h) { return h ? h->launches : 0; }

extern "C" int64_t qs_handover_timeouts(QsHandle* h) {
    if (!h) return -1;
    int v = 0;
    if (cudaSetDevice(h->device) != cudaSuccess) return -1;
    if (cudaMemcpy(&v, h->st.ready + h->cfg.num_envs, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return v;
}

#ifdef QS_TIMELINE
// debug build: copies the [64][4096][8] stamp buffer to the host (synchronises)
extern "C" int qs_debug_timeline(QsHandle* h, unsigned long long* out_host) {
    if (!h || !h->tl) return fail(QS_ERR_INVALID_ARG, "no timeline");
    QS_CUDA(cudaMemcpy(out_host, h->tl, sizeof(unsigned long long) * 64 * 4096 * 16, cudaMemcpyDeviceToHost));
    return QS_OK;
}
extern "C" int qs_debug_timeline_rewind(QsHandle* h) { if (h) h->tl_next = 0; return QS_OK; }
#endif

// ------------------------------------------------------------------------------------------
// training wrappers (qs_wrap.cuh)
// ------------------------------------------------------------------------------------------
extern "C" int qs_wrap_enable(QsHandle* h, const QsWrapConfig* cfg) {
    if (!h || !cfg) return fail(QS_ERR_INVALID_ARG, "null argument");
    if (h->wrap_on) return fail(QS_ERR_INVALID_ARG, "wrappers already enabled");
    if (cfg->use_replay && h->cfg.scenario == QS_SCENARIO_HOST_TABLES)
        return fail(QS_ERR_UNSUPPORTED, "collision-event replay on the device needs device-side scenarios (host scenario objects are not part of a snapshot)");
    if (cfg->use_replay && (cfg->replay_buffer_size < 1 || cfg->replay_buffer_size > 64)) return fail(QS_ERR_INVALID_ARG, "replay_buffer_size must be 1..64");
    QS_CUDA(cudaSetDevice(h->device));
    WrapState& w = h->wrap;
    memset(&w, 0, sizeof(w));
    const long long A = h->A, E = h->cfg.num_envs, N = h->cfg.num_agents, M = h->M > 0 ? h->M : 1;
    QS_ALLOC0(w.acc, sizeof(float4) * 6 * A);
    QS_ALLOC0(w.ep_steps, sizeof(int) * E);
    QS_ALLOC0(w.true_reward, sizeof(float) * A);
    QS_ALLOC0(w.agg, sizeof(float) * QS_WRAP_AGG);
    QS_CUDA(cudaMallocHost((void**)&h->wrap_agg_host, sizeof(float) * QS_WRAP_AGG));
    w.replay_on = cfg->use_replay ? 1 : 0;
    w.replay_prob = cfg->replay_prob;
    w.always_active = cfg->replay_always_active ? 1 : 0;
    if (w.replay_on) {
        w.buffer = cfg->replay_buffer_size;
        w.slots = RP_KEEP + w.buffer;
        QS_ALLOC0(w.snap_slots, sizeof(float4) * E * w.slots * NUM_SLOTS * N);
        QS_ALLOC0(w.snap_obs, sizeof(float) * E * w.slots * N * h->D);
        QS_ALLOC0(w.snap_env, sizeof(int32_t) * E * w.slots * SNAP_ENV_I32);
        QS_ALLOC0(w.snap_obst, sizeof(float2) * E * w.slots * M);
        QS_ALLOC0(w.rp, sizeof(int4) * E);
        QS_ALLOC0(w.rq, sizeof(int4) * E);
        QS_ALLOC0(w.crash_now, sizeof(float) * E);
        QS_ALLOC0(w.crash_hist, sizeof(float) * E * 100);
        QS_CUDA(cudaMalloc((void**)&w.ev_state, sizeof(int32_t) * E * w.buffer));
        QS_CUDA(cudaMemset(w.ev_state, 0xff, sizeof(int32_t) * E * w.buffer));          // -1: empty
        std::vector<int4> rp((size_t)E, make_int4(0, 0, 0, -(1 << 30)));
        QS_CUDA(cudaMemcpy(w.rp, rp.data(), sizeof(int4) * E, cudaMemcpyHostToDevice));
        if (w.always_active) {
            std::vector<int4> rq((size_t)E, make_int4(0, 1, 0, 0));
            QS_CUDA(cudaMemcpy(w.rq, rq.data(), sizeof(int4) * E, cudaMemcpyHostToDevice));
        }
    }
    h->wrap_on = true;
    return QS_OK;
}

static int launch_wrap(QsHandle* h, const float* actions_dev, const float* terms_dev, float* obs_dev, uint8_t* dones_dev, void* stream, bool chain = false);

extern "C" int qs_wrap_step(QsHandle* h, const float* actions_dev, float* obs_dev, float* rewards_dev, uint8_t* dones_dev, void* stream) {
    if (!h || !h->wrap_on) return fail(QS_ERR_INVALID_ARG, "qs_wrap_enable first");
    h->in_wrap_step = 1;
    int rc = qs_step(h, actions_dev, obs_dev, rewards_dev, dones_dev, h->d_terms, stream);
    h->in_wrap_step = 0;
    if (rc != QS_OK) return rc;
    return launch_wrap(h, actions_dev, h->d_terms, obs_dev, dones_dev, stream, h->step_wrap_chain != 0);
}

extern "C" int qs_wrap_apply(QsHandle* h, const float* actions_dev, const float* rew_terms_dev, float* obs_dev, const uint8_t* dones_dev,
                             void* stream) {
    if (!h || !h->wrap_on || !actions_dev || !rew_terms_dev || !obs_dev || !dones_dev) return fail(QS_ERR_INVALID_ARG, "null argument / wrappers not enabled");
    if ((((uintptr_t)actions_dev | (uintptr_t)rew_terms_dev) & 15u) != 0) return fail(QS_ERR_INVALID_ARG, "actions / terms must be 16-byte aligned");
    return launch_wrap(h, actions_dev, rew_terms_dev, obs_dev, (uint8_t*)dones_dev, stream);
}

static int launch_wrap(QsHandle* h, const float* actions_dev, const float* terms_dev, float* obs_dev, uint8_t* dones_dev, void* stream, bool chain) {
    int rc = QS_OK;
    static int probe = -1;                 // QS_WRAP_PROBE (tuning): 1 = skip the wrapper kernel, 2 = launch it without PDL
    if (probe < 0) { const char* e = getenv("QS_WRAP_PROBE"); probe = e ? atoi(e) : 0; }
    if (probe == 1) return QS_OK;
    WrapParams q;
    fill_params(h, q.sp);
    q.w = h->wrap;
    q.sp.actions = (const float4*)actions_dev;
    q.sp.rew_terms = const_cast<float*>(terms_dev);
    q.sp.dones = dones_dev;
    q.sp.obs = obs_dev;
    q.chain = chain ? 1 : 0;
#ifdef QS_TIMELINE
    q.sp.tl = h->tl; q.sp.tl_slot = (h->tl_next + 63) % 64;      // the slot of the step grid this kernel follows
#endif
    const int kBlock = chain ? h->wrap_block : 128;            // chained: block b covers the envs of step block b
    const int envs_per_block = kBlock / h->NP;
    const int grid = (h->cfg.num_envs + envs_per_block - 1) / envs_per_block;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(grid); lc.blockDim = dim3(kBlock); lc.dynamicSmemBytes = 0; lc.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // launched behind the step grid's late trigger
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = attr; lc.numAttrs = probe == 2 ? 0 : 1;
    using WrapFn = void (*)(WrapParams);
    WrapFn fn = nullptr;
    rc = dispatch_np(h->NP, [&](auto np) { fn = (WrapFn)qs_wrap_kernel<decltype(np)::value>; return QS_OK; });
    if (rc != QS_OK) return rc;
    {
        const char* e = getenv("QS_CARVEOUT");
        if (e && atoi(e) >= 0) QS_CUDA(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e)));
        const char* d = getenv("QS_WRAP_SMEM_KB");          // experiment: dummy dynamic shared memory of the wrapper kernel
        if (e || d) { lc.dynamicSmemBytes = d ? (size_t)atoi(d) * 1024 : 0; if (lc.dynamicSmemBytes > 48 * 1024) QS_CUDA(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lc.dynamicSmemBytes)); }
    }
    const cudaError_t lerr = cudaLaunchKernelEx(&lc, fn, q);
    if (lerr != cudaSuccess) return fail(QS_ERR_CUDA, std::string("cudaLaunchKernelEx(wrap): ") + cudaGetErrorString(lerr));
    h->launches += 1;
    note_async(h, (cudaStream_t)stream, false);
    h->last_was_wrap = chain ? 1 : 0;
    return QS_OK;
}

extern "C" int qs_wrap_read(QsHandle* h, float* agg_host, int reset, void* stream) {
    if (!h || !h->wrap_on || !agg_host) return fail(QS_ERR_INVALID_ARG, "null argument / wrappers not enabled");
    QS_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    QS_CUDA(cudaMemcpyAsync(h->wrap_agg_host, h->wrap.agg, sizeof(float) * QS_WRAP_AGG, cudaMemcpyDeviceToHost, s));
    if (reset) QS_CUDA(cudaMemsetAsync(h->wrap.agg, 0, sizeof(float) * QS_WRAP_AGG, s));
    QS_CUDA(cudaStreamSynchronize(s));
    memcpy(agg_host, h->wrap_agg_host, sizeof(float) * QS_WRAP_AGG);
    note_async(h, s, false);
    return QS_OK;
}

extern "C" int qs_wrap_true_reward(QsHandle* h, float* out_dev, void* stream) {
    if (!h || !h->wrap_on || !out_dev) return fail(QS_ERR_INVALID_ARG, "null argument / wrappers not enabled");
    QS_CUDA(cudaSetDevice(h->device));
    QS_CUDA(cudaMemcpyAsync(out_dev, h->wrap.true_reward, sizeof(float) * h->A, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    note_async(h, (cudaStream_t)stream, false);
    return QS_OK;
}

extern "C" int qs_set_chained(QsHandle* h, int on) {
    if (!h) return fail(QS_ERR_INVALID_ARG, "null argument");
    h->chained = on ? 1 : 0;
    h->last_was_step = 0;
    h->last_was_wrap = 0;
    return QS_OK;
}

extern "C" int qs_set_obstacle_randomization(QsHandle* h, const float* densities_host, int n_densities, const float* sizes_host, int n_sizes) {
    if (!h) return fail(QS_ERR_INVALID_ARG, "null argument");
    if (n_densities == 0 && n_sizes == 0) { h->obst_random = 0; return QS_OK; }
    if (!h->cfg.use_obstacles || h->cfg.scenario == QS_SCENARIO_HOST_TABLES)
        return fail(QS_ERR_INVALID_ARG, "obstacle randomisation needs use_obstacles and a device-side obstacle scenario");
    if (n_densities < 1 || n_sizes < 1 || n_densities > QS_MAX_OBST_CHOICES || n_sizes > QS_MAX_OBST_CHOICES || !densities_host || !sizes_host)
        return fail(QS_ERR_INVALID_ARG, "1..16 densities and sizes are needed");
    const double area = (double)h->cfg.obst_grid[0] * h->cfg.obst_grid[1];
    for (int k = 0; k < n_densities; ++k) {
        const int m = (int)((double)densities_host[k] * area);                   // quadrotor_multi.py:128
        if (m < 0 || m > h->M) return fail(QS_ERR_INVALID_ARG, "a density needs more pillars than QsConfig.num_obstacles (the table size)");
        if ((int)area - m < h->cfg.num_agents) return fail(QS_ERR_INVALID_ARG, "a density leaves fewer free cells than drones");
        h->obst_counts[k] = m;
    }
    for (int k = 0; k < n_sizes; ++k) {
        if (!(sizes_host[k] >= 0.f)) return fail(QS_ERR_INVALID_ARG, "negative pillar size");
        h->obst_radii[k] = sizes_host[k] * 0.5f;
    }
    h->n_obst_counts = n_densities; h->n_obst_radii = n_sizes;
    h->obst_random = 1;
    return QS_OK;
}

extern "C" int qs_set_dynamics(QsHandle* h, const uint8_t* env_mask_dev, const float* rows_dev, int at_next_reset, void* stream) {
    if (!h || !rows_dev) return fail(QS_ERR_INVALID_ARG, "null argument");
    if (((uintptr_t)rows_dev & 15u) != 0) return fail(QS_ERR_INVALID_ARG, "rows must be 16-byte aligned");
    QS_CUDA(cudaSetDevice(h->device));
    DevState& st = h->st;
    const long long A = h->A, E = h->cfg.num_envs;
    if (st.dyn == nullptr) {
        // first use: both tables start as Crazyflie rows (the constants compiled into the other kernels), so that envs
        // outside a mask keep flying the default model
        std::vector<float> row(QS_DYN_ROW, 0.f);
        const float cf[QS_DYN_ROW] = {MASS, INV_MASS, IXX, IYY, IZZ, INV_IXX, INV_IYY, INV_IZZ, THRUST_MAX, THRUST_MAX, THRUST_MAX, THRUST_MAX,
                                      TORQUE_MAX, TORQUE_MAX, TORQUE_MAX, TORQUE_MAX, PROP_ARM_XY, -PROP_ARM_XY, -PROP_ARM_XY, -PROP_ARM_XY,
                                      -PROP_ARM_XY, PROP_ARM_XY, PROP_ARM_XY, PROP_ARM_XY, 0.f, 0.f, 0.f, 0.f, MOTOR_TAU_UP, MOTOR_TAU_DOWN,
                                      1.0f, OU_SIGMA, 0.f, 0.f, 0.f, 0.f, ARM, 0.f, 0.f, 0.f};
        std::vector<float> all((size_t)A * QS_DYN_ROW);
        for (long long a = 0; a < A; ++a) memcpy(&all[(size_t)a * QS_DYN_ROW], cf, sizeof(cf));
        float4 *d0 = nullptr, *d1 = nullptr;
        int* pend = nullptr;
        QS_CUDA(cudaMalloc((void**)&d0, sizeof(float) * QS_DYN_ROW * A));
        QS_CUDA(cudaMalloc((void**)&d1, sizeof(float) * QS_DYN_ROW * A));
        QS_CUDA(cudaMalloc((void**)&pend, sizeof(int) * E));
        QS_CUDA(cudaMemcpy(d0, all.data(), sizeof(float) * QS_DYN_ROW * A, cudaMemcpyHostToDevice));
        QS_CUDA(cudaMemcpy(d1, all.data(), sizeof(float) * QS_DYN_ROW * A, cudaMemcpyHostToDevice));
        QS_CUDA(cudaMemset(pend, 0, sizeof(int) * E));
        st.dyn = d0; st.next_dyn = d1; st.dyn_pending = pend;
    }
    const long long n = A * (QS_DYN_ROW / 4);
    k_set_dynamics<<<blocks_for(n > E ? n : E), 256, 0, (cudaStream_t)stream>>>(st, h->cfg.num_envs, h->cfg.num_agents, env_mask_dev,
                                                                               (const float4*)rows_dev, at_next_reset ? 1 : 0);
    QS_CUDA(cudaGetLastError());
    h->launches += 1;
    note_async(h, (cudaStream_t)stream, false);
    return QS_OK;
}

extern "C" int qs_set_reward_coeffs(QsHandle* h, const float* coeffs_host) {
    if (!h || !coeffs_host) return fail(QS_ERR_INVALID_ARG, "null argument");
    for (int k = 0; k < QS_NUM_REW_COEFF; ++k) {
        if (!(coeffs_host[k] == coeffs_host[k])) return fail(QS_ERR_INVALID_ARG, "reward coefficient is NaN");
        h->rew[k] = coeffs_host[k];
    }
    return QS_OK;
}


extern "C" int qs_set_next_episode(QsHandle* h, const uint8_t* env_mask_dev, const float* goals_dev, const float* spawn_dev,
                                   const float* obst_xy_dev, void* stream) {
    if (!h || !goals_dev) return fail(QS_ERR_INVALID_ARG, "null argument");
    if (obst_xy_dev && !h->cfg.use_obstacles) return fail(QS_ERR_INVALID_ARG, "obstacle table given but use_obstacles = 0");
    QS_CUDA(cudaSetDevice(h->device));
    const long long n = h->A > (long long)h->cfg.num_envs * h->M ? h->A : (long long)h->cfg.num_envs * h->M;
    k_set_next_episode<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(h->st, h->cfg.num_envs, h->cfg.num_agents, h->M,
                                                                         env_mask_dev, goals_dev, spawn_dev, obst_xy_dev);
    QS_CUDA(cudaGetLastError());
    h->launches += 1;
    note_async(h, (cudaStream_t)stream, false);
    return QS_OK;
}

extern "C" int qs_set_goals(QsHandle* h, const uint8_t* env_mask_dev, const float* goals_dev, void* stream) {
    if (!h || !goals_dev) return fail(QS_ERR_INVALID_ARG, "null argument");
    QS_CUDA(cudaSetDevice(h->device));
    k_set_goals<<<blocks_for(h->A), 256, 0, (cudaStream_t)stream>>>(h->st, h->cfg.num_envs, h->cfg.num_agents, env_mask_dev,
                                                                     goals_dev);
    QS_CUDA(cudaGetLastError());
    h->launches += 1;
    note_async(h, (cudaStream_t)stream, false);
    return QS_OK;
}

extern "C" int qs_reset(QsHandle* h, const uint8_t* env_mask_dev, float* obs_dev, void* stream) {
    if (!h || !obs_dev) return fail(QS_ERR_INVALID_ARG, "null argument");
    QS_CUDA(cudaSetDevice(h->device));
    StepParams p;
    fill_params(h, p);
    p.obs = obs_dev;
    p.env_mask = env_mask_dev;
    int rc = launch_reset(h, p, (cudaStream_t)stream);
    if (rc == QS_OK && h->pregen_every > 0) rc = launch_pregen(h, (cudaStream_t)stream);
    return rc;
}

extern "C" int qs_step(QsHandle* h, const float* actions_dev, float* obs_dev, float* rewards_dev, uint8_t* dones_dev,
                       float* rew_terms_dev, void* stream) {
    if (!h || !actions_dev || !obs_dev || !rewards_dev || !dones_dev) return fail(QS_ERR_INVALID_ARG, "null argument");
    if (((uintptr_t)actions_dev & 15u) != 0) return fail(QS_ERR_INVALID_ARG, "actions must be 16-byte aligned");
    QS_CUDA(cudaSetDevice(h->device));
    StepParams p;
    fill_params(h, p);
    p.actions = (const float4*)actions_dev;
    p.obs = obs_dev; p.rewards = rewards_dev; p.dones = dones_dev; p.rew_terms = rew_terms_dev;
    return launch_step(h, p, (cudaStream_t)stream);
}

extern "C" int qs_rollout(QsHandle* h, int num_steps, const float* actions_dev, float* obs_dev, float* rewards_dev,
                          uint8_t* dones_dev, int last_obs_only, void* stream) {
    if (!h || !actions_dev || !obs_dev || !rewards_dev || !dones_dev) return fail(QS_ERR_INVALID_ARG, "null argument");
    if (num_steps < 1) return fail(QS_ERR_INVALID_ARG, "num_steps must be >= 1");
    if (((uintptr_t)actions_dev & 15u) != 0) return fail(QS_ERR_INVALID_ARG, "actions must be 16-byte aligned");
    QS_CUDA(cudaSetDevice(h->device));
    StepParams p;
    fill_params(h, p);
    p.actions = (const float4*)actions_dev;
    p.obs = obs_dev; p.rewards = rewards_dev; p.dones = dones_dev; p.rew_terms = nullptr;
    p.T = num_steps; p.last_obs_only = last_obs_only ? 1 : 0;
    return launch_step(h, p, (cudaStream_t)stream);
}

#ifndef QS_ZERO_COPY_DEFAULT
#define QS_ZERO_COPY_DEFAULT true          // measured on c3: 148 -> 134 us per host-buffer step (profiles/r01_notes.md)
#endif
// true when the host pointer is page-locked (cudaHostAlloc / cudaHostRegister): DMA can use it directly.  `dev` receives the
// device alias of a mapped buffer (or null).  The last few answers are cached: a rollout worker passes the same buffers on
// every step, and the two driver queries per buffer cost more host time than enqueueing the step.
static bool is_pinned(const void* p, void** dev = nullptr) {
    struct Entry { const void* p; bool pinned; void* dev; };
    static thread_local Entry cache[16];
    static thread_local int next = 0;
    for (int k = 0; k < 16; ++k)
        if (cache[k].p == p && p != nullptr) {
            if (dev) *dev = cache[k].dev;
            return cache[k].pinned;
        }
    cudaPointerAttributes at;
    bool pinned = false;
    void* d = nullptr;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) cudaGetLastError();
    else pinned = at.type == cudaMemoryTypeHost;
    if (pinned && cudaHostGetDevicePointer(&d, (void*)p, 0) != cudaSuccess) { cudaGetLastError(); d = nullptr; }
    cache[next] = {p, pinned, d};
    next = (next + 1) % 16;
    if (dev) *dev = d;
    return pinned;
}

static int step_host_impl(QsHandle* h, const float* actions_host, float* obs_host, float* rewards_host, uint8_t* dones_host,
                          float* rew_terms_host, bool sync);

extern "C" int qs_step_host(QsHandle* h, const float* actions_host, float* obs_host, float* rewards_host, uint8_t* dones_host,
                            float* rew_terms_host) {
    return step_host_impl(h, actions_host, obs_host, rewards_host, dones_host, rew_terms_host, true);
}

extern "C" int qs_step_host_async(QsHandle* h, const float* actions_host, float* obs_host, float* rewards_host, uint8_t* dones_host,
                                  float* rew_terms_host) {
    return step_host_impl(h, actions_host, obs_host, rewards_host, dones_host, rew_terms_host, false);
}

extern "C" int qs_wait(QsHandle* h) {
    if (!h) return fail(QS_ERR_INVALID_ARG, "null argument");
    QS_CUDA(cudaSetDevice(h->device));
    QS_CUDA(cudaStreamSynchronize(h->own_stream));
    return QS_OK;
}

static int step_host_impl(QsHandle* h, const float* actions_host, float* obs_host, float* rewards_host, uint8_t* dones_host,
                          float* rew_terms_host, bool sync) {
    if (!h || !actions_host || !obs_host || !rewards_host || !dones_host) return fail(QS_ERR_INVALID_ARG, "null argument");
    QS_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = h->own_stream;
    join_caller_stream(h);
    const long long A = h->A;
    // pageable buffers go through the handle's pinned staging; page-locked caller buffers are used as they are
    void *da = nullptr, *dob = nullptr, *dr = nullptr, *dd = nullptr, *dt = nullptr;
    const bool pa = is_pinned(actions_host, &da), po = is_pinned(obs_host, &dob), pr = is_pinned(rewards_host, &dr),
               pd = is_pinned(dones_host, &dd), pt = rew_terms_host && is_pinned(rew_terms_host, &dt);
    if (!sync && !(pa && po && pr && pd && (!rew_terms_host || pt)))
        return fail(QS_ERR_INVALID_ARG, "qs_step_host_async needs page-locked buffers (pageable ones would need a copy after the wait)");
    // Zero-copy path (all caller buffers page-locked and mapped): the kernel reads the actions from, and writes its
    // outputs straight to, host memory — coalesced 128-bit stores over PCIe overlap the transfer with the step and save
    // the four copy launches (QS_ZERO_COPY=0 falls back to explicit copies).
    const char* zc_env = getenv("QS_ZERO_COPY");          // read per call: bench.py times both paths in one process
    const bool zero_copy = zc_env ? atoi(zc_env) != 0 : QS_ZERO_COPY_DEFAULT;
    if (zero_copy && pa && po && pr && pd && (!rew_terms_host || pt)) {
        const bool ok = da && dob && dr && dd && (!rew_terms_host || dt);
        static int zc_bulk = -1;                           // QS_ZC_BULK=1 (experiment): observation tiles leave through the bulk-copy engine
        if (zc_bulk < 0) { const char* e = getenv("QS_ZC_BULK"); zc_bulk = e ? atoi(e) : 0; }
        if (ok) {
            StepParams p;
            fill_params(h, p);
            p.actions = (const float4*)da;
            p.obs = (float*)dob; p.rewards = (float*)dr; p.dones = (uint8_t*)dd; p.rew_terms = (float*)dt;
            int rc0 = launch_step(h, p, s, /*obs_in_device_memory=*/zc_bulk != 0);
            if (rc0 != QS_OK) return rc0;
            if (sync) QS_CUDA(cudaStreamSynchronize(s));
            return QS_OK;
        }
    }
    const float* a_src = actions_host;
    if (!pa) { memcpy(h->h_actions, actions_host, sizeof(float) * 4 * A); a_src = h->h_actions; }
    QS_CUDA(cudaMemcpyAsync(h->d_actions, a_src, sizeof(float) * 4 * A, cudaMemcpyHostToDevice, s));
    int rc = qs_step(h, h->d_actions, h->d_obs, h->d_rewards, h->d_dones, rew_terms_host ? h->d_terms : nullptr, s);
    if (rc != QS_OK) return rc;
    QS_CUDA(cudaMemcpyAsync(po ? obs_host : h->h_obs, h->d_obs, sizeof(float) * h->D * A, cudaMemcpyDeviceToHost, s));
    QS_CUDA(cudaMemcpyAsync(pr ? rewards_host : h->h_rewards, h->d_rewards, sizeof(float) * A, cudaMemcpyDeviceToHost, s));
    QS_CUDA(cudaMemcpyAsync(pd ? dones_host : h->h_dones, h->d_dones, A, cudaMemcpyDeviceToHost, s));
    if (rew_terms_host)
        QS_CUDA(cudaMemcpyAsync(pt ? rew_terms_host : h->h_terms, h->d_terms, sizeof(float) * QS_NUM_TERMS * A, cudaMemcpyDeviceToHost, s));
    if (!sync) return QS_OK;
    QS_CUDA(cudaStreamSynchronize(s));
    if (!po) memcpy(obs_host, h->h_obs, sizeof(float) * h->D * A);
    if (!pr) memcpy(rewards_host, h->h_rewards, sizeof(float) * A);
    if (!pd) memcpy(dones_host, h->h_dones, A);
    if (rew_terms_host && !pt) memcpy(rew_terms_host, h->h_terms, sizeof(float) * QS_NUM_TERMS * A);
    return QS_OK;
}

extern "C" int qs_reset_host(QsHandle* h, const uint8_t* env_mask_host, float* obs_host) {
    if (!h || !obs_host) return fail(QS_ERR_INVALID_ARG, "null argument");
    QS_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = h->own_stream;
    join_caller_stream(h);
    const long long A = h->A;
    if (env_mask_host) {
        memcpy(h->h_mask, env_mask_host, h->cfg.num_envs);
        QS_CUDA(cudaMemcpyAsync(h->d_mask, h->h_mask, h->cfg.num_envs, cudaMemcpyHostToDevice, s));
        // rows of unmasked envs keep the caller's values
        memcpy(h->h_obs, obs_host, sizeof(float) * h->D * A);
        QS_CUDA(cudaMemcpyAsync(h->d_obs, h->h_obs, sizeof(float) * h->D * A, cudaMemcpyHostToDevice, s));
    }
    int rc = qs_reset(h, env_mask_host ? h->d_mask : nullptr, h->d_obs, s);
    if (rc != QS_OK) return rc;
    QS_CUDA(cudaMemcpyAsync(h->h_obs, h->d_obs, sizeof(float) * h->D * A, cudaMemcpyDeviceToHost, s));
    QS_CUDA(cudaStreamSynchronize(s));
    memcpy(obs_host, h->h_obs, sizeof(float) * h->D * A);
    return QS_OK;
}

extern "C" int qs_get_state(QsHandle* h, float* agent_f32_dev, uint32_t* agent_u32_dev, int32_t* env_i32_dev,
                            float* obst_xy_dev, void* stream) {
    if (!h || !agent_f32_dev || !agent_u32_dev || !env_i32_dev) return fail(QS_ERR_INVALID_ARG, "null argument");
    QS_CUDA(cudaSetDevice(h->device));
    const long long n = h->A > (long long)h->cfg.num_envs * h->M ? h->A : (long long)h->cfg.num_envs * h->M;
    k_get_state<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(h->st, h->cfg.num_envs, h->cfg.num_agents, h->M, agent_f32_dev,
                                                                  agent_u32_dev, env_i32_dev, h->M > 0 ? obst_xy_dev : nullptr);
    QS_CUDA(cudaGetLastError());
    h->launches += 1;
    note_async(h, (cudaStream_t)stream, false);
    return QS_OK;
}

extern "C" int qs_set_state(QsHandle* h, const uint8_t* env_mask_dev, const float* agent_f32_dev, const uint32_t* agent_u32_dev,
                            const int32_t* env_i32_dev, const float* obst_xy_dev, void* stream) {
    if (!h || !agent_f32_dev || !agent_u32_dev || !env_i32_dev) return fail(QS_ERR_INVALID_ARG, "null argument");
    QS_CUDA(cudaSetDevice(h->device));
    const long long n = h->A > (long long)h->cfg.num_envs * h->M ? h->A : (long long)h->cfg.num_envs * h->M;
    k_set_state<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(h->st, h->cfg.num_envs, h->cfg.num_agents, h->M, env_mask_dev,
                                                                  agent_f32_dev, agent_u32_dev, env_i32_dev,
                                                                  h->M > 0 ? obst_xy_dev : nullptr);
    QS_CUDA(cudaGetLastError());
    h->launches += 1;
    note_async(h, (cudaStream_t)stream, false);
    return QS_OK;
}

extern "C" int qs_read_episode_stats(QsHandle* h, int32_t* env_stats_dev, float* agent_stats_dev, void* stream) {
    if (!h) return fail(QS_ERR_INVALID_ARG, "null argument");
    QS_CUDA(cudaSetDevice(h->device));
    const long long n1 = h->A, n2 = (long long)h->cfg.num_envs * QS_NUM_ENV_STATS;
    k_read_stats<<<blocks_for(n1 > n2 ? n1 : n2), 256, 0, (cudaStream_t)stream>>>(h->st, h->cfg.num_envs, h->cfg.num_agents,
                                                                                   env_stats_dev, agent_stats_dev);
    QS_CUDA(cudaGetLastError());
    h->launches += 1;
    note_async(h, (cudaStream_t)stream, false);
    return QS_OK;
}
