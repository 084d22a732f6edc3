#!/usr/bin/env python
"""bench.py — env agent-steps/s of the vectorised QuadSwarm env step on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--config c2|c3|c4] [--impl reference]

One "step" = one control step (2 physics sub-steps + collisions + observations, auto-reset included) of every env
of the workload.  Default workload = BASELINE.json configs[2] ("c3"): 8 drones x 4096 envs PER GPU (weak scaling),
12 pillars, K=2 neighbour obs, floor obs, downwash — the configuration the north-star target is quoted on.

JSON keys (one line, rank 0):
  value        whole-job agent-steps/s with actions/observations resident in HBM; one kernel launch per control step,
               replayed from a CUDA graph; inputs (action ring) and outputs (observation rollout ring) are larger than L2.
  e2e          the same metric through the reference-facing call with HOST numpy buffers (qs_step_host: H2D actions,
               kernel, D2H observations / rewards / dones inside the timed region).
  roofline     algorithmic bytes per launch (292 + 4 D + 8 M / N per agent-step) / mean launch time.
  cpu_baseline the UNMODIFIED reference (oracle/_ref, numba path) on one host core, bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # BASELINE configs -> (E per GPU, env kwargs, scenario, reward coefficients)
    'c2': dict(E=1024, kw=dict(num_agents=8, neighbor_visible_num=6, obs_repr='xyz_vxyz_R_omega'),
               mode='static_same_goal', rew=dict(quadcol_bin=5.0, quadcol_bin_smooth_max=10.0),
               desc='8 drones x 1024 envs, obstacle-free, K=6, static_same_goal'),
    'c3': dict(E=4096, kw=dict(num_agents=8, neighbor_visible_num=2, obs_repr='xyz_vxyz_R_omega_floor',
                               use_obstacles=True, obst_density=0.2, obst_size=0.6, obst_spawn_area=(8.0, 8.0),
                               use_downwash=True),
               mode='o_random', rew=dict(quadcol_bin=5.0, quadcol_bin_smooth_max=4.0, quadcol_bin_obst=5.0),
               desc='8 drones x 4096 envs, 12 pillars (8x8 m, density 0.2, size 0.6), K=2, floor obs, downwash, o_random'),
    'c4': dict(E=2048, kw=dict(num_agents=32, neighbor_visible_num=6, obs_repr='xyz_vxyz_R_omega'),
               mode='swarm_vs_swarm', rew=dict(quadcol_bin=5.0, quadcol_bin_smooth_max=10.0),
               desc='32 drones x 2048 envs, all-pairs collisions, K=6, swarm_vs_swarm'),
    'c5': dict(E=4096, kw=dict(num_agents=8, neighbor_visible_num=6, obs_repr='xyz_vxyz_R_omega'),
               mode='static_same_goal', rew=dict(quadcol_bin=5.0, quadcol_bin_smooth_max=10.0),
               desc='8 drones x 4096 envs per GPU, obstacle-free, K=6'),
}


def measured_traffic(config):
    """dram__bytes_read.sum + dram__bytes_write.sum of the step kernel per launch, from the committed ncu --set full
    capture of this workload (profiles/r01_traffic.json), or None."""
    p = os.path.join(ROOT, 'profiles', 'r01_traffic.json')
    try:
        return json.load(open(p)).get(config, {}).get('dram_bytes_per_launch')
    except Exception:
        return None


def hbm_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []
        self.proc = None

    def start(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), f'--query-gpu={q}', '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for s in self.samples:
            f = [x.strip() for x in s.split(',')]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def make_episode_tables(cfg, E, seed):
    """Synthetic episode tables of the workload's scenario, generated once on the host (scenarios.py)."""
    from quad_swarm_rl_b200.scenarios import create_scenario, obstacle_map_given_density
    kw = cfg['kw']
    N = kw['num_agents']
    use_obst = kw.get('use_obstacles', False)
    rs = np.random.RandomState(seed)
    goals = np.zeros((E, N, 3), np.float32)
    spawn = np.zeros((E, N, 3), np.float32)
    M = int(kw.get('obst_density', 0.2) * 64) if use_obst else 0
    obst = np.zeros((E, max(M, 1), 2), np.float32)
    sc = create_scenario(cfg['mode'], N, rng=rs, use_obstacles=use_obst)
    for e in range(E):
        if use_obst:
            obst_map, pos_arr, cells = obstacle_map_given_density(rs, kw['obst_spawn_area'], kw['obst_density'])
            sc.reset(obst_map=obst_map, cell_centers=cells)
            obst[e, :M] = np.asarray(pos_arr)[:, :2]
        else:
            sc.reset()
        goals[e] = sc.goals
        spawn[e] = sc.goals if sc.spawn_points is None else sc.spawn_points
    return goals, spawn, (obst[:, :M] if use_obst else None)


# ------------------------------------------------------------------------------------------
# reference CPU arm
# ------------------------------------------------------------------------------------------
def _ref_kwargs(cfg):
    kw = dict(cfg['kw'])
    rew = dict(pos=1.0, effort=0.05, spin=0.1, vel=0.0, crash=1.0, orient=1.0, yaw=0.0,
               quadcol_bin=0.0, quadcol_bin_smooth_max=0.0, quadcol_bin_obst=0.0)
    rew.update(cfg['rew'])
    return dict(num_agents=kw['num_agents'], neighbor_visible_num=kw['neighbor_visible_num'],
                obs_repr=kw['obs_repr'], use_obstacles=kw.get('use_obstacles', False),
                obst_density=kw.get('obst_density', 0.2), obst_size=kw.get('obst_size', 0.6),
                obst_spawn_area=kw.get('obst_spawn_area', (8.0, 8.0)), use_downwash=kw.get('use_downwash', False),
                quads_mode=cfg['mode'], rew_coeff=rew, use_numba=True)


def effective_cpus():
    """Host cores this process may actually use: cpu_count, capped by the affinity mask and the cgroup CPU quota."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        quota, period = open('/sys/fs/cgroup/cpu.max').read().split()[:2]                  # cgroup v2
        if quota != 'max':
            n = min(n, max(1, int(float(quota) / float(period))))
    except Exception:
        try:
            q = int(open('/sys/fs/cgroup/cpu/cpu.cfs_quota_us').read())                    # cgroup v1
            per = int(open('/sys/fs/cgroup/cpu/cpu.cfs_period_us').read())
            if q > 0:
                n = min(n, max(1, q // per))
        except Exception:
            pass
    return n


def _ref_worker(args, barrier=None):
    """One process = one reference env (how Sample Factory's rollout workers run it)."""
    cfg_name, n_warm, n_steps, seed = args
    from oracle import ref_harness as rh
    cfg = CONFIGS[cfg_name]
    kind = 'reference' if rh.reference_available() else 'port'
    N = cfg['kw']['num_agents']
    rs = np.random.RandomState(seed)
    if kind == 'reference':
        env = rh.make_reference_env(**_ref_kwargs(cfg))
        env.reset()
        step = lambda: env.step([a for a in rs.uniform(-1, 1, (N, 4)).astype(np.float32)])
    else:
        from tests.parity_util import cfg_to_oracle, make_tables
        from oracle import quadswarm_oracle as qo
        ocfg = cfg_to_oracle(cfg['kw'])
        t = make_tables(rs, 1, N, ocfg.num_obstacles, ocfg.use_obstacles, episodes=1)[0]
        src = qo.TableEpisodeSource([dict(goals=t['goals'][0], spawn=t['spawn'][0],
                                          obst_xy=None if t['obst'] is None else t['obst'][0])])
        env = qo.OracleEnv(ocfg, qo.PhiloxRng(seed), src)
        env.reset()
        step = lambda: env.step(rs.uniform(-1, 1, (N, 4)))
    for _ in range(n_warm):
        step()
    if barrier is not None:
        barrier.wait()                 # every process has imported, JIT-compiled and warmed up before any of them is timed
    t0 = time.perf_counter()
    for _ in range(n_steps):
        step()
    return time.perf_counter() - t0, kind


def _ref_proc(args, barrier, q):
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    try:
        q.put(_ref_worker(args, barrier))
    except Exception as e:             # never leave the others waiting at the barrier
        try:
            barrier.abort()
        except Exception:
            pass
        q.put((float('nan'), f'error: {e!r}'))


def cpu_baseline_single(cfg_name, budget_s=12.0):
    """Reference on ONE host core, bounded sample (about budget_s seconds)."""
    dt, kind = _ref_worker((cfg_name, 30, 50, 0))
    per = dt / 50
    n = int(max(100, min(5000, budget_s / per)))
    dt, kind = _ref_worker((cfg_name, 0, n, 1))
    N = CONFIGS[cfg_name]['kw']['num_agents']
    return dict(value=N * n / dt, unit='agent-steps/s', cores=1, kind=kind,
                sample=f'1 env x {n} control steps of workload {cfg_name}, random actions, 1 process (numba path)')


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation on all host cores (one env per process, the way Sample
    Factory's rollout workers run it).  The K bench steps are a BOUNDED SAMPLE of the workload: every process advances
    its env by n_proc = clamp(K, 100, 2500) control steps in total (so the run ends within minutes whatever K is);
    the rate, not the step count, is what is compared."""
    import multiprocessing as mp
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    N = cfg['kw']['num_agents']
    P = effective_cpus()
    n_proc = int(min(max(args.steps, 100), 2500))
    n_warm = int(min(max(args.warmup, 3), 50))
    for k in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS', 'NUMBA_NUM_THREADS'):
        os.environ.setdefault(k, '1')  # one env per process, one thread per process (inherited by the workers)
    ctx = mp.get_context('spawn')
    barrier, q = ctx.Barrier(P), ctx.Queue()
    procs = [ctx.Process(target=_ref_proc, args=((args.config, n_warm, n_proc, 100 + r), barrier, q)) for r in range(P)]
    for pr in procs:
        pr.start()
    res = [q.get() for _ in procs]
    for pr in procs:
        pr.join()
    bad = [r for r in res if not (r[0] == r[0])]
    if bad:
        raise RuntimeError(f'reference worker failed: {bad[0][1]}')
    wall = max(r[0] for r in res)
    kind = res[0][1]
    value = P * N * n_proc / wall
    line = {
        'impl': 'reference', 'metric': 'env agent-steps/sec', 'value': value, 'unit': 'agent-steps/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * wall / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': f"{args.config}: {cfg['desc']}",
                   'note': f'bounded sample: {P} processes (= usable host cores of {os.cpu_count()} logical) x 1 env each x {n_proc} control steps (numba path), timed after all processes warmed up'},
        'cpu_baseline': {'value': value, 'unit': 'agent-steps/s', 'cores': P, 'kind': kind,
                         'sample': f'{P} processes x 1 env x {n_proc} control steps of workload {args.config}'},
        'e2e': {'value': value, 'unit': 'agent-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------
def run_cuda_arm(args):
    import torch
    import torch.distributed as dist
    from quad_swarm_rl_b200.engine import QuadSwarmEngine

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)

    cfg = CONFIGS[args.config]
    E = args.envs or cfg['E']
    kw = cfg['kw']
    N = kw['num_agents']
    # Episodes are generated ON THE DEVICE inside the reset path of the kernels (o_random: pillars, spawn and goal cells;
    # static_same_goal / swarm_vs_swarm: goal formations), and swarm_vs_swarm's goal swaps every 4-6 s happen inside the
    # step kernel: no host work per episode or per tick.  --host-tables uploads host-generated tables once instead.
    dev_scn = None if args.host_tables else cfg['mode']
    # Optional (--groups G > 1): the E envs of this GPU are stepped as G independent blocks, each with its own chain of
    # kernel launches and no edge between the chains (double-buffered sampling, as Sample Factory runs its rollout
    # workers).  Every env still advances one control step per bench step; results do not depend on the grouping (global
    # env ids key the RNG: tests/test_gpu_api.py).  Measured gain on c3: 9.17 -> 8.93 us/step with 2 or 4 groups; the
    # default stays 1 group = one kernel launch per control step.
    groups = max(1, args.groups)
    while E % groups:
        groups -= 1
    Eg = E // groups
    engs = []
    for gi in range(groups):
        e_ = QuadSwarmEngine(num_envs=Eg, seed=args.seed, device=local_rank, env_id_offset=rank * E + gi * Eg,
                             rew_coeff=cfg['rew'], ep_time=args.ep_time, device_scenario=dev_scn, **kw)
        if dev_scn is None:
            goals, spawn, obst = make_episode_tables(cfg, Eg, seed=1000 + rank * 64 + gi)
            e_.set_next_episode(goals, spawn, obst)
        e_.reset()
        engs.append(e_)
    eng = engs[0]
    A, D, M = E * N, eng.D, eng.M

    # synthetic inputs: i.i.d. U(-1,1)^4 actions, pre-generated ring larger than L2; observation rollout ring
    pow2 = lambda x: 1 << int(np.ceil(np.log2(max(1, x))))
    R_act = max(8, min(1024, pow2(140e6 / (A * 16))))
    R_obs = max(8, min(1024, pow2(140e6 / (A * D * 4))))
    g = torch.Generator(device=dev)
    g.manual_seed(args.seed * 1000 + rank)
    act_ring = (torch.rand((R_act, E, N, 4), device=dev, generator=g) * 2 - 1).contiguous()
    obs_ring = torch.empty((R_obs, E, N, D), device=dev)
    rew_ring = torch.empty((R_obs, E, N), device=dev)
    done_ring = torch.empty((R_obs, E, N), dtype=torch.uint8, device=dev)
    metrics = torch.zeros(64, device=dev)

    counter = [0]
    side = [torch.cuda.Stream(device=dev) for _ in range(groups - 1)]

    def one_step():
        """one control step of every env: one kernel launch per env group, each on its own stream"""
        k = counter[0]
        cur = torch.cuda.current_stream(dev)
        for gi, e_ in enumerate(engs):
            sl = slice(gi * Eg, (gi + 1) * Eg)
            with torch.cuda.stream(cur if gi == 0 else side[gi - 1]):
                e_.step(act_ring[k % R_act, sl], obs_out=obs_ring[k % R_obs, sl], rewards_out=rew_ring[k % R_obs, sl],
                        dones_out=done_ring[k % R_obs, sl])
        counter[0] = k + 1

    def fork():
        cur = torch.cuda.current_stream(dev)
        for s_ in side:
            s_.wait_stream(cur)

    def join():
        cur = torch.cuda.current_stream(dev)
        for s_ in side:
            cur.wait_stream(s_)

    # CUDA graph of G consecutive steps (G divides both rings' periods so replays stay consistent)
    # graph of G consecutive steps; G is a power of two (ring indices baked into the graph stay periodic) and no larger
    # than the number of timed steps, so that short runs are still graph-replayed
    G = max(R_act, R_obs)
    while G > max(1, args.steps):
        G //= 2
    stream = torch.cuda.Stream(device=dev)
    graph = None
    with torch.cuda.stream(stream):
        fork()
        for _ in range(3):
            one_step()
        join()
        stream.synchronize()
        if not args.no_graph and G >= 4:
            counter[0] = 0
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                fork()
                for _ in range(G):
                    one_step()
                join()
            counter[0] = 0

        def run_steps(n):
            done = 0
            if graph is not None:
                while n - done >= G and counter[0] % G == 0:
                    graph.replay()
                    done += G
                    counter[0] += G
            if done < n:
                fork()
                while done < n:
                    one_step()
                    done += 1
                join()

        def sync_all():
            stream.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        run_steps(max(3, args.warmup))
        # realign to a graph boundary so that the timed region is mostly graph replays
        if graph is not None and counter[0] % G != 0:
            run_steps(G - counter[0] % G)
        sync_all()
        launches0 = sum(e_.launch_count for e_ in engs)
        clocks = ClockSampler(local_rank)
        if rank == 0:
            clocks.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        n_done = 0
        chunk = G if graph is not None else 100
        while n_done < args.steps:
            n = min(chunk, args.steps - n_done) if world > 1 else args.steps - n_done
            run_steps(n)
            n_done += n
            if world > 1:
                # optional cross-GPU metrics gather (north star: NCCL only for this): one small all-reduce per chunk
                metrics[0] = rew_ring[(counter[0] - 1) % R_obs].sum()
                dist.all_reduce(metrics)
        ev1.record(stream)
        sync_all()
        ms = ev0.elapsed_time(ev1)
        clk = clocks.stop() if rank == 0 else None
        launches = sum(e_.launch_count for e_ in engs) - launches0
        if graph is not None:
            launches = args.steps * groups             # replayed launches are not seen by the host-side counter
        if any(e_.handover_timeouts for e_ in engs):   # overlapping step grids must never have lost a hand-over
            raise RuntimeError("a per-block hand-over between step grids timed out: results of this run are invalid")
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())

    # ---- e2e: reference-facing call with HOST buffers (one engine holding all E envs of this GPU)
    for e_ in engs:
        e_.close()
    eng = QuadSwarmEngine(num_envs=E, seed=args.seed, device=local_rank, env_id_offset=rank * E, rew_coeff=cfg['rew'],
                          ep_time=args.ep_time, device_scenario=dev_scn, **kw)
    if dev_scn is None:
        goals, spawn, obst = make_episode_tables(cfg, E, seed=1000 + rank)
        eng.set_next_episode(goals, spawn, obst)
    eng.reset()
    n_e2e = max(10, min(args.steps, args.e2e_steps))
    # page-locked host buffers (numpy views of pinned torch tensors): the DMA engine reads / writes them directly
    a_pin = torch.empty((8, E, N, 4), dtype=torch.float32).pin_memory()
    a_pin.copy_(torch.from_numpy(np.random.RandomState(5 + rank).uniform(-1, 1, (8, E, N, 4)).astype(np.float32)))
    a_host = a_pin.numpy()
    obs_h = torch.zeros((E, N, D), dtype=torch.float32).pin_memory().numpy()
    rew_h = torch.zeros((E, N), dtype=torch.float32).pin_memory().numpy()
    done_h = torch.zeros((E, N), dtype=torch.uint8).pin_memory().numpy()
    for k in range(3):
        eng.step_host(a_host[k % 8], obs_h, rew_h, done_h)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(n_e2e):
        eng.step_host(a_host[k % 8], obs_h, rew_h, done_h)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    # the same loop with explicit cudaMemcpyAsync H2D / D2H around the kernel (QS_ZERO_COPY=0) instead of the default, in
    # which the kernel reads / writes the mapped page-locked host buffers itself; reported beside it
    os.environ['QS_ZERO_COPY'] = '0'
    for k in range(3):
        eng.step_host(a_host[k % 8], obs_h, rew_h, done_h)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(n_e2e):
        eng.step_host(a_host[k % 8], obs_h, rew_h, done_h)
    torch.cuda.synchronize()
    e2e_copy_s = time.perf_counter() - t0
    os.environ.pop('QS_ZERO_COPY', None)
    if world > 1:
        t = torch.tensor([e2e_s, e2e_copy_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s, e2e_copy_s = float(t[0].item()), float(t[1].item())

    # ---- extras (rank 0, single GPU): the same kernel (a) T steps per launch (qs_rollout: env block stays in registers,
    #      every observation still written) and (b) at a 4x larger batch, where the GPU has enough warps to fill its
    #      issue slots.  They explain the headline (which is latency-bound at 1.7 warps per SM sub-partition); they are
    #      not the headline.
    extra = {}
    if rank == 0 and world == 1 and not args.no_extras:
        def _events(fn, reps):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); [fn() for _ in range(reps)]; e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) * 1e-3 / reps
        T = 64
        acts = act_ring[:T].contiguous()
        o = torch.empty((T, E, N, D), device=dev); r = torch.empty((T, E, N), device=dev)
        d_ = torch.empty((T, E, N), dtype=torch.uint8, device=dev)
        eng.rollout(acts, obs_out=o, rewards_out=r, dones_out=d_)
        sec = _events(lambda: eng.rollout(acts, obs_out=o, rewards_out=r, dones_out=d_), 8)
        b_alg_ = 292 + 4 * D + (8.0 * M / N if M else 0.0)
        extra['rollout'] = {'steps_per_launch': T, 'us_per_step': sec / T * 1e6, 'agent_steps_per_s': A * T / sec,
                            'roofline_frac': b_alg_ * A * T / sec / 1e9 / hbm_peak()[0],
                            'note': 'qs_rollout: T control steps per launch, all observations written'}
        del o, r, d_
        E4 = 4 * E
        big = QuadSwarmEngine(num_envs=E4, seed=args.seed, device=local_rank, rew_coeff=cfg['rew'], ep_time=args.ep_time,
                              device_scenario=dev_scn, **kw)
        if dev_scn is None:
            g4, s4, o4 = make_episode_tables(cfg, min(E4, 512), seed=77)
            rep = E4 // min(E4, 512)
            big.set_next_episode(np.tile(g4, (rep, 1, 1)), np.tile(s4, (rep, 1, 1)), None if o4 is None else np.tile(o4, (rep, 1, 1)))
        big.reset()
        Rb = 8
        ab = (torch.rand((Rb, E4, N, 4), device=dev) * 2 - 1).contiguous()
        ob = torch.empty((Rb, E4, N, D), device=dev); rb = torch.empty((Rb, E4, N), device=dev)
        db = torch.empty((Rb, E4, N), dtype=torch.uint8, device=dev)
        st2 = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(st2):
            for k in range(Rb):
                big.step(ab[k], obs_out=ob[k], rewards_out=rb[k], dones_out=db[k])
            st2.synchronize()
            gb = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gb, stream=st2):
                for k in range(Rb):
                    big.step(ab[k], obs_out=ob[k], rewards_out=rb[k], dones_out=db[k])
            gb.replay(); st2.synchronize()
            sec = _events(gb.replay, 40) / Rb
        extra['large_batch'] = {'envs': E4, 'agents': E4 * N, 'us_per_step': sec * 1e6, 'agent_steps_per_s': E4 * N / sec,
                                'roofline_frac': b_alg_ * E4 * N / sec / 1e9 / hbm_peak()[0],
                                'note': 'same kernel, one launch per control step, 4x the envs of the headline workload'}
        big.close()
        del ab, ob, rb, db

    if rank == 0:
        value = world * A * args.steps / (ms * 1e-3)
        peak, peak_src = hbm_peak()
        b_alg = 292 + 4 * D + (8.0 * M / N if M else 0.0)
        per_launch_bytes = b_alg * A
        launch_s = ms * 1e-3 / args.steps
        achieved = per_launch_bytes / launch_s / 1e9
        cpu = cpu_baseline_single(args.config) if (world == 1 and not args.no_cpu_baseline) else None
        line = {
            'metric': 'env agent-steps/sec', 'value': value, 'unit': 'agent-steps/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f"{args.config}: {cfg['desc']}", 'envs_per_gpu': E, 'drones': N, 'obs_dim': D,
                       'agents_per_gpu': A, 'ep_len': eng.ep_len,
                       'episodes': 'generated on the device at every auto-reset' if dev_scn else 'host-generated tables, uploaded once',
                       'l2': f'inputs larger than L2: action ring {R_act} x {A * 16 / 1e6:.2f} MB, observation rollout ring '
                             f'{R_obs} x {A * D * 4 / 1e6:.2f} MB; env state ({A * 192 / 1e6:.1f} MB) is L2-resident by nature',
                       'launch': (('one kernel per control step' if groups == 1 else
                                   f'{groups} independent env blocks of {Eg} envs (double-buffered sampling), one kernel per block per control step')
                                  + (f', CUDA graph of {G} steps' if graph is not None else '')),
                       'parallelism': f'dp{world} (envs sharded, no step-time collective)'},
            'clocks': clk,
            'e2e': {'value': world * A * n_e2e / e2e_s, 'unit': 'agent-steps/s', 'h2d_bytes_per_step': A * 16,
                    'd2h_bytes_per_step': A * (4 * D + 4 + 1), 'steps': n_e2e,
                    'explicit_copies_value': world * A * n_e2e / e2e_copy_s,
                    'note': 'qs_step_host with page-locked numpy buffers, stream sync every step.  value: the kernel reads the '
                            'actions from and writes obs/rewards/dones to the mapped host buffers itself (zero-copy: the bytes '
                            'listed cross PCIe inside the timed region, no separate copy launches); explicit_copies_value: '
                            'cudaMemcpyAsync H2D actions, step kernel, cudaMemcpyAsync D2H obs/rewards/dones (QS_ZERO_COPY=0)'},
            'gpu_launches': int(launches),
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': measured_traffic(args.config) if E == cfg['E'] else None, 'peak_source': peak_src, 'alg_bytes_per_agent_step': b_alg,
                         'alg_bytes_per_launch': per_launch_bytes / groups, 'launch_us': launch_s * 1e6,
                         'note': f'{groups} concurrent launches of {Eg} envs per control step; achieved = bytes of all of them / step time'},
            'cpu_baseline': cpu,
        }
        line.update(extra)
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100000)
    ap.add_argument('--warmup', type=int, default=1024)
    ap.add_argument('--impl', default='cuda', choices=['cuda', 'reference'])
    ap.add_argument('--config', default='c3', choices=sorted(CONFIGS))
    ap.add_argument('--envs', type=int, default=0, help='envs per GPU (default: the config\'s)')
    ap.add_argument('--ep-time', type=float, default=15.0)
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--e2e-steps', type=int, default=300)
    ap.add_argument('--no-extras', action='store_true', help='skip the rollout / large-batch explanatory measurements')
    ap.add_argument('--groups', type=int, default=1, help='independent env blocks per GPU, each with its own launch chain')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--host-tables', action='store_true', help='use host-generated episode tables even where a device generator exists')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_cuda_arm(args)


if __name__ == '__main__':
    main()
