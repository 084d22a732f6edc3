"""Batched (device-tensor) versions of the wrappers around the env step.

`QuadrotorEnvMultiBatched` steps E envs per launch and hands CUDA tensors to the sampler; the reference's per-env
wrappers (`swarm_rl/env_wrappers/reward_shaping.py`, `gym_art/quadrotor_multi/quad_experience_replay.py`) walk Python
lists of dicts per agent, which at thousands of envs would dominate the step.  The classes here keep the same
semantics per env but hold their state as tensors on the GPU and work with masks:

* `BatchedRewardShaping`   reward-coefficient annealing, cumulative reward terms, true_reward, action statistics and the
                           `episode_extra_stats` keys, aggregated over the envs that finished (reward_shaping.py:52-123 +
                           quadrotor_multi.py:626-718);
* `BatchedExperienceReplay` collision-event replay: a checkpoint of every env every 0.5 s, the checkpoint from 1.5 s
                           before a collision goes into the env's 20-slot buffer, a finished env restarts from a buffered
                           event with probability p (quad_experience_replay.py:66-209) — snapshots are rows of
                           `qs_get_state`, replaying is `qs_set_state` with an env mask.

PyTorch is used for the bookkeeping tensors only; the env state never leaves the device.
"""
import torch

from . import _lib as L

T_POS, T_ACTION, T_CRASH, T_ORIENT, T_SPIN, T_QUADCOL, T_PROX, T_OBST = range(8)      # QS_TERM_* of include/quadswarm.h


class BatchedRewardShaping:
    """reward_shaping.py:19-123 for E envs at once.  step() returns (obs, rewards, dones, truncated, infos) like the env;
    when episodes end, infos['episode_extra_stats'] holds the reference's keys averaged over the agents of the envs
    that finished (Sample Factory averages them over episodes anyway) and infos['true_reward'] the per-agent values."""

    def __init__(self, env, reward_shaping_scheme=None, annealing=None):
        self.env = env
        self.engine = env.engine
        self.reward_shaping_scheme = reward_shaping_scheme
        self.annealing = annealing
        self.training_info = {}                      # Sample Factory writes approx_total_training_steps here
        self.reward_shaping_updated = True
        E, N, dev = env.num_envs, env.num_agents_per_env, self.engine.device
        self.cum_raw = torch.zeros((E, N, L.QS_NUM_TERMS), device=dev)       # rewraw_* sums of the running episode
        self.cum_rew = torch.zeros((E, N, L.QS_NUM_TERMS), device=dev)       # rew_* sums (coefficients in force per step)
        self.act_sum = torch.zeros((E, N, 4), device=dev)
        self.act_sq = torch.zeros((E, N, 4), device=dev)
        self.steps = torch.zeros((E,), device=dev)
        self.num_agents = env.num_agents

    def __getattr__(self, name):
        return getattr(self.env, name)

    def reset(self, **kw):
        out = self.env.reset(**kw)
        for t in (self.cum_raw, self.cum_rew, self.act_sum, self.act_sq, self.steps):
            t.zero_()
        return out

    def _coeff_vector(self):
        c = self.engine.rew_coeff
        # weights of the raw terms, in QS_TERM_* order; proximity is delivered already weighted
        return torch.tensor([c['pos'], c['effort'], c['crash'], c['orient'], c['spin'], c['quadcol_bin'], 1.0,
                             c['quadcol_bin_obst']], device=self.engine.device)

    def step(self, actions):
        env, eng = self.env, self.engine
        if self.reward_shaping_updated and self.reward_shaping_scheme:
            for key, weight in self.reward_shaping_scheme['quad_rewards'].items():
                eng.rew_coeff[key] = weight                       # pushed to the device before the next launch
            self.reward_shaping_updated = False
        coeff = self._coeff_vector()
        a = torch.as_tensor(actions, dtype=torch.float32, device=eng.device).reshape(env.num_envs, env.num_agents_per_env, 4)
        obs, rew, term, trunc, infos = env.step(a, with_terms=True)
        terms = eng.rew_terms
        self.cum_raw += terms
        self.cum_rew += terms * coeff
        self.act_sum += a
        self.act_sq += a * a
        self.steps += 1
        done_env = term.view(env.num_envs, -1)[:, 0]
        if bool(done_env.any()):                                  # one host sync per step, as any sampler needs for `dones`
            infos = dict(infos)
            infos.update(self._finish(done_env))
        return obs, rew, term, trunc, infos

    def _finish(self, done_env):
        env, eng = self.env, self.engine
        idx = torch.nonzero(done_env).flatten()
        raw, rw = self.cum_raw[idx], self.cum_rew[idx]
        true_reward = raw[..., T_POS] + 1000.0 * raw[..., T_QUADCOL]                      # reward_shaping.py:80-84
        steps = self.steps[idx].clamp(min=1).view(-1, 1, 1)
        a_mean = self.act_sum[idx] / steps
        a_std = (self.act_sq[idx] / steps - a_mean * a_mean).clamp(min=0).sqrt()
        stats = {
            'rewraw_main': true_reward.mean(), 'rewraw_pos': raw[..., T_POS].mean(), 'rewraw_action': raw[..., T_ACTION].mean(),
            'rewraw_crash': raw[..., T_CRASH].mean(), 'rewraw_orient': raw[..., T_ORIENT].mean(),
            'rewraw_spin': raw[..., T_SPIN].mean(), 'rewraw_quadcol': raw[..., T_QUADCOL].mean(),
            'rew_main': rw[..., T_POS].mean(), 'rew_pos': rw[..., T_POS].mean(), 'rew_action': rw[..., T_ACTION].mean(),
            'rew_crash': rw[..., T_CRASH].mean(), 'rew_orient': rw[..., T_ORIENT].mean(), 'rew_spin': rw[..., T_SPIN].mean(),
            'rew_quadcol': rw[..., T_QUADCOL].mean(), 'rew_proximity': rw[..., T_PROX].mean(),
        }
        if env.use_obstacles:
            stats['rewraw_quadcol_obstacle'] = raw[..., T_OBST].mean()
            stats['rew_quadcol_obstacle'] = rw[..., T_OBST].mean()
        for k in range(4):
            stats[f'z_action{k}_mean'] = a_mean[..., k].mean()
            stats[f'z_action{k}_std'] = a_std[..., k].mean()
        # statistics latched by the kernel at the episode end (quadrotor_multi.py:626-718)
        es, ags = eng.episode_stats()
        es, ags = es[idx].float(), ags[idx]
        names = L.ENV_STAT_KEYS
        for k in range(7):
            stats[names[k]] = es[:, k].mean()
        if env.use_obstacles:
            for k in range(7, 11):
                stats[names[k]] = es[:, k].mean()
        for k, col in (('1s', 0), ('3s', 1), ('5s', 2)):
            stats[f'distance_to_goal_{k}'] = ags[..., col].mean()
        fl = ags[..., 3].long()
        no_col_agent, no_col_obst, reached = (fl & 1) != 0, (fl & 2) != 0, (fl & 4) != 0
        col_flag = no_col_agent & no_col_obst
        stats['metric/agent_success_rate'] = (col_flag & reached).float().mean()
        stats['metric/agent_deadlock_rate'] = (col_flag & ~reached).float().mean()
        stats['metric/agent_col_rate'] = 1.0 - col_flag.float().mean()
        stats['metric/agent_neighbor_col_rate'] = 1.0 - no_col_agent.float().mean()
        stats['metric/agent_obst_col_rate'] = 1.0 - no_col_obst.float().mean()
        keys = list(stats)
        vals = torch.stack([stats[k].float() for k in keys]).cpu().numpy()                # ONE device -> host copy
        out = {k: float(v) for k, v in zip(keys, vals)}
        # per-scenario copies of the headline keys (reward_shaping.py:95-98, quadrotor_multi.py:680-718)
        scn_ids = es[:, names.index('scenario')].long()
        for sid in torch.unique(scn_ids).tolist():
            name = L.SCENARIO_NAMES.get(int(sid)) or env.quads_mode
            m = scn_ids == sid
            sub = torch.stack([rw[m][..., T_POS].mean(), rw[m][..., T_CRASH].mean(), es[m][:, 1].mean(),
                               ags[m][..., 0].mean()]).cpu().numpy()
            out[f'Scenario_{name}/rew_pos'], out[f'Scenario_{name}/rew_crash'] = float(sub[0]), float(sub[1])
            out[f'{name}/num_collisions'], out[f'{name}/distance_to_goal_1s'] = float(sub[2]), float(sub[3])
        approx = self.training_info.get('approx_total_training_steps', 0)
        out['z_approx_total_training_steps'] = approx
        if self.annealing:                                       # linear from 0 to the final value (reward_shaping.py:110-118)
            for sched in self.annealing:
                eng.rew_coeff[sched.coeff_name] = min(sched.final_value * approx / sched.anneal_env_steps, sched.final_value)
                out[f'z_anneal_{sched.coeff_name}'] = eng.rew_coeff[sched.coeff_name]
        for t in (self.cum_raw, self.cum_rew, self.act_sum, self.act_sq):
            t[idx] = 0
        self.steps[idx] = 0
        return {'episode_extra_stats': out, 'true_reward': true_reward, 'done_envs': idx}


class BatchedExperienceReplay:
    """quad_experience_replay.py:66-209 for E envs at once (each env keeps its own 20-slot buffer, as each wrapped env
    of the reference does).  Episodes of replayed envs start at the tick of their snapshot, so envs stop running in
    lock-step; everything here is per env and masked."""

    CP_EVERY = 50            # 0.5 s checkpoints (quad_experience_replay.py:17-21)
    STEPS_AGO = 3            # the checkpoint from 1.5 s before the collision (:87,:157)
    KEEP = 6                 # 3 s of checkpoints (:84)
    COOLDOWN = 500           # at most one event per 5 s (:154)
    MAX_REPLAYS = 10         # an event is dropped after 10 replays (:56-57)

    def __init__(self, env, replay_buffer_sample_prob=0.75, buffer_size=20, always_active=False, seed=0):
        if env.device_scenario is None:
            raise ValueError("BatchedExperienceReplay needs device-side scenarios (host scenario objects are not snapshotted)")
        self.env, self.engine = env, env.engine
        self.p = float(replay_buffer_sample_prob)
        E, dev = env.num_envs, self.engine.device
        self.E, self.B = E, buffer_size
        st = self.engine.get_state()
        self._keys = [k for k, v in st.items() if v is not None]
        D = self.engine.D
        self.ring = {k: torch.zeros((self.KEEP,) + tuple(st[k].shape), dtype=st[k].dtype, device=dev) for k in self._keys}
        self.ring_obs = torch.zeros((self.KEEP, E, env.num_agents_per_env, D), device=dev)
        self.ring_pos = torch.zeros(E, dtype=torch.long, device=dev)          # next slot to write
        self.ring_cnt = torch.zeros(E, dtype=torch.long, device=dev)
        self.buf = {k: torch.zeros((buffer_size,) + tuple(st[k].shape), dtype=st[k].dtype, device=dev) for k in self._keys}
        self.buf_obs = torch.zeros((buffer_size, E, env.num_agents_per_env, D), device=dev)
        self.buf_valid = torch.zeros((buffer_size, E), dtype=torch.bool, device=dev)
        self.buf_replayed = torch.zeros((buffer_size, E), dtype=torch.long, device=dev)
        self.buf_pos = torch.zeros(E, dtype=torch.long, device=dev)
        self.tick = torch.zeros(E, dtype=torch.long, device=dev)
        self.saved = torch.zeros(E, dtype=torch.bool, device=dev)             # saved_in_replay_buffer of the running episode
        self.last_added = torch.full((E,), -10 ** 9, dtype=torch.long, device=dev)
        # can_drones_fly (quadrotor_multi.py:281-287): fewer than one floor crash per episode over >= 10 episodes
        self.active = torch.full((E,), bool(always_active), dtype=torch.bool, device=dev)
        self.crash_hist = torch.zeros((100, E), device=dev)
        self.crash_n = torch.zeros(E, dtype=torch.long, device=dev)
        self.crash_now = torch.zeros(E, device=dev)
        self.gen = torch.Generator(device=dev)
        self.gen.manual_seed(seed)
        self.episode_counter = 0
        self.replayed_events = 0

    def __getattr__(self, name):
        return getattr(self.env, name)

    def reset(self, **kw):
        out = self.env.reset(**kw)
        self.tick.zero_(); self.saved.zero_(); self.ring_cnt.zero_(); self.last_added.fill_(-10 ** 9)
        return out

    def _ar(self):
        return torch.arange(self.E, device=self.engine.device)

    def step(self, actions, with_terms=True):
        env, eng = self.env, self.engine
        obs, rew, term, trunc, infos = env.step(actions, with_terms=True)      # collisions are read from the raw terms
        E, N = self.E, env.num_agents_per_env
        obs3 = obs.view(E, N, -1)
        done = term.view(E, N)[:, 0]
        terms = eng.rew_terms
        self.tick += 1
        self.crash_now += terms[..., T_CRASH].sum(dim=1) * eng.rew_coeff['crash'] / N     # infos[0]['rewards']['rew_crash'] analogue
        running = ~done
        # 1. checkpoints every 0.5 s of envs whose episode is not a replay
        cp = running & self.active & ~self.saved & (self.tick % self.CP_EVERY == 0)
        state = None
        if bool(cp.any()):
            state = eng.get_state()
            idx = torch.nonzero(cp).flatten()
            slot = self.ring_pos[idx]
            for k in self._keys:
                self.ring[k][slot, idx] = state[k][idx]
            self.ring_obs[slot, idx] = obs3[idx]
            self.ring_pos[idx] = (slot + 1) % self.KEEP
            self.ring_cnt[idx] = (self.ring_cnt[idx] + 1).clamp(max=self.KEEP)
        # 2. a collision after the grace period stores the checkpoint from 1.5 s earlier
        col = (terms[..., T_QUADCOL] < 0).any(dim=1)
        if env.use_obstacles:
            col |= (terms[..., T_OBST] < 0).any(dim=1)
        ev = running & col & self.active & ~self.saved & (self.tick > 150) & (self.tick - self.last_added > self.COOLDOWN) \
            & (self.ring_cnt >= self.STEPS_AGO)
        if bool(ev.any()):
            idx = torch.nonzero(ev).flatten()
            src = (self.ring_pos[idx] - self.STEPS_AGO) % self.KEEP
            # first free slot of the env's buffer, else round-robin (quad_experience_replay.py:36-45)
            free = ~self.buf_valid[:, idx]
            first_free = torch.argmax(free.int(), dim=0)
            dst = torch.where(free.any(dim=0), first_free, self.buf_pos[idx])
            for k in self._keys:
                self.buf[k][dst, idx] = self.ring[k][src, idx]
            self.buf_obs[dst, idx] = self.ring_obs[src, idx]
            self.buf_valid[dst, idx] = True
            self.buf_replayed[dst, idx] = 0
            self.buf_pos[idx] = (dst + 1) % self.B
            self.saved[idx] = True
            self.last_added[idx] = self.tick[idx]
        # 3. finished envs: replay a buffered event with probability p, else keep the fresh episode the kernel started
        if bool(done.any()):
            obs = self._new_episodes(done, obs3).view(obs.shape)
            infos = dict(infos)
            infos['replay'] = {'replay/replay_rate': self.replayed_events / max(self.episode_counter, 1),
                               'replay/replay_buffer_size': float(self.buf_valid.sum(dim=0).float().mean())}
        return obs, rew, term, trunc, infos

    def _new_episodes(self, done, obs3):
        eng = self.engine
        idx = torch.nonzero(done).flatten()
        n = idx.numel()
        self.episode_counter += n
        # can_drones_fly bookkeeping (quadrotor_multi.py:356-359)
        slot = self.crash_n[idx] % 100
        self.crash_hist[slot, idx] = self.crash_now[idx]
        self.crash_n[idx] += 1
        cnt = self.crash_n[idx].clamp(max=100)
        mean = self.crash_hist[:, idx].sum(dim=0) / cnt.clamp(min=1)
        self.active[idx] |= (cnt >= 10) & (mean.abs() < 1)
        self.crash_now[idx] = 0
        # fresh-episode defaults
        self.tick[idx] = 0
        self.saved[idx] = False
        self.ring_cnt[idx] = 0
        self.last_added[idx] = -10 ** 9
        have = self.buf_valid[:, idx].any(dim=0)
        u = torch.rand(n, device=eng.device, generator=self.gen)
        rep = have & self.active[idx] & (u < self.p)
        if not bool(rep.any()):
            return obs3
        ridx = idx[rep]
        w = self.buf_valid[:, ridx].float().t()                                  # [n_rep, B]
        pick = torch.multinomial(w, 1, generator=self.gen).flatten()
        self.replayed_events += int(ridx.numel())
        cur = eng.get_state()
        new = {}
        for k in self._keys:
            t = cur[k].clone()
            t[ridx] = self.buf[k][pick, ridx]
            new[k] = t
        # quad_experience_replay.py:188-190: accurate per-replay collision statistics
        for c in (0, 1, 7, 8):
            new['env_i32'][ridx, 4 + c] = 0
        # counters the snapshot must not rewind: the RNG step counter and the episode index stay those of the live env
        new['env_i32'][ridx, 1] = cur['env_i32'][ridx, 1]
        new['env_i32'][ridx, 3] = cur['env_i32'][ridx, 3]
        new['env_i32'][ridx, 4 + L.QS_NUM_ENV_STATS + 16] = cur['env_i32'][ridx, 4 + L.QS_NUM_ENV_STATS + 16]   # episode number
        mask = torch.zeros(self.E, dtype=torch.uint8, device=eng.device)
        mask[ridx] = 1
        eng.set_state(new, env_mask=mask)
        obs3 = obs3.clone()
        obs3[ridx] = self.buf_obs[pick, ridx]
        self.tick[ridx] = new['env_i32'][ridx, 0].long()
        self.saved[ridx] = True                                                   # a replayed episode is not checkpointed again
        self.buf_replayed[pick, ridx] += 1
        self.buf_valid[pick, ridx] &= self.buf_replayed[pick, ridx] < self.MAX_REPLAYS
        return obs3
