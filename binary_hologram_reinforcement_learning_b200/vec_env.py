"""Vectorised hologram environments: E envs resident on one GPU, one batched launch per step.

The reference drives a single env through SB3's implicit ``DummyVecEnv``
(train-PPO.py:296-298; ``make_vec_env(lambda: env, n_envs=1)`` at
optimize_hyperparameter.py:317).  This class offers the same VecEnv protocol
(``reset`` / ``step_async`` / ``step_wait`` / ``step``, auto-reset with
``terminal_observation``) for E independent envs that share one engine, so the
E flip evaluations of a step are one ``bh_step_batch`` call (two kernel
launches) instead of E full re-simulations.
"""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional, Sequence

import numpy as np

from .engine import (HoloEngine, RULE_ENV, RESULT_DTYPE, VecBook, pinned_empty, DeviceArray, OBS_PINNED_HOST,
                     OBS_CONTEXT, OBS_COMMITTED_ONLY, OBS_SYNC)
from .envs import BinaryHologramEnv, WL_MONO, RW, goal_bonus


class ShardedLoader:
    """Items ``index, index + count, ...`` of a shared, re-iterable loader: E environments built on
    one ``trainloader`` walk disjoint images instead of E copies of the same episode."""

    def __init__(self, loader, index: int, count: int):
        self.loader, self.index, self.count = loader, int(index), int(count)

    def __iter__(self):
        n = 0
        for k, item in enumerate(self.loader):
            if k % self.count == self.index:
                n += 1
                yield item
        if n == 0:                       # fewer images than environments: share them
            yield from self.loader

    def __getattr__(self, name):         # e.g. target_function helpers hung on the loader
        return getattr(self.loader, name)


class HologramVecEnv:
    def __init__(self, n_envs: int, target_function: Callable, trainloaders, max_steps=10000,
                 T_PSNR=30, T_steps=1, T_PSNR_DIFF=0.1, *, IPS=256, CH=8, wl: Sequence[float] = WL_MONO,
                 crop_margin=0, reward_mode="psnr", device=0, pad=1, relative=True, method="asm",
                 z=2e-3, pixel_pitch=7.56e-6, obs_mode="views", recon_obs="eager", verbose=False,
                 seed: Optional[int] = None, resync_every: int = 1024, num_samples: int = 10000,
                 obs_buffers: int = 2):
        """``recon_obs``: "eager" (default; the reference returns the reconstruction of the evaluated
        flip on every step, env.py:176-181): after each step one kernel writes the changed colour
        planes of all E envs into a pinned host block, double buffered (``obs_buffers``) so the
        observation of step k stays valid while step k+1 runs; "device": same, zero-copy device
        views; "lazy": only ``refresh_recon(i)`` moves data.  One ``trainloader`` is sharded over the
        envs (env i sees items i, i+E, ...); pass a list for one loader per env."""
        self.num_envs = int(n_envs)
        self.obs_mode = obs_mode
        self.z, self.pixel_pitch = float(z), float(pixel_pitch)
        nsim = IPS - 2 * crop_margin
        self.engine = HoloEngine(nsim, CH, wl, n_env=self.num_envs, device=device, dx=pixel_pitch,
                                 z=z, pad=pad, relative=relative, method=method)
        if not isinstance(trainloaders, (list, tuple)):
            trainloaders = ([trainloaders] if self.num_envs == 1 else
                            [ShardedLoader(trainloaders, i, self.num_envs) for i in range(self.num_envs)])
        ss = np.random.SeedSequence(seed)
        self.envs: List[BinaryHologramEnv] = [
            BinaryHologramEnv(target_function, trainloaders[i], max_steps, T_PSNR, T_steps,
                              T_PSNR_DIFF, IPS=IPS, CH=CH, wl=wl, crop_margin=crop_margin,
                              reward_mode=reward_mode, recon_obs=recon_obs, device=device, pad=pad,
                              relative=relative, method=method, verbose=verbose, engine=self.engine,
                              env_index=i, rng=np.random.default_rng(c), resync_every=resync_every,
                              num_samples=num_samples)
            for i, c in enumerate(ss.spawn(self.num_envs))]
        self.observation_space = self.envs[0].observation_space
        self.action_space = self.envs[0].action_space
        self._actions = None
        self._res = np.empty(self.num_envs, dtype=RESULT_DTYPE)
        self._eids = np.arange(self.num_envs, dtype=np.int32)
        # the step path passes raw addresses: ndarray.ctypes costs more than a microsecond per use
        self._actions = np.zeros(self.num_envs, dtype=np.int64)
        self._ptrs = (self._eids.ctypes.data, self._actions.ctypes.data, self._res.ctypes.data)
        self._psnr_after = self._res["psnr_after"]
        self._no_event = bytes(self.num_envs)
        self._resync_every = int(resync_every)
        self._resync_wait = 0                  # steps until a re-propagation can be due (see _step_fast)
        self._sim = np.empty(self.num_envs, dtype=np.int64)
        # episode statistics of finished episodes: reward, steps, flips, psnr0, psnr1
        self.episode_stats: List[np.ndarray] = []
        self._ep_reward = np.zeros(self.num_envs)
        # vectorised bookkeeping (env.py:154-260 evaluated for all envs at once); used when no
        # per-step printing, cropping or rank-table reward is involved
        self._fast = (crop_margin == 0 and reward_mode in ("psnr", "group") and not verbose)
        self._obs_shape = (self.num_envs, 1, len(wl), nsim, nsim)
        self._n_obs = max(1, min(int(obs_buffers), 4))
        self._sub_eids = np.zeros(self.num_envs, dtype=np.int32)
        self._blocks_by_mode = {}
        self._cur = 0
        self._obs_cache = [None] * self.num_envs
        self.set_recon_obs(recon_obs)
        self._group = reward_mode == "group"
        self._changes = self._ranks = None      # (E, num_samples) tables of env_group.py:90-143
        self._sorted = [None] * self.num_envs
        E = self.num_envs
        self._steps = np.zeros(E, dtype=np.int64)
        self._flips = np.zeros(E, dtype=np.int64)
        self._prev = np.zeros(E)
        self._init = np.zeros(E)
        self._sustained = np.zeros(E, dtype=np.int64)
        self._ar = np.arange(E)
        self._state = self._record = None
        self._last_cand = np.full(E, -1, dtype=np.int64)
        self._tdiff = np.full(E, float(T_PSNR_DIFF))
        self._tpsnr = np.full(E, float(T_PSNR))
        self._maxsteps = np.full(E, int(max_steps), dtype=np.int64)
        self._rewards = np.zeros(E)
        self._change = np.zeros(E)
        self._diff = np.zeros(E)
        self._event = np.zeros(E, dtype=np.uint8)
        self._no_done = np.zeros(E, dtype=bool)
        self._book = None

    # ------------------------------------------------------------------
    def _pack(self, obs_list):
        if self.obs_mode == "views":
            return obs_list
        return {k: np.stack([o[k] for o in obs_list]) for k in obs_list[0]}

    def _adopt(self, i: int):
        """Re-home env i's host mirrors in the stacked arrays and load its scalars."""
        env = self.envs[i]
        if self._state is None:
            shape = (self.num_envs,) + env.state.shape
            self._state = np.zeros(shape, dtype=np.int8)
            self._record = np.zeros(shape, dtype=np.int8)
            self._state2d = self._state.reshape(self.num_envs, -1)
            self._record2d = self._record.reshape(self.num_envs, -1)
            b = VecBook()
            b.state, b.state_record = self._state.ctypes.data, self._record.ctypes.data
            b.stride = self._state2d.shape[1]
            b.prev_psnr, b.init_psnr = self._prev.ctypes.data, self._init.ctypes.data
            b.steps, b.flips = self._steps.ctypes.data, self._flips.ctypes.data
            b.t_psnr_diff, b.t_psnr = self._tdiff.ctypes.data, self._tpsnr.ctypes.data
            b.max_steps, b.reward_scale = self._maxsteps.ctypes.data, float(RW)
            b.rewards, b.psnr_change, b.psnr_diff = (self._rewards.ctypes.data, self._change.ctypes.data,
                                                     self._diff.ctypes.data)
            b.last_candidate, b.event = self._last_cand.ctypes.data, self._event.ctypes.data
            self._book = b
        self._state[i] = env.state
        self._record[i] = env.state_record
        env.state, env.state_record = self._state[i], self._record[i]
        self._steps[i] = env.steps
        self._flips[i] = env.flip_count
        self._resync_wait = 0
        self._prev[i] = env.previous_psnr
        self._init[i] = env.initial_psnr
        self._sustained[i] = env.psnr_sustained_steps
        self._last_cand[i] = -1
        self._tdiff[i], self._tpsnr[i], self._maxsteps[i] = env.T_PSNR_DIFF, env.T_PSNR, env.max_steps
        self._obs_cache[i] = env._obs()
        if self._group:
            if self._changes is None:
                self._changes = np.zeros((self.num_envs, env.num_samples))
                self._ranks = np.zeros((self.num_envs, env.num_samples))
            self._changes[i] = env._psnr_change_arr
            self._ranks[i] = env.importance_ranks
            # sorted view for the nearest-value lookup of env_group.py:254 (np.argmin semantics:
            # the lowest original index among the nearest values)
            order = np.argsort(env._psnr_change_arr, kind="stable")
            sv = env._psnr_change_arr[order]
            self._sorted[i] = (sv, order, np.searchsorted(sv, sv, side="left"))

    # -- observation blocks ------------------------------------------------
    def set_recon_obs(self, mode: str):
        """Select how obs["recon_image"] is kept ("eager" / "device" / "lazy"); may be switched between steps.
        Blocks: [E][1][G][nsim][nsim], written plane by plane by bh_recon_batch."""
        if mode not in ("eager", "device", "lazy"):
            raise ValueError(f"recon_obs must be 'eager', 'device' or 'lazy', not {mode!r}")
        self.recon_obs = mode
        self._obs_live = mode in ("eager", "device")
        for env in self.envs:
            env.recon_obs = mode
            env._vec = self if self._obs_live else None
        if not self._obs_live:
            for env in self.envs:
                if env._engine is not None and env.steps is not None:
                    env._recon_buf = None
                    env._attach_obs()
                    env.refresh_recon()
            self._refresh_obs_cache()
            return
        if mode not in self._blocks_by_mode:
            if mode == "device":
                blocks = [self.engine.recon_device_block(b) for b in range(self._n_obs)]
                ptrs = [0] * self._n_obs
            else:
                blocks = [pinned_empty(self._obs_shape, np.float32) for _ in range(self._n_obs)]
                ptrs = [b.ctypes.data for b in blocks]
            self._blocks_by_mode[mode] = (blocks, ptrs)
        self._blocks, self._block_ptrs = self._blocks_by_mode[mode]
        self._obs_kind = OBS_CONTEXT if mode == "device" else OBS_PINNED_HOST
        self._views = [[blk[i] for i in range(self.num_envs)] for blk in self._blocks]
        self._stale_all = [True] * self._n_obs          # a (re)selected block holds nothing current
        if self.envs[0].steps is not None:               # already reset: fill the current block now
            from .engine import OBS_FULL
            self.engine.recon_batch(self.num_envs, self._block_ptrs[self._cur], self._obs_kind, self._cur,
                                    OBS_SYNC | OBS_COMMITTED_ONLY | OBS_FULL, env_ids_ptr=self._eids.ctypes.data)
            self._stale_all[self._cur] = False
            for env in self.envs:
                self._attach_env_obs(env)
            self._refresh_obs_cache()

    def _refresh_obs_cache(self):
        for i, env in enumerate(self.envs):
            if self._obs_cache[i] is not None:
                self._obs_cache[i] = env._obs()

    def _attach_env_obs(self, env: BinaryHologramEnv):
        env._recon_buf = self._views[self._cur][env._e]

    def _env_was_reset(self, env: BinaryHologramEnv):
        """A reset / clone re-propagated env: write its committed reconstruction into the current block."""
        self._attach_env_obs(env)
        self.engine.recon_batch(1, self._block_ptrs[self._cur], self._obs_kind, self._cur,
                                OBS_SYNC | OBS_COMMITTED_ONLY, env_ids_ptr=env._ptrs[0])

    def _publish_step(self, n: int, eids_ptr: int = 0, rest: Optional[np.ndarray] = None):
        """After a step of n envs: advance to the next block and bring it up to date for every env."""
        b = self._cur = (self._cur + 1) % self._n_obs
        eng, ptr, kind = self.engine, self._block_ptrs[b], self._obs_kind
        full = 0
        if self._stale_all[b]:                        # first use of this block since it was (re)selected
            from .engine import OBS_FULL
            full, self._stale_all[b] = OBS_FULL, False
        if rest is not None and rest.size:            # envs without an engine step this time (cropped-out pixels)
            self._rest_eids = np.ascontiguousarray(rest, dtype=np.int32)
            if n:
                eng.recon_batch(n, ptr, kind, b, full, env_ids_ptr=eids_ptr)
            eng.recon_batch(int(rest.size), ptr, kind, b, OBS_SYNC | OBS_COMMITTED_ONLY | full,
                            env_ids_ptr=self._rest_eids.ctypes.data)
        else:
            eng.recon_batch(n, ptr, kind, b, OBS_SYNC | full, env_ids_ptr=eids_ptr)
        views = self._views[b]
        for i, env in enumerate(self.envs):
            env._recon_buf = views[i]
            oc = self._obs_cache[i]
            if oc is not None:
                oc["recon_image"] = views[i]

    def _terminal_obs(self, env: BinaryHologramEnv):
        out = {}
        for k, v in env._obs().items():
            if isinstance(v, DeviceArray):            # host copy of the final reconstruction
                v = self.engine.recon(env._e, env._last_candidate)[None]
            out[k] = np.array(v)
        return out

    def sync_envs(self):
        """Push the vectorised counters back into the per-env objects."""
        if not self._fast:
            return
        for i, env in enumerate(self.envs):
            env.steps, env.flip_count = int(self._steps[i]), int(self._flips[i])
            env.previous_psnr, env.psnr_sustained_steps = float(self._prev[i]), int(self._sustained[i])
            env._last_candidate = int(self._last_cand[i])

    def refresh_recon(self, i: int) -> np.ndarray:
        """obs["recon_image"] of env i, the last evaluated flip included (env.py:176-181)."""
        self.sync_envs()
        return self.envs[i].refresh_recon()

    def _reset_env(self, i: int):
        obs = self.envs[i].reset(z=self.z, pixel_pitch=self.pixel_pitch)[0]
        self._adopt(i)
        return self.envs[i]._obs()

    def reset(self):
        obs = [self._reset_env(i) for i in range(self.num_envs)]
        self._ep_reward[:] = 0
        return self._pack(obs)

    def reset_groups(self, group_size: int):
        """GRPO-style groups: envs [k*M, (k+1)*M) start from the reset state of env k*M.

        Only the leaders load a target, run the initial-hologram function, propagate and (for
        the rank-table reward) score their candidates; members are device-side clones.
        """
        if self.num_envs % group_size:
            raise ValueError("num_envs must be a multiple of group_size")
        obs = [None] * self.num_envs
        for lead in range(0, self.num_envs, group_size):
            obs[lead] = self._reset_env(lead)
            for m in range(lead + 1, lead + group_size):
                self.envs[m].clone_from(self.envs[lead])
                self._adopt(m)
                obs[m] = self.envs[m]._obs()
        self._ep_reward[:] = 0
        return self._pack(obs)

    def _nearest(self, i: int, x: float) -> int:
        """argmin(|psnr_change_list - x|) of env i in O(log n): same index as np.argmin."""
        sv, order, run_start = self._sorted[i]
        n = sv.shape[0]
        j = int(np.searchsorted(sv, x, side="left"))
        if j == 0:
            return int(order[0])
        if j == n:
            return int(order[run_start[n - 1]])
        dl, dr = abs(sv[j - 1] - x), abs(sv[j] - x)
        if dr < dl:
            return int(order[j])
        left = int(order[run_start[j - 1]])
        if dl < dr:
            return left
        return min(left, int(order[j]))

    def _step_fast(self):
        """env.py:154-260 for all envs with numpy; per-env Python only on episode events."""
        acts, envs, E = self._actions, self.envs, self.num_envs
        # scoring on the GPU + mirrors, counters, psnr_change, reward, event mask in one foreign call
        self.engine.vec_step_ptrs(E, self._ptrs, RULE_ENV, self._book)
        if self._obs_live:                                           # env.py:176-181
            self._publish_step(E, self._ptrs[0])
        res, psnr_after = self._res, self._psnr_after
        diff = self._diff
        if self._group:                                              # env_group.py:254-255
            rewards = np.empty(E)
            for i in range(E):
                rewards[i] = self._ranks[i, self._nearest(i, self._change[i])]
        else:
            rewards = self._rewards.copy()                           # env.py:188
        infos = [{} for _ in range(E)]
        if self._resync_every > 0:
            # an env is re-propagated when its kept-flip count reaches a multiple of resync_every.  Counts
            # grow by at most one per step, so nothing can be due before `_resync_wait` more steps:
            # the (numpy) test runs only then, not on every step
            self._resync_wait -= 1
            if self._resync_wait <= 0:
                due = (res["accept"] != 0) & (self._flips % self._resync_every == 0)
                for i in np.flatnonzero(due):
                    # the device now decides against the re-propagated PSNR: follow it on the host
                    self.engine.resync(int(i))
                    self._prev[i] = self.engine.metrics(int(i))[0]
                self._resync_wait = int((self._resync_every - self._flips % self._resync_every).min())
        if self._event.tobytes() == self._no_event:
            self._ep_reward += rewards
            return self._pack(self._obs_cache), rewards, self._no_done.copy(), infos
        dones = np.zeros(E, dtype=bool)
        event = self._event
        for i in np.flatnonzero(event):                              # env.py:216-260
            env = envs[i]
            ratio = self._flips[i] / self._steps[i]
            linear = 100 + (-200.0 / 1500.0) * (self._steps[i] - 1000)     # env_group.py:294-299
            if diff[i] >= env.T_PSNR_DIFF or (psnr_after[i] >= env.T_PSNR and diff[i] < 0.1):
                self._sustained[i] += 1
                if self._sustained[i] >= env.T_steps and diff[i] >= env.T_PSNR_DIFF:
                    rewards[i] += linear if self._group else goal_bonus(ratio, -595.2)
            if self._steps[i] >= env.max_steps:
                rewards[i] += linear if self._group else goal_bonus(ratio, -595.24)
            term = self._steps[i] >= env.max_steps or self._sustained[i] >= env.T_steps
            trunc = self._steps[i] >= env.max_steps
            if term or trunc:
                dones[i] = True
                self.sync_envs()
                infos[i] = {"terminal_observation": self._terminal_obs(env),
                            "TimeLimit.truncated": bool(trunc and not term)}
                self.episode_stats.append(np.array(
                    [self._ep_reward[i] + rewards[i], self._steps[i], self._flips[i], self._init[i],
                     self._prev[i]]))
                self._ep_reward[i] = -rewards[i]
                self._reset_env(i)
        self._ep_reward += rewards
        return self._pack(self._obs_cache), rewards, dones, infos

    def step_async(self, actions):
        self._actions[:] = np.asarray(actions).reshape(self.num_envs)

    def step_wait(self):
        if self._fast:
            return self._step_fast()
        acts = self._actions
        envs = self.envs
        if envs[0].crop_margin == 0:
            self.engine.step_batch(acts, self._eids, RULE_ENV, out=self._res)
            inside = None
            if self._obs_live:
                self._publish_step(self.num_envs, self._eids.ctypes.data)
        else:
            sim, inside = envs[0]._map_actions(acts)
            if inside.all():
                self.engine.step_batch(sim, self._eids, RULE_ENV, out=self._res)
                inside = None
                if self._obs_live:
                    self._publish_step(self.num_envs, self._eids.ctypes.data)
            else:
                m = int(inside.sum())
                if m:
                    self._sub_eids[:m] = self._eids[inside]
                    self._res[inside] = self.engine.step_batch(sim[inside], self._sub_eids[:m], RULE_ENV)
                if self._obs_live:
                    self._publish_step(m, self._sub_eids.ctypes.data, rest=self._eids[~inside])
            self._sim[:] = sim
        obs_list, infos = [], []
        rewards = np.empty(self.num_envs, dtype=np.float64)
        dones = np.zeros(self.num_envs, dtype=bool)
        res = self._res
        for i, env in enumerate(envs):
            if inside is None or inside[i]:
                sim_a = int(acts[i]) if env.crop_margin == 0 else int(self._sim[i])
                obs, r, term, trunc, info = env._finish_step(
                    int(acts[i]), float(res["psnr_after"][i]), bool(res["accept"][i]), sim_a)
            else:
                obs, r, term, trunc, info = env._finish_step(int(acts[i]), env.previous_psnr, True, -1)
            rewards[i] = r
            self._ep_reward[i] += r
            if term or trunc:
                dones[i] = True
                info = dict(info)
                info["terminal_observation"] = self._terminal_obs(env)
                info["TimeLimit.truncated"] = bool(trunc and not term)
                self.episode_stats.append(np.array(
                    [self._ep_reward[i], env.steps, env.flip_count, env.initial_psnr, env.previous_psnr]))
                self._ep_reward[i] = 0.0
                obs = self._reset_env(i)
            obs_list.append(obs)
            infos.append(info)
        return self._pack(obs_list), rewards, dones, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        self.engine.close()


def as_sb3_vec_env(vec: HologramVecEnv):
    """Wrap a :class:`HologramVecEnv` in Stable-Baselines3's ``VecEnv`` interface.

    The reference hands a single env to SB3, which wraps it in ``DummyVecEnv``
    (train-PPO.py:296-298).  With this adapter ``PPO("MultiInputPolicy", as_sb3_vec_env(vec))``
    trains on E GPU-resident envs whose steps are one batched launch.  Needs
    stable_baselines3 (not shipped in the build image); observations are stacked per key.
    """
    from stable_baselines3.common.vec_env import VecEnv

    class _HologramSB3VecEnv(VecEnv):
        def __init__(self, inner: HologramVecEnv):
            self.inner = inner
            inner.obs_mode = "stacked"
            super().__init__(inner.num_envs, inner.observation_space, inner.action_space)

        def reset(self):
            return self.inner.reset()

        def step_async(self, actions):
            self.inner.step_async(actions)

        def step_wait(self):
            obs, rewards, dones, infos = self.inner.step_wait()
            return obs, rewards.astype(np.float32), dones, infos

        def close(self):
            self.inner.close()

        def seed(self, seed=None):
            return [None] * self.num_envs

        def _targets(self, indices):
            idx = range(self.num_envs) if indices is None else ([indices] if isinstance(indices, int) else indices)
            return [self.inner.envs[i] for i in idx]

        def get_attr(self, attr_name, indices=None):
            self.inner.sync_envs()
            return [getattr(e, attr_name) for e in self._targets(indices)]

        def set_attr(self, attr_name, value, indices=None):
            for e in self._targets(indices):
                setattr(e, attr_name, value)

        def env_method(self, method_name, *args, indices=None, **kwargs):
            return [getattr(e, method_name)(*args, **kwargs) for e in self._targets(indices)]

        def env_is_wrapped(self, wrapper_class, indices=None):
            return [False for _ in self._targets(indices)]

    return _HologramSB3VecEnv(vec)
