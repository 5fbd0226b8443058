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

from .engine import HoloEngine, RULE_ENV, RESULT_DTYPE
from .envs import BinaryHologramEnv, WL_MONO


class HologramVecEnv:
    def __init__(self, n_envs: int, target_function: Callable, trainloaders, max_steps=10000,
                 T_PSNR=30, T_steps=1, T_PSNR_DIFF=0.1, *, IPS=256, CH=8, wl: Sequence[float] = WL_MONO,
                 crop_margin=0, reward_mode="psnr", device=0, pad=1, relative=True, method="asm",
                 z=2e-3, pixel_pitch=7.56e-6, obs_mode="views", recon_obs="lazy", verbose=False,
                 seed: Optional[int] = None, resync_every: int = 1024, num_samples: int = 10000):
        self.num_envs = int(n_envs)
        self.obs_mode = obs_mode
        self.z, self.pixel_pitch = float(z), float(pixel_pitch)
        nsim = IPS - 2 * crop_margin
        self.engine = HoloEngine(nsim, CH, wl, n_env=self.num_envs, device=device, dx=pixel_pitch,
                                 z=z, pad=pad, relative=relative, method=method)
        if not isinstance(trainloaders, (list, tuple)):
            trainloaders = [trainloaders] * self.num_envs
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
        self._sim = np.empty(self.num_envs, dtype=np.int64)
        # episode statistics of finished episodes: reward, steps, flips, psnr0, psnr1
        self.episode_stats: List[np.ndarray] = []
        self._ep_reward = np.zeros(self.num_envs)

    # ------------------------------------------------------------------
    def _pack(self, obs_list):
        if self.obs_mode == "views":
            return obs_list
        return {k: np.stack([o[k] for o in obs_list]) for k in obs_list[0]}

    def reset(self):
        obs = [e.reset(z=self.z, pixel_pitch=self.pixel_pitch)[0] for e in self.envs]
        self._ep_reward[:] = 0
        return self._pack(obs)

    def step_async(self, actions):
        self._actions = np.asarray(actions, dtype=np.int64).reshape(self.num_envs)

    def step_wait(self):
        acts = self._actions
        envs = self.envs
        if envs[0].crop_margin == 0:
            self.engine.step_batch(acts, self._eids, RULE_ENV, out=self._res)
            inside = None
        else:
            sim, inside = envs[0]._map_actions(acts)
            if inside.all():
                self.engine.step_batch(sim, self._eids, RULE_ENV, out=self._res)
                inside = None
            elif inside.any():
                self._res[inside] = self.engine.step_batch(sim[inside], self._eids[inside], RULE_ENV)
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
                info["terminal_observation"] = obs
                info["TimeLimit.truncated"] = bool(trunc and not term)
                self.episode_stats.append(np.array(
                    [self._ep_reward[i], env.steps, env.flip_count, env.initial_psnr, env.previous_psnr]))
                self._ep_reward[i] = 0.0
                obs = env.reset(z=self.z, pixel_pitch=self.pixel_pitch)[0]
            obs_list.append(obs)
            infos.append(info)
        return self._pack(obs_list), rewards, dones, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        self.engine.close()
