"""``BinaryHologramEnv`` -- the reference's gymnasium environment on the CUDA engine.

Same constructor, ``reset``/``step`` signatures, observation/action spaces and
reward semantics as the reference (env.py:37-260, env_1024_24.py:95-186,
env_1024_24_128.py:100-201, env_group.py:90-320); the arithmetic of the reward
(tt.simulate + abs**2 + mean + tt.relativeLoss, env.py:170-174) runs in
``libbholo_b200.so``.  A step no longer re-uploads and re-propagates the whole
stack: the engine keeps the fields of every frame resident in HBM and scores a
flip with the incremental delta kernel.

Behaviours of SURVEY.md appendix B marked "P" are preserved (early return of a
rejected step, recon_image showing the rejected flip, >= vs > accept rules, the
two bonus constants, state_record counting every attempt).  Documented fixes:
B-13 (the RGB step implements DBS_1024_24.py:324-363), B-15 (a seed can be
injected), B-16 (no empty_cache/gc in reset).
"""
from __future__ import annotations

import time
from typing import Callable, Iterable, Optional, Sequence

import numpy as np

from . import spaces
from .engine import (HoloEngine, RULE_ENV, RESULT_DTYPE, pinned_empty, OBS_PINNED_HOST, OBS_CONTEXT,
                     OBS_COMMITTED_ONLY, OBS_SYNC)

RW = 800                                   # env.py:29
WL_MONO = (515e-9,)                        # env.py:124
WL_RGB = (638e-9, 515e-9, 450e-9)          # env_1024_24.py:135-138


def _to_numpy(x) -> np.ndarray:
    if isinstance(x, np.ndarray):
        return x
    if hasattr(x, "detach"):               # torch tensor
        return x.detach().cpu().numpy()
    return np.asarray(x)


def goal_bonus(success_ratio: float, const: float) -> float:
    """env.py:230-235 (const -595.2) / env.py:249-254 (const -595.24)."""
    return (1828.57 * (success_ratio ** 3) - 3733.33 * (success_ratio ** 2)
            + 2800 * success_ratio + const)


def importance_reward_table(psnr_changes: Sequence[float], verbose: bool = False) -> np.ndarray:
    """env_group.py:121-141: rank of each candidate -> degree-5 polynomial reward."""
    n = len(psnr_changes)
    step_poly = np.array([10000, 9000, 8000, 5000, 2500, 1])
    rewards_poly = np.array([-0.5, -0.48, -0.45, -0.35, 0, 1])
    poly = np.poly1d(np.polyfit(step_poly, rewards_poly, len(step_poly) - 1))
    if verbose:                                           # env_group.py:128-129
        print("Polynomial Reward Function Equation:")
        print(poly)
    order = np.argsort(psnr_changes)
    x_val = 10000 - (10000 - 1) * (np.arange(n) / (n - 1))
    ranks = np.zeros(n)
    ranks[order] = poly(x_val)
    return ranks


class BinaryHologramEnv(spaces.Env):
    """Drop-in for the reference's ``BinaryHologramEnv``.

    Positional arguments are the reference's (env.py:38).  Keyword-only
    arguments select the variant:

    IPS, CH        image side and frame count (env.py:27-28; 1024/24 for RGB)
    wl             one wavelength per colour group (env_1024_24.py:135-147)
    crop_margin    env_1024_24_128.py: simulate the centre (IPS-2m)^2 window
    reward_mode    "psnr" (env.py:188) or "group" (env_group.py:254-255)
    recon_obs      "eager" (default, the reference's behaviour, env.py:176-181): obs["recon_image"]
                   is current after every step -- the changed colour plane is written into a pinned
                   host buffer by one kernel (bh_recon_batch); "device": the same, but the
                   observation is a zero-copy view of device memory (``__cuda_array_interface__``)
                   for a policy on the same GPU; "lazy": only on ``refresh_recon()``
    engine/env_index  share one multi-environment engine (vectorised envs)
    """

    metadata = {"render_modes": []}

    def __init__(self, target_function: Callable, trainloader: Iterable, max_steps=10000,
                 T_PSNR=30, T_steps=1, T_PSNR_DIFF=0.1, *, IPS: int = 256, CH: int = 8,
                 wl: Sequence[float] = WL_MONO, crop_margin: int = 0, reward_mode: str = "psnr",
                 recon_obs: str = "eager", device: int = 0, pad: int = 1, relative: bool = True,
                 method: str = "asm", verbose: bool = True, num_samples: int = 10000,
                 engine: Optional[HoloEngine] = None, env_index: int = 0,
                 rng: Optional[np.random.Generator] = None, resync_every: int = 1024,
                 action_mode: str = "discrete"):
        super().__init__()
        self.IPS, self.CH, self.wl = int(IPS), int(CH), tuple(wl)
        self.G = len(self.wl)
        self.crop_margin = int(crop_margin)
        self.Nsim = self.IPS - 2 * self.crop_margin
        self.reward_mode, self.recon_obs = reward_mode, recon_obs
        self.device, self.pad, self.relative, self.method = device, pad, relative, method
        self.verbose, self.num_samples = verbose, int(num_samples)
        self.rng = rng
        self.resync_every = int(resync_every)

        img_shape = (1, self.G, self.IPS, self.IPS)
        self.observation_space = spaces.Dict({           # env.py:42-48
            "state_record": spaces.Box(low=0, high=1, shape=(1, CH, IPS, IPS), dtype=np.int8),
            "state": spaces.Box(low=0, high=1, shape=(1, CH, IPS, IPS), dtype=np.int8),
            "pre_model": spaces.Box(low=0, high=1, shape=(1, CH, IPS, IPS), dtype=np.float32),
            "recon_image": spaces.Box(low=0, high=1, shape=(1, self.G, self.Nsim, self.Nsim),
                                      dtype=np.float32),
            "target_image": spaces.Box(low=0, high=1, shape=img_shape, dtype=np.float32),
        })
        self.num_pixels = CH * IPS * IPS                  # env.py:51-52
        self.action_mode = action_mode
        if action_mode == "multidiscrete":                # env_md.py:54,160: [channel, row, col]
            self.action_space = spaces.MultiDiscrete([CH, IPS, IPS])
        else:
            self.action_space = spaces.Discrete(self.num_pixels)

        self.target_function = target_function
        self.trainloader = trainloader
        self.max_steps, self.T_PSNR = max_steps, T_PSNR
        self.T_steps, self.T_PSNR_DIFF = T_steps, T_PSNR_DIFF

        self.state = self.state_record = self.observation = None
        self.steps = self.psnr_sustained_steps = self.flip_count = None
        self.next_print_thresholds = 0
        self.total_start_time = None
        self.target_image = self.target_image_np = None
        self.initial_psnr = self.previous_psnr = None
        self.max_psnr_diff = float("-inf")
        self.data_iter = iter(self.trainloader)           # env.py:83
        self.episode_num_count = 0
        self.current_file = None

        self._engine = engine
        self._own_engine = engine is None
        self._e = int(env_index)
        self._z = self._dx = None
        self._commits = 0
        self._last_candidate = -1
        self._res = np.empty(1, dtype=RESULT_DTYPE)
        self._act = np.empty(1, dtype=np.int64)
        self._eid = np.array([self._e], dtype=np.int32)
        self._ptrs = (self._eid.ctypes.data, self._act.ctypes.data, self._res.ctypes.data)
        self._recon_buf = None                            # pinned, allocated with the engine
        self._vec = None                                  # set by HologramVecEnv: it refreshes all envs at once
        if recon_obs not in ("eager", "lazy", "device"):
            raise ValueError(f"recon_obs must be 'eager', 'lazy' or 'device', not {recon_obs!r}")

    # ------------------------------------------------------------------
    def _ensure_engine(self, z: float, dx: float):
        if self._engine is not None:
            if self._engine.z == z and self._engine.dx == dx:
                return
            if not self._own_engine:
                raise ValueError("shared engine was built for a different z / pixel pitch")
            self._engine.close()
        self._engine = HoloEngine(self.Nsim, self.CH, self.wl, n_env=1, device=self.device, dx=dx,
                                  z=z, pad=self.pad, relative=self.relative, method=self.method)

    @property
    def engine(self) -> HoloEngine:
        return self._engine

    # -- observation buffer ---------------------------------------------
    def _attach_obs(self):
        """Make ``_recon_buf`` the (1, G, Nsim, Nsim) block obs["recon_image"] aliases."""
        if self._vec is not None:
            self._vec._attach_env_obs(self)
        elif self.recon_obs == "device":
            self._recon_buf = self._engine.recon_device_block(0)[self._e]
        elif not isinstance(self._recon_buf, np.ndarray):
            self._recon_buf = pinned_empty((1, self.G, self.Nsim, self.Nsim), np.float32)

    def _publish_recon(self, stepped: bool, reset: bool = False):
        """Bring obs["recon_image"] up to date after a step (env.py:176-181) or a reset."""
        if self._vec is not None:
            if reset:
                self._vec._env_was_reset(self)
            return
        if self.recon_obs == "lazy":
            if not stepped:
                self._engine.recon(self._e, -1, out=self._recon_buf[0])
            return
        flags = OBS_SYNC | (0 if stepped else OBS_COMMITTED_ONLY)
        if self.recon_obs == "device":
            self._engine.recon_batch(1, 0, OBS_CONTEXT, 0, flags, env_ids_ptr=self._ptrs[0])
        else:
            self._engine.recon_batch(1, self._recon_buf.ctypes.data, OBS_PINNED_HOST, 0, flags,
                                     env_ids_ptr=self._ptrs[0])

    def _crop(self, a: np.ndarray) -> np.ndarray:
        m = self.crop_margin
        return a if m == 0 else a[..., m:-m, m:-m]

    def _next_target(self):
        try:                                              # env.py:95-102
            target, cur = next(self.data_iter)
        except StopIteration:
            if self.verbose:
                print("\033[40;93m[INFO] Reached the end of dataset. Restarting from the beginning.\033[0m")
            self.data_iter = iter(self.trainloader)
            target, cur = next(self.data_iter)
        return target, cur

    # -- env.py:90-152 ---------------------------------------------------
    def reset(self, seed=None, options=None, z=2e-3, pixel_pitch=7.56e-6, crop_margin=None):
        if crop_margin is not None and int(crop_margin) != self.crop_margin:
            # env_1024_24_128.py:100 takes the margin at reset; a new margin is a new FFT side, i.e. a new engine
            if not self._own_engine:
                raise ValueError("crop_margin of a shared engine is fixed at construction (it sets the FFT side)")
            if self._engine is not None:
                self._engine.close()
                self._engine = None
            self.crop_margin = int(crop_margin)
            self.Nsim = self.IPS - 2 * self.crop_margin
            self._recon_buf = None
            if self._vec is not None:
                raise ValueError("crop_margin of a vectorised env is fixed at construction")
            self.observation_space.spaces["recon_image"] = spaces.Box(
                low=0, high=1, shape=(1, self.G, self.Nsim, self.Nsim), dtype=np.float32)
        if seed is not None:
            self.rng = np.random.default_rng(seed)
        self._ensure_engine(float(z), float(pixel_pitch))
        self._attach_obs()
        self.episode_num_count += 1

        self.target_image, self.current_file = self._next_target()
        if self.verbose:
            print(f"\033[40;93m[Episode Start] Currently using dataset file: {self.current_file}, "
                  f"Episode count: {self.episode_num_count}\033[0m")
        tgt_in = self.target_image
        if hasattr(tgt_in, "cuda"):                       # env.py:106
            try:
                tgt_in = tgt_in.cuda(self.device)
                self.target_image = tgt_in
            except Exception:
                pass
        self.target_image_np = np.ascontiguousarray(_to_numpy(tgt_in), dtype=np.float32)
        model_output = self.target_function(tgt_in)       # env.py:109-111
        self.observation = np.ascontiguousarray(_to_numpy(model_output), dtype=np.float32)

        self.max_psnr_diff = float("-inf")
        self.steps = 0
        self.flip_count = 0
        self.psnr_sustained_steps = 0
        self.state = (self.observation >= 0.5).astype(np.int8)       # env.py:120
        self.state_record = np.zeros_like(self.state)                 # env.py:121

        eng, e = self._engine, self._e
        eng.set_target(e, self._crop(self.target_image_np[0]))
        eng.load_state(e, self._crop(self.state[0]))                  # env.py:123-128
        self.initial_psnr, mse, _ = eng.metrics(e)                    # env.py:131-132
        self.previous_psnr = self.initial_psnr
        self._commits = 0
        self._last_candidate = -1
        self._publish_recon(False, reset=True)

        if self.reward_mode == "group":                               # env_group.py:190-199
            t0 = time.time()
            self.psnr_change_list, self.importance_ranks, pos = self._calculate_pixel_importance()
            self._psnr_change_arr = np.asarray(self.psnr_change_list, dtype=np.float64)
            if self.verbose:
                print(f"\nTime taken for psnr_change_list: {time.time() - t0:.2f} seconds")
            self.T_PSNR_DIFF = pos / 4
            if self.verbose:
                print(f"\033[94m[Dynamic Threshold] T_PSNR_DIFF set to: {self.T_PSNR_DIFF:.6f}\033[0m")

        obs = self._obs()
        if self.verbose:                                              # env.py:142-145
            print(f"\033[92mInitial PSNR: {self.initial_psnr:.6f}\033[0m"
                  f"\nInitial MSE: {mse:.6f}\033[0m")
        self.next_print_thresholds = [self.initial_psnr + i * 0.01 for i in range(1, 21)]
        self.total_start_time = time.time()
        return obs, {"state": self.state}

    def clone_from(self, leader: "BinaryHologramEnv"):
        """Start an episode from ``leader``'s freshly reset state without propagating again.

        Group rollouts (env_group.py + GRPO-style groups): M members share one target and one
        initial hologram; the device state is copied (bh_clone_env), the host mirrors are copied,
        and for the rank-table reward the leader's 10 000-candidate table is shared.
        """
        if leader._engine is not self._engine:
            raise ValueError("clone_from needs envs that share one engine")
        self.episode_num_count += 1
        self.target_image, self.current_file = leader.target_image, leader.current_file
        self.target_image_np, self.observation = leader.target_image_np, leader.observation
        self.state, self.state_record = leader.state.copy(), np.zeros_like(leader.state)
        self.max_psnr_diff = float("-inf")
        self.steps = self.flip_count = self.psnr_sustained_steps = 0
        self.initial_psnr = self.previous_psnr = leader.initial_psnr
        self._commits, self._last_candidate = 0, -1
        self._engine.clone_env(leader._e, self._e)
        self._attach_obs()
        self._publish_recon(False, reset=True)
        if self.reward_mode == "group":
            self.psnr_change_list, self.importance_ranks = leader.psnr_change_list, leader.importance_ranks
            self._psnr_change_arr, self.T_PSNR_DIFF = leader._psnr_change_arr, leader.T_PSNR_DIFF
        self.next_print_thresholds = [self.initial_psnr + i * 0.01 for i in range(1, 21)]
        self.total_start_time = time.time()
        return self._obs(), {"state": self.state}

    def _obs(self):
        return {"state_record": self.state_record,                   # env.py:135-140
                "state": self.state,
                "pre_model": self.observation,
                "recon_image": self._recon_buf,
                "target_image": self.target_image_np}

    def refresh_recon(self) -> np.ndarray:
        """Materialise obs["recon_image"] (the last evaluated flip included, env.py:176-181).
        Only needed with ``recon_obs="lazy"``; the other modes keep the observation current."""
        if self.recon_obs == "lazy":
            self._engine.recon(self._e, self._last_candidate, out=self._recon_buf[0])
        return self._recon_buf

    # -- env_group.py:90-143 --------------------------------------------
    def _calculate_pixel_importance(self):
        rng = self.rng if self.rng is not None else np.random
        if hasattr(rng, "integers"):
            actions = rng.integers(0, self.num_pixels, size=self.num_samples)
        else:
            actions = np.array([rng.randint(self.num_pixels) for _ in range(self.num_samples)])
        self.importance_actions = actions
        sim_actions, inside = self._map_actions(actions)
        psnr = np.full(actions.shape[0], self.initial_psnr, dtype=np.float64)
        if inside.any():
            psnr[inside] = self._engine.eval_flips(sim_actions[inside], env=self._e)
        changes = psnr - self.initial_psnr
        positive = float(np.sum(changes[changes > 0]))
        return list(changes), importance_reward_table(changes, self.verbose), positive

    def _map_actions(self, actions: np.ndarray):
        """Full-grid action -> engine (cropped-grid) action; inside = within the window."""
        a = np.asarray(actions, dtype=np.int64)
        if self.crop_margin == 0:
            return a, np.ones(a.shape, dtype=bool)
        n2 = self.IPS * self.IPS
        ch, pix = a // n2, a % n2
        r, c = pix // self.IPS - self.crop_margin, pix % self.IPS - self.crop_margin
        inside = (r >= 0) & (r < self.Nsim) & (c >= 0) & (c < self.Nsim)
        return ch * self.Nsim * self.Nsim + r * self.Nsim + c, inside

    # -- env.py:154-260 --------------------------------------------------
    def step(self, action, z=2e-3, pixel_pitch=7.56e-6):
        if np.ndim(action) == 1 and len(action) == 3:     # env_md.py:160 (channel, row, col)
            ch, r, c = (int(v) for v in action)
            if not (0 <= ch < self.CH and 0 <= r < self.IPS and 0 <= c < self.IPS):
                raise ValueError(f"action {action} outside MultiDiscrete([{self.CH}, {self.IPS}, {self.IPS}])")
            action = (ch * self.IPS + r) * self.IPS + c
        action = int(action)
        if not 0 <= action < self.num_pixels:
            raise ValueError(f"action {action} outside Discrete({self.num_pixels})")
        sim_action = action
        if self.crop_margin:
            sim, inside = self._map_actions(np.array([action]))
            if not inside[0]:
                # a pixel outside the simulated window leaves the reconstruction unchanged
                return self._finish_step(action, self.previous_psnr, True, -1)
            sim_action = int(sim[0])
        self._act[0] = sim_action
        self._engine.step_batch_ptrs(1, self._ptrs, RULE_ENV)
        res = self._res[0]
        return self._finish_step(action, float(res["psnr_after"]), bool(res["accept"]), sim_action)

    def _finish_step(self, action: int, psnr_after: float, accepted: bool, sim_action: int):
        IPS = self.IPS
        self.steps += 1
        channel = action // (IPS * IPS)
        pixel_index = action % (IPS * IPS)
        row, col = pixel_index // IPS, pixel_index % IPS
        self.state_record[0, channel, row, col] += 1                  # env.py:165
        self.flip_count += 1
        self._last_candidate = -1 if accepted else sim_action
        self._publish_recon(sim_action >= 0)                          # env.py:176-181
        obs = self._obs()

        psnr_change = psnr_after - self.previous_psnr                 # env.py:184-185
        psnr_diff = psnr_after - self.initial_psnr
        if self.reward_mode == "group":                               # env_group.py:254-255
            idx = int(np.argmin(np.abs(self._psnr_change_arr - psnr_change)))
            reward = float(self.importance_ranks[idx])
        else:
            reward = psnr_change * RW                                 # env.py:188

        if not accepted:                                              # env.py:191-196
            self.flip_count -= 1
            return obs, reward, False, False, {}
        self.state[0, channel, row, col] = 1 - self.state[0, channel, row, col]   # env.py:164
        self._commits += 1
        resynced_psnr = None
        if self.resync_every > 0 and self._commits % self.resync_every == 0:
            # the device decides the next flips against the re-propagated PSNR: follow it on the host
            # (env.py:184-196: a kept flip never has a negative change)
            self._engine.resync(self._e)
            resynced_psnr = self._engine.metrics(self._e)[0]

        self.max_psnr_diff = max(self.max_psnr_diff, psnr_diff)
        success_ratio = self.flip_count / self.steps if self.steps > 0 else 0

        def _block():
            dt = time.time() - self.total_start_time
            print(f"Step: {self.steps:<6} | Initial PSNR: {self.initial_psnr:.6f}"
                  f"\nPSNR After: {psnr_after:.6f} | Change: {psnr_change:.6f} | Diff: {psnr_diff:.6f}"
                  f"\nReward: {reward:.2f} | Success Ratio: {success_ratio:.6f} | Flip Count: {self.flip_count}"
                  f"\nFlip Pixel: Channel={channel}, Row={row}, Col={col}"
                  f"\nTime taken for this data: {dt:.2f} seconds")

        while self.next_print_thresholds and psnr_after >= self.next_print_thresholds[0]:
            self.next_print_thresholds.pop(0)                         # env.py:203-212
            if self.verbose:
                _block()
        self.previous_psnr = psnr_after if resynced_psnr is None else resynced_psnr   # env.py:214

        if psnr_diff >= self.T_PSNR_DIFF or (psnr_after >= self.T_PSNR and psnr_diff < 0.1):
            if self.verbose:
                _block()
            self.psnr_sustained_steps += 1                            # env.py:216-225
            if self.psnr_sustained_steps >= self.T_steps and psnr_diff >= self.T_PSNR_DIFF:
                if self.reward_mode == "group":                       # env_group.py:294-299
                    reward += 100 + (-200.0 / 1500.0) * (self.steps - 1000)
                else:
                    reward += goal_bonus(success_ratio, -595.2)       # env.py:230-235
        if self.steps >= self.max_steps:                              # env.py:237-254
            if self.verbose:
                _block()
            if self.reward_mode == "group":                           # env_group.py:311-315
                reward += 100 + (-200.0 / 1500.0) * (self.steps - 1000)
            else:
                reward += goal_bonus(success_ratio, -595.24)
        terminated = self.steps >= self.max_steps or self.psnr_sustained_steps >= self.T_steps
        truncated = self.steps >= self.max_steps                      # env.py:257-260
        return obs, reward, terminated, truncated, {}

    def close(self):
        if self._own_engine and self._engine is not None:
            self._engine.close()
            self._engine = None


# ---------------------------------------------------------------------------
# variants named after the reference modules
# ---------------------------------------------------------------------------
class BinaryHologramEnvRGB(BinaryHologramEnv):
    """env_1024_24.py: 1024^2, 24 frames = 3 colour thirds of 8 (638/515/450 nm)."""

    def __init__(self, target_function, trainloader, max_steps=10000, T_PSNR=30, T_steps=1,
                 T_PSNR_DIFF=0.1, **kw):
        kw.setdefault("IPS", 1024); kw.setdefault("CH", 24); kw.setdefault("wl", WL_RGB)
        super().__init__(target_function, trainloader, max_steps, T_PSNR, T_steps, T_PSNR_DIFF, **kw)


class BinaryHologramEnvRGBCrop(BinaryHologramEnvRGB):
    """env_1024_24_128.py: as RGB but the centre 896^2 window is simulated."""

    def __init__(self, target_function, trainloader, max_steps=10000, T_PSNR=30, T_steps=1,
                 T_PSNR_DIFF=0.1, **kw):
        kw.setdefault("crop_margin", 64)
        super().__init__(target_function, trainloader, max_steps, T_PSNR, T_steps, T_PSNR_DIFF, **kw)


class BinaryHologramEnvGroup(BinaryHologramEnv):
    """env_group.py: reward = rank table of 10 000 candidate flips scored at reset."""

    def __init__(self, target_function, trainloader, max_steps=10000, T_PSNR=30, T_steps=1,
                 T_PSNR_DIFF=0.1, **kw):
        kw.setdefault("reward_mode", "group")
        super().__init__(target_function, trainloader, max_steps, T_PSNR, T_steps, T_PSNR_DIFF, **kw)
