"""Observation path at 1024^2 x 24, 8 vectorised envs (GPU): env steps/s of HologramVecEnv.step with
recon_obs = eager (pinned host, double buffered) / device / lazy, the PCIe rate of the eager path, and a
check of the eager observation against bh_get_recon."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import binary_hologram_reinforcement_learning_b200 as bh  # noqa: E402

N, F, G, E = 1024, 24, 3, 8
out = {}
for mode in ("eager", "device", "lazy"):
    loaders = [bh.SyntheticLoader(N, F, G, seeds=(i,)) for i in range(E)]

    def target_function(t):
        key = np.ascontiguousarray(t[0, 0, 0, :4], dtype=np.float32).tobytes()
        for ld in loaders:
            if key in ld._pre:
                return ld._pre[key][None]
        raise KeyError

    vec = bh.HologramVecEnv(E, target_function, loaders, max_steps=10 ** 9, T_PSNR_DIFF=1e9, IPS=N, CH=F,
                            wl=bh.WL_RGB, recon_obs=mode, verbose=False, seed=0)
    vec.reset()
    rng = np.random.default_rng(1)
    acts = rng.integers(0, F * N * N, size=(600, E), dtype=np.int64)
    for i in range(40):
        obs, *_ = vec.step(acts[i])
    if mode == "eager":       # the block equals the per-env reference path (committed I + rejected candidate)
        worst = 0.0
        for j in range(E):
            ref = vec.engine.recon(j, int(vec._last_cand[j]))
            worst = max(worst, float(np.abs(obs[j]["recon_image"][0] - ref).max()))
        out["eager_vs_get_recon_max_abs"] = worst
    n = 200 if mode != "lazy" else 500
    vec.engine.stream_sync()
    t0 = time.perf_counter()
    for i in range(40, 40 + n):
        vec.step(acts[i])
    vec.engine.stream_sync()
    dt = time.perf_counter() - t0
    out[mode + "_env_steps_per_s"] = E * n / dt
    out[mode + "_ms_per_vec_step"] = 1e3 * dt / n
    vec.close()
print(json.dumps(out))
