#!/usr/bin/env python
"""Secondary measurements: one figure per BASELINE.json config (the bench line covers configs[1]).

    python scripts/bench_configs.py > profiles/<round>_configs.json
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import binary_hologram_reinforcement_learning_b200 as bh  # noqa: E402

out = {}


def problem(N, F, G, seed=0):
    pre, tgt = bh.synthetic_problem(N, F, G, seed)
    return pre, tgt, (pre >= 0.5).astype(np.int8)


# -- configs[0]: DBS.py greedy at 256^2 x 8 (DBS.py:242-294) --------------------------------------
N, F = 256, 8
pre, tgt, st = problem(N, F, 1)
eng = bh.HoloEngine(N, F, bh.WL_MONO)
eng.set_target(0, tgt); eng.load_state(0, st)
p0 = eng.metrics(0)[0]
order = np.random.default_rng(0).permutation(F * N * N)
eng.dbs_run(order[:4096])                                   # warm (graphs, tables)
eng.load_state(0, st)
t0 = time.perf_counter()
acc, _, nacc, fin = eng.dbs_run(order, resync_every=1024)   # the whole image: 524 288 candidates
dt = time.perf_counter() - t0
out["config0_DBS_256x8_greedy"] = {
    "candidates": int(order.size), "seconds": dt, "candidates_per_s": order.size / dt,
    "accepted": int(nacc), "psnr_initial": p0, "psnr_final": fin}
# env steps at 256^2 x 8 through the single-env API (eager recon copy, as the reference env)
ld = bh.SyntheticLoader(N, F, 1, seeds=(0,))
env = bh.BinaryHologramEnv(ld.target_function, ld, verbose=False, T_PSNR_DIFF=1e9, max_steps=10 ** 9)
env.reset()
acts = np.random.default_rng(1).integers(0, F * N * N, size=3000)
for a in acts[:200]:
    env.step(int(a))
t0 = time.perf_counter()
for a in acts[200:]:
    env.step(int(a))
dt = time.perf_counter() - t0
out["config0_env_256x8_single_env_steps_per_s_eager_recon"] = (len(acts) - 200) / dt
env.close(); eng.close()

# -- configs[2]/[3]: 896^2 x 24 cropped sweep (dbs-1024-1024-24-6464.py, DBS_1024_24-128.py) ------
N, F = 896, 24
pre, tgt, st = problem(N, F, 3, seed=1)
eng = bh.HoloEngine(N, F, bh.WL_RGB)
eng.set_target(0, tgt); eng.load_state(0, st)
p0 = eng.metrics(0)[0]
rng = np.random.default_rng(2)
cand = rng.permutation(F * N * N)
eng.eval_flips(cand[:256])
t0 = time.perf_counter()
for i in range(64):
    eng.eval_flips(cand[i * 128:(i + 1) * 128])            # 128 candidates per call
dt = time.perf_counter() - t0
out["config2_sweep_896x24_128_per_call_flip_evals_per_s"] = 64 * 128 / dt
t0 = time.perf_counter()
eng.eval_flips(cand[:65536])
out["config2_sweep_896x24_delta_kernel_flip_evals_per_s"] = 65536 / (time.perf_counter() - t0)
eng.sweep_all(0)
t0 = time.perf_counter()
pm = eng.sweep_all(0)                                      # all 19 267 584 candidates, host output
dt_map = time.perf_counter() - t0
t0 = time.perf_counter()
r = bh.sweep_engine(eng, 0, pre, cand, p0, psnr_map=pm)    # decile statistics on the host
dt_stats = time.perf_counter() - t0
eng.sweep_stats(pre, bh.OUTPUT_BINS, 0)                    # warm (first launch loads the kernel)
t0 = time.perf_counter()
att, imp, gn, _ = eng.sweep_stats(pre, bh.OUTPUT_BINS, 0)   # sweep + decile statistics on the device
dt_dev = time.perf_counter() - t0
assert np.array_equal(att, r["attempted"]) and np.array_equal(imp, r["improved"])
out["config3_full_sweep_896x24_device_stats"] = {
    "seconds_incl_pre_model_upload": dt_dev, "flip_evals_per_s_end_to_end": cand.size / dt_dev,
    "max_rel_gain_diff_vs_host": float(np.max(np.abs(gn - r["gains"]) / np.maximum(np.abs(r["gains"]), 1e-30)))}
out["config3_full_sweep_896x24"] = {
    "candidates": int(cand.size), "sweep_all_seconds_incl_d2h": dt_map, "host_decile_stats_seconds": dt_stats,
    "flip_evals_per_s_end_to_end": cand.size / (dt_map + dt_stats), "improving_fraction": r["flip_count"] / cand.size}
eng.close()

# -- configs[4]: env_group reset scoring + group rollouts (env_group.py:90-143) -------------------
N, F, E = 256, 8, 64
loaders = [bh.SyntheticLoader(N, F, 1, seeds=(i % 4,)) for i in range(E)]
tf = lambda t: next(l for l in loaders if np.ascontiguousarray(t[0, 0, 0, :4]).tobytes() in l._pre).target_function(t)
vec = bh.HologramVecEnv(E, tf, loaders, max_steps=10 ** 9, T_PSNR_DIFF=1e9, IPS=N, CH=F,
                        reward_mode="group", num_samples=10000, seed=0)
t0 = time.perf_counter()
vec.reset()                                                # 64 x 10 000 candidate scorings
dt = time.perf_counter() - t0
out["config4_group_reset_64_envs_seconds"] = dt
out["config4_group_reset_candidates_per_s"] = E * 10000 / dt
t0 = time.perf_counter()
vec.reset_groups(8)                                        # 8 leaders score, 56 members are device clones
out["config4_group_reset_8_groups_of_8_seconds"] = time.perf_counter() - t0
acts = np.random.default_rng(3).integers(0, F * N * N, size=(600, E))
for i in range(100):
    vec.step(acts[i])
t0 = time.perf_counter()
for i in range(100, 600):
    vec.step(acts[i])
dt = time.perf_counter() - t0
out["config4_group_rollout_64_envs_256x8_env_steps_per_s"] = 500 * E / dt
vec.close()

print(json.dumps(out, indent=1))
