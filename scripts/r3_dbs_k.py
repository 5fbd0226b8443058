"""Greedy DBS at 1024^2 x 24 from a fresh random state: candidates/s and candidates consumed per iteration for fixed
speculation depths (k_spec) and the adaptive default (0).  One JSON line per depth."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import binary_hologram_reinforcement_learning_b200 as bh

N, F = 1024, 24
pre, tgt = bh.synthetic_problem(N, F, 3, 0)
state = (pre >= 0.5).astype(np.int8)
eng = bh.HoloEngine(N, F, bh.WL_RGB, n_env=1)
eng.set_target(0, tgt)
rng = np.random.default_rng(3)
order = rng.permutation(F * N * N)[:30000]
for k in (0, 1, 2, 3, 4, 6, 8, 12, 16):
    eng.load_state(0, state)
    eng.stream_sync()
    l0 = eng.launch_count
    t0 = time.perf_counter()
    acc, _, nacc, psnr = eng.dbs_run(order, env=0, k_spec=k, resync_every=0)
    dt = time.perf_counter() - t0
    iters = (eng.launch_count - l0) / 2
    print(json.dumps({"k_spec": k, "cand_per_s": round(order.size / dt), "us_per_iteration": round(1e6 * dt / iters, 2),
                      "consumed_per_iteration": round(order.size / iters, 2), "accept_rate": round(nacc / order.size, 4),
                      "psnr": psnr}), flush=True)
