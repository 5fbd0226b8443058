"""Seconds-long GPU sanity run of the Python step paths (single env, vectorised env with periodic
re-propagation, greedy DBS with the early stop) against the CPU oracle.  No torch, no pytest:
`python tests/gpu_sanity.py` on a B200.  Lives under tests/ because it uses the oracle as checker."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import binary_hologram_reinforcement_learning_b200 as bh  # noqa: E402
from oracle import hologram_oracle as O  # noqa: E402

t0 = time.time()
N, F = 64, 8
cfg = O.HoloConfig(N=N, F=F)
ld = bh.SyntheticLoader(N, F, 1, seeds=(3,))
env = bh.BinaryHologramEnv(ld.target_function, ld, IPS=N, CH=F, verbose=False, max_steps=10 ** 6, T_PSNR_DIFF=1e9)
env.reset()
pre, tgt = bh.synthetic_problem(N, F, 1, 3)
ref = O.OracleEnv(cfg, max_steps=10 ** 6, T_PSNR_DIFF=1e9)
ref.reset(pre, tgt)
rng = np.random.default_rng(0)
for _ in range(40):
    a = int(rng.integers(0, F * N * N))
    r = env.step(a)[1]
    r_ref = ref.step(a)[0]
    assert abs(r - r_ref) <= 1e-5 * abs(r_ref) + 800 * 2e-7, (r, r_ref)
assert np.array_equal(env.state[0], ref.state)
env.close()
print("single env ok", round(time.time() - t0, 2), flush=True)

E = 3
loaders = [bh.SyntheticLoader(N, F, 1, seeds=(10 + i,)) for i in range(E)]
tf = lambda t: next(l for l in loaders if np.ascontiguousarray(t[0, 0, 0, :4]).tobytes() in l._pre).target_function(t)
vec = bh.HologramVecEnv(E, tf, loaders, max_steps=10 ** 6, T_PSNR_DIFF=1e9, IPS=N, CH=F, resync_every=3)
vec.reset()
refs = []
for i in range(E):
    e = O.OracleEnv(cfg, max_steps=10 ** 6, T_PSNR_DIFF=1e9)
    e.reset(*bh.synthetic_problem(N, F, 1, 10 + i))
    refs.append(e)
for _ in range(30):
    acts = rng.integers(0, F * N * N, size=E)
    rewards = vec.step(acts)[1]
    for i in range(E):
        r_ref = refs[i].step(int(acts[i]))[0]
        assert abs(rewards[i] - r_ref) <= 1e-5 * abs(r_ref) + 800 * 2e-7, (i, rewards[i], r_ref)
for i in range(E):
    assert np.array_equal(vec.engine.state(i), refs[i].state)
vec.close()
print("vec env ok", round(time.time() - t0, 2), flush=True)

env = bh.BinaryHologramEnv(ld.target_function, ld, IPS=N, CH=F, verbose=False)
res = bh.optimize_with_random_pixel_flips(env, max_datasets=0, rng=np.random.default_rng(4), verbose=False,
                                          psnr_diff_threshold=0.05, max_candidates=2000)[0]
st_ref, acc_ref, tr_ref = O.dbs_greedy(cfg, (pre >= 0.5).astype(np.int8), tgt, res["order"][:res["steps"]])
assert res["stopped_on_threshold"] and np.array_equal(res["accepted"].astype(bool), acc_ref)
assert np.array_equal(res["state"], st_ref)
env.close()
print("dbs early stop ok at step", res["steps"], round(time.time() - t0, 2), flush=True)
print("SANITY_OK")
