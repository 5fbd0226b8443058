"""pad = 2 (zero-padded, linear propagation) at the BASELINE shape: timings of the hot kernels."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import binary_hologram_reinforcement_learning_b200 as bh
N, F, E = 1024, 24, 4
t0 = time.time()
eng = bh.HoloEngine(N, F, bh.WL_RGB, n_env=E, pad=2)
print("create (host tables 2048^2 x 3)", round(time.time() - t0, 2), "s")
pre, tgt = bh.synthetic_problem(N, F, 3, 0)
st = (pre >= 0.5).astype(np.int8)
for e in range(E):
    eng.set_target(e, tgt); eng.load_state(e, st)
print("psnr", eng.metrics(0)[0], "propagate ms", eng.time_propagate(0, 3), eng.time_propagate_passes(0, 3))
rng = np.random.default_rng(0)
sets = torch.from_numpy(rng.integers(0, F * N * N, size=(64, E), dtype=np.int64)).cuda()
envs = torch.arange(E, dtype=torch.int32, device="cuda")
ms = eng.time_eval(E, envs.data_ptr(), sets.data_ptr(), 64, 256)
print("k_eval pad=2: ms per", E, "candidates", ms, "GB/s", 16 * N * N * E / ms / 1e6)
acts = rng.integers(0, F * N * N, size=2000)
t0 = time.perf_counter(); eng.eval_flips(acts); print("eval_flips/s", 2000 / (time.perf_counter() - t0))
