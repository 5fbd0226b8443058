"""Driver for profiling k_rollout_t: 8 envs at 1024^2 x 24, rollouts of 512 vectorised steps (the bench's launch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import binary_hologram_reinforcement_learning_b200 as bh
from binary_hologram_reinforcement_learning_b200.engine import RULE_ENV, RESULT_DTYPE

N, F, E, R = 1024, 24, 8, int(os.environ.get("ROLLOUT", 512))
eng = bh.HoloEngine(N, F, bh.WL_RGB, n_env=E)
for e in range(E):
    pre, tgt = bh.synthetic_problem(N, F, 3, e)
    eng.set_target(e, tgt); eng.load_state(e, (pre >= 0.5).astype(np.int8))
rng = np.random.default_rng(0)
acts = torch.from_numpy(rng.integers(0, F * N * N, size=(3, R, E), dtype=np.int64)).cuda()
envs = torch.arange(E, dtype=torch.int32, device="cuda")
res = torch.zeros(R * E * 40, dtype=torch.uint8, device="cuda")
for i in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.rollout_device(E, envs.data_ptr(), acts[i].data_ptr(), R, RULE_ENV, res.data_ptr())
    e1.record(); e1.synchronize()
    eng.rollout_status()
    acc = float(np.frombuffer(res.cpu().numpy().tobytes(), dtype=RESULT_DTYPE)["accept"].mean())
    print("rollout", i, "ms", e0.elapsed_time(e1), "us per vectorised step", 1e3 * e0.elapsed_time(e1) / R, "accept rate", acc)
eng.close()
