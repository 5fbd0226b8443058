"""Small driver for profiling the step kernels (1024^2 x 24, 8 envs): a few vectorised steps with fresh actions."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import binary_hologram_reinforcement_learning_b200 as bh
from binary_hologram_reinforcement_learning_b200.engine import RULE_ENV

N, F, E = 1024, 24, 8
eng = bh.HoloEngine(N, F, bh.WL_RGB, n_env=E)
for e in range(E):
    pre, tgt = bh.synthetic_problem(N, F, 3, e)
    eng.set_target(e, tgt); eng.load_state(e, (pre >= 0.5).astype(np.int8))
rng = np.random.default_rng(0)
acts = torch.from_numpy(rng.integers(0, F * N * N, size=(64, E), dtype=np.int64)).cuda()
envs = torch.arange(E, dtype=torch.int32, device="cuda")
print("step us", 1e3 * eng.time_step(E, envs.data_ptr(), acts.data_ptr(), 64, 16, RULE_ENV, True))
eng.close()
