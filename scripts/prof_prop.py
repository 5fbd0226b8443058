"""Small driver for profiling the propagation passes alone (1024^2 x 24, one env)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import binary_hologram_reinforcement_learning_b200 as bh

N, F = int(os.environ.get("PN", 1024)), 24
pre, tgt = bh.synthetic_problem(N, F, 3, 0)
eng = bh.HoloEngine(N, F, bh.WL_RGB, n_env=1)
eng.set_target(0, tgt)
eng.load_state(0, (pre >= 0.5).astype(np.int8))
print("psnr", eng.metrics(0)[0], "ms/propagation", eng.time_propagate(0, 5))
print("per-pass ms (rows_fwd, cols, rows_inv, intensity), live:", eng.time_propagate_passes(0, 5))
eng.close()
