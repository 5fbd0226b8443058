"""Where the host time of dbs_sweep goes at 1024^2 x 24 with a 64-pixel crop (device time of the sweep: 2.5 ms)."""
import cProfile, pstats, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import binary_hologram_reinforcement_learning_b200 as bh
ld = bh.SyntheticLoader(1024, 24, 3, seeds=(1, 2, 3))
pr = cProfile.Profile()
pr.enable()
t0 = time.perf_counter()
rs = bh.dbs_sweep(ld.target_function, ld, 2e-3, 7.56e-6, 64, CH=24, wl=bh.WL_RGB, max_datasets=2,
                  rng=np.random.default_rng(9), verbose=False, device=0)
dt = time.perf_counter() - t0
pr.disable()
print("3 images", dt, [r["seconds"] for r in rs])
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
