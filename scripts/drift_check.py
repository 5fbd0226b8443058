"""fp32 drift of the incrementally updated fields: PSNR before/after a full re-propagation."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import binary_hologram_reinforcement_learning_b200 as bh
for N, F, wl, n_cand in [(256, 8, bh.WL_MONO, 200000), (1024, 24, bh.WL_RGB, 60000)]:
    pre, tgt = bh.synthetic_problem(N, F, len(wl), 0)
    eng = bh.HoloEngine(N, F, wl)
    eng.set_target(0, tgt); eng.load_state(0, (pre >= 0.5).astype(np.int8))
    order = np.random.default_rng(0).permutation(F * N * N)[:n_cand]
    done = 0
    for chunk in (1000, 9000, 40000, n_cand - 50000):
        acc, _, nacc, _ = eng.dbs_run(order[done:done + chunk], resync_every=0)
        done += chunk
        p_inc, _, s_inc = eng.metrics(0)
        st = eng.state(0)
        eng2 = bh.HoloEngine(N, F, wl); eng2.set_target(0, tgt); eng2.load_state(0, st)
        p_ref, _, s_ref = eng2.metrics(0); eng2.close()
        print(f"N={N} candidates={done} psnr_incremental={p_inc:.9f} psnr_repropagated={p_ref:.9f} "
              f"diff_dB={p_inc - p_ref:+.3e} rel_sumI2={(s_inc[0] - s_ref[0]) / s_ref[0]:+.2e}")
    eng.close()
