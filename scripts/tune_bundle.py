"""Time the bundled candidate-list evaluation (BHOLO_BUNDLE_VARIANT; 9 = off, lists go through k_eval_t).

(a) windows of K random candidates of ONE 1024^2 x 24 environment (a DBS speculation window),
(b) eval_flips of 4096 random candidates through the host API (sorted by frame inside),
(c) a whole greedy DBS of a 256^2 x 8 hologram (524 288 candidates, adaptive window).
"""
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) > 1:
    import numpy as np
    import torch
    import binary_hologram_reinforcement_learning_b200 as bh

    out = {"variant": os.environ.get("BHOLO_BUNDLE_VARIANT", "0")}
    N, F = 1024, 24
    eng = bh.HoloEngine(N, F, bh.WL_RGB, n_env=1)
    pre, tgt = bh.synthetic_problem(N, F, 3, 0)
    eng.set_target(0, tgt)
    eng.load_state(0, (pre >= 0.5).astype(np.int8))
    rng = np.random.default_rng(0)
    for K in (8, 32, 64, 128):
        sets = torch.from_numpy(rng.integers(0, F * N * N, size=(64, K), dtype=np.int64)).cuda()
        ms = min(eng.time_eval(K, 0, sets.data_ptr(), 64, 128) for _ in range(3))
        out[f"window_{K}_us_per_candidate"] = round(1e3 * ms / K, 3)
    acts = rng.integers(0, F * N * N, size=4096, dtype=np.int64)
    eng.eval_flips(acts[:256])
    t0 = time.perf_counter()
    eng.eval_flips(acts)
    out["eval_flips_4096_per_s"] = round(4096 / (time.perf_counter() - t0))
    eng.close()

    N, F = 256, 8
    eng = bh.HoloEngine(N, F, [515e-9], n_env=1)
    pre, tgt = bh.synthetic_problem(N, F, 1, 0)
    eng.set_target(0, tgt)
    st = (pre >= 0.5).astype(np.int8)
    order = np.random.default_rng(1).permutation(F * N * N).astype(np.int64)
    eng.load_state(0, st)
    eng.dbs_run(order[:20000])
    eng.load_state(0, st)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    acc, _, n_acc, psnr = eng.dbs_run(order)
    out["dbs_256x8_full_s"] = round(time.perf_counter() - t0, 3)
    out["dbs_256x8_accepted"] = int(n_acc)
    out["dbs_256x8_final_psnr"] = round(float(psnr), 6)
    print(json.dumps(out))
else:
    for v in sys.argv[1:] or [9, 0, 1, 2, 3, 4]:
        env = dict(os.environ, BHOLO_BUNDLE_VARIANT=str(v))
        r = subprocess.run([sys.executable, __file__, "run"], env=env, capture_output=True, text=True)
        print(r.stdout.strip() or r.stderr.strip()[-400:], flush=True)
