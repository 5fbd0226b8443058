"""Time k_eval variants (BHOLO_EVAL_VARIANT) at 1024^2 x 24, 8 candidates (one per env) per launch."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import numpy as np, torch
    import binary_hologram_reinforcement_learning_b200 as bh
    N, F, E = 1024, 24, 8
    eng = bh.HoloEngine(N, F, bh.WL_RGB, n_env=E)
    pre, tgt = bh.synthetic_problem(N, F, 3, 0)
    st = (pre >= 0.5).astype(np.int8)
    for e in range(E):
        eng.set_target(e, tgt); eng.load_state(e, st)
    rng = np.random.default_rng(0)
    sets = torch.from_numpy(rng.integers(0, F * N * N, size=(64, E), dtype=np.int64)).cuda()
    envs = torch.arange(E, dtype=torch.int32, device="cuda")
    ms = min(eng.time_eval(E, envs.data_ptr(), sets.data_ptr(), 64, 512) for _ in range(3))
    one = torch.from_numpy(rng.integers(0, F * N * N, size=(64, 1), dtype=np.int64)).cuda()
    ms1 = min(eng.time_eval(1, envs.data_ptr(), one.data_ptr(), 64, 512) for _ in range(3))
    print(json.dumps({"variant": os.environ.get("BHOLO_EVAL_VARIANT", "0"), "ms_8": ms,
                      "gbs_8": 16 * N * N * E / ms / 1e6, "ms_1": ms1, "gbs_1": 16 * N * N / ms1 / 1e6}))
else:
    for v in [0, 1, 2, 3, 4, 5, 6]:
        env = dict(os.environ, BHOLO_EVAL_VARIANT=str(v))
        print(subprocess.run([sys.executable, __file__, "run"], env=env, capture_output=True, text=True).stdout.strip())
