"""Round-2 timing sweep of the delta kernels at 1024^2 x 24, 8 envs (GPU; one subprocess per variant).

    python scripts/r2_tune.py           # prints one JSON line per configuration

Per configuration (BHOLO_EVAL_VARIANT x BHOLO_EVAL_FP64 x BHOLO_COMMIT_VARIANT): back-to-back k_eval
launches (8 candidates and 1 candidate), the eval -> commit chain of the step path at the env accept rule,
the eval-only chain, and k_commit alone with 8 kept flips per launch.
"""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) > 1 and sys.argv[1] == "run":
    import numpy as np
    import torch
    import binary_hologram_reinforcement_learning_b200 as bh
    from binary_hologram_reinforcement_learning_b200.engine import RULE_ENV
    N, F, E = 1024, 24, 8
    eng = bh.HoloEngine(N, F, bh.WL_RGB, n_env=E)
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    for e in range(E):
        pre, tgt = bh.synthetic_problem(N, F, 3, e)
        eng.set_target(e, tgt); eng.load_state(e, (pre >= 0.5).astype(np.int8))
    rng = np.random.default_rng(0)
    sets = torch.from_numpy(rng.integers(0, F * N * N, size=(64, E), dtype=np.int64)).cuda()
    envs = torch.arange(E, dtype=torch.int32, device="cuda")
    one = torch.from_numpy(rng.integers(0, F * N * N, size=(64, 1), dtype=np.int64)).cuda()
    fresh = [torch.from_numpy(rng.integers(0, F * N * N, size=(600, E), dtype=np.int64)).cuda() for _ in range(3)]
    torch.cuda.synchronize()
    best = lambda f: min(f() for _ in range(3))
    ms8 = best(lambda: eng.time_eval(E, envs.data_ptr(), sets.data_ptr(), 64, 512))
    ms1 = best(lambda: eng.time_eval(1, envs.data_ptr(), one.data_ptr(), 64, 512))
    # fresh actions for every step (a revisited pixel is almost never kept again: the chain would time no commits)
    it = iter(fresh)
    step = best(lambda: eng.time_step(E, envs.data_ptr(), next(it).data_ptr(), 600, 512, RULE_ENV, True))
    evo = best(lambda: eng.time_step(E, envs.data_ptr(), sets.data_ptr(), 64, 512, RULE_ENV, False))
    com = best(lambda: eng.time_commit(E, envs.data_ptr(), sets.data_ptr(), 64, 512))
    print(json.dumps({"eval": os.environ.get("BHOLO_EVAL_VARIANT", "0"), "fp64": bool(os.environ.get("BHOLO_EVAL_FP64")),
                      "commit": os.environ.get("BHOLO_COMMIT_VARIANT", "0"),
                      "eval8_us": round(ms8 * 1e3, 2), "eval8_gbs": round(16 * N * N * E / ms8 / 1e6),
                      "eval1_us": round(ms1 * 1e3, 2), "step_us": round(step * 1e3, 2), "eval_only_chain_us": round(evo * 1e3, 2),
                      "commit8_us": round(com * 1e3, 2), "commit8_gbs": round(24 * N * N * E / com / 1e6)}))
else:
    grid = [(0, 0, 0), (0, 1, 0)]
    if "--full" in sys.argv:
        grid += [(v, 0, 0) for v in (1, 2, 4, 6, 7)] + [(0, 0, c) for c in (1, 2, 4, 5, 6)]
    for ev, fp64, cv in grid:
        env = dict(os.environ, BHOLO_EVAL_VARIANT=str(ev), BHOLO_COMMIT_VARIANT=str(cv))
        env.pop("BHOLO_EVAL_FP64", None)
        if fp64:
            env["BHOLO_EVAL_FP64"] = "1"
        r = subprocess.run([sys.executable, __file__, "run"], env=env, capture_output=True, text=True)
        print(r.stdout.strip() or r.stderr[-800:], flush=True)
