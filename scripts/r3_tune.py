"""Round-2 (session 3) timing of the step kernels at 1024^2 x 24, 8 envs: register-staged kernels (BHOLO_PIPE=0)
against the TMA-pipelined ones with S stages (BHOLO_PIPE=S), each also as two sub-batches on two streams
(BHOLO_SPLIT=2, experiment inside bh_time_step).  One subprocess per configuration, one JSON line each.
Also checks that the pipelined evaluation is bit-identical to the register-staged one."""
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
    # checksum of 256 evaluations through the public list API (bit-identity across kernel variants)
    cand = rng.integers(0, F * N * N, size=256, dtype=np.int64)
    from binary_hologram_reinforcement_learning_b200.engine import RULE_NEVER
    ids = np.arange(E, dtype=np.int32)
    chk = 0.0
    for j in range(16):
        r = eng.step_batch(cand[j * E:(j + 1) * E], ids, RULE_NEVER)
        chk += float(np.sum(r["psnr_after"] * (1 + np.arange(E))))
    r = eng.step_batch(cand[:E], ids, RULE_ENV)
    r2 = eng.step_batch(cand[E:2 * E], ids, RULE_ENV)
    chk2 = float(np.sum(r["psnr_after"])) + float(np.sum(r["accept"])) + float(np.sum(r2["psnr_after"]))
    best = lambda f: min(f() for _ in range(3))
    ms8 = best(lambda: eng.time_eval(E, envs.data_ptr(), sets.data_ptr(), 64, 512))
    ms1 = best(lambda: eng.time_eval(1, envs.data_ptr(), one.data_ptr(), 64, 512))
    it = iter(fresh)
    step = best(lambda: eng.time_step(E, envs.data_ptr(), next(it).data_ptr(), 600, 512, RULE_ENV, True))
    com = best(lambda: eng.time_commit(E, envs.data_ptr(), sets.data_ptr(), 64, 512))
    psnr = [eng.metrics(e)[0] for e in range(E)]
    print(json.dumps({"pipe": os.environ.get("BHOLO_PIPE", "12"), "split": os.environ.get("BHOLO_SPLIT", "1"),
                      "fused": os.environ.get("BHOLO_FUSED", "1"),
                      "eval8_us": round(ms8 * 1e3, 2), "eval8_gbs": round(16 * N * N * E / ms8 / 1e6),
                      "eval1_us": round(ms1 * 1e3, 2), "step_us": round(step * 1e3, 2),
                      "commit8_us": round(com * 1e3, 2), "commit8_gbs": round(24 * N * N * E / com / 1e6),
                      "chk_eval": chk, "chk_step": chk2, "psnr_sum_after": float(np.sum(psnr))}))
else:
    configs = [("0", "1", "0"), ("0", "1", "1"), ("12", "1", "0"), ("8", "1", "0"), ("6", "1", "0"), ("4", "1", "0")]
    if len(sys.argv) > 1:
        configs = [tuple(x.split(":")) for x in sys.argv[1:]]
    for pipe, split, fused in configs:
        env = dict(os.environ, BHOLO_PIPE=pipe, BHOLO_SPLIT=split, BHOLO_FUSED=fused)
        r = subprocess.run([sys.executable, __file__, "run"], env=env, capture_output=True, text=True)
        print(r.stdout.strip() or r.stderr[-1200:], flush=True)
