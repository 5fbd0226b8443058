"""Persistent rollout kernel (bh_rollout_device) against the two-kernel step chain at 1024^2 x 24: 8 envs x 512 steps
(the bench's `value` region) and 1 env (the sequential greedy DBS chain), device-resident actions, CUDA events."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import binary_hologram_reinforcement_learning_b200 as bh
from binary_hologram_reinforcement_learning_b200.engine import RULE_ENV, RULE_DBS

N, F = 1024, 24
for E, steps, rule in ((8, 512, RULE_ENV), (1, 2048, RULE_DBS), (2, 1024, RULE_DBS), (4, 1024, RULE_DBS)):
    eng = bh.HoloEngine(N, F, bh.WL_RGB, n_env=E)
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    for e in range(E):
        pre, tgt = bh.synthetic_problem(N, F, 3, e)
        eng.set_target(e, tgt); eng.load_state(e, (pre >= 0.5).astype(np.int8))
    rng = np.random.default_rng(1)
    acts = torch.from_numpy(rng.integers(0, F * N * N, size=(3, steps, E), dtype=np.int64)).cuda()
    envs = torch.arange(E, dtype=torch.int32, device="cuda")
    res = torch.zeros(steps * E * 40, dtype=torch.uint8, device="cuda")
    out = {"envs": E, "steps": steps, "rule": rule}
    with torch.cuda.stream(stream):
        for name in ("chain", "rollout", "chain", "rollout"):
            k = 0 if name == "chain" else 1
            a = acts[(k + (0 if "chain2" not in out else 2)) % 3]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            if name == "chain":
                for t in range(steps):
                    eng.step_batch_device(E, envs.data_ptr(), a.data_ptr() + t * E * 8, rule, res.data_ptr() + t * E * 40)
            else:
                eng.rollout_device(E, envs.data_ptr(), a.data_ptr(), steps, rule, res.data_ptr())
            e1.record(stream)
            e1.synchronize()
            if name == "rollout":
                eng.rollout_status()
            ms = e0.elapsed_time(e1)
            r = np.frombuffer(res.cpu().numpy().tobytes(), dtype=bh.engine.RESULT_DTYPE)
            key = name if name not in out else name + "2"
            out[key] = {"us_per_step": round(1e3 * ms / steps, 2), "steps_per_s": round(E * steps / ms * 1e3),
                        "accept_rate": round(float(r["accept"].mean()), 3)}
    print(json.dumps(out), flush=True)
    eng.close()
