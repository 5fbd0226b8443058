"""Barrier tuning of k_rollout_t: CTAs per environment cap and poll back-off, 1 / 2 / 8 envs at 1024^2 x 24."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "run":
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    import binary_hologram_reinforcement_learning_b200 as bh
    from binary_hologram_reinforcement_learning_b200.engine import RULE_DBS, RESULT_DTYPE
    N, F = 1024, 24
    out = {"cpe": os.environ.get("BHOLO_ROLLOUT_CPE", "-"), "backoff": os.environ.get("BHOLO_ROLLOUT_BACKOFF", "0")}
    for E, steps in ((1, 2048), (2, 1024), (8, 512)):
        eng = bh.HoloEngine(N, F, bh.WL_RGB, n_env=E)
        for e in range(E):
            pre, tgt = bh.synthetic_problem(N, F, 3, e)
            eng.set_target(e, tgt); eng.load_state(e, (pre >= 0.5).astype(np.int8))
        rng = np.random.default_rng(1)
        acts = torch.from_numpy(rng.integers(0, F * N * N, size=(3, steps, E), dtype=np.int64)).cuda()
        envs = torch.arange(E, dtype=torch.int32, device="cuda")
        res = torch.zeros(steps * E * 40, dtype=torch.uint8, device="cuda")
        best = 1e9
        for i in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.rollout_device(E, envs.data_ptr(), acts[i].data_ptr(), steps, RULE_DBS, res.data_ptr())
            e1.record(); e1.synchronize(); eng.rollout_status()
            best = min(best, e0.elapsed_time(e1))
        acc = float(np.frombuffer(res.cpu().numpy().tobytes(), dtype=RESULT_DTYPE)["accept"].mean())
        out[f"E{E}_us_per_step"] = round(1e3 * best / steps, 2)
        out[f"E{E}_acc"] = round(acc, 3)
        eng.close()
    print(json.dumps(out))
else:
    for cpe, bo in ((0, 0), (0, 32), (0, 100), (0, 300), (148, 0), (148, 100), (74, 0), (74, 100)):
        env = dict(os.environ, BHOLO_ROLLOUT_BACKOFF=str(bo))
        if cpe:
            env["BHOLO_ROLLOUT_CPE"] = str(cpe)
        r = subprocess.run([sys.executable, __file__, "run"], env=env, capture_output=True, text=True)
        print(r.stdout.strip() or r.stderr[-600:], flush=True)
