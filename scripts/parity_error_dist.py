"""Error distribution of dPSNR / reward of the CUDA delta paths against the float64 oracle (GPU).

    python scripts/parity_error_dist.py [out.json]

Reads the committed oracle fixtures tests/golden/flips_<shape>.npz (tests/golden/make_flip_oracle.py)
and scores the same flips on the B200 through the C ABI by three routes:
  k_eval_t         bh_step_batch(rule = never), one candidate per launch
  k_eval_bundle_t  bh_eval_flips (host list sorted by frame, bundled kernel)
  sweep_all        bh_sweep_all (FFT correlations), indexed at the same actions
for the per-quad arithmetic in float (default) and in double (BHOLO_EVAL_FP64=1).  The quantity the
reward is made of is dPSNR = psnr_after - psnr_before (env.py:184-188, reward = 800 * dPSNR); errors
are reported relative to |dPSNR_oracle| (median / p90 / p99 / max) and in absolute dB.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import binary_hologram_reinforcement_learning_b200 as bh  # noqa: E402
from binary_hologram_reinforcement_learning_b200.engine import RULE_NEVER  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def dist(err_abs, ref):
    rel = err_abs / np.maximum(np.abs(ref), 1e-300)
    q = lambda a, p: float(np.quantile(a, p))
    return {"rel_median": q(rel, .5), "rel_p90": q(rel, .9), "rel_p99": q(rel, .99), "rel_max": float(rel.max()),
            "abs_dB_median": q(err_abs, .5), "abs_dB_p99": q(err_abs, .99), "abs_dB_max": float(err_abs.max()),
            "reward_abs_max": float(800.0 * err_abs.max()),
            # relative error of the bulk: flips whose |dPSNR| is at least a tenth of the median
            "rel_max_excluding_tiny": float(rel[np.abs(ref) >= 0.1 * np.median(np.abs(ref))].max())}


def run_case(name, fp32):
    d = np.load(os.path.join(GOLD, f"flips_{name}.npz"))
    N, F, wl, seed = int(d["N"]), int(d["F"]), tuple(float(w) for w in d["wl"]), int(d["seed"])
    pre, tgt = bh.synthetic_problem(N, F, len(wl), seed=seed)
    st = (pre >= 0.5).astype(np.int8)
    if fp32:
        os.environ.pop("BHOLO_EVAL_FP64", None)
    else:
        os.environ["BHOLO_EVAL_FP64"] = "1"
    eng = bh.HoloEngine(N, F, wl, n_env=1, device=0)
    eng.set_target(0, tgt)
    eng.load_state(0, st)
    psnr0 = eng.metrics(0)[0]
    acts = d["actions"]
    ref = d["psnr_after"] - float(d["psnr0"])
    out = {"psnr0_err_dB": abs(psnr0 - float(d["psnr0"])), "median_abs_dpsnr_dB": float(np.median(np.abs(ref))),
           "flips": int(acts.size)}
    one = np.array([psnr0 * 0 + eng.step_batch(acts[i:i + 1], np.zeros(1, np.int32), RULE_NEVER)["psnr_after"][0]
                    for i in range(acts.size)])
    out["k_eval_t"] = dist(np.abs((one - psnr0) - ref), ref)
    lst = eng.eval_flips(acts, env=0)
    out["k_eval_bundle_t"] = dist(np.abs((lst - psnr0) - ref), ref)
    out["bundle_equals_single_bitwise"] = bool(np.array_equal(lst, one))
    sw = eng.sweep_all(0).reshape(-1)[acts]
    out["sweep_all"] = dist(np.abs((sw - psnr0) - ref), ref)
    # decisions (env rule, ties kept) against the oracle
    out["decisions_differ_k_eval_t"] = int(np.sum(((one - psnr0) >= 0) != (ref >= 0)))
    eng.close()
    return out


def main():
    res = {"what": "error of dPSNR (GPU) vs float64 oracle over random single-pixel flips; reward = 800 * dPSNR",
           "cases": {}}
    for name in ("mono256", "rgb896", "rgb1024"):
        if not os.path.exists(os.path.join(GOLD, f"flips_{name}.npz")):
            continue
        res["cases"][name] = {"fp64_quads": run_case(name, False), "fp32_quads": run_case(name, True)}
        c = res["cases"][name]
        for k in ("fp64_quads", "fp32_quads"):
            e = c[k]["k_eval_t"]
            print(f"{name} {k}: median|dPSNR| {c[k]['median_abs_dpsnr_dB']:.2e} dB; k_eval_t rel med {e['rel_median']:.2e} "
                  f"p99 {e['rel_p99']:.2e} max {e['rel_max']:.2e} (bulk max {e['rel_max_excluding_tiny']:.2e}); "
                  f"abs max {e['abs_dB_max']:.2e} dB; sweep rel p99 {c[k]['sweep_all']['rel_p99']:.2e}; "
                  f"bundle==single {c[k]['bundle_equals_single_bitwise']}; decisions differ {c[k]['decisions_differ_k_eval_t']}",
                  flush=True)
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "parity_error_dist.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        json.dump(res, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
