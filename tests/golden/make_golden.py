"""Generate the golden fixtures of tests/golden/ from the float64 CPU oracle.

    python tests/golden/make_golden.py

The reference holds no tests, golden vectors or logs for this path and its
arithmetic lives in the absent package torchOptics, so these
vectors pin OUR restatement (oracle/hologram_oracle.py), not the reference:
parity stays "unpinned" until the torchOptics source is available.  Inputs are
regenerated from seeds by synthetic_problem(); only outputs are stored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import hologram_oracle as O  # noqa: E402

CASES = {
    # name: (N, F, wl, pad, relative, seed)
    "mono64": (64, 8, O.WL_MONO, 1, True, 11),
    "rgb64": (64, 6, O.WL_RGB, 1, True, 12),
    "mono32_pad2": (32, 4, O.WL_MONO, 2, True, 13),
    "mono64_abs": (64, 8, O.WL_MONO, 1, False, 14),
}
N_STEPS, N_DBS, N_SWEEP = 150, 200, 120


def make_case(name):
    N, F, wl, pad, relative, seed = CASES[name]
    cfg = O.HoloConfig(N=N, F=F, wl=wl, pad=pad, relative=relative)
    pre, tgt = O.synthetic_problem(N, F, len(wl), seed)
    rng = np.random.default_rng(seed + 1000)
    out = {}
    state = (pre >= 0.5).astype(np.int8)
    recon = O.reconstruct(cfg, state)
    psnr, mse = O.score(cfg, recon, tgt)
    out["initial_psnr"], out["initial_mse"] = psnr, mse
    out["loss_sums"] = np.array(O.loss_sums(recon, tgt))
    U = O.propagate_group(cfg, state[:cfg.Fg], 0)
    out["field_samples"] = U[0, ::max(1, N // 8), ::max(1, N // 8)].astype(np.complex128)
    out["h_samples"] = cfg.h(0)[:4, :4].copy()
    # env trajectory (env.py:154-260)
    env = O.OracleEnv(cfg, max_steps=N_STEPS - 10, T_PSNR_DIFF=1e9)
    env.reset(pre, tgt)
    actions = rng.integers(0, F * N * N, size=N_STEPS)
    rewards, psnrs, accs, terms = [], [], [], []
    for a in actions:
        r, term, trunc, p, acc = env.step(int(a))
        rewards.append(r); psnrs.append(p); accs.append(acc); terms.append(term)
    out["env_actions"] = actions
    out["env_rewards"] = np.array(rewards)
    out["env_psnr"] = np.array(psnrs)
    out["env_accepted"] = np.array(accs)
    out["env_terminated"] = np.array(terms)
    out["env_final_state_sum"] = int(env.state.sum())
    # greedy DBS (DBS.py:247-294)
    order = rng.permutation(F * N * N)[:N_DBS]
    st, acc, trace = O.dbs_greedy(cfg, state, tgt, order)
    out["dbs_order"], out["dbs_accepted"], out["dbs_trace"] = order, acc, trace
    out["dbs_final_state_sum"] = int(st.sum())
    # score-and-revert sweep (dbs-1024-1024-24-6464.py:330-395)
    sw_order = rng.permutation(F * N * N)[:N_SWEEP]
    ps, p0, att, imp, gain = O.sweep(cfg, state, tgt, pre, sw_order)
    out["sweep_order"], out["sweep_psnr"] = sw_order, ps
    out["sweep_attempted"], out["sweep_improved"], out["sweep_gain"] = att, imp, gain
    return out


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    for name in CASES:
        data = make_case(name)
        np.savez_compressed(os.path.join(here, f"{name}.npz"), **data)
        print(name, "psnr0=%.6f" % data["initial_psnr"], "env accepts=%d" % data["env_accepted"].sum(),
              "dbs accepts=%d" % data["dbs_accepted"].sum())
