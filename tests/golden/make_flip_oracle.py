"""Float64 oracle values of random single-pixel flips at the BASELINE shapes (test infrastructure).

    python tests/golden/make_flip_oracle.py            # writes tests/golden/flips_<name>.npz

For every shape: the seeded synthetic problem of oracle.synthetic_problem, K random actions and, from
the float64 restatement (oracle/hologram_oracle.py: propagate_group + delta_terms + the closed-form
loss), the PSNR before, the change of the two loss sums and the PSNR after each flip.  The GPU script
scripts/parity_error_dist.py and tests/test_gpu_parity.py::test_reward_parity_bound compare the CUDA
kernels with these numbers; generating them needs no GPU and minutes of CPU, so they are committed.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import hologram_oracle as O  # noqa: E402

CASES = {
    # name: (N, F, wavelengths, seed, flips)
    "mono256": (256, 8, O.WL_MONO, 11, 2048),
    "rgb896": (896, 24, O.WL_RGB, 12, 2048),
    "rgb1024": (1024, 24, O.WL_RGB, 13, 2048),
}


def make(name):
    N, F, wl, seed, K = CASES[name]
    cfg = O.HoloConfig(N=N, F=F, wl=wl)
    pre, tgt = O.synthetic_problem(N, F, len(wl), seed=seed)
    st = (pre >= 0.5).astype(np.int8)
    t64 = tgt.astype(np.float64)
    U = [O.propagate_group(cfg, st[g * cfg.Fg:(g + 1) * cfg.Fg], g) for g in range(cfg.G)]
    I = np.stack([O.group_mean_intensity(u) for u in U])
    sii, sit, stt = O.loss_sums(I, t64)
    n = cfg.G * N * N
    psnr0 = O.psnr_from_mse(O.mse_from_sums(sii, sit, stt, n, True))
    psnr0_abs = O.psnr_from_mse(O.mse_from_sums(sii, sit, stt, n, False))
    rng = np.random.default_rng(seed + 1000)
    actions = rng.integers(0, F * N * N, size=K, dtype=np.int64)
    d_sii, d_sit = np.zeros(K), np.zeros(K)
    psnr_rel, psnr_abs = np.zeros(K), np.zeros(K)
    t0 = time.time()
    for i, a in enumerate(actions):
        f, r, c = cfg.decode(int(a))
        g = cfg.group_of(f)
        s = 1 - 2 * int(st[f, r, c])
        d_sii[i], d_sit[i], _ = O.delta_terms(cfg, U[g][f - g * cfg.Fg], I[g], t64[g], g, r, c, s)
        psnr_rel[i] = O.psnr_from_mse(O.mse_from_sums(sii + d_sii[i], sit + d_sit[i], stt, n, True))
        psnr_abs[i] = O.psnr_from_mse(O.mse_from_sums(sii + d_sii[i], sit + d_sit[i], stt, n, False))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), f"flips_{name}.npz")
    np.savez_compressed(out, N=N, F=F, wl=np.array(wl), seed=seed, actions=actions, psnr0=psnr0,
                        psnr0_abs=psnr0_abs, d_sii=d_sii, d_sit=d_sit, psnr_after=psnr_rel,
                        psnr_after_abs=psnr_abs, sums=np.array([sii, sit, stt]))
    print(f"{name}: psnr0 {psnr0:.6f}, median |dPSNR| {np.median(np.abs(psnr_rel - psnr0)):.3e} dB, "
          f"{time.time() - t0:.1f} s -> {out}", flush=True)


if __name__ == "__main__":
    for nm in (sys.argv[1:] or CASES):
        make(nm)
