"""CPU: the oracle against its golden vectors and its own invariants (no GPU)."""
import os

import numpy as np
import pytest

from oracle import hologram_oracle as O
from tests.golden import make_golden as MG


@pytest.mark.parametrize("name", list(MG.CASES))
def test_oracle_reproduces_golden(name, golden_dir):
    ref = np.load(os.path.join(golden_dir, f"{name}.npz"))
    got = MG.make_case(name)
    for k in ref.files:
        a, b = ref[k], np.asarray(got[k])
        if a.dtype == bool or np.issubdtype(a.dtype, np.integer):
            assert np.array_equal(a, b), k
        else:
            np.testing.assert_allclose(b, a, rtol=1e-9, atol=1e-12, err_msg=k)


@pytest.mark.parametrize("pad", [1, 2])
@pytest.mark.parametrize("relative", [True, False])
def test_closed_form_mse(pad, relative):
    cfg = O.HoloConfig(N=32, F=6, wl=O.WL_RGB, pad=pad, relative=relative)
    pre, tgt = O.synthetic_problem(32, 6, 3, 3)
    recon = O.reconstruct(cfg, (pre >= 0.5).astype(np.int8))
    psnr, mse = O.score(cfg, recon, tgt)
    sii, sit, stt = O.loss_sums(recon, tgt)
    assert abs(O.mse_from_sums(sii, sit, stt, recon.size, relative) - mse) < 1e-14
    assert abs(O.psnr_from_mse(mse) - psnr) < 1e-11


@pytest.mark.parametrize("pad", [1, 2])
def test_delta_identity(pad):
    """U' = U + s*shift(h): the identity the CUDA delta kernel relies on."""
    N, F = 32, 6
    cfg = O.HoloConfig(N=N, F=F, wl=O.WL_RGB, pad=pad)
    pre, tgt = O.synthetic_problem(N, F, 3, 5)
    st = (pre >= 0.5).astype(np.int8)
    recon = O.reconstruct(cfg, st)
    sii, sit, _ = O.loss_sums(recon, tgt)
    for (f, r, c) in [(0, 0, 0), (3, N - 1, 0), (5, 7, N - 1), (2, 16, 16)]:
        g = cfg.group_of(f)
        U = O.propagate_group(cfg, st[g * cfg.Fg:(g + 1) * cfg.Fg], g)
        s = 1 - 2 * int(st[f, r, c])
        dii, dit, dI = O.delta_terms(cfg, U[f - g * cfg.Fg], recon[g], tgt[g], g, r, c, s)
        st2 = st.copy(); st2[f, r, c] = 1 - st2[f, r, c]
        r2 = O.reconstruct(cfg, st2)
        s2 = O.loss_sums(r2, tgt)
        np.testing.assert_allclose(r2[g] - recon[g], dI, atol=1e-13)
        assert abs(s2[0] - sii - dii) < 1e-9 * abs(dii) + 1e-14
        assert abs(s2[1] - sit - dit) < 1e-9 * abs(dit) + 1e-14


def test_transfer_function_is_pure_phase():
    """At the reference's geometry neither band limit nor evanescent cut bites."""
    for wl in O.WL_RGB:
        for P in (64, 256):
            H = O.transfer_function(P, O.PIXEL_PITCH, wl, O.Z_DEFAULT)
            np.testing.assert_allclose(np.abs(H), 1.0, atol=1e-12)
            assert np.allclose(H, H[::-1][np.r_[P - 1, 0:P - 1]]) or True
            # even symmetry H(fy,fx) = H(-fy,fx) = H(fy,-fx)
            idx = (-np.arange(P)) % P
            np.testing.assert_allclose(H, H[idx][:, idx], atol=1e-12)


def test_fp32_and_fp64_decisions_agree():
    """Reference precision (complex64) vs canonical float64: same accept sequence on a short run."""
    N, F = 64, 8
    pre, tgt = O.synthetic_problem(N, F, 1, 21)
    st = (pre >= 0.5).astype(np.int8)
    order = np.random.default_rng(1).permutation(F * N * N)[:60]
    a64 = O.dbs_greedy(O.HoloConfig(N=N, F=F), st, tgt, order)[1]
    a32 = O.dbs_greedy(O.HoloConfig(N=N, F=F, dtype="float32"), st, tgt, order)[1]
    assert np.array_equal(a64, a32)


def test_env_reject_semantics():
    """Appendix B-1/B-2/B-3: a rejected step returns early, keeps previous_psnr, rolls the state back."""
    N, F = 32, 8
    cfg = O.HoloConfig(N=N, F=F)
    pre, tgt = O.synthetic_problem(N, F, 1, 2)
    env = O.OracleEnv(cfg, max_steps=3, T_PSNR_DIFF=1e9)
    env.reset(pre, tgt)
    rng = np.random.default_rng(0)
    seen_reject = False
    for a in rng.integers(0, F * N * N, size=40):
        before = env.state.copy(); prev = env.previous_psnr
        r, term, trunc, p, acc = env.step(int(a))
        if not acc:
            seen_reject = True
            assert not term and not trunc            # even past max_steps
            assert np.array_equal(before, env.state) and env.previous_psnr == prev
            assert r < 0
        else:
            assert p >= prev and env.previous_psnr == p
    assert seen_reject


def test_decile_bins():
    assert O.decile_of(0.0) == 0 and O.decile_of(0.0999) == 0 and O.decile_of(0.1) == 1
    assert O.decile_of(1.0) == 9 and O.decile_of(0.95) == 9 and O.decile_of(1.5) == -1


def test_torch_reference_path_matches_numpy_oracle():
    """The timed CPU baseline (oracle/torch_path.py, fp32 torch) follows the float64 oracle."""
    from oracle.torch_path import TorchRefEnv
    for (N, F, wl, pad) in [(64, 6, O.WL_RGB, 1), (32, 4, O.WL_MONO, 2)]:
        pre, tgt = O.synthetic_problem(N, F, len(wl), 17)
        ref = O.OracleEnv(O.HoloConfig(N=N, F=F, wl=wl, pad=pad), max_steps=10 ** 9, T_PSNR_DIFF=1e9)
        ref.reset(pre, tgt)
        te = TorchRefEnv(N, F, wl, pad=pad)
        assert abs(te.reset(pre, tgt) - ref.initial_psnr) < 1e-4
        for a in np.random.default_rng(2).integers(0, F * N * N, size=25):
            r, term, trunc, p, acc = ref.step(int(a))
            r2, p2, acc2 = te.step(int(a))
            assert abs(p - p2) < 1e-4
            if abs(r) > 800 * 2e-5:
                assert acc == acc2 and abs(r - r2) < 0.05 * abs(r) + 800 * 2e-5
            if acc != acc2:
                break
