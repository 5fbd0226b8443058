"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle.

Tolerances (BASELINE.json north_star): PSNR within 1e-3 dB, per-step reward within
1e-5 relative, accept/reject sequences identical except documented near-ties.

The reward is 800 * dPSNR (env.py:184-188).  Measured against the float64 oracle over 2048 random flips per
BASELINE shape (profiles/r2_parity_error_dist.json, scripts/parity_error_dist.py): the relative error of dPSNR has
median 4e-7, p90 2.5e-6, p99 2e-5; what is left above 1e-5 relative are flips whose dPSNR is itself tiny, with an
ABSOLUTE error of at most 4e-12 dB at 1024^2 x 24 (1.1e-10 dB at 256^2 x 8).  That error is set by the fp32
storage of the fields U (6e-8 relative per element, summed in quadrature over the support of the impulse
response and divided by n * mse with n = G N^2), not by the kernel arithmetic -- forming the per-quad sums in
double changes it by 10 %.  The asserted bound is therefore
    |dPSNR_gpu - dPSNR_oracle| <= 1e-5 |dPSNR_oracle| + DPSNR_C / (G N^2)   dB,   DPSNR_C = 3e-5
(9.5e-12 dB at 1024^2 x 3, i.e. 7.6e-9 reward units; 20 000 x below the floor round 1 used).
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import binary_hologram_reinforcement_learning_b200 as bh
from binary_hologram_reinforcement_learning_b200.engine import RESULT_DTYPE, RULE_DBS, RULE_ENV, RULE_NEVER
from oracle import hologram_oracle as O
from tests.golden import make_golden as MG

PSNR_TOL = 1e-3          # dB, north_star
DPSNR_C = 3e-5           # see the module docstring


def dpsnr_floor(N, G):
    """Absolute error bound (dB) of one flip's dPSNR from the fp32 fields."""
    return DPSNR_C / (G * N * N)


def reward_tol(r_ref, N, G):
    return 1e-5 * abs(r_ref) + 800.0 * dpsnr_floor(N, G)


def near_tie(N, G):
    """|dPSNR| (dB) below which the fp32 fields may flip an accept decision."""
    return 4.0 * dpsnr_floor(N, G)


def _problem(N, F, wl, seed):
    pre, tgt = bh.synthetic_problem(N, F, len(wl), seed)
    return pre, tgt, (pre >= 0.5).astype(np.int8)


def _engine(N, F, wl, pad=1, relative=True, n_env=1):
    return bh.HoloEngine(N, F, wl, n_env=n_env, pad=pad, relative=relative)


# ---------------------------------------------------------------------------
# propagation: tt.simulate + abs**2 + mean + relativeLoss
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("N,F,wl,pad", [
    (32, 4, O.WL_MONO, 1), (64, 8, O.WL_MONO, 1), (128, 6, O.WL_RGB, 1), (256, 8, O.WL_MONO, 1),
    (512, 3, O.WL_RGB, 1), (896, 3, O.WL_RGB, 1), (1024, 3, O.WL_RGB, 1),
    (32, 4, O.WL_MONO, 2), (64, 6, O.WL_RGB, 2), (128, 2, O.WL_MONO, 2), (256, 2, O.WL_MONO, 2),
    (896, 1, O.WL_MONO, 2), (1024, 1, O.WL_MONO, 2),
])
def test_propagation_matches_oracle(N, F, wl, pad):
    pre, tgt, st = _problem(N, F, wl, seed=N + F)
    cfg = O.HoloConfig(N=N, F=F, wl=wl, pad=pad)
    eng = _engine(N, F, wl, pad)
    eng.set_target(0, tgt)
    eng.load_state(0, st)
    recon_ref = O.reconstruct(cfg, st)
    psnr_ref, mse_ref = O.score(cfg, recon_ref, tgt)
    psnr, mse, sums = eng.metrics(0)
    assert abs(psnr - psnr_ref) < 1e-4, (psnr, psnr_ref)
    np.testing.assert_allclose(sums, O.loss_sums(recon_ref, tgt), rtol=2e-6)
    np.testing.assert_allclose(eng.recon(0), recon_ref, rtol=0, atol=2e-5 * recon_ref.max())
    for f in sorted({0, F - 1}):
        g = cfg.group_of(f)
        U_ref = O.propagate_group(cfg, st[g * cfg.Fg:(g + 1) * cfg.Fg], g)[f - g * cfg.Fg]
        U = eng.field(0, f)
        assert np.abs(U - U_ref).max() < 1e-5 * np.abs(U_ref).max()
    assert np.array_equal(eng.state(0), st)
    eng.close()


def test_simulate_operator_matches_oracle():
    rng = np.random.default_rng(0)
    for N, pad in [(64, 1), (256, 1), (64, 2)]:
        x = rng.random((3, N, N)).astype(np.float32)
        H = O.transfer_function(N * pad, O.PIXEL_PITCH, 515e-9, O.Z_DEFAULT)
        ref = O.simulate(x.astype(np.float64), H, pad)
        got = bh.simulate(x, 515e-9, pad=pad)
        assert np.abs(got - ref).max() < 1e-5 * np.abs(ref).max()
        xc = (x + 1j * rng.random((3, N, N))).astype(np.complex64)
        refc = O.simulate(xc.astype(np.complex128), H, pad)
        assert np.abs(bh.simulate(xc, 515e-9, pad=pad) - refc).max() < 1e-5 * np.abs(refc).max()


def test_unsupported_shapes_fail_loudly():
    with pytest.raises(bh.HoloError):
        bh.HoloEngine(96, 8, bh.WL_MONO)          # FFT side 96 not in the plan table
    with pytest.raises(bh.HoloError):
        bh.HoloEngine(64, 7, bh.WL_RGB)           # F not a multiple of G
    eng = _engine(32, 4, O.WL_MONO)
    with pytest.raises(bh.HoloError):
        eng.eval_flips(np.array([4 * 32 * 32]))   # action out of range
    with pytest.raises(bh.HoloError):
        eng.step_batch(np.array([0, 1]), np.array([0, 0]))   # same env twice
    eng.close()


# ---------------------------------------------------------------------------
# candidate scoring (env_group.py:96-119, dbs-...-6464.py:337-371)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("N,F,wl,pad,relative", [
    (64, 8, O.WL_MONO, 1, True), (64, 6, O.WL_RGB, 1, True), (32, 4, O.WL_MONO, 2, True),
    (64, 8, O.WL_MONO, 1, False), (128, 6, O.WL_RGB, 1, True),
])
def test_eval_flips_matches_full_resimulation(N, F, wl, pad, relative):
    pre, tgt, st = _problem(N, F, wl, seed=7)
    cfg = O.HoloConfig(N=N, F=F, wl=wl, pad=pad, relative=relative)
    eng = _engine(N, F, wl, pad, relative)
    eng.set_target(0, tgt)
    eng.load_state(0, st)
    rng = np.random.default_rng(3)
    actions = np.concatenate([rng.integers(0, F * N * N, size=60),
                              [0, N - 1, N * N - 1, F * N * N - 1, (F - 1) * N * N + N * (N - 1)]])
    got = eng.eval_flips(actions)
    ref, p0, *_ = O.sweep(cfg, st, tgt, pre, actions)
    psnr0 = eng.metrics(0)[0]
    assert np.abs(got - ref).max() < 1e-4
    d_got, d_ref = got - psnr0, ref - p0
    np.testing.assert_allclose(d_got, d_ref, rtol=1e-4, atol=1e-7)
    assert np.array_equal(eng.state(0), st)                      # score-and-revert: state untouched
    assert abs(eng.metrics(0)[0] - psnr0) == 0
    eng.close()


@pytest.mark.parametrize("N,F,wl,relative,pad", [
    (64, 8, O.WL_MONO, True, 1), (64, 6, O.WL_RGB, True, 1), (128, 6, O.WL_RGB, False, 1),
    (256, 8, O.WL_MONO, True, 1), (64, 8, O.WL_MONO, True, 2), (128, 6, O.WL_RGB, True, 2),
])
def test_sweep_all_correlation_matches_oracle_and_delta_kernel(N, F, wl, relative, pad):
    """bh_sweep_all (FFT correlations) == bh_eval_flips (delta kernel) == full re-simulation."""
    pre, tgt, st = _problem(N, F, wl, seed=23)
    cfg = O.HoloConfig(N=N, F=F, wl=wl, relative=relative, pad=pad)
    eng = _engine(N, F, wl, pad, relative)
    eng.set_target(0, tgt)
    eng.load_state(0, st)
    psnr0 = eng.metrics(0)[0]
    pm = eng.sweep_all(0)
    assert pm.shape == (F, N, N) and np.all(np.isfinite(pm))
    rng = np.random.default_rng(4)
    actions = np.concatenate([rng.integers(0, F * N * N, size=300), [0, N * N - 1, F * N * N - 1]])
    d_map = pm.reshape(-1)[actions] - psnr0
    d_delta = eng.eval_flips(actions) - psnr0
    np.testing.assert_allclose(d_map, d_delta, rtol=2e-4, atol=2e-8)
    ref, p0, *_ = O.sweep(cfg, st, tgt, pre, actions[:40])
    np.testing.assert_allclose(d_map[:40], ref - p0, rtol=2e-4, atol=1e-7)
    # every pixel scored: the improving fraction agrees with the delta kernel on a sample
    assert abs((d_map > 0).mean() - (d_delta > 0).mean()) < 1e-9
    assert np.array_equal(eng.state(0), st) and eng.metrics(0)[0] == psnr0
    # the sweep driver picks the map for an exhaustive order and the decile stats agree
    order = rng.permutation(F * N * N)
    r_map = bh.sweep_engine(eng, 0, pre, order, psnr0)
    r_del = bh.sweep_engine(eng, 0, pre, order[:2000], psnr0, exhaustive=False)
    np.testing.assert_allclose(r_map["psnr_after"][:2000], r_del["psnr_after"], rtol=0, atol=1e-8)
    assert r_map["attempted"].sum() == F * N * N
    eng.close()


# ---------------------------------------------------------------------------
# golden fixtures (tests/golden/*.npz, made by the float64 oracle)
# ---------------------------------------------------------------------------
def _decisions_equal_up_to_near_ties(acc, acc_ref, delta_ref, N, G):
    """Greedy runs keep thousands of flips without re-propagating; the incrementally updated fields collect
    one fp32 rounding per kept flip and frame, so the near-tie band of a long run is wider than one flip's."""
    bad = np.flatnonzero(np.asarray(acc, bool) != np.asarray(acc_ref, bool))
    return all(abs(delta_ref[i]) < 25.0 * dpsnr_floor(N, G) for i in bad), bad


@pytest.mark.parametrize("name", list(MG.CASES))
def test_golden_env_trajectory(name, golden_dir):
    """env.py:154-260 semantics through BinaryHologramEnv.step on the CUDA engine."""
    N, F, wl, pad, relative, seed = MG.CASES[name]
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    ld = bh.SyntheticLoader(N, F, len(wl), seeds=(seed,))
    env = bh.BinaryHologramEnv(ld.target_function, ld, max_steps=MG.N_STEPS - 10, T_PSNR_DIFF=1e9,
                               IPS=N, CH=F, wl=wl, pad=pad, relative=relative, verbose=False)
    obs, info = env.reset()
    assert abs(env.initial_psnr - float(g["initial_psnr"])) < 1e-4
    assert obs["recon_image"].shape == (1, len(wl), N, N) and obs["state"].dtype == np.int8
    prev = float(g["initial_psnr"])
    for i, a in enumerate(g["env_actions"]):
        flips_before = env.flip_count
        obs, r, term, trunc, _ = env.step(int(a))
        acc_ref = bool(g["env_accepted"][i])
        d_ref = float(g["env_psnr"][i]) - prev
        assert isinstance(r, float) and isinstance(term, bool) and isinstance(trunc, bool)
        got_acc = env.flip_count == flips_before + 1      # kept flips count (env.py:167,194)
        if got_acc != acc_ref:
            assert abs(d_ref) < near_tie(N, len(wl)), (i, d_ref)
            pytest.skip("near-tie divergence: trajectories legitimately differ from here")
        r_ref = float(g["env_rewards"][i])
        assert abs(r - r_ref) <= reward_tol(r_ref, N, len(wl)), (i, r, r_ref)
        assert term == bool(g["env_terminated"][i])
        if acc_ref:
            prev = float(g["env_psnr"][i])
    assert int(env.state.sum()) == int(g["env_final_state_sum"])
    assert np.array_equal(env.engine.state(0), env.state[0])
    env.close()


@pytest.mark.parametrize("name", list(MG.CASES))
@pytest.mark.parametrize("k_spec", [1, 0, 16, 128])    # 128: frame-sorted, bundled windows
def test_golden_dbs_greedy(name, k_spec, golden_dir):
    """DBS.py:247-294: the device loop reproduces the sequential accept sequence for any batch depth."""
    N, F, wl, pad, relative, seed = MG.CASES[name]
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    pre, tgt, st = _problem(N, F, wl, seed)
    eng = _engine(N, F, wl, pad, relative)
    eng.set_target(0, tgt)
    eng.load_state(0, st)
    acc, trace, nacc, final = eng.dbs_run(g["dbs_order"], k_spec=k_spec, trace=True)
    ref_acc, ref_trace = g["dbs_accepted"], g["dbs_trace"]
    prev = np.concatenate([[float(g["initial_psnr"])], ref_trace[:-1]])
    running = np.maximum.accumulate(np.concatenate([[float(g["initial_psnr"])], np.where(ref_acc, ref_trace, -np.inf)]))[:-1]
    ok, bad = _decisions_equal_up_to_near_ties(acc, ref_acc, ref_trace - running, N, len(wl))
    assert ok, bad
    if bad.size == 0:
        np.testing.assert_allclose(trace, ref_trace, rtol=0, atol=1e-4)
        assert nacc == int(ref_acc.sum())
        assert int(eng.state(0).sum()) == int(g["dbs_final_state_sum"])
        assert abs(final - ref_trace[ref_acc][-1]) < 1e-4
    eng.close()


@pytest.mark.parametrize("name", list(MG.CASES))
def test_golden_sweep(name, golden_dir):
    N, F, wl, pad, relative, seed = MG.CASES[name]
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    pre, tgt, st = _problem(N, F, wl, seed)
    eng = _engine(N, F, wl, pad, relative)
    eng.set_target(0, tgt)
    eng.load_state(0, st)
    r = bh.sweep_engine(eng, 0, pre, g["sweep_order"], eng.metrics(0)[0])
    np.testing.assert_allclose(r["psnr_after"], g["sweep_psnr"], rtol=0, atol=1e-4)
    assert np.array_equal(r["attempted"], g["sweep_attempted"])
    d_ref = g["sweep_psnr"] - float(g["initial_psnr"])
    if np.all(np.abs(d_ref) > near_tie(N, len(wl))):
        assert np.array_equal(r["improved"], g["sweep_improved"])
        np.testing.assert_allclose(r["gains"], g["sweep_gain"], rtol=1e-4, atol=1e-7)
    eng.close()


# ---------------------------------------------------------------------------
# vectorised envs, group reward, cropped env
# ---------------------------------------------------------------------------
def test_vec_env_matches_independent_oracle_envs():
    N, F, wl, E = 64, 6, O.WL_RGB, 4
    loaders = [bh.SyntheticLoader(N, F, 3, seeds=(100 + i,)) for i in range(E)]
    tf = lambda t: next(l for l in loaders if np.ascontiguousarray(t[0, 0, 0, :4]).tobytes() in l._pre).target_function(t)
    vec = bh.HologramVecEnv(E, tf, loaders, max_steps=10 ** 6, T_PSNR_DIFF=1e9, IPS=N, CH=F, wl=wl)
    vec.reset()
    cfg = O.HoloConfig(N=N, F=F, wl=wl)
    refs = []
    for i in range(E):
        pre, tgt = bh.synthetic_problem(N, F, 3, 100 + i)
        e = O.OracleEnv(cfg, max_steps=10 ** 6, T_PSNR_DIFF=1e9)
        e.reset(pre, tgt)
        refs.append(e)
        assert abs(vec.envs[i].initial_psnr - e.initial_psnr) < 1e-4
    assert vec._fast
    rng = np.random.default_rng(5)
    for step in range(40):
        acts = rng.integers(0, F * N * N, size=E)
        obs, rewards, dones, infos = vec.step(acts)
        for i in range(E):
            r, term, trunc, p, acc = refs[i].step(int(acts[i]))
            assert abs(rewards[i] - r) <= reward_tol(r, N, len(wl))
            assert not dones[i]
    vec.sync_envs()
    for i in range(E):
        assert np.array_equal(vec.envs[i].state[0], refs[i].state)
        assert np.array_equal(vec.envs[i].state_record[0], refs[i].state_record)
        assert np.array_equal(vec.engine.state(i), refs[i].state)
        assert vec.envs[i].flip_count == refs[i].flip_count and vec.envs[i].steps == refs[i].steps
        rec = vec.refresh_recon(i)[0]
        np.testing.assert_allclose(rec, refs[i].recon, atol=3e-5 * refs[i].recon.max())
    vec.close()


@pytest.mark.parametrize("verbose", [False, True])
def test_vec_env_episode_end_bonus_and_autoreset(verbose, capsys):
    """max_steps bonus (env.py:237-254), SB3-style auto-reset; fast (numpy) and per-env paths agree."""
    N, F, wl, E, MAXS = 64, 8, O.WL_MONO, 3, 12
    loaders = [bh.SyntheticLoader(N, F, 1, seeds=(200 + i,)) for i in range(E)]
    tf = lambda t: next(l for l in loaders if np.ascontiguousarray(t[0, 0, 0, :4]).tobytes() in l._pre).target_function(t)
    vec = bh.HologramVecEnv(E, tf, loaders, max_steps=MAXS, T_PSNR_DIFF=1e9, IPS=N, CH=F, wl=wl,
                            verbose=verbose)
    assert vec._fast == (not verbose)
    vec.reset()
    cfg = O.HoloConfig(N=N, F=F, wl=wl)
    probs = [bh.synthetic_problem(N, F, 1, 200 + i) for i in range(E)]
    refs = []
    for i in range(E):
        e = O.OracleEnv(cfg, max_steps=MAXS, T_PSNR_DIFF=1e9)
        e.reset(*probs[i])
        refs.append(e)
    rng = np.random.default_rng(8)
    n_done = 0
    for step in range(40):
        acts = rng.integers(0, F * N * N, size=E)
        obs, rewards, dones, infos = vec.step(acts)
        for i in range(E):
            r, term, trunc, p, acc = refs[i].step(int(acts[i]))
            assert abs(rewards[i] - r) <= reward_tol(r, N, len(wl)), (step, i, rewards[i], r)
            assert bool(dones[i]) == bool(term or trunc)
            if dones[i]:
                n_done += 1
                assert "terminal_observation" in infos[i] and infos[i]["TimeLimit.truncated"] is False
                assert infos[i]["terminal_observation"]["state"].sum() == refs[i].state.sum()
                refs[i].reset(*probs[i])
                assert np.array_equal(obs[i]["state"][0], refs[i].state)
    assert n_done >= E and len(vec.episode_stats) == n_done
    capsys.readouterr()
    vec.close()


def test_group_env_importance_table():
    N, F = 64, 8
    ld = bh.SyntheticLoader(N, F, 1, seeds=(31,))
    env = bh.BinaryHologramEnvGroup(ld.target_function, ld, IPS=N, CH=F, verbose=False,
                                    num_samples=300, rng=np.random.default_rng(9))
    env.reset()
    cfg = O.HoloConfig(N=N, F=F)
    pre, tgt = bh.synthetic_problem(N, F, 1, 31)
    st = (pre >= 0.5).astype(np.int8)
    p0_ref, _ = O.score(cfg, O.reconstruct(cfg, st), tgt)
    assert abs(env.initial_psnr - p0_ref) < 1e-4
    ch, ranks, pos = O.pixel_importance(cfg, st, tgt, p0_ref, None, 300,
                                        actions=env.importance_actions)
    np.testing.assert_allclose(env.psnr_change_list, ch, rtol=1e-4, atol=1e-7)
    assert abs(env.T_PSNR_DIFF - pos / 4) < 1e-5 * pos
    if np.min(np.abs(np.diff(np.sort(ch)))) > 1e-7:
        np.testing.assert_allclose(env.importance_ranks, ranks, atol=1e-9)
    obs, r, term, trunc, _ = env.step(int(env.importance_actions[0]))
    assert -0.6 < r < 1.1                                           # a table value (env_group.py:255)
    env.close()


def test_cropped_env_simulates_the_window_only():
    """env_1024_24_128.py:139-181 at a small shape: crop 8 of 64 -> 48^2 simulated."""
    N, F, m = 64, 6, 8
    ld = bh.SyntheticLoader(N, F, 3, seeds=(41,))
    env = bh.BinaryHologramEnv(ld.target_function, ld, IPS=N, CH=F, wl=O.WL_RGB, crop_margin=m,
                               verbose=False, T_PSNR_DIFF=1e9)
    with pytest.raises(bh.HoloError):
        env.reset()                                                 # 48 is not a supported FFT side
    env.close()
    N, m = 80, 8                                                    # 80 - 16 = 64
    ld = bh.SyntheticLoader(N, F, 3, seeds=(41,))
    env = bh.BinaryHologramEnv(ld.target_function, ld, IPS=N, CH=F, wl=O.WL_RGB, crop_margin=m,
                               verbose=False, T_PSNR_DIFF=1e9)
    obs, _ = env.reset()
    pre, tgt = bh.synthetic_problem(N, F, 3, 41)
    cfg = O.HoloConfig(N=N - 2 * m, F=F, wl=O.WL_RGB)
    st = (pre >= 0.5).astype(np.int8)[:, m:-m, m:-m]
    p_ref, _ = O.score(cfg, O.reconstruct(cfg, st), tgt[:, m:-m, m:-m])
    assert abs(env.initial_psnr - p_ref) < 1e-4
    obs, r, *_ = env.step(0)                                        # outside the window: no change
    assert r == 0.0 and obs["state"][0, 0, 0, 0] != (pre[0, 0, 0] >= 0.5)
    env.close()


# ---------------------------------------------------------------------------
# size-independent properties at the full BASELINE shape (1024^2 x 24, 3 colours)
# ---------------------------------------------------------------------------
def test_full_size_properties():
    N, F, wl = 1024, 24, O.WL_RGB
    pre, tgt, st = _problem(N, F, wl, seed=0)
    eng = _engine(N, F, wl, n_env=2)
    for e in range(2):
        eng.set_target(e, tgt)
        eng.load_state(e, st)
    p0, mse0, sums0 = eng.metrics(0)
    assert eng.metrics(1)[0] == p0                                  # same inputs -> same bits
    # one colour group against the oracle (full 24-frame oracle run is too slow for a unit test)
    cfg = O.HoloConfig(N=N, F=8, wl=(wl[1],), dtype="float64")
    rec_g = O.reconstruct(cfg, st[8:16])[0]
    np.testing.assert_allclose(eng.recon(0)[1], rec_g, atol=2e-5 * rec_g.max())
    rng = np.random.default_rng(1)
    acts = rng.integers(0, F * N * N, size=64)
    # (1) scoring does not depend on batch composition or order
    a = eng.eval_flips(acts)
    b = eng.eval_flips(acts[::-1])[::-1]
    assert np.array_equal(a, b)
    c = np.array([eng.eval_flips(acts[i:i + 1])[0] for i in range(8)])
    assert np.array_equal(a[:8], c)
    # (2) delta evaluation == full re-propagation of the flipped state
    for i in range(3):
        st2 = st.copy().reshape(-1); st2[acts[i]] = 1 - st2[acts[i]]
        eng.load_state(1, st2.reshape(F, N, N))
        assert abs(eng.metrics(1)[0] - a[i]) < 2e-6, (eng.metrics(1)[0], a[i])
    eng.load_state(1, st)
    # (3) commit + commit of the same pixel is the identity (up to fp32 rounding of U)
    eng.commit_flip(0, int(acts[0]))
    assert abs(eng.metrics(0)[0] - a[0]) < 1e-9
    assert eng.state(0).reshape(-1)[acts[0]] != st.reshape(-1)[acts[0]]
    eng.commit_flip(0, int(acts[0]))
    assert abs(eng.metrics(0)[0] - p0) < 1e-7
    assert np.array_equal(eng.state(0), st)
    # (4) a batch step over both envs == two single steps; accepted flips stick, rejected revert
    res = eng.step_batch(acts[:2], np.array([0, 1]), RULE_ENV)
    for e in range(2):
        flipped = eng.state(e).reshape(-1)[acts[e]] != st.reshape(-1)[acts[e]]
        assert flipped == bool(res["accept"][e])
        assert bool(res["accept"][e]) == (res["psnr_after"][e] - p0 >= 0)
        if res["accept"][e]:
            assert eng.metrics(e)[0] == res["psnr_after"][e]
    # (5) resync after incremental updates agrees with the running sums
    order = rng.permutation(F * N * N)[:400]
    acc, tr, nacc, fin = eng.dbs_run(order, env=0, k_spec=0, trace=True)
    assert nacc == int(acc.sum()) and nacc > 50
    before = eng.metrics(0)
    eng.resync(0)
    after = eng.metrics(0)
    assert abs(before[0] - after[0]) < 5e-6
    np.testing.assert_allclose(before[2], after[2], rtol=1e-6)
    # accepted candidates strictly increase the PSNR trace
    inc = tr[acc.astype(bool)]
    assert np.all(np.diff(inc) > 0)
    eng.close()


# ---------------------------------------------------------------------------
# drop-in shims: the reference's own call sequence (env.py:123-132) through compat/torchOptics
# ---------------------------------------------------------------------------
def test_reference_call_sequence_through_compat_shims():
    import sys
    import torch
    compat = os.path.join(os.path.dirname(bh.__file__), "compat")
    sys.path.insert(0, compat)
    try:
        import torchOptics.optics as tt
        import torchOptics.metrics as tm
        import env as ref_env_module
    finally:
        sys.path.remove(compat)
    assert ref_env_module.BinaryHologramEnv is bh.BinaryHologramEnv and ref_env_module.IPS == 256
    N, F = 256, 8
    pre, tgt = bh.synthetic_problem(N, F, 1, 5)
    state = (pre >= 0.5).astype(np.int8)[None]
    # --- verbatim shape of env.py:123-132 ---
    binary = torch.tensor(state, dtype=torch.float32).cuda()
    binary = tt.Tensor(binary, meta={'dx': (7.56e-6, 7.56e-6), 'wl': 515e-9})
    sim = tt.simulate(binary, 2e-3).abs() ** 2
    result = torch.mean(sim, dim=1, keepdim=True)
    target = torch.from_numpy(tgt[None]).cuda()
    psnr = tt.relativeLoss(result, target, tm.get_PSNR)
    # ---
    cfg = O.HoloConfig(N=N, F=F)
    psnr_ref, _ = O.score(cfg, O.reconstruct(cfg, state[0]), tgt)
    assert isinstance(psnr, float) and abs(psnr - psnr_ref) < PSNR_TOL
    # numpy int8 input as in env_1024_24.py:149-151
    red = tt.Tensor(state[:, :4], meta={'wl': (638e-9), 'dx': (7.56e-6, 7.56e-6)})
    U = tt.simulate(red, 2e-3)
    H = O.transfer_function(N, O.PIXEL_PITCH, 638e-9, O.Z_DEFAULT)
    ref = O.simulate(state[0, :4].astype(np.float64), H, 1)
    assert np.abs(U.cpu().numpy()[0] - ref).max() < 1e-5 * np.abs(ref).max()


# ---------------------------------------------------------------------------
# DBS drivers (reference entry points) incl. checkpoint / resume
# ---------------------------------------------------------------------------
def test_dbs_greedy_driver_and_resume(golden_dir, tmp_path):
    """optimize_with_random_pixel_flips(env): DBS.py:202-307 through the env, resumable."""
    N, F, wl, pad, relative, seed = MG.CASES["mono64"]
    g = np.load(os.path.join(golden_dir, "mono64.npz"))
    order = g["dbs_order"]

    def make_env():
        ld = bh.SyntheticLoader(N, F, 1, seeds=(seed,))
        return bh.BinaryHologramEnv(ld.target_function, ld, IPS=N, CH=F, verbose=False)

    env = make_env()
    res = bh.optimize_with_random_pixel_flips(env, max_datasets=0, order=order, verbose=False)
    assert len(res) == 1 and res[0]["complete"]
    r = res[0]
    assert np.array_equal(r["accepted"].astype(bool), g["dbs_accepted"])
    assert r["flip_count"] == int(g["dbs_accepted"].sum())
    assert int(env.state.sum()) == int(g["dbs_final_state_sum"])          # host mirror follows the device
    assert np.array_equal(env.engine.state(0), env.state[0])
    assert r["improved_bin_counts"].sum() == r["flip_count"]
    env.close()
    # same run in two calls through a checkpoint
    ck = str(tmp_path / "dbs_ck.npz")
    env = make_env()
    part = bh.dbs_greedy_env(env, max_datasets=0, order=order, verbose=False, segment=64,
                             checkpoint=ck, max_segments=2)[0]
    assert not part["complete"] and part["steps"] == 128 and os.path.exists(ck)
    env.close()
    env = make_env()                                                       # fresh process stand-in
    rest = bh.dbs_greedy_env(env, max_datasets=0, order=order, verbose=False, segment=64, checkpoint=ck)[0]
    assert rest["complete"] and np.array_equal(rest["accepted"], r["accepted"])
    np.testing.assert_allclose(rest["psnr_trace"], r["psnr_trace"], rtol=0, atol=1e-6)
    assert np.array_equal(rest["state"], r["state"])
    env.close()


def test_dbs_sweep_driver_with_crop():
    """dbs-1024-1024-24-6464.py:194-478 at a small shape: crop 8 of 80 -> 64^2 x 6, all pixels scored."""
    N, F, m = 80, 6, 8
    ld = bh.SyntheticLoader(N, F, 3, seeds=(61,))
    res = bh.optimize_with_random_pixel_flips(ld.target_function, ld, 2e-3, 7.56e-6, m, CH=F,
                                              max_datasets=0, rng=np.random.default_rng(1), verbose=False)
    assert len(res) == 1
    r = res[0]
    n = F * 64 * 64
    assert r["steps"] == n and r["attempted"].sum() == n and np.array_equal(r["attempted"], r["bin_counts"])
    pre, tgt = bh.synthetic_problem(N, F, 3, 61)
    cfg = O.HoloConfig(N=64, F=F, wl=O.WL_RGB)
    st = (pre >= 0.5).astype(np.int8)[:, m:-m, m:-m]
    sub = r["order"][:50]
    ref, p0, *_ = O.sweep(cfg, st, tgt[:, m:-m, m:-m], pre[:, m:-m, m:-m], sub)
    assert abs(r["initial_psnr"] - p0) < 1e-4
    np.testing.assert_allclose(r["psnr_after"][:50] - r["initial_psnr"], ref - p0, rtol=2e-4, atol=1e-7)
    assert 0 < r["flip_count"] < n and r["improved"].sum() == r["flip_count"]


# ---------------------------------------------------------------------------
# golden fixtures at the BASELINE shapes (tests/golden/large_*.npz, make_golden_large.py)
# ---------------------------------------------------------------------------
from tests.golden import make_golden_large as MGL  # noqa: E402


def _need(golden_dir, name):
    path = os.path.join(golden_dir, f"{name}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{name}.npz not generated")
    return np.load(path)


@pytest.mark.parametrize("name", list(MGL.CASES))
def test_large_golden(name, golden_dir):
    g = _need(golden_dir, name)
    N, F, wl, seed, n_steps, n_dbs, n_sweep = MGL.CASES[name]
    G = len(wl)
    ld = bh.SyntheticLoader(N, F, G, seeds=(seed,))
    env = bh.BinaryHologramEnv(ld.target_function, ld, max_steps=n_steps - 20, T_PSNR_DIFF=1e9,
                               IPS=N, CH=F, wl=wl, verbose=False, recon_obs="lazy")
    env.reset()
    eng = env.engine
    # propagation
    psnr0, mse0, sums = eng.metrics(0)
    assert abs(psnr0 - float(g["initial_psnr"])) < 1e-4
    assert abs(mse0 - float(g["initial_mse"])) < 1e-5 * float(g["initial_mse"])
    np.testing.assert_allclose(sums, g["loss_sums"], rtol=2e-6)
    rec = eng.recon(0)[:, ::N // 8, ::N // 8]
    np.testing.assert_allclose(rec, g["recon_samples"], atol=2e-5 * g["recon_samples"].max())
    for i, f in enumerate(g["field_frames"]):
        U = eng.field(0, int(f))[::N // 8, ::N // 8]
        assert np.abs(U - g["field_samples"][i]).max() < 1e-5 * np.abs(g["field_samples"][i]).max()
    # sweep: delta kernel and correlation map against the oracle
    d_ref = g["sweep_psnr"] - float(g["initial_psnr"])
    d_eval = eng.eval_flips(g["sweep_order"]) - psnr0
    assert np.all(np.abs(d_eval - d_ref) <= 1e-5 * np.abs(d_ref) + dpsnr_floor(N, G))
    d_map = eng.sweep_all(0).reshape(-1)[g["sweep_order"]] - psnr0
    assert np.all(np.abs(d_map - d_ref) <= 3e-5 * np.abs(d_ref) + 3.5 * dpsnr_floor(N, G))
    # greedy DBS prefix on a second context (the env keeps its own state for the trajectory)
    eng2 = bh.HoloEngine(N, F, wl)
    pre, tgt = bh.synthetic_problem(N, F, G, seed)
    eng2.set_target(0, tgt)
    eng2.load_state(0, (pre >= 0.5).astype(np.int8))
    acc, tr, nacc, fin = eng2.dbs_run(g["dbs_order"], trace=True)
    prev = np.maximum.accumulate(np.concatenate([[float(g["initial_psnr"])],
                                                 np.where(g["dbs_accepted"], g["dbs_trace"], -np.inf)]))[:-1]
    ok, bad = _decisions_equal_up_to_near_ties(acc, g["dbs_accepted"], g["dbs_trace"] - prev, N, G)
    assert ok, bad
    if bad.size == 0:
        np.testing.assert_allclose(tr, g["dbs_trace"], rtol=0, atol=1e-4)
        assert int(eng2.state(0).sum()) == int(g["dbs_final_state_sum"])
    eng2.close()
    # env trajectory incl. the max_steps bonus
    prev = float(g["initial_psnr"])
    for i, a in enumerate(g["env_actions"]):
        flips_before = env.flip_count
        obs, r, term, trunc, _ = env.step(int(a))
        acc_ref = bool(g["env_accepted"][i])
        if (env.flip_count == flips_before + 1) != acc_ref:
            assert abs(float(g["env_psnr"][i]) - prev) < near_tie(N, G), i
            break                                            # legitimately diverged on a near-tie
        r_ref = float(g["env_rewards"][i])
        assert abs(r - r_ref) <= reward_tol(r_ref, N, G), (i, r, r_ref)
        assert term == bool(g["env_terminated"][i])
        if acc_ref:
            prev = float(g["env_psnr"][i])
    else:
        assert int(env.state.sum()) == int(g["env_final_state_sum"])
    env.close()


def test_large_golden_dbs64_exhaustive(golden_dir):
    """Every pixel of a 64^2 x 8 stack visited once (DBS.py:243-294): 32768 sequential decisions."""
    g = _need(golden_dir, "large_dbs64_exhaustive")
    N, F, seed = 64, 8, 41
    pre, tgt, st = _problem(N, F, O.WL_MONO, seed)
    eng = _engine(N, F, O.WL_MONO)
    eng.set_target(0, tgt)
    eng.load_state(0, st)
    acc, _, nacc, fin = eng.dbs_run(g["order"], resync_every=1024)
    ref = np.unpackbits(g["accepted"])[:int(g["n"])]
    mism = int(np.count_nonzero(acc != ref))
    assert mism <= 0.002 * ref.size, mism                   # near-tie divergences only
    assert abs(fin - float(g["final_psnr"])) < 5e-3
    assert abs(nacc - int(g["n_accepted"])) <= 0.002 * ref.size
    if mism == 0:
        assert np.array_equal(np.packbits(eng.state(0).astype(np.uint8)), g["final_state"])
    eng.close()


def test_large_golden_dbs256_first_50k(golden_dir):
    """first 50 000 candidates of the greedy DBS of a 256^2 x 8 stack (DBS.py:243-294)."""
    g = _need(golden_dir, "large_dbs256_50k")
    N, F, seed = 256, 8, 61
    pre, tgt, st = _problem(N, F, O.WL_MONO, seed)
    eng = _engine(N, F, O.WL_MONO)
    eng.set_target(0, tgt)
    eng.load_state(0, st)
    acc, _, nacc, fin = eng.dbs_run(g["order"], resync_every=1024)
    ref = np.unpackbits(g["accepted"])[:int(g["n"])]
    mism = int(np.count_nonzero(acc != ref))
    assert mism <= 0.002 * ref.size, mism                   # near-tie divergences only
    assert abs(fin - float(g["final_psnr"])) < 5e-3
    assert abs(nacc - int(g["n_accepted"])) <= 0.002 * ref.size
    if mism == 0:
        assert np.array_equal(np.packbits(eng.state(0).astype(np.uint8)), g["final_state"])
    eng.close()


def test_large_golden_group256(golden_dir):
    """env_group.py:90-143 at 256^2 x 8 with the reference's 10 000 candidates."""
    g = _need(golden_dir, "large_group256")
    N, F, seed = 256, 8, 51
    ld = bh.SyntheticLoader(N, F, 1, seeds=(seed,))
    env = bh.BinaryHologramEnvGroup(ld.target_function, ld, IPS=N, CH=F, verbose=False,
                                    rng=np.random.default_rng(seed))
    env.reset()
    assert np.array_equal(env.importance_actions, g["actions"])
    np.testing.assert_allclose(env.psnr_change_list, g["psnr_changes"], rtol=2e-4, atol=2e-8)
    assert abs(env.T_PSNR_DIFF - float(g["positive_sum"]) / 4) < 1e-4 * float(g["positive_sum"])
    # rank table: identical wherever the ordering of the changes is unambiguous
    order_ref = np.argsort(g["psnr_changes"])
    gaps = np.diff(g["psnr_changes"][order_ref])
    if gaps.min() > 1e-9:
        np.testing.assert_allclose(env.importance_ranks, g["ranks"], atol=1e-9)
    else:
        assert np.mean(np.abs(env.importance_ranks - g["ranks"]) > 1e-6) < 0.01
    env.close()


def test_group_vec_env_fast_path_matches_per_env_path():
    """env_group.py reward (rank-table lookup) vectorised across envs == the per-env implementation."""
    N, F, E = 64, 8, 3
    def make(verbose):
        loaders = [bh.SyntheticLoader(N, F, 1, seeds=(300 + i,)) for i in range(E)]
        tf = lambda t: next(l for l in loaders if np.ascontiguousarray(t[0, 0, 0, :4]).tobytes() in l._pre).target_function(t)
        return bh.HologramVecEnv(E, tf, loaders, max_steps=25, IPS=N, CH=F, reward_mode="group",
                                 num_samples=200, seed=5, verbose=verbose)
    import contextlib, io
    fast, slow = make(False), make(True)
    assert fast._fast and not slow._fast
    with contextlib.redirect_stdout(io.StringIO()):
        fast.reset(); slow.reset()
        rng = np.random.default_rng(0)
        for step in range(60):
            acts = rng.integers(0, F * N * N, size=E)
            _, r1, d1, _ = fast.step(acts)
            _, r2, d2, _ = slow.step(acts)
            np.testing.assert_allclose(r1, r2, rtol=1e-12, atol=1e-12)
            assert np.array_equal(d1, d2)
    assert len(fast.episode_stats) == len(slow.episode_stats) > 0
    fast.close(); slow.close()


def test_batch_step_beyond_inline_limit_and_pinned_obs():
    """More than 32 envs per step take the copied-task path (bh_step_batch); same results as the oracle."""
    N, F, E = 32, 4, 40
    eng = _engine(N, F, O.WL_MONO, n_env=E)
    cfg = O.HoloConfig(N=N, F=F)
    refs = []
    for e in range(E):
        pre, tgt = bh.synthetic_problem(N, F, 1, 400 + e)
        eng.set_target(e, tgt)
        eng.load_state(e, (pre >= 0.5).astype(np.int8))
        r = O.OracleEnv(cfg, max_steps=10 ** 9, T_PSNR_DIFF=1e9)
        r.reset(pre, tgt)
        refs.append(r)
    rng = np.random.default_rng(2)
    for step in range(6):
        acts = rng.integers(0, F * N * N, size=E)
        res = eng.step_batch(acts, np.arange(E, dtype=np.int32), RULE_ENV)
        for e in range(E):
            rew, _, _, p, acc = refs[e].step(int(acts[e]))
            assert abs(res["psnr_after"][e] - p) < 1e-4
            if abs(rew) > 800 * 2e-6:
                assert bool(res["accept"][e]) == acc
    from binary_hologram_reinforcement_learning_b200.engine import pinned_empty
    buf = pinned_empty((1, N, N), np.float32)
    eng.recon(3, -1, out=buf)
    np.testing.assert_allclose(buf, refs[3].means, atol=3e-5 * refs[3].means.max())
    eng.close()


def test_sharded_sweep_equals_full_sweep():
    """Candidate ranges sharded over ranks: the concatenated shards are the full sweep, bit for bit."""
    N, F, m = 80, 6, 8
    def run(shard):
        ld = bh.SyntheticLoader(N, F, 3, seeds=(71,))
        return bh.dbs_sweep(ld.target_function, ld, 2e-3, 7.56e-6, m, CH=F, max_datasets=0,
                            rng=np.random.default_rng(9), verbose=False, shard=shard)[0]
    full = run(None)
    parts = [run((r, 3)) for r in range(3)]
    assert np.array_equal(np.concatenate([p["order"] for p in parts]), full["order"])
    assert np.array_equal(np.concatenate([p["psnr_after"] for p in parts]), full["psnr_after"])
    for key in ("attempted", "improved"):
        assert np.array_equal(sum(p[key] for p in parts), full[key])
    # the full sweep takes its statistics on the device (2^-40 fixed-point sums), the shards on the host
    np.testing.assert_allclose(sum(p["gains"] for p in parts), full["gains"], rtol=1e-8, atol=1e-9)
    # a shard small enough to take the delta-kernel path agrees with the correlation map
    small = run((0, 64))
    np.testing.assert_allclose(small["psnr_after"], full["psnr_after"][:small["steps"]], rtol=0, atol=1e-8)


def test_group_rollouts_from_cloned_reset_state():
    """GRPO-style groups: members cloned on the device step exactly like independently reset envs."""
    N, F, E, M = 64, 8, 6, 3
    loaders = [bh.SyntheticLoader(N, F, 1, seeds=(500 + i // M,)) for i in range(E)]
    tf = lambda t: next(l for l in loaders if np.ascontiguousarray(t[0, 0, 0, :4]).tobytes() in l._pre).target_function(t)
    kw = dict(max_steps=10 ** 6, T_PSNR_DIFF=1e9, IPS=N, CH=F, recon_obs="lazy")   # count propagation launches only
    grp = bh.HologramVecEnv(E, tf, loaders, **kw)
    ind = bh.HologramVecEnv(E, tf, [bh.SyntheticLoader(N, F, 1, seeds=(500 + i // M,)) for i in range(E)], **kw)
    launches0 = grp.engine.launch_count
    grp.reset_groups(M)
    assert grp.engine.launch_count - launches0 == (E // M) * 4          # one propagation per group leader
    ind.reset()
    for i in range(E):
        assert grp.envs[i].initial_psnr == ind.envs[i].initial_psnr
        assert np.array_equal(grp.engine.state(i), ind.engine.state(i))
    rng = np.random.default_rng(3)
    for _ in range(25):
        acts = rng.integers(0, F * N * N, size=E)
        _, r1, d1, _ = grp.step(acts)
        _, r2, d2, _ = ind.step(acts)
        assert np.array_equal(r1, r2) and np.array_equal(d1, d2)
    for i in range(E):
        assert np.array_equal(grp.envs[i].state, ind.envs[i].state)
        assert np.array_equal(grp.engine.state(i), ind.engine.state(i))
    grp.close(); ind.close()


def test_multidiscrete_action_equals_flat_action():
    """env_md.py:160: action (channel, row, col) addresses the same pixel as the flat index."""
    N, F = 64, 8
    def make(mode):
        ld = bh.SyntheticLoader(N, F, 1, seeds=(77,))
        e = bh.BinaryHologramEnv(ld.target_function, ld, IPS=N, CH=F, verbose=False, T_PSNR_DIFF=1e9,
                                 action_mode=mode)
        e.reset()
        return e
    a, b = make("multidiscrete"), make("discrete")
    rng = np.random.default_rng(0)
    for _ in range(20):
        ch, r, c = int(rng.integers(F)), int(rng.integers(N)), int(rng.integers(N))
        ra = a.step(np.array([ch, r, c]))
        rb = b.step((ch * N + r) * N + c)
        assert ra[1] == rb[1] and ra[2:4] == rb[2:4]
    assert np.array_equal(a.state, b.state)
    with pytest.raises(ValueError):
        a.step(np.array([F, 0, 0]))
    a.close(); b.close()


@pytest.mark.parametrize("N,F,wl,n", [(64, 6, O.WL_RGB, 301), (256, 8, O.WL_MONO, 301), (896, 3, O.WL_RGB, 601)])
def test_bundled_list_evaluation_is_bit_identical_to_single_candidates(N, F, wl, n):
    """Candidate lists (host list, single-env device window) run through k_eval_bundle_t, which shares the
    U / I / T loads between candidates of a frame; its exact fixed-point sums equal k_eval_t's bit for bit."""
    import torch
    E = 3
    eng = _engine(N, F, wl, n_env=E)
    for e in range(E):
        pre, tgt, st = _problem(N, F, wl, seed=900 + e)
        eng.set_target(e, tgt)
        eng.load_state(e, st)
    rng = np.random.default_rng(5)
    # n is ragged (not a multiple of any bundle size); 896 is the one size that is not row-regular: its
    # lists take the bundled kernel's generic loop from 512 candidates on
    acts = rng.integers(0, F * N * N, size=n)
    acts[10:40] = rng.integers(0, N * N, size=30) + 2 * N * N          # a long same-frame run
    acts[50] = acts[51] = acts[52]                                     # duplicates
    envs = rng.integers(0, E, size=n).astype(np.int32)
    single = np.array([eng.eval_flips(acts[i:i + 1], env_ids=envs[i:i + 1])[0] for i in range(n)])
    listed = eng.eval_flips(acts, env_ids=envs)
    assert np.array_equal(single, listed)
    # one environment, device-resident window of K candidates (the DBS speculation window)
    for K in (2, 7, 64, 128, 300):
        d_act = torch.from_numpy(acts[:K].astype(np.int64)).cuda()
        d_res = torch.zeros(K * 40, dtype=torch.uint8, device="cuda")
        eng.eval_flips_device(K, 0, d_act.data_ptr(), d_res.data_ptr(), env=1)
        torch.cuda.synchronize()
        res = d_res.cpu().numpy().view(RESULT_DTYPE)
        ref = np.array([eng.eval_flips(acts[i:i + 1], env=1)[0] for i in range(K)])
        assert np.array_equal(res["psnr_after"], ref)
        assert np.array_equal(res["action"], acts[:K])
    eng.close()


# ---------------------------------------------------------------------------
# round 2: batched observation path (bh_recon_batch), env.py:176-181
# ---------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["eager", "device"])
def test_observation_is_current_after_every_step(mode):
    """obs["recon_image"] after EVERY step equals the oracle's reconstruction of the evaluated flip (kept or
    rejected, appendix B-2), for every env, through the plane-wise double-buffered observation blocks; the
    observation of step k is still intact after step k+1 (double buffering)."""
    N, F, wl, E = 64, 6, O.WL_RGB, 3
    loaders = [bh.SyntheticLoader(N, F, 3, seeds=(300 + i,)) for i in range(E)]
    tf = lambda t: next(l for l in loaders if np.ascontiguousarray(t[0, 0, 0, :4]).tobytes() in l._pre).target_function(t)
    vec = bh.HologramVecEnv(E, tf, loaders, max_steps=10 ** 6, T_PSNR_DIFF=1e9, IPS=N, CH=F, wl=wl, recon_obs=mode)
    obs = vec.reset()
    cfg = O.HoloConfig(N=N, F=F, wl=wl)
    refs = []
    for i in range(E):
        pre, tgt = bh.synthetic_problem(N, F, 3, 300 + i)
        e = O.OracleEnv(cfg, max_steps=10 ** 6, T_PSNR_DIFF=1e9)
        e.reset(pre, tgt)
        refs.append(e)

    def host(x):
        if isinstance(x, np.ndarray):
            return x
        import torch
        return torch.as_tensor(x, device="cuda").cpu().numpy()

    for i in range(E):
        np.testing.assert_allclose(host(obs[i]["recon_image"])[0], refs[i].recon, atol=3e-5 * refs[i].recon.max())
    rng = np.random.default_rng(8)
    prev_obs, prev_copy, kept, rejected = None, None, 0, 0
    for step in range(30):
        acts = rng.integers(0, F * N * N, size=E)
        obs, rewards, dones, infos = vec.step(acts)
        if prev_obs is not None:              # step k's observation survives step k+1
            for i in range(E):
                assert np.array_equal(host(prev_obs[i]), prev_copy[i])
        for i in range(E):
            _, _, _, _, acc = refs[i].step(int(acts[i]))
            kept += int(acc); rejected += int(not acc)
            rec = host(obs[i]["recon_image"])
            assert rec.shape == (1, 3, N, N)
            np.testing.assert_allclose(rec[0], refs[i].recon, atol=3e-5 * refs[i].recon.max())
        prev_obs = [o["recon_image"] for o in obs]
        prev_copy = [host(p).copy() for p in prev_obs]
    assert kept > 10 and rejected > 10
    vec.close()


@pytest.mark.gpu
def test_single_env_eager_observation_and_reset():
    """Single env (own engine), default recon_obs: the observation follows every step and a reset."""
    N, F = 64, 8
    ld = bh.SyntheticLoader(N, F, 1, seeds=(41, 42))
    env = bh.BinaryHologramEnv(ld.target_function, ld, IPS=N, CH=F, verbose=False, max_steps=10 ** 6, T_PSNR_DIFF=1e9)
    assert env.recon_obs == "eager"
    cfg = O.HoloConfig(N=N, F=F)
    rng = np.random.default_rng(3)
    for ep, seed in enumerate((41, 42)):
        obs, _ = env.reset()
        pre, tgt = bh.synthetic_problem(N, F, 1, seed)
        ref = O.OracleEnv(cfg, max_steps=10 ** 6, T_PSNR_DIFF=1e9)
        ref.reset(pre, tgt)
        np.testing.assert_allclose(obs["recon_image"][0], ref.recon, atol=3e-5 * ref.recon.max())
        for _ in range(12):
            a = int(rng.integers(0, F * N * N))
            obs, *_ = env.step(a)
            ref.step(a)
            np.testing.assert_allclose(obs["recon_image"][0], ref.recon, atol=3e-5 * ref.recon.max())
    env.close()


# ---------------------------------------------------------------------------
# round 2: reward parity at the documented bound, fixed-point range
# ---------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["mono256", "rgb896", "rgb1024"])
def test_reward_parity_bound(name, golden_dir):
    """dPSNR (reward / 800, env.py:184-188) of 2048 random flips per BASELINE shape against the float64 oracle
    (tests/golden/make_flip_oracle.py), through k_eval_t, k_eval_bundle_t and bh_sweep_all: every flip inside
    1e-5 |dPSNR| + DPSNR_C / (G N^2) dB, and the distribution of the relative error as measured
    (profiles/r2_parity_error_dist.json): median < 1.5e-6, p90 < 8e-6, p99 < 8e-5."""
    path = os.path.join(golden_dir, f"flips_{name}.npz")
    if not os.path.exists(path):
        pytest.skip("fixture missing: python tests/golden/make_flip_oracle.py")
    d = np.load(path)
    N, F, wl, seed = int(d["N"]), int(d["F"]), tuple(float(w) for w in d["wl"]), int(d["seed"])
    G = len(wl)
    pre, tgt, st = _problem(N, F, wl, seed)
    eng = _engine(N, F, wl)
    eng.set_target(0, tgt)
    eng.load_state(0, st)
    psnr0 = eng.metrics(0)[0]
    assert abs(psnr0 - float(d["psnr0"])) < 1e-5
    acts, ref = d["actions"], d["psnr_after"] - float(d["psnr0"])
    floor = dpsnr_floor(N, G)
    zero = np.zeros(1, np.int32)
    single = np.array([eng.step_batch(acts[i:i + 1], zero, RULE_NEVER)["psnr_after"][0] for i in range(512)])
    listed = eng.eval_flips(acts)
    assert np.array_equal(listed[:512], single)                      # k_eval_t == k_eval_bundle_t, bit for bit
    err = np.abs((listed - psnr0) - ref)
    assert np.all(err <= 1e-5 * np.abs(ref) + floor), float((err - 1e-5 * np.abs(ref)).max() / floor)
    rel = err / np.abs(ref)
    assert np.median(rel) < 1.5e-6 and np.quantile(rel, 0.9) < 8e-6 and np.quantile(rel, 0.99) < 8e-5
    # decisions of the env rule (keep iff dPSNR >= 0) agree wherever the oracle is outside the near-tie band
    clear = np.abs(ref) > near_tie(N, G)
    assert np.array_equal(((listed - psnr0) >= 0)[clear], (ref >= 0)[clear])
    sw = eng.sweep_all(0).reshape(-1)[acts] - psnr0
    assert np.all(np.abs(sw - ref) <= 3e-5 * np.abs(ref) + 3.5 * floor)
    # the change of the loss sums themselves (relative to the largest change seen)
    res = np.concatenate([eng.step_batch(acts[i:i + 1], zero, RULE_NEVER) for i in range(256)])
    for key in ("d_sii", "d_sit"):
        scale = np.abs(d[key][:256]).max()
        assert np.abs(res[key] - d[key][:256]).max() < 2e-6 * scale
    eng.close()


@pytest.mark.gpu
def test_fixed_point_range_dark_target_and_tiny_fields():
    """The 2^-40 fixed-point sums of the delta kernels (bh_delta.cuh): per-quad terms far below 2^-41 round
    to zero, so a pathological case -- an almost black target (1e-4) and a state with a single lit pixel per
    frame (|U| ~ |h| ~ 1e-2, |dI| ~ 1e-5) -- must stay inside the documented absolute bound N^2 * 2^-43 on
    both sums, and the PSNR change inside the usual band."""
    N, F = 256, 8
    cfg = O.HoloConfig(N=N, F=F)
    rng = np.random.default_rng(17)
    tgt = (1e-4 * rng.random((1, N, N))).astype(np.float32)
    st = np.zeros((F, N, N), np.int8)
    for f in range(F):
        st[f, rng.integers(0, N), rng.integers(0, N)] = 1
    eng = _engine(N, F, O.WL_MONO)
    eng.set_target(0, tgt)
    eng.load_state(0, st)
    U = O.propagate_group(cfg, st, 0)
    I = O.group_mean_intensity(U)
    t64 = tgt.astype(np.float64)
    sii, sit, stt = O.loss_sums(I[None], t64)
    acts = rng.integers(0, F * N * N, size=64)
    zero = np.zeros(1, np.int32)
    bound = N * N * 2.0 ** -43
    for a in acts:
        f, r, c = cfg.decode(int(a))
        s = 1 - 2 * int(st[f, r, c])
        d_sii, d_sit, _ = O.delta_terms(cfg, U[f], I, t64[0], 0, r, c, s)
        res = eng.step_batch(np.array([a]), zero, RULE_NEVER)[0]
        # fp32 fields contribute a relative error; the fixed-point rounding the absolute one
        assert abs(res["d_sii"] - d_sii) <= 2e-6 * abs(d_sii) + bound, (res["d_sii"], d_sii)
        assert abs(res["d_sit"] - d_sit) <= 2e-6 * abs(d_sit) + bound, (res["d_sit"], d_sit)
    eng.close()


@pytest.mark.gpu
def test_dbs_windows_keep_candidates_of_untouched_colour_groups_exactly():
    """A speculation window goes on after its first kept flip through candidates of other colour groups
    (DBS_1024_24.py:324-352: only the flipped group changes).  Decisions AND the PSNR trace must be those of
    the one-candidate-at-a-time loop, bit for bit, for every depth."""
    N, F = 128, 12
    pre, tgt, st = _problem(N, F, O.WL_RGB, 11)
    order = np.random.default_rng(4).permutation(F * N * N)[:6000]
    runs = {}
    for k in (1, 3, 16, 128):
        eng = _engine(N, F, O.WL_RGB)
        eng.set_target(0, tgt)
        eng.load_state(0, st)
        acc, tr, nacc, fin = eng.dbs_run(order, k_spec=k, resync_every=0, trace=True)
        runs[k] = (acc.copy(), tr.copy(), nacc, fin, eng.state(0).copy())
        eng.close()
    a1, t1, n1, f1, s1 = runs[1]
    assert 0.2 < n1 / order.size < 0.8                  # a regime where windows do hold several kept flips
    for k in (3, 16, 128):
        a, t, n, f, s = runs[k]
        assert np.array_equal(a, a1), k
        assert np.array_equal(t, t1), k                 # identical float64 values, not just close
        assert n == n1 and f == f1 and np.array_equal(s, s1), k


def test_batched_dbs_equals_the_sequential_loop_of_every_image():
    """bh_dbs_run_batch: several images in flight, one candidate per image and launch; each image's decisions,
    PSNR trace and final hologram are those of its own sequential greedy loop (bh_dbs_run, no speculation)."""
    N, F, wl, E, n = 64, 6, O.WL_RGB, 3, 1500
    eng = _engine(N, F, wl, n_env=E)
    one = _engine(N, F, wl)
    rng = np.random.default_rng(4)
    orders = np.stack([rng.permutation(F * N * N)[:n] for _ in range(E)])
    probs = [_problem(N, F, wl, 800 + e) for e in range(E)]
    for e, (pre, tgt, st) in enumerate(probs):
        eng.set_target(e, tgt)
        eng.load_state(e, st)
    acc, tr, nacc, fin = eng.dbs_run_batch(orders, trace=True)
    for e, (pre, tgt, st) in enumerate(probs):
        one.set_target(0, tgt)
        one.load_state(0, st)
        a1, t1, n1, f1 = one.dbs_run(orders[e], k_spec=1, trace=True)
        assert np.array_equal(acc[e], a1) and n1 == nacc[e]
        np.testing.assert_array_equal(tr[e], t1)
        assert fin[e] == f1
        assert np.array_equal(eng.state(e), one.state(0))
    eng.close(); one.close()


# ---------------------------------------------------------------------------
# round 2, session 4: persistent open-loop rollout kernel (bh_rollout_device)
# ---------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.timeout(300)
@pytest.mark.parametrize("N,F,wl,E,steps,rule", [
    (32, 4, O.WL_MONO, 1, 40, RULE_ENV),          # one unit per image: one CTA per env
    (64, 6, O.WL_RGB, 3, 60, RULE_ENV),
    (256, 8, O.WL_MONO, 5, 80, RULE_DBS),
    (1024, 24, O.WL_RGB, 8, 48, RULE_ENV),        # the bench shape: 37 CTAs per env
    (1024, 3, O.WL_RGB, 1, 24, RULE_DBS),         # one env owns the whole chip (look-ahead variant of the kernel)
    (256, 12, O.WL_RGB, 2, 150, RULE_DBS),        # two envs: look-ahead, many same-group neighbours
    (128, 4, O.WL_MONO, 1, 200, RULE_ENV),        # one colour group: every kept flip invalidates the look-ahead
    (896, 3, O.WL_RGB, 2, 12, RULE_ENV),          # not row regular: falls back to the two-kernel chain
])
def test_rollout_kernel_is_bit_identical_to_the_step_chain(N, F, wl, E, steps, rule):
    """bh_rollout_device (one persistent cooperative launch, per-environment barriers) against `steps` calls of
    bh_step_batch_device on a twin context: result records, final holograms, fields, reconstructions and running
    sums are identical bit for bit.  The action lists revisit a pixel in consecutive steps (kept and rejected),
    contain idle slots (-1) and use both list layouts (step-major, env-major)."""
    import torch
    rng = np.random.default_rng(N + 7 * E)
    acts = rng.integers(0, F * N * N, size=(steps, E)).astype(np.int64)
    acts[3] = acts[2]                                   # same pixels again right away
    acts[4] = acts[2]                                   # and a third time
    acts[7, 0] = acts[5, 0]                             # same pixel two steps later
    if E > 1:
        acts[9, 1] = -1                                 # idle slot
    envs = np.arange(E, dtype=np.int32)[::-1].copy()    # slot e drives environment E-1-e
    engs = []
    for _ in range(2):
        eng = _engine(N, F, wl, n_env=E)
        for e in range(E):
            pre, tgt, st = _problem(N, F, wl, seed=40 + e)
            eng.set_target(e, tgt)
            eng.load_state(e, st)
        engs.append(eng)
    chain, roll = engs
    d_envs = torch.from_numpy(envs).cuda()
    d_acts = torch.from_numpy(acts).cuda()
    d_res_a = torch.zeros(steps * E * 40, dtype=torch.uint8, device="cuda")
    for t in range(steps):
        chain.step_batch_device(E, d_envs.data_ptr(), d_acts.data_ptr() + t * E * 8, rule, d_res_a.data_ptr() + t * E * 40)
    chain.stream_sync()
    regular = 1024 % N == 0
    if regular:
        # env-major lists: actions[e][t], results[e][t]
        d_acts_t = d_acts.t().contiguous()
        d_res_b = torch.zeros(steps * E * 40, dtype=torch.uint8, device="cuda")
        roll.rollout_device(E, d_envs.data_ptr(), d_acts_t.data_ptr(), steps, rule, d_res_b.data_ptr(),
                            act_strides=(1, steps), res_strides=(1, steps))
        roll.rollout_status()
        res_b = d_res_b.cpu().numpy().view(RESULT_DTYPE).reshape(E, steps).T
    else:
        d_res_b = torch.zeros(steps * E * 40, dtype=torch.uint8, device="cuda")
        roll.rollout_device(E, d_envs.data_ptr(), d_acts.data_ptr(), steps, rule, d_res_b.data_ptr())
        roll.rollout_status()
        res_b = d_res_b.cpu().numpy().view(RESULT_DTYPE).reshape(steps, E)
    res_a = d_res_a.cpu().numpy().view(RESULT_DTYPE).reshape(steps, E)
    for name in ("action", "accept", "sgn", "psnr_after", "d_sii", "d_sit"):
        assert np.array_equal(res_a[name], res_b[name]), name
    assert 0 < res_a["accept"].sum() < steps * E         # both decisions occur
    for e in range(E):
        assert np.array_equal(chain.state(e), roll.state(e))
        assert chain.metrics(e) == roll.metrics(e)
        assert np.array_equal(chain.recon(e), roll.recon(e))
        for f in range(0, F, max(1, F // 3)):
            assert np.array_equal(chain.field(e, f), roll.field(e, f))
    # a second rollout on the same context continues from the committed state (ring and sums re-initialised)
    if regular:
        more = rng.integers(0, F * N * N, size=(10, E)).astype(np.int64)
        d_more = torch.from_numpy(more).cuda()
        for t in range(10):
            chain.step_batch_device(E, d_envs.data_ptr(), d_more.data_ptr() + t * E * 8, rule, 0 if False else d_res_a.data_ptr())
        roll.rollout_device(E, d_envs.data_ptr(), d_more.data_ptr(), 10, rule, 0)
        roll.rollout_status()
        chain.stream_sync()
        for e in range(E):
            assert np.array_equal(chain.state(e), roll.state(e))
            assert chain.metrics(e) == roll.metrics(e)
    chain.close(); roll.close()
