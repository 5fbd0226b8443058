"""CPU tests of the Python host logic (envs.py, vec_env.py, dbs.py) with the oracle-backed engine
stand-in of tests/oracle_engine.py in place of the CUDA engine: rewards, bonuses, termination,
rejected-step quirks, auto-reset, the rank-table reward, group clones, the DBS drivers and their
checkpoints, all against oracle.hologram_oracle.  The GPU suite runs the
same comparisons through the C ABI; here the scoring itself is the oracle's, so any difference is
a host-logic bug."""
import os

import numpy as np
import pytest

import binary_hologram_reinforcement_learning_b200 as bh
from binary_hologram_reinforcement_learning_b200 import dbs, envs, vec_env
from oracle import hologram_oracle as O
from tests.oracle_engine import OracleEngine, pinned_stub


@pytest.fixture(autouse=True)
def _oracle_engine(monkeypatch):
    for mod in (envs, vec_env, dbs):
        if hasattr(mod, "HoloEngine"):
            monkeypatch.setattr(mod, "HoloEngine", OracleEngine)
        if hasattr(mod, "pinned_empty"):
            monkeypatch.setattr(mod, "pinned_empty", pinned_stub)


def _loader_fn(loaders):
    def tf(t):
        key = np.ascontiguousarray(t[0, 0, 0, :4]).tobytes()
        return next(l for l in loaders if key in l._pre).target_function(t)
    return tf


@pytest.mark.parametrize("N,F,wl,kw", [
    (16, 4, O.WL_MONO, dict(max_steps=30, T_PSNR_DIFF=1e9)),          # ends on max_steps (bonus -595.24)
    (16, 6, O.WL_RGB, dict(max_steps=10 ** 6, T_PSNR_DIFF=0.02)),     # ends on the PSNR goal (bonus -595.2)
    (16, 4, O.WL_MONO, dict(max_steps=25, T_PSNR_DIFF=0.01, T_steps=3)),
])
def test_single_env_matches_oracle_env(N, F, wl, kw):
    G = len(wl)
    ld = bh.SyntheticLoader(N, F, G, seeds=(11,))
    env = bh.BinaryHologramEnv(ld.target_function, ld, IPS=N, CH=F, wl=wl, verbose=False, recon_obs="eager", **kw)
    obs, info = env.reset()
    pre, tgt = bh.synthetic_problem(N, F, G, 11)
    ref = O.OracleEnv(O.HoloConfig(N=N, F=F, wl=wl), **kw)
    ref.reset(pre, tgt)
    assert env.initial_psnr == ref.initial_psnr and info["state"] is env.state
    assert obs["state"].shape == (1, F, N, N) and obs["recon_image"].shape == (1, G, N, N)
    rng = np.random.default_rng(3)
    ended = False
    for step in range(400):
        a = int(rng.integers(0, F * N * N))
        obs, r, term, trunc, info = env.step(a)
        r_ref, term_ref, trunc_ref, p, acc = ref.step(a)
        assert r == pytest.approx(r_ref, rel=1e-12, abs=1e-12)
        assert (term, trunc) == (term_ref, trunc_ref)
        assert np.array_equal(obs["state"][0], ref.state) and np.array_equal(obs["state_record"][0], ref.state_record)
        # appendix B-2: the observation shows the evaluated flip even when it was rolled back
        np.testing.assert_allclose(obs["recon_image"][0], ref.recon, rtol=0, atol=1e-6)
        assert env.flip_count == ref.flip_count and env.steps == ref.steps
        assert env.previous_psnr == ref.previous_psnr
        if term or trunc:
            ended = True
            break
    assert ended
    env.close()


@pytest.mark.parametrize("verbose", [False, True])
def test_vec_env_autoreset_and_bonus_match_oracle_envs(verbose, capsys):
    """Fast path (bh_vec_book_update in C + numpy) and per-env path give the oracle's rewards and dones."""
    N, F, wl, E, MAXS = 16, 4, O.WL_MONO, 3, 9
    loaders = [bh.SyntheticLoader(N, F, 1, seeds=(200 + i,)) for i in range(E)]
    vec = bh.HologramVecEnv(E, _loader_fn(loaders), loaders, max_steps=MAXS, T_PSNR_DIFF=1e9, IPS=N, CH=F,
                            wl=wl, verbose=verbose)
    assert vec._fast == (not verbose)
    vec.reset()
    probs = [bh.synthetic_problem(N, F, 1, 200 + i) for i in range(E)]
    refs = []
    for i in range(E):
        e = O.OracleEnv(O.HoloConfig(N=N, F=F, wl=wl), max_steps=MAXS, T_PSNR_DIFF=1e9)
        e.reset(*probs[i])
        refs.append(e)
    rng = np.random.default_rng(8)
    n_done = 0
    for step in range(60):
        acts = rng.integers(0, F * N * N, size=E)
        obs, rewards, dones, infos = vec.step(acts)
        for i in range(E):
            r, term, trunc, p, acc = refs[i].step(int(acts[i]))
            assert rewards[i] == pytest.approx(r, rel=1e-12, abs=1e-12), (step, i)
            assert bool(dones[i]) == bool(term or trunc)
            if dones[i]:
                n_done += 1
                assert infos[i]["TimeLimit.truncated"] is False
                assert np.array_equal(infos[i]["terminal_observation"]["state"][0], refs[i].state)
                refs[i].reset(*probs[i])
            assert np.array_equal(obs[i]["state"][0], refs[i].state)
            assert np.array_equal(obs[i]["state_record"][0], refs[i].state_record)
    assert n_done >= E and len(vec.episode_stats) == n_done
    for row in vec.episode_stats:                      # [reward, steps, flips, psnr0, psnr1]
        assert row[1] >= MAXS and 0 <= row[2] <= row[1] and row[4] >= row[3]
    capsys.readouterr()
    vec.close()


def test_vec_env_goal_termination_fast_path():
    """env.py:216-235 through the event mask of bh_vec_book_update: success bonus and sustained steps."""
    N, F, E = 16, 4, 2
    kw = dict(max_steps=10 ** 6, T_PSNR_DIFF=0.015, T_steps=2)
    loaders = [bh.SyntheticLoader(N, F, 1, seeds=(300 + i,)) for i in range(E)]
    vec = bh.HologramVecEnv(E, _loader_fn(loaders), loaders, IPS=N, CH=F, **kw)
    assert vec._fast
    vec.reset()
    probs = [bh.synthetic_problem(N, F, 1, 300 + i) for i in range(E)]
    refs = [O.OracleEnv(O.HoloConfig(N=N, F=F), **kw) for _ in range(E)]
    for e, p in zip(refs, probs):
        e.reset(*p)
    rng = np.random.default_rng(4)
    n_done = 0
    for step in range(300):
        acts = rng.integers(0, F * N * N, size=E)
        _, rewards, dones, infos = vec.step(acts)
        for i in range(E):
            r, term, trunc, p, acc = refs[i].step(int(acts[i]))
            assert rewards[i] == pytest.approx(r, rel=1e-12, abs=1e-12), (step, i)
            assert bool(dones[i]) == bool(term or trunc)
            if dones[i]:
                n_done += 1
                refs[i].reset(*probs[i])
        if n_done >= 3:
            break
    assert n_done >= 3
    vec.close()


def test_group_env_and_vec_group_rewards_match_oracle():
    """env_group.py:90-143,254-255: rank table at reset, nearest-value reward; vec fast path + clones."""
    N, F, S = 16, 4, 60
    ld = bh.SyntheticLoader(N, F, 1, seeds=(31,))
    env = bh.BinaryHologramEnvGroup(ld.target_function, ld, IPS=N, CH=F, verbose=False, num_samples=S,
                                    rng=np.random.default_rng(9))
    env.reset()
    pre, tgt = bh.synthetic_problem(N, F, 1, 31)
    ref = O.OracleEnv(O.HoloConfig(N=N, F=F), reward_mode="group")
    ref.reset(pre, tgt, rng=np.random.default_rng(9), num_samples=S)
    np.testing.assert_allclose(env.psnr_change_list, ref.psnr_change_list, rtol=0, atol=1e-12)
    np.testing.assert_allclose(env.importance_ranks, ref.importance_ranks, rtol=0, atol=1e-12)
    assert env.T_PSNR_DIFF == pytest.approx(ref.T_PSNR_DIFF, abs=1e-12)
    rng = np.random.default_rng(1)
    for _ in range(40):
        a = int(rng.integers(0, F * N * N))
        _, r, term, trunc, _ = env.step(a)
        r_ref, term_ref, trunc_ref, _, _ = ref.step(a)
        assert r == pytest.approx(r_ref, abs=1e-12) and (term, trunc) == (term_ref, trunc_ref)
        if term or trunc:
            break
    env.close()
    # vectorised: the O(log n) nearest-rank lookup equals np.argmin; members of a group are clones
    E = 4
    loaders = [bh.SyntheticLoader(N, F, 1, seeds=(40 + i,)) for i in range(E)]
    vec = bh.HologramVecEnv(E, _loader_fn(loaders), loaders, max_steps=10 ** 6, T_PSNR_DIFF=1e9, IPS=N, CH=F,
                            reward_mode="group", num_samples=S, seed=5)
    vec.reset_groups(2)
    assert vec._fast
    for lead in (0, 2):
        assert np.array_equal(vec.envs[lead + 1].state, vec.envs[lead].state)
        assert vec.envs[lead + 1].importance_ranks is vec.envs[lead].importance_ranks
        assert np.array_equal(vec.engine.state(lead + 1), vec.engine.state(lead))
    for step in range(25):
        acts = rng.integers(0, F * N * N, size=E)
        prev = vec._prev.copy()
        _, rewards, dones, _ = vec.step(acts)
        for i in range(E):
            change = vec._res["psnr_after"][i] - prev[i]
            idx = int(np.argmin(np.abs(vec.envs[i]._psnr_change_arr - change)))
            if not dones[i]:
                assert rewards[i] == vec.envs[i].importance_ranks[idx]
    vec.close()


def test_cropped_env_maps_actions_and_ignores_pixels_outside_the_window():
    """env_1024_24_128.py: the centre window is simulated; a flip outside it changes nothing but the mirrors."""
    N, F, m = 24, 4, 4
    ld = bh.SyntheticLoader(N, F, 1, seeds=(5,))
    env = bh.BinaryHologramEnv(ld.target_function, ld, IPS=N, CH=F, crop_margin=m, verbose=False,
                               max_steps=10 ** 6, T_PSNR_DIFF=1e9, recon_obs="eager")
    obs, _ = env.reset()
    assert env.engine.N == N - 2 * m and obs["recon_image"].shape == (1, 1, N - 2 * m, N - 2 * m)
    pre, tgt = bh.synthetic_problem(N, F, 1, 5)
    ref = O.OracleEnv(O.HoloConfig(N=N - 2 * m, F=F), max_steps=10 ** 6, T_PSNR_DIFF=1e9)
    ref.reset(pre[:, m:-m, m:-m], tgt[:, m:-m, m:-m])
    assert env.initial_psnr == ref.initial_psnr
    p0 = env.previous_psnr
    _, r, term, trunc, _ = env.step(0)                                 # corner pixel: outside
    assert r == 0.0 and env.previous_psnr == p0 and env.state[0, 0, 0, 0] == 1 - int(pre[0, 0, 0] >= 0.5)
    a_full = (1 * N + (m + 3)) * N + (m + 5)                           # frame 1, window pixel (3, 5)
    _, r, _, _, _ = env.step(a_full)
    r_ref, _, _, _, acc = ref.step((1 * (N - 2 * m) + 3) * (N - 2 * m) + 5)
    assert r == pytest.approx(r_ref, abs=1e-12)
    env.close()


def test_dbs_greedy_driver_checkpoint_and_resume(tmp_path):
    """DBS.py:202-305 driver: same decisions as the oracle loop; a run cut into segments resumes bit-exactly."""
    N, F = 16, 4
    ld = bh.SyntheticLoader(N, F, 1, seeds=(21,))
    pre, tgt = bh.synthetic_problem(N, F, 1, 21)
    cfg = O.HoloConfig(N=N, F=F)
    order = np.random.default_rng(6).permutation(F * N * N)
    st_ref, acc_ref, tr_ref = O.dbs_greedy(cfg, (pre >= 0.5).astype(np.int8), tgt, order)

    def run(**kw):
        env = bh.BinaryHologramEnv(ld.target_function, ld, IPS=N, CH=F, verbose=False)
        out = bh.optimize_with_random_pixel_flips(env, 2e-3, 7.56e-6, max_datasets=0, rng=np.random.default_rng(6),
                                                  verbose=False, **kw)
        env.close()
        return out

    full = run()[0]
    assert np.array_equal(full["order"], order)
    assert np.array_equal(full["accepted"].astype(bool), acc_ref)
    assert np.array_equal(full["state"].reshape(F, N, N), st_ref)
    assert full["final_psnr"] == pytest.approx(tr_ref[acc_ref][-1], abs=1e-12)
    ck = str(tmp_path / "dbs.npz")
    part = run(checkpoint=ck, segment=300, max_segments=2)[0]
    assert os.path.exists(ck) and not part.get("complete", True)
    rest = run(checkpoint=ck, segment=300)[0]
    assert np.array_equal(rest["accepted"], full["accepted"]) and np.array_equal(rest["state"], full["state"])
    assert rest["final_psnr"] == full["final_psnr"]
