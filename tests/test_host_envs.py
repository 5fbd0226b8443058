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


def test_group_env_and_vec_group_rewards_match_oracle(capsys):
    """env_group.py:90-143,254-255: rank table at reset, nearest-value reward; vec fast path + clones."""
    N, F, S = 16, 4, 60
    ld = bh.SyntheticLoader(N, F, 1, seeds=(31,))
    env = bh.BinaryHologramEnvGroup(ld.target_function, ld, IPS=N, CH=F, verbose=True, num_samples=S,
                                    rng=np.random.default_rng(9))
    env.reset()
    head = capsys.readouterr().out                         # env_group.py:128-129,194-199 / log_py/'Dynamic Threshold.py':16-20
    assert "Polynomial Reward Function Equation:" in head and "Time taken for psnr_change_list:" in head
    import re
    assert float(re.search(r"\[Dynamic Threshold\] T_PSNR_DIFF set to: ([0-9.]+)", head).group(1)) == pytest.approx(
        env.T_PSNR_DIFF, abs=1e-6)
    assert re.search(r"\[Episode Start\].*dataset file: \('(?:.*/)?(.*?)',\)", head).group(1) == "synthetic_0031.png"
    env.verbose = False
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
    # decile statistics of the kept flips (DBS_1024_24.py:398-416)
    assert full["improved_bin_counts"].sum() == full["flip_count"] == acc_ref.sum()
    assert full["psnr_improvements"].sum() == pytest.approx(full["final_psnr"] - full["initial_psnr"], abs=1e-9)
    d = dbs.decile_index(pre.ravel()[order[acc_ref]])
    assert np.array_equal(full["improved_bin_counts"], np.bincount(d, minlength=10))
    ck = str(tmp_path / "dbs.npz")
    part = run(checkpoint=ck, segment=300, max_segments=2)[0]
    assert os.path.exists(ck) and not part.get("complete", True)
    rest = run(checkpoint=ck, segment=300)[0]
    assert np.array_equal(rest["accepted"], full["accepted"]) and np.array_equal(rest["state"], full["state"])
    assert rest["final_psnr"] == full["final_psnr"]


def test_dbs_resume_adopts_stored_order_and_skips_finished_images(tmp_path):
    """A restarted process draws a NEW permutation and starts at the loader's first image again: the checkpoint
    carries the image number and the order of the image in progress, so the run continues where it stopped."""
    N, F = 16, 4

    def run(rng, **kw):
        ld = bh.SyntheticLoader(N, F, 1, seeds=(21, 22))
        env = bh.BinaryHologramEnv(ld.target_function, ld, IPS=N, CH=F, verbose=False)
        out = bh.optimize_with_random_pixel_flips(env, 2e-3, 7.56e-6, max_datasets=1, rng=rng, verbose=False, **kw)
        env.close()
        return out

    ref = run(np.random.default_rng(6))
    assert len(ref) == 2 and all(r["complete"] for r in ref)
    ck = str(tmp_path / "dbs2.npz")
    # 1024 candidates per image, 300 per segment: image 1 takes 4 segments, the call stops inside image 2
    part = run(np.random.default_rng(6), checkpoint=ck, segment=300, max_segments=6)
    assert len(part) == 2 and part[0]["complete"] and not part[1]["complete"] and part[1]["steps"] == 600
    # the "restarted process": an unrelated rng stream, a fresh loader and env
    rest = run(np.random.default_rng(12345), checkpoint=ck, segment=300)
    assert len(rest) == 1 and rest[0]["complete"] and rest[0]["file"] == ref[1]["file"]
    assert np.array_equal(rest[0]["order"], ref[1]["order"])
    assert np.array_equal(rest[0]["accepted"], ref[1]["accepted"]) and np.array_equal(rest[0]["state"], ref[1]["state"])
    assert rest[0]["final_psnr"] == ref[1]["final_psnr"]
    # everything is finished now: another restart has nothing left to do
    assert run(np.random.default_rng(1), checkpoint=ck, segment=300) == []


def test_dbs_sweep_driver_full_sharded_and_partial():
    """dbs-1024-1024-24-6464.py:194-478: crop, score every flip against the fixed state, decile statistics;
    the whole-image path, the rank-sharded path and a candidate subset agree with the oracle."""
    N, F, m = 16, 4, 2
    n = F * (N - 2 * m) ** 2
    ld = bh.SyntheticLoader(N, F, 1, seeds=(61,))
    pre, tgt = bh.synthetic_problem(N, F, 1, 61)
    cfg = O.HoloConfig(N=N - 2 * m, F=F)
    cst, ctg, cpre = (pre >= 0.5).astype(np.int8)[:, m:-m, m:-m], tgt[:, m:-m, m:-m], pre[:, m:-m, m:-m]

    def run(**kw):
        return bh.optimize_with_random_pixel_flips(ld.target_function, ld, 2e-3, 7.56e-6, m, CH=F, wl=O.WL_MONO,
                                                   max_datasets=0, rng=np.random.default_rng(1), verbose=False, **kw)[0]

    full = run()
    ref_psnr, p0, att, imp, gain = O.sweep(cfg, cst, ctg, cpre, full["order"])
    assert full["steps"] == n and full["initial_psnr"] == p0
    np.testing.assert_allclose(full["psnr_after"], ref_psnr, rtol=0, atol=1e-12)
    assert np.array_equal(full["attempted"], att) and np.array_equal(full["improved"], imp)
    np.testing.assert_allclose(full["gains"], gain, rtol=1e-12, atol=1e-15)
    assert np.array_equal(full["attempted"], full["bin_counts"]) and full["flip_count"] == imp.sum()
    parts = [run(shard=(r, 3)) for r in range(3)]
    assert np.array_equal(np.concatenate([p["order"] for p in parts]), full["order"])
    np.testing.assert_allclose(np.concatenate([p["psnr_after"] for p in parts]), ref_psnr, rtol=0, atol=1e-12)
    assert np.array_equal(sum(p["attempted"] for p in parts), att)
    assert np.array_equal(sum(p["improved"] for p in parts), imp)
    np.testing.assert_allclose(sum(p["gains"] for p in parts), gain, rtol=1e-12, atol=1e-15)
    some = run(max_candidates=40)                              # a candidate subset goes through eval_flips
    assert some["steps"] == 40 and np.array_equal(some["order"], full["order"][:40])
    np.testing.assert_allclose(some["psnr_after"], ref_psnr[:40], rtol=0, atol=1e-12)
    r40 = O.sweep(cfg, cst, ctg, cpre, full["order"][:40])
    assert np.array_equal(some["attempted"], r40[2]) and np.array_equal(some["improved"], r40[3])


# the reference's own log parsers, restated: log_py/DBS_psnr_log.py:23-30 and log_py/valid_log.py:24-31
DBS_STEP_RE = (r"Step: (\d+)\s+"
               r"PSNR Before: [\d.]+\s+\|\s+PSNR After: [\d.]+\s+\|\s+Change: [\d.e+-]+\s+\|\s+Diff: ([\d.e+-]+)\s+"
               r"Success Ratio: ([\d.e+-]+)\s+\|\s+Flip Count: (\d+).*?"
               r"Time taken for this data: ([\d.]+) seconds")
ENV_STEP_RE = (r"Step: (\d+)\s+\| Initial PSNR: ([\d.]+)\s+"
               r"PSNR After: ([\d.]+)\s+\|\s+Change: ([\d.e+-]+)\s+\|\s+Diff: ([\d.e+-]+)\s+"
               r"Reward: ([\d.]+)\s+\|\s+Success Ratio: ([\d.e+-]+)\s+\|\s+Flip Count: (\d+).*?"
               r"Time taken for this data: ([\d.]+) seconds")


def test_reference_log_parsers_read_our_stdout(capsys):
    """the stdout blocks are the reference's wire format.  The regexes of its log_py/ parsers
    find the step blocks of the greedy DBS (DBS.py:283-289,299-305) and of the env (env.py:206-212)."""
    import re
    N, F = 16, 4
    ld = bh.SyntheticLoader(N, F, 1, seeds=(7,))
    env = bh.BinaryHologramEnv(ld.target_function, ld, IPS=N, CH=F, verbose=True, max_steps=10 ** 6, T_PSNR_DIFF=1e9)
    res = bh.optimize_with_random_pixel_flips(env, 2e-3, 7.56e-6, max_datasets=0, rng=np.random.default_rng(2),
                                              verbose=True)[0]
    out = capsys.readouterr().out
    assert re.search(r"Starting pixel flip optimization for file \S+\.png with initial PSNR: \d+\.\d{6}", out)
    blocks = list(re.finditer(DBS_STEP_RE, out, re.DOTALL))
    assert len(blocks) >= 2                                    # threshold blocks (+0.5 dB each) and the summary
    steps = [int(m.group(1)) for m in blocks]
    assert steps == sorted(steps) and steps[-1] == res["steps"]
    for m in blocks[:-1]:                                      # threshold block: counters at that accepted flip
        j = int(m.group(1))
        assert bool(res["accepted"][j - 1]) and int(m.group(4)) == int(res["accepted"][:j].sum())
        assert float(m.group(2)) == pytest.approx(res["psnr_trace"][j - 1] - res["initial_psnr"], abs=1e-6)
    assert int(blocks[-1].group(4)) == res["flip_count"]
    assert re.search(r"Flip Pixel: Channel=\d+, Row=\d+, Col=\d+\nTime taken for this data: ", out)
    fin = re.search(r"Optimization completed\. Final PSNR improvement: (-?\d+\.\d{6})", out)
    assert float(fin.group(1)) == pytest.approx(res["psnr_trace"][-1] - res["initial_psnr"], abs=1e-6)   # B-11

    env.reset()
    head = capsys.readouterr().out
    # log_py/valid_log.py:12 / DBS_psnr_log.py:12 split episodes on this line; comp.py:25 takes the file name
    assert re.search(r"\[Episode Start\] Currently using dataset file: \((.+?)\), Episode count: \d+", head)
    assert re.search(r"Currently using dataset file:\s*\('.*?([^/]+\.png)',\)", head).group(1) == "synthetic_0007.png"
    rng = np.random.default_rng(0)
    for _ in range(300):
        _, _, term, trunc, _ = env.step(int(rng.integers(0, F * N * N)))
    out = capsys.readouterr().out
    blocks = list(re.finditer(ENV_STEP_RE, out, re.DOTALL))
    assert blocks, out[:500]
    for m in blocks:
        assert float(m.group(2)) == pytest.approx(env.initial_psnr, abs=1e-6)
        assert float(m.group(5)) == pytest.approx(float(m.group(3)) - env.initial_psnr, abs=2e-6)
    env.close()


# log_py/'dbs 평균.py':25-33 -- the range lines of the sweep's progress blocks
RANGE_RE = (r"Range\s*([\d\.]+-[\d\.]+):\s*Total Pixels\s*=\s*(\d+),\s*"
            r"Improved Pixels\s*=\s*(\d+),\s*Attempted Pixels\s*=\s*(\d+),\s*"
            r"Improvement Ratio\s*=\s*([\d\.]+),\s*"
            r"Improvement Ratio \(in range\)\s*=\s*([\d\.]+),\s*"
            r"Improvement Ratio \(to total improved\)\s*=\s*([\d\.]+),\s*"
            r"Total PSNR Improvement\s*=\s*([\d\.]+),\s*"
            r"Average PSNR Improvement\s*=\s*([\d\.]+)")


def test_sweep_stdout_replays_the_reference_blocks(capsys):
    """dbs-1024-1024-24-6464.py:396-447,462-478: a step block + cumulative range lines every `log_every`
    candidates (counters include that candidate), then the summary; parsed with the reference's regex."""
    import re
    N, F, m, every = 16, 4, 2, 100
    ld = bh.SyntheticLoader(N, F, 1, seeds=(61,))
    res = bh.optimize_with_random_pixel_flips(ld.target_function, ld, 2e-3, 7.56e-6, m, CH=F, wl=O.WL_MONO,
                                              max_datasets=0, rng=np.random.default_rng(1), verbose=True,
                                              log_every=every)[0]
    out = capsys.readouterr().out
    n = res["steps"]
    better = res["psnr_after"] > res["initial_psnr"]
    blocks = list(re.finditer(r"Step: (\d+)\nPSNR Before: ([\d.]+) \| PSNR After: ([\d.]+) \| Change: (-?[\d.]+)\n"
                              r"Success Ratio: ([\d.]+) \| Flip Count: (\d+)\n"
                              r"Flip Pixel: Channel=(\d+), Row=(\d+), Col=(\d+)\n"
                              r"Time taken for this data: [\d.]+ seconds(\npre_value: ([\d.]+))?", out))
    assert [int(b.group(1)) for b in blocks] == list(range(every, n + 1, every)) + [n]
    side = N - 2 * m
    for b in blocks:
        k = int(b.group(1))
        assert int(b.group(6)) == int(better[:k].sum())                    # counters include candidate k
        assert float(b.group(3)) == pytest.approx(res["psnr_after"][k - 1], abs=1e-6)
        a = int(res["order"][k - 1])
        assert (int(b.group(7)), int(b.group(8)), int(b.group(9))) == (a // side ** 2, (a % side ** 2) // side, a % side)
    assert all(b.group(10) for b in blocks[:-1]) and blocks[-1].group(10) is None   # pre_value only in progress blocks
    lines = [re.search(RANGE_RE, l) for l in out.splitlines()]
    lines = [l for l in lines if l]
    assert len(lines) == 10 * (n // every)
    last10 = lines[-10:]
    k = (n // every) * every
    assert sum(int(l.group(4)) for l in last10) == k and sum(int(l.group(3)) for l in last10) == int(better[:k].sum())
    assert [int(l.group(2)) for l in last10] == list(res["bin_counts"])
    assert re.search(r"\S+\.png Optimization completed\. Final PSNR improvement: -?\d+\.\d{6}\n"
                     r"Time taken for this data: [\d.]+ seconds\n\nPre-model output range statistics:\nRange 0\.0-0\.1: "
                     r"Total Pixels = \d+, Improved Pixels = \d+, Improvement Ratio \(in range\)", out)


def test_rgb_greedy_stdout_has_the_range_statistics(capsys, tmp_path):
    """DBS_1024_24.py:372-396,430-469: threshold blocks every 0.1 dB followed by the cumulative range lines
    (flips kept BEFORE that step), summary without 'Diff', final 'Pre-model output range statistics'."""
    import re
    N, F = 16, 6
    ld = bh.SyntheticLoader(N, F, 3, seeds=(9,))
    env = bh.BinaryHologramEnv(ld.target_function, ld, IPS=N, CH=F, wl=O.WL_RGB, verbose=False)
    res = bh.optimize_with_random_pixel_flips(env, 2e-3, 7.56e-6, max_datasets=0, rng=np.random.default_rng(3),
                                              verbose=True, save_dir=str(tmp_path))[0]
    out = capsys.readouterr().out
    # DBS_1024_24.py:282-287,441-451: before / after reconstructions; log_py/'log dbs.py':16 splits on this line
    assert out.count("RGB data saved to ") == 2
    before = np.load(tmp_path / "episode_synthetic_0009png_rgb_before.npy")
    after = np.load(tmp_path / "episode_synthetic_0009_rgb_after.npy")
    assert before.shape == after.shape == (1, 3, N, N) and not np.array_equal(before, after)
    assert np.array_equal(np.load(tmp_path / "episode_synthetic_0009_state_after.npy"), res["state"])
    assert out.index("Pre-model output range statistics:") < out.rindex("RGB data saved to ") < out.rindex("Range 0.0-0.1")
    blocks = list(re.finditer(DBS_STEP_RE, out, re.DOTALL))
    assert blocks and all(res["accepted"][int(b.group(1)) - 1] for b in blocks)
    first = int(blocks[0].group(1))
    rng_lines = re.findall(r"Range \d\.\d-\d\.\d: Total Pixels = (\d+), Improved Pixels = (\d+), Improvement Ratio \(in range\)", out)
    assert len(rng_lines) == 10 * (len(blocks) + 1)            # one table per threshold block + the final one
    assert sum(int(i) for _, i in rng_lines[:10]) == int(res["accepted"][:first - 1].sum())
    assert sum(int(i) for _, i in rng_lines[-10:]) == res["flip_count"]
    assert re.search(r"PSNR Before: [\d.]+ \| PSNR After: [\d.]+ \| Change: -?[\d.]+\nSuccess Ratio", out)   # :432
    assert "Pre-model output range statistics:" in out
    env.close()


def test_gpu_suite_driver_tests_also_hold_with_the_oracle_engine(golden_dir, tmp_path):
    """The DBS-driver tests of the GPU suite, executed here with the oracle engine: the driver code they
    exercise (checkpoints, mirrors, decile statistics, crop) is host logic and must not depend on the GPU."""
    from tests import test_gpu_parity as G
    G.test_dbs_greedy_driver_and_resume(golden_dir, tmp_path)
    G.test_multidiscrete_action_equals_flat_action()


def test_dbs_psnr_diff_threshold_stops_at_the_first_candidate_that_reaches_it(capsys):
    """DBS_01.py:204,320-325 / DBS_ratio_0.5.py:204: leave the image once a candidate lifts the PSNR by the
    threshold; state and decisions are those of the sequential loop cut at that candidate."""
    N, F, thr = 16, 4, 0.3
    ld = bh.SyntheticLoader(N, F, 1, seeds=(21,))
    pre, tgt = bh.synthetic_problem(N, F, 1, 21)
    cfg = O.HoloConfig(N=N, F=F)
    order = np.random.default_rng(6).permutation(F * N * N)
    st0 = (pre >= 0.5).astype(np.int8)
    _, acc_ref, tr_ref = O.dbs_greedy(cfg, st0, tgt, order)
    p0 = O.score(cfg, O.reconstruct(cfg, st0), tgt)[0]
    stop = int(np.flatnonzero(tr_ref - p0 >= thr)[0]) + 1
    assert 1 < stop < order.size
    st_ref, _, _ = O.dbs_greedy(cfg, st0, tgt, order[:stop])
    env = bh.BinaryHologramEnv(ld.target_function, ld, IPS=N, CH=F, verbose=False)
    res = bh.optimize_with_random_pixel_flips(env, 2e-3, 7.56e-6, max_datasets=0, rng=np.random.default_rng(6),
                                              verbose=True, psnr_diff_threshold=thr)[0]
    out = capsys.readouterr().out
    assert res["stopped_on_threshold"] and not res["complete"] and res["steps"] == stop
    assert np.array_equal(res["accepted"].astype(bool), acc_ref[:stop]) and res["accepted"][-1]
    assert np.array_equal(res["state"].reshape(F, N, N), st_ref) and np.array_equal(env.state[0], st_ref)
    assert res["final_psnr"] - res["initial_psnr"] >= thr
    assert f"PSNR diff threshold {thr} reached at step {stop}. Moving to next dataset." in out
    env.close()


def test_crop_margin_as_a_reset_argument_and_as_dbs_argument():
    """env_1024_24_128.py:100 takes crop_margin at reset, DBS_1024_24-128.py:187 as 4th argument of the DBS."""
    N, F, m = 24, 4, 4
    ld = bh.SyntheticLoader(N, F, 1, seeds=(5,))
    env = bh.BinaryHologramEnv(ld.target_function, ld, IPS=N, CH=F, verbose=False, T_PSNR_DIFF=1e9)
    obs, _ = env.reset()
    assert env.engine.N == N and obs["recon_image"].shape == (1, 1, N, N)
    first = env.engine
    obs, _ = env.reset(crop_margin=m)                          # new FFT side -> new engine
    assert first.closed and env.engine.N == N - 2 * m and obs["recon_image"].shape == (1, 1, N - 2 * m, N - 2 * m)
    assert env.observation_space["recon_image"].shape == (1, 1, N - 2 * m, N - 2 * m)
    pre, tgt = bh.synthetic_problem(N, F, 1, 5)
    cfg = O.HoloConfig(N=N - 2 * m, F=F)
    cst, ctg = (pre >= 0.5).astype(np.int8)[:, m:-m, m:-m], tgt[:, m:-m, m:-m]
    assert env.initial_psnr == O.score(cfg, O.reconstruct(cfg, cst), ctg)[0]
    env.close()
    env = bh.BinaryHologramEnv(ld.target_function, ld, IPS=N, CH=F, verbose=False)
    res = bh.optimize_with_random_pixel_flips(env, 2e-3, 7.56e-6, m, max_datasets=0, rng=np.random.default_rng(2),
                                              max_candidates=150, verbose=False)[0]
    st_ref, acc_ref, _ = O.dbs_greedy(cfg, cst, ctg, res["order"])
    assert res["order"].max() < F * (N - 2 * m) ** 2 and np.array_equal(res["accepted"].astype(bool), acc_ref)
    assert np.array_equal(env.state[0][:, m:-m, m:-m], st_ref)          # the window of the full-size mirror
    assert np.array_equal(env.state[0][:, :m], (pre >= 0.5).astype(np.int8)[:, :m])    # margin untouched
    shared = OracleEngine(N, F, O.WL_MONO, n_env=2)
    env2 = bh.BinaryHologramEnv(ld.target_function, ld, IPS=N, CH=F, verbose=False, engine=shared, env_index=1)
    with pytest.raises(ValueError):
        env2.reset(crop_margin=m)
    env.close()


@pytest.mark.parametrize("R", [1, 3, 7])
def test_vec_env_repropagates_exactly_when_a_flip_count_reaches_a_multiple(R):
    """resync_every: the incremental fields are re-propagated every R kept flips of an env.  The fast path
    only looks at the counters when a multiple can have been reached; the calls must be exactly the due ones."""
    N, F, E = 16, 4, 3
    loaders = [bh.SyntheticLoader(N, F, 1, seeds=(500 + i,)) for i in range(E)]
    vec = bh.HologramVecEnv(E, _loader_fn(loaders), loaders, max_steps=25, T_PSNR_DIFF=1e9, IPS=N, CH=F,
                            resync_every=R)
    assert vec._fast
    vec.reset()
    calls, expected = [], []
    real_resync = vec.engine.resync
    step_no = [0]

    def spy(env):
        calls.append((step_no[0], int(env)))
        real_resync(env)
    vec.engine.resync = spy
    rng = np.random.default_rng(12)
    for step in range(90):                                  # crosses several auto-resets (max_steps = 25)
        step_no[0] = step
        flips_before = vec._flips.copy()
        vec.step(rng.integers(0, F * N * N, size=E))
        acc = vec._res["accept"] != 0
        for i in range(E):
            if acc[i] and (flips_before[i] + 1) % R == 0:
                expected.append((step, i))
    # auto-reset calls env.reset -> load_state, not resync, so the spy only sees the periodic ones
    assert calls == expected and len(calls) > 0
    vec.close()


@pytest.mark.parametrize("name", ["mono64", "rgb64", "mono32_pad2", "mono64_abs"])
def test_golden_env_trajectories_through_the_env_with_the_oracle_engine(name, golden_dir):
    """The golden env trajectories (tests/golden/*.npz) replayed through BinaryHologramEnv.step on the CPU:
    same test body as the GPU suite, scoring by the oracle engine -- pins the env's host logic to the fixtures."""
    from tests import test_gpu_parity as G
    G.test_golden_env_trajectory(name, golden_dir)


def test_dataset_iterator_restart_and_image_count_quirks(capsys):
    """Appendix B-8: the env restarts its loader silently (one INFO line) at exhaustion, env.py:96-102.
    B-9: `while db_num <= max_datasets` processes max_datasets + 1 images, DBS.py:208."""
    N, F = 16, 4
    ld = bh.SyntheticLoader(N, F, 1, seeds=(1, 2))
    env = bh.BinaryHologramEnv(ld.target_function, ld, IPS=N, CH=F, verbose=True)
    names = []
    for _ in range(3):
        env.reset()
        names.append(env.current_file[0])
    out = capsys.readouterr().out
    assert names == ["synthetic_0001.png", "synthetic_0002.png", "synthetic_0001.png"]
    assert out.count("[INFO] Reached the end of dataset. Restarting from the beginning.") == 1
    assert "Episode count: 3" in out and env.episode_num_count == 3
    env.close()
    ld = bh.SyntheticLoader(N, F, 1, seeds=(1, 2, 3))
    env = bh.BinaryHologramEnv(ld.target_function, ld, IPS=N, CH=F, verbose=False)
    res = bh.optimize_with_random_pixel_flips(env, max_datasets=1, rng=np.random.default_rng(0), max_candidates=30,
                                              verbose=False)
    assert [r["file"] for r in res] == ["synthetic_0001", "synthetic_0002"]     # B-9 (and B-10 fixed: own names)
    env.close()
