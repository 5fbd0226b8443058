"""CPU: host logic, the C-ABI library surface and the native FFT check (no GPU needed)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

import binary_hologram_reinforcement_learning_b200 as bh
from binary_hologram_reinforcement_learning_b200 import dbs, dist, engine, envs, spaces, _build
from oracle import hologram_oracle as O


def test_library_builds_loads_and_exports_every_declared_symbol():
    path = _build.build()
    assert os.path.exists(path)
    lib = engine.load_library()
    header = open(os.path.join(ROOT, "include", "bholo.h")).read()
    declared = set(re.findall(r"\b(bh_[a-z_0-9]+)\s*\(", header))
    assert declared == set(engine.ABI_SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None, name
    assert lib.bh_abi_version() == 1
    assert ctypes.sizeof(engine.BhResult) == 40


def test_sass_is_sm100a():
    out = subprocess.run(["cuobjdump", "-lelf", _build.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(bh.HoloError):
        bh.HoloEngine(64, 8, bh.WL_MONO)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "binary_hologram_reinforcement_learning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f


def test_native_fft_passes_on_host(tmp_path):
    """The exact __host__ __device__ pass functions the kernels run, executed on the CPU."""
    exe = str(tmp_path / "host_check")
    src = os.path.join(ROOT, "tests", "native", "host_check.cu")
    subprocess.run(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", exe, src], check=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0 and "HOST_CHECK_OK" in res.stdout, res.stdout


def test_decile_index_matches_oracle():
    v = np.concatenate([np.linspace(0, 1, 101), [0.1, 0.2, 0.30000000001, 1.0, 0.0]])
    got = dbs.decile_index(v)
    exp = np.array([O.decile_of(float(x)) for x in v])
    assert np.array_equal(got, exp)


def test_importance_table_matches_oracle():
    rng = np.random.default_rng(0)
    ch = list(rng.normal(size=500) * 1e-4)
    np.testing.assert_allclose(envs.importance_reward_table(ch), O.importance_reward_table(ch),
                               rtol=1e-12, atol=1e-12)


def test_goal_bonus_constants():
    assert abs(envs.goal_bonus(1.0, -595.2) - 300.04) < 1e-9      # env.py:227-235
    assert abs(envs.goal_bonus(0.5, -595.24) - O.goal_bonus(0.5, -595.24)) < 1e-12


def test_vec_book_update_follows_env_py_without_a_gpu():
    """bh_vec_book_update (the host half of bh_vec_step): env.py:163-167,184-196,214 and appendix B-1/3/6/7."""
    lib = engine.load_library()
    E, stride = 3, 12
    state = np.zeros((E, stride), np.int8); record = np.zeros((E, stride), np.int8)
    record[2, 7] = 127                                         # B-7: the int8 attempt counter wraps
    prev = np.array([10.0, 20.0, 29.95]); init = prev.copy()
    steps = np.zeros(E, np.int64); flips = np.zeros(E, np.int64)
    tdiff = np.full(E, 0.1); tpsnr = np.full(E, 30.0); maxs = np.array([100, 3, 100], np.int64)
    rewards = np.zeros(E); change = np.zeros(E); diff = np.zeros(E)
    last = np.full(E, -7, np.int64); event = np.zeros(E, np.uint8)
    b = engine.VecBook()
    b.state, b.state_record, b.stride = state.ctypes.data, record.ctypes.data, stride
    b.prev_psnr, b.init_psnr, b.steps, b.flips = prev.ctypes.data, init.ctypes.data, steps.ctypes.data, flips.ctypes.data
    b.t_psnr_diff, b.t_psnr, b.max_steps, b.reward_scale = tdiff.ctypes.data, tpsnr.ctypes.data, maxs.ctypes.data, 800.0
    b.rewards, b.psnr_change, b.psnr_diff = rewards.ctypes.data, change.ctypes.data, diff.ctypes.data
    b.last_candidate, b.event = last.ctypes.data, event.ctypes.data

    def update(env_ids, actions, psnr_after, accept):
        n = len(actions)
        res = np.zeros(n, engine.RESULT_DTYPE)
        res["psnr_after"], res["accept"], res["action"] = psnr_after, accept, actions
        ids = np.asarray(env_ids, np.int32); act = np.asarray(actions, np.int64)
        rc = lib.bh_vec_book_update(n, ids.ctypes.data, act.ctypes.data, res.ctypes.data, ctypes.addressof(b))
        assert rc == 0

    # step 1: env0 kept with a big gain (success event), env1 rejected, env2 kept above T_PSNR with diff < 0.1
    update([0, 1, 2], [2, 5, 7], [10.5, 19.9, 30.02], [1, 0, 1])
    np.testing.assert_allclose(rewards, [400.0, -80.0, 0.07 * 800], rtol=0, atol=1e-9)       # env.py:188
    assert list(steps) == [1, 1, 1] and list(flips) == [1, 0, 1]                            # B-6
    assert state[0, 2] == 1 and state[1, 5] == 0 and state[2, 7] == 1                       # rejected flip rolled back
    assert record[0, 2] == 1 and record[1, 5] == 1 and record[2, 7] == -128                 # every attempt counts; wraps
    np.testing.assert_allclose(prev, [10.5, 20.0, 30.02])                                   # env.py:214 / :196
    assert list(last) == [-1, 5, -1] and list(event) == [1, 0, 1]
    np.testing.assert_allclose(diff, [0.5, -0.1, 0.07], atol=1e-12)
    # step 2: a subset in another order; ties are kept by the env rule (decided on the device), no event for env1
    update([1, 0], [5, 2], [20.0, 10.5], [1, 1])
    np.testing.assert_allclose(rewards[:2], [0.0, 0.0], atol=1e-12)
    assert list(steps) == [2, 2, 1] and list(flips) == [2, 1, 1]
    assert state[1, 5] == 1 and state[0, 2] == 0 and record[1, 5] == 2 and record[0, 2] == 2
    assert list(event[:2]) == [0, 1]                           # env0: psnr_diff 0.5 >= T_PSNR_DIFF again
    # step 3: env1 reaches max_steps on a rejected flip -> no event (B-1) ...
    update([1], [6], [19.0], [0])
    assert steps[1] == 3 and event[0] == 0 and last[1] == 6 and prev[1] == 20.0
    # ... and on the next kept flip the max_steps branch has to run although the gain is tiny
    update([1], [6], [20.01], [1])
    assert steps[1] == 4 and event[0] == 1 and flips[1] == 2 and abs(rewards[0] - 8.0) < 1e-9
    # an incomplete book is refused
    b2 = engine.VecBook()
    ids = np.zeros(1, np.int32); act = np.zeros(1, np.int64); res = np.zeros(1, engine.RESULT_DTYPE)
    assert lib.bh_vec_book_update(1, ids.ctypes.data, act.ctypes.data, res.ctypes.data, ctypes.addressof(b2)) == -1
    assert b"incomplete" in lib.bh_last_error(None)


def test_synthetic_inputs_match_oracle_copy():
    a = bh.synthetic_problem(32, 6, 3, 7)
    b = O.synthetic_problem(32, 6, 3, 7)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    ld = bh.SyntheticLoader(32, 6, 3, seeds=(7,))
    tgt, name = next(iter(ld))
    assert tgt.shape == (1, 3, 32, 32) and np.array_equal(ld.target_function(tgt)[0], a[0])


def test_spaces_and_action_mapping():
    ld = bh.SyntheticLoader(32, 6, 3, seeds=(1,))
    env = bh.BinaryHologramEnv(ld.target_function, ld, IPS=32, CH=6, wl=bh.WL_RGB, crop_margin=4,
                               verbose=False)
    assert env.action_space.n == 6 * 32 * 32
    assert env.observation_space["state"].shape == (1, 6, 32, 32)
    assert env.observation_space["recon_image"].shape == (1, 3, 24, 24)
    a = np.array([0, 4 * 32 + 4, 32 * 32 + 5 * 32 + 27, 6 * 32 * 32 - 1])
    sim, inside = env._map_actions(a)
    assert list(inside) == [False, True, True, False]
    assert sim[1] == 0 and sim[2] == 24 * 24 + 1 * 24 + 23


def test_shard_helpers():
    n = 19267584
    cover = []
    for r in range(8):
        lo, hi = dist.shard_range(n, r, 8)
        cover.append((lo, hi))
    assert cover[0][0] == 0 and cover[-1][1] == n
    assert all(cover[i][1] == cover[i + 1][0] for i in range(7))
    assert np.array_equal(np.sort(np.concatenate([dist.shard_indices(13, r, 4) for r in range(4)])),
                          np.arange(13))


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np
from binary_hologram_reinforcement_learning_b200 import dist
dist.init_process_group("gloo")
rank, world, _ = dist.env_info()
rows = np.arange((rank + 1) * 5, dtype=np.float64).reshape(rank + 1, 5) + 100 * rank
allr = dist.gather_episode_stats(rows)
assert allr.shape == (3, 5), allr.shape
assert allr[0, 0] == 0 and allr[1, 0] == 100 and allr[2, 0] == 105
h, = dist.reduce_histograms(np.array([1, 2, 3]) * (rank + 1))
assert list(h) == [3, 6, 9]
assert dist.max_over_ranks(float(rank)) == 1.0
lo, hi = dist.shard_range(10, rank, world)
assert (lo, hi) == ((0, 5) if rank == 0 else (5, 10))
# the sharded score-and-revert sweep of SURVEY 8e, with the oracle-backed engine stand-in: every rank scores
# its contiguous slice of the globally defined order, the decile histograms are all-reduced, and the result
# equals the unsharded sweep
import binary_hologram_reinforcement_learning_b200 as bh
from binary_hologram_reinforcement_learning_b200 import dbs
from tests.oracle_engine import OracleEngine
dbs.HoloEngine = OracleEngine
N, F, m = 16, 4, 2
def sweep(shard):
    ld = bh.SyntheticLoader(N, F, 1, seeds=(61,))
    return bh.dbs_sweep(ld.target_function, ld, 2e-3, 7.56e-6, m, CH=F, wl=bh.WL_MONO, max_datasets=0,
                        rng=np.random.default_rng(1), verbose=False, shard=shard)[0]
mine, full = sweep((rank, world)), sweep(None)
att, imp, gn = dist.reduce_histograms(mine["attempted"], mine["improved"], mine["gains"])
assert np.array_equal(att, full["attempted"]) and np.array_equal(imp, full["improved"])
assert np.allclose(gn, full["gains"], rtol=1e-12, atol=1e-15)
lo, hi = dist.shard_range(full["order"].size, rank, world)
assert np.array_equal(mine["order"], full["order"][lo:hi])
assert np.array_equal(mine["psnr_after"], full["psnr_after"][lo:hi])
dist.barrier()
print("GLOO_OK", rank)
"""


def test_two_rank_gloo_stats_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    import socket
    with socket.socket() as sk:                       # a free rendezvous port
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"GLOO_OK {r}" in o, o


def test_image_folder_loader(tmp_path):
    """Dataset512-shaped loader (DBS.py:172-199): (1, C, N, N) float32 in [0,1] and (path,)."""
    PIL = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(0)
    for i, shape in enumerate([(80, 100, 3), (40, 50, 3)]):
        PIL.fromarray((rng.random(shape) * 255).astype("uint8")).save(tmp_path / f"{i:04d}.png")
    ld = bh.ImageFolderLoader(str(tmp_path), 64)
    items = list(ld)
    assert len(items) == 2 and len(ld) == 2
    t, p = items[0]
    assert t.shape == (1, 3, 64, 64) and t.dtype == np.float32 and 0 <= t.min() and t.max() <= 1
    assert isinstance(p, tuple) and p[0].endswith("0000.png")     # torch default_collate of one string
    assert items[1][0].shape == (1, 3, 64, 64)              # smaller image tiled up, then cropped
    full = bh.load_image(p[0])
    assert np.array_equal(t[0], full[:, 8:72, 18:82])       # centre crop
    g = next(iter(bh.ImageFolderLoader(str(tmp_path), 32, gray=True, random_crop=True, seed=3)))[0]
    assert g.shape == (1, 1, 32, 32)


def test_bench_reference_arm_contract():
    """bench.py --impl reference: exactly one JSON line on stdout with the contract's keys (CPU only)."""
    import json
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and d["vs_baseline"] is None


def test_sweep_shard_partition_is_a_partition():
    """dbs_sweep(shard=(rank, world)) slices one globally defined order."""
    n, world = 1000, 8
    per = (n + world - 1) // world
    cover = np.concatenate([np.arange(min(n, r * per), min(n, (r + 1) * per)) for r in range(world)])
    assert np.array_equal(cover, np.arange(n))
    assert [dist.shard_range(n, r, world) for r in range(world)] == \
        [(min(n, r * per), min(n, (r + 1) * per)) for r in range(world)]


def test_variant_modules_and_multidiscrete_action_space():
    """compat/env_md.py (MultiDiscrete action, env_md.py:54) and env_05.py (T_PSNR_DIFF 0.5) resolve."""
    compat = os.path.join(ROOT, "binary_hologram_reinforcement_learning_b200", "compat")
    sys.path.insert(0, compat)
    try:
        import env_md, env_05
    finally:
        sys.path.remove(compat)
    ld = bh.SyntheticLoader(32, 4, 1, seeds=(1,))
    e = env_md.BinaryHologramEnv(ld.target_function, ld, IPS=32, CH=4, verbose=False)
    assert list(e.action_space.nvec) == [4, 32, 32]
    e5 = env_05.BinaryHologramEnv(ld.target_function, ld, IPS=32, CH=4, verbose=False)
    assert e5.T_PSNR_DIFF == 0.5 and e5.action_space.n == 4 * 32 * 32


def test_cpulist_parser_and_numa_binding_is_a_noop_without_topology(monkeypatch):
    """dist.bind_to_gpu_numa_node: sysfs cpulists parse ("0-3,8,10-11"), and without a GPU / exposed topology the
    call changes nothing and returns None (the B200 boxes of this pool are single-node VMs)."""
    from binary_hologram_reinforcement_learning_b200 import dist as D
    assert D._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert D._parse_cpulist("") == set()
    assert D._parse_cpulist("5") == {5}
    before = os.sched_getaffinity(0)
    assert D.bind_to_gpu_numa_node(0) is None
    assert os.sched_getaffinity(0) == before
    # a node whose cores are outside the container's cpuset must not empty the affinity mask
    monkeypatch.setattr(D, "gpu_numa_node", lambda local_rank: 10 ** 6)
    assert D.bind_to_gpu_numa_node(0) is None
    assert os.sched_getaffinity(0) == before


def test_bench_cpu_arm_reports_stock_and_tuned_allocator(monkeypatch):
    """The reference arm prints the stock figure as `value` and the allocator-tuned best case beside it; the
    mallopt switch is reversible (bench.tune_host_allocator)."""
    import importlib
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    bench = importlib.import_module("bench")
    assert bench.tune_host_allocator(True) is True
    assert bench.tune_host_allocator(False) is True
    calls = []

    def fake_rate(seconds_budget=None, samples=None, sample_steps=8, warm=1):
        calls.append(samples)
        return (10.0 if len(calls) == 1 else 25.0), [0.8] * (samples or 1), 4
    monkeypatch.setattr(bench, "cpu_reference_rate", fake_rate)
    lines = []
    monkeypatch.setattr(bench, "emit", lambda line: lines.append(line))

    class A:
        steps, warmup, gpus, envs, rollout = 3, 1, 1, 8, 512
    assert bench.run_reference(A()) == 0
    line = lines[0]
    assert line["impl"] == "reference" and line["value"] == 10.0 and line["e2e"]["value"] == 10.0
    assert line["cpu_baseline"]["tuned_allocator"]["value"] == 25.0
    assert line["config"] == bench.bench_config(8, 512)
