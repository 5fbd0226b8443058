#!/usr/bin/env python
"""Benchmark of the hologram reward / DBS hot path (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): env_1024_24 semantics, 1024 x 1024 pixels,
24 binary frames in 3 colour groups, E = 8 vectorised environments per GPU,
uniformly random single-pixel flips, keep iff dPSNR >= 0 (env.py:184-196).
One bench "step" = one rollout of ROLLOUT = 512 vectorised env steps
(train-PPO.py:300 n_steps=512), i.e. 4096 env steps per GPU.

  value   env steps/s, device-resident inputs (actions pre-uploaded), CUDA events
  e2e     same metric through HologramVecEnv.step(): host numpy actions copied
          from pinned memory every step, results (PSNR, accept) read back every
          step, Python reward logic included
  roofline  delta-eval kernel k_eval: 16*N^2 algorithmic HBM bytes per candidate
          (U 8 B/px + I 4 + T 4; the shifted impulse response is L2 resident)
  cpu_baseline  the reference-shaped torch CPU path (oracle/torch_path.py)

Under torchrun (N > 1) every rank owns one GPU and its own 8 envs (weak
scaling, no data-path collective); NCCL is used for the max-over-ranks timing
and one all-gather of episode statistics after the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_SIDE, FRAMES, GROUPS = 1024, 24, 3
ENVS_PER_GPU = 8
ROLLOUT = 512
WORKLOAD = ("env_1024_24 x 8 vectorised envs per GPU: 1024x1024, 24 binary frames in 3 colour "
            "groups (638/515/450 nm), random single-pixel flips, keep iff dPSNR>=0; "
            "1 step = rollout of 512 vectorised env steps (4096 env steps per GPU)")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--rollout", type=int, default=ROLLOUT)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed regions run."""
    FIELDS = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.lines, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-i", str(self.device), "-lms", "100"], stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, windows):
        if self.proc is not None:
            self.proc.terminate()
        rows = [l for t, l in self.lines if any(a - 0.05 <= t <= b + 0.05 for a, b in windows)]
        if not rows:
            rows = [l for _, l in self.lines]
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            p = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(p[2])); smax.append(float(p[3]))
            except Exception:
                continue
            for name, val in zip(names, p[6:10]):
                if val.lower() == "active":
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)),
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic_bytes():
    """Per-launch DRAM bytes of k_eval from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        with open(path) as f:
            d = json.load(f)
            return d.get("k_eval_dram_bytes_per_launch_warm_cache", d.get("k_eval_dram_bytes_per_launch"))
    except Exception:
        return None


# ---------------------------------------------------------------------------
# CPU arm: the reference-shaped torch path on the host cores
# ---------------------------------------------------------------------------
def cpu_reference_rate(seconds_budget=None, samples=None, sample_steps=8, warm=1):
    """Env steps/s of the torch CPU restatement at the bench workload.

    Either runs for ~seconds_budget (cpu_baseline leg) or exactly `samples`
    samples of `sample_steps` env steps (reference arm).  Returns (rate, per-sample
    seconds list, cores).
    """
    import torch
    from oracle.torch_path import TorchRefEnv
    from oracle import hologram_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    pre, tgt = O.synthetic_problem(N_SIDE, FRAMES, GROUPS, seed=0)
    env = TorchRefEnv(N_SIDE, FRAMES, O.WL_RGB, threads=cores)
    env.reset(pre, tgt)
    rng = np.random.default_rng(12345)
    n2 = FRAMES * N_SIDE * N_SIDE
    for _ in range(warm * sample_steps):
        env.step(int(rng.integers(0, n2)))
    times = []
    t_start = time.perf_counter()
    while True:
        t0 = time.perf_counter()
        for _ in range(sample_steps):
            env.step(int(rng.integers(0, n2)))
        times.append(time.perf_counter() - t0)
        if samples is not None and len(times) >= samples:
            break
        if samples is None and time.perf_counter() - t_start >= seconds_budget:
            break
    rate = sample_steps * len(times) / sum(times)
    return rate, times, cores


def comparator_rates(seconds=4.0):
    """Other shapes of the same reference-shaped torch path, part of the cpu_baseline leg:
    the 256^2 x 8 full step on the host cores, and the call sequence "simply run on the B200"
    (torch.fft = cuFFT, the colour group re-uploaded every step as env.py:170 does, blocking .item())."""
    import torch
    from oracle.torch_path import TorchRefEnv
    from oracle import hologram_oracle as O

    def rate(N, F, wl, device):
        pre, tgt = O.synthetic_problem(N, F, len(wl), 0)
        env = TorchRefEnv(N, F, wl, device=device, threads=os.cpu_count())
        env.reset(pre, tgt)
        rng = np.random.default_rng(5)
        for _ in range(5):
            env.step(int(rng.integers(0, F * N * N)))
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            env.step(int(rng.integers(0, F * N * N)))
            n += 1
        if device != "cpu":
            torch.cuda.synchronize()
        return n / (time.perf_counter() - t0)

    out = {"torch_cpu_256x8_full_step_env_steps_per_s": rate(256, 8, O.WL_MONO, "cpu")}
    if torch.cuda.is_available():
        out["torch_on_b200_1024x24_single_group_env_steps_per_s"] = rate(N_SIDE, FRAMES, O.WL_RGB, "cuda")
        out["torch_on_b200_256x8_full_step_env_steps_per_s"] = rate(256, 8, O.WL_MONO, "cuda")
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    # one sample = one vectorised step of 8 envs on the CPU path; shrunk for very long runs so that
    # K + W samples still finish within a few minutes (~35 ms per env step on 16 cores)
    sample_steps = 8 if args.steps + args.warmup <= 400 else max(1, 3200 // (args.steps + args.warmup))
    import torch  # noqa: F401
    rate, times, cores = cpu_reference_rate(samples=args.steps, sample_steps=sample_steps,
                                            warm=max(1, args.warmup))
    ms = 1000.0 * sum(times) / len(times)
    sample = (f"{len(times)} timed samples of {sample_steps} env steps of one 1024x1024x24 env "
              f"(per-colour-group re-simulation, float32 torch CPU, {cores} threads); "
              f"ms_per_step is per sample, not per 4096-step rollout")
    line = {
        "impl": "reference", "metric": "env_steps_per_s", "value": rate, "unit": "env steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "N": N_SIDE, "frames": FRAMES,
                                        "groups": GROUPS, "pad": 1, "relative": True},
        "cpu_baseline": {"value": rate, "unit": "env steps/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": rate, "unit": "env steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ---------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------
def run_b200(args):
    import torch
    import binary_hologram_reinforcement_learning_b200 as bh
    from binary_hologram_reinforcement_learning_b200 import dist as bdist
    from binary_hologram_reinforcement_learning_b200.engine import RULE_ENV, RESULT_DTYPE

    rank, world, local = bdist.env_info()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        bdist.init_process_group("nccl")
    E, R, K, W = args.envs, args.rollout, args.steps, max(args.warmup, 0)
    n_pix = FRAMES * N_SIDE * N_SIDE

    # ---- environments (public API) --------------------------------------
    loaders = [bh.SyntheticLoader(N_SIDE, FRAMES, GROUPS, seeds=(rank * E + i,)) for i in range(E)]

    def target_function(t):
        key = np.ascontiguousarray(t[0, 0, 0, :4], dtype=np.float32).tobytes()
        for ld in loaders:
            if key in ld._pre:
                return ld._pre[key][None]
        raise KeyError("unknown synthetic target")

    vec = bh.HologramVecEnv(E, target_function, loaders, max_steps=10 ** 9, T_PSNR_DIFF=1e9,
                            IPS=N_SIDE, CH=FRAMES, wl=bh.WL_RGB, device=local, recon_obs="lazy",
                            obs_mode="views", verbose=False, seed=rank)
    eng = vec.engine
    stream = torch.cuda.Stream(device=local)
    eng.set_stream(stream.cuda_stream)
    vec.reset()
    psnr0 = [e.initial_psnr for e in vec.envs]

    rng = np.random.default_rng(1000 + rank)
    sampler = ClockSampler(local)
    sampler.start()
    windows = []

    # ---- e2e: HologramVecEnv.step with host actions ----------------------
    e2e_steps = R * K
    acts_host = rng.integers(0, n_pix, size=(R * (K + W), E), dtype=np.int64)
    for i in range(R * W):
        vec.step(acts_host[i])
    torch.cuda.synchronize()
    bdist.barrier()
    t0w = time.time()
    t0 = time.perf_counter()
    reward_sum = 0.0
    for i in range(R * W, R * (W + K)):
        _, rewards, _, _ = vec.step(acts_host[i])
        reward_sum += float(rewards[0])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    windows.append((t0w, time.time()))
    bdist.barrier()
    e2e_s = bdist.max_over_ranks(e2e_s)
    e2e_value = world * E * e2e_steps / e2e_s
    h2d = R * E * (8 + 4)                       # actions int64 + env ids int32 per vec step
    d2h = R * E * RESULT_DTYPE.itemsize         # bh_result records per vec step

    # ---- e2e with the reference's eager observation: recon_image copied to the host every step
    n_eager = 64
    t0 = time.perf_counter()
    for i in range(n_eager):
        vec.step(acts_host[i % len(acts_host)])
        for j in range(E):
            vec.refresh_recon(j)                  # 12.6 MB D2H per env into pinned memory
    torch.cuda.synchronize()
    eager_s = bdist.max_over_ranks(time.perf_counter() - t0)

    # ---- value: device-resident actions, CUDA events ---------------------
    d_acts = torch.from_numpy(rng.integers(0, n_pix, size=(R * (K + W), E), dtype=np.int64)).cuda(local)
    d_envs = torch.arange(E, dtype=torch.int32, device=f"cuda:{local}")
    d_res = torch.zeros((R, E, RESULT_DTYPE.itemsize), dtype=torch.uint8, device=f"cuda:{local}")
    a_ptr, e_ptr, r_ptr = d_acts.data_ptr(), d_envs.data_ptr(), d_res.data_ptr()
    row_a, row_r = E * 8, E * RESULT_DTYPE.itemsize

    def rollout(base_row):
        for s in range(R):
            eng.step_batch_device(E, e_ptr, a_ptr + (base_row + s) * row_a, RULE_ENV, r_ptr + s * row_r)

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        for w in range(W):
            rollout(w * R)
        torch.cuda.synchronize()
        bdist.barrier()
        launches0 = eng.launch_count
        t0w = time.time()
        ev0.record(stream)
        for k in range(K):
            rollout((W + k) * R)
        ev1.record(stream)
        torch.cuda.synchronize()
        windows.append((t0w, time.time()))
    bdist.barrier()
    launches = eng.launch_count - launches0
    dev_ms = bdist.max_over_ranks(ev0.elapsed_time(ev1))
    value = world * E * R * K / (dev_ms / 1000.0)
    res_host = d_res.cpu().numpy().view(RESULT_DTYPE).reshape(R, E)
    accept_rate = float(res_host["accept"].mean())

    # ---- roofline of the dominant kernel (k_eval), measured live ---------
    n_sets = 64
    d_sets = torch.from_numpy(rng.integers(0, n_pix, size=(n_sets, E), dtype=np.int64)).cuda(local)
    with torch.cuda.stream(stream):
        t0w = time.time()
        eval_ms = eng.time_eval(E, e_ptr, d_sets.data_ptr(), n_sets, 512)
        windows.append((t0w, time.time()))
    alg_bytes = 16.0 * N_SIDE * N_SIDE * E
    achieved = alg_bytes / (eval_ms / 1000.0) / 1e9
    peak, peak_src = measured_peak_gbs()

    # ---- secondary figures: sweep flip-evals/s, propagation ---------------
    n_cand = 8192
    cand = rng.integers(0, n_pix, size=n_cand, dtype=np.int64)
    eng.eval_flips(cand[:1024], env=0)
    t0 = time.perf_counter()
    eng.eval_flips(cand, env=0)
    sweep_s = time.perf_counter() - t0
    with torch.cuda.stream(stream):
        prop_ms = eng.time_propagate(0, 5)
        prop_pass_ms = eng.time_propagate_passes(0, 3)
    # exhaustive sweep of all 24*1024^2 candidates of env 0 by FFT correlation (device output)
    d_map = torch.empty(n_pix, dtype=torch.float64, device=f"cuda:{local}")
    eng.sweep_all_device(0, d_map.data_ptr())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.sweep_all_device(0, d_map.data_ptr())
    torch.cuda.synchronize()
    sweep_all_s = time.perf_counter() - t0
    # greedy DBS (device loop, speculative batches) on env 1
    order = rng.permutation(n_pix)[:20000]
    t0 = time.perf_counter()
    _, _, dbs_nacc, _ = eng.dbs_run(order, env=min(1, E - 1), k_spec=0, resync_every=0)
    dbs_s = time.perf_counter() - t0
    prop_bytes = FRAMES * (N_SIDE ** 2 + 40.0 * N_SIDE ** 2) + 8.0 * GROUPS * N_SIDE ** 2
    clocks = sampler.stop(windows)

    # ---- the one collective: all-gather of per-env episode statistics ----
    vec.sync_envs()
    stats = np.array([[vec._ep_reward[i], vec.envs[i].steps, vec.envs[i].flip_count, psnr0[i],
                       vec.envs[i].previous_psnr] for i in range(E)])
    all_stats = bdist.gather_episode_stats(stats)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, times, cores = cpu_reference_rate(seconds_budget=args.cpu_seconds)
        cpu = {"value": rate, "unit": "env steps/s", "cores": cores, "kind": "port",
               "sample": (f"{8 * len(times)} env steps of one 1024x1024x24 env, per-colour-group "
                          f"re-simulation, float32 torch CPU ({cores} threads), "
                          f"{sum(times):.1f} s (oracle/torch_path.py)"),
               "comparators": comparator_rates()}

    if rank == 0:
        line = {
            "metric": "env_steps_per_s", "value": value, "unit": "env steps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": dev_ms / max(K, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": E, "rollout": R, "N": N_SIDE,
                       "frames": FRAMES, "groups": GROUPS, "pad": 1, "relative": True,
                       "l2": ("inputs larger than L2: 8 envs x 226 MB resident per GPU, every step "
                              "streams a random frame (16.8 MB) of each env"),
                       "episodes": "max_steps raised so no episode ends inside the timed region",
                       "accept_rate": accept_rate},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic_bytes(),
                         "kernel": "k_eval (delta evaluation of 8 candidates, one per env)",
                         "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": eval_ms,
                         "peak_source": peak_src},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "env steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "api": "HologramVecEnv.step (bh_vec_step)",
                    "checksum_reward_env0": reward_sum,
                    "with_eager_recon_obs": {"value": world * E * n_eager / eager_s, "unit": "env steps/s",
                                             "d2h_bytes_per_env_step": 4 * GROUPS * N_SIDE * N_SIDE}},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "extra": {
                "flip_evals_per_s_kernel": E / (eval_ms / 1000.0) * world,
                "flip_evals_per_s_sweep_api": n_cand / sweep_s * world,
                "flip_evals_per_s_exhaustive_sweep_all": n_pix / sweep_all_s * world,
                "sweep_all_ms_25M_candidates": sweep_all_s * 1e3,
                "dbs_greedy_candidates_per_s": len(order) / dbs_s * world,
                "dbs_greedy_accept_rate": dbs_nacc / len(order),
                "propagate_ms_24_frames": prop_ms,
                "propagate_pass_ms": prop_pass_ms,
                "propagate_gbs_algorithmic": prop_bytes / (prop_ms / 1000.0) / 1e9,
                "episode_stats_rows_gathered": int(all_stats.shape[0]),
                "psnr_env0_initial": psnr0[0],
            },
        }
        emit(line)
    vec.close()
    if world > 1:
        import torch.distributed as tdist
        tdist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def quiet_stdout():
    """Route fd 1 to stderr so library banners (NCCL prints its version on stdout) cannot
    precede the JSON line; emit() writes the line to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    args = parse()
    quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1", "--master-port", "29517",
               os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
