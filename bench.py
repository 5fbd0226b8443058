#!/usr/bin/env python
"""Benchmark of the hologram reward / DBS hot path (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): env_1024_24 semantics, 1024 x 1024 pixels,
24 binary frames in 3 colour groups, E = 8 vectorised environments per GPU,
uniformly random single-pixel flips, keep iff dPSNR >= 0 (env.py:184-196).
One bench "step" = one rollout of ROLLOUT = 512 vectorised env steps
(train-PPO.py:300 n_steps=512), i.e. 4096 env steps per GPU.

  value     env steps/s, device-resident inputs (actions pre-uploaded), CUDA events
  e2e       same metric through HologramVecEnv.step() as the reference returns it (env.py:176-181):
            host numpy actions in, results (PSNR, accept) AND obs["recon_image"] of every env current on
            the host after every step (recon_obs="eager", the default); e2e.lazy_obs / e2e.device_obs are
            the same loop without the observation copy / with a device-resident observation
  roofline  delta-eval kernel k_eval: 16*N^2 algorithmic HBM bytes per candidate
            (U 8 B/px + I 4 + T 4; the shifted impulse response is L2 resident); roofline_commit
            (24*N^2 per kept flip) and roofline_propagate (N^2 + 32 P^2 per frame) beside it
  dbs       greedy DBS candidates/s at three accept-rate regimes (metric's second half)
  cpu_baseline  the reference-shaped torch CPU path (oracle/torch_path.py) + comparators
  parity_check  the first vectorised steps of two envs replayed on the float64 oracle

Under torchrun (N > 1) every rank owns one GPU and its own 8 envs (weak
scaling, no data-path collective); NCCL is used for the max-over-ranks timing,
one all-gather of episode statistics and one all-reduce of sweep histograms after the timed region.
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


def bench_config(envs: int, rollout: int) -> dict:
    """The workload description: identical for the repo arm and the reference arm."""
    return {"workload": WORKLOAD, "envs_per_gpu": envs, "rollout": rollout, "N": N_SIDE, "frames": FRAMES,
            "groups": GROUPS, "pad": 1, "relative": True, "recon_obs": "eager (env.py:176-181)",
            "l2": ("inputs larger than L2: 8 envs x 226 MB resident per GPU, every step streams a random "
                   "frame (16.8 MB) of each env"),
            "episodes": "max_steps raised so no episode ends inside the timed region",
            "value_path": ("open-loop: the rollout's pre-drawn random actions are device resident and the 512 vectorised "
                           "steps run in one persistent launch (bh_rollout_device); e2e: one HologramVecEnv.step call "
                           "per vectorised step, host actions in, results + recon_image out")}


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
    ap.add_argument("--no-extras", action="store_true", help="skip the dbs / group / sweep blocks")
    ap.add_argument("--dbs-full", action="store_true", help="also time one full greedy pass (25.2 M candidates)")
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


def ncu_traffic_bytes(key=None):
    """Per-launch DRAM bytes of a kernel from the committed ncu capture, if any (default: k_eval)."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        with open(path) as f:
            d = json.load(f)
            if key:
                return d.get(key)
            return d.get("k_eval_dram_bytes_per_launch_warm_cache", d.get("k_eval_dram_bytes_per_launch"))
    except Exception:
        return None


# ---------------------------------------------------------------------------
# CPU arm: the reference-shaped torch path on the host cores
# ---------------------------------------------------------------------------
def tune_host_allocator(on=True):
    """glibc serves the 64 MB temporaries of every torch CPU step by mmap / munmap (page faults on first
    touch, every step); raising M_MMAP_THRESHOLD recycles them from the heap.  Measured: 2-3 x on the
    torch CPU path.  Process-wide, so the stock figure is always taken first."""
    import ctypes
    try:
        libc = ctypes.CDLL("libc.so.6")
        if on:
            libc.mallopt(-3, 1 << 30)                 # M_MMAP_THRESHOLD
            libc.mallopt(-1, (1 << 31) - 1)           # M_TRIM_THRESHOLD
        else:
            libc.mallopt(-3, 128 * 1024)
            libc.mallopt(-1, 128 * 1024)
        return True
    except Exception:
        return False


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
    the 256^2 x 8 full step on the host cores, the call sequence "simply run on the B200"
    (torch.fft = cuFFT, the colour group re-uploaded every step as env.py:170 does, blocking .item()),
    and the DEVICE-RESIDENT torch.fft propagation of the 24 x 1024^2 stack (+ abs^2 + mean + relative loss) --
    the on-GPU bar SURVEY 2.1 sets for the hand-written FFT passes."""
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
        # device-resident propagation with cuFFT: nothing crosses PCIe inside the timed region
        pre, tgt = O.synthetic_problem(N_SIDE, FRAMES, GROUPS, 0)
        x = torch.from_numpy((pre >= 0.5).astype(np.float32)).cuda()
        T = torch.from_numpy(tgt).cuda()
        H = torch.stack([torch.from_numpy(O.transfer_function(N_SIDE, O.PIXEL_PITCH, w, O.Z_DEFAULT)
                                          .astype(np.complex64)) for w in O.WL_RGB]).cuda()
        Fg = FRAMES // GROUPS

        def propagate():
            means = []
            for g in range(GROUPS):
                U = torch.fft.ifft2(torch.fft.fft2(x[g * Fg:(g + 1) * Fg]) * H[g])
                means.append((U.abs() ** 2).mean(dim=0))
            I = torch.stack(means)
            s = (I * T).sum() / (I * I).sum()
            return 10.0 * torch.log10(1.0 / ((s * I - T) ** 2).mean())

        for _ in range(3):
            propagate()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            p = propagate()
        e1.record()
        torch.cuda.synchronize()
        out["torch_fft_device_resident_propagate_ms_24_frames"] = e0.elapsed_time(e1) / 10.0
        out["torch_fft_device_resident_psnr"] = float(p.item())
    return out


def oracle_replay(seed: int, actions: np.ndarray):
    """float64 replay of one env's first steps (oracle/hologram_oracle.py: propagate_group + delta_terms +
    closed-form loss; env.py:154-196 decision rule).  Returns (rewards, accepted)."""
    from oracle import hologram_oracle as O
    cfg = O.HoloConfig(N=N_SIDE, F=FRAMES, wl=O.WL_RGB)
    pre, tgt = O.synthetic_problem(N_SIDE, FRAMES, GROUPS, seed=seed)
    st = (pre >= 0.5).astype(np.int8)
    t64 = tgt.astype(np.float64)
    U = [O.propagate_group(cfg, st[g * cfg.Fg:(g + 1) * cfg.Fg], g) for g in range(cfg.G)]
    I = [O.group_mean_intensity(u) for u in U]
    sii, sit, stt = O.loss_sums(np.stack(I), t64)
    n = cfg.G * N_SIDE * N_SIDE
    prev = O.psnr_from_mse(O.mse_from_sums(sii, sit, stt, n, True))
    rewards, accepted = [], []
    for a in actions:
        f, r, c = cfg.decode(int(a))
        g = cfg.group_of(f)
        s = 1 - 2 * int(st[f, r, c])
        d_sii, d_sit, dI = O.delta_terms(cfg, U[g][f - g * cfg.Fg], I[g], t64[g], g, r, c, s)
        after = O.psnr_from_mse(O.mse_from_sums(sii + d_sii, sit + d_sit, stt, n, True))
        change = after - prev
        rewards.append(change * 800.0)
        acc = not (change < 0)                                  # env.py:191
        accepted.append(acc)
        if acc:
            h = cfg.h(g)
            yy = (np.arange(N_SIDE) - r) % cfg.P
            xx = (np.arange(N_SIDE) - c) % cfg.P
            U[g][f - g * cfg.Fg] += s * h[np.ix_(yy, xx)]
            I[g] = I[g] + dI
            sii, sit, prev = sii + d_sii, sit + d_sit, after
            st[f, r, c] = 1 - st[f, r, c]
    return np.array(rewards), np.array(accepted)


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
    tuned = None
    if tune_host_allocator(True):                     # after the stock figure: same routine, heap-recycled temporaries
        r2, t2, _ = cpu_reference_rate(samples=max(2, min(args.steps, 8)), sample_steps=sample_steps, warm=1)
        tuned = {"value": r2, "unit": "env steps/s", "samples": len(t2),
                 "what": "same routine after mallopt(M_MMAP_THRESHOLD = 1 GiB): the 64 MB temporaries of a step are "
                         "recycled from the heap instead of mmap / munmap + page faults per step; not what the "
                         "reference's scripts do, reported as the best case of the CPU path"}
        tune_host_allocator(False)
    sample = (f"{len(times)} timed samples of {sample_steps} env steps of one 1024x1024x24 env "
              f"(per-colour-group re-simulation, float32 torch CPU, {cores} threads); "
              f"ms_per_step is per sample, not per 4096-step rollout")
    line = {
        "impl": "reference", "metric": "env_steps_per_s", "value": rate, "unit": "env steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": bench_config(args.envs, args.rollout),
        "cpu_baseline": {"value": rate, "unit": "env steps/s", "cores": cores, "kind": "port",
                         "sample": sample, "allocator": "glibc defaults (fresh process, as the reference's scripts run)",
                         "tuned_allocator": tuned},
        "e2e": {"value": rate, "unit": "env steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ---------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------
def make_vec(bh, E, rank, local, recon_obs, N=N_SIDE, F=FRAMES, wl=None, seed0=0, **kw):
    wl = bh.WL_RGB if wl is None else wl
    loaders = [bh.SyntheticLoader(N, F, len(wl), seeds=(seed0 + rank * E + i,)) for i in range(E)]

    def target_function(t):
        key = np.ascontiguousarray(t[0, 0, 0, :4], dtype=np.float32).tobytes()
        for ld in loaders:
            if key in ld._pre:
                return ld._pre[key][None]
        raise KeyError("unknown synthetic target")

    return bh.HologramVecEnv(E, target_function, loaders, max_steps=10 ** 9, T_PSNR_DIFF=1e9, IPS=N, CH=F,
                             wl=wl, device=local, recon_obs=recon_obs, obs_mode="views", verbose=False,
                             seed=rank, **kw)


def dbs_block(eng, env, rng, peak, torch, n_env):
    """Greedy DBS (DBS_1024_24.py:313-422) candidates/s at three accept-rate regimes of ONE 1024^2 x 24 target,
    plus the dataset form (bh_dbs_run_batch: one candidate per image and launch, n_env images in flight).

    fresh: the random initial hologram.  mid / late: after rounds of a parallel pre-pass that is NOT the
    reference's algorithm, only a quick way to a well-optimised state for the measurement: score every pixel
    (bh_sweep_all), flip the best improving pixel of every 32 x 32 tile of every frame directly in the
    device-resident state, re-propagate (a round is undone if it lowered the PSNR).  Roofline per regime:
    (16 N^2 per scored candidate + 24 N^2 per kept flip) / time -- discarded speculation is not credited."""
    from binary_hologram_reinforcement_learning_b200.engine import DeviceArray
    n2 = N_SIDE * N_SIDE
    n_pix = FRAMES * n2
    d_map = torch.empty(n_pix, dtype=torch.float64, device="cuda")
    state_all = torch.as_tensor(DeviceArray(eng.device_ptr("state"), (eng.n_env, FRAMES, N_SIDE, N_SIDE), "|i1", owner=eng),
                                device="cuda")
    state = state_all[env]
    out = {}

    def measure(label, n_cand):
        order = np.unique(rng.integers(0, n_pix, size=n_cand))
        rng.shuffle(order)
        eng.stream_sync()
        t0 = time.perf_counter()
        acc, _, nacc, psnr = eng.dbs_run(order, env=env, k_spec=0, resync_every=1024)
        dt = time.perf_counter() - t0
        useful = (order.size * 16.0 + nacc * 24.0) * n2
        out[label] = {"candidates": int(order.size), "candidates_per_s": order.size / dt,
                      "accept_rate": nacc / order.size, "psnr": psnr,
                      "hbm_frac_useful_bytes": useful / dt / 1e9 / peak}

    def prepass(rounds, tile=32):
        for _ in range(rounds):
            eng.sweep_all_device(env, d_map.data_ptr())
            eng.stream_sync()
            p0 = eng.metrics(env)[0]
            oy, ox = (int(v) for v in rng.integers(0, tile, size=2))
            gain = torch.roll(d_map.view(FRAMES, N_SIDE, N_SIDE) - p0, shifts=(oy, ox), dims=(1, 2))
            t = gain.view(FRAMES, N_SIDE // tile, tile, N_SIDE // tile, tile).permute(0, 1, 3, 2, 4)
            t = t.reshape(FRAMES, N_SIDE // tile, N_SIDE // tile, tile * tile)
            best, arg = t.max(dim=-1)
            f, ty, tx = torch.nonzero(best > 0, as_tuple=True)
            if f.numel() == 0:
                break
            a = arg[f, ty, tx]
            y = (ty * tile + a // tile - oy) % N_SIDE
            x = (tx * tile + a % tile - ox) % N_SIDE
            state[f, y, x] = 1 - state[f, y, x]
            torch.cuda.synchronize()
            eng.resync(env)
            if eng.metrics(env)[0] < p0:                 # interacting flips: undo the round
                state[f, y, x] = 1 - state[f, y, x]
                torch.cuda.synchronize()
                eng.resync(env)
                tile = min(128, tile * 2)

    measure("fresh", 40000)
    t0 = time.perf_counter()
    prepass(40)
    measure("mid", 40000)
    prepass(160)
    out["prepass_seconds"] = time.perf_counter() - t0
    measure("late", 40000)
    # dataset form: n_env images in flight, no speculation (state of the envs as the rollouts left it)
    n_it = 6000
    orders = np.stack([rng.permutation(n_pix)[:n_it] for _ in range(n_env)])
    eng.stream_sync()
    t0 = time.perf_counter()
    acc, _, nacc, _ = eng.dbs_run_batch(orders, resync_every=0)
    dt = time.perf_counter() - t0
    out["batch_images_in_flight"] = {"images": n_env, "candidates": int(orders.size), "candidates_per_s": orders.size / dt,
                                     "accept_rate": float(nacc.sum()) / orders.size,
                                     "hbm_frac_useful_bytes": (orders.size * 16.0 + float(nacc.sum()) * 24.0) * n2 / dt / 1e9 / peak}
    return out


def group_block(bh, bdist, rank, local):
    """BASELINE configs[4]: env_group reward (10 000 scored candidates per reset, env_group.py:90-143) with
    GRPO-style groups of M = 8 members cloned from one reset state, 512-step rollouts (lazy observation)."""
    out = {}
    for name, N, F, wl, E in (("256x8_64envs", 256, 8, bh.WL_MONO, 64), ("1024x24_8envs", 1024, 24, bh.WL_RGB, 8)):
        M = 8
        loaders = [bh.SyntheticLoader(N, F, len(wl), seeds=(9000 + rank * E + i // M,)) for i in range(E)]

        def tf(t, loaders=loaders):
            key = np.ascontiguousarray(t[0, 0, 0, :4], dtype=np.float32).tobytes()
            for ld in loaders:
                if key in ld._pre:
                    return ld._pre[key][None]
            raise KeyError

        vec = bh.HologramVecEnv(E, tf, loaders, max_steps=10 ** 9, T_PSNR_DIFF=1e9, IPS=N, CH=F, wl=wl,
                                device=local, reward_mode="group", recon_obs="lazy", verbose=False, seed=rank)
        vec.reset_groups(M)                              # warm (tables, allocations)
        vec.engine.stream_sync()
        cand = np.random.default_rng(5 + rank).integers(0, F * N * N, size=10000)
        vec.engine.eval_flips(cand, env=0)
        t0 = time.perf_counter()
        vec.engine.eval_flips(cand, env=0)               # the scoring itself (env_group.py:96-119)
        t_score = time.perf_counter() - t0
        bdist.barrier()
        t0 = time.perf_counter()
        vec.reset_groups(M)
        vec.engine.stream_sync()
        t_reset = bdist.max_over_ranks(time.perf_counter() - t0)
        rng = np.random.default_rng(77 + rank)
        acts = rng.integers(0, F * N * N, size=(512 + 32, E), dtype=np.int64)
        for i in range(32):
            vec.step(acts[i])
        vec.engine.stream_sync()
        bdist.barrier()
        t0 = time.perf_counter()
        for i in range(32, 32 + 512):
            vec.step(acts[i])
        vec.engine.stream_sync()
        t_roll = bdist.max_over_ranks(time.perf_counter() - t0)
        world = bdist.env_info()[1]
        out[name] = {"envs_per_gpu": E, "group_size": M,
                     "reset_groups_s_incl_host_mirrors_and_synthetic_targets": t_reset,
                     "importance_scoring_ms_10000_candidates": 1e3 * t_score,
                     "candidates_scored_per_s": world * 10000 / t_score,
                     "rollout_env_steps_per_s": world * E * 512 / t_roll}
        vec.close()
    return out


def sharded_sweep_block(bh, bdist, rank, world, local, per_rank=2):
    """BASELINE configs[3]: the 64-pixel-crop sweep of 1024^2 x 24 targets (dbs-1024-1024-24-6464.py:330-395),
    TARGETS sharded over the ranks (north_star: "independent environments and target images are sharded
    across the 8 B200s"): every rank sweeps its own `per_rank` images -- all 19.3 M candidates of an image in
    one device call, decile statistics on the device -- and the decile histograms of all images meet in one
    NCCL all-reduce.  Weak scaling: the images per rank are fixed.  (The candidate-sharded form of ONE image,
    dbs_sweep(shard=(rank, world)), is covered by the 2-rank gloo test; a whole image is 2.5 ms of device
    time, so splitting it buys nothing.)"""
    ld = bh.SyntheticLoader(N_SIDE, FRAMES, GROUPS, seeds=tuple(4242 + rank * per_rank + i for i in range(per_rank)))
    items = list(ld)                                   # synthetic targets generated outside the timed region
    bdist.barrier()
    t0 = time.perf_counter()
    rs = bh.dbs_sweep(ld.target_function, items, 2e-3, 7.56e-6, 64, CH=FRAMES, wl=bh.WL_RGB, max_datasets=per_rank - 1,
                      rng=np.random.default_rng(9 + rank), verbose=False, device=local)
    att = np.sum([r["attempted"] for r in rs], axis=0)
    imp = np.sum([r["improved"] for r in rs], axis=0)
    gains = np.sum([r["gains"] for r in rs], axis=0)
    t1 = time.perf_counter()
    att, imp, gains = bdist.reduce_histograms(att, imp, gains)
    t2 = time.perf_counter()
    dt = bdist.max_over_ranks(t2 - t0)
    n_img = FRAMES * 896 * 896
    images = per_rank * world
    return {"images": images, "images_per_rank": per_rank, "candidates": n_img * images,
            "seconds_incl_engine_setup_and_host_statistics": dt,
            "what_is_timed": "engine construction (tables), crop, upload, propagation, device sweep + decile statistics, read-back of the 19.3 M dPSNR values, the reference's candidate permutation, all-reduce",
            "flip_evals_per_s": n_img * images / dt,
            "per_image_seconds_rank0": [float(r["seconds"]) for r in rs],
            "histogram_all_reduce_ms": 1e3 * bdist.max_over_ranks(t2 - t1),
            "attempted_total": int(np.sum(att)), "improved_total": int(np.sum(imp)),
            "scaling": "weak (images per rank fixed)"}


def run_b200(args):
    import torch
    import binary_hologram_reinforcement_learning_b200 as bh
    from binary_hologram_reinforcement_learning_b200 import dist as bdist
    from binary_hologram_reinforcement_learning_b200.engine import RULE_ENV, RESULT_DTYPE

    rank, world, local = bdist.env_info()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    # one process per GPU: run on the cores of the GPU's NUMA node before any pinned host memory is allocated
    numa = None if os.environ.get("BHOLO_NO_NUMA_BIND") else bdist.bind_to_gpu_numa_node(local)
    if os.environ.get("BHOLO_TOPO_DEBUG"):
        print(f"[rank {rank}] gpu {local} numa node {bdist.gpu_numa_node(local)} bound {numa} "
              f"cpus {sorted(os.sched_getaffinity(0))[:4]}..({len(os.sched_getaffinity(0))})", file=sys.stderr, flush=True)
    if world > 1:
        bdist.init_process_group("nccl")
    E, R, K, W = args.envs, args.rollout, args.steps, max(args.warmup, 0)
    n_pix = FRAMES * N_SIDE * N_SIDE
    plane_bytes = 4 * N_SIDE * N_SIDE
    peak, peak_src = measured_peak_gbs()

    # ---- environments (public API), reference-faithful observation ---------
    vec = make_vec(bh, E, rank, local, "eager")
    eng = vec.engine
    stream = torch.cuda.Stream(device=local)
    eng.set_stream(stream.cuda_stream)
    vec.reset()
    psnr0 = [e.initial_psnr for e in vec.envs]

    rng = np.random.default_rng(1000 + rank)
    sampler = ClockSampler(local)
    sampler.start()
    windows = []

    # ---- e2e: HologramVecEnv.step, host actions in, results + recon_image on the host every step ----
    acts_host = rng.integers(0, n_pix, size=(R * (K + W), E), dtype=np.int64)
    n_chk = min(32, R * W) if W > 0 else 0               # steps recorded for the oracle replay
    chk_rewards, chk_accept = np.zeros((n_chk, E)), np.zeros((n_chk, E), dtype=bool)
    for i in range(R * W):
        obs, rewards, _, _ = vec.step(acts_host[i])
        if i < n_chk:
            chk_rewards[i], chk_accept[i] = rewards, vec._res["accept"] != 0
    eng.stream_sync()
    bdist.barrier()
    planes0 = eng.recon_planes_written
    t0w = time.time()
    t0 = time.perf_counter()
    reward_sum, obs_sum = 0.0, 0.0
    for i in range(R * W, R * (W + K)):
        obs, rewards, _, _ = vec.step(acts_host[i])
        reward_sum += float(rewards[0])
        obs_sum += float(obs[0]["recon_image"][0, 0, 0, 0])          # the observation is host memory
    eng.stream_sync()
    e2e_s = time.perf_counter() - t0
    windows.append((t0w, time.time()))
    planes = eng.recon_planes_written - planes0
    bdist.barrier()
    e2e_s = bdist.max_over_ranks(e2e_s)
    e2e_value = world * E * R * K / e2e_s
    h2d = R * E * (8 + 4)                       # actions int64 + env ids int32 per vec step
    d2h = R * E * RESULT_DTYPE.itemsize + planes * plane_bytes / max(K, 1)     # records + recon planes per bench step

    # ---- the same loop without the observation copy, and with a device-resident observation ----
    def short_loop(mode, n_roll):
        vec.set_recon_obs(mode)
        a = rng.integers(0, n_pix, size=(R * (n_roll + 1), E), dtype=np.int64)
        for i in range(R):
            vec.step(a[i])
        eng.stream_sync()
        bdist.barrier()
        t0 = time.perf_counter()
        for i in range(R, R * (n_roll + 1)):
            vec.step(a[i])
        eng.stream_sync()
        return world * E * R * n_roll / bdist.max_over_ranks(time.perf_counter() - t0)

    lazy_value = short_loop("lazy", max(2, min(K, 5)))
    device_value = short_loop("device", max(2, min(K, 5)))
    vec.set_recon_obs("lazy")

    # ---- value: device-resident actions, CUDA events ---------------------
    d_acts = torch.from_numpy(rng.integers(0, n_pix, size=(R * (K + W + 2), E), dtype=np.int64)).cuda(local)
    d_envs = torch.arange(E, dtype=torch.int32, device=f"cuda:{local}")
    d_res = torch.zeros((R, E, RESULT_DTYPE.itemsize), dtype=torch.uint8, device=f"cuda:{local}")
    a_ptr, e_ptr, r_ptr = d_acts.data_ptr(), d_envs.data_ptr(), d_res.data_ptr()
    row_a, row_r = E * 8, E * RESULT_DTYPE.itemsize

    def rollout(base_row):
        # one persistent cooperative launch per 512-step rollout (k_rollout_t: per-environment barriers instead
        # of two launches per vectorised step; bit-identical to the step chain, tests/test_gpu_parity.py)
        eng.rollout_device(E, e_ptr, a_ptr + base_row * row_a, R, RULE_ENV, r_ptr)

    def rollout_step_chain(base_row):
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
        eng.rollout_status()                          # raises if a barrier of the persistent kernel gave up
        launches = eng.launch_count - launches0       # K: one k_rollout_t launch per bench step
        res_host = d_res.cpu().numpy().view(RESULT_DTYPE).reshape(R, E).copy()      # last timed rollout
        # a rollout of fresh actions through one k_eval + one k_commit launch per vectorised step (what step() uses)
        ec0, ec1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        rollout_step_chain((W + K) * R)
        ec0.record(stream)
        rollout_step_chain((W + K + 1) * R)
        ec1.record(stream)
        torch.cuda.synchronize()
        step_chain_ms = ec0.elapsed_time(ec1)
        chain_accept = float(d_res.cpu().numpy().view(RESULT_DTYPE).reshape(R, E)["accept"].mean())
    bdist.barrier()
    dev_ms = bdist.max_over_ranks(ev0.elapsed_time(ev1))
    value = world * E * R * K / (dev_ms / 1000.0)
    accept_rate = float(res_host["accept"].mean())

    # ---- rooflines of the delta kernels, measured live ---------------------
    n_sets = 64
    d_sets = torch.from_numpy(rng.integers(0, n_pix, size=(n_sets, E), dtype=np.int64)).cuda(local)
    d_fresh = torch.from_numpy(rng.integers(0, n_pix, size=(600, E), dtype=np.int64)).cuda(local)
    with torch.cuda.stream(stream):
        t0w = time.time()
        eval_ms = eng.time_eval(E, e_ptr, d_sets.data_ptr(), n_sets, 512)
        chain_ms = eng.time_step(E, e_ptr, d_fresh.data_ptr(), 600, 512, RULE_ENV, True)
        commit_ms = eng.time_commit(E, e_ptr, d_sets.data_ptr(), n_sets, 512)
        windows.append((t0w, time.time()))
    alg_bytes = 16.0 * N_SIDE * N_SIDE * E
    achieved = alg_bytes / (eval_ms / 1000.0) / 1e9
    commit_bytes = 24.0 * N_SIDE * N_SIDE * E
    commit_gbs = commit_bytes / (commit_ms / 1000.0) / 1e9
    # whole step: one evaluation + accept_rate kept flips per env
    step_bytes = (16.0 + 24.0 * accept_rate) * N_SIDE * N_SIDE * E
    step_gbs = step_bytes / (dev_ms / (K * R) / 1000.0) / 1e9
    chain_bytes = (16.0 + 24.0 * chain_accept) * N_SIDE * N_SIDE * E

    # ---- propagation (reset / re-sync) ------------------------------------
    with torch.cuda.stream(stream):
        prop_ms = min(eng.time_propagate(0, 10) for _ in range(3))
        prop_pass_ms = eng.time_propagate_passes(0, 3)
    prop_bytes_survey = FRAMES * (N_SIDE ** 2 + 32.0 * N_SIDE ** 2) + 8.0 * GROUPS * N_SIDE ** 2
    prop_bytes_design = FRAMES * (N_SIDE ** 2 + 40.0 * N_SIDE ** 2) + 8.0 * GROUPS * N_SIDE ** 2     # + U written

    extra = {"accept_rate": accept_rate, "psnr_env0_initial": psnr0[0]}
    dbs = group = sharded = None
    if not args.no_extras:
        # ---- secondary figures: candidate lists, exhaustive sweep -------------
        n_cand = 8192
        cand = rng.integers(0, n_pix, size=n_cand, dtype=np.int64)
        eng.eval_flips(cand[:1024], env=0)
        t0 = time.perf_counter()
        eng.eval_flips(cand, env=0)
        sweep_s = time.perf_counter() - t0
        d_map = torch.empty(n_pix, dtype=torch.float64, device=f"cuda:{local}")
        eng.sweep_all_device(0, d_map.data_ptr())
        eng.stream_sync()
        t0 = time.perf_counter()
        eng.sweep_all_device(0, d_map.data_ptr())
        eng.stream_sync()
        sweep_all_s = time.perf_counter() - t0
        del d_map
        extra.update({"flip_evals_per_s_kernel": E / (eval_ms / 1000.0) * world,
                      "flip_evals_per_s_sweep_api": n_cand / sweep_s * world,
                      "flip_evals_per_s_exhaustive_sweep_all": n_pix / sweep_all_s * world,
                      "sweep_all_ms_25M_candidates": sweep_all_s * 1e3})
        # ---- greedy DBS at three regimes (env 1), the metric's second half ----
        with torch.cuda.stream(stream):
            dbs = dbs_block(eng, min(1, E - 1), rng, peak, torch, E)
        if args.dbs_full:
            order = np.random.default_rng(31).permutation(n_pix)
            eng.load_state(E - 1, vec.envs[E - 1].state[0] * 0 + (vec.envs[E - 1].observation[0] >= 0.5))
            t0 = time.perf_counter()
            acc, _, nacc, psnr = eng.dbs_run(order, env=E - 1, k_spec=0, resync_every=1024)
            dt = time.perf_counter() - t0
            dbs["full_pass"] = {"candidates": int(n_pix), "seconds": dt, "candidates_per_s": n_pix / dt,
                                "accept_rate": nacc / n_pix, "final_psnr": psnr,
                                "accept_rate_by_tenth": [float(acc[i * n_pix // 10:(i + 1) * n_pix // 10].mean())
                                                         for i in range(10)]}
    clocks = sampler.stop(windows)

    # ---- the one collective: all-gather of per-env episode statistics ----
    vec.sync_envs()
    stats = np.array([[vec._ep_reward[i], vec.envs[i].steps, vec.envs[i].flip_count, psnr0[i],
                       vec.envs[i].previous_psnr] for i in range(E)])
    all_stats = bdist.gather_episode_stats(stats)
    extra["episode_stats_rows_gathered"] = int(all_stats.shape[0])
    vec.close()
    del vec, eng, d_acts, d_res
    torch.cuda.empty_cache()

    if not args.no_extras:
        group = group_block(bh, bdist, rank, local)
        sharded = sharded_sweep_block(bh, bdist, rank, world, local)

    cpu = parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, times, cores = cpu_reference_rate(seconds_budget=0.6 * args.cpu_seconds)
        tuned = None
        if tune_host_allocator(True):
            r2, t2, _ = cpu_reference_rate(seconds_budget=0.4 * args.cpu_seconds)
            tuned = {"value": r2, "unit": "env steps/s", "env_steps": 8 * len(t2),
                     "what": "same routine after mallopt(M_MMAP_THRESHOLD = 1 GiB), the best case of the CPU path"}
            tune_host_allocator(False)
        cpu = {"value": rate, "unit": "env steps/s", "cores": cores, "kind": "port",
               "sample": (f"{8 * len(times)} env steps of one 1024x1024x24 env, per-colour-group "
                          f"re-simulation, float32 torch CPU ({cores} threads), "
                          f"{sum(times):.1f} s (oracle/torch_path.py)"),
               "allocator": ("glibc defaults, inside the bench process (after CUDA start-up and the pinned "
                             "allocations glibc recycles large blocks more often than in a fresh process: the "
                             "reference arm, a fresh process, measures 2.3 x less with the same routine)"),
               "tuned_allocator": tuned,
               "comparators": comparator_rates()}
    if rank == 0 and n_chk > 0 and not args.no_cpu_baseline:
        # ---- self-check: the first vectorised steps of envs 0 and 1 on the float64 oracle ----
        # documented bound of the reward error (tests/test_gpu_parity.py): 1e-5 |r| + 800 * 3e-5 / (G N^2), the
        # second term being the absolute error of dPSNR that the fp32 fields leave (profiles/r2_notes.md, 6)
        floor = 800.0 * 3e-5 / (GROUPS * N_SIDE * N_SIDE)
        worst, worst_abs, worst_bound, equal, n_cmp = 0.0, 0.0, 0.0, True, 0
        for e in (0, min(1, E - 1)):
            r_ref, a_ref = oracle_replay(rank * E + e, acts_host[:n_chk, e])
            for i in range(n_chk):
                if bool(chk_accept[i, e]) != bool(a_ref[i]):
                    equal = False                        # trajectories differ from here on
                    break
                err = abs(chk_rewards[i, e] - r_ref[i])
                worst = max(worst, err / max(abs(r_ref[i]), 1e-300))
                worst_abs = max(worst_abs, err)
                worst_bound = max(worst_bound, err / (1e-5 * abs(r_ref[i]) + floor))
                n_cmp += 1
        parity = {"envs": [0, min(1, E - 1)], "steps_each": n_chk, "steps_compared": n_cmp,
                  "decisions_equal": equal, "max_rel_reward_err": worst, "max_abs_reward_err": worst_abs,
                  "max_err_over_documented_bound": worst_bound,
                  "documented_bound": "1e-5 |r| + 800 * 3e-5 / (G N^2) reward units",
                  "oracle": "float64 delta replay (oracle/hologram_oracle.py)"}

    if rank == 0:
        line = {
            "metric": "env_steps_per_s", "value": value, "unit": "env steps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": dev_ms / max(K, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(E, R),
            # dominant kernel of the timed region: ONE k_rollout_t launch per bench step (R vectorised steps)
            "roofline": {"bound": "hbm", "achieved": step_gbs, "peak": peak, "unit": "GB/s",
                         "frac": step_gbs / peak, "traffic": ncu_traffic_bytes("k_rollout_dram_bytes_per_launch"),
                         "kernel": "k_rollout_t (persistent: evaluation + commit of R x E env steps, per-environment barriers)",
                         "algorithmic_bytes_per_launch": step_bytes * R,
                         "algorithmic_bytes_per_env_step": "(16 + 24 * accept_rate) N^2  (SURVEY 8d: 16 N^2 per evaluated, 24 N^2 per kept flip)",
                         "ms_per_launch": dev_ms / K, "us_per_vec_step": 1e3 * dev_ms / (K * R),
                         "peak_source": peak_src},
            # the kernels of the step-at-a-time API (HologramVecEnv.step, the e2e path, the greedy DBS windows)
            "roofline_eval": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                              "frac": achieved / peak, "traffic": ncu_traffic_bytes(),
                              "kernel": "k_eval (delta evaluation of 8 candidates, one per env)",
                              "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": eval_ms},
            "roofline_commit": {"bound": "hbm", "achieved": commit_gbs, "peak": peak, "unit": "GB/s",
                                "frac": commit_gbs / peak, "kernel": "k_commit (8 kept flips per launch)",
                                "algorithmic_bytes_per_launch": commit_bytes, "ms_per_launch": commit_ms},
            "roofline_step_chain": {"bound": "hbm", "achieved": chain_bytes / (step_chain_ms / R / 1000.0) / 1e9, "peak": peak,
                                    "unit": "GB/s", "frac": chain_bytes / (step_chain_ms / R / 1000.0) / 1e9 / peak,
                                    "accept_rate": chain_accept,
                                    "what": "the same rollout as one k_eval + one k_commit launch per vectorised step",
                                    "us_per_vec_step": 1e3 * step_chain_ms / R,
                                    "env_steps_per_s": world * E * R / (step_chain_ms / 1000.0),
                                    "us_per_vec_step_fresh_chain_hook": 1e3 * chain_ms},
            "roofline_propagate": {"bound": "hbm", "ms_24_frames": prop_ms, "pass_ms": prop_pass_ms,
                                   "achieved": prop_bytes_survey / (prop_ms / 1000.0) / 1e9, "peak": peak, "unit": "GB/s",
                                   "frac": prop_bytes_survey / (prop_ms / 1000.0) / 1e9 / peak,
                                   "model": "SURVEY 8d: N^2 + 32 P^2 B per frame + 8 G N^2 = 0.856 GB per 24 frames",
                                   "frac_design_model_40P2": prop_bytes_design / (prop_ms / 1000.0) / 1e9 / peak},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "env steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "api": "HologramVecEnv.step (bh_vec_step + bh_recon_batch), recon_obs=eager",
                    "recon_planes_per_env_step": planes / max(1, K * R * E),
                    "pcie_d2h_gb_per_s": d2h * K / e2e_s / 1e9,
                    "checksum_reward_env0": reward_sum, "checksum_obs": obs_sum,
                    "lazy_obs": {"value": lazy_value, "unit": "env steps/s",
                                 "what": "same loop, recon_image only on refresh_recon() (round-1 headline)"},
                    "device_obs": {"value": device_value, "unit": "env steps/s",
                                   "what": "same loop, recon_image current in device memory (zero-copy view)"}},
            "parity_check": parity,
            "dbs": dbs,
            "group": group,
            "sharded_sweep": sharded,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "host_binding": {"gpu_numa_node_rank0": bdist.gpu_numa_node(local), "bound_to_node_rank0": numa,
                             "cpus_rank0": len(os.sched_getaffinity(0))},
            "extra": extra,
        }
        emit(line)
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
