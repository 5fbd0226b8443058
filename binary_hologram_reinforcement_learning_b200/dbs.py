"""Direct binary search drivers on the CUDA engine.

``optimize_with_random_pixel_flips`` keeps the reference's two call shapes:

* ``(env, z=2e-3[, pixel_pitch])`` -- greedy DBS (DBS.py:202-307,
  DBS_1024_24.py:206-469): visit every pixel once in a random order, keep a flip
  iff the PSNR strictly improves.
* ``(target_function, trainloader, z, pixel_pitch, crop_margin)`` -- the
  score-and-revert sweep (dbs-1024-1024-24-6464.py:194-478, range.py:195-421):
  every candidate is scored against the fixed base state and binned by the
  decile of the pre-model output.

The reference re-simulates a colour group per candidate on the host loop; here
the loop runs on the device (speculative batches through ``bh_dbs_run``) or as
batched ``bh_eval_flips`` launches.  Both return their results (the reference
only prints); the printed blocks keep the reference's format.
"""
from __future__ import annotations

import os
import time
from typing import Callable, Iterable, List, Optional

import numpy as np

from .engine import HoloEngine
from .envs import BinaryHologramEnv, WL_MONO, WL_RGB, _to_numpy

OUTPUT_BINS = np.round(np.linspace(0, 1.0, 11), decimals=10)   # dbs-...-6464.py:197


def decile_index(pre_values: np.ndarray) -> np.ndarray:
    """Half-open deciles, last one closed (dbs-...-6464.py:377-391); -1 = outside [0,1]."""
    v = np.asarray(pre_values, dtype=np.float64)
    idx = np.searchsorted(OUTPUT_BINS, v, side="right") - 1
    idx[v == OUTPUT_BINS[-1]] = len(OUTPUT_BINS) - 2
    idx[(v < OUTPUT_BINS[0]) | (v > OUTPUT_BINS[-1])] = -1
    return idx


def _decode(action, N: int):
    """(channel, row, col) of a flat action on an N x N grid (DBS.py:248-251)."""
    a = int(action)
    return a // (N * N), (a % (N * N)) // N, a % N


def _accepted_bins(pre, perm, accepted, trace, initial_psnr):
    """Accepted flips and their PSNR gains per decile of the pre-model output (DBS_1024_24.py:398-416)."""
    nb = len(OUTPUT_BINS) - 1
    acc_idx = np.flatnonzero(accepted)
    if not acc_idx.size:
        return np.zeros(nb, dtype=np.int64), np.zeros(nb)
    prev_psnr = np.concatenate([[initial_psnr], trace[acc_idx[:-1]]])
    d = decile_index(pre.ravel()[perm[acc_idx]])
    ok = d >= 0
    return (np.bincount(d[ok], minlength=nb),
            np.bincount(d[ok], weights=(trace[acc_idx] - prev_psnr)[ok], minlength=nb))


def _emit_sweep_log(file_name, initial_psnr, perm, psnr_all, cpre, N, bin_counts, t0, log_every=5000):
    """The stdout of the reference's score-and-revert sweep, replayed from the per-candidate results.

    dbs-1024-1024-24-6464.py:396-431: every ``log_every`` candidates a step block (the candidate of that
    step, counters including it, its pre-model value) and the cumulative range lines WITH ``Attempted
    Pixels`` -- the lines log_py/'dbs 평균.py':25-31 parses; then the summary of :439-447,462-478 (range
    lines without the attempt counts).  The sweep never accepts, so ``PSNR Before`` is the initial PSNR.
    """
    n = int(perm.shape[0])
    nb = len(OUTPUT_BINS) - 1
    pre_vals = np.asarray(cpre).ravel()[perm]
    d = decile_index(pre_vals)
    better = psnr_all > initial_psnr
    gain = np.where(better, psnr_all - initial_psnr, 0.0)
    att, imp, gn = np.zeros(nb, dtype=np.int64), np.zeros(nb, dtype=np.int64), np.zeros(nb)
    flips = 0

    def step_block(k, with_pre):
        ch, row, col = _decode(perm[k - 1], N)
        tail = f"\npre_value: {float(pre_vals[k - 1]):.6f}" if with_pre else ""
        print(f"Step: {k}"
              f"\nPSNR Before: {initial_psnr:.6f} | PSNR After: {psnr_all[k - 1]:.6f} | Change: {psnr_all[k - 1] - initial_psnr:.6f}"
              f"\nSuccess Ratio: {flips / k:.6f} | Flip Count: {flips}"
              f"\nFlip Pixel: Channel={ch}, Row={row}, Col={col}"
              f"\nTime taken for this data: {time.time() - t0:.2f} seconds{tail}")

    step = max(1, int(log_every))
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        dd, bb = d[lo:hi], better[lo:hi]
        ok = dd >= 0
        att += np.bincount(dd[ok], minlength=nb)
        imp += np.bincount(dd[ok & bb], minlength=nb)
        gn += np.bincount(dd[ok & bb], weights=gain[lo:hi][ok & bb], minlength=nb)
        flips += int(np.count_nonzero(bb))
        if hi - lo == step:                                # ...6464.py:396 `steps % 5000 == 0`
            step_block(hi, True)
            _print_bins(bin_counts, imp, gn, att)
    if n:
        step_block(n, False)                               # ...6464.py:439-445
    last = psnr_all[-1] if n else initial_psnr
    print(f"{file_name}.png Optimization completed. Final PSNR improvement: {last - initial_psnr:.6f}")
    print(f"Time taken for this data: {time.time() - t0:.2f} seconds\n")
    print("Pre-model output range statistics:")
    _print_bins(bin_counts, imp, gn)
    print("\n")


def _file_stem(name) -> str:
    if isinstance(name, (list, tuple)):
        name = name[0]
    return os.path.splitext(os.path.basename(str(name)))[0]


def _permutation(n: int, rng) -> np.ndarray:
    """np.arange + np.random.shuffle (DBS.py:243-244); a Generator can be injected."""
    order = np.arange(n, dtype=np.int64)
    (rng if rng is not None else np.random).shuffle(order)
    return order


def _print_bins(bin_counts, improved, gains, attempted=None):
    total_improved = int(np.sum(improved))
    for i in range(len(OUTPUT_BINS) - 1):
        tc, ic = int(bin_counts[i]), int(improved[i])
        ratio_in = ic / tc if tc > 0 else 0
        ratio_tot = ic / total_improved if total_improved > 0 else 0
        tot = float(gains[i]) if ic > 0 else 0
        avg = tot / ic if ic > 0 else 0
        mid = ""
        if attempted is not None:
            ac = int(attempted[i])
            mid = f"Attempted Pixels = {ac}, Improvement Ratio = {(ic / ac if ac > 0 else 0):.6f}, "
        print(f"Range {OUTPUT_BINS[i]:.1f}-{OUTPUT_BINS[i + 1]:.1f}: "
              f"Total Pixels = {tc}, Improved Pixels = {ic}, {mid}"
              f"Improvement Ratio (in range) = {ratio_in:.6f}, "
              f"Improvement Ratio (to total improved) = {ratio_tot:.6f}, "
              f"Total PSNR Improvement = {tot:.6f}, "
              f"Average PSNR Improvement = {avg:.8f}")


def bin_population(pre_model: np.ndarray) -> np.ndarray:
    """Pixels per decile of the pre-model output (DBS_1024_24.py:290-301)."""
    idx = decile_index(pre_model.ravel())
    return np.bincount(idx[idx >= 0], minlength=len(OUTPUT_BINS) - 1)


# ---------------------------------------------------------------------------
# greedy DBS
# ---------------------------------------------------------------------------
def dbs_greedy_env(env: BinaryHologramEnv, z=2e-3, pixel_pitch=7.56e-6, crop_margin: Optional[int] = None, *,
                   psnr_diff_threshold: Optional[float] = None, range_stats: Optional[bool] = None,
                   max_datasets=None,
                   order: Optional[np.ndarray] = None, rng=None, k_spec: int = 0,
                   resync_every: int = 1024, segment: int = 1 << 20, verbose: bool = True,
                   max_candidates: Optional[int] = None, save_dir: Optional[str] = None,
                   checkpoint: Optional[str] = None, max_segments: Optional[int] = None) -> List[dict]:
    """DBS.py:202-307 / DBS_1024_24.py:206-469 on the device-resident engine.

    ``checkpoint``: path of an .npz written after every segment (image number, binary state,
    candidate order, cursor, decisions so far).  A later call with the same path skips the
    images that were already finished, and continues the image in progress from its cursor
    WITH THE STORED ORDER (a freshly drawn permutation would never match it) -- the reference keeps
    nothing but the before/after reconstructions (DBS_1024_24.py:282-287,446-451).  An explicit
    ``order`` must equal the stored one, otherwise the image starts over.
    ``max_segments`` bounds the work of one call (the result then has ``complete = False``
    and a later call with the same checkpoint continues).
    ``crop_margin`` (DBS_1024_24-128.py:187): optimise the centre window only; passed to
    ``env.reset`` (env_1024_24_128.py:100).  ``psnr_diff_threshold`` (DBS_01.py:204,320-325,
    DBS_ratio_0.5.py:204): leave an image as soon as a candidate lifts the PSNR by that much.
    ``range_stats``: print the per-decile tables of the kept flips (DBS_ratio.py, DBS_1024_24.py:379-396,
    453-469); default: for colour holograms only, as in the reference's scripts.
    """
    results = []
    db_num = 0
    segments_run, out_of_budget = 0, False               # max_segments bounds the whole call
    stats = (env.G > 1) if range_stats is None else bool(range_stats)
    if max_datasets is None:
        max_datasets = 800 if env.G == 1 else 10          # DBS.py:205, DBS_1024_24.py:208
    if checkpoint and os.path.exists(checkpoint):
        # a restarted process begins at the loader's first image again: skip what the checkpoint has finished
        with np.load(checkpoint) as ck0:
            done_images = (int(ck0["db_num"]) - (0 if int(ck0["image_done"]) else 1)) if "db_num" in ck0.files else 0
        for _ in range(max(0, done_images)):
            env._next_target()
            env.episode_num_count += 1
            db_num += 1
    while db_num <= max_datasets:                         # DBS.py:208 (runs max+1 images)
        try:
            obs, info = env.reset(z=z, pixel_pitch=pixel_pitch, crop_margin=crop_margin)
            db_num += 1
        except Exception as e:                            # DBS.py:212-214
            print(f"An error occurred during reset: {e}")
            break
        t0 = time.time()
        eng, e = env.engine, env._e
        initial_psnr = env.initial_psnr
        file_name = _file_stem(env.current_file)
        pre = env._crop(env.observation[0])
        bin_counts = bin_population(pre)
        if save_dir:                                      # DBS_1024_24.py:282-287
            os.makedirs(save_dir, exist_ok=True)
            before_path = os.path.join(save_dir, f"episode_{file_name}png_rgb_before.npy")
            np.save(before_path, eng.recon(e)[None])
            if verbose:
                print(f"RGB data saved to {before_path}")
        if verbose:
            print(f"Starting pixel flip optimization for file {file_name}.png with initial PSNR: {initial_psnr:.6f}")
        n = eng.num_pixels
        perm = _permutation(n, rng) if order is None else np.asarray(order, dtype=np.int64)
        if max_candidates is not None:
            perm = perm[:max_candidates]
        accepted = np.zeros(perm.shape[0], dtype=np.uint8)
        trace = np.zeros(perm.shape[0], dtype=np.float64)
        step_gain = 0.5 if env.G == 1 else 0.1            # DBS.py:226, DBS_1024_24.py:304
        thresholds = [initial_psnr + i * step_gain for i in range(1, 21 if env.G == 1 else 101)]
        previous = initial_psnr
        flip_count = 0
        last_change = last_ratio = None                   # set by the threshold blocks (stale in the summary)
        start = 0
        if checkpoint and os.path.exists(checkpoint):
            ck = np.load(checkpoint)
            same_image = ck["fname"].item() == file_name and not ("image_done" in ck.files and int(ck["image_done"]))
            if same_image and order is None and (max_candidates is None or ck["order"].shape[0] == max_candidates):
                perm = ck["order"]                         # adopt the stored order of the image in progress
                accepted = np.zeros(perm.shape[0], dtype=np.uint8)
                trace = np.zeros(perm.shape[0], dtype=np.float64)
            if same_image and ck["order"].shape == perm.shape and np.array_equal(ck["order"], perm):
                start = int(ck["cursor"])
                accepted[:start], trace[:start] = ck["accepted"][:start], ck["trace"][:start]
                eng.load_state(e, ck["state"])             # re-propagates from the saved hologram
                flip_count = int(np.count_nonzero(accepted[:start]))
                hits = np.flatnonzero(accepted[:start])
                previous = trace[hits[-1]] if hits.size else initial_psnr
                if verbose:
                    print(f"Resuming {file_name}.png from candidate {start} (PSNR {eng.metrics(e)[0]:.6f})")
        done_upto = start
        stopped = False
        seg_len = segment if psnr_diff_threshold is None else min(segment, 4096)
        for lo in range(start, perm.shape[0], seg_len):
            if max_segments is not None and segments_run >= max_segments:
                out_of_budget = True
                break
            segments_run += 1
            hi = min(perm.shape[0], lo + seg_len)
            state_before = eng.state(e) if psnr_diff_threshold is not None else None
            acc, tr, nacc, psnr_now = eng.dbs_run(perm[lo:hi], env=e, k_spec=k_spec,
                                                  resync_every=resync_every, trace=True)
            if psnr_diff_threshold is not None:           # DBS_01.py:320-325
                reach = np.flatnonzero(tr - initial_psnr >= psnr_diff_threshold)
                if reach.size:
                    k = int(reach[0]) + 1
                    if k < hi - lo:                       # later flips were applied: replay the prefix only
                        eng.load_state(e, state_before)
                        acc, tr, nacc, psnr_now = eng.dbs_run(perm[lo:lo + k], env=e, k_spec=k_spec,
                                                              resync_every=resync_every, trace=True)
                    hi, stopped = lo + k, True
            done_upto = hi
            accepted[lo:hi], trace[lo:hi] = acc, tr
            flip_count += nacc
            if verbose and nacc:
                hit = np.flatnonzero(acc) + lo
                for j in hit:
                    while thresholds and trace[j] >= thresholds[0]:
                        thresholds.pop(0)
                        pa = trace[j]
                        prev_acc = hit[hit < j]
                        pb = trace[prev_acc[-1]] if prev_acc.size else previous
                        n_acc = int(np.count_nonzero(accepted[:j + 1]))
                        last_change, last_ratio = pa - pb, n_acc / (j + 1)
                        ch, row, col = _decode(perm[j], eng.N)
                        print(f"Step: {j + 1}"                       # DBS.py:283-289, DBS_1024_24.py:372-378
                              f"\nPSNR Before: {pb:.6f} | PSNR After: {pa:.6f} | Change: {last_change:.6f} | Diff: {pa - initial_psnr:.6f}"
                              f"\nSuccess Ratio: {last_ratio:.6f} | Flip Count: {n_acc}"
                              f"\nFlip Pixel: Channel={ch}, Row={row}, Col={col}"
                              f"\nTime taken for this data: {time.time() - t0:.2f} seconds")
                        if stats:                                    # DBS_1024_24.py:379-396 (flips before this one)
                            imp_j, gn_j = _accepted_bins(pre, perm, accepted[:j], trace, initial_psnr)
                            _print_bins(bin_counts + imp_j, imp_j, gn_j)
                            print("\n")
                previous = trace[hit[-1]]
            if stopped:
                if verbose:
                    print(f"PSNR diff threshold {psnr_diff_threshold} reached at step {hi}. Moving to next dataset.")
                break
            if checkpoint:
                tmp = checkpoint + ".tmp.npz"
                np.savez_compressed(tmp, fname=np.array(file_name), order=perm, cursor=np.array(hi),
                                    accepted=accepted, trace=trace, state=eng.state(e), db_num=np.array(db_num),
                                    image_done=np.array(int(hi == perm.shape[0])))
                os.replace(tmp, checkpoint)
        # host mirrors follow the device state
        new_state = eng.state(e)
        env._crop(env.state[0])[...] = new_state
        final_psnr, _, _ = eng.metrics(e)
        env.previous_psnr = final_psnr
        # decile statistics of the accepted flips (DBS_1024_24.py:390-416)
        improved, gains = _accepted_bins(pre, perm, accepted, trace, initial_psnr)
        steps = int(done_upto)
        accepted, trace = accepted[:steps], trace[:steps]
        dt = time.time() - t0
        out = dict(complete=steps == perm.shape[0], stopped_on_threshold=stopped, file=file_name, initial_psnr=initial_psnr, final_psnr=final_psnr, steps=steps,
                   flip_count=int(flip_count), accepted=accepted, psnr_trace=trace, order=perm,
                   seconds=dt, bin_counts=bin_counts + improved, improved_bin_counts=improved,
                   psnr_improvements=gains, state=new_state)
        results.append(out)
        if verbose:
            # the reference's summary shows the LAST EVALUATED candidate (possibly rejected) next to the stale
            # change / ratio of the last threshold block (DBS.py:297-305, appendix B-11)
            last = trace[-1] if steps else initial_psnr
            ch, row, col = _decode(perm[steps - 1], eng.N) if steps else (0, 0, 0)
            if last_change is None:                       # no threshold block was printed: the reference raises here
                last_change, last_ratio = 0.0, (flip_count / steps if steps else 0.0)
            change = (f"Change: {last_change:.6f} | Diff: {last - initial_psnr:.6f}" if env.G == 1   # DBS.py:301
                      else f"Change: {last - initial_psnr:.6f}")                                       # DBS_1024_24.py:432
            print(f"Step: {steps}"
                  f"\nPSNR Before: {final_psnr:.6f} | PSNR After: {last:.6f} | {change}"
                  f"\nSuccess Ratio: {last_ratio:.6f} | Flip Count: {flip_count}"
                  f"\nFlip Pixel: Channel={ch}, Row={row}, Col={col}"
                  f"\nTime taken for this data: {dt:.2f} seconds")
            print(f"{file_name}.png Optimization completed. Final PSNR improvement: {last - initial_psnr:.6f}")
            print(f"Time taken for this data: {dt:.2f} seconds\n")
            if stats:
                print("Pre-model output range statistics:")
        if save_dir:                                      # DBS_1024_24.py:441-451 (+ the hologram itself)
            after_path = os.path.join(save_dir, f"episode_{file_name}_rgb_after.npy")
            np.save(after_path, eng.recon(e)[None])
            np.save(os.path.join(save_dir, f"episode_{file_name}_state_after.npy"), new_state)
            if verbose:
                print(f"RGB data saved to {after_path}")
        if verbose and stats:                             # DBS_1024_24.py:453-469
            _print_bins(out["bin_counts"], improved, gains)
            print("\n")
        if out_of_budget:
            break
    return results


# ---------------------------------------------------------------------------
# score-and-revert sweep
# ---------------------------------------------------------------------------
def sweep_engine(eng: HoloEngine, env_index: int, pre_model: np.ndarray, order: np.ndarray,
                 initial_psnr: float, psnr_map: Optional[np.ndarray] = None, exhaustive="auto"):
    """Score ``order`` against the fixed state of one env; decile statistics.

    Large sweeps (at least a quarter of all pixels) are scored by ``bh_sweep_all`` --
    every pixel at once through FFT correlations -- and gathered in ``order``; small ones
    by the batched delta kernel (``bh_eval_flips``).  ``psnr_map`` passes a map already
    computed for this state.  Returns dict(psnr_after, attempted, improved, gains, flip_count).
    """
    order = np.asarray(order, dtype=np.int64)
    use_map = psnr_map is not None or exhaustive is True or (
        exhaustive == "auto" and eng.Fg % 2 == 0
        and order.shape[0] * 4 >= eng.num_pixels)
    if use_map:
        if psnr_map is None:
            psnr_map = eng.sweep_all(env_index)
        psnr_after = psnr_map.reshape(-1)[order]
    else:
        psnr_after = eng.eval_flips(order, env=env_index)
    better = psnr_after > initial_psnr                    # dbs-...-6464.py:381,393
    d = decile_index(pre_model.ravel()[order])
    ok = d >= 0
    nb = len(OUTPUT_BINS) - 1
    attempted = np.bincount(d[ok], minlength=nb)
    improved = np.bincount(d[ok & better], minlength=nb)
    gains = np.bincount(d[ok & better], weights=(psnr_after - initial_psnr)[ok & better], minlength=nb)
    return dict(psnr_after=psnr_after, attempted=attempted, improved=improved, gains=gains,
                flip_count=int(np.count_nonzero(better)))


def dbs_sweep(target_function: Callable, trainloader: Iterable, z=2e-3, pixel_pitch=7.56e-6,
              crop_margin=64, *, CH=24, wl=WL_RGB, max_datasets=1, order: Optional[np.ndarray] = None,
              rng=None, device=0, pad=1, relative=True, verbose=True,
              max_candidates: Optional[int] = None, chunk: int = 1 << 18,
              shard: Optional[tuple] = None, log_every: int = 5000) -> List[dict]:
    """dbs-1024-1024-24-6464.py:194-478: crop, then score every flip and always revert.

    ``shard=(rank, world)`` scores only this rank's contiguous slice of the
    (globally defined) candidate order -- the multi-GPU partition of SURVEY 8e.
    """
    results = []
    data_iter = iter(trainloader)
    db_num = 0
    eng = None
    if verbose:
        print(OUTPUT_BINS)
    while db_num <= max_datasets:                         # ...6464.py:202
        try:
            target_image, current_file = next(data_iter)
            db_num += 1
        except Exception as e:
            print(f"An error occurred during reset: {e}")
            break
        t0 = time.time()
        tgt_in = target_image
        if hasattr(tgt_in, "cuda"):
            try:
                tgt_in = tgt_in.cuda(device)
            except Exception:
                pass
        target_np = np.ascontiguousarray(_to_numpy(tgt_in), dtype=np.float32)
        observation = np.ascontiguousarray(_to_numpy(target_function(tgt_in)), dtype=np.float32)
        state = (observation >= 0.5).astype(np.int8)     # ...6464.py:217
        m = int(crop_margin)
        crop = (lambda a: a[..., m:-m, m:-m]) if m > 0 else (lambda a: a)
        cstate, ctarget, cpre = crop(state[0]), crop(target_np[0]), crop(observation[0])
        N = cstate.shape[-1]
        if eng is None or eng.N != N or eng.z != float(z) or eng.dx != float(pixel_pitch):
            if eng is not None:
                eng.close()
            eng = HoloEngine(N, cstate.shape[0], wl, n_env=1, device=device, dx=pixel_pitch, z=z,
                             pad=pad, relative=relative)
        eng.set_target(0, ctarget)
        eng.load_state(0, cstate)
        initial_psnr, _, _ = eng.metrics(0)
        file_name = _file_stem(current_file)
        if verbose:
            print(f"Starting pixel flip optimization for file {file_name}.png with initial PSNR: {initial_psnr:.6f}")
        n = eng.num_pixels
        perm = _permutation(n, rng) if order is None else np.asarray(order, dtype=np.int64)
        if max_candidates is not None:
            perm = perm[:max_candidates]
        lo_all, hi_all = 0, perm.shape[0]
        if shard is not None:
            rank, world = shard
            per = (perm.shape[0] + world - 1) // world
            lo_all, hi_all = min(perm.shape[0], rank * per), min(perm.shape[0], (rank + 1) * per)
        nb = len(OUTPUT_BINS) - 1
        attempted = np.zeros(nb, dtype=np.int64)
        improved = np.zeros(nb, dtype=np.int64)
        gains = np.zeros(nb)
        psnr_all = np.empty(hi_all - lo_all, dtype=np.float64)
        flip_count = 0
        psnr_map = None
        if (eng.Fg % 2 == 0 and shard is None and max_candidates is None
                and perm.shape[0] == n):
            # the whole image in one call: correlation sweep + decile statistics on the device
            att, imp, gn, psnr_map = eng.sweep_stats(cpre, OUTPUT_BINS, 0, want_map=True)
            bin_counts = att.copy()                        # every pixel is a candidate: population = attempts per bin
            psnr_all = psnr_map.reshape(-1)[perm]
            flip_count = int(imp.sum())
            dt = time.time() - t0
            results.append(dict(file=file_name, initial_psnr=initial_psnr, psnr_after=psnr_all, order=perm,
                                attempted=att, improved=imp, gains=gn, bin_counts=bin_counts,
                                flip_count=flip_count, steps=int(n), seconds=dt))
            if verbose:
                _emit_sweep_log(file_name, initial_psnr, perm, psnr_all, cpre, N, bin_counts, t0, log_every)
            continue
        bin_counts = bin_population(cpre)
        if eng.Fg % 2 == 0 and (hi_all - lo_all) * 4 >= n:
            psnr_map = eng.sweep_all(0)                    # every candidate in one call
        for lo in range(lo_all, hi_all, chunk):
            hi = min(hi_all, lo + chunk)
            r = sweep_engine(eng, 0, cpre, perm[lo:hi], initial_psnr, psnr_map=psnr_map,
                             exhaustive=False)
            psnr_all[lo - lo_all:hi - lo_all] = r["psnr_after"]
            attempted += r["attempted"]; improved += r["improved"]; gains += r["gains"]
            flip_count += r["flip_count"]
        dt = time.time() - t0
        results.append(dict(file=file_name, initial_psnr=initial_psnr, psnr_after=psnr_all,
                            order=perm[lo_all:hi_all], attempted=attempted, improved=improved,
                            gains=gains, bin_counts=bin_counts, flip_count=flip_count,
                            steps=int(hi_all - lo_all), seconds=dt))
        if verbose:
            _emit_sweep_log(file_name, initial_psnr, perm[lo_all:hi_all], psnr_all, cpre, N, bin_counts, t0, log_every)
    if eng is not None:
        eng.close()
    return results


def optimize_with_random_pixel_flips(first, *args, **kw):
    """Reference entry point (DBS.py:202; dbs-1024-1024-24-6464.py:194).

    ``first`` is an env -> greedy DBS; ``first`` is a callable ``target_function``
    (followed by the trainloader) -> score-and-revert sweep.
    """
    if isinstance(first, BinaryHologramEnv):
        return dbs_greedy_env(first, *args, **kw)
    if callable(first):
        return dbs_sweep(first, *args, **kw)
    raise TypeError("expected a BinaryHologramEnv or a target_function")
