"""CPU oracle for the binary-hologram reward / direct-binary-search hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and there only as the checker
or as the timed CPU baseline.

PARITY UNPINNED.  The arithmetic of the reference path lives in the third-party
package ``torchOptics`` (``import torchOptics.optics as tt`` /
``torchOptics.metrics as tm`` at reference ``env.py:24-25``), which is a
git-ignored local directory of the reference repo (``.gitignore:3``), has no
pinned version in ``requirements.txt`` and is not installed here.  The reference
holds no tests, golden vectors or logs for this path.  This file is therefore a
documented *restatement* of the operator semantics implied by the reference's
call sites, exposed with the two remaining unknowns as
options: ``pad`` (1 = circular, FFT side P = N; 2 = linear, zero padded to
P = 2N) and ``relative`` (True: scale-invariant loss with s = sum(I*T)/sum(I*I)).

Every function cites the reference lines it follows.  Arithmetic is numpy;
``dtype`` selects float64/complex128 (canonical oracle) or float32/complex64
(the reference's working precision, used for the timed CPU baseline).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Iterable, List, Optional, Sequence, Tuple

import numpy as np

try:  # scipy keeps single precision and can use all host threads
    import scipy.fft as _fft
    _HAVE_SCIPY = True
except Exception:  # pragma: no cover
    import numpy.fft as _fft
    _HAVE_SCIPY = False

# ----------------------------------------------------------------------------
# constants
# ----------------------------------------------------------------------------
PIXEL_PITCH = 7.56e-6          # env.py:124, env_1024_24.py:95
WL_MONO = (515e-9,)            # env.py:124
WL_RGB = (638e-9, 515e-9, 450e-9)  # env_1024_24.py:135-138
Z_DEFAULT = 2e-3               # env.py:90,154
RW = 800                       # env.py:29,188


def _fft2(x, workers=None):
    if _HAVE_SCIPY:
        return _fft.fft2(x, axes=(-2, -1), workers=workers)
    return _fft.fft2(x, axes=(-2, -1))


def _ifft2(x, workers=None):
    if _HAVE_SCIPY:
        return _fft.ifft2(x, axes=(-2, -1), workers=workers)
    return _fft.ifft2(x, axes=(-2, -1))


# ----------------------------------------------------------------------------
# operator boundary: tt.simulate / tt.relativeLoss / tm.get_PSNR
# ----------------------------------------------------------------------------
def transfer_function(P: int, dx: float, wl: float, z: float,
                      method: str = "asm") -> np.ndarray:
    """Angular-spectrum transfer function H on the P x P FFT grid (complex128).

    Restates what ``tt.simulate(field, z)`` multiplies by (call sites
    env.py:127,172; DBS_1024_24.py:248-250,328).  ``H = exp(i 2 pi z
    sqrt(1/wl^2 - fx^2 - fy^2))`` with ``fx = fftfreq(P, dx)``, zero where the
    radicand is not positive.  At every configuration of the reference
    (dx = 7.56 um, z = 2 mm, wl in 450..638 nm) the band-limit and evanescent
    masks are no-ops, so H is pure phase.
    ``method='fresnel'`` gives the paraxial kernel as a documented alternative.
    """
    f = np.fft.fftfreq(P, d=dx)
    fy = f[:, None]
    fx = f[None, :]
    if method == "asm":
        rad = 1.0 / (wl * wl) - fx * fx - fy * fy
        ok = rad > 0.0
        phase = 2.0 * math.pi * z * np.sqrt(np.where(ok, rad, 0.0))
        H = np.where(ok, np.exp(1j * phase), 0.0 + 0.0j)
    elif method == "fresnel":
        phase = 2.0 * math.pi * z / wl - math.pi * wl * z * (fx * fx + fy * fy)
        H = np.exp(1j * phase)
    else:
        raise ValueError(f"unknown method {method!r}")
    return H.astype(np.complex128)


def impulse_response(H: np.ndarray) -> np.ndarray:
    """h = ifft2(H): field produced at the output plane by a unit pixel at (0,0)."""
    return np.fft.ifft2(H)


def simulate(x: np.ndarray, H: np.ndarray, pad: int = 1, workers=None) -> np.ndarray:
    """``tt.simulate``: U = crop_N(ifft2(fft2(pad_P(x)) * H)) over the last two axes.

    x: (..., N, N) real or complex.  H: (P, P) with P = pad * N.  With pad = 2
    the input is centred in the P x P canvas and the centre N x N window is
    returned.  The output dtype follows H's precision.
    """
    N = x.shape[-1]
    P = H.shape[-1]
    assert P == pad * N and x.shape[-2] == N
    cdt = H.dtype
    rdt = np.float32 if cdt == np.complex64 else np.float64
    if np.iscomplexobj(x):
        rdt = cdt
    if pad == 1:
        X = _fft2(x.astype(rdt, copy=False), workers=workers)
        return _ifft2(X * H, workers=workers).astype(cdt, copy=False)
    o = (P - N) // 2
    canvas = np.zeros(x.shape[:-2] + (P, P), dtype=rdt)
    canvas[..., o:o + N, o:o + N] = x
    U = _ifft2(_fft2(canvas, workers=workers) * H, workers=workers)
    return np.ascontiguousarray(U[..., o:o + N, o:o + N]).astype(cdt, copy=False)


def mse_loss(a: np.ndarray, b: np.ndarray) -> float:
    """``F.mse_loss`` (env.py:131): mean squared difference over every element."""
    d = a.astype(np.float64) - b.astype(np.float64) if a.dtype == np.float64 else a - b
    return float(np.mean(d * d))


def get_psnr(a: np.ndarray, b: np.ndarray) -> float:
    """``tm.get_PSNR`` (env.py:132,174): 10 log10(1 / mse), peak 1.0, python float."""
    m = mse_loss(a, b)
    return float(10.0 * math.log10(1.0 / m))


def relative_loss(recon: np.ndarray, target: np.ndarray, fn: Callable,
                  relative: bool = True) -> float:
    """``tt.relativeLoss(recon, target, fn)`` (env.py:131-132,174; DBS.py:270).

    Inferred semantics: one global scale ``s = sum(recon*target)/sum(recon^2)``
    over every channel, then ``fn(s*recon, target)``.  ``relative=False`` gives
    s = 1 (the second documented unknown).
    """
    if relative:
        s = np.sum(recon * target, dtype=recon.dtype) / np.sum(recon * recon, dtype=recon.dtype)
        return fn(s * recon, target)
    return fn(recon, target)


def loss_sums(I: np.ndarray, T: np.ndarray) -> Tuple[float, float, float]:
    """(sum I^2, sum I*T, sum T^2) in float64 -- the three sufficient statistics."""
    I64 = I.astype(np.float64)
    T64 = T.astype(np.float64)
    return float(np.sum(I64 * I64)), float(np.sum(I64 * T64)), float(np.sum(T64 * T64))


def mse_from_sums(sii: float, sit: float, stt: float, n: int, relative: bool = True) -> float:
    """Closed form of relative_loss(.., mse): (stt - sit^2/sii)/n  (SURVEY.md 8c)."""
    if relative:
        return (stt - sit * sit / sii) / n
    return (sii - 2.0 * sit + stt) / n


def psnr_from_mse(mse: float) -> float:
    return 10.0 * math.log10(1.0 / mse)


# ----------------------------------------------------------------------------
# configuration of one hologram problem
# ----------------------------------------------------------------------------
@dataclass
class HoloConfig:
    """Shape/physics of one environment.  F frames in G colour groups of F/G.

    Mono 256^2 x 8: env.py:27-28.  RGB 1024^2 x 24 in thirds with one
    wavelength each: env_1024_24.py:29-30,135-147.
    """
    N: int
    F: int
    wl: Tuple[float, ...] = WL_MONO
    dx: float = PIXEL_PITCH
    z: float = Z_DEFAULT
    pad: int = 1
    relative: bool = True
    method: str = "asm"
    dtype: str = "float64"   # "float64" canonical, "float32" reference precision
    workers: Optional[int] = None
    _H: list = field(default_factory=list, repr=False)
    _h: list = field(default_factory=list, repr=False)

    @property
    def G(self) -> int:
        return len(self.wl)

    @property
    def Fg(self) -> int:
        return self.F // self.G

    @property
    def P(self) -> int:
        return self.pad * self.N

    @property
    def rdtype(self):
        return np.float32 if self.dtype == "float32" else np.float64

    @property
    def cdtype(self):
        return np.complex64 if self.dtype == "float32" else np.complex128

    def H(self, g: int) -> np.ndarray:
        if not self._H:
            for w in self.wl:
                self._H.append(transfer_function(self.P, self.dx, w, self.z, self.method)
                               .astype(self.cdtype))
        return self._H[g]

    def h(self, g: int) -> np.ndarray:
        """float64 impulse response of group g (always complex128)."""
        if not self._h:
            for w in self.wl:
                self._h.append(impulse_response(
                    transfer_function(self.P, self.dx, w, self.z, self.method)))
        return self._h[g]

    def group_of(self, frame: int) -> int:
        # DBS_1024_24.py:237-238,324,334,344: thirds of the frame axis
        return frame // self.Fg

    def decode(self, action: int) -> Tuple[int, int, int]:
        # env.py:158-161
        n2 = self.N * self.N
        channel = int(action) // n2
        pix = int(action) % n2
        return channel, pix // self.N, pix % self.N


def propagate_group(cfg: HoloConfig, state_g: np.ndarray, g: int) -> np.ndarray:
    """Fields of one colour group: simulate((Fg,N,N) binary) -> complex (Fg,N,N)."""
    return simulate(state_g.astype(cfg.rdtype), cfg.H(g), cfg.pad, workers=cfg.workers)


def group_mean_intensity(U: np.ndarray) -> np.ndarray:
    """``sim.abs()**2`` then ``torch.mean(dim=1)`` (env.py:127-128,172-173)."""
    return np.mean(U.real * U.real + U.imag * U.imag, axis=0)


def reconstruct(cfg: HoloConfig, state: np.ndarray) -> np.ndarray:
    """Frame-averaged reconstruction (G,N,N) of a (F,N,N) binary stack.

    Mono: env.py:123-128.  RGB: env_1024_24.py:140-162 (three simulate calls
    on channel thirds, concatenated).
    """
    out = np.empty((cfg.G, cfg.N, cfg.N), dtype=cfg.rdtype)
    for g in range(cfg.G):
        U = propagate_group(cfg, state[g * cfg.Fg:(g + 1) * cfg.Fg], g)
        out[g] = group_mean_intensity(U)
    return out


def score(cfg: HoloConfig, recon: np.ndarray, target: np.ndarray) -> Tuple[float, float]:
    """(psnr, mse) of a reconstruction -- env.py:131-132."""
    t = target.astype(cfg.rdtype, copy=False)
    mse = relative_loss(recon, t, mse_loss, cfg.relative)
    psnr = relative_loss(recon, t, get_psnr, cfg.relative)
    return psnr, mse


# ----------------------------------------------------------------------------
# delta identity (what the CUDA delta kernel relies on)
# ----------------------------------------------------------------------------
def delta_terms(cfg: HoloConfig, U_f: np.ndarray, I_g: np.ndarray, T_g: np.ndarray,
                g: int, r: int, c: int, s: int) -> Tuple[float, float, np.ndarray]:
    """Exact change of (sum I^2, sum I*T) when pixel (r,c) of one frame flips by s=+-1.

    U'[y,x] = U[y,x] + s*h[(y-r) mod P, (x-c) mod P]; dI = (|U'|^2-|U|^2)/Fg.
    Returns (d_sii, d_sit, dI) in float64.
    """
    P, N = cfg.P, cfg.N
    h = cfg.h(g)
    yy = (np.arange(N) - r) % P
    xx = (np.arange(N) - c) % P
    hs = h[np.ix_(yy, xx)]
    U64 = U_f.astype(np.complex128)
    dI = (2.0 * s * (U64.real * hs.real + U64.imag * hs.imag)
          + hs.real ** 2 + hs.imag ** 2) / cfg.Fg
    I64 = I_g.astype(np.float64)
    d_sii = float(np.sum(dI * (2.0 * I64 + dI)))
    d_sit = float(np.sum(dI * T_g.astype(np.float64)))
    return d_sii, d_sit, dI


# ----------------------------------------------------------------------------
# environment (env.py / env_1024_24.py intent / env_1024_24_128.py / env_group.py)
# ----------------------------------------------------------------------------
def goal_bonus(success_ratio: float, const: float) -> float:
    # env.py:230-235 (const -595.2) and env.py:249-254 (const -595.24)
    return (1828.57 * (success_ratio ** 3) - 3733.33 * (success_ratio ** 2)
            + 2800 * success_ratio + const)


class OracleEnv:
    """Restatement of ``BinaryHologramEnv`` on numpy arrays.

    reset: env.py:90-152 (mono) / env_1024_24.py:95-186 (RGB: one simulate per
    colour third, one relativeLoss over the 3-channel image).
    step: env.py:154-260.  For G > 1 the reference step is broken
    (env_1024_24.py:214-237 NameError); the evident intent,
    DBS_1024_24.py:324-363, is implemented: only the flipped group is
    re-simulated, the other cached group means are reused, and the cached mean
    is replaced on accept.  The rejected-step quirks of appendix B are kept.
    """

    def __init__(self, cfg: HoloConfig, max_steps=10000, T_PSNR=30, T_steps=1,
                 T_PSNR_DIFF=0.1, reward_mode="psnr"):
        self.cfg = cfg
        self.max_steps = max_steps
        self.T_PSNR = T_PSNR
        self.T_steps = T_steps
        self.T_PSNR_DIFF = T_PSNR_DIFF
        self.reward_mode = reward_mode  # "psnr" (env.py) or "group" (env_group.py)

    # -- env.py:90-152 ---------------------------------------------------
    def reset(self, pre_model: np.ndarray, target: np.ndarray,
              rng: Optional[np.random.Generator] = None, num_samples: int = 10000):
        cfg = self.cfg
        self.pre_model = pre_model
        self.target = target.astype(cfg.rdtype)
        self.state = (pre_model >= 0.5).astype(np.int8)           # env.py:120
        self.state_record = np.zeros_like(self.state)              # env.py:121
        self.steps = 0
        self.flip_count = 0
        self.psnr_sustained_steps = 0
        self.max_psnr_diff = float("-inf")
        self.means = reconstruct(cfg, self.state)                  # env.py:127-128
        self.initial_psnr, self.initial_mse = score(cfg, self.means, self.target)
        self.previous_psnr = self.initial_psnr
        self.recon = self.means.copy()
        if self.reward_mode == "group":                            # env_group.py:192-198
            ch, ranks, pos = pixel_importance(cfg, self.state, self.target,
                                              self.initial_psnr, rng, num_samples)
            self.psnr_change_list, self.importance_ranks = ch, ranks
            self.T_PSNR_DIFF = pos / 4
        return self.state

    # -- env.py:154-260 --------------------------------------------------
    def step(self, action: int):
        cfg = self.cfg
        self.steps += 1
        channel, row, col = cfg.decode(action)
        self.state[channel, row, col] = 1 - self.state[channel, row, col]
        self.state_record[channel, row, col] += 1
        self.flip_count += 1

        g = cfg.group_of(channel)
        U = propagate_group(cfg, self.state[g * cfg.Fg:(g + 1) * cfg.Fg], g)
        mean_after = group_mean_intensity(U)
        recon_after = self.means.copy()
        recon_after[g] = mean_after
        psnr_after = relative_loss(recon_after, self.target, get_psnr, cfg.relative)
        self.recon = recon_after            # obs["recon_image"], also on reject (appendix B-2)

        psnr_change = psnr_after - self.previous_psnr
        psnr_diff = psnr_after - self.initial_psnr
        if self.reward_mode == "group":     # env_group.py:254-255
            idx = int(np.argmin(np.abs(np.array(self.psnr_change_list) - psnr_change)))
            reward = float(self.importance_ranks[idx])
        else:
            reward = psnr_change * RW       # env.py:188

        if psnr_change < 0:                 # env.py:191-196
            self.state[channel, row, col] = 1 - self.state[channel, row, col]
            self.flip_count -= 1
            return reward, False, False, psnr_after, False

        self.means[g] = mean_after          # DBS_1024_24.py:355-363 intent
        self.max_psnr_diff = max(self.max_psnr_diff, psnr_diff)
        success_ratio = self.flip_count / self.steps if self.steps > 0 else 0
        self.previous_psnr = psnr_after     # env.py:214

        if psnr_diff >= self.T_PSNR_DIFF or (psnr_after >= self.T_PSNR and psnr_diff < 0.1):
            self.psnr_sustained_steps += 1  # env.py:216-225
            if self.psnr_sustained_steps >= self.T_steps and psnr_diff >= self.T_PSNR_DIFF:
                if self.reward_mode == "group":   # env_group.py:294-299
                    reward += 100 + (-200.0 / 1500.0) * (self.steps - 1000)
                else:
                    reward += goal_bonus(success_ratio, -595.2)
        if self.steps >= self.max_steps:    # env.py:237-254
            if self.reward_mode == "group":       # env_group.py:311-315
                reward += 100 + (-200.0 / 1500.0) * (self.steps - 1000)
            else:
                reward += goal_bonus(success_ratio, -595.24)
        terminated = self.steps >= self.max_steps or self.psnr_sustained_steps >= self.T_steps
        truncated = self.steps >= self.max_steps
        return reward, terminated, truncated, psnr_after, True


# ----------------------------------------------------------------------------
# env_group.py:90-143 -- candidate scoring at reset
# ----------------------------------------------------------------------------
def importance_reward_table(psnr_changes: Sequence[float]) -> np.ndarray:
    """Rank -> reward via the degree-5 polynomial of env_group.py:121-141."""
    num_samples = len(psnr_changes)
    step_poly = np.array([10000, 9000, 8000, 5000, 2500, 1])
    rewards_poly = np.array([-0.5, -0.48, -0.45, -0.35, 0, 1])
    poly = np.poly1d(np.polyfit(step_poly, rewards_poly, len(step_poly) - 1))
    sorted_indices = np.argsort(psnr_changes)
    ranks = np.zeros(num_samples)
    for rank, idx in enumerate(sorted_indices):
        x_val = 10000 - (10000 - 1) * (rank / (num_samples - 1))
        ranks[idx] = poly(x_val)
    return ranks


def pixel_importance(cfg: HoloConfig, state: np.ndarray, target: np.ndarray,
                     initial_psnr: float, rng: np.random.Generator,
                     num_samples: int = 10000, actions: Optional[np.ndarray] = None):
    """env_group.py:90-143: score random single flips against the fixed state."""
    if actions is None:
        actions = rng.integers(0, cfg.F * cfg.N * cfg.N, size=num_samples)
    means = reconstruct(cfg, state)
    changes: List[float] = []
    positive = 0.0
    st = state.copy()
    for a in actions:
        ch, r, c = cfg.decode(int(a))
        st[ch, r, c] = 1 - st[ch, r, c]
        g = cfg.group_of(ch)
        rec = means.copy()
        rec[g] = group_mean_intensity(propagate_group(cfg, st[g * cfg.Fg:(g + 1) * cfg.Fg], g))
        d = relative_loss(rec, target.astype(cfg.rdtype), get_psnr, cfg.relative) - initial_psnr
        changes.append(d)
        if d > 0:
            positive += d
        st[ch, r, c] = 1 - st[ch, r, c]
    return changes, importance_reward_table(changes), positive


# ----------------------------------------------------------------------------
# DBS.py:242-294, DBS_1024_24.py:313-422 -- greedy direct binary search
# ----------------------------------------------------------------------------
def dbs_greedy(cfg: HoloConfig, state: np.ndarray, target: np.ndarray,
               order: Iterable[int]):
    """Visit pixels in ``order``; flip, re-simulate the flipped group, keep iff the
    PSNR strictly improves (DBS.py:273; ties rejected), else un-flip.

    Returns (final_state, accepted flags, psnr trace after each candidate).
    """
    st = state.copy()
    tgt = target.astype(cfg.rdtype)
    means = reconstruct(cfg, st)
    previous = relative_loss(means, tgt, get_psnr, cfg.relative)
    accepted, trace = [], []
    for a in order:
        ch, r, c = cfg.decode(int(a))
        st[ch, r, c] = 1 - st[ch, r, c]
        g = cfg.group_of(ch)
        mean_after = group_mean_intensity(
            propagate_group(cfg, st[g * cfg.Fg:(g + 1) * cfg.Fg], g))
        rec = means.copy()
        rec[g] = mean_after
        psnr_after = relative_loss(rec, tgt, get_psnr, cfg.relative)
        trace.append(psnr_after)
        if psnr_after > previous:
            means[g] = mean_after
            previous = psnr_after
            accepted.append(True)
        else:
            st[ch, r, c] = 1 - st[ch, r, c]
            accepted.append(False)
    return st, np.array(accepted, dtype=bool), np.array(trace, dtype=np.float64)


# ----------------------------------------------------------------------------
# dbs-1024-1024-24-6464.py:330-395, range.py:294-335 -- score-and-revert sweep
# ----------------------------------------------------------------------------
OUTPUT_BINS = np.round(np.linspace(0, 1.0, 11), decimals=10)  # ...6464.py:197


def decile_of(pre_value: float) -> int:
    """Half-open deciles, last one closed (...6464.py:377-391)."""
    nb = len(OUTPUT_BINS) - 1
    for i in range(nb):
        if i == nb - 1:
            if OUTPUT_BINS[i] <= pre_value <= OUTPUT_BINS[i + 1]:
                return i
        elif OUTPUT_BINS[i] <= pre_value < OUTPUT_BINS[i + 1]:
            return i
    return -1


def sweep(cfg: HoloConfig, state: np.ndarray, target: np.ndarray, pre_model: np.ndarray,
          order: Iterable[int]):
    """Score every candidate against the FIXED base state; never accept.

    Returns (psnr_after per candidate, initial_psnr, attempted[10], improved[10],
    improvement_sum[10]).
    """
    st = state.copy()
    tgt = target.astype(cfg.rdtype)
    means = reconstruct(cfg, st)
    previous = relative_loss(means, tgt, get_psnr, cfg.relative)
    nb = len(OUTPUT_BINS) - 1
    attempted = np.zeros(nb, dtype=np.int64)
    improved = np.zeros(nb, dtype=np.int64)
    gain = np.zeros(nb, dtype=np.float64)
    out = []
    for a in order:
        ch, r, c = cfg.decode(int(a))
        st[ch, r, c] = 1 - st[ch, r, c]
        g = cfg.group_of(ch)
        rec = means.copy()
        rec[g] = group_mean_intensity(propagate_group(cfg, st[g * cfg.Fg:(g + 1) * cfg.Fg], g))
        psnr_after = relative_loss(rec, tgt, get_psnr, cfg.relative)
        st[ch, r, c] = 1 - st[ch, r, c]                      # ...6464.py:371
        out.append(psnr_after)
        b = decile_of(float(pre_model[ch, r, c]))
        if b >= 0:
            attempted[b] += 1
            if psnr_after > previous:
                improved[b] += 1
                gain[b] += psnr_after - previous
    return np.array(out, dtype=np.float64), previous, attempted, improved, gain


# ----------------------------------------------------------------------------
# synthetic inputs -- shared by tests and bench so both sides see
# the same bytes.  Pure numpy; no reference data is read.
# ----------------------------------------------------------------------------
def _box_blur(img: np.ndarray, radius: int) -> np.ndarray:
    """Circular box blur of half-width ``radius`` along the last two axes."""
    out = img
    for ax in (-2, -1):
        acc = np.zeros_like(out)
        for s in range(-radius, radius + 1):
            acc += np.roll(out, s, axis=ax)
        out = acc / (2 * radius + 1)
    return out


def synthetic_problem(N: int, F: int, G: int, seed: int = 0):
    """(pre_model (F,N,N) f32 in [0,1), target (G,N,N) f32 in [0,1]).

    Target: blurred uniform noise rescaled to [0,1] (natural-image-like
    spectrum); pre-model output: uniform so all ten deciles are populated;
    state = pre >= 0.5 (env.py:120).
    """
    rng = np.random.default_rng(seed)
    t = rng.random((G, N, N))
    t = _box_blur(_box_blur(t, max(1, N // 64)), max(1, N // 64))
    t = (t - t.min()) / (t.max() - t.min())
    pre = rng.random((F, N, N), dtype=np.float32)
    return pre, t.astype(np.float32)
