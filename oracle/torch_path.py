"""Reference-shaped torch CPU path: the timed CPU baseline of bench.py.

TEST / BENCHMARK INFRASTRUCTURE ONLY (see hologram_oracle.py header; parity
unpinned).  The reference executes its arithmetic with torch (complex64
``torch.fft`` inside torchOptics); this module restates the same call sequence
with torch CPU ops so the CPU baseline is timed on the library the reference
itself would use on host cores:

    tt.Tensor(state slice) -> tt.simulate -> .abs()**2 -> torch.mean(dim=1)
    -> torch.cat with the cached group means -> tt.relativeLoss(.., get_PSNR)

(env.py:170-174; DBS_1024_24.py:324-363 for the per-colour-group form).
tests/test_oracle.py checks it against the float64 numpy oracle.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import hologram_oracle as O


class TorchRefEnv:
    """Single env, full re-simulation of the flipped colour group per step, fp32."""

    def __init__(self, N, F, wl, dx=O.PIXEL_PITCH, z=O.Z_DEFAULT, pad=1, relative=True, threads=None,
                 device="cpu"):
        """device="cuda" gives the "reference simply run on a GPU" comparator: torch.fft (cuFFT),
        the whole colour group re-uploaded every step (env.py:170) and a blocking .item()."""
        if threads:
            torch.set_num_threads(int(threads))
        self.N, self.F, self.wl, self.G = N, F, tuple(wl), len(wl)
        self.Fg, self.pad, self.relative = F // len(wl), pad, relative
        self.device = torch.device(device)
        P = N * pad
        self.H = [torch.from_numpy(O.transfer_function(P, dx, w, z).astype(np.complex64)).to(self.device)
                  for w in wl]

    def simulate(self, x: torch.Tensor, g: int) -> torch.Tensor:
        """tt.simulate on a (1, Fg, N, N) float tensor."""
        N, P = self.N, self.N * self.pad
        if self.pad == 1:
            return torch.fft.ifft2(torch.fft.fft2(x) * self.H[g])
        o = (P - N) // 2
        canvas = torch.zeros(x.shape[:-2] + (P, P), dtype=x.dtype, device=x.device)
        canvas[..., o:o + N, o:o + N] = x
        return torch.fft.ifft2(torch.fft.fft2(canvas) * self.H[g])[..., o:o + N, o:o + N]

    def relative_psnr(self, recon: torch.Tensor, target: torch.Tensor) -> float:
        if self.relative:
            s = (recon * target).sum() / (recon * recon).sum()
            recon = s * recon
        mse = torch.mean((recon - target) ** 2)
        return float(10.0 * math.log10(1.0 / mse.item()))

    def reset(self, pre_model: np.ndarray, target: np.ndarray):
        self.state = (pre_model >= 0.5).astype(np.int8)[None]           # env.py:120
        self.target = torch.from_numpy(np.ascontiguousarray(target, dtype=np.float32))[None].to(self.device)
        Fg = self.Fg
        self.means = []
        for g in range(self.G):                                          # env_1024_24.py:149-159
            x = torch.tensor(self.state[:, g * Fg:(g + 1) * Fg], dtype=torch.float32).to(self.device)
            self.means.append(torch.mean(self.simulate(x, g).abs() ** 2, dim=1, keepdim=True))
        self.previous_psnr = self.initial_psnr = self.relative_psnr(torch.cat(self.means, dim=1), self.target)
        return self.initial_psnr

    def step(self, action: int):
        """env.py:154-196 with the per-group re-simulation of DBS_1024_24.py:324-363."""
        N, Fg = self.N, self.Fg
        ch, pix = divmod(int(action), N * N)
        r, c = divmod(pix, N)
        self.state[0, ch, r, c] = 1 - self.state[0, ch, r, c]
        g = ch // Fg
        x = torch.tensor(self.state[:, g * Fg:(g + 1) * Fg], dtype=torch.float32).to(self.device)   # env.py:170
        mean_after = torch.mean(self.simulate(x, g).abs() ** 2, dim=1, keepdim=True)
        rgb = torch.cat([mean_after if k == g else self.means[k] for k in range(self.G)], dim=1)
        psnr_after = self.relative_psnr(rgb, self.target)
        change = psnr_after - self.previous_psnr
        reward = change * O.RW
        if change < 0:                                                   # env.py:191-196
            self.state[0, ch, r, c] = 1 - self.state[0, ch, r, c]
            return reward, psnr_after, False
        self.means[g] = mean_after
        self.previous_psnr = psnr_after
        return reward, psnr_after, True
