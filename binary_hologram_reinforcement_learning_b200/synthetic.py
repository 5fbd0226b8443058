"""Synthetic targets / pre-model outputs of the reference's shapes.

The reference needs DIV2K images and git-ignored U-Net weights
(DBS.py:311-312,329); neither is available offline, so benchmarks and tests use
a seeded stand-in: a blurred-noise target in [0,1] and a uniform pre-model
output (so all ten deciles of dbs-1024-1024-24-6464.py:197 are populated).
"""
from __future__ import annotations

import numpy as np


def _box_blur(img: np.ndarray, radius: int) -> np.ndarray:
    out = img
    for ax in (-2, -1):
        acc = np.zeros_like(out)
        for s in range(-radius, radius + 1):
            acc += np.roll(out, s, axis=ax)
        out = acc / (2 * radius + 1)
    return out


def synthetic_problem(N: int, F: int, G: int, seed: int = 0):
    """(pre_model (F,N,N) f32 in [0,1), target (G,N,N) f32 in [0,1])."""
    rng = np.random.default_rng(seed)
    t = rng.random((G, N, N))
    t = _box_blur(_box_blur(t, max(1, N // 64)), max(1, N // 64))
    t = (t - t.min()) / (t.max() - t.min())
    pre = rng.random((F, N, N), dtype=np.float32)
    return pre, t.astype(np.float32)


class SyntheticLoader:
    """Iterable of (target (1,G,N,N) float32, (name,)) like the reference's DataLoader."""

    def __init__(self, N: int, F: int, G: int, seeds=(0,)):
        self.N, self.F, self.G, self.seeds = N, F, G, tuple(seeds)
        self._pre = {}

    def __iter__(self):
        for s in self.seeds:
            pre, tgt = synthetic_problem(self.N, self.F, self.G, s)
            self._pre[tgt[0, 0, :4].tobytes()] = pre
            yield tgt[None], (f"synthetic_{s:04d}.png",)

    def target_function(self, target):
        """Stand-in for BinaryNet: returns the seeded pre-model output of this target."""
        t = target.detach().cpu().numpy() if hasattr(target, "detach") else np.asarray(target)
        return self._pre[np.ascontiguousarray(t[0, 0, 0, :4], dtype=np.float32).tobytes()][None]
