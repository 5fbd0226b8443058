"""Target-image loader of the reference's shape (``Dataset512``, DBS.py:172-199).

The reference reads DIV2K PNGs with ``tt.imread``, centre- or random-crops N x N and
yields ``(target, path)`` through a torch DataLoader with batch size 1, so an env sees
``target`` of shape (1, C, N, N) in [0, 1] and the 1-tuple ``(path,)`` that torch's default collate
makes of a batch of one string (the reference's logs print it as ``('.../0001.png',)`` and its
log parsers split on that form, log_py/valid_log.py:12, log_py/comp.py:25).  ``ImageFolderLoader`` yields
exactly that from a directory of images (PIL is the only dependency); any iterable of
the same shape -- including the reference's own DataLoader -- works as ``trainloader``.
"""
from __future__ import annotations

import glob
import os
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np


def load_image(path: str, gray: bool = False) -> np.ndarray:
    """Image file -> float32 (C, H, W) in [0, 1] (the ``tt.imread`` contract)."""
    from PIL import Image
    img = Image.open(path).convert("L" if gray else "RGB")
    arr = np.asarray(img, dtype=np.float32) / 255.0
    return arr[None] if gray else np.ascontiguousarray(arr.transpose(2, 0, 1))


def crop_to(img: np.ndarray, N: int, rng: Optional[np.random.Generator] = None) -> np.ndarray:
    """Centre crop (rng None) or random crop of N x N; smaller images are tiled up first."""
    C, H, W = img.shape
    if H < N or W < N:
        reps = (1, -(-N // H), -(-N // W))
        img = np.tile(img, reps)
        C, H, W = img.shape
    if rng is None:
        y0, x0 = (H - N) // 2, (W - N) // 2
    else:
        y0, x0 = int(rng.integers(0, H - N + 1)), int(rng.integers(0, W - N + 1))
    return np.ascontiguousarray(img[:, y0:y0 + N, x0:x0 + N])


class ImageFolderLoader:
    """Iterable of ``(target (1, C, N, N) float32, (path,))`` over the images of a directory."""

    def __init__(self, target_dir: str, N: int, gray: bool = False, random_crop: bool = False,
                 shuffle: bool = False, seed: Optional[int] = None,
                 patterns: Sequence[str] = ("*.png", "*.jpg", "*.jpeg", "*.bmp")):
        self.N, self.gray, self.random_crop, self.shuffle = N, gray, random_crop, shuffle
        self.rng = np.random.default_rng(seed)
        self.target_list: List[str] = sorted(
            p for pat in patterns for p in glob.glob(os.path.join(target_dir, pat)))
        if not self.target_list:
            raise FileNotFoundError(f"no images under {target_dir}")

    def __len__(self) -> int:
        return len(self.target_list)

    def __iter__(self) -> Iterator[Tuple[np.ndarray, List[str]]]:
        order = np.arange(len(self.target_list))
        if self.shuffle:
            self.rng.shuffle(order)
        for i in order:
            path = self.target_list[i]
            img = crop_to(load_image(path, self.gray), self.N, self.rng if self.random_crop else None)
            yield img[None], (path,)
