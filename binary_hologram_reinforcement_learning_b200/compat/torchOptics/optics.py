"""tt.Tensor / tt.simulate / tt.relativeLoss / tt.imread (reference call sites env.py:124-132,171-174).

``simulate`` is the hand-written sm_100a propagation (bh_simulate through the C ABI) on the
tensor's CUDA device; there is no CPU path.  ``relativeLoss`` is a few torch reductions and is
only here so unmodified reference scripts resolve -- the engine computes the same quantity
inside its kernels.
"""
import ctypes as C

import numpy as np
import torch

from binary_hologram_reinforcement_learning_b200 import engine as _eng

PAD = 1          # documented unknown (i): 1 = circular, 2 = zero padded to 2N
RELATIVE = True  # documented unknown (ii): scale-invariant loss


class Tensor(torch.Tensor):
    """A torch tensor carrying ``meta = {'dx': (dy, dx), 'wl': wavelength(s)}``."""

    @staticmethod
    def __new__(cls, data, meta=None):
        if isinstance(data, np.ndarray):
            data = torch.from_numpy(np.ascontiguousarray(data))
        data = torch.as_tensor(data)
        if not (data.is_floating_point() or data.is_complex()):
            data = data.float()
        if torch.cuda.is_available() and not data.is_cuda:
            data = data.cuda()
        obj = torch.Tensor._make_subclass(cls, data)
        obj.meta = dict(meta or {})
        return obj


def simulate(field, z, pad=None, method="asm"):
    """ifft2(fft2(field) * H_asm(z)) over the last two axes; complex64 result on the same device."""
    meta = getattr(field, "meta", {})
    wl = meta.get("wl", 515e-9)
    wl = float(wl[0] if isinstance(wl, (tuple, list)) else wl)
    dx = meta.get("dx", (7.56e-6, 7.56e-6))
    dx = float(dx[0] if isinstance(dx, (tuple, list)) else dx)
    x = torch.as_tensor(field)
    if not x.is_cuda:
        raise RuntimeError("tt.simulate needs a CUDA tensor (no CPU fallback)")
    cplx = x.is_complex()
    x = x.to(torch.complex64 if cplx else torch.float32).contiguous()
    N = x.shape[-1]
    out = torch.empty(x.shape, dtype=torch.complex64, device=x.device)
    lib = _eng.load_library()
    stream = torch.cuda.current_stream(x.device).cuda_stream
    rc = lib.bh_simulate(x.device.index or 0, C.c_void_p(stream), C.c_void_p(x.data_ptr()), int(cplx),
                         int(x.numel() // (N * N)), N, wl, dx, float(z), int(pad or PAD),
                         0 if method == "asm" else 1, C.c_void_p(out.data_ptr()), 0)
    if rc != 0:
        raise _eng.HoloError(f"bh_simulate failed ({rc}): {lib.bh_last_error(None).decode()}")
    res = Tensor(out, meta)
    return res


def relativeLoss(recon, target, fn):
    """fn(s * recon, target) with the global scale s = sum(recon*target)/sum(recon^2)."""
    recon = torch.as_tensor(recon).as_subclass(torch.Tensor)
    target = torch.as_tensor(target, device=recon.device).as_subclass(torch.Tensor)
    if RELATIVE:
        s = (recon * target).sum() / (recon * recon).sum()
        recon = s * recon
    return fn(recon, target)


def imread(path, meta=None, gray=False):
    """Image file -> Tensor (1, C, H, W) float in [0, 1] (DBS.py:191)."""
    from PIL import Image
    img = Image.open(path).convert("L" if gray else "RGB")
    arr = np.asarray(img, dtype=np.float32) / 255.0
    arr = arr[None, None] if gray else arr.transpose(2, 0, 1)[None]
    return Tensor(arr, meta)
